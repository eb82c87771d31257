/* ulc_oracle.c - CPU restatement of the ulc-codec block encode/decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ulc_codec_b200/ (the product) imports, links or
 * executes this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg have something to check the CUDA path against on a box where
 * /root/reference does not exist.
 *
 * Pinning: the reference ships no golden vectors or tests, so this
 * restatement is pinned against the reference ITSELF, compiled unmodified into oracle/_ref by
 * oracle/Makefile: tests/test_oracle_vs_ref.py demands BIT-IDENTICAL output (DCT-II/IV, MDCT,
 * IMDCT, decoded PCM, encoder decisions and bytes) against the stock scalar build, and
 * tests/golden/ holds reference-generated fixtures for boxes without oracle/_ref.
 * To be bit-identical the float arithmetic below follows the reference's *operation order*
 * (plain IEEE f32 mul/add, no FMA contraction: build with -ffp-contract=off); structure,
 * naming and control flow are this repository's own.
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* Trig: minimax sin/cos of x*pi/2 (libfourier/source/Fourier_Helper.h:218-240, scalar path:  */
/* the VFMA macro expands to an unfused x*y+a).                                               */
/* ------------------------------------------------------------------------------------------ */
static float sin_halfpi(float x) {
	float x2 = x * x;
	float r = x2 * 0x1.3C8B08p-13f + -0x1.3237ECp-8f;
	r = x2 * r + 0x1.4667ACp-4f;
	r = x2 * r + -0x1.4ABBB8p-1f;
	r = x2 * r + 0x1.921FB4p0f;
	return x * r;
}
static float cos_halfpi(float x) {
	float x2 = x * x;
	float r = x2 * 0x1.C2EA62p-11f + -0x1.550810p-6f;
	r = x2 * r + 0x1.03BDD6p-2f;
	r = x2 * r + -0x1.3BD3B2p0f;
	r = x2 * r + 1.0f;
	return r;
}

/* A quarter-wave rotator: yields (cos,sin)((k+1/2)/span * pi/2) for k = 0,1,2,... by the same
 * f32 rotation recurrence every libfourier scalar loop uses (e.g. Fourier_DCT4_Template.h:49-68,
 * Fourier_IMDCT_Template.h:76-91, Fourier_MDCT_Template.h:131-150).                            */
typedef struct { float c, s, wc, ws; } Rotor;
static Rotor rotor_start(int span) {
	Rotor r;
	r.c  = cos_halfpi(0.5f / span);
	r.s  = sin_halfpi(0.5f / span);
	r.wc = cos_halfpi(1.0f / span);
	r.ws = sin_halfpi(1.0f / span);
	return r;
}
static void rotor_step(Rotor *r) {
	float c = r->c, s = r->s;
	r->c = r->wc * c - r->ws * s;
	r->s = r->ws * c + r->wc * s;
}

/* ------------------------------------------------------------------------------------------ */
/* DCT-II / DCT-IV, unnormalised, mutually recursive radix-2                                   */
/* (libfourier/source/Fourier_DCT4_Template.h:6-116, Fourier_DCT2_Template.h:6-79;            */
/*  8-point leaves Fourier_DCT4.c:64-111, Fourier_DCT2.c:64-109).                             */
/* ------------------------------------------------------------------------------------------ */
static void dct2_rec(float *x, float *t, int n);
static void dct4_rec(float *x, float *t, int n);

static void dct4_leaf8(float *x) {
	static const float R = 0x1.6A09E6p-1f;                    /* sqrt(1/2) */
	static const float C13 = 0x1.D906BCp-1f, S13 = 0x1.87DE2Ap-2f;
	static const float CK[4] = { 0x1.FD88DAp-1f, 0x1.E9F416p-1f, 0x1.C38B30p-1f, 0x1.8BC806p-1f };
	static const float SK[4] = { 0x1.917A6Cp-4f, 0x1.294062p-2f, 0x1.E2B5D4p-2f, 0x1.44CF32p-1f };
	float p[4], q[4];
	/* input rotations; odd lanes carry the DST sign flip */
	p[0] =  CK[0]*x[0] + SK[0]*x[7];   q[0] =  SK[0]*x[0] - CK[0]*x[7];
	p[1] =  CK[1]*x[1] + SK[1]*x[6];   q[1] = -SK[1]*x[1] + CK[1]*x[6];
	p[2] =  CK[2]*x[2] + SK[2]*x[5];   q[2] =  SK[2]*x[2] - CK[2]*x[5];
	p[3] =  CK[3]*x[3] + SK[3]*x[4];   q[3] = -SK[3]*x[3] + CK[3]*x[4];
	float pa = p[0] + p[3], pd = p[0] - p[3];
	float pb = p[1] + p[2], pe = p[1] - p[2];
	float qa = q[3] + q[0], qd = q[3] - q[0];
	float qb = q[2] + q[1], qe = q[2] - q[1];
	float sx = pa + pb, sy = pa - pb;
	float tx = C13*pd + S13*pe, ty = S13*pd - C13*pe;
	float ux = qa + qb, uy = qa - qb;
	float vx = C13*qd + S13*qe, vy = S13*qd - C13*qe;
	x[0] = sx;
	x[1] = tx - vy;
	x[2] = tx + vy;
	x[3] = (sy + uy) * R;
	x[4] = (sy - uy) * R;
	x[5] = ty - vx;
	x[6] = ty + vx;
	x[7] = ux;
}

static void dct2_leaf8(float *x) {
	static const float R = 0x1.6A09E6p-1f;
	static const float C14 = 0x1.F6297Cp-1f, S14 = 0x1.8F8B84p-3f;
	static const float C34 = 0x1.A9B662p-1f, S34 = 0x1.1C73B4p-1f;
	static const float C64 = 0x1.87DE2Ap-2f, S64 = 0x1.D906BCp-1f;
	float s0 = x[0]+x[7], d0 = x[0]-x[7];
	float s1 = x[1]+x[6], d1 = x[1]-x[6];
	float s2 = x[2]+x[5], d2 = x[2]-x[5];
	float s3 = x[3]+x[4], d3 = x[3]-x[4];
	float ee = s0+s3, eo = s0-s3;
	float fe = s1+s2, fo = s1-s2;
	float gx =  C34*d3 + S34*d0, gy = -S34*d3 + C34*d0;
	float hx =  C14*d2 + S14*d1, hy = -S14*d2 + C14*d1;
	float a0 = ee + fe, b0 = ee - fe;
	float c0 =  C64*fo + S64*eo, e0 = -S64*fo + C64*eo;
	float a1 = gy + hx, c1 = gy - hx;
	float e1 = gx + hy, b1 = gx - hy;
	x[0] = a0;
	x[4] = b0 * R;
	x[2] = c0;
	x[6] = e0;
	x[1] = (a1 + e1) * R;
	x[5] = b1;
	x[3] = c1;
	x[7] = (a1 - e1) * R;
}

static void dct4_rec(float *x, float *t, int n) {
	if(n == 8) { dct4_leaf8(x); return; }
	int h = n / 2;
	Rotor r = rotor_start(n);
	for(int i = 0; i < h; i++) {
		float a = x[i], b = x[n-1-i];
		t[i] = r.c*a + r.s*b;
		t[h+i] = (i & 1) ? (-r.s*a + r.c*b) : (r.s*a - r.c*b);
		rotor_step(&r);
	}
	dct2_rec(t,     x,     h);
	dct2_rec(t + h, x + h, h);
	x[0] = t[0];
	for(int i = 0; i < h-1; i++) {
		float a = t[i+1], b = t[n-1-i];
		x[2*i+1] = a + b;
		x[2*i+2] = a - b;
	}
	x[n-1] = t[h];
}

static void dct2_rec(float *x, float *t, int n) {
	if(n == 8) { dct2_leaf8(x); return; }
	int h = n / 2;
	for(int i = 0; i < h; i++) {
		float a = x[i], b = x[n-1-i];
		t[i]   = a + b;
		t[h+i] = a - b;
	}
	dct2_rec(t,     x,     h);
	dct4_rec(t + h, x + h, h);
	for(int i = 0; i < h; i++) { x[2*i] = t[i]; x[2*i+1] = t[h+i]; }
}

ORACLE_API void ulco_dct4(float *buf, float *tmp, int n) { dct4_rec(buf, tmp, n); }
ORACLE_API void ulco_dct2(float *buf, float *tmp, int n) { dct2_rec(buf, tmp, n); }

/* ------------------------------------------------------------------------------------------ */
/* IMDCT with sine-window overlap (libfourier/source/Fourier_IMDCT_Template.h:22-94, scalar). */
/*   out[n]     n output samples                                                              */
/*   in[n]      coefficients                                                                  */
/*   lap[n/2]   carried state: first half of the previous DCT-IV output                       */
/* ------------------------------------------------------------------------------------------ */
ORACLE_API void ulco_imdct(float *out, const float *in, float *lap, float *tmp, int n, int overlap) {
	int h = n / 2, flat = (n - overlap) / 2;
	memcpy(tmp, in, sizeof(float) * n);
	dct4_rec(tmp, out, n);
	for(int i = 0; i < flat; i++) {
		out[i]     = lap[h-1-i];
		out[n-1-i] = tmp[h+i];
	}
	if(flat < h) {
		Rotor r = rotor_start(overlap);
		for(int i = flat; i < h; i++) {
			float a = lap[h-1-i], b = tmp[h+i];
			out[i]     = r.c*a - r.s*b;
			out[n-1-i] = r.s*a + r.c*b;
			rotor_step(&r);
		}
	}
	memcpy(lap, tmp, sizeof(float) * h);
}

/* ------------------------------------------------------------------------------------------ */
/* MDCT + MDST of one (sub)block (libfourier/source/Fourier_MDCT_Template.h:45-187, scalar).  */
/*   fresh[n]  new samples;  lap[n] carried windowed tail of the previous call (updated)      */
/* ------------------------------------------------------------------------------------------ */
ORACLE_API void ulco_mdct_mdst(float *mdct, float *mdst, const float *fresh, float *lap, float *tmp,
                               int n, int overlap) {
	int h = n / 2, flat = (n - overlap) / 2;
	Rotor r = rotor_start(overlap ? overlap : 1);
	for(int i = 0; i < h; i++) {
		float a = lap[i], br = lap[n-1-i];
		float c = fresh[i], dr = fresh[n-1-i];
		if(i < flat) {
			lap[i] = 0.0f;
			lap[n-1-i] = dr;
			mdct[h-1-i] = c;
			mdst[h-1-i] = c;
		} else {
			lap[i]     = r.s * c;
			lap[n-1-i] = r.c * dr;
			c *= r.c; dr *= r.s;
			mdct[h-1-i] = c + dr;
			mdst[h-1-i] = c - dr;
			rotor_step(&r);
		}
		mdct[h+i] = -a + br;
		mdst[h+i] =  a + br;
	}
	for(int i = 1; i < n; i += 2) mdst[i] = -mdst[i];
	dct4_rec(mdct, tmp, n);
	dct4_rec(mdst, tmp, n);
	for(int i = 0; i < h; i++) { float v = mdst[i]; mdst[i] = mdst[n-1-i]; mdst[n-1-i] = v; }
}

/* ------------------------------------------------------------------------------------------ */
/* Shared helpers (libulc/ulcHelper.h)                                                        */
/* ------------------------------------------------------------------------------------------ */
/* Sub-block layout of a window-control code (ulcHelper.h:24-46, FormatSpecs.md:35-51):
 * fills shift[] (size = N >> shift) and returns the count; *transient = index of the
 * sub-block that carries the overlap scaling, -1 when none does (layout 0: "unused",
 * ulcHelper.h:29, but the decoder takes it - found by tests/grammar.py, not by any golden). */
ORACLE_API int ulco_subblocks(int window_ctrl, int shift[4], int *transient) {
	static const uint16_t layout[16] = {
		0x0000, 0x0008, 0x0019, 0x0091, 0x012A, 0x01A2, 0x02A1, 0x0A21,
		0x123B, 0x12B3, 0x13B2, 0x1B32, 0x23B1, 0x2B31, 0x3B21, 0xB321,
	};
	unsigned pat = layout[(window_ctrl >> 4) & 15];
	int n = 0;
	*transient = -1;
	do {
		shift[n] = pat & 7;
		if(pat & 8) *transient = n;
		n++;
	} while(pat >>= 4);
	return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Decoder (libulc/ulcDecoder.c)                                                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	int n_chan, block_size, last_sub_size;
	uint32_t rng;                 /* per-decoder here; the reference's is process-global (:75-81) */
	float *coef, *work, *lap;     /* [N], [2*nChan*N], [nChan*N/2] */
} OracleDecoder;

ORACLE_API void *ulco_decoder_new(int n_chan, int block_size) {
	/* parameter checks: ulcDecoder.c:26-35 */
	if(n_chan < 1 || n_chan > 255) return NULL;
	if(block_size < 256 || block_size > 32768 || (block_size & (block_size-1))) return NULL;
	OracleDecoder *d = calloc(1, sizeof *d);
	d->n_chan = n_chan; d->block_size = block_size;
	d->last_sub_size = 0;                                       /* :52 */
	d->rng = 1234567u;                                          /* :76 */
	d->coef = calloc(block_size, sizeof(float));
	d->work = calloc((size_t)2 * n_chan * block_size, sizeof(float));
	d->lap  = calloc((size_t)n_chan * block_size / 2, sizeof(float));
	return d;
}
ORACLE_API void ulco_decoder_free(void *p) {
	OracleDecoder *d = p;
	if(!d) return;
	free(d->coef); free(d->work); free(d->lap); free(d);
}
ORACLE_API uint32_t ulco_decoder_rng(void *p) { return ((OracleDecoder*)p)->rng; }

static uint32_t rng_next(uint32_t *s) {                         /* ulcDecoder.c:75-81 */
	uint32_t x = *s;
	x ^= x << 13; x ^= x >> 17; x ^= x << 5;
	return *s = x;
}

typedef struct { const uint8_t *p; int bits; } NybReader;
static int nyb(NybReader *r) {                                  /* ulcDecoder.c:82-88: low nybble first */
	int v = (r->p[r->bits >> 3] >> (r->bits & 4)) & 15;
	r->bits += 4;
	return v;
}
/* quantiser field after the (implicit or explicit) Fh: >=0 index, -1 stop, -2 noise tail (:89-95) */
static int read_quant(NybReader *r) {
	int q = nyb(r);
	if(q == 15) return -2;
	if(q == 14) { q += nyb(r); if(q == 29) return -1; }
	return q;
}
static float quant_scale(int q) {                               /* :96-98 ; q<0 is undefined in the reference */
	return 0x1.0p-31f * (float)(((uint32_t)1 << 26) >> (q & 31));
}

/* One sub-block of coefficients (ulcDecoder.c:99-197).  Returns 0 on a run that overshoots. */
static int parse_subblock(OracleDecoder *d, float *dst, int left, NybReader *r) {
	int q = read_quant(r);
	if(q == -1) { memset(dst, 0, sizeof(float) * left); return 1; }
	float scale = quant_scale(q);
	while(left > 0) {
		int v = nyb(r), n;
		switch(v) {
		case 0:                                                  /* 0h,X: 1..16 zeros */
			n = nyb(r) + 1;
			if(n > left) return 0;
			memset(dst, 0, sizeof(float) * n); dst += n; left -= n;
			break;
		case 1:                                                  /* 1h,Y,X: 33..288 zeros */
			n = nyb(r) << 4; n |= nyb(r); n += 33;
			if(n > left) return 0;
			memset(dst, 0, sizeof(float) * n); dst += n; left -= n;
			break;
		case 8: {                                                /* 8h,Z,Y,X: 16..527 noise */
			n = nyb(r) << 4; n |= nyb(r);
			int x = nyb(r);
			n = ((n << 1) | (x & 1)) + 16;
			int lv = (x >> 1) + 1;
			if(n > left) return 0;
			float p = (float)(lv*lv) * scale * 0.25f;
			for(int i = 0; i < n; i++) {
				if(rng_next(&d->rng) & 0x80000000u) p = -p;        /* sign toggles, :156-160 */
				*dst++ = p;
			}
			left -= n;
		} break;
		case 15:
			q = read_quant(r);
			if(q >= 0) { scale = quant_scale(q); break; }
			if(q == -2) {                                          /* Fh,Fh,Z,Y,X: decaying noise to end */
				int lv = nyb(r) + 1;
				n = nyb(r) << 4; n |= nyb(r);
				float p = (float)(lv*lv) * scale * (1.0f/16);
				float decay = 1.0f + (float)(n*n) * -0x1.0p-19f;
				while(left--) {
					if(rng_next(&d->rng) & 0x80000000u) p = -p;
					*dst++ = p; p *= decay;
				}
			} else memset(dst, 0, sizeof(float) * left);           /* Fh,Eh,Fh: zeros to end */
			return 1;
		default: {                                               /* +-2..+-7 */
			int s = (v ^ 8) - 8;
			s = s < 0 ? -(s*s) : s*s;
			*dst++ = (float)s * scale; left--;
		} break;
		}
	}
	return 1;
}

/* Parse only: header + dense coefficients for every channel/sub-block, no transform.
 * coefs[nChan*N] in coding order.  Returns bits consumed (0 = corrupt).                  */
ORACLE_API int ulco_parse_block(void *p, const uint8_t *src, float *coefs, int *window_ctrl) {
	OracleDecoder *d = p;
	NybReader r = { src, 0 };
	int wc = nyb(&r);
	wc |= (wc & 8) ? (nyb(&r) << 4) : 0x10;                      /* ulcDecoder.c:211-216 */
	*window_ctrl = wc;
	int shift[4], tr, nsub = ulco_subblocks(wc, shift, &tr);
	for(int c = 0; c < d->n_chan; c++)
		for(int j = 0; j < nsub; j++) {
			int sz = d->block_size >> shift[j];
			if(!parse_subblock(d, coefs, sz, &r)) return 0;
			coefs += sz;
		}
	return r.bits;
}

/* Full block decode (ulcDecoder.c:198-302): returns bits consumed, 0 when corrupt.
 * out = N*nChan interleaved floats.                                                       */
ORACLE_API int ulco_decode_block(void *p, float *out, const uint8_t *src) {
	OracleDecoder *d = p;
	int N = d->block_size, C = d->n_chan;
	NybReader r = { src, 0 };
	int wc = nyb(&r);
	wc |= (wc & 8) ? (nyb(&r) << 4) : 0x10;
	int shift[4], tr, nsub = ulco_subblocks(wc, shift, &tr);
	float *planar = d->work + (size_t)C * N;                     /* planar result, [C][N] */
	int last = 0;
	for(int c = 0; c < C; c++) {
		float *dst = planar + (size_t)c * N;
		float *lap = d->lap + (size_t)c * (N/2);
		last = d->last_sub_size;                                  /* reset per channel, :219 */
		for(int j = 0; j < nsub; j++) {
			int S = N >> shift[j];
			if(!parse_subblock(d, d->coef, S, &r)) return 0;
			int ov = S;
			if(j == tr) ov >>= (wc & 7);
			if(ov > last) ov = last;
			last = S;
			if(S == N) { ulco_imdct(dst, d->coef, lap, d->work, S, ov); break; }
			float *fresh = d->work + S;
			ulco_imdct(fresh, d->coef, lap, d->work, S, ov);
			/* delay line in the upper part of the lap buffer, newest sample at the lowest
			 * address (ulcDecoder.c:248-272) */
			int held = (N - S) / 2;
			float *fifo = lap + N/2;                                /* one past the oldest sample */
			if(S <= held) {
				for(int i = 0; i < S; i++) *dst++ = fifo[-1-i];
				memmove(fifo - (held - S), fifo - held, sizeof(float) * (held - S));
				for(int i = 0; i < S; i++) fifo[-(held - S) - 1 - i] = fresh[i];
			} else {
				for(int i = 0; i < held; i++) *dst++ = fifo[-1-i];
				for(int i = 0; i < S - held; i++) *dst++ = fresh[i];
				for(int i = 0; i < held; i++) fifo[-1-i] = fresh[S - held + i];
			}
		}
	}
	for(int c = 1; c < C; c += 2) {                               /* M/S undo, :281-289 */
		float *m = planar + (size_t)(c-1) * N, *s = planar + (size_t)c * N;
		for(int i = 0; i < N; i++) { float a = m[i], b = s[i]; m[i] = a + b; s[i] = a - b; }
	}
	for(int c = 0; c < C; c++)                                    /* interleave, :292-297 */
		for(int i = 0; i < N; i++) out[(size_t)i * C + c] = planar[(size_t)c * N + i];
	d->last_sub_size = last;
	return r.bits;
}

/* tools/WavIO_Helper.c:57-63 */
ORACLE_API void ulco_float_to_pcm16(int16_t *dst, const float *src, size_t n) {
	for(size_t i = 0; i < n; i++) {
		float v = src[i] * 0x1.0p+15f;
		if(v < -32768.0f) v = -32768.0f;
		if(v >  32767.0f) v =  32767.0f;
		dst[i] = (int16_t)lrintf(v);
	}
}

/* ========================================================================================== */
/* Encoder                                                                                    */
/* ========================================================================================== */
#define ORACLE_COEF_EPS 0x1.0p-31f                              /* include/ulcEncoder.h:36 */
#define N_BARK 25

typedef struct { float sum, cnt; } SegSum;                      /* struct ULC_TransientData_t */

typedef struct {
	int rate_hz, n_chan, block_size;
	int window_ctrl, next_window_ctrl;
	float complexity;
	float env[3];                 /* post-mask HP, post-mask BP, block-mask envelopes (TransientFilter) */
	SegSum seg[16];               /* [0..7] previous block, [8..15] current (TransientBuffer) */
	float *samples;               /* [2][C][N]: previous block, newest block (SampleBuffer) */
	float *coef;                  /* [C*N] normalised MDCT (TransformBuffer) */
	float *noise;                 /* [C*N] {w, w*logN} pairs (TransformNoise) */
	float *lap;                   /* [C][N] (TransformFwdLap) */
	float *tmp;                   /* [max(2,C)*N] (TransformTemp) - also the output byte buffer */
	int   *rank;                  /* [C*N] keys as floats, then ranks (TransformIndex) */
	float *dbg_key, *dbg_mask;    /* test taps: final sort keys [C*N], masking curve [N/2] */
	int    dbg_n_nz, dbg_n_out;   /* test taps: codeable coefficients, coefficients requested from the last pass */
} OracleEncoder;

ORACLE_API void *ulco_encoder_new(int rate_hz, int n_chan, int block_size) {
	if(n_chan < 1 || n_chan > 255) return NULL;                  /* ulcEncoder.c:29-34 */
	if(block_size < 256 || block_size > 32768 || (block_size & (block_size-1))) return NULL;
	OracleEncoder *e = calloc(1, sizeof *e);
	size_t cn = (size_t)n_chan * block_size;
	e->rate_hz = rate_hz; e->n_chan = n_chan; e->block_size = block_size;
	e->next_window_ctrl = 0x10;                                  /* ulcEncoder.c:70 */
	e->samples = calloc(2*cn, sizeof(float));
	e->coef    = calloc(cn, sizeof(float));
	e->noise   = calloc(cn, sizeof(float));
	e->lap     = calloc(cn, sizeof(float));
	e->tmp     = calloc((size_t)(n_chan < 2 ? 2 : n_chan) * block_size, sizeof(float));
	e->rank    = calloc(cn, sizeof(int));
	e->dbg_key = calloc(cn, sizeof(float));
	e->dbg_mask = calloc(block_size/2, sizeof(float));
	return e;
}
ORACLE_API void ulco_encoder_free(void *p) {
	OracleEncoder *e = p;
	if(!e) return;
	free(e->samples); free(e->coef); free(e->noise); free(e->lap); free(e->tmp); free(e->rank); free(e->dbg_key); free(e->dbg_mask); free(e);
}
ORACLE_API int   ulco_encoder_window_ctrl(void *p) { return ((OracleEncoder*)p)->window_ctrl; }
ORACLE_API float ulco_encoder_complexity(void *p)  { return ((OracleEncoder*)p)->complexity; }
ORACLE_API const float *ulco_encoder_coefs(void *p) { return ((OracleEncoder*)p)->coef; }
ORACLE_API const int   *ulco_encoder_ranks(void *p) { return ((OracleEncoder*)p)->rank; }
ORACLE_API const float *ulco_encoder_noise(void *p) { return ((OracleEncoder*)p)->noise; }
ORACLE_API const float *ulco_encoder_keys(void *p)  { return ((OracleEncoder*)p)->dbg_key; }
ORACLE_API const float *ulco_encoder_mask(void *p)  { return ((OracleEncoder*)p)->dbg_mask; }
ORACLE_API int ulco_encoder_n_nz(void *p)  { return ((OracleEncoder*)p)->dbg_n_nz; }
ORACLE_API int ulco_encoder_n_out(void *p) { return ((OracleEncoder*)p)->dbg_n_out; }

/* ulcHelper.h:127-136 */
static float fast_log(float x) {
	uint32_t b; memcpy(&b, &x, 4);
	int32_t e = (int32_t)(b >> 23) - 127;
	b = (127u << 23) | (b & 0x7FFFFFu);
	memcpy(&x, &b, 4);
	return -1.7417939f + (2.8212026f + (-1.4699568f + (0.44717955f - 0.056570851f*x)*x)*x)*x + 0.6931471806f*e;
}
/* ulcHelper.h:51-91 */
static int compand_u(float v) { return (v >= 0.5f) ? (int)(0.5f + sqrtf(v - 0.25f)) : 0; }
static int compand_clip(float v, int limit) {
	int q = compand_u(v < 0 ? -v : v);
	if(q > limit) q = limit;
	return v < 0.0f ? -q : q;
}
/* ulcHelper.h:96-120 */
static float bark_to_line(float bark, float nyquist, int n) {
	float hz = 600.0f * sinhf(bark * (1.0f/6.0f));
	return (hz * (float)n / nyquist) - 0.5f;
}
static float line_to_bark(int line, float nyquist, int n) {
	float hz = ((float)line + 0.5f) * nyquist / (float)n;
	return 6.0f * asinhf(hz * (1.0f/600.0f));
}

/* ---- transient detector (libulc/ulcEncoder_WindowControl.c) ---- */
static void transient_filter(OracleEncoder *e, float *energy) {
	int N = e->block_size, C = e->n_chan, lag = N/2;
	const float *blocks = e->samples;
	memset(energy, 0, sizeof(float) * 2 * N);
	for(int c = 0; c < C; c++) {                                  /* :55-70, UpdateFilters :31-40 */
		const float *old = blocks + (size_t)c*N + (N - lag);
		const float *new = blocks + (size_t)c*N + (size_t)C*N;
		float t0, t1 = old[-1], t2 = old[0];
		int n_old = N - lag - 1;
		for(int n = 0; n < N; n++) {
			float x = (n < n_old) ? old[1+n] : new[n - n_old];
			t0 = t1; t1 = t2; t2 = x;
			float hp = -t0 + 2*t1 - t2, bp = -t0 + t2;
			energy[2*n+0] += hp*hp;
			energy[2*n+1] += bp*bp;
		}
	}
	float post_hp = e->env[0], post_bp = e->env[1];               /* :73-88 */
	float k_hp = expf(-0x1.CC845Cp6f / e->rate_hz), k_bp = expf(-0x1.596344p8f / e->rate_hz);
	for(int n = 0; n < N; n++) {
		float d_hp = sqrtf(energy[2*n+0]) - post_hp, d_bp = sqrtf(energy[2*n+1]) - post_bp;
		post_hp += d_hp * (1.0f - k_hp);
		post_bp += d_bp * (1.0f - k_bp);
		energy[2*n+0] = post_hp; energy[2*n+1] = post_bp;
	}
	e->env[0] = post_hp; e->env[1] = post_bp;
	float pre_hp = post_hp, pre_bp = post_bp;                     /* :92-104 */
	float r_hp = expf(-0x1.CC845Cp7f / e->rate_hz), r_bp = expf(-0x1.596344p8f / e->rate_hz);
	for(int n = N-1; n >= 0; n--) {
		float d_hp = energy[2*n+0] - pre_hp, d_bp = energy[2*n+1] - pre_bp;
		pre_hp += d_hp * (1.0f - r_hp);
		pre_bp += d_bp * (1.0f - r_bp);
		float a = d_hp*pre_bp, b = d_bp*pre_hp;
		energy[2*n+0] = a*a + b*b;
	}
	int bin = N / 8;                                              /* :107-134 */
	float k_blk = expf(-0x1.1AF110p-6f * N / e->rate_hz);
	for(int s = 0; s < 8; s++) {
		e->seg[s] = e->seg[8+s];
		SegSum acc = { 0.0f, 0.0f };
		float m = e->env[2];
		for(int n = 0; n < bin; n++) {
			float d = energy[2*(s*bin+n)] - m;
			m += d * (1.0f - k_blk);
			acc.sum += m; acc.cnt += 1;
		}
		e->env[2] = m;
		e->seg[8+s] = acc;
	}
}

static int window_ctrl_decide(OracleEncoder *e) {                /* :140-239 */
	int N = e->block_size;
	transient_filter(e, e->tmp);
	const SegSum *cur = e->seg + 8;
	int log2sub = 31 - __builtin_clz(N/8);
	int n_seg = 8, seg_sz = 1;
	if(log2sub < 6) { int sh = 6 - log2sub; n_seg >>= sh; seg_sz <<= sh; log2sub = 6; }
	int decimation = 1; float tr_ratio = 0.0f;
	for(;;) {
		log2sub++;
		int best = 0; float max_ratio = -1000.0f;
		for(int s = 0; s < n_seg; s++) {
			SegSum L = {0,0}, R = {0,0};
			const SegSum *src = cur + s*seg_sz;
			for(int n = 0; n < seg_sz; n++) {
				L.sum += src[n-seg_sz].sum; L.cnt += src[n-seg_sz].cnt;
				R.sum += src[n].sum;        R.cnt += src[n].cnt;
			}
			float l = L.sum ? logf(L.sum / L.cnt) : -100.0f;
			float r = R.sum ? logf(R.sum / R.cnt) : -100.0f;
			float ratio = r - l; if(ratio < 0) ratio = -ratio;
			if(ratio > max_ratio) { best = s; max_ratio = ratio; }
		}
		if(max_ratio - tr_ratio < 0x1.62E430p-1f) break;
		decimation = n_seg + best;
		tr_ratio = max_ratio;
		if(n_seg > 1 && tr_ratio < 0x1.62E430p-1f) { n_seg /= 2; seg_sz *= 2; } else break;
	}
	if(tr_ratio < 0x1.62E430p-2f) return 0x10;
	tr_ratio *= 0x1.715476p0f;
	int ov = (tr_ratio < 0.5f) ? 0 : (tr_ratio >= 6.5f) ? 7 : (int)lrintf(tr_ratio);
	if(log2sub - ov < 6) ov = log2sub - 6;
	return ov + 8*(decimation != 1) + 16*decimation;
}

/* ---- Bark-band analysis (libulc/ulcEncoder_Psyopt.c) ---- */
typedef struct { int end; double lg, wlg, w; } BandAcc;           /* struct LineSum_t, :16-51 */
static void band_acc_to(BandAcc *a, const float *v, const float *lg, int end) {
	for(int i = a->end; i < end; i++) {
		double x = v[i], l = lg[i];
		a->lg += l; a->wlg += l*x; a->w += x;
	}
	a->end = end;
}
static void band_edges(float b0, float b1, float nyq, int n, int *beg, int *end) {
	int lo = (int)floorf(bark_to_line(b0, nyq, n)), hi = (int)ceilf(bark_to_line(b1, nyq, n));
	if(lo < 0) lo = 0;
	if(hi < 0) hi = 0;
	if(lo > n-1) lo = n-1;
	if(hi > n) hi = n;
	*beg = lo; *end = hi;
}
static float bark_interp(const float *band, int line, float nyq, int n) {
	float bb = line_to_bark(line, nyq, n);
	int idx = (int)bb;
	float fr = bb - (float)idx;
	float l = (idx < N_BARK) ? band[idx] : band[N_BARK-1];
	float r = (idx+1 < N_BARK) ? band[idx+1] : l;
	return l*(1.0f-fr) + r*fr;
}

/* :168-250 - in: data[0..n/2) = sum of Re^2+Im^2 of line pairs; out: data[0..n) = {w, w*(level+ln2)} pairs */
static void noise_log_spectrum(float *data, float *work, int n, int rate_hz) {
	float nyq = (float)rate_hz * 0.5f;
	n /= 2;
	float *lg = work, *band = work + n;
	for(int i = 0; i < n; i++) lg[i] = fast_log(0x1.0p-126f + data[i]);
	float level = -100.0f;
	BandAcc lo = {0,0,0,0}, hi = {0,0,0,0};
	for(int b = 0; b < N_BARK; b++) {
		int beg, end;
		band_edges((float)b, (float)b + 2.0f, nyq, n, &beg, &end);
		band_acc_to(&lo, data, lg, beg);
		band_acc_to(&hi, data, lg, end);
		double s_lg = hi.lg - lo.lg, s_wlg = hi.wlg - lo.wlg, s_w = hi.w - lo.w;
		if(s_w > 0.0) {
			double inv = 1.0 / (double)(end - beg);
			s_wlg = s_wlg / s_w;
			s_lg  = s_lg * inv;
			level = 0.5f * (float)(log(s_w * inv) + s_lg - s_wlg);
		}
		band[b] = level;
	}
	for(int i = 0; i < n; i++) {
		float v = bark_interp(band, i, nyq, n);
		float w = expf(0.5f * v);
		data[2*i+0] = w;
		data[2*i+1] = w * (v + 0x1.62E430p-1f);
	}
}

/* :60-155 - amp2[0..N/2) channel-summed line energies (trashed); mask[0..N/2) out */
static void psycho_masks(float *mask, float *amp2, float *work, int N, int rate_hz, int window_ctrl) {
	float nyq = (float)rate_hz * 0.5f;
	N /= 2;
	for(int i = 0; i < N; i++) mask[i] = fast_log(0x1.0p-126f + amp2[i]);
	int shift[4], tr, nsub = ulco_subblocks(window_ctrl, shift, &tr);
	for(int j = 0; j < nsub; j++) {
		int n = N >> shift[j];
		float ratio = 0.0f, *band = work;
		BandAcc lo = {0,0,0,0}, hi = {0,0,0,0};
		for(int b = 0; b < N_BARK; b++) {
			int beg, end;
			band_edges((float)b - 0.75f, (float)b + 0.25f, nyq, n, &beg, &end);
			band_acc_to(&lo, amp2, mask, beg);
			band_acc_to(&hi, amp2, mask, end);
			double s_lg = hi.lg - lo.lg, s_wlg = hi.wlg - lo.wlg, s_w = hi.w - lo.w;
			if(s_w > 0.0) {
				s_wlg = s_wlg / s_w;
				s_lg  = s_lg / (double)(end - beg);
				ratio = (float)(s_wlg - s_lg - log(s_w));
			}
			band[b] = ratio;
		}
		for(int i = 0; i < n; i++) mask[i] = bark_interp(band, i, nyq, n);
		mask += n; amp2 += n;
	}
}

/* ---- ranking: the reference's heapsort, comparison for comparison, so that ties fall the same way
 * (libulc/ulcEncoder_BlockTransform.c:20-77) ---- */
static void sift(const float *key, int *ord, int root, int n) {
	int child = 2*root + 1;
	while(child < n) {
		int ri = ord[root], ci = ord[child];
		if(child+1 < n && key[ord[child+1]] < key[ci]) { child++; ci = ord[child]; }
		if(key[ci] > key[ri]) return;
		ord[root] = ci; ord[child] = ri;
		root = child; child = 2*root + 1;
	}
}
static void rank_by_key(int *rank_out, const float *key, int *ord, int n) {
	for(int i = 0; i < n; i++) ord[i] = i;
	for(int i = n/2 - 1; i >= 0; i--) sift(key, ord, i, n);
	for(int i = n-1; i > 0; i--) {
		rank_out[ord[0]] = i;
		ord[0] = ord[i];
		sift(key, ord, 0, i);
	}
	rank_out[ord[0]] = 0;
}

/* ---- block transform (libulc/ulcEncoder_BlockTransform.c:82-357).  Returns the number of codeable coefs. ---- */
static int transform_block(OracleEncoder *e, const float *pcm) {
	int N = e->block_size, C = e->n_chan;
	float *prev = e->samples, *cur = e->samples + (size_t)C*N;
	memcpy(prev, cur, sizeof(float) * C * N);                     /* :92 */
	for(int c = 0; c < C; c++) for(int n = 0; n < N; n++) cur[c*N+n] = pcm[n*C+c];
	for(int c = 1; c < C; c += 2)                                 /* M/S, :102-111 */
		for(int n = 0; n < N; n++) {
			float a = cur[(c-1)*N+n], b = cur[c*N+n];
			cur[(c-1)*N+n] = (a+b) * 0.5f;
			cur[c*N+n]     = (a-b) * 0.5f;
		}
	int wc = e->window_ctrl = e->next_window_ctrl;                /* :114-128 */
	int next_wc = e->next_window_ctrl = window_ctrl_decide(e);
	int next_ov; {
		int sh[4], tr; ulco_subblocks(next_wc, sh, &tr);
		next_ov = N >> sh[0];
		if(tr == 0 && (next_wc >> 4) != 0) next_ov >>= (next_wc & 7);   /* first sub-block carries the flag */
	}
	int shift[4], tr, nsub = ulco_subblocks(wc, shift, &tr);
	int has_tr = (wc >> 4) != 0;                                  /* layout 0 has no flagged sub-block */

	int n_nz = 0;
	float *key = (float*)e->rank;
	float *smp = e->samples, *mdct = e->coef, *noise = e->noise, *lap = e->lap;
	float *work = e->tmp, *amp2 = e->tmp + N, *mask = e->samples; /* aliasing as in :139-151 */
	memset(noise, 0, sizeof(float) * C * N);
	memset(amp2, 0, sizeof(float) * (N/2));
	float sum_sq = 0.0f, sum_abs = 0.0f;
	for(int c = 0; c < C; c++) {
		for(int j = 0; j < nsub; j++) {
			int S = N >> shift[j], ov;
			if(j+1 < nsub) { ov = N >> shift[j+1]; if(has_tr && tr == j+1) ov >>= (wc & 7); }
			else ov = next_ov;
			if(ov > S) ov = S;
			/* sample FIFO beside the centred lap segment (:175-224) */
			int held = (N - S) / 2;
			float *fifo = lap + (N + S)/2;
			if(held < S) {
				memcpy(work, fifo, sizeof(float) * held);
				memcpy(work + held, smp, sizeof(float) * (S - held));
				memcpy(fifo, smp + (S - held), sizeof(float) * held);
			} else {
				memcpy(work, fifo, sizeof(float) * S);
				memmove(fifo, fifo + S, sizeof(float) * (held - S));
				memcpy(fifo + (held - S), smp, sizeof(float) * S);
			}
			float *mdst = smp;                                      /* :227 */
			ulco_mdct_mdst(mdct, mdst, work, lap + (N - S)/2, work, S, ov);
			float norm = 2.0f / S;                                  /* :243-281 */
			for(int n = 0; n < S; n++) {
				float re = (mdct[n] *= norm), re2 = re*re;
				float im = mdst[n] * norm,    im2 = im*im;
				float are = re < 0 ? -re : re;
				float a2 = re2 + im2;
				if(are < 0.5f*ORACLE_COEF_EPS) key[n] = -INFINITY;
				else { key[n] = fast_log(re2); n_nz++; }
				noise[n/2] += a2;
				amp2[n/2]  += a2;
				sum_sq += re2; sum_abs += are;
			}
			noise_log_spectrum(noise, work, S, e->rate_hz);
			smp += S; mdct += S; key += S; amp2 += S/2; noise += S;
		}
		lap += N; amp2 -= N/2;
	}
	mdct -= (size_t)C*N; key -= (size_t)C*N;
	float cx = sum_sq;                                             /* :310-325 */
	if(cx) {
		float scale = 0x1.62E430p-1f * (31 - __builtin_clz(N));
		cx = logf(sum_abs*sum_abs / sum_sq) / scale;
		if(cx < 0.0f) cx = 0.0f;
		if(cx > 1.0f) cx = 1.0f;
	}
	e->complexity = cx;
	psycho_masks(mask, amp2, work, N, e->rate_hz, wc);             /* :329 */
	for(int c = 0; c < C; c++)                                     /* :337-345 */
		for(int n = 0; n < N; n++)
			key[c*N+n] = 2*key[c*N+n] + mask[n/2] + -0x1.62E430p0f*(c & 1);
	memcpy(e->dbg_key, key, sizeof(float) * C * N);
	memcpy(e->dbg_mask, mask, sizeof(float) * (N/2));
	e->dbg_n_nz = n_nz;
	rank_by_key(e->rank, (const float*)e->rank, (int*)e->tmp, C*N); /* :350-354 */
	return n_nz;
}

/* ---- noise-fill parameter fits (libulc/ulcEncoder_NoiseFill.c) ---- */
static int noise_q(const float *pairs, int band, int n, float q) { /* :15-36 */
	pairs += band / 2 * 2;
	n = (n + (band & 1) + 1) / 2;
	float s = 0.0f, sw = 0.0f;
	for(int i = 0; i < n; i++) { float w = pairs[2*i], wy = pairs[2*i+1]; s += wy; sw += w; }
	if(s == 0.0f) return 0;
	float amp = expf(s / sw);
	int r = compand_u(amp * q);
	return r < 8 ? r : 8;
}
static void hf_ext_params(const float *pairs, int band, int n, float q, int *nq, int *decay_q) { /* :41-94 */
	pairs += band / 2 * 2;
	n = (n + (band & 1) + 1) / 2;
	float sx = 0, sx2 = 0, sxy = 0, sy = 0, sw = 0;
	for(int i = 0; i < n; i++) {
		float x = i * 2.0f, w = pairs[2*i], wy = pairs[2*i+1];
		sx += w*x; sx2 += w*x*x; sxy += x*wy; sy += wy; sw += w;
	}
	float det = sw*sx2 - sx*sx;
	if(det == 0.0f) { *nq = *decay_q = 0; return; }
	float amp = (sx2*sy - sx*sxy) / det, dec = (sw*sxy - sx*sy) / det;
	amp = expf(amp);
	dec = (dec < 0.0f) ? expf(dec) : 1.0f;
	int a = compand_u(amp * q * 4.0f); if(a > 16) a = 16;
	int d = compand_u((dec - 1.0f) * -0x1.0p19f);
	if(!d) return;
	if(d > 0xFF) d = 0xFF;
	*nq = a; *decay_q = d;
}

/* ---- nybble coder (libulc/ulcEncoder_Encode.c) ---- */
typedef struct { uint8_t *dst; int bits; } NybWriter;
static void put(NybWriter *w, int x) {                            /* :23-29 */
	uint8_t *p = &w->dst[w->bits >> 3];
	*p = (uint8_t)((*p >> 4) | ((x & 15) << 4));
	w->bits += 4;
}
static void put_quant(NybWriter *w, int qi, int lead) {           /* :32-45 */
	int s = qi - 5;
	if(lead) put(w, 0xF);
	if(s < 0xE) put(w, s);
	else { put(w, 0xE); put(w, s - 0xE); }
}
static int make_quant(float peak) {                                /* :83-88 */
	int q = (int)(0x1.657006p2f + -0x1.715476p0f * logf(peak));
	if(q < 5) q = 5;
	if(q > 31) q = 31;
	return q;
}
/* one quantiser zone [cur, end): returns the index after the last coded coefficient (:92-196) */
static int put_zone(const OracleEncoder *e, NybWriter *w, int cur, int end, float q, int next_coded, int n_out) {
	const float *coef = e->coef, *pairs = e->noise; const int *rank = e->rank;
	for(;;) {
		while(cur < end && rank[cur] >= n_out) cur++;
		if(cur >= end) break;
		float sc = coef[cur] * q;
		if((sc < 0 ? -sc : sc) < 2.5f) { cur++; continue; }
		int gap = cur - next_coded;
		while(gap) {
			int n, v;
			if(gap <= 2) {
				int q1 = compand_clip(coef[next_coded] * q, 7), q2 = 0;
				if(gap >= 2) q2 = compand_clip(coef[next_coded+1] * q, 7);
				if(abs(q1) > 1 && (gap < 2 || abs(q2) > 1)) {
					put(w, q1);
					if(gap >= 2) put(w, q2);
					next_coded += gap;
					break;
				}
			}
			int nq = 0;
			if(gap >= 16) {
				v = gap - 16; if(v > 0x1FF) v = 0x1FF;
				n = v + 16;
				nq = noise_q(pairs, next_coded, n, q);
			}
			if(nq) {
				put(w, 0x8); put(w, v >> 5); put(w, v >> 1); put(w, (v & 1) | ((nq-1) << 1));
			} else if(gap < 33) {
				v = gap - 1; if(v > 0xF) v = 0xF;
				n = v + 1;
				put(w, 0x0); put(w, v);
			} else {
				v = gap - 33; if(v > 0xFF) v = 0xFF;
				n = v + 33;
				put(w, 0x1); put(w, v >> 4); put(w, v);
			}
			next_coded += n; gap -= n;
		}
		put(w, compand_clip(coef[cur] * q, 7));
		next_coded++; cur++;
	}
	return next_coded;
}
static void put_subblock(const OracleEncoder *e, NybWriter *w, int idx, int S, int n_out) { /* :200-316 */
	const float *coef = e->coef; const int *rank = e->rank;
	int end = idx + S, next_coded = idx, prev_q = -1, zone_start = -1;
	float lo = 1000.0f, hi = -1000.0f;
	do {
		while(idx < end && rank[idx] >= n_out) idx++;
		float new_lo = 0.0f, new_hi = hi, level = 0.0f;
		if(idx < end) {
			level = coef[idx] < 0 ? -coef[idx] : coef[idx];
			new_lo = level < lo ? level : lo;
			new_hi = level > hi ? level : hi;
			if(zone_start == -1) zone_start = idx;
		}
		if(new_hi > new_lo * 4.0f) {
			int qi = make_quant(hi);
			if(qi != prev_q) { put_quant(w, qi, prev_q != -1); prev_q = qi; }
			next_coded = put_zone(e, w, zone_start, idx, (float)(1u << qi), next_coded, n_out);
			zone_start = idx;
			lo = hi = level;
		} else { lo = new_lo; hi = new_hi; }
	} while(++idx <= end);
	int n = end - next_coded;
	if(n > 4) {
		if(prev_q != -1) put(w, 0xF);
		int nq = 0, dq = 0;
		if(prev_q != -1 && n >= 16) hf_ext_params(e->noise, next_coded, n, (float)(1u << prev_q), &nq, &dq);
		if(nq) { put(w, 0xF); put(w, nq-1); put(w, dq >> 4); put(w, dq); }
		else   { put(w, 0xE); put(w, 0xF); }
	} else if(n > 0) { put(w, 0x0); put(w, n-1); }
}
static int encode_pass(const OracleEncoder *e, uint8_t *dst, int n_out) {   /* :319-360 */
	NybWriter w = { dst, 0 };
	((OracleEncoder*)e)->dbg_n_out = n_out;
	int wc = e->window_ctrl;
	put(&w, wc);
	if(wc & 8) put(&w, wc >> 4);
	int shift[4], tr, nsub = ulco_subblocks(wc, shift, &tr), idx = 0;
	for(int c = 0; c < e->n_chan; c++)
		for(int j = 0; j < nsub; j++) {
			int S = e->block_size >> shift[j];
			put_subblock(e, &w, idx, S, n_out);
			idx += S;
		}
	if(w.bits & 4) dst[w.bits >> 3] >>= 4;
	return (w.bits + 7) & ~7;
}

/* ---- rate control (libulc/ulcEncoder.c:93-158).  mode 0 = VBR(quality), 1 = CBR(kbps), 2 = ABR(kbps, avg) ---- */
static int cbr_core(OracleEncoder *e, uint8_t *dst, float kbps, int max_coef) {
	int size = 0, n_out = -1;
	int budget = (int)((e->block_size * kbps) * 1000.0f / e->rate_hz);
	int lo = 0, hi = max_coef;
	if(lo < hi) do {
		n_out = (lo + hi) / 2u;
		size = encode_pass(e, dst, n_out);
		if(size < budget) lo = n_out;
		else if(size > budget) hi = n_out - 1;
		else { lo = n_out; break; }
	} while(lo < hi - 1);
	if(lo != n_out) size = encode_pass(e, dst, lo);
	return size;
}
/* Returns bits (multiple of 8); bytes are at *out (valid until the next call). pcm = N*C interleaved floats. */
ORACLE_API int ulco_encode_block(void *p, const float *pcm, int mode, float param, float avg_complexity, const uint8_t **out) {
	OracleEncoder *e = p;
	uint8_t *dst = (uint8_t*)e->tmp;
	int max_coef = transform_block(e, pcm), size;
	if(mode == 1) size = cbr_core(e, dst, param, max_coef);
	else if(mode == 2) size = cbr_core(e, dst, param * e->complexity / avg_complexity, max_coef);
	else {
		float target = 0x1.E4EFB7p3f * logf(100.0f / param);
		int n_out = max_coef;
		if(target > 0.0f) {
			float f = (e->n_chan * e->block_size) * e->complexity / target;
			if(f < max_coef) n_out = (int)f;
		}
		size = encode_pass(e, dst, n_out);
	}
	*out = dst;
	return size;
}
