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
 * sub-block that carries the overlap scaling.                                             */
ORACLE_API int ulco_subblocks(int window_ctrl, int shift[4], int *transient) {
	static const uint16_t layout[16] = {
		0x0000, 0x0008, 0x0019, 0x0091, 0x012A, 0x01A2, 0x02A1, 0x0A21,
		0x123B, 0x12B3, 0x13B2, 0x1B32, 0x23B1, 0x2B31, 0x3B21, 0xB321,
	};
	unsigned pat = layout[(window_ctrl >> 4) & 15];
	int n = 0;
	*transient = 0;
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
