/* ref_harness.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the UNMODIFIED reference libulc (objects compiled by oracle/Makefile from
 * /root/reference) through its public C API and dumps everything the parity tests
 * compare against.  One process = one stream, because the reference decoder's
 * noise RNG is a process-global static (libulc/ulcDecoder.c:75-81).
 *
 *   ulcref_harness enc <vbr|cbr|abr> <param> <avgComplexity> <blockSize> <nChan> <rateHz>
 *                      <in.s16> <out.bytes> <out.meta> [<out.coefs> [<out.ranks>]]
 *       in.s16    raw interleaved little-endian PCM16; converted with x * 2^-15 exactly as
 *                 tools/WavIO_Helper.c:49-55 does; zero padded at EOF (tools/WavIO_Reader.c:142-145)
 *       nBlocks = ceil(nSamples/blockSize) + 2                   (tools/ulcEncodeTool.c:98)
 *       out.bytes concatenated byte-aligned blocks                (tools/ulcEncodeTool.c:162-168)
 *       out.meta  per block: u32 nBytes, i32 WindowCtrl, f32 BlockComplexity, u32 0
 *       out.coefs per block: f32 TransformBuffer[nChan*blockSize] (normalised MDCT coefficients)
 *       out.ranks per block: i32 TransformIndex[nChan*blockSize]
 *
 *   ulcref_harness dec <blockSize> <nChan> <nBlocks> <in.bytes> <out.f32> <out.bits>
 *       decodes nBlocks consecutive blocks the way tools/ulcDecodeTool.c:135-165 does
 *       (advance by (bits+7)/8 bytes); out.f32 = interleaved float PCM, out.bits = i32 per block
 *       (0 = corrupt; decoding stops there and the remaining output is left zero).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ulcEncoder.h"
#include "ulcDecoder.h"

static void *slurp(const char *path, size_t *len, size_t pad) {
	FILE *f = fopen(path, "rb");
	if(!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
	fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
	char *p = calloc(1, (size_t)n + pad + 1);
	if(n && fread(p, 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read %s\n", path); exit(2); }
	fclose(f);
	*len = (size_t)n;
	return p;
}

static int do_enc(int argc, char **argv) {
	if(argc < 11) return 1;
	const char *mode = argv[2];
	float param = (float)atof(argv[3]);
	float avgc  = (float)atof(argv[4]);
	int N = atoi(argv[5]), nChan = atoi(argv[6]), rate = atoi(argv[7]);
	size_t len; int16_t *pcm = slurp(argv[8], &len, 0);
	size_t nSamples = len / (2u * (size_t)nChan);
	size_t nBlocks  = (nSamples + N - 1) / N + 2;
	FILE *fb = fopen(argv[9], "wb"), *fm = fopen(argv[10], "wb");
	FILE *fc = argc > 11 ? fopen(argv[11], "wb") : NULL;
	FILE *fr = argc > 12 ? fopen(argv[12], "wb") : NULL;
	struct ULC_EncoderState_t st;
	st.RateHz = rate; st.nChan = nChan; st.BlockSize = N;
	if(ULC_EncoderState_Init(&st) <= 0) { fprintf(stderr, "encoder init failed\n"); return 3; }
	float *blk = aligned_alloc(64, sizeof(float) * N * nChan);   /* the tools align to 64 B (tools/ulc_Helper.h:6) */
	for(size_t b = 0; b < nBlocks; b++) {
		for(size_t i = 0; i < (size_t)N * nChan; i++) {
			size_t s = b * (size_t)N * nChan + i;
			blk[i] = (s < nSamples * nChan) ? (float)pcm[s] * 0x1.0p-15f : 0.0f;
		}
		int bits = 0; const void *out;
		if(!strcmp(mode, "vbr"))      out = ULC_EncodeBlock_VBR(&st, blk, &bits, param);
		else if(!strcmp(mode, "abr")) out = ULC_EncodeBlock_ABR(&st, blk, &bits, param, avgc);
		else                          out = ULC_EncodeBlock_CBR(&st, blk, &bits, param);
		uint32_t nb = (uint32_t)(bits + 7) / 8u;
		fwrite(out, 1, nb, fb);
		struct { uint32_t nb; int32_t wc; float cx; uint32_t z; } m = { nb, st.WindowCtrl, st.BlockComplexity, 0 };
		fwrite(&m, sizeof m, 1, fm);
		if(fc) fwrite(st.TransformBuffer, sizeof(float), (size_t)N * nChan, fc);
		if(fr) fwrite(st.TransformIndex,  sizeof(int),   (size_t)N * nChan, fr);
	}
	ULC_EncoderState_Destroy(&st);
	fclose(fb); fclose(fm); if(fc) fclose(fc); if(fr) fclose(fr);
	return 0;
}

static int do_dec(int argc, char **argv) {
	if(argc < 8) return 1;
	int N = atoi(argv[2]), nChan = atoi(argv[3]); size_t nBlocks = (size_t)atol(argv[4]);
	size_t len; uint8_t *bytes = slurp(argv[5], &len, (size_t)4 * N * nChan + 64);
	FILE *fo = fopen(argv[6], "wb"), *fbits = fopen(argv[7], "wb");
	struct ULC_DecoderState_t st;
	st.nChan = nChan; st.BlockSize = N;
	if(ULC_DecoderState_Init(&st) <= 0) { fprintf(stderr, "decoder init failed\n"); return 3; }
	float *out = aligned_alloc(64, sizeof(float) * (size_t)N * nChan);  /* AVX build stores with aligned moves */
	memset(out, 0, sizeof(float) * (size_t)N * nChan);
	size_t pos = 0; int dead = 0;
	for(size_t b = 0; b < nBlocks; b++) {
		int32_t bits = 0;
		if(!dead && pos < len) {
			bits = ULC_DecodeBlock(&st, out, bytes + pos);
			if(!bits) { dead = 1; memset(out, 0, sizeof(float) * N * nChan); }
			pos += (size_t)(bits + 7) / 8u;
		} else memset(out, 0, sizeof(float) * N * nChan);
		fwrite(out, sizeof(float), (size_t)N * nChan, fo);
		fwrite(&bits, sizeof bits, 1, fbits);
	}
	ULC_DecoderState_Destroy(&st);
	fclose(fo); fclose(fbits);
	return 0;
}

int main(int argc, char **argv) {
	int r = 1;
	if(argc > 1 && !strcmp(argv[1], "enc")) r = do_enc(argc, argv);
	if(argc > 1 && !strcmp(argv[1], "dec")) r = do_dec(argc, argv);
	if(r == 1) fprintf(stderr, "usage: see header of oracle/ref_harness.c\n");
	return r;
}
