// ulcb_decode_ws.cu - warp-specialised batched decoder, BlockSize 2048 stereo (the benchmark path).
//
// One 96-thread CTA walks one stream front to back (libulc/ulcDecoder.c:198-302 is sequential per stream);
// its three warps form a pipeline with no CTA-wide barrier:
//
//   warp 0  PARSE      nybble grammar of block k+1 (ulcDecoder.c:99-197) -> dequantised coefficients as bf16
//                      (every value the format can express is n * 2^-e with n <= 64: exact in bf16) plus one
//                      8-byte record per noise run.  The stream is read in 256-nybble windows, one 32-bit word
//                      per lane; token lengths and stop codes of all 8 start offsets come from SWAR nybble
//                      compares, "which offset does the chain enter my word at" is a warp scan of 4-state
//                      transition maps (composition = one PRMT), counts are ordinary shuffle scans.
//   warp 1  TRANSFORM  Mid channel of block k    } noise fill from the records + sign-parity table
//   warp 2  TRANSFORM  Side channel of block k   } (ulcDecoder.c:75-81,156-184), IMDCT, un-fold, M/S, PCM out
//
// A long block's IMDCT (Fourier_IMDCT_Template.h:22-94) never leaves the register file: the DCT-IV is a
// 1024-point complex FFT between two twiddles, done as 32 x 32 (32-point FFT per lane, twiddle, transpose
// through 4 KB of shared memory, 32-point FFT per lane).  With X = DCT-IV output, lane/register k holds
// y[k] = (X[2k], -X[N-1-2k]); the un-fold pairs X[N/2+i] with the pending tail sample i, and those turn out
// to be y[k] and the previous block's y[k] at the SAME k, so the pending tail lives in a lane-private layout
// ("k-layout": slot k holds time sample 2k-N/2 for k >= N/4, N/2-1-2k below) and is only re-ordered when a
// window-switched block (generic shared-memory path, time order) is met.
//
// Hand-offs use named barriers (bar.arrive / bar.sync): FULL (parse -> transforms), EMPTY (transforms have
// loaded the coefficients -> parse may overwrite), XCHG (the two transform warps swap half of their output
// so that each interleaves and stores one half of the block: 8- or 16-byte stores, 256 B+ per warp instruction).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <mutex>
#include "ulcb_dct.cuh"
#include "ulcb_decode.h"

#ifdef ULCB_PHASE_TIMING
__device__ unsigned long long g_ulcbWsPhase[32];
#define WTK(i) do { long long t_ = clock64(); if (lane == 0) acc[i] += t_ - tPrev; tPrev = t_; } while (0)
#define WCNT(i) do { if (lane == 0) acc[i] += 1; } while (0)
extern "C" int ULCB_DebugDecodeWsPhaseTimes(unsigned long long *out32) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(out32, g_ulcbWsPhase, sizeof(unsigned long long) * 32);
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_ulcbWsPhase, z, sizeof z);
    return (int)e;
}
#else
#define WTK(i) do { } while (0)
#define WCNT(i) do { } while (0)
#endif

namespace {

constexpr int N = 2048, H = N / 2;
constexpr int NTHR = 96;
constexpr int MAXRUNS = N / 16 + 8;                 // per channel: a noise run spans >= 16 coefficients; + one tail per sub-block
constexpr int ZROW = 33;                              // padded row of the 32 x 32 complex work buffer
constexpr unsigned FULLM = 0xffffffffu;
enum { BAR_FULL = 1, BAR_EMPTY = 2, BAR_XCHG = 3 };

__device__ __forceinline__ void bar_sync(int id, int n)   { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { __threadfence_block(); asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }

// Decimation layouts, one nybble per sub-block, LSB first: bits0-2 shift, bit3 transient
// (libulc/ulcHelper.h:24-46, FormatSpecs.md:35-51)
__constant__ uint16_t kLayoutWs[16] = {
    0x0000, 0x0008, 0x0019, 0x0091, 0x012A, 0x01A2, 0x02A1, 0x0A21,
    0x123B, 0x12B3, 0x13B2, 0x1B32, 0x23B1, 0x2B31, 0x3B21, 0xB321,
};

struct Run  { uint32_t a, b; };                     // a: start | n << 12; b: draw offset in block | bf16 amplitude << 16
struct Tail { int start, end; float amp, l2; };     // decaying noise to the end of a sub-block (ulcDecoder.c:172-186); l2 = decay factor r
struct Meta { int live, wc, nRuns[2], nTails[2]; uint32_t drawBase; int pad; Tail tail[2][8]; };

struct Smem {
    uint16_t coef[2][N];          // bf16 dequantised coefficients, per channel
    float    pend[2][H];          // pending tail (k-layout or time order, see above)
    float2   zbuf[2][32 * ZROW];  // per transform warp: FFT work buffer / short-block work buffer / output swap
    Run      runs[2][2][MAXRUNS];  // [block parity][channel]: the transform warps read a block's records after releasing its coefficients
    Meta     meta[2];
    uint32_t wfifo[128];          // parse warp: the next 128 words of the stream (cp.async, 32-word groups), circular
};

__device__ __forceinline__ float quant_scale(int q) {      // ulcDecoder.c:96-98
    return 0x1.0p-31f * (float)((1u << 26) >> (q & 31));
}

// ================================================================================================
// PARSE warp
// ================================================================================================
// bit 3 of every nybble of x that equals c
__device__ __forceinline__ uint32_t nyb_eq(uint32_t x, uint32_t c) {
    uint32_t t = x ^ (c * 0x11111111u);
    return ~(((t & 0x77777777u) + 0x77777777u) | t) & 0x88888888u;
}

struct Win {                       // one lane's 8 nybbles of the current 256-nybble window, pre-digested
    uint32_t w, nx;                // own word, next word (tokens spill up to 4 nybbles)
    uint32_t L;                    // token length if a token starts at nybble o (nybble-packed)
    uint32_t exl, exh;             // chain exit (0..3 nybbles into the next word, 4 = stop code met) for start o, one per
                                   // byte; exl (starts 0..3) is the word's transition map
    uint32_t is0, is1, is8, isF, stop;
};

// nybbles n0..n3 of the low half of x -> one per byte
__device__ __forceinline__ uint32_t spread4(uint32_t x) {
    const uint32_t y = __byte_perm(x, 0u, 0x1100);
    return (y & 0x000f000fu) | ((y >> 4) & 0x0f000f00u);
}
// bytes (< 8 each) -> PRMT selector nybbles
__device__ __forceinline__ uint32_t pack_sel(uint32_t t) {
    const uint32_t s = t | (t >> 4);
    return __byte_perm(s, 0u, 0x4420);
}

// Token lengths, stop codes and the word's transition map from its raw nybbles.
__device__ __forceinline__ void win_digest(Win &W) {
    const uint32_t x = W.w, x1 = __funnelshift_r(W.w, W.nx, 4), x2 = __funnelshift_r(W.w, W.nx, 8);
    W.is0 = nyb_eq(x, 0); W.is1 = nyb_eq(x, 1); W.is8 = nyb_eq(x, 8); W.isF = nyb_eq(x, 15);
    const uint32_t e1 = nyb_eq(x1, 14), f1 = nyb_eq(x1, 15);
    W.stop = W.isF & ((e1 & nyb_eq(x2, 15)) | f1);           // Fh,Eh,Fh / Fh,Fh,...
    W.L = 0x11111111u + (W.is0 >> 3) + ((W.is1 >> 3) * 2u) + ((W.is8 >> 3) * 3u) + (W.isF >> 3) +
          ((W.isF & (e1 | f1)) >> 3) + ((W.isF & f1) >> 2);
    // T[o] = where the chain that starts at nybble o stands after 1, 2, 4, 8 tokens: o + len, kept once it has
    // left the word (8 + spill, spill <= 3) or met a stop code (12).  Gather = PRMT over the 8 byte table (Tl, Th).
    const uint32_t sl = spread4(W.stop >> 3) * 255u, sh = spread4(W.stop >> 19) * 255u;
    uint32_t Tl = ((spread4(W.L) + 0x03020100u) & ~sl) | (0x0c0c0c0cu & sl);
    uint32_t Th = ((spread4(W.L >> 16) + 0x07060504u) & ~sh) | (0x0c0c0c0cu & sh);
#pragma unroll
    for (int it = 0; it < 3; it++) {
        const uint32_t gl = __byte_perm(Tl, Th, pack_sel(Tl & 0x07070707u)), gh = __byte_perm(Tl, Th, pack_sel(Th & 0x07070707u));
        const uint32_t kl = ((Tl >> 3) & 0x01010101u) * 255u, kh = ((Th >> 3) & 0x01010101u) * 255u;
        Tl = (Tl & kl) | (gl & ~kl);
        Th = (Th & kh) | (gh & ~kh);
    }
    W.exl = Tl - 0x08080808u; W.exh = Th - 0x08080808u;     // exit state per start offset: 0..3 = spill, 4 = stopped
}

struct StreamSrc { const uint32_t *abase; long long nWordsValid, endNyb; };

struct Parser {
    Smem &sm;
    const StreamSrc &src;
    const int lane;
    Win W;
    long long curW = -1;           // word index of the window start (any word: a chain's window starts at its first word)
    long long reqUpTo = -100;      // last 32-word group requested
    bool digested = false;
    bool emptyPending = false;     // the EMPTY hand-off of the previous block has not been waited for yet
#ifdef ULCB_PHASE_TIMING
    long long cacc[10] = {0}, ct = 0;
#define CTK(i) do { long long t_ = clock64(); cacc[i] += t_ - ct; ct = t_; } while (0)
#else
#define CTK(i) do { } while (0)
#endif

    __device__ Parser(Smem &s, const StreamSrc &r, int l) : sm(s), src(r), lane(l) {
        asm volatile("cp.async.wait_all;" ::: "memory");               // the previous stream's look-ahead windows
        __syncwarp();
    }

    // The stream is copied global -> shared memory with cp.async in aligned 32-word groups into a 128-word circular
    // buffer, up to three groups ahead of the parse (a stream is read front to back), so the HBM/L2 latency is hidden
    // and costs no registers.  A window is any 33 consecutive words of it.
    __device__ __forceinline__ void request(long long g32) {              // words [32 g32, 32 g32 + 32)
        const long long g = g32 * 32 + lane;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.wfifo[g & 127]);
        const bool ok = g < src.nWordsValid;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src.abase + (ok ? g : 0)), "r"(ok ? 4 : 0) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // Open the window that starts at word w0.
    __device__ __forceinline__ void open(long long w0) {
        const long long g0 = w0 >> 5;                                      // the window spans groups g0 and g0 + 1
        if (g0 > reqUpTo + 1 || g0 + 3 < reqUpTo) reqUpTo = g0 - 1;         // not a step forward: start over from here
        while (reqUpTo < g0 + 3) request(++reqUpTo);                        // g0 + 4 would overwrite g0
        const int ahead = (int)(reqUpTo - (g0 + 1));                        // groups issued after the window's last
        if (ahead >= 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (ahead == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        W.w = sm.wfifo[(w0 + lane) & 127]; W.nx = sm.wfifo[(w0 + lane + 1) & 127];
        __syncwarp();
        curW = w0; digested = false;
    }
    // A chain's window starts at the word that holds its first token, so no lane of the window is wasted on what came
    // before it; a window that merely has to contain `pos` (header peek) is kept if it already does.
    __device__ __forceinline__ void ensure_start(long long pos) { if ((pos >> 3) != curW) open(pos >> 3); }
    __device__ __forceinline__ void ensure(long long pos) {
        const long long w = pos >> 3;
        if (curW < 0 || w < curW || w >= curW + 32) open(w);
    }
    // the 4 nybbles that start at pos, low nybble first (a window's last lane also holds the next window's first word)
    __device__ __forceinline__ uint32_t peek16(long long pos) {
        ensure(pos);
        const int rel = (int)(pos - curW * 8);
        const uint32_t lo = __shfl_sync(FULL, W.w, rel >> 3), hi = __shfl_sync(FULL, W.nx, rel >> 3);
        return __funnelshift_r(lo, hi, (rel & 7) * 4) & 0xffffu;
    }
    static constexpr unsigned FULL = 0xffffffffu;

    // Resolve and emit the token chain that starts at window nybble tp: `remain` coefficients still to come in this
    // sub-block, quantiser qcur, next coefficient index idx0 (channel-relative), next noise draw blkDraws (block-relative).
    // Returns through the reference arguments; result: 0 = chain continues in the next window, 1 = sub-block complete,
    // -1 = corrupt.
    struct ChainOut { int endPos, kind, coefs, draws, q, zyx; };

    __device__ int chain(int tp, int remain, int qcur, int idx0, int blkDraws, int c, Run *runsOut, int &nRuns, ChainOut &out) {
        const int s0 = tp >> 3, e0 = tp & 7;
        CTK(9);
        if (!digested) { win_digest(W); digested = true; }
        CTK(0);
        // ---- entry offsets: inclusive scan of transition maps ----
        uint32_t m = lane < s0 ? 0x03020100u : (lane == s0 ? ((((e0 & 4) ? W.exh : W.exl) >> (8 * (e0 & 3))) & 0xffu) * 0x01010101u : W.exl);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, m, d);
            const uint32_t mm = __byte_perm(m, 0x04040404u, pack_sel(t));       // new[e] = own[t[e]]; state 4 (stopped) absorbs
            if (lane >= d) m = mm;
        }
        CTK(1);
        const uint32_t mex = __shfl_up_sync(FULL, m, 1);
        const int entry = (lane == s0) ? e0 : (int)(mex & 0xffu);
        const bool on = lane >= s0 && (lane == s0 || entry != 4);

        // ---- walk my word from the entry offset: token-start mask ----
        uint32_t starts = 0;
        bool hasStop = false;
        if (on) {
            int p = entry;
#pragma unroll 2
            for (int it = 0; it < 8; it++) {
                if (p < 8 && !hasStop) {
                    const uint32_t bit = 8u << (4 * p);
                    starts |= bit;
                    if (W.stop & bit) hasStop = true;
                    else p += (W.L >> (4 * p)) & 7;
                }
            }
        }
        CTK(2);
        // ---- coefficient / noise-draw counts and the last quantiser my tokens set ----
        int cs = __popc(starts & ~(W.is0 | W.is1 | W.is8 | W.isF)), ds = 0;
        for (uint32_t sp = starts & (W.is0 | W.is1 | W.is8); sp; sp &= sp - 1) {
            const int o4 = (__ffs(sp) - 1) & ~3;
            const uint32_t x = __funnelshift_r(W.w, W.nx, o4);
            const int v0 = x & 15, v1 = (x >> 4) & 15, b12 = (v1 << 4) | ((x >> 8) & 15);
            const int n = v0 == 0 ? v1 + 1 : v0 == 1 ? b12 + 33 : ((b12 << 1) | ((x >> 12) & 1)) + 16;
            cs += n;
            if (v0 == 8) ds += n;
        }
        int lq = 0;
        {
            const uint32_t qm = starts & W.isF & ~W.stop;
            if (qm) {
                const int o4 = (31 - __clz(qm)) & ~3;
                const uint32_t x = __funnelshift_r(W.w, W.nx, o4);
                const int v1 = (x >> 4) & 15;
                lq = 32 | (v1 < 14 ? v1 : 14 + ((x >> 8) & 15));
            }
        }
        CTK(3);
        uint32_t inc = (uint32_t)cs | ((uint32_t)ds << 16);
        int il = lq;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, inc, d);
            const int tl = __shfl_up_sync(FULL, il, d);
            if (lane >= d) { inc += t; if (!il) il = tl; }
        }
        const int ic = inc & 0xffff, ec = ic - cs, ed = (int)(inc >> 16) - ds;
        int el = __shfl_up_sync(FULL, il, 1);
        if (lane == 0) el = 0;
        const int qIn = el ? (el & 31) : qcur;

        CTK(4);
        // ---- end word: the first on-chain word that holds a stop code or completes the sub-block ----
        const unsigned endBal = __ballot_sync(FULL, on && (hasStop || ic >= remain));
        const int endLane = endBal ? __ffs(endBal) - 1 : 32;
        int endP = 8, kind = -1, zyx = 0, ok = 1, cEnd = 0, dEnd = 0, qEnd = qIn;
        if (lane == endLane) {                                     // its owner walks it token by token
            int cc = ec, dd = ed, p = entry, q = qIn;
            for (;;) {
                const uint32_t x = __funnelshift_r(W.w, W.nx, 4 * (p & 7)) ;
                // tokens of the end word start at p < 8 by construction
                const int v0 = x & 15, v1 = (x >> 4) & 15, v2 = (x >> 8) & 15, b12 = (v1 << 4) | v2;
                int len = 1, n = 1, dr = 0;
                if (v0 == 0) { len = 2; n = v1 + 1; }
                else if (v0 == 1) { len = 3; n = b12 + 33; }
                else if (v0 == 8) { len = 4; n = ((b12 << 1) | ((x >> 12) & 1)) + 16; dr = n; }
                else if (v0 == 15) {
                    n = 0;
                    if (v1 < 14) { len = 2; q = v1; }
                    else if (v1 == 14) { len = 3; if (v2 == 15) { kind = 1; p += 3; break; } q = 14 + v2; }
                    else { kind = 2; zyx = (x >> 8) & 0xfff; p += 5; break; }
                }
                if (cc + n > remain) { ok = 0; break; }            // run overshoot: corrupt (ulcDecoder.c:127,139,154)
                cc += n; dd += dr; p += len;
                if (cc == remain) { kind = 0; break; }
                if (p >= 8) { ok = 0; break; }                     // cannot happen for a well-formed end word
            }
            endP = p; cEnd = cc; dEnd = dd; qEnd = q;
        }
        if (endLane < 32) {
            endP = __shfl_sync(FULL, endP, endLane); kind = __shfl_sync(FULL, kind, endLane);
            zyx = __shfl_sync(FULL, zyx, endLane);   ok = __shfl_sync(FULL, ok, endLane);
            cEnd = __shfl_sync(FULL, cEnd, endLane); dEnd = __shfl_sync(FULL, dEnd, endLane);
            qEnd = __shfl_sync(FULL, qEnd, endLane);
            if (!ok) return -1;
        }

        CTK(5);
        // ---- emit ----
        if (emptyPending) { bar_sync(BAR_EMPTY, NTHR); emptyPending = false; }   // first write of this block: the transform warps must have loaded the last one
        const bool emit = on && lane <= endLane;
        const int lim = (lane == endLane) ? endP : 8;              // tokens starting before lim are mine
        const uint32_t below = lim >= 8 ? 0xffffffffu : ((1u << (4 * lim)) - 1u);
        const int nNoise = emit ? __popc(starts & W.is8 & below) : 0;
        const unsigned b1 = __ballot_sync(FULL, nNoise >= 1), b2 = __ballot_sync(FULL, nNoise >= 2);
        const unsigned lt = (1u << lane) - 1u;
        int slot = nRuns + __popc(b1 & lt) + __popc(b2 & lt);
        nRuns += __popc(b1) + __popc(b2);
        if (emit) {
            // Straight-line, predicated: one pass over the 8 nybble positions.  Values are assembled as bf16 bit patterns
            // with integer arithmetic (exact small integer via the 2^23 trick, then the quantiser subtracted from the
            // exponent field): coefficient = sign * v^2 * 2^-(5+q), noise amplitude = lv^2 * 2^-(7+q), zero when q > 26
            // (ulcDecoder.c:96-98).
            int idx = idx0 + ec, draw = blkDraws + ed;
            int eb = qIn > 26 ? -1 : (5 + qIn) << 7;                               // bf16 exponent bias of the quantiser
            uint16_t *cf = sm.coef[c];
            const uint32_t mine = starts & below, special = W.is0 | W.is1 | W.is8 | W.isF, qtok = W.isF & ~W.stop;
#pragma unroll 2
            for (int o = 0; o < 8; o++) {
                const uint32_t bit = 8u << (4 * o);
                const bool tok = (mine & bit) != 0;
                const uint32_t x = __funnelshift_r(W.w, W.nx, 4 * o);
                const int v0 = x & 15, v1 = (x >> 4) & 15, v2 = (x >> 8) & 15, b12 = (v1 << 4) | v2;
                const bool s8 = (W.is8 & bit) != 0;
                const int sv = (v0 ^ 8) - 8;
                const int m2 = s8 ? (int)(((x >> 13) & 7) + 1) * (int)(((x >> 13) & 7) + 1) : sv * sv;   // lv^2 or v^2, 1..64
                const uint32_t fb = __float_as_uint(__uint_as_float(0x4B000000u | (uint32_t)m2) - 8388608.0f) >> 16;
                const uint32_t val = eb < 0 ? 0u : (fb - (uint32_t)eb - (s8 ? 0x100u : 0u)) | ((uint32_t)(v0 & 8 & (s8 ? 0 : 8)) << 12);
                int n = (special & bit) ? 0 : 1;
                n = (W.is0 & bit) ? v1 + 1 : n;
                n = (W.is1 & bit) ? b12 + 33 : n;
                n = s8 ? ((b12 << 1) | (int)((x >> 12) & 1)) + 16 : n;
                if (tok && !(special & bit) && idx < N) cf[idx] = (uint16_t)val;   // coefficient
                if (tok && s8 && slot < MAXRUNS && idx + n <= N) {                 // noise run: one record, filled by the transform warp
                    Run r; r.a = (uint32_t)idx | ((uint32_t)n << 12); r.b = (uint32_t)draw | (val << 16);
                    runsOut[slot] = r;
                }
                const bool nz = tok && s8;
                slot += nz ? 1 : 0; draw += nz ? n : 0;
                const int nq = v1 < 14 ? v1 : 14 + v2;
                eb = (tok && (qtok & bit)) ? (nq > 26 ? -1 : (5 + nq) << 7) : eb;
                idx += tok ? n : 0;
            }
        }
        CTK(6);
        if (endLane < 32) {
            out.endPos = endLane * 8 + endP; out.kind = kind; out.coefs = cEnd; out.draws = dEnd; out.q = qEnd; out.zyx = zyx;
            return 1;
        }
        // chain runs off this window
        out.endPos = 256 + (int)(__shfl_sync(FULL, m, 31) & 0xffu);
        out.coefs = __shfl_sync(FULL, ic, 31); out.draws = (int)(__shfl_sync(FULL, inc, 31) >> 16);
        const int l31 = __shfl_sync(FULL, il, 31);
        out.q = l31 ? (l31 & 31) : qcur; out.kind = -1; out.zyx = 0;
        return 0;
    }
};

__device__ void parse_warp(Smem &sm, const UlcbDecodeArgs &args, int stream, const StreamSrc &src, long long pos0, int nBlocks, int lane, unsigned bufBase) {
    Parser P(sm, src, lane);
    long long pos = pos0;
    uint32_t drawBase = 0;
    int deadAt = -1;
#ifdef ULCB_PHASE_TIMING
    long long acc[8] = {0}, tPrev = clock64();
#endif
    for (int blk = 0; blk < nBlocks; blk++) {
        WTK(0);
        P.emptyPending = blk > 0;                                        // waited for at the block's first write (Parser::chain) or its end
        WTK(1);
        // runs / meta are double-buffered by a CTA-lifetime block counter: the transform warps may still read the previous
        // block's records (of this or of the previous stream) when the parse of the next block starts
        const unsigned par = (bufBase + (unsigned)blk) & 1u;
        Meta &mt = sm.meta[par];
        int nRunsC[2] = {0, 0}, nTailsC[2] = {0, 0};
        if (deadAt < 0 && pos >= src.endNyb) deadAt = blk;             // stream exhausted
        bool live = deadAt < 0;
        int wc = 0x10;
        uint32_t rawBits = 0, blkDraws = 0;
        if (live) {
            if (blk == 0) {   // clear both channels' coefficients once per stream; afterwards the transform warps clear what they load
                uint4 *z = reinterpret_cast<uint4 *>(&sm.coef[0][0]);
#pragma unroll
                for (int i = 0; i < (int)(sizeof(sm.coef) / 16 / 32); i++) z[lane + 32 * i] = make_uint4(0, 0, 0, 0);
                __syncwarp();
            }
            const long long blkStart = pos;
            uint32_t hdr = P.peek16(pos);                               // block header (1-2 nybbles) + the first quantiser (1-2)
            int hdrLeft = 4;
            const int h0 = hdr & 15;
            if (h0 & 8) { wc = h0 | (int)(((hdr >> 4) & 15) << 4); pos += 2; hdr >>= 8; hdrLeft = 2; }
            else { wc = h0 | 0x10; pos += 1; hdr >>= 4; hdrLeft = 3; }
            const unsigned layout = kLayoutWs[wc >> 4];
            bool ok = true;
            for (int c = 0; c < 2 && ok; c++) {
                int coefBase = 0, nRuns = 0, nTails = 0;
                for (unsigned pat = layout;; ) {
                    const int S = N >> (pat & 7);
                    if (hdrLeft < 2) { hdr = P.peek16(pos); hdrLeft = 4; }
                    int q = hdr & 15;                                  // sub-block header: bare quantiser (ulcDecoder.c:103-109)
                    int used = 1;
                    bool silent = false;
                    if (q == 15) q = 30;                               // invalid lead-in; x86 shift semantics -> scale 0
                    else if (q == 14) { q += (hdr >> 4) & 15; used = 2; if (q == 29) silent = true; }
                    pos += used; hdr >>= 4 * used; hdrLeft -= used;
                    if (!silent) hdrLeft = 0;                          // the token chains move pos: peek again afterwards
                    if (!silent) {
                        int remain = S;
                        for (;;) {
                            WTK(0);
                            P.ensure_start(pos);
                            WTK(2);
                            const int tp = (int)(pos - P.curW * 8);
                            Parser::ChainOut co;
                            WCNT(4);
                            const int r = P.chain(tp, remain, q, coefBase + (S - remain), (int)blkDraws, c, sm.runs[par][c], nRuns, co);
                            WTK(3);
                            if (r < 0) { ok = false; break; }
                            pos = P.curW * 8 + co.endPos;
                            blkDraws += (uint32_t)co.draws;
                            if (r == 1) {
                                if (co.kind == 2) {                    // Fh,Fh,Z,Y,X: decaying noise to the end (ulcDecoder.c:172-186)
                                    const int left = remain - co.coefs;
                                    if (left > 0) {
                                        const int lv = (co.zyx & 15) + 1;
                                        const int nn = (((co.zyx >> 4) & 15) << 4) | ((co.zyx >> 8) & 15);
                                        const int start = coefBase + (S - remain) + co.coefs;
                                        if (lane == 0 && nRuns < MAXRUNS) {
                                            Run rr; rr.a = (uint32_t)start | ((uint32_t)left << 12);
                                            rr.b = blkDraws | 0x3f800000u;          // unit amplitude; the tail's level is applied at load
                                            sm.runs[par][c][nRuns] = rr;
                                            Tail t; t.start = start; t.end = start + left;
                                            t.amp = (float)(lv * lv) * quant_scale(co.q) * (1.0f / 16);
                                            const float decay = 1.0f + (float)(nn * nn) * -0x1.0p-19f;
                                            t.l2 = decay;
                                            mt.tail[c][nTails] = t;
                                        }
                                        nRuns++; nTails++;
                                        blkDraws += (uint32_t)left;
                                    }
                                }
                                break;
                            }
                            remain -= co.coefs; q = co.q;
                            if (pos >= src.endNyb + 16) { ok = false; break; }   // ran off the stream
                        }
                    }
                    coefBase += S;
                    pat >>= 4;
                    if (!pat || !ok) break;
                }
                nRunsC[c] = min(nRuns, MAXRUNS); nTailsC[c] = nTails;      // stored with the rest of the meta, after the EMPTY wait
            }
            // a block whose bytes run past the end of the stream (it parsed from the zero padding) is not decodable; neither
            // is one whose noise draws lie beyond the sign table the context was created for (maxBlocksPerStream)
            if (ok && (pos > src.endNyb || (unsigned long long)drawBase + blkDraws + 128ull > args.tb.nSignBits)) ok = false;
            if (!ok) { live = false; deadAt = blk; }
            else {
                rawBits = (uint32_t)(pos - blkStart) * 4u;               // what ULC_DecodeBlock returns (ulcDecoder.c:301)
                pos = (pos + 1) & ~1ll;                                   // blocks are byte aligned (ulcEncoder_Encode.c:357-358)
            }
        }
        if (P.emptyPending) { bar_sync(BAR_EMPTY, NTHR); P.emptyPending = false; }   // block without a single token chain
        if (lane == 0) {
            mt.live = live; mt.wc = wc; mt.drawBase = drawBase;
            mt.nRuns[0] = nRunsC[0]; mt.nRuns[1] = nRunsC[1]; mt.nTails[0] = nTailsC[0]; mt.nTails[1] = nTailsC[1];
            if (args.blockBits) args.blockBits[(size_t)stream * args.bitsStride + blk] = live ? rawBits : 0u;
        }
        if (live) drawBase += blkDraws;
        __syncwarp();
        bar_arrive(BAR_FULL, NTHR);
    }
    if (nBlocks > 0) bar_sync(BAR_EMPTY, NTHR);
#ifdef ULCB_PHASE_TIMING
    if (lane == 0) { for (int i = 0; i < 8; i++) atomicAdd(&g_ulcbWsPhase[i], (unsigned long long)acc[i]); atomicAdd(&g_ulcbWsPhase[8], (unsigned long long)nBlocks);
        for (int i = 0; i < 7; i++) atomicAdd(&g_ulcbWsPhase[9 + i], (unsigned long long)P.cacc[i]); }
#endif
    if (lane == 0) {
        if (args.status) args.status[stream] = deadAt;                   // -1 = every block decoded
        if (args.bytesConsumed) args.bytesConsumed[stream] = (uint64_t)((pos - pos0) >> 1);
    }
}

// ================================================================================================
// TRANSFORM warps
// ================================================================================================
__host__ __device__ constexpr int brev5(int x) {
    return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}
__device__ __forceinline__ constexpr float cos32(int k) {
    constexpr float t[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f,
                             0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f, 0.0f, -0.19509032201612825f,
                             -0.38268343236508978f, -0.55557023301960218f, -0.70710678118654757f, -0.83146961230254524f,
                             -0.92387953251128674f, -0.98078528040323043f};
    return t[k];
}
__device__ __forceinline__ constexpr float sin32(int k) { return k < 8 ? cos32(8 - k) : -cos32(24 - k); }

// 32-point forward FFT (e^{-2 pi i nk/32}) on registers, radix-2 decimation in frequency:
// natural order in, X[brev5(r)] left in register r.
__device__ __forceinline__ void fft32(float (&xr)[32], float (&xi)[32]) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int half = 16 >> s;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            if ((i & half) == 0) {
                const int j = i + half, k = (i & (half - 1)) << s;
                const float ar = xr[i], ai = xi[i], br = xr[j], bi = xi[j];
                xr[i] = ar + br; xi[i] = ai + bi;
                const float dr = ar - br, di = ai - bi;
                if (k == 0)      { xr[j] = dr; xi[j] = di; }
                else if (k == 8) { xr[j] = di; xi[j] = -dr; }
                else {
                    const float c = cos32(k), sn = sin32(k);
                    xr[j] = dr * c + di * sn;
                    xi[j] = di * c - dr * sn;
                }
            }
        }
    }
}

// time index held by k-layout slot k
__device__ __forceinline__ int slot_time(int k) { return k >= N / 4 ? 2 * k - H : H - 1 - 2 * k; }

template <bool PCM16>
__device__ __forceinline__ void store_pairs(void *outBlk, int slotIdx, float l0, float r0, float l1, float r1);
template <>
__device__ __forceinline__ void store_pairs<false>(void *outBlk, int slotIdx, float l0, float r0, float l1, float r1) {
    __stcs(reinterpret_cast<float4 *>(outBlk) + slotIdx, make_float4(l0, r0, l1, r1));
}
__device__ __forceinline__ uint32_t pcm16_pair(float l, float r) {   // tools/WavIO_Helper.c:57-63: clamp, lrintf
    l = fminf(fmaxf(l * 32768.0f, -32768.0f), 32767.0f);
    r = fminf(fmaxf(r * 32768.0f, -32768.0f), 32767.0f);
    return (uint32_t)(uint16_t)(int16_t)__float2int_rn(l) | ((uint32_t)(uint16_t)(int16_t)__float2int_rn(r) << 16);
}
template <>
__device__ __forceinline__ void store_pairs<true>(void *outBlk, int slotIdx, float l0, float r0, float l1, float r1) {
    __stcs(reinterpret_cast<uint2 *>(outBlk) + slotIdx, make_uint2(pcm16_pair(l0, r0), pcm16_pair(l1, r1)));
}

// noise fill of channel c from the run records (ulcDecoder.c:156-160,179-184): sign of draw g+i of a run that
// starts at draw g is parity[g+i+1] ^ parity[g].  One lane per run, 64 coefficients per round of table reads.
// STASH = true: the coefficient words have already been moved into the FFT tile (word n in the .x of slot [n / 32][n % 32]).
template <bool STASH>
__device__ void noise_fill(Smem &sm, const Meta &mt, const Run *runs, int c, const uint32_t *__restrict__ par, int lane) {
    const int nR = mt.nRuns[c];
    auto word = [&](int n) -> uint32_t * {
        return STASH ? reinterpret_cast<uint32_t *>(&sm.zbuf[c][(n >> 5) * ZROW + (n & 31)].x)
                     : reinterpret_cast<uint32_t *>(sm.coef[c]) + n;
    };
    auto half = [&](int i) -> uint16_t * { return reinterpret_cast<uint16_t *>(word(i >> 1)) + (i & 1); };
    for (int r0 = 0; r0 < nR; r0 += 32) {
        if (r0 + lane >= nR) continue;
        const Run rr = runs[r0 + lane];
        int i = rr.a & 0xfff;
        const int end = i + (int)(rr.a >> 12);
        const uint32_t g0 = mt.drawBase + (rr.b & 0xffffu);
        const uint32_t amp = rr.b >> 16, amp2 = amp | (amp << 16);
        const uint32_t flip = 0u - ((__ldg(&par[g0 >> 5]) >> (g0 & 31)) & 1u);
        for (uint32_t g = g0; i < end; g += 64) {
            const uint32_t wi = g >> 5, sh = (g & 31) + 1;                        // bits g+1 .. g+64
            const uint32_t a0 = __ldg(&par[wi]), a1 = __ldg(&par[wi + 1]), a2 = __ldg(&par[wi + 2]);
            uint32_t lo = (sh < 32 ? __funnelshift_r(a0, a1, sh) : a1) ^ flip;
            uint32_t hi = (sh < 32 ? __funnelshift_r(a1, a2, sh) : a2) ^ flip;
            int left = min(end - i, 64);
            if (i & 1) {                                                            // unaligned head
                *half(i) = (uint16_t)(amp | ((lo & 1u) << 15));
                lo = __funnelshift_r(lo, hi, 1); hi >>= 1; i++; left--;
            }
            for (; left >= 2; left -= 2, i += 2) {
                *word(i >> 1) = amp2 | ((lo & 1u) << 15) | ((lo & 2u) << 30);
                lo = __funnelshift_r(lo, hi, 2); hi >>= 2;
            }
            if (left) { *half(i) = (uint16_t)(amp | ((lo & 1u) << 15)); i++; }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ float tail_gain(const Meta &mt, int c, int nT, int idx) {
    float g = 1.0f;
    for (int t = 0; t < nT; t++) {
        const Tail &tl = mt.tail[c][t];
        if (idx >= tl.start && idx < tl.end) {
            // the reference's running product p *= r, step by step (ulcDecoder.c:178-184): in f32 it drifts up to 1e-4 away from
            // amp * r^k for slow decays, so the closed form is not an option; tails are rare and this kernel is not the default
            // path, so each value simply re-runs the chain from the start of the tail
            g = tl.amp;
            const float r = tl.l2;                           // the decay factor itself
            for (int k = idx - tl.start; k > 0; k--) g = __fmul_rn(g, r);
        }
    }
    return g;
}

template <bool PCM16>
__device__ void transform_warp(Smem &sm, const UlcbDecodeArgs &args, int stream, int nBlocks, int c, int lane, unsigned bufBase) {
    const UlcbTables &tb = args.tb;
    float *pend = sm.pend[c];
    float2 *zb = sm.zbuf[c];
    float *scr = reinterpret_cast<float *>(sm.zbuf[c]);
    const float2 *__restrict__ E = tb.dctE[0];
    const float2 *__restrict__ Wt = tb.fftW;
    const float2 *__restrict__ win = tb.win;

#pragma unroll
    for (int j = 0; j < 32; j++) pend[lane + 32 * j] = 0.0f;
    bool kLayout = true;                               // zeros read the same either way
    int lastSub = 0;
    __syncwarp();

#ifdef ULCB_PHASE_TIMING
    long long acc[12] = {0}, tPrev = clock64();
#endif
    for (int blk = 0; blk < nBlocks; blk++) {
        WTK(0);
        bar_sync(BAR_FULL, NTHR);
        WTK(1);
        const unsigned par = (bufBase + (unsigned)blk) & 1u;
        const Meta &mt = sm.meta[par];
        const int live = mt.live, wc = mt.wc, nT = mt.nTails[c];
        char *outBlk = reinterpret_cast<char *>(args.out) +
            ((size_t)stream * args.outStride + (size_t)blk * N * 2) * (PCM16 ? 2 : 4);
        if (!live) {
            // corrupt / exhausted stream: silence (the reference tool aborts here, ulcDecodeTool.c:153-157)
            bar_arrive(BAR_EMPTY, NTHR);
#pragma unroll
            for (int j = 0; j < 16; j++) store_pairs<PCM16>(outBlk, c * 512 + lane + 32 * j, 0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const Run *runs = sm.runs[par][c];
        WTK(2);
        float *dbg = args.dbgCoefs ? args.dbgCoefs + (((size_t)stream * args.nBlocksAll + blk) * 2 + c) * N : nullptr;

        if ((wc >> 4) <= 1) {                             // layouts 1 (N/1*) and 0 (N/1, "unused" but parseable, ulcHelper.h:29)
            // ======================= long block: register path =======================
            if (!kLayout) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = pend[slot_time(lane + 32 * j)];
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j++) pend[lane + 32 * j] = v[j];
                __syncwarp();
                kLayout = true;
            }
            {
                // z[n] = (v[2n] + i v[N-1-2n]) E[n], n = lane + 32 j, stored at zb[j][lane]; v[N-1-2n] is the odd half of
                // word 1023-n, which lane 31-lane holds at j' = 31-j
                // First only MOVE the raw words out of the hand-off buffer (into the .x of the slot their z will occupy: this
                // lane's own column) and clear it, so the parse warp gets its buffer back after ~100 instructions instead of
                // after the whole twiddle pass with its table loads.
                uint32_t *cw = reinterpret_cast<uint32_t *>(sm.coef[c]);
#pragma unroll 4
                for (int j = 0; j < 32; j++) {
                    const int n = lane + 32 * j;
                    const uint32_t wv = cw[n];
                    cw[n] = 0u;                                           // leave the buffer clear for the parse warp's next block
                    zb[j * ZROW + lane].x = __uint_as_float(wv);
                }
                bar_arrive(BAR_EMPTY, NTHR);                              // the coefficient buffer is free again
                __syncwarp();
                noise_fill<true>(sm, mt, runs, c, tb.signParity, lane);   // into the moved words; the records are double-buffered
#pragma unroll 1
                for (int j = 0; j < 16; j++) {
                    const int ja = j, jb = 31 - j;
                    const int na = lane + 32 * ja, nb = lane + 32 * jb;
                    const uint32_t wa = __float_as_uint(zb[ja * ZROW + lane].x), wb = __float_as_uint(zb[jb * ZROW + lane].x);
                    float ea = __uint_as_float(wa << 16), oa = __uint_as_float(wa & 0xffff0000u);
                    float eb = __uint_as_float(wb << 16), ob = __uint_as_float(wb & 0xffff0000u);
                    if (nT) {
                        ea *= tail_gain(mt, c, nT, 2 * na); oa *= tail_gain(mt, c, nT, 2 * na + 1);
                        eb *= tail_gain(mt, c, nT, 2 * nb); ob *= tail_gain(mt, c, nT, 2 * nb + 1);
                    }
                    if (dbg) {
                        reinterpret_cast<float2 *>(dbg)[na] = make_float2(ea, oa);
                        reinterpret_cast<float2 *>(dbg)[nb] = make_float2(eb, ob);
                    }
                    const float ia = __shfl_xor_sync(FULLM, ob, 31);      // v[N-1-2na]
                    const float ib = __shfl_xor_sync(FULLM, oa, 31);      // v[N-1-2nb]
                    const float2 e0 = __ldg(&E[na]), e1 = __ldg(&E[nb]);
                    zb[ja * ZROW + lane] = make_float2(ea * e0.x - ia * e0.y, ea * e0.y + ia * e0.x);
                    zb[jb * ZROW + lane] = make_float2(eb * e1.x - ib * e1.y, eb * e1.y + ib * e1.x);
                }
            }
            WTK(3);
            // 1024-point FFT as 32 x 32.  ONE copy of the pass code (instruction-cache footprint), run twice:
            //   pass 0: lane l transforms column l (n = l + 32 j over j), multiplies by W_1024^(l k1), stores A[l][k1] at zb[k1][l]
            //   pass 1: lane k1 transforms row k1 (over l), multiplies by the DCT-IV post-twiddle E[k], k = k1 + 32 k2, stores at zb[k2][k1]
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                float xr[32], xi[32];
                const float2 *srcp = zb + (pass ? lane * ZROW : lane);
                const int ldR = pass ? 1 : ZROW;
#pragma unroll
                for (int r = 0; r < 32; r++) { const float2 v = srcp[r * ldR]; xr[r] = v.x; xi[r] = v.y; }
                fft32(xr, xi);                            // register r now holds output index brev5(r)
                const float2 *__restrict__ tab = pass ? E : Wt;
                const int mulK = pass ? 32 : lane, addK = pass ? lane : 0;
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    const float2 w = __ldg(&tab[brev5(r) * mulK + addK]);
                    const float a = xr[r], b = xi[r];
                    xr[r] = a * w.x - b * w.y; xi[r] = a * w.y + b * w.x;
                }
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 32; r++) zb[brev5(r) * ZROW + lane] = make_float2(xr[r], xi[r]);
                __syncwarp();
            }
            WTK(6);
            // un-fold (Fourier_IMDCT_Template.h:66-93) against the pending tail, all at the same k = lane + 32 k2
            int ov = (wc >> 4) ? N >> (wc & 7) : N;       // layout 1: the long block is its own transient sub-block; layout 0: no overlap scaling
            if (ov > lastSub) ov = lastSub;
            lastSub = N;
            const int flat = (N - ov) >> 1;
            const int wofs = (ov >> 1) - flat;
            if (flat == 0) {
                // steady state (full overlap): sine window over the whole half block, win[N/2 + i]
                const float2 *__restrict__ wv = win + H;
#pragma unroll 2
                for (int k2 = 0; k2 < 16; k2++) {         // k < 512: i = N/2-1-2k, X[N/2+i] = -Im y, next tail = Re y
                    const int k = lane + 32 * k2;
                    const float2 y = zb[k2 * ZROW + lane];
                    const float2 w = __ldg(&wv[H - 1 - 2 * k]);
                    const float a = pend[k], b = -y.y;
                    pend[k] = y.x;
                    zb[k2 * ZROW + lane] = make_float2(w.x * a - w.y * b, w.y * a + w.x * b);
                }
#pragma unroll 2
                for (int k2 = 16; k2 < 32; k2++) {        // k >= 512: i = 2k-N/2, X[N/2+i] = Re y, next tail = -Im y
                    const int k = lane + 32 * k2;
                    const float2 y = zb[k2 * ZROW + lane];
                    const float2 w = __ldg(&wv[2 * k - H]);
                    const float a = pend[k], b = y.x;
                    pend[k] = -y.y;
                    zb[k2 * ZROW + lane] = make_float2(w.x * a - w.y * b, w.y * a + w.x * b);
                }
            } else {
#pragma unroll 2
                for (int k2 = 0; k2 < 32; k2++) {
                    const int k = lane + 32 * k2;
                    const float2 y = zb[k2 * ZROW + lane];
                    const bool up = k2 >= 16;
                    const int i = up ? 2 * k - H : H - 1 - 2 * k;
                    const float b  = up ? y.x : -y.y;         // X[N/2 + i]
                    const float np = up ? -y.y : y.x;         // X[N/2 - 1 - i]
                    const float a = pend[k];
                    pend[k] = np;
                    float lo, hi;
                    if (i < flat) { lo = a; hi = b; }
                    else { const float2 w = __ldg(&win[wofs + i]); lo = w.x * a - w.y * b; hi = w.y * a + w.x * b; }
                    zb[k2 * ZROW + lane] = make_float2(lo, hi);   // time samples i and N-1-i
                }
            }
            WTK(7);
            bar_sync(BAR_XCHG, 64);
            WTK(9);
            {
                // warp 0 stores the first half of the block (samples i = .x), warp 1 the second (samples N-1-i = .y); M/S undo
                // (ulcDecoder.c:281-289).  Pair (t, t+1) = slots k_e = p + (c ? 0 : 512) and k_o = (c ? 1023 : 511) - p, p = lane + 32 q.
                const float *m0 = reinterpret_cast<const float *>(sm.zbuf[0]) + c, *s0 = reinterpret_cast<const float *>(sm.zbuf[1]) + c;
#pragma unroll 2
                for (int q = 0; q < 16; q++) {
                    const int ie = ((c ? q : q + 16) * ZROW + lane) * 2, io = ((c ? 31 - q : 15 - q) * ZROW + 31 - lane) * 2;
                    const float me = m0[ie], se = s0[ie], mo = m0[io], so = s0[io];
                    store_pairs<PCM16>(outBlk, c * 512 + lane + 32 * q, me + se, me - se, mo + so, mo - so);
                }
            }
            bar_sync(BAR_XCHG, 64);
            WTK(10);
        } else {
            // ======================= window-switched block: generic shared-memory path, time order =======================
            if (kLayout) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = pend[lane + 32 * j];
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j++) pend[slot_time(lane + 32 * j)] = v[j];
                __syncwarp();
                kLayout = false;
            }
            noise_fill<false>(sm, mt, runs, c, tb.signParity, lane);
            int off = 0;
            for (unsigned pat = kLayoutWs[wc >> 4];;) {
                const int sh = pat & 7, S = N >> sh, hS = S >> 1;
                for (int i = lane; i < S; i += 32) {
                    float x = __uint_as_float((uint32_t)sm.coef[c][off + i] << 16);
                    sm.coef[c][off + i] = 0;
                    if (nT) x *= tail_gain(mt, c, nT, off + i);
                    if (dbg) dbg[off + i] = x;
                    scr[i] = x;
                }
                __syncwarp();
                if (!(pat >> 4)) bar_arrive(BAR_EMPTY, NTHR);          // last sub-block loaded
                dct4_inplace<8, true>(scr, S, 1, S, sh, tb);

                int ov = S;
                if (pat & 8) ov >>= (wc & 7);
                if (ov > lastSub) ov = lastSub;
                lastSub = S;
                const int flat = (S - ov) >> 1, held = H - hS;
                const float2 *__restrict__ wv = win + (ov >> 1);
                float uLo[16], uHi[16], tn[16];
#pragma unroll
                for (int r = 0; r < 16; r++) {
                    const int i = lane + 32 * r;
                    if (i < hS) {
                        const float a = pend[held + i], b = scr[hS + i];
                        if (i < flat) { uLo[r] = a; uHi[r] = b; }
                        else { const float2 w = __ldg(&wv[i - flat]); uLo[r] = w.x * a - w.y * b; uHi[r] = w.y * a + w.x * b; }
                        tn[r] = scr[hS - 1 - i];
                    }
                }
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 16; r++) {
                    const int i = lane + 32 * r;
                    if (i < hS) { scr[i] = uLo[r]; scr[S - 1 - i] = uHi[r]; pend[held + i] = tn[r]; }
                }
                __syncwarp();
                {   // delay line (ulcDecoder.c:248-272): [pend[0..held) ++ u[0..S)] rotated left by S, three reversals
                    const int tot = held + S;
                    auto at = [&](int p) -> float & { return (p < held) ? pend[p] : scr[p - held]; };
                    auto reverse = [&](int lo, int hi) {
                        const int np = (hi - lo) >> 1;
                        for (int k = lane; k < np; k += 32) { float &x = at(lo + k), &y = at(hi - 1 - k); const float t = x; x = y; y = t; }
                        __syncwarp();
                    };
                    reverse(0, S); reverse(S, tot); reverse(0, tot);
                }
                bar_sync(BAR_XCHG, 64);
                {
                    const float *m = reinterpret_cast<const float *>(sm.zbuf[0]), *s = reinterpret_cast<const float *>(sm.zbuf[1]);
                    for (int p = lane; p < (hS >> 1); p += 32) {
                        const int t = c * hS + 2 * p;
                        const float m0 = m[t], m1 = m[t + 1], s0 = s[t], s1 = s[t + 1];
                        store_pairs<PCM16>(outBlk, (off + t) >> 1, m0 + s0, m0 - s0, m1 + s1, m1 - s1);
                    }
                }
                bar_sync(BAR_XCHG, 64);
                off += S;
                pat >>= 4;
                if (!pat) break;
            }
        }
    }
#ifdef ULCB_PHASE_TIMING
    if (lane == 0 && c == 0) { for (int i = 0; i < 12; i++) atomicAdd(&g_ulcbWsPhase[16 + i], (unsigned long long)acc[i]); }
#endif
}

#ifndef ULCB_WS_MINB
#define ULCB_WS_MINB 5
#endif
template <bool PCM16>
__global__ void __launch_bounds__(NTHR, ULCB_WS_MINB)
ulcb_decode_ws_kernel(UlcbDecodeArgs args) {
    __shared__ __align__(16) Smem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned bufBase = 0;                                               // blocks this CTA has handled so far (buffer parity)
    for (int stream = blockIdx.x; stream < args.nStreams; stream += gridDim.x) {
        const int nBlocks = args.nBlocks ? (int)args.nBlocks[stream] : args.nBlocksAll;
        if (warp == 0) {
            const uint8_t *src = args.bytes + args.streamOffs[stream];
            const long long srcLen = (long long)(args.streamOffs[stream + 1] - args.streamOffs[stream]);
            const uintptr_t mis = reinterpret_cast<uintptr_t>(src) & 3;
            StreamSrc ss;
            ss.abase = reinterpret_cast<const uint32_t *>(src - mis);
            ss.nWordsValid = ((long long)mis + srcLen + 3) >> 2;
            ss.endNyb = 2 * ((long long)mis + srcLen);                  // aligned-space nybble index of the stream end
            parse_warp(sm, args, stream, ss, 2 * (long long)mis, nBlocks, lane, bufBase);
        } else {
            transform_warp<PCM16>(sm, args, stream, nBlocks, warp - 1, lane, bufBase);
        }
        bufBase += (unsigned)nBlocks;
    }
}

}  // namespace

// Fast path applicability: BlockSize 2048, stereo, fresh streams (no carried per-stream state).
bool ulcb_decode_ws_supported(const UlcbDecodeArgs &args, int nChan) {
    return args.tb.N == N && nChan == 2 && args.state == nullptr;
}

cudaError_t ulcb_decode_ws_launch(const UlcbDecodeArgs &args, int pcm16, int smCount, cudaStream_t stream) {
    void (*k)(UlcbDecodeArgs) = pcm16 ? ulcb_decode_ws_kernel<true> : ulcb_decode_ws_kernel<false>;
    // occupancy / carve-out are per device: cached per (device, variant), set up under a lock
    static std::mutex mu;
    static int perSmTab[64][2];
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    int &slot = perSmTab[dev & 63][pcm16 ? 1 : 0];
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!slot) {
            int carve = cudaSharedmemCarveoutMaxShared;
            if (const char *ev = getenv("ULCB_WS_CARVEOUT")) carve = atoi(ev);
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            if (e != cudaSuccess) return e;
            int n = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, NTHR, 0);
            if (e != cudaSuccess) return e;
            slot = n > 0 ? n : 1;
        }
    }
    // resident CTAs per SM: the most that fit, unless one or two fewer cut the last wave's idle tail by more than
    // the ~10 % of throughput each step down costs (measured: profiles/)
    const int rmax = slot;
    int best = rmax; double bestEff = 0.0;
    for (int r = rmax; r >= 1 && r >= rmax - 2; r--) {
        const long long slots = (long long)smCount * r;
        const long long waves = (args.nStreams + slots - 1) / slots;
        const double eff = (double)args.nStreams / (double)(waves * slots) * (1.0 - 0.10 * (rmax - r));
        if (eff > bestEff + 1e-9) { bestEff = eff; best = r; }
    }
    if (const char *e = getenv("ULCB_WS_PERSM")) { int v = atoi(e); if (v >= 1 && v <= rmax) best = v; }
    long long grid = (long long)smCount * best;
    if (grid > args.nStreams) grid = args.nStreams;
    k<<<(unsigned)grid, NTHR, 0, stream>>>(args);
    return cudaGetLastError();
}
