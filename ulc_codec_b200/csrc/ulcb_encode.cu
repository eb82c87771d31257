// ulcb_encode.cu - batched ulc-codec block encoder for sm_100a.
//
// One CTA owns one stream and walks it block by block with every intermediate in shared memory
// (global workspace for the largest block sizes).  Per block (reference: libulc/ulcEncoder.c:93-158):
//
//   ingest + M/S                 ulcEncoder_BlockTransform.c:87-111
//   transient detector           ulcEncoder_WindowControl.c:31-239      -> window control of the NEXT block
//   fold + window                ulcEncoder_BlockTransform.c:156-237, Fourier_MDCT_Template.h:45-187
//   DCT-IV x2 per (chan,sub)     Fourier_DCT4_Template.h / Fourier_DCT2_Template.h
//   normalise, keys, sums        ulcEncoder_BlockTransform.c:243-325
//   noise spectrum, masking      ulcEncoder_Psyopt.c:60-250
//   ranks                        ulcEncoder_BlockTransform.c:20-77,350-354
//   rate control                 ulcEncoder.c:93-158
//   nybble coder                 ulcEncoder_Encode.c:23-360, ulcEncoder_NoiseFill.c:15-94
//
// This is not a translation of the reference's control flow:
//  * The lapping FIFO + carried lap buffer are replaced by their closed form: every sub-block transform
//    reads a 2S-sample frame of the M/S sample timeline (two previous blocks are kept) and applies the
//    left/right sine windows directly.  The lap VALUES are the same single f32 products the reference
//    stores, so the result is bit-identical (tests/test_encode_gpu.py).
//  * The recursive DCT-IV/DCT-II pair is evaluated breadth-first: one pass per recursion level down,
//    8-point leaves, one pass per level up, all nodes of a level in parallel.  Same flow graph, same
//    rotor values (tabulated from the reference's recurrence), same operation order per node => bit-identical.
//  * Heapsort is replaced by an LSD radix sort of order-preserving key bits.
//  * The serial nybble coder becomes: selection -> zone walk (one thread per sub-block) -> per-slot token
//    strings -> prefix sum of lengths -> parallel nybble scatter.  CBR/ABR run their binary search on the
//    device with size-only passes.
//
// This file is compiled with -fmad=false: float and double expressions keep the reference's operation order.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "ulcb_encode.h"

#ifdef ULCB_PHASE_TIMING
__device__ unsigned long long g_ulcbPhase[32];
#define TICK(i) do { if (tid == 0) { long long t_ = clock64(); phaseAcc[i] += t_ - tPrev; tPrev = t_; } } while (0)
#else
#define TICK(i) do { } while (0)
#endif

namespace {

constexpr float kLn2f = 0x1.62E430p-1f;

// Decimation layouts, one nybble per sub-block, LSB first: bits0-2 shift, bit3 transient
// (libulc/ulcHelper.h:24-46, FormatSpecs.md:35-51)
__constant__ uint16_t kLayoutE[16] = {
    0x0000, 0x0008, 0x0019, 0x0091, 0x012A, 0x01A2, 0x02A1, 0x0A21,
    0x123B, 0x12B3, 0x13B2, 0x1B32, 0x23B1, 0x2B31, 0x3B21, 0xB321,
};

struct SubInfo { int off, S, lgS, sh, ovL, ovR; };

constexpr int ULCB_ENC_MAXCH = 8;      // most channels the run-time-channel-count instantiation takes

// MAXU = channels x sub-blocks (4 at most) = units of the coder and of the Bark analysis; + one amplitude unit per sub-block
template <int MAXU>
struct EncSmallT {
    static constexpr int MAXBU = MAXU + 4;
    float   env[3];
    float   seg[32];                 // {sum,cnt} x 16: [0..7] previous block, [8..15] current
    SubInfo sub[4];
    int     nsub;
    int     wcPrev, wcCur, wcNext;
    int     wcRing[2];               // decisions published by the detector warp
    int     nNz;
    float   cx;
    int     unitSel[MAXU + 1];
    int     unitZones[MAXU];
    int     unitLastQ[MAXU];
    int     tailLen[MAXU];
    uint32_t tailTok[MAXU];
    int     passTotal;
    int     selBin, selBelow;        // radix-select hand-off
    int     scan[40];
    float   band[MAXBU][ULCB_NBARK];
};
// f64 scratch of the Bark analysis per unit: lo[MAXBU][NBARK][3] then stage[MAXBU][3][32]
__host__ __device__ constexpr size_t band_scratch_bytes(int maxbu) { return sizeof(double) * (size_t)maxbu * (ULCB_NBARK * 3 + 3 * 32); }

// ---- small exact helpers -------------------------------------------------------------------------
// logf as the host libm evaluates it (glibc >= 2.28, sysdeps/ieee754/flt-32/e_logf.c: the Arm optimized-routines algorithm -
// 16-entry {1/c, log c} table, cubic in f64, one rounding to f32).  Restated from the published algorithm with the table read
// out of the host's libm; checked on the host against libm logf over EVERY positive finite float (2 139 095 039 values): 0
// mismatches, with and without fused multiply-adds ((float)log((double)x), used before, differs on 416 909 of them).
__constant__ uint64_t kLogfTab[16][2] = {
    {0x3ff661ec79f8f3beull, 0xbfd57bf7808caadeull}, {0x3ff571ed4aaf883dull, 0xbfd2bef0a7c06ddbull}, {0x3ff49539f0f010b0ull, 0xbfd01eae7f513a67ull},
    {0x3ff3c995b0b80385ull, 0xbfcb31d8a68224e9ull}, {0x3ff30d190c8864a5ull, 0xbfc6574f0ac07758ull}, {0x3ff25e227b0b8ea0ull, 0xbfc1aa2bc79c8100ull},
    {0x3ff1bb4a4a1a343full, 0xbfba4e76ce8c0e5eull}, {0x3ff12358f08ae5baull, 0xbfb1973c5a611cccull}, {0x3ff0953f419900a7ull, 0xbfa252f438e10c1eull},
    {0x3ff0000000000000ull, 0x0000000000000000ull}, {0x3fee608cfd9a47acull, 0x3faaa5aa5df25984ull}, {0x3feca4b31f026aa0ull, 0x3fbc5e53aa362eb4ull},
    {0x3feb2036576afce6ull, 0x3fc526e57720db08ull}, {0x3fe9c2d163a1aa2dull, 0x3fcbc2860d224770ull}, {0x3fe886e6037841edull, 0x3fd1058bc8a07ee1ull},
    {0x3fe767dcf5534862ull, 0x3fd4043057b6ee09ull},
};
__device__ __forceinline__ float logf_ref(float x) {
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {               // zero, subnormal, negative, inf, nan
        if (ix * 2u == 0u) return -INFINITY;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return __uint_as_float(0x7fc00000u);
        ix = __float_as_uint(x * 0x1p23f) - (23u << 23);
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15, k = (int)tmp >> 23;
    const double z = (double)__uint_as_float(ix - (tmp & 0xff800000u));
    const double invc = __longlong_as_double((long long)kLogfTab[i][0]), logc = __longlong_as_double((long long)kLogfTab[i][1]);
    const double r = fma(z, invc, -1.0);
    const double y0 = fma((double)k, 0x1.62e42fefa39efp-1, logc);
    const double r2 = r * r;
    double y = fma(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    y = fma(-0x1.00ea348b88334p-2, r2, y);
    y = fma(y, r2, y0 + r);
    return (float)y;
}
// expf as the host libm evaluates it (glibc >= 2.27, sysdeps/ieee754/flt-32/e_expf.c: the Arm optimized-routines
// algorithm - 32-entry 2^(i/32) table, cubic in f64, one rounding to f32).  Restated from the published algorithm;
// checked exhaustively on the host against libm expf over every float in [-88, 88]: 2 mismatches in 2.2e9.
__constant__ uint64_t kExp2Tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};
__device__ __forceinline__ float expf_ref(float x) {
    if (!(fabsf(x) < 88.0f)) return (float)exp((double)x);
    const double Nn = 32.0;
    const double InvLn2N = 0x1.71547652b82fep+0 * Nn;
    const double C0 = 0x1.c6af84b912394p-5 / Nn / Nn / Nn, C1 = 0x1.ebfce50fac4f3p-3 / Nn / Nn, C2 = 0x1.62e42ff0c52d6p-1 / Nn;
    const double SHIFT = 0x1.8p+52;
    double z = InvLn2N * (double)x;
    double kd = z + SHIFT;
    uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd -= SHIFT;
    double r = z - kd;
    uint64_t t = kExp2Tab[ki & 31] + (ki << 47);
    double s = __longlong_as_double((long long)t);
    z = C0 * r + C1;
    double r2 = r * r;
    double y = C2 * r + 1.0;
    y = z * r2 + y;
    y = y * s;
    return (float)y;
}

__device__ __forceinline__ float fast_log(float x) {              // libulc/ulcHelper.h:127-136
    uint32_t b = __float_as_uint(x);
    int e = (int)(b >> 23) - 127;
    x = __uint_as_float((127u << 23) | (b & 0x7FFFFFu));
    return -1.7417939f + (2.8212026f + (-1.4699568f + (0.44717955f - 0.056570851f * x) * x) * x) * x + 0.6931471806f * (float)e;
}
__device__ __forceinline__ int compand_u(float v) {               // ulcHelper.h:51-60
    return (v >= 0.5f) ? (int)(0.5f + sqrtf(v - 0.25f)) : 0;
}
__device__ __forceinline__ int compand_clip(float v, int limit) { // ulcHelper.h:62-91
    int q = compand_u(fabsf(v));
    if (q > limit) q = limit;
    return v < 0.0f ? -q : q;
}
__device__ __forceinline__ int first_overlap(int wc, int N) {
    unsigned pat = kLayoutE[(wc >> 4) & 15];
    int ov = N >> (pat & 7);
    if ((pat & 8) && (wc >> 4) != 0) ov >>= (wc & 7);
    return ov;
}
__device__ __forceinline__ int last_subsize(int wc, int N) {
    unsigned pat = kLayoutE[(wc >> 4) & 15];
    int sh = pat & 7;
    while (pat >>= 4) sh = pat & 7;
    return N >> sh;
}
__device__ __forceinline__ bool node_is_dct4(int d, int node) {   // root = DCT-IV; DCT-IV -> (II, II); DCT-II -> (II, IV)
    unsigned h = (1u << d) | (unsigned)node;
    int ones = __ffs(~h) - 1;                                     // trailing ones of the heap index
    return ones & 1;
}

// The CTA is NTHR worker threads + one extra warp that runs the transient detector one block ahead.
// Workers synchronise among themselves on named barrier 1; named barrier 2 joins everybody once per block (the detector warp
// arrives from its own loop, so this one is the unaligned barrier.sync: bar.sync promises one call site for all its threads).
// (bar.sync is warp-aligned: a warp counts as arrived as soon as any of its lanes executes it, so each warp is
// reconverged explicitly first - single-lane serial sections precede several of these barriers.)
template <int NTHR> __device__ __forceinline__ void wsync() { __syncwarp(); asm volatile("bar.sync 1, %0;" ::"n"(NTHR) : "memory"); }
template <int NTHR> __device__ __forceinline__ void fsync() { __syncwarp(); asm volatile("barrier.sync 2, %0;" ::"n"(NTHR + 32) : "memory"); }
template <int NTHR, bool PRE> __device__ __forceinline__ void csync() { if (PRE) wsync<NTHR>(); else fsync<NTHR>(); }

// Worker-wide exclusive scans of one int per thread.  `scr` = smem int[2][20]: the warp totals go to the half `flip` selects and
// every warp scans them for itself after ONE barrier; consecutive calls alternate halves, so the next call's writes cannot
// overtake this call's reads (a warp can only get there through the barrier of the call in between).
template <int NTHR>
__device__ __forceinline__ int block_excl_sum(int v, int *scr, int &flip, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = NTHR / 32;
    int *mine = scr + 20 * flip; flip ^= 1;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) mine[warp] = inc;
    wsync<NTHR>();
    const int t = lane < NW ? mine[lane] : 0;
    int ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= d) ti += u; }
    total = __shfl_sync(0xffffffffu, ti, 31);
    return inc - v + __shfl_sync(0xffffffffu, ti - t, warp);
}
template <int NTHR>
__device__ __forceinline__ int block_excl_max(int v, int *scr, int &flip) {     // v >= 0; identity 0
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = NTHR / 32;
    int *mine = scr + 20 * flip; flip ^= 1;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = max(inc, t); }
    int exl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exl = 0;
    if (lane == 31) mine[warp] = inc;
    wsync<NTHR>();
    int ti = lane < NW ? mine[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= d) ti = max(ti, u); }
    int te = __shfl_up_sync(0xffffffffu, ti, 1);                       // max over the warps before `lane`
    if (lane == 0) te = 0;
    return max(exl, __shfl_sync(0xffffffffu, te, warp));
}

// ---- pointers into the per-CTA work area ------------------------------------------------------------
struct Work {
    float *R;            // [3][CN] sample timeline, rotating: block k lives at R + (k % 3)*CN (M/S planar)
    float *XA, *XB;      // DCT "X" buffers: MDCT / MDST, [CN] each; XA becomes the normalised coefficients
    float *TA, *TB;      // DCT "T" buffers; afterwards TA = keys -> ranks/selection list, TB = noise pairs
    float *amp2, *mask;  // [N/2]
    uint16_t *cnt;       // radix counters [16][NTHR]
    uint8_t *encScratch; // 3*CN bytes for the coder (aliases cnt, which is dead after the sort; global for large blocks)
    double (*lo)[ULCB_NBARK][3];   // Bark analysis: per-unit band sums ...
    double (*stage)[3][32];        // ... and the 32-line staging of the three f64 chains
};

// ---- transient detector (ulcEncoder_WindowControl.c:140-239) decision on the 16 segment sums ----------
__device__ int window_decide(const float *seg, int log2N) {
    const float *cur = seg + 16;                      // entries are {sum,cnt} pairs; cur = entry 8
    int log2sub = log2N - 3, nSeg = 8, segSz = 1;
    if (log2sub < 6) { int sh = 6 - log2sub; nSeg >>= sh; segSz <<= sh; log2sub = 6; }
    int decimation = 1; float trRatio = 0.0f;
    for (;;) {
        log2sub++;
        int best = 0; float maxRatio = -1000.0f;
        for (int s = 0; s < nSeg; s++) {
            float Ls = 0.0f, Lc = 0.0f, Rs = 0.0f, Rc = 0.0f;
            const float *src = cur + 2 * s * segSz;
            for (int n = 0; n < segSz; n++) {
                Ls += src[2 * (n - segSz)]; Lc += src[2 * (n - segSz) + 1];
                Rs += src[2 * n];           Rc += src[2 * n + 1];
            }
            float l = (Ls != 0.0f) ? logf_ref(Ls / Lc) : -100.0f;
            float r = (Rs != 0.0f) ? logf_ref(Rs / Rc) : -100.0f;
            float ratio = r - l; if (ratio < 0.0f) ratio = -ratio;
            if (ratio > maxRatio) { best = s; maxRatio = ratio; }
        }
        if (maxRatio - trRatio < kLn2f) break;
        decimation = nSeg + best;
        trRatio = maxRatio;
        if (nSeg > 1 && trRatio < kLn2f) { nSeg /= 2; segSz *= 2; } else break;
    }
    if (trRatio < 0x1.62E430p-2f) return 0x10;
    trRatio *= 0x1.715476p0f;
    int ov = (trRatio < 0.5f) ? 0 : (trRatio >= 6.5f) ? 7 : __float2int_rn(trRatio);
    if (log2sub - ov < 6) ov = log2sub - 6;
    return ov + 8 * (decimation != 1) + 16 * decimation;
}

// ---- 8-point leaves (Fourier_DCT4.c:64-111, Fourier_DCT2.c:64-109) --------------------------------------
__device__ __forceinline__ void dct4_leaf8(float *x) {
    const float R = 0x1.6A09E6p-1f, C13 = 0x1.D906BCp-1f, S13 = 0x1.87DE2Ap-2f;
    const float CK0 = 0x1.FD88DAp-1f, CK1 = 0x1.E9F416p-1f, CK2 = 0x1.C38B30p-1f, CK3 = 0x1.8BC806p-1f;
    const float SK0 = 0x1.917A6Cp-4f, SK1 = 0x1.294062p-2f, SK2 = 0x1.E2B5D4p-2f, SK3 = 0x1.44CF32p-1f;
    float x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = x[4], x5 = x[5], x6 = x[6], x7 = x[7];
    float p0 = CK0 * x0 + SK0 * x7, q0 = SK0 * x0 - CK0 * x7;
    float p1 = CK1 * x1 + SK1 * x6, q1 = -SK1 * x1 + CK1 * x6;
    float p2 = CK2 * x2 + SK2 * x5, q2 = SK2 * x2 - CK2 * x5;
    float p3 = CK3 * x3 + SK3 * x4, q3 = -SK3 * x3 + CK3 * x4;
    float pa = p0 + p3, pd = p0 - p3, pb = p1 + p2, pe = p1 - p2;
    float qa = q3 + q0, qd = q3 - q0, qb = q2 + q1, qe = q2 - q1;
    float sx = pa + pb, sy = pa - pb;
    float tx = C13 * pd + S13 * pe, ty = S13 * pd - C13 * pe;
    float ux = qa + qb, uy = qa - qb;
    float vx = C13 * qd + S13 * qe, vy = S13 * qd - C13 * qe;
    x[0] = sx; x[1] = tx - vy; x[2] = tx + vy; x[3] = (sy + uy) * R;
    x[4] = (sy - uy) * R; x[5] = ty - vx; x[6] = ty + vx; x[7] = ux;
}
__device__ __forceinline__ void dct2_leaf8(float *x) {
    const float R = 0x1.6A09E6p-1f;
    const float C14 = 0x1.F6297Cp-1f, S14 = 0x1.8F8B84p-3f, C34 = 0x1.A9B662p-1f, S34 = 0x1.1C73B4p-1f;
    const float C64 = 0x1.87DE2Ap-2f, S64 = 0x1.D906BCp-1f;
    float x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = x[4], x5 = x[5], x6 = x[6], x7 = x[7];
    float s0 = x0 + x7, d0 = x0 - x7, s1 = x1 + x6, d1 = x1 - x6;
    float s2 = x2 + x5, d2 = x2 - x5, s3 = x3 + x4, d3 = x3 - x4;
    float ee = s0 + s3, eo = s0 - s3, fe = s1 + s2, fo = s1 - s2;
    float gx = C34 * d3 + S34 * d0, gy = -S34 * d3 + C34 * d0;
    float hx = C14 * d2 + S14 * d1, hy = -S14 * d2 + C14 * d1;
    float a0 = ee + fe, b0 = ee - fe;
    float c0 = C64 * fo + S64 * eo, e0 = -S64 * fo + C64 * eo;
    float a1 = gy + hx, c1 = gy - hx, e1 = gx + hy, b1 = gx - hy;
    x[0] = a0; x[4] = b0 * R; x[2] = c0; x[6] = e0;
    x[1] = (a1 + e1) * R; x[5] = b1; x[3] = c1; x[7] = (a1 - e1) * R;
}

template <class SM>
__device__ __forceinline__ int find_sub(const SM &sm, int n) {     // sub-block holding in-channel index n
    int j = 0;
    const int ns = sm.nsub;
    if (ns > 1) {
        j = (n >= sm.sub[1].off);
        if (ns > 2) { j += (n >= sm.sub[2].off); if (ns > 3) j += (n >= sm.sub[3].off); }
    }
    return j;
}

// ---- nybble sinks ----------------------------------------------------------------------------------
struct CountSink { int n; __device__ __forceinline__ void put(int) { n++; } };
struct WriteSink {                                   // a thread's nybbles are contiguous: gather a word, one atomicOr per word
    uint32_t *words; int pos, cap; uint32_t acc = 0;
    __device__ __forceinline__ void put(int x) {
        if (pos < cap) acc |= (uint32_t)(x & 15) << ((pos & 7) * 4);
        pos++;
        if ((pos & 7) == 0) flush_word(pos - 1);
    }
    __device__ __forceinline__ void flush_word(int lastPos) {
        if (acc) atomicOr(&words[lastPos >> 3], acc);
        acc = 0;
    }
    __device__ __forceinline__ void finish() { if (pos & 7) flush_word(pos); }
};

// Counts like CountSink and keeps the thread's first 64 nybbles in its own 8-word slot, so the write pass is a shifted
// copy instead of a second walk over the items (noise-fill decisions, quantiser headers, ...).
struct StageSink {
    uint32_t *slot; int n = 0; uint32_t acc = 0;
    __device__ __forceinline__ void put(int x) {
        if (n < 64) {
            acc |= (uint32_t)(x & 15) << ((n & 7) * 4);
            if ((n & 7) == 7) { slot[n >> 3] = acc; acc = 0; }
        }
        n++;
    }
    __device__ __forceinline__ void finish() { if (n < 64 && (n & 7)) slot[n >> 3] = acc; }
};

// noise amplitude code of a run (ulcEncoder_NoiseFill.c:15-36)
__device__ int noise_q(const float *pairs, int band, int n, float q) {
    pairs += band & ~1;
    n = (n + (band & 1) + 1) >> 1;
    float s = 0.0f, sw = 0.0f;
#pragma unroll 4
    for (int i = 0; i < n; i++) { float2 p = *reinterpret_cast<const float2 *>(pairs + 2 * i); s += p.y; sw += p.x; }
    if (s == 0.0f) return 0;
    float amp = expf_ref(s / sw);
    int r = compand_u(amp * q);
    return r < 8 ? r : 8;
}
// HF extension fit (ulcEncoder_NoiseFill.c:41-94)
__device__ void hf_ext_params(const float *pairs, int band, int n, float q, int &nq, int &dq) {
    pairs += band & ~1;
    n = (n + (band & 1) + 1) >> 1;
    float sx = 0.0f, sx2 = 0.0f, sxy = 0.0f, sy = 0.0f, sw = 0.0f;
    float xf = 0.0f;                                              // (float)i * 2.0f, exact: small even integers
#pragma unroll 4
    for (int i = 0; i < n; i++, xf += 2.0f) {
        float2 p = *reinterpret_cast<const float2 *>(pairs + 2 * i);
        float x = xf, w = p.x, wy = p.y;
        sx += w * x; sx2 += w * x * x; sxy += x * wy; sy += wy; sw += w;
    }
    float det = sw * sx2 - sx * sx;
    if (det == 0.0f) { nq = dq = 0; return; }
    float amp = (sx2 * sy - sx * sxy) / det, dec = (sw * sxy - sx * sy) / det;
    amp = expf_ref(amp);
    dec = (dec < 0.0f) ? expf_ref(dec) : 1.0f;
    int a = compand_u(amp * q * 4.0f); if (a > 16) a = 16;
    int d = compand_u((dec - 1.0f) * -0x1.0p19f);
    if (!d) return;
    if (d > 0xFF) d = 0xFF;
    nq = a; dq = d;
}

// tokens of one kept coefficient: the gap [nc, cur) then the coefficient (ulcEncoder_Encode.c:114-192)
template <class Sink>
__device__ void put_item(Sink &w, const float *coef, const float *pairs, int nc, int cur, float q) {
    int gap = cur - nc;
    while (gap) {
        int n, v = 0;
        if (gap <= 2) {
            int q1 = compand_clip(coef[nc] * q, 7), q2 = 0;
            if (gap >= 2) q2 = compand_clip(coef[nc + 1] * q, 7);
            if (abs(q1) > 1 && (gap < 2 || abs(q2) > 1)) {
                w.put(q1);
                if (gap >= 2) w.put(q2);
                break;
            }
        }
        int nq = 0;
        n = 0;
        if (gap >= 16) {
            v = gap - 16; if (v > 0x1FF) v = 0x1FF;
            n = v + 16;
            nq = noise_q(pairs, nc, n, q);
        }
        if (nq) { w.put(0x8); w.put(v >> 5); w.put(v >> 1); w.put((v & 1) | ((nq - 1) << 1)); }
        else if (gap < 33) { v = gap - 1; if (v > 0xF) v = 0xF; n = v + 1; w.put(0x0); w.put(v); }
        else { v = gap - 33; if (v > 0xFF) v = 0xFF; n = v + 33; w.put(0x1); w.put(v >> 4); w.put(v); }
        nc += n; gap -= n;
    }
    w.put(compand_clip(coef[cur] * q, 7));
}

// slot info byte: bits0-4 quantiser index (5..31, so non-zero = selected), bit5 emit quantiser, bit6 first zone of its sub-block
enum { INF_Q = 31, INF_HDR = 32, INF_FIRST = 64 };

template <class Sink>
__device__ __forceinline__ void put_quant(Sink &w, int inf) {            // ulcEncoder_Encode.c:32-45
    int s = (inf & INF_Q) - 5;
    if (!(inf & INF_FIRST)) w.put(0xF);
    if (s < 0xE) w.put(s);
    else { w.put(0xE); w.put(s - 0xE); }
}

// ------------------------------------------------------------------------------------------------
// PRE = the window controls come from the detector pre-pass (ulcb_detect_kernel): no detector warp, no sample ring in shared
// memory (the fold reads the PCM input directly), scratch aliased so that two CTAs fit one SM.
template <int NCHT, int NTHR, bool PRE = false>   // NCHT = channels, 0 = taken from the arguments (3 .. ULCB_ENC_MAXCH)
struct Encoder {
    static constexpr int MAXCH = NCHT ? NCHT : ULCB_ENC_MAXCH;
    using EncSmall = EncSmallT<4 * MAXCH>;
    static constexpr int MAXU = 4 * MAXCH;
    const UlcbEncodeArgs &args;
    const UlcbEncTables &tb;
    EncSmall &sm;
    Work w;
    const int nchRt, N, CN, H, tid;
    int scanFlip = 0;            // which half of sm.scan the next worker-wide scan uses
    __device__ __forceinline__ int nch() const { return NCHT ? NCHT : nchRt; }
    __device__ Encoder(const UlcbEncodeArgs &a, EncSmall &s, const Work &wk)
        : args(a), tb(a.tb), sm(s), w(wk), nchRt(a.nChan), N(a.tb.N), CN((NCHT ? NCHT : a.nChan) * a.tb.N), H(a.tb.N >> 1), tid(threadIdx.x) {}

    // Coefficients per thread in the rank and coder phases: thread t owns [t*E, (t+1)*E).  E is a power of two, so it divides every
    // sub-block size (N/8 >= 32 >= E) and a thread's slots never straddle a sub-block or a channel; with a channel count that is
    // not a power of two the last threads own nothing.
    __device__ __forceinline__ int slot_items() const { const int e = CN / NTHR; return (e & (e - 1)) ? 1 << (32 - __clz(e)) : e; }
    __device__ __forceinline__ float *blk(int k) const { return w.R + ((k + 3) % 3) * CN; }
    // log of the noise line sums of unit (c, sub-block s): a dead sample block, or (PRE) the upper half of the unit's TB range,
    // which the line sums themselves leave free
    __device__ __forceinline__ float *lg_lines(int k, int c, const SubInfo &s) const {
        return PRE ? w.TB + c * N + s.off + (s.S >> 1) : blk(k - 2) + c * H + (s.off >> 1);
    }
    // packed nybbles of the block being coded: a dead sample block, or (PRE) the amp2 / mask lines, dead once the keys are made
    __device__ __forceinline__ uint32_t *out_words(int k) const { return reinterpret_cast<uint32_t *>(PRE ? w.amp2 : blk(k - 2)); }

    // ---------------- ingest ----------------
    __device__ void ingest(int k, const void *pcm, long long validSamples) {
        float *R0 = blk(k);
        for (int n = tid; n < N; n += NTHR) {
            float v[MAXCH];
            bool ok = n < validSamples;
#pragma unroll
            for (int c = 0; c < nch(); c++) {
                float x = 0.0f;
                if (ok) {
                    if (args.pcmFloat) x = reinterpret_cast<const float *>(pcm)[(size_t)n * nch() + c];
                    else x = (float)reinterpret_cast<const int16_t *>(pcm)[(size_t)n * nch() + c] * 0x1.0p-15f;   // tools/WavIO_Helper.c:49-55
                }
                v[c] = x;
            }
#pragma unroll
            for (int c = 1; c < nch(); c += 2) {                       // M/S (ulcEncoder_BlockTransform.c:102-111)
                float a = v[c - 1], b = v[c];
                v[c - 1] = (a + b) * 0.5f;
                v[c]     = (a - b) * 0.5f;
            }
#pragma unroll
            for (int c = 0; c < nch(); c++) R0[c * N + n] = v[c];
        }
    }

    // ---------------- per-block setup (worker thread 0): window controls and the sub-block table ----------------
    __device__ void begin_block(int decision) {
        if (tid == 0) {
            sm.wcPrev = sm.wcCur;
            sm.wcCur  = sm.wcNext;
            sm.wcNext = decision;
            // sub-block table of the block being coded now (ulcEncoder_BlockTransform.c:156-172)
            const int wc = sm.wcCur;
            unsigned pat = kLayoutE[(wc >> 4) & 15];
            int j = 0, off = 0;
            int trs[4];
            do {
                SubInfo &s = sm.sub[j];
                s.sh = pat & 7; s.S = N >> s.sh; s.lgS = tb.log2N - s.sh; s.off = off;
                trs[j] = ((pat & 8) && (wc >> 4) != 0) ? (wc & 7) : 0;
                off += s.S; j++;
            } while (pat >>= 4);
            sm.nsub = j;
            for (int i = 0; i < j; i++) {
                int ov = (i + 1 < j) ? (sm.sub[i + 1].S >> trs[i + 1]) : first_overlap(sm.wcNext, N);
                if (ov > sm.sub[i].S) ov = sm.sub[i].S;
                sm.sub[i].ovR = ov;
            }
            for (int i = 0; i < j; i++) {
                if (i > 0) sm.sub[i].ovL = sm.sub[i - 1].ovR;
                else { int ov = first_overlap(wc, N), sp = last_subsize(sm.wcPrev, N); sm.sub[0].ovL = ov < sp ? ov : sp; }
            }
            sm.nNz = 0;
        }
        wsync<NTHR>();
    }

    // ---------------- fold + window (closed form of the lapping FIFO and Fourier_MDCT_Template.h:45-187) ----------------
    __device__ __forceinline__ float lap_weight(float x, int u, int hv) const {      // u = offset from the fold point, hv = overlap / 2
        if (u < -hv) return 0.0f;
        if (u >= hv) return x;
        if (u < 0) return __ldg(&tb.rot[hv + (u + hv)]).y * x;
        return __ldg(&tb.rot[hv + (hv - 1 - u)]).x * x;
    }
    __device__ __forceinline__ float lap_value(const float *R2, const float *R1, int c, int rel, int u, int ovL) const {
        // rel = index on the 2N timeline [block k-2 | block k-1]
        const int hv = ovL >> 1;
        if (u < -hv) return 0.0f;
        float x = (rel < N) ? R2[c * N + rel] : R1[c * N + rel - N];
        return lap_weight(x, u, hv);
    }
    // M/S sample point t of the stream straight from the PCM input (zero outside [0, nSamples)): what ingest() would have put
    // into the sample ring (ulcEncoder_BlockTransform.c:87-111, tools/WavIO_Helper.c:49-55)
    __device__ __forceinline__ void load_ms(const unsigned char *pcm, long long t, long long nSamples, float *v) const {
        if (t < 0 || t >= nSamples) {
#pragma unroll
            for (int c = 0; c < MAXCH; c++) v[c] = 0.0f;
            return;
        }
        if (NCHT == 2 && !args.pcmFloat) {
            const uint32_t wv = __ldg(reinterpret_cast<const uint32_t *>(pcm) + t);
            v[0] = (float)(int16_t)(wv & 0xffffu) * 0x1.0p-15f; v[1] = (float)(int16_t)(wv >> 16) * 0x1.0p-15f;
        } else {
#pragma unroll
            for (int c = 0; c < MAXCH; c++)
                if (c < nch()) v[c] = args.pcmFloat ? __ldg(reinterpret_cast<const float *>(pcm) + (size_t)t * nch() + c)
                                                    : (float)__ldg(reinterpret_cast<const int16_t *>(pcm) + (size_t)t * nch() + c) * 0x1.0p-15f;
        }
#pragma unroll
        for (int c = 1; c < MAXCH; c += 2)
            if (c < nch()) { float a = v[c - 1], b = v[c]; v[c - 1] = (a + b) * 0.5f; v[c] = (a - b) * 0.5f; }
    }
    // fold + window of output block b with the two input blocks b-2, b-1 read from the PCM input (PRE): the same values and
    // operations as fold(), one thread per line pair for every channel
    __device__ void fold_pre(int b, const unsigned char *pcm, long long nSamples) {
        const long long t0 = (long long)(b - 2) * N;                  // stream position of timeline index 0
        for (int r = tid; r < H; r += NTHR) {
            const int j = find_sub(sm, 2 * r);
            const SubInfo s = sm.sub[j];
            const int hS = s.S >> 1, i = r - (s.off >> 1);
            const int hrel = H + s.off, hv = s.ovL >> 1;
            float xa[MAXCH], xb[MAXCH], xc[MAXCH], xd[MAXCH];
            if (i - hS >= -hv) load_ms(pcm, t0 + hrel - hS + i, nSamples, xa);
            else {
#pragma unroll
                for (int c = 0; c < MAXCH; c++) xa[c] = 0.0f;
            }
            load_ms(pcm, t0 + hrel + hS - 1 - i, nSamples, xb);
            load_ms(pcm, t0 + hrel + hS + i, nSamples, xc);
            load_ms(pcm, t0 + hrel + hS + s.S - 1 - i, nSamples, xd);
            const int flat = (s.S - s.ovR) >> 1;
            float2 rt = make_float2(0.0f, 0.0f);
            if (i >= flat) rt = __ldg(&tb.rot[(s.ovR >> 1) + (i - flat)]);
#pragma unroll
            for (int c = 0; c < MAXCH; c++) {
                if (c >= nch()) break;
                const float a = lap_weight(xa[c], i - hS, hv), br = lap_weight(xb[c], hS - 1 - i, hv);
                const float cc = xc[c], dr = xd[c];
                float mc, ms;
                if (i < flat) { mc = cc; ms = cc; }
                else { float c2 = cc * rt.x, d2 = dr * rt.y; mc = c2 + d2; ms = c2 - d2; }
                float mc2 = -a + br, ms2 = a + br;
                const int base = c * N + s.off;
                const int lo = hS - 1 - i, hi = hS + i;
                w.XA[base + lo] = mc;
                w.XA[base + hi] = mc2;
                w.XB[base + lo] = (lo & 1) ? -ms : ms;
                w.XB[base + hi] = (hi & 1) ? -ms2 : ms2;
            }
        }
    }
    __device__ void fold(int k) {
        const float *R1 = blk(k - 1), *R2 = blk(k - 2);
        for (int P = tid; P < nch() * H; P += NTHR) {
            const int c = P >> (tb.log2N - 1), r = P & (H - 1);
            const int j = find_sub(sm, 2 * r);
            const SubInfo s = sm.sub[j];
            const int hS = s.S >> 1, i = r - (s.off >> 1);
            const int hrel = H + s.off;                               // fold point on the 2N timeline
            float a  = lap_value(R2, R1, c, hrel - hS + i, i - hS, s.ovL);
            float br = lap_value(R2, R1, c, hrel + hS - 1 - i, hS - 1 - i, s.ovL);
            auto fresh = [&](int q) -> float { int rel = hrel + hS + q; return (rel < N) ? R2[c * N + rel] : R1[c * N + rel - N]; };
            float cc = fresh(i), dr = fresh(s.S - 1 - i);
            const int flat = (s.S - s.ovR) >> 1;
            float mc, ms;
            if (i < flat) { mc = cc; ms = cc; }
            else {
                float2 rt = __ldg(&tb.rot[(s.ovR >> 1) + (i - flat)]);
                float c2 = cc * rt.x, d2 = dr * rt.y;
                mc = c2 + d2; ms = c2 - d2;
            }
            float mc2 = -a + br, ms2 = a + br;
            const int base = c * N + s.off;
            const int lo = hS - 1 - i, hi = hS + i;
            w.XA[base + lo] = mc;
            w.XA[base + hi] = mc2;
            w.XB[base + lo] = (lo & 1) ? -ms : ms;                    // odd-index sign flip of the MDST input (:170-176)
            w.XB[base + hi] = (hi & 1) ? -ms2 : ms2;
        }
    }

    // ---------------- breadth-first DCT-IV of every (buffer, channel, sub-block) ----------------
    __device__ __forceinline__ void locate(int P, int &hb, int &cbase, int &j, int &lp) const {
        // pair index P in [0, CN) -> buffer half, channel base, sub-block, local pair index
        const int lgH = tb.log2N - 1;
        const int cc = P >> lgH, r = P & (H - 1);                     // cc = hb*nch() + c
        hb = cc >= nch(); const int c = hb ? cc - nch() : cc;
        j = find_sub(sm, 2 * r);
        lp = r - (sm.sub[j].off >> 1);
        cbase = c * N + sm.sub[j].off;
    }
    __device__ void dct_down(int m, int lgm) {
        const int hm = m >> 1;
#pragma unroll 4
        for (int P = tid; P < CN; P += NTHR) {
            int hb, cbase, j, lp; locate(P, hb, cbase, j, lp);
            const int lgS = sm.sub[j].lgS;
            if (lgS < lgm) continue;
            const int d = lgS - lgm, node = lp >> (lgm - 1), i = lp & (hm - 1);
            float *X = hb ? w.XB : w.XA, *T = hb ? w.TB : w.TA;
            const float *src = ((d & 1) ? T : X) + cbase + node * m;
            float *dst = ((d & 1) ? X : T) + cbase + node * m;
            float a = src[i], b = src[m - 1 - i];
            if (node_is_dct4(d, node)) {                              // Fourier_DCT4_Template.h:49-68
                float2 rt = __ldg(&tb.rot[hm + i]);
                dst[i] = rt.x * a + rt.y * b;
                float p = rt.y * a, q = rt.x * b;
                dst[hm + i] = (i & 1) ? (q - p) : (p - q);
            } else {                                                  // Fourier_DCT2_Template.h:24-40
                dst[i] = a + b;
                dst[hm + i] = a - b;
            }
        }
    }
    __device__ void dct_leaves(int inX = -1) {                        // inX: -1 = by depth parity, else 1 = X buffers, 0 = T buffers
        for (int Lf = tid; Lf < (CN >> 2); Lf += NTHR) {              // 2*CN/8 leaves
            const int half = CN >> 3;
            int hb = Lf >= half; int q = Lf - (hb ? half : 0);
            int pos = q << 3;                                         // element position in [0, CN)
            int c = pos >> tb.log2N, n = pos & (N - 1);
            int j = find_sub(sm, n);
            const SubInfo &s = sm.sub[j];
            int d = s.lgS - 3, node = (n - s.off) >> 3;
            const bool useT = inX < 0 ? (d & 1) : !inX;
            float *buf = (hb ? (useT ? w.TB : w.XB) : (useT ? w.TA : w.XA)) + pos;
            if (node_is_dct4(d, node)) dct4_leaf8(buf); else dct2_leaf8(buf);
        }
    }
    __device__ void dct_up(int m, int lgm) {
        const int hm = m >> 1;
#pragma unroll 4
        for (int P = tid; P < CN; P += NTHR) {
            int hb, cbase, j, lp; locate(P, hb, cbase, j, lp);
            const int lgS = sm.sub[j].lgS;
            if (lgS < lgm) continue;
            const int d = lgS - lgm, node = lp >> (lgm - 1), i = lp & (hm - 1);
            float *X = hb ? w.XB : w.XA, *T = hb ? w.TB : w.TA;
            const float *src = ((d & 1) ? X : T) + cbase + node * m;  // children live at depth d+1
            float *dst = ((d & 1) ? T : X) + cbase + node * m;
            if (node_is_dct4(d, node)) {                              // Fourier_DCT4_Template.h:100-114
                if (i == hm - 1) { dst[0] = src[0]; dst[m - 1] = src[hm]; }
                else {
                    float a = src[i + 1], b = src[m - 1 - i];
                    dst[2 * i + 1] = a + b;
                    dst[2 * i + 2] = a - b;
                }
            } else {                                                  // Fourier_DCT2_Template.h:62-77
                dst[2 * i] = src[i];
                dst[2 * i + 1] = src[hm + i];
            }
        }
    }
    // The same passes for a block that is ONE sub-block (no window switching, the common case): the recursion depth is
    // uniform across the CTA, so buffer parity, node size and the pair index need no per-element sub-block lookup.
    // Identical arithmetic per element.
    __device__ void dct_down_long(int m, int lgm, bool srcIsX) {
        const int hm = m >> 1, lgH = tb.log2N - 1, d = tb.log2N - lgm;
        const float *const S0 = srcIsX ? w.XA : w.TA, *const S1 = srcIsX ? w.XB : w.TB;
        float *const D0 = srcIsX ? w.TA : w.XA, *const D1 = srcIsX ? w.TB : w.XB;
#pragma unroll 4
        for (int P = tid; P < CN; P += NTHR) {
            const int cc = P >> lgH, r = P & (H - 1);
            const bool hb = cc >= nch();
            const int node = r >> (lgm - 1), i = r & (hm - 1);
            const int off = (hb ? cc - nch() : cc) * N + (node << lgm);
            const float *src = (hb ? S1 : S0) + off;
            float *dst = (hb ? D1 : D0) + off;
            float a = src[i], b = src[m - 1 - i];
            if (node_is_dct4(d, node)) {                              // Fourier_DCT4_Template.h:49-68
                float2 rt = __ldg(&tb.rot[hm + i]);
                dst[i] = rt.x * a + rt.y * b;
                float p = rt.y * a, q = rt.x * b;
                dst[hm + i] = (i & 1) ? (q - p) : (p - q);
            } else {                                                  // Fourier_DCT2_Template.h:24-40
                dst[i] = a + b;
                dst[hm + i] = a - b;
            }
        }
    }
    // Two recursion levels (node sizes m and m/2) in one pass: the level-m pairs i and m/2-1-i feed exactly the level-m/2
    // pairs i of both children, so a thread carries its four values through both butterflies in registers - half the
    // shared-memory traffic and half the barriers of two single-level passes.  Same operations per value.
    __device__ void dct_down2_long(int lgm, bool srcIsX) {
        const int m = 1 << lgm, hm = m >> 1, qm = m >> 2, lgH = tb.log2N - 1, d = tb.log2N - lgm;
        const float *const S0 = srcIsX ? w.XA : w.TA, *const S1 = srcIsX ? w.XB : w.TB;
        float *const D0 = srcIsX ? w.TA : w.XA, *const D1 = srcIsX ? w.TB : w.XB;
        auto fly = [&](bool dct4, int half, int i, float a, float b, float &lo, float &hi) {
            if (dct4) {                                               // Fourier_DCT4_Template.h:49-68
                const float2 rt = __ldg(&tb.rot[half + i]);
                lo = rt.x * a + rt.y * b;
                const float p = rt.y * a, q = rt.x * b;
                hi = (i & 1) ? (q - p) : (p - q);
            } else { lo = a + b; hi = a - b; }                        // Fourier_DCT2_Template.h:24-40
        };
        if (((H >> 1) % NTHR) == 0) {
            // The quarter-pairs of one (buffer, channel) are a multiple of the thread count: a thread meets the same node, the same
            // pair and the same rotors in every (buffer, channel), so all of that is set up once.  The lower child of any node is a
            // DCT-II; the upper one is a DCT-IV exactly when the node itself is a DCT-II.
            for (int r = tid; r < (H >> 1); r += NTHR) {
                const int node = r >> (lgm - 2), i = r & (qm - 1), i2 = hm - 1 - i;
                const bool t0 = node_is_dct4(d, node);
                float2 r0 = make_float2(0.f, 0.f), r1 = r0, rU = r0;
                if (t0) { r0 = __ldg(&tb.rot[hm + i]); r1 = __ldg(&tb.rot[hm + i2]); }
                else rU = __ldg(&tb.rot[qm + i]);
                const int offn = node << lgm;
                const bool odd = i & 1;                               // i2 = hm-1-i has the opposite parity
#pragma unroll 2
                for (int cc = 0; cc < 2 * nch(); cc++) {
                    const bool hb = cc >= nch();
                    const int off = (hb ? cc - nch() : cc) * N + offn;
                    const float *src = (hb ? S1 : S0) + off;
                    float *dst = (hb ? D1 : D0) + off;
                    const float a0 = src[i], b0 = src[m - 1 - i], a1 = src[i2], b1 = src[hm + i];
                    float lo0, hi0, lo1, hi1, o0, o1, o2, o3;
                    if (t0) {                                         // Fourier_DCT4_Template.h:49-68
                        lo0 = r0.x * a0 + r0.y * b0; { const float p = r0.y * a0, q = r0.x * b0; hi0 = odd ? (q - p) : (p - q); }
                        lo1 = r1.x * a1 + r1.y * b1; { const float p = r1.y * a1, q = r1.x * b1; hi1 = odd ? (p - q) : (q - p); }
                    } else { lo0 = a0 + b0; hi0 = a0 - b0; lo1 = a1 + b1; hi1 = a1 - b1; }     // Fourier_DCT2_Template.h:24-40
                    o0 = lo0 + lo1; o1 = lo0 - lo1;
                    if (!t0) { o2 = rU.x * hi0 + rU.y * hi1; const float p = rU.y * hi0, q = rU.x * hi1; o3 = odd ? (q - p) : (p - q); }
                    else { o2 = hi0 + hi1; o3 = hi0 - hi1; }
                    dst[i] = o0; dst[qm + i] = o1; dst[hm + i] = o2; dst[hm + qm + i] = o3;
                }
            }
            return;
        }
#pragma unroll 2
        for (int Q = tid; Q < (CN >> 1); Q += NTHR) {
            const int cc = Q >> (lgH - 1), r = Q & ((H >> 1) - 1);
            const bool hb = cc >= nch();
            const int node = r >> (lgm - 2), i = r & (qm - 1), i2 = hm - 1 - i;
            const int off = (hb ? cc - nch() : cc) * N + (node << lgm);
            const float *src = (hb ? S1 : S0) + off;
            float *dst = (hb ? D1 : D0) + off;
            const float a0 = src[i], b0 = src[m - 1 - i], a1 = src[i2], b1 = src[hm + i];
            const bool t0 = node_is_dct4(d, node);
            float lo0, hi0, lo1, hi1;
            fly(t0, hm, i, a0, b0, lo0, hi0);                         // level m, pair i      -> mid[i], mid[hm+i]
            fly(t0, hm, i2, a1, b1, lo1, hi1);                        // level m, pair hm-1-i -> mid[hm-1-i], mid[m-1-i]
            float o0, o1, o2, o3;
            fly(node_is_dct4(d + 1, 2 * node), qm, i, lo0, lo1, o0, o1);          // lower child, pair i
            fly(node_is_dct4(d + 1, 2 * node + 1), qm, i, hi0, hi1, o2, o3);      // upper child, pair i
            dst[i] = o0; dst[qm + i] = o1; dst[hm + i] = o2; dst[hm + qm + i] = o3;
        }
    }
    __device__ void dct_up_long(int m, int lgm, bool srcIsX) {
        const int hm = m >> 1, lgH = tb.log2N - 1, d = tb.log2N - lgm;
        const float *const S0 = srcIsX ? w.XA : w.TA, *const S1 = srcIsX ? w.XB : w.TB;
        float *const D0 = srcIsX ? w.TA : w.XA, *const D1 = srcIsX ? w.TB : w.XB;
#pragma unroll 4
        for (int P = tid; P < CN; P += NTHR) {
            const int cc = P >> lgH, r = P & (H - 1);
            const bool hb = cc >= nch();
            const int node = r >> (lgm - 1), i = r & (hm - 1);
            const int off = (hb ? cc - nch() : cc) * N + (node << lgm);
            const float *src = (hb ? S1 : S0) + off;
            float *dst = (hb ? D1 : D0) + off;
            if (node_is_dct4(d, node)) {                              // Fourier_DCT4_Template.h:100-114
                if (i == hm - 1) { dst[0] = src[0]; dst[m - 1] = src[hm]; }
                else {
                    float a = src[i + 1], b = src[m - 1 - i];
                    dst[2 * i + 1] = a + b;
                    dst[2 * i + 2] = a - b;
                }
            } else {                                                  // Fourier_DCT2_Template.h:62-77
                dst[2 * i] = src[i];
                dst[2 * i + 1] = src[hm + i];
            }
        }
    }
    // Two levels up in one pass: node size m from its four grandchildren (size m/4).  A DCT-IV node has children (II, II), a
    // DCT-II node (II, IV); a DCT-II result is the interleave of its children, a DCT-IV result their butterfly
    // (Fourier_DCT4_Template.h:100-114, Fourier_DCT2_Template.h:62-77), so four consecutive outputs 4k..4k+3 depend on six
    // grandchild values.  Same additions on the same operands as two single-level passes; half the barriers, one 128-bit store.
    // Reads one buffer pair, writes the other (the grandchildren share the parity of the node's own depth).
    __device__ void dct_up2_long(int lgm, bool srcIsX) {
        const int m = 1 << lgm, hm = m >> 1, qm = m >> 2, lgH = tb.log2N - 1, d = tb.log2N - lgm;
        const float *const S0 = srcIsX ? w.XA : w.TA, *const S1 = srcIsX ? w.XB : w.TB;
        float *const D0 = srcIsX ? w.TA : w.XA, *const D1 = srcIsX ? w.TB : w.XB;
        if (((H >> 1) % NTHR) == 0) {                                  // as in dct_down2_long: node and position are the thread's own
            for (int r = tid; r < (H >> 1); r += NTHR) {
                const int node = r >> (lgm - 2), k = r & (qm - 1);
                const bool t0 = node_is_dct4(d, node), first = k == 0, last = k == qm - 1;
                const int offn = node << lgm;
#pragma unroll 2
                for (int cc = 0; cc < 2 * nch(); cc++) {
                    const bool hb = cc >= nch();
                    const int off = (hb ? cc - nch() : cc) * N + offn;
                    const float *AA = (hb ? S1 : S0) + off, *AB = AA + qm, *BA = AA + hm, *BB = BA + qm;
                    float4 o;
                    const float ab = AB[k], bbr = BB[qm - 1 - k];
                    if (t0) {
                        o.x = first ? AA[0] : AA[k] - BA[qm - k];
                        o.y = ab + bbr;
                        o.z = ab - bbr;
                        o.w = last ? BA[0] : AA[k + 1] + BA[qm - 1 - k];
                    } else {
                        o.x = AA[k];
                        o.y = first ? BA[0] : BA[k] - BB[qm - k];
                        o.z = ab;
                        o.w = last ? BB[0] : BA[k + 1] + bbr;
                    }
                    *reinterpret_cast<float4 *>((hb ? D1 : D0) + off + 4 * k) = o;
                }
            }
            return;
        }
#pragma unroll 2
        for (int Q = tid; Q < (CN >> 1); Q += NTHR) {
            const int cc = Q >> (lgH - 1), r = Q & ((H >> 1) - 1);
            const bool hb = cc >= nch();
            const int node = r >> (lgm - 2), k = r & (qm - 1);
            const int off = (hb ? cc - nch() : cc) * N + (node << lgm);
            const float *AA = (hb ? S1 : S0) + off, *AB = AA + qm, *BA = AA + hm, *BB = BA + qm;
            float4 o;
            const float ab = AB[k], bbr = BB[qm - 1 - k];
            if (node_is_dct4(d, node)) {
                o.x = (k == 0) ? AA[0] : AA[k] - BA[qm - k];
                o.y = ab + bbr;
                o.z = ab - bbr;
                o.w = (k == qm - 1) ? BA[0] : AA[k + 1] + BA[qm - 1 - k];
            } else {
                o.x = AA[k];
                o.y = (k == 0) ? BA[0] : BA[k] - BB[qm - k];
                o.z = ab;
                o.w = (k == qm - 1) ? BB[0] : BA[k + 1] + bbr;
            }
            *reinterpret_cast<float4 *>((hb ? D1 : D0) + off + 4 * k) = o;
        }
    }
    __device__ void transform() {
        const int lgN = tb.log2N;
        if (sm.nsub == 1) {
            // Two recursion levels per pass, down to the 8-point leaves and up again; an odd number of levels leaves one single
            // pass at node size 16 either side of the leaves.  Every pass reads one buffer pair and writes the other; the number
            // of passes is even, so the result is back in X.
            bool srcIsX = true;
            int lgm = lgN;
            for (; lgm >= 5; lgm -= 2) { dct_down2_long(lgm, srcIsX); wsync<NTHR>(); srcIsX = !srcIsX; }
            if (lgm == 4) { dct_down_long(16, 4, srcIsX); wsync<NTHR>(); srcIsX = !srcIsX; }
            dct_leaves(srcIsX ? 1 : 0);
            wsync<NTHR>();
            if (lgm == 4) { dct_up_long(16, 4, srcIsX); wsync<NTHR>(); srcIsX = !srcIsX; lgm = 6; } else lgm = 5;
            for (; lgm <= lgN; lgm += 2) { dct_up2_long(lgm, srcIsX); wsync<NTHR>(); srcIsX = !srcIsX; }
            return;
        }
        for (int lgm = lgN; lgm >= 4; lgm--) { dct_down(1 << lgm, lgm); wsync<NTHR>(); }
        dct_leaves();
        wsync<NTHR>();
        for (int lgm = 4; lgm <= lgN; lgm++) { dct_up(1 << lgm, lgm); wsync<NTHR>(); }
    }

    // ---------------- normalise, keys, line sums (ulcEncoder_BlockTransform.c:243-281) ----------------
    __device__ void analyse(int k) {
        int nz = 0;
        for (int p = tid; p < H; p += NTHR) {
            const int j = find_sub(sm, 2 * p);
            const SubInfo s = sm.sub[j];
            const float norm = 2.0f / (float)s.S;
            const int n0 = 2 * p - s.off;                             // even local index
            float amp = 0.0f;
#pragma unroll
            for (int c = 0; c < nch(); c++) {
                const int base = c * N + s.off;
                float re0 = w.XA[base + n0] * norm, re1 = w.XA[base + n0 + 1] * norm;
                float im0 = w.XB[base + s.S - 1 - n0] * norm, im1 = w.XB[base + s.S - 2 - n0] * norm;   // MDST is stored reversed (:184-186)
                w.XA[base + n0] = re0; w.XA[base + n0 + 1] = re1;
                float a0 = re0 * re0 + im0 * im0, a1 = re1 * re1 + im1 * im1;
                float k0, k1;
                if (fabsf(re0) < 0.5f * 0x1.0p-31f) k0 = -INFINITY; else { k0 = fast_log(re0 * re0); nz++; }
                if (fabsf(re1) < 0.5f * 0x1.0p-31f) k1 = -INFINITY; else { k1 = fast_log(re1 * re1); nz++; }
                w.TA[base + n0] = k0; w.TA[base + n0 + 1] = k1;
                float ns = a0 + a1;
                amp += a0; amp += a1;
                w.TB[base + (n0 >> 1)] = ns;
                lg_lines(k, c, s)[n0 >> 1] = fast_log(0x1.0p-126f + ns);
            }
            w.amp2[p] = amp;
            w.mask[p] = fast_log(0x1.0p-126f + amp);
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, d);
        if ((tid & 31) == 0 && nz) atomicAdd(&sm.nNz, nz);
    }

    // ---------------- serial sums: complexity + Bark band analysis ----------------
    // Bark analysis of one unit by one warp.  The three running f64 sums (struct LineSum_t, ulcEncoder_Psyopt.c:16-51)
    // must be accumulated in index order; lanes 0,1,2 each own one sum, so one DADD per line serves all three, and the
    // per-line conversions / products are done 32 lines at a time by the whole warp.  The 50 band edges are walked
    // from a host-sorted list; the 25 band levels are evaluated in parallel afterwards (one lane per band).
    __device__ void band_unit(int unit, const float *v, const float *lg, int n, int sh, int kind) {
        const int lane = tid & 31;
        const int16_t *ev = tb.events + ((sh * 2 + kind) * 2 * ULCB_NBARK) * 2;
        const int16_t *ed = tb.edges + ((sh * 2 + kind) * ULCB_NBARK) * 2;
        double (*lo)[3] = w.lo[unit];
        if (!args.exactBands) {
            // Segmented sums.  The reference adds the lines of a unit one by one in f64 (LineSum_t); a dependent f64 add costs
            // ~30 cycles here and there are n of them.  This path gives every lane a contiguous run of lines, scans the lane
            // totals and takes the band sums as differences of the prefix sums at the band edges: the same sums, associated
            // differently - the band sums differ from the reference's in the last bits of an f64, the band levels (rounded to
            // f32 from there) practically never (conformance).  Opt-in (ULCB_ENC_FAST_BANDS=1); the default is the
            // sequential chains below.
            double (*hiS)[3] = reinterpret_cast<double (*)[3]>(w.stage[unit]);   // prefix sums at the band ends (lo takes those at the starts)
            const int len = (n + 31) >> 5, beg = lane * len, end = (beg + len < n) ? beg + len : n;
            int e = 0;
            {                                                             // first edge at or after this lane's first line
                int lo_ = 0, hi_ = 2 * ULCB_NBARK;
                while (lo_ < hi_) { const int mid = (lo_ + hi_) >> 1; if (__ldg(&ev[2 * mid]) < beg) lo_ = mid + 1; else hi_ = mid; }
                e = lo_;
            }
            const int e0 = e;
            int pe = e < 2 * ULCB_NBARK ? __ldg(&ev[2 * e]) : 0x7fffffff;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            auto edges_at = [&](int t) {
                while (pe == t) {
                    const int code = __ldg(&ev[2 * e + 1]);
                    double *dst = (code & 0x80) ? hiS[code & 0x7F] : lo[code & 0x7F];
                    dst[0] = a0; dst[1] = a1; dst[2] = a2;                // relative to the lane's first line; the base comes below
                    e++;
                    pe = e < 2 * ULCB_NBARK ? __ldg(&ev[2 * e]) : 0x7fffffff;
                }
            };
            if (beg < n) {
                for (int t = beg; t < end; t++) {
                    edges_at(t);
                    const double X = (double)v[t], L = (double)lg[t];
                    a0 += L; a1 += L * X; a2 += X;
                }
                if (end == n) edges_at(n);                                // the line past the end sees the totals
            }
            const int e1 = e;
            double b0 = a0, b1 = a1, b2 = a2;                             // inclusive scan over the lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double u0 = __shfl_up_sync(0xffffffffu, b0, d), u1 = __shfl_up_sync(0xffffffffu, b1, d), u2 = __shfl_up_sync(0xffffffffu, b2, d);
                if (lane >= d) { b0 += u0; b1 += u1; b2 += u2; }
            }
            b0 = __shfl_up_sync(0xffffffffu, b0, 1); b1 = __shfl_up_sync(0xffffffffu, b1, 1); b2 = __shfl_up_sync(0xffffffffu, b2, 1);
            if (lane == 0) { b0 = b1 = b2 = 0.0; }                        // sum of the lanes before this one
            for (int q = e0; q < e1; q++) {
                const int code = __ldg(&ev[2 * q + 1]);
                double *dst = (code & 0x80) ? hiS[code & 0x7F] : lo[code & 0x7F];
                dst[0] += b0; dst[1] += b1; dst[2] += b2;
            }
            __syncwarp();
            if (lane < ULCB_NBARK) {
#pragma unroll
                for (int q = 0; q < 3; q++) lo[lane][q] = hiS[lane][q] - lo[lane][q];
            }
            __syncwarp();
        } else {
        double (*stage)[32] = w.stage[unit];
        const int src = lane < 3 ? lane : 0;
        double acc = 0.0;
        // The 2*NBARK band-edge events (position-sorted) live two per lane.  Lanes 0..2 run the three dependent DADD
        // chains over a group of 32 lines and leave the running sum BEFORE each line in `stage`; the lanes whose
        // events fall inside the group then pick their snapshots up in parallel (band starts first, then band ends,
        // which is the list order for a band that starts and ends on the same line).
        int p0 = 0x7fffffff, c0 = 0, p1 = 0x7fffffff, c1 = 0;
        if (lane < 2 * ULCB_NBARK) { p0 = __ldg(&ev[2 * lane]); c0 = __ldg(&ev[2 * lane + 1]); }
        if (lane + 32 < 2 * ULCB_NBARK) { p1 = __ldg(&ev[2 * (lane + 32)]); c1 = __ldg(&ev[2 * (lane + 32) + 1]); }
        auto apply = [&](int pos, int code, int base, bool ends) {
            if (pos >= base && pos < base + 32 && ((code & 0x80) != 0) == ends) {
                const int t = pos - base, b = code & 0x7F;
#pragma unroll
                for (int q = 0; q < 3; q++) { const double val = stage[q][t]; lo[b][q] = ends ? val - lo[b][q] : val; }
            }
        };
        auto events_in = [&](int base) {
            const bool mine = (p0 >= base && p0 < base + 32) || (p1 >= base && p1 < base + 32);
            if (__any_sync(0xffffffffu, mine)) {
                apply(p0, c0, base, false); apply(p1, c1, base, false);
                __syncwarp();
                apply(p0, c0, base, true); apply(p1, c1, base, true);
            }
        };
        float xN = v[lane], lN = lg[lane];
        for (int base = 0; base < n; base += 32) {
            const double X = (double)xN, L = (double)lN;
            if (base + 32 < n) { xN = v[base + 32 + lane]; lN = lg[base + 32 + lane]; }
            __syncwarp();
            stage[0][lane] = L; stage[1][lane] = L * X; stage[2][lane] = X;
            __syncwarp();
            if (lane < 3) {                                               // the other lanes have no chain (and must not touch the stage)
#pragma unroll
                for (int t0 = 0; t0 < 32; t0 += 8) {
                    double a[8];
#pragma unroll
                    for (int t = 0; t < 8; t++) a[t] = stage[src][t0 + t];
#pragma unroll
                    for (int t = 0; t < 8; t++) { stage[src][t0 + t] = acc; acc += a[t]; }
                }
            }
            __syncwarp();
            events_in(base);
        }
        __syncwarp();
        if (lane < 3) stage[src][0] = acc;                                // events on the line past the end see the totals
        __syncwarp();
        events_in(n);
        __syncwarp();
        }
        // band levels: lanes 0..24; a band without energy inherits the previous band's level
        float level = 0.0f; bool ok = false;
        if (lane < ULCB_NBARK) {
            const double dLg0 = lo[lane][0], dWlg0 = lo[lane][1], dW = lo[lane][2];
            if (dW > 0.0) {
                ok = true;
                const int beg = __ldg(&ed[2 * lane]), end = __ldg(&ed[2 * lane + 1]);
                if (kind == 0) {                                      // noise level (ulcEncoder_Psyopt.c:206-222)
                    double inv = 1.0 / (double)(end - beg);
                    double dWlg = dWlg0 / dW;
                    double dLg = dLg0 * inv;
                    level = 0.5f * (float)(log(dW * inv) + dLg - dWlg);
                } else {                                              // masking (ulcEncoder_Psyopt.c:117-131)
                    double dWlg = dWlg0 / dW;
                    double dLg = dLg0 / (double)(end - beg);
                    level = (float)(dWlg - dLg - log(dW));
                }
            }
        }
        const unsigned okMask = __ballot_sync(0xffffffffu, ok);
        const unsigned below = okMask & ((2u << lane) - 1u);          // valid bands at or below this one
        const int from = below ? 31 - __clz(below) : 0;
        const float got = __shfl_sync(0xffffffffu, level, from);
        if (lane < ULCB_NBARK) sm.band[unit][lane] = below ? got : (kind ? 0.0f : -100.0f);
    }
    __device__ void serial_sums(int k) {
        const int U = nch() * sm.nsub, nTasks = U + sm.nsub + 1;
        constexpr int NW = NTHR / 32;
        const int warp = tid >> 5, lane = tid & 31;
        for (int task = warp; task < nTasks; task += NW) {
            if (task == nTasks - 1) {                                 // block complexity (ulcEncoder_BlockTransform.c:310-325)
                if (lane == 0 && !(args.dbgFlags & 1)) {
                    float sq = 0.0f, ab = 0.0f;
                    const float4 *x4 = reinterpret_cast<const float4 *>(w.XA);
#pragma unroll 8
                    for (int i = 0; i < (CN >> 2); i++) {
                        float4 r = x4[i];
                        float a0 = r.x * r.x, a1 = r.y * r.y, a2 = r.z * r.z, a3 = r.w * r.w;
                        sq += a0; sq += a1; sq += a2; sq += a3;
                        ab += fabsf(r.x); ab += fabsf(r.y); ab += fabsf(r.z); ab += fabsf(r.w);
                    }
                    float cx = sq;
                    if (cx != 0.0f) {
                        float scale = kLn2f * (float)tb.log2N;
                        cx = logf_ref(ab * ab / sq) / scale;
                        if (cx < 0.0f) cx = 0.0f;
                        if (cx > 1.0f) cx = 1.0f;
                    }
                    sm.cx = cx;
                }
            } else if (args.mode == ULCB_ENC_ANALYSE || (args.dbgFlags & 2) || ((args.dbgFlags & 4) && task != U)) {
            } else if (task < U) {
                int c = task / sm.nsub, j = task - c * sm.nsub;
                const SubInfo &s = sm.sub[j];
                band_unit(task, w.TB + c * N + s.off, lg_lines(k, c, s), s.S >> 1, s.sh, 0);
            } else {
                int j = task - U;
                const SubInfo &s = sm.sub[j];
                band_unit(task, w.amp2 + (s.off >> 1), w.mask + (s.off >> 1), s.S >> 1, s.sh, 1);
            }
        }
    }
    __device__ __forceinline__ float bark_interp(const float *band, int sh, int line) const {    // ulcEncoder_Psyopt.c:136-152
        const int t = tb.lineOff[sh] + line;
        const int idx = __ldg(&tb.lineBand[t]);
        const float fr = __ldg(&tb.lineFrac[t]);
        float l = (idx < ULCB_NBARK) ? band[idx] : band[ULCB_NBARK - 1];
        float r = (idx + 1 < ULCB_NBARK) ? band[idx + 1] : l;
        return l * (1.0f - fr) + r * fr;
    }
    __device__ void spread() {
        const int U = nch() * sm.nsub;
        for (int p = tid; p < H; p += NTHR) {
            const int j = find_sub(sm, 2 * p);
            const SubInfo &s = sm.sub[j];
            const int line = p - (s.off >> 1);
#pragma unroll
            for (int c = 0; c < nch(); c++) {
                float v = bark_interp(sm.band[c * sm.nsub + j], s.sh, line);
                float wgt = expf_ref(0.5f * v);
                float *dst = w.TB + c * N + s.off + 2 * line;
                dst[0] = wgt;
                dst[1] = wgt * (v + kLn2f);
            }
            w.mask[p] = bark_interp(sm.band[U + j], s.sh, line);
        }
    }

    // ---------------- final keys + ranks ----------------
    __device__ void make_keys(float *dbgKeys) {
        uint32_t *uk = reinterpret_cast<uint32_t *>(w.TA);
        uint16_t *idxA = reinterpret_cast<uint16_t *>(w.XB);
        for (int i = tid; i < CN; i += NTHR) {
            const int c = i >> tb.log2N, n = i & (N - 1);
            float key = w.TA[i];
            key = 2.0f * key + w.mask[n >> 1] + -0x1.62E430p0f * (float)(c & 1);     // ulcEncoder_BlockTransform.c:337-345
            if (dbgKeys) dbgKeys[i] = key;
            uint32_t b = __float_as_uint(key);
            b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);          // ascending order-preserving
            uk[i] = ~b;                                               // descending importance
            idxA[i] = (uint16_t)i;
        }
    }
    __device__ void rank_sort() {
        // LSD radix sort, 4 bits per pass, stable: equal keys stay in index order
        const uint32_t *uk = reinterpret_cast<const uint32_t *>(w.TA);
        uint16_t *src = reinterpret_cast<uint16_t *>(w.XB), *dst = src + CN;
        const int E = slot_items(), Ev = tid * E < CN ? E : 0;
        for (int shift = 0; shift < 32; shift += 4) {
#pragma unroll
            for (int d = 0; d < 16; d++) w.cnt[d * NTHR + tid] = 0;
            for (int e = 0; e < Ev; e++) {
                int dg = (uk[src[tid * E + e]] >> shift) & 15;
                w.cnt[dg * NTHR + tid]++;
            }
            wsync<NTHR>();
            int loc = 0;
#pragma unroll
            for (int q = 0; q < 16; q++) loc += w.cnt[tid * 16 + q];
            int tot, base = block_excl_sum<NTHR>(loc, sm.scan, scanFlip, tot);
#pragma unroll
            for (int q = 0; q < 16; q++) { int t = w.cnt[tid * 16 + q]; w.cnt[tid * 16 + q] = (uint16_t)base; base += t; }
            wsync<NTHR>();
            for (int e = 0; e < Ev; e++) {
                uint16_t ix = src[tid * E + e];
                int dg = (uk[ix] >> shift) & 15;
                int pos = w.cnt[dg * NTHR + tid]++;
                dst[pos] = ix;
            }
            wsync<NTHR>();
            uint16_t *t = src; src = dst; dst = t;
        }
        // 8 passes: the sorted order is back in the first array.  Turn order into ranks (uint16 at TA).
        wsync<NTHR>();
        uint16_t *rank = reinterpret_cast<uint16_t *>(w.TA);
        // keys are dead now; but rank[] aliases them, so stage through registers
        const uint16_t *ord = reinterpret_cast<const uint16_t *>(w.XB);
        wsync<NTHR>();
        for (int r = tid; r < CN; r += NTHR) rank[ord[r]] = (uint16_t)r;
        wsync<NTHR>();
    }

    // VBR needs only the predicate "rank < nOut" (encode_pass, :236-268 of ulcEncoder_Encode.c scan coefficients whose
    // TransformIndex is below the cut), and nOut is known before the ranks are: an MSB-first radix SELECT of the nOut-th
    // key (4 histogram passes of 8 bits, no scatter) replaces the 8-pass sort.  Keys equal to the threshold are taken in
    // index order, which is the order the stable sort gives them.  Writes rank = 0 (kept) / 0xFFFF (dropped).
    static constexpr int MAXE = 16;
    __device__ bool can_select() const { return slot_items() <= MAXE; }
    __device__ void rank_select(int nOut) {
        const uint32_t *uk = reinterpret_cast<const uint32_t *>(w.TA);
        const int i0 = tid * slot_items(), E = i0 < CN ? slot_items() : 0;      // E = this thread's keys (all or none)
        uint32_t key[MAXE];
#pragma unroll
        for (int e = 0; e < MAXE; e++) key[e] = e < E ? uk[i0 + e] : 0xffffffffu;
        int *hist = reinterpret_cast<int *>(w.cnt);                     // 4 passes x 256 bins, cleared together
        uint32_t prefix = 0, pmask = 0;
        int need = nOut;                                                 // keys still to take among those matching the prefix
        if (nOut > 0) {
            for (int i = tid; i < 4 * 256; i += NTHR) hist[i] = 0;
            wsync<NTHR>();
            const int lane = tid & 31;
            for (int shift = 24; shift >= 0; shift -= 8, hist += 256) {
#pragma unroll
                for (int e = 0; e < MAXE; e++)
                    if (e < E && (key[e] & pmask) == prefix) atomicAdd(&hist[(key[e] >> shift) & 255], 1);
                wsync<NTHR>();
                // every warp finds the bin that holds the need-th key for itself: no hand-off, one barrier per pass
                int c[8], sum = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) { c[q] = hist[lane * 8 + q]; sum += c[q]; }
                int inc = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
                const int exc = inc - sum;
                const bool hit = exc < need && need <= inc;             // exactly one lane
                int run = exc, bin = lane * 8 + 7;
                if (hit) for (int q = 0; q < 8; q++) { if (need <= run + c[q]) { bin = lane * 8 + q; break; } run += c[q]; }
                const unsigned who = __ballot_sync(0xffffffffu, hit);
                const int srcLane = who ? __ffs(who) - 1 : 31;
                bin = __shfl_sync(0xffffffffu, bin, srcLane);
                run = __shfl_sync(0xffffffffu, run, srcLane);
                need -= run;
                prefix |= (uint32_t)bin << shift;
                pmask |= 0xffu << shift;
            }
        }
        // prefix = the nOut-th smallest key; `need` of the keys equal to it are kept, first in index order
        int eq = 0;
#pragma unroll
        for (int e = 0; e < MAXE; e++) eq += (e < E && key[e] == prefix) ? 1 : 0;
        int tot, base = block_excl_sum<NTHR>(eq, sm.scan, scanFlip, tot);
        wsync<NTHR>();                                                   // every key is in registers: TA may be overwritten
        uint16_t *rank = reinterpret_cast<uint16_t *>(w.TA);
#pragma unroll
        for (int e = 0; e < MAXE; e++) {
            if (e < E) {
                bool keep = false;
                if (nOut > 0) {
                    if (key[e] < prefix) keep = true;
                    else if (key[e] == prefix) { keep = base < need; base++; }
                }
                rank[i0 + e] = keep ? (uint16_t)0 : (uint16_t)0xFFFF;
            }
        }
        wsync<NTHR>();
    }

    // ---------------- one coder pass (ulcEncoder_Encode.c:319-360); returns the size in bits ----------------
    __device__ int encode_pass(int k, int nOut, bool write) {
        const uint16_t *rank = reinterpret_cast<const uint16_t *>(w.TA);
        uint16_t *selList = reinterpret_cast<uint16_t *>(w.TA) + CN;
        uint8_t  *info = reinterpret_cast<uint8_t *>(w.XB);                       // [CN]
        uint16_t *zoneStart = reinterpret_cast<uint16_t *>(info + CN);           // [CN]
        uint8_t  *zoneQi = info + 3 * CN;                                         // [CN]
        uint32_t *outW = out_words(k);                                            // packed nybbles
        uint16_t *jump = reinterpret_cast<uint16_t *>(w.encScratch);             // [CN] next zone start if a zone starts here
        uint8_t  *qcand = w.encScratch + 2 * CN;                                  // [CN] quantiser of that zone
        const float *coef = w.XA, *pairs = w.TB;
        const int Eall = slot_items(), U = nch() * sm.nsub, i0 = tid * Eall, E = i0 < CN ? Eall : 0;   // E = this thread's slots (all or none)
        const int capNyb = 8 * ((CN >> 1) + 4);                                   // 2*CN + 16 bytes

#ifdef ULCB_PHASE_TIMING
        long long et0 = clock64();
#define ETK(i) do { if (tid == 0) { long long t_ = clock64(); atomicAdd(&g_ulcbPhase[i], (unsigned long long)(t_ - et0)); et0 = t_; } } while (0)
#else
#define ETK(i) do { } while (0)
#endif
        // G1: selection, compaction in index order
        int cntSel = 0;
        for (int e = 0; e < E; e++) { info[i0 + e] = 0; cntSel += (int)rank[i0 + e] < nOut; }
        int nSel, base = block_excl_sum<NTHR>(cntSel, sm.scan, scanFlip, nSel);
        {
            int b = base;
            if (E) {                                                  // a unit can only start at a thread's first slot (E divides every sub-block size)
                const int c = i0 >> tb.log2N, n = i0 & (N - 1);
                for (int j = 0; j < sm.nsub; j++) if (n == sm.sub[j].off) sm.unitSel[c * sm.nsub + j] = b;
            }
            for (int e = 0; e < E; e++) {
                const int i = i0 + e;
                if ((int)rank[i] < nOut) selList[b++] = (uint16_t)i;
            }
            if (tid == 0) sm.unitSel[U] = nSel;
        }
        if (write) for (int i = tid; i < (CN >> 1) + 4; i += NTHR) outW[i] = 0u;
        wsync<NTHR>();

        ETK(22);
        // G2: greedy quantiser zones (ulcEncoder_Encode.c:200-269).  A zone that starts at selected coefficient m runs
        // until the first later one that would push max/min of the levels above 4; that end only depends on m, so
        // every selected coefficient computes "the zone that would start here" in parallel (jump pointer + quantiser),
        // and one thread per (channel, sub-block) then just follows the pointers from the first selected coefficient.
        // The scan for a zone's end is monotone (the running max only grows, the running min only shrinks), so whole groups of
        // eight selected coefficients are skipped on their {min, max} summary while the zone stays open with the group included;
        // the group that closes it is then scanned one by one.  (The summaries borrow the zone tables, written after this loop.)
        float2 *gsum = reinterpret_cast<float2 *>(zoneStart);
        for (int g = tid; g * 8 + 8 <= nSel; g += NTHR) {
            float mn = fabsf(coef[selList[g * 8]]), mx = mn;
#pragma unroll
            for (int q = 1; q < 8; q++) { const float l = fabsf(coef[selList[g * 8 + q]]); mn = l < mn ? l : mn; mx = l > mx ? l : mx; }
            gsum[g] = make_float2(mn, mx);
        }
        wsync<NTHR>();
        for (int m = tid; m < nSel; m += NTHR) {
            int u = 0;
            while (m >= sm.unitSel[u + 1]) u++;
            const int m1 = sm.unitSel[u + 1];
            float lo = fabsf(coef[selList[m]]), hi = lo;
            int mm = m + 1;
            bool open = true;
            auto one = [&]() {                                        // try to take selected coefficient mm into the zone
                const float level = fabsf(coef[selList[mm]]);
                const float newLo = level < lo ? level : lo, newHi = level > hi ? level : hi;
                if (newHi > newLo * 4.0f) open = false;
                else { lo = newLo; hi = newHi; mm++; }
            };
            while (open && mm < m1 && (mm & 7)) one();
            while (open && mm + 8 <= m1) {
                const float2 g = gsum[mm >> 3];
                const float newLo = g.x < lo ? g.x : lo, newHi = g.y > hi ? g.y : hi;
                if (newHi > newLo * 4.0f) break;
                lo = newLo; hi = newHi; mm += 8;
            }
            while (open && mm < m1) one();
            int q = (int)(0x1.657006p2f + -0x1.715476p0f * logf_ref(hi));     // :83-88
            if (q < 5) q = 5;
            if (q > 31) q = 31;
            jump[m] = (uint16_t)mm;
            qcand[m] = (uint8_t)q;
        }
        wsync<NTHR>();
        ETK(23);
        if (tid < U) {
            const int m0 = sm.unitSel[tid], m1 = sm.unitSel[tid + 1];
            int z = 0, prevQ = -1;
            for (int m = m0; m < m1; m = jump[m]) {
                prevQ = qcand[m];
                zoneStart[m0 + z] = (uint16_t)m;
                zoneQi[m0 + z] = (uint8_t)prevQ;
                z++;
            }
            sm.unitZones[tid] = z;
            sm.unitLastQ[tid] = prevQ;
        }
        wsync<NTHR>();

        ETK(24);
        // G2b: per selected coefficient, its zone's quantiser and header flags
        for (int m = tid; m < nSel; m += NTHR) {
            int u = 0;
            while (m >= sm.unitSel[u + 1]) u++;
            const int zb = sm.unitSel[u];
            int loz = 0, hiz = sm.unitZones[u] - 1;                   // last zone whose start <= m
            while (loz < hiz) { int mid = (loz + hiz + 1) >> 1; if ((int)zoneStart[zb + mid] <= m) loz = mid; else hiz = mid - 1; }
            int q = zoneQi[zb + loz];
            int inf = q;
            if ((int)zoneStart[zb + loz] == m) {
                if (loz == 0) inf |= INF_HDR | INF_FIRST;
                else if (zoneQi[zb + loz - 1] != q) inf |= INF_HDR;
            }
            info[selList[m]] = (uint8_t)inf;
        }
        wsync<NTHR>();

        ETK(25);
        // G3: kept coefficients and the last kept index before each thread's slots
        uint32_t keptMask = 0; int lastKeptP1 = 0;
        for (int e = 0; e < E; e++) {
            const int i = i0 + e, inf = info[i];
            if (inf) {
                float sc = coef[i] * (float)(1u << (inf & INF_Q));
                if (!(fabsf(sc) < 2.5f)) { keptMask |= 1u << e; lastKeptP1 = i + 1; }   // :114
            }
        }
        const int prevKeptP1 = block_excl_max<NTHR>(lastKeptP1, sm.scan, scanFlip);

        // unit of this thread's slots (a thread's slots never straddle a sub-block: E divides every sub-block size)
        const int c = i0 >> tb.log2N, nloc = i0 & (N - 1), j = find_sub(sm, nloc), u = c * sm.nsub + j;
        const int uOff = c * N + sm.sub[j].off, uEnd = uOff + sm.sub[j].S;
        const bool ownsTail = E && (i0 + E == uEnd);

        auto walk = [&](auto &sink, bool cacheTail) {
            int nc = prevKeptP1 > uOff ? prevKeptP1 : uOff;
            for (int e = 0; e < E; e++) {
                const int i = i0 + e, inf = info[i];
                if (inf & INF_HDR) put_quant(sink, inf);
                if (keptMask & (1u << e)) { put_item(sink, coef, pairs, nc, i, (float)(1u << (inf & INF_Q))); nc = i + 1; }
            }
            if (ownsTail) {                                           // :273-312
                if (cacheTail) {
                    uint32_t tok = 0; int len = 0;
                    auto tput = [&](int x) { tok |= (uint32_t)(x & 15) << (4 * len); len++; };
                    const int n = uEnd - nc, pq = sm.unitLastQ[u];
                    if (n > 4) {
                        if (pq != -1) tput(0xF);
                        int nq = 0, dq = 0;
                        if (pq != -1 && n >= 16) hf_ext_params(pairs, nc, n, (float)(1u << pq), nq, dq);
                        if (nq) { tput(0xF); tput(nq - 1); tput(dq >> 4); tput(dq); }
                        else { tput(0xE); tput(0xF); }
                    } else if (n > 0) { tput(0x0); tput(n - 1); }
                    sm.tailTok[u] = tok; sm.tailLen[u] = len;
                }
                const uint32_t tok = sm.tailTok[u]; const int len = sm.tailLen[u];
                for (int q = 0; q < len; q++) sink.put((tok >> (4 * q)) & 15);
            }
        };

        ETK(26);
        // write passes stage their nybbles (the zone scratch is dead by now: 32 bytes per thread)
        uint32_t *stageSlot = reinterpret_cast<uint32_t *>(w.encScratch) + tid * 8;
        CountSink cs{0};
        if (write) { StageSink ss{stageSlot}; walk(ss, true); ss.finish(); cs.n = ss.n; }
        else walk(cs, true);
        __syncwarp();
        ETK(27);
        int total, offs = block_excl_sum<NTHR>(cs.n, sm.scan, scanFlip, total);
        ETK(28);
        const int hdr = (sm.wcCur & 8) ? 2 : 1;
        if (write) {
            if (cs.n <= 64) {                                         // shifted copy of the staged nybbles to hdr + offs
                const int p0 = hdr + offs, sh = (p0 & 7) * 4, nW = (cs.n + 7) >> 3;
                uint32_t prev = 0;
                for (int q = 0; q <= nW; q++) {
                    const uint32_t cur = q < nW ? stageSlot[q] : 0u;
                    const uint32_t v = sh ? ((cur << sh) | (prev >> (32 - sh))) : cur;
                    const int wi = (p0 >> 3) + q;
                    if (v && wi * 8 < capNyb) atomicOr(&outW[wi], v);
                    prev = cur;
                }
            } else {
                WriteSink ws{outW, hdr + offs, capNyb};
                walk(ws, false);
                ws.finish();
            }
            if (tid == 0) { WriteSink h{outW, 0, capNyb}; h.put(sm.wcCur); if (sm.wcCur & 8) h.put(sm.wcCur >> 4); h.finish(); }
        }
        wsync<NTHR>();
        ETK(29);
        return ((hdr + total) * 4 + 7) & ~7;
    }
};


// ---- transient detector warp (ulcEncoder_WindowControl.c:31-239) -----------------------------------------
// Runs on its own warp, one block ahead of the workers.  It reads its analysis window (second half of block
// b-1, first half of block b, one sample either side) straight from the PCM input, so it shares nothing with
// the workers but the published decision.  The tap energies are computed by the 32 lanes; the three one-pole
// envelope recurrences are serial by definition and run on lane 0 in the reference's operation order.
template <int NCHT>
struct Detector {
    static constexpr int MAXCH = NCHT ? NCHT : ULCB_ENC_MAXCH;
    const UlcbEncodeArgs &args;
    float2 *E;                    // [3][N] private scratch: the block buffers of the three pipelined passes
    float *stg;                   // [nch()][N+2] analysis window, M/S
    float *seg;                   // [32] {sum,cnt} x 16 (shared memory)
    float env[3];
    const unsigned char *pcm; long long nSamples; const float *prevBlk;   // prevBlk: block before this call (M/S planar) or nullptr

    // One iteration of the software-pipelined detector.  The three envelope passes of a block depend on each other
    // (post-mask forwards -> pre-mask backwards from the post-mask end value -> block-mask forwards), but the passes of
    // DIFFERENT blocks do not: iteration `it` runs post(it), pre(it-1) and blk(it-2) as three independent dependency
    // chains in ONE loop on lane 0, so the recurrences' add/mul latencies overlap instead of adding up.  Each pass keeps
    // the reference's operation order, so the decisions are unchanged.  E holds three block buffers that rotate through
    // the roles post (in place) -> pre (in place) -> blk (read).  Returns the decision for block it-2 (it >= 2).
    float preHp = 0.0f, preBp = 0.0f;          // post-mask end values of the block whose pre-mask pass comes next
    int nchRt = NCHT;
    __device__ __forceinline__ int nch() const { return NCHT ? NCHT : nchRt; }
    __device__ int iterate(int it, int nBlocks) {
        const int N = args.tb.N, H = N >> 1, lane = threadIdx.x & 31, W = N + 2;
        const UlcbEncTables &tb = args.tb;
        const bool doPost = it < nBlocks, doPre = it >= 1 && it - 1 < nBlocks, doBlk = it >= 2;
        float2 *P = E + (size_t)(it % 3) * N, *Q = E + (size_t)((it + 2) % 3) * N, *R = E + (size_t)((it + 1) % 3) * N;
        const int b = it;
#ifdef ULCB_PHASE_TIMING
        long long dt0 = clock64();
#define DTICK(i) do { if (lane == 0) { long long t_ = clock64(); atomicAdd(&g_ulcbPhase[i], (unsigned long long)(t_ - dt0)); dt0 = t_; } } while (0)
#else
#define DTICK(i) do { } while (0)
#endif
        if (it + 1 < nBlocks && nSamples > 0) {
            // the next iteration's window comes straight from HBM: ask for its lines now, a whole iteration early
            const size_t elem = args.pcmFloat ? 4 : 2;
            const long long s0 = (long long)b * N + H - 1;
            const long long lo = s0 < 0 ? 0 : s0, hi = (s0 + W < nSamples ? s0 + W : nSamples);
            for (long long o = lo * nch() * (long long)elem + 128ll * lane; o < hi * nch() * (long long)elem; o += 128 * 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pcm + o));
        }
        if (doPost) {
            // stage the window: samples (b-1)*N + H - 1 .. b*N + H of the call's timeline
            const long long sBeg = (long long)(b - 1) * N + H - 1;
            int i0 = 0;
            if (sBeg < 0) {                                               // reaches back into the block before this call
                i0 = (-sBeg < (long long)W) ? (int)(-sBeg) : W;
                for (int i = lane; i < i0; i += 32)
#pragma unroll
                    for (int c = 0; c < nch(); c++) stg[c * W + i] = prevBlk ? prevBlk[c * N + (int)(N + sBeg + i)] : 0.0f;
            }
            if (nSamples > 0 && nch() == 2 && !args.pcmFloat) {
                // stereo PCM16: one 32-bit load per sample pair, 16 of them in flight per lane
                const uint32_t *p32 = reinterpret_cast<const uint32_t *>(pcm);
                for (int ib = i0 + lane; ib < W; ib += 512) {
                    uint32_t wv[16]; bool ok[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const int i = ib + 32 * u;
                        const long long sv = sBeg + i;
                        ok[u] = i < W && sv < nSamples;
                        wv[u] = ok[u] ? __ldg(p32 + sv) : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const int i = ib + 32 * u;
                        if (i < W) {
                            const float a = (float)(int16_t)(wv[u] & 0xffffu) * 0x1.0p-15f, bb = (float)(int16_t)(wv[u] >> 16) * 0x1.0p-15f;
                            stg[i] = ok[u] ? (a + bb) * 0.5f : 0.0f;
                            stg[W + i] = ok[u] ? (a - bb) * 0.5f : 0.0f;
                        }
                    }
                }
            } else if (nSamples > 0) {
                for (int ib = i0 + lane; ib < W; ib += 256) {             // 8 independent loads in flight per lane
                    float v[8][MAXCH]; bool ok[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int i = ib + 32 * u;
                        const long long sv = sBeg + i;
                        ok[u] = i < W && sv < nSamples;
                        const size_t sc = ok[u] ? (size_t)sv : 0;
#pragma unroll
                        for (int q = 0; q < nch(); q++)
                            v[u][q] = args.pcmFloat ? reinterpret_cast<const float *>(pcm)[sc * nch() + q]
                                                    : (float)reinterpret_cast<const int16_t *>(pcm)[sc * nch() + q] * 0x1.0p-15f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int i = ib + 32 * u;
#pragma unroll
                        for (int c = 1; c < nch(); c += 2) { float a = v[u][c - 1], bb = v[u][c]; v[u][c - 1] = (a + bb) * 0.5f; v[u][c] = (a - bb) * 0.5f; }
                        if (i < W)
#pragma unroll
                            for (int c = 0; c < nch(); c++) stg[c * W + i] = ok[u] ? v[u][c] : 0.0f;
                    }
                }
            } else {
                for (int i = i0 + lane; i < W; i += 32)
#pragma unroll
                    for (int c = 0; c < nch(); c++) stg[c * W + i] = 0.0f;
            }
            __syncwarp();
            DTICK(16);
#pragma unroll 4
            for (int n = lane; n < N; n += 32) {
                float ehp = 0.0f, ebp = 0.0f;
#pragma unroll
                for (int c = 0; c < nch(); c++) {
                    float t0 = stg[c * W + n], t1 = stg[c * W + n + 1], t2 = stg[c * W + n + 2];
                    float hp = (-t0 + 2.0f * t1) - t2, bp = -t0 + t2;
                    ehp += hp * hp; ebp += bp * bp;
                }
                P[n] = make_float2(sqrtf(ehp), sqrtf(ebp));
            }
        }
        __syncwarp();
        DTICK(17);
        // The five one-pole chains are the same three instructions (d = x - s; s += d * c) on different data, so they
        // run SIMD across lanes: lane 0/1 = post-mask HP/BP of block it (forwards, envelope written back in place),
        // lane 2/3 = pre-mask HP/BP of block it-1 (backwards; their cross products give the error energy, which lane 2
        // writes over the HP slot), lane 4 = block-mask of block it-2 (forwards) and its running segment sum.
        int decision = 0;
        {
            const bool isPost = lane < 2, isPre = lane == 2 || lane == 3, isBlk = lane == 4;
            float *base = reinterpret_cast<float *>(isPost ? P : isPre ? Q : R);
            const int comp = isBlk ? 0 : (lane & 1);                       // .x (HP / energy) or .y (BP)
            const int dir = isPre ? -1 : 1;
            float *ptr = base + comp + (isPre ? 2 * (N - 1) : 0);
            const float c = lane == 0 ? tb.c1PostHp : lane == 1 ? tb.c1PostBp : lane == 2 ? tb.c1PreHp : lane == 3 ? tb.c1PreBp : tb.c1Blk;
            float st = lane == 0 ? env[0] : lane == 1 ? env[1] : lane == 2 ? preHp : lane == 3 ? preBp : env[2];
            // Lanes 0-3 write their updated envelope back in place; the pre-mask lanes also park d = x - s in the (by now
            // dead) staging planes, and the error energy (dh*qbp)^2 + (db*qhp)^2 is formed afterwards by all 32 lanes.
            const bool store = (isPost && doPost) || (isPre && doPre);
            float *dpark = stg + (lane == 3 ? W : 0) + (N - 1);         // d of sample n at plane[n]; the pre-mask runs backwards
            const int bin = N >> 3;
            for (int sg = 0; sg < 8; sg++) {
                float acc = 0.0f;
                for (int n0 = 0; n0 < bin; n0 += 8) {                    // bin >= 32
                    float xs[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) xs[q] = ptr[2 * dir * q];   // every slot is read before it is overwritten
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float d = xs[q] - st;
                        st += d * c;
                        if (store) ptr[2 * dir * q] = st;
                        if (isPre) dpark[-q] = d;
                        acc += st;
                    }
                    ptr += 16 * dir;
                    dpark -= 8;
                }
                if (isBlk && doBlk) {
                    seg[2 * sg] = seg[16 + 2 * sg]; seg[2 * sg + 1] = seg[16 + 2 * sg + 1];
                    seg[16 + 2 * sg] = acc; seg[16 + 2 * sg + 1] = (float)bin;   // the count is one 1.0f per sample (:126)
                }
            }
            __syncwarp();
            if (doPre) {                                                 // error energy of block it-1 over the HP slot (:98-103)
#pragma unroll 4
                for (int n = lane; n < N; n += 32) {
                    const float2 sv = Q[n];                              // (qhp, qbp) after this sample's update
                    const float a = stg[n] * sv.y, bb = stg[W + n] * sv.x;
                    Q[n].x = a * a + bb * bb;
                }
            }
            // carry the envelope states
            const float s0 = __shfl_sync(0xffffffffu, st, 0), s1 = __shfl_sync(0xffffffffu, st, 1), s4 = __shfl_sync(0xffffffffu, st, 4);
            if (doPost) { env[0] = s0; env[1] = s1; preHp = s0; preBp = s1; }
            if (doBlk) env[2] = s4;
            __syncwarp();
            DTICK(18);
            if (doBlk && lane == 0) decision = window_decide(seg, tb.log2N);
            DTICK(21);
        }
        return __shfl_sync(0xffffffffu, decision, 0);
    }
};

template <int NCHT, int NTHR, bool GWS, bool PRE>
__global__ void __launch_bounds__(PRE ? NTHR : NTHR + 32, PRE ? (NTHR == 256 ? 4 : 2) : 1) ulcb_encode_kernel(UlcbEncodeArgs args) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int NCH = NCHT ? NCHT : args.nChan;
    using EncSmall = typename Encoder<NCHT, NTHR, PRE>::EncSmall;
    const int N = args.tb.N, CN = NCH * N, H = N >> 1;
    const int tid = threadIdx.x;
    const bool isDetector = !PRE && tid >= NTHR;                      // the extra warp
    EncSmall &sm = *reinterpret_cast<EncSmall *>(smraw);
    Work wk;
    float2 *detE = nullptr; float *detStage = nullptr;
    {
        constexpr int MAXBU = EncSmall::MAXBU;
        unsigned char *p = smraw + ((sizeof(EncSmall) + 15) & ~(size_t)15);
        if (PRE) {
            // one region serves the Bark analysis' f64 scratch, then the radix counters, then the coder scratch
            size_t u = band_scratch_bytes(MAXBU);
            if (u < 32 * (size_t)NTHR) u = 32 * (size_t)NTHR;
            if (!GWS && u < 3 * (size_t)CN) u = 3 * (size_t)CN;
            u = (u + 15) & ~(size_t)15;
            wk.cnt = reinterpret_cast<uint16_t *>(p); wk.encScratch = p;
            wk.lo = reinterpret_cast<double (*)[ULCB_NBARK][3]>(p);
            wk.stage = reinterpret_cast<double (*)[3][32]>(p + sizeof(double) * MAXBU * ULCB_NBARK * 3);
            p += u;
            wk.amp2 = reinterpret_cast<float *>(p); p += sizeof(float) * H;
            wk.mask = reinterpret_cast<float *>(p); p += sizeof(float) * H + 16;        // + the four spare words of the packed block
            float *big = GWS ? args.workspace + (size_t)blockIdx.x * args.wsFloats : reinterpret_cast<float *>(p);
            wk.R = nullptr; wk.XA = big; wk.XB = wk.XA + CN; wk.TA = wk.XB + CN; wk.TB = wk.TA + CN;
            if (GWS) wk.encScratch = reinterpret_cast<uint8_t *>(big + 4 * CN);
        } else {
            wk.lo = reinterpret_cast<double (*)[ULCB_NBARK][3]>(p); p += sizeof(double) * MAXBU * ULCB_NBARK * 3;
            wk.stage = reinterpret_cast<double (*)[3][32]>(p); p += sizeof(double) * MAXBU * 3 * 32;
            wk.cnt = reinterpret_cast<uint16_t *>(p); p += 32 * NTHR;
            wk.amp2 = reinterpret_cast<float *>(p); p += sizeof(float) * H;
            wk.mask = reinterpret_cast<float *>(p); p += sizeof(float) * H;
            if (!GWS) { detE = reinterpret_cast<float2 *>(p); p += sizeof(float2) * N * 3; }
            detStage = reinterpret_cast<float *>(p); p += sizeof(float) * (NCH < 2 ? 2 : NCH) * (N + 4);   // >= 2 planes: the pre-mask lanes park d there
            float *big = GWS ? args.workspace + (size_t)blockIdx.x * args.wsFloats : reinterpret_cast<float *>(p);
            if (GWS) detE = reinterpret_cast<float2 *>(big + 8 * CN + 64);
            wk.R = big; wk.XA = big + 3 * CN; wk.XB = wk.XA + CN; wk.TA = wk.XB + CN; wk.TB = wk.TA + CN;
            wk.encScratch = GWS ? reinterpret_cast<uint8_t *>(big + 7 * CN) : reinterpret_cast<uint8_t *>(wk.cnt);
        }
    }
    Encoder<NCHT, NTHR, PRE> enc(args, sm, wk);

    for (int stream = blockIdx.x; stream < args.nStreams; stream += gridDim.x) {
        const int nBlocks = args.nBlocks ? (int)args.nBlocks[stream] : args.nBlocksAll;
        const long long nSamples = args.nSamples ? (long long)args.nSamples[stream] : (long long)nBlocks * N;
        const size_t elem = args.pcmFloat ? 4 : 2;
        const unsigned char *pcm = reinterpret_cast<const unsigned char *>(args.pcm) + (size_t)stream * args.inStride * elem;
        UlcbEncStreamState *st = args.state ? args.state + stream : nullptr;
        const bool resume = !PRE && st && st->valid;
        csync<NTHR, PRE>();                                           // the previous stream is finished everywhere

        if (!PRE && isDetector) {
            // ================= transient detector warp: decisions for blocks 0..nBlocks-1, one block ahead =================
            const int lane = tid & 31;
            Detector<NCHT> det{args, detE, detStage, sm.seg, {0.0f, 0.0f, 0.0f}, pcm, nSamples,
                              resume ? args.stateSamples + (size_t)stream * 2 * CN + CN : nullptr};
            det.nchRt = NCH;
            if (lane == 0)
                for (int i = 0; i < 16; i++) { sm.seg[i] = 0.0f; sm.seg[16 + i] = resume ? st->seg[i] : 0.0f; }
            if (resume) for (int i = 0; i < 3; i++) det.env[i] = st->env[i];      // every lane: the chains run SIMD across lanes
            __syncwarp();
#ifdef ULCB_PHASE_TIMING
            long long detAcc = 0, detWait = 0;
#endif
            for (int it = 0; it < nBlocks + 2; it++) {
#ifdef ULCB_PHASE_TIMING
                long long t0 = clock64();
#endif
                const int d = det.iterate(it, nBlocks);               // decision of block it-2
                if (it < 2) continue;
                if (lane == 0) sm.wcRing[(it - 2) & 1] = d;
#ifdef ULCB_PHASE_TIMING
                long long t1 = clock64(); detAcc += t1 - t0;
#endif
                fsync<NTHR>();                                        // publish decision it-2; workers start that block
#ifdef ULCB_PHASE_TIMING
                detWait += clock64() - t1;
#endif
            }
#ifdef ULCB_PHASE_TIMING
            if (lane == 0) { atomicAdd(&g_ulcbPhase[13], (unsigned long long)detAcc); atomicAdd(&g_ulcbPhase[14], (unsigned long long)detWait); }
#endif
            if (st && lane == 0) {
                for (int i = 0; i < 3; i++) st->env[i] = det.env[i];
                for (int i = 0; i < 16; i++) st->seg[i] = sm.seg[16 + i];
            }
            continue;
        }

        // ================= workers =================
        uint8_t *out = args.out + (size_t)stream * args.outStride;
        uint64_t outPos = 0;
        int overflow = 0;
        uint32_t k0 = 0;
        if (resume) {
            k0 = st->calls;
            const float *ss = args.stateSamples + (size_t)stream * 2 * CN;
            for (int i = tid; i < CN; i += NTHR) { enc.blk((int)(k0 % 3) - 2)[i] = ss[i]; enc.blk((int)(k0 % 3) - 1)[i] = ss[CN + i]; }
            if (tid == 0) { sm.wcNext = st->nextWindowCtrl; sm.wcCur = st->windowCtrl; sm.wcPrev = st->prevWindowCtrl; }
        } else {
            if (!PRE) for (int i = tid; i < 3 * CN; i += NTHR) wk.R[i] = 0.0f;
            if (tid == 0) { sm.wcNext = 0x10; sm.wcCur = 0x10; sm.wcPrev = 0x10; }    // libulc/ulcEncoder.c:70
        }
        const int kb = (int)(k0 % 3);                                         // only k mod 3 matters for the rotating buffers
        wsync<NTHR>();

#ifdef ULCB_PHASE_TIMING
        long long phaseAcc[12] = {0}; long long tPrev = clock64();
#endif
        for (int b = 0; b < nBlocks; b++) {
            const int k = kb + b;
            if (!PRE) {
                long long valid = nSamples - (long long)b * N;
                enc.ingest(k, pcm + (size_t)b * N * NCH * elem, valid < 0 ? 0 : valid);
                TICK(0);
                fsync<NTHR>();                                        // decision b is published (and the ingest is visible)
                TICK(1);
                enc.begin_block(sm.wcRing[b & 1]);
                TICK(10);
                enc.fold(k);
            } else {
                enc.begin_block(__ldg(args.wcAll + (size_t)stream * args.wcStride + b));
                TICK(10);
                {   // the next block folds input block b, which nothing has touched since the detector pre-pass: ask for its lines now
                    const long long lo = (long long)b * N * NCH * (long long)elem, hi = min((long long)(b + 1) * N, nSamples) * NCH * (long long)elem;
                    for (long long o = lo + 128ll * tid; o < hi; o += 128ll * NTHR) asm volatile("prefetch.global.L2 [%0];" ::"l"(pcm + o));
                }
                enc.fold_pre(b, pcm, nSamples);
            }
            wsync<NTHR>();
            TICK(2);
            enc.transform();
            TICK(3);
            enc.analyse(k);
            wsync<NTHR>();
            TICK(4);
            const size_t dbgBlk = ((size_t)stream * args.nBlocksAll + b);
            if (args.dbgCoefs) for (int i = tid; i < CN; i += NTHR) args.dbgCoefs[dbgBlk * CN + i] = wk.XA[i];
            enc.serial_sums(k);
            wsync<NTHR>();
            TICK(5);
            if (args.mode == ULCB_ENC_ANALYSE) {                      // first pass of two-pass ABR: complexity and window control only
                if (tid == 0) {
                    const size_t mi = (size_t)stream * args.metaStride + b;
                    if (args.blockBytes) args.blockBytes[mi] = 0u;
                    if (args.complexity) args.complexity[mi] = sm.cx;
                    if (args.windowCtrl) args.windowCtrl[mi] = sm.wcCur;
                }
                wsync<NTHR>();
                continue;
            }
            enc.spread();
            wsync<NTHR>();
            TICK(6);
            if (args.dbgNoise) for (int i = tid; i < CN; i += NTHR) args.dbgNoise[dbgBlk * CN + i] = wk.TB[i];
            if (args.dbgMask) for (int i = tid; i < H; i += NTHR) args.dbgMask[dbgBlk * H + i] = wk.mask[i];
            enc.make_keys(args.dbgKeys ? args.dbgKeys + dbgBlk * CN : nullptr);
            wsync<NTHR>();
            // ---- rate control (libulc/ulcEncoder.c:93-158) ----
            const int maxCoef = sm.nNz;
            const float cx = sm.cx;
            int size, nOutUsed, nOutVbr = maxCoef;
            if (args.mode == ULCB_ENC_VBR && args.param > 0.0f) {
                float f = (float)(NCH * N) * cx / args.param;
                if (f < (float)maxCoef) nOutVbr = (int)f;
            }
            if (args.mode == ULCB_ENC_VBR && !args.dbgRanks && enc.can_select()) enc.rank_select(nOutVbr);
            else enc.rank_sort();
            TICK(7);
            if (args.dbgRanks) for (int i = tid; i < CN; i += NTHR) args.dbgRanks[dbgBlk * CN + i] = reinterpret_cast<const uint16_t *>(wk.TA)[i];

            if (args.mode == ULCB_ENC_VBR) {
                const int nOut = nOutVbr;
                nOutUsed = nOut;
                size = enc.encode_pass(k, nOut, true);
            } else {
                float kbps = args.param;
                if (args.mode == ULCB_ENC_ABR) kbps = args.param * cx / (args.avgPerStream ? args.avgPerStream[stream] : args.avgComplexity);
                const int budget = (int)(((float)N * kbps) * 1000.0f / (float)args.tb.rateHz);
                int lo = 0, hi = maxCoef, nOut = -1;
                if (lo < hi) do {
                    nOut = (int)((unsigned)(lo + hi) / 2u);
                    size = enc.encode_pass(k, nOut, false);
                    if (size < budget) lo = nOut;
                    else if (size > budget) hi = nOut - 1;
                    else { lo = nOut; break; }
                } while (lo < hi - 1);
                nOutUsed = lo;
                size = enc.encode_pass(k, lo, true);
            }
            if (args.dbgCounts && tid == 0) { args.dbgCounts[dbgBlk * 2] = maxCoef; args.dbgCounts[dbgBlk * 2 + 1] = nOutUsed; }
            TICK(8);

            // ---- emit ----
            const int nBytes = size >> 3;
            if (!overflow && outPos + (uint64_t)nBytes > args.outStride) overflow = 1;
            if (!overflow) {
                const uint8_t *ob = reinterpret_cast<const uint8_t *>(enc.out_words(k));
                for (int i = tid; i < nBytes; i += NTHR) out[outPos + i] = ob[i];
            }
            if (tid == 0) {
                const size_t mi = (size_t)stream * args.metaStride + b;
                if (args.blockBytes) args.blockBytes[mi] = overflow ? 0u : (uint32_t)nBytes;
                if (args.complexity) args.complexity[mi] = cx;
                if (args.windowCtrl) args.windowCtrl[mi] = sm.wcCur;
            }
            if (!overflow) outPos += (uint64_t)nBytes;
            wsync<NTHR>();
            TICK(9);
        }
#ifdef ULCB_PHASE_TIMING
        if (tid == 0) { for (int i = 0; i < 12; i++) atomicAdd(&g_ulcbPhase[i], (unsigned long long)phaseAcc[i]); atomicAdd(&g_ulcbPhase[12], (unsigned long long)nBlocks); }
#endif

        // ---- save state ----
        if (!PRE && st) {
            const int k = kb + nBlocks;
            float *ss = args.stateSamples + (size_t)stream * 2 * CN;
            for (int i = tid; i < CN; i += NTHR) { ss[i] = enc.blk(k - 2)[i]; ss[CN + i] = enc.blk(k - 1)[i]; }
            if (tid == 0) {
                st->valid = 1; st->calls = k0 + (uint32_t)nBlocks;
                st->nextWindowCtrl = sm.wcNext; st->windowCtrl = sm.wcCur; st->prevWindowCtrl = sm.wcPrev;
                st->complexity = sm.cx;
            }
        }
        if (tid == 0) {
            if (args.totalBytes) args.totalBytes[stream] = outPos;
            if (args.status) args.status[stream] = overflow;
        }
    }
}


// ---- transient detector pre-pass (ulcEncoder_WindowControl.c:31-239) ------------------------------------------------
// The detector is the only part of the encoder that is serial along a stream (three one-pole envelopes carried from block to
// block); everything it feeds is a per-block decision.  This kernel computes the decisions of every block up front so that
// the encode kernel needs neither the detector warp nor its 64 KB of buffers.
//
// One CTA takes 32 streams and walks their blocks together; the serial recurrences run ONE LANE PER STREAM, so a warp
// advances 32 envelopes per instruction instead of five.  Four roles are software-pipelined over the blocks and meet at one
// __syncthreads per step:
//   energy warps (DET_EW)  block j    tap energies sqrt(sum hp^2), sqrt(sum bp^2) of every sample      -> A[j % 3]
//   post warp              block j-1  post-masking envelopes, forwards, in place                        A[(j-1) % 3]
//   pre warp               block j-2  pre-masking envelopes backwards from the post end values, error energy -> R[j % 2]
//   block warp             block j-3  block-masking envelope over the error energy, 8 segment sums, the decision
// A / R live in global memory as [sample][32 lanes] so that every access of a chain warp is one coalesced row; the chains
// load two batches ahead of the arithmetic.  Every chain keeps the reference's operation order (same values as the in-kernel
// detector above, which stays for calls that carry state).
constexpr int DET_EW = 16, DET_THREADS = (DET_EW + 3) * 32;
constexpr int DET_CHUNK = 16;          // samples per staged chunk: [16][32 lanes] is one contiguous piece of A / R
constexpr int DET_STAGES = 8;          // chunks in flight per chain warp (divides N / DET_CHUNK for every block size >= 256)

// mbarrier / TMA bulk-copy helpers (the chain warps stage their rows through shared memory: the recurrences leave no room to hide
// a global-memory latency per batch, a ring of bulk copies runs DET_STAGES chunks ahead instead)
__device__ __forceinline__ unsigned det_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void det_mbar_wait(unsigned long long *b, unsigned parity) {
    unsigned done = 0;
    for (unsigned spins = 0; !done; spins++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(det_smem_u32(b)), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();                 // a copy that never lands is a bug, not a reason to hang the GPU
    }
}
__device__ __forceinline__ void det_fetch(void *dst, const void *src, unsigned bytes, unsigned long long *b) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the slot's last readers (generic proxy) come first
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(det_smem_u32(b)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(det_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(det_smem_u32(b)) : "memory");
}

// raw sample point t of a stream (zero outside [0, nSamples)) and its conversion to M/S floats
// (ulcEncoder_BlockTransform.c:87-111, tools/WavIO_Helper.c:49-55); split so that a batch of loads can be issued before any use
template <int NCH, bool FLT> struct DetRaw { uint32_t w[(NCH == 2 && !FLT) ? 1 : NCH]; };
template <int NCH, bool FLT>
__device__ __forceinline__ DetRaw<NCH, FLT> det_load_raw(const unsigned char *pcm, long long t, long long nSamples, bool want = true) {
    DetRaw<NCH, FLT> r;
    const bool ok = want && t >= 0 && t < nSamples;
    if (NCH == 2 && !FLT) r.w[0] = ok ? __ldg(reinterpret_cast<const uint32_t *>(pcm) + t) : 0u;
    else {
#pragma unroll
        for (int c = 0; c < NCH; c++)
            r.w[c < ((NCH == 2 && !FLT) ? 1 : NCH) ? c : 0] =
                !ok ? 0u : FLT ? __ldg(reinterpret_cast<const uint32_t *>(pcm) + (size_t)t * NCH + c)
                               : (uint32_t)(int)__ldg(reinterpret_cast<const int16_t *>(pcm) + (size_t)t * NCH + c);
    }
    return r;
}
template <int NCH, bool FLT>
__device__ __forceinline__ void det_cvt_ms(const DetRaw<NCH, FLT> &r, float *v) {
    if (NCH == 2 && !FLT) {
        v[0] = (float)(int16_t)(r.w[0] & 0xffffu) * 0x1.0p-15f; v[1 % NCH] = (float)(int16_t)(r.w[0] >> 16) * 0x1.0p-15f;
    } else {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const uint32_t x = r.w[c < ((NCH == 2 && !FLT) ? 1 : NCH) ? c : 0];
            v[c] = FLT ? __uint_as_float(x) : (float)(int)x * 0x1.0p-15f;
        }
    }
    if (NCH == 2) { const float a = v[0], b = v[1]; v[0] = (a + b) * 0.5f; v[1] = (a - b) * 0.5f; }
}

template <int NCH, bool FLT>
__global__ void __launch_bounds__(DET_THREADS, 1) ulcb_detect_kernel(UlcbEncodeArgs args, int streamBase, float2 *scrA, float *scrR, int32_t *wcAll, uint64_t wcStride) {
    extern __shared__ __align__(128) unsigned char detsm[];
    // shared memory: three chunk rings (post, pre: float2 rows; block: float rows), the transposition tiles, the segment sums
    float2 (*ringP)[DET_CHUNK][32] = reinterpret_cast<float2 (*)[DET_CHUNK][32]>(detsm);
    float2 (*ringQ)[DET_CHUNK][32] = ringP + DET_STAGES;
    float  (*ringE)[DET_CHUNK][32] = reinterpret_cast<float (*)[DET_CHUNK][32]>(ringQ + DET_STAGES);
    float2 (*tile)[32][33] = reinterpret_cast<float2 (*)[32][33]>(ringE + DET_STAGES);        // [DET_EW]
    float (*segs)[33] = reinterpret_cast<float (*)[33]>(tile + DET_EW);                        // [32 lanes] {sum,cnt} x 16
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(segs + 32);              // [3][DET_STAGES]
    const UlcbEncTables &tb = args.tb;
    const int N = tb.N, H = N >> 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = streamBase + blockIdx.x * 32, sMine = s0 + lane;   // scratch is per CTA of this launch, streams are absolute
    const bool live = sMine < args.nStreams;
    const int nbAll = args.nBlocksAll;
    const int nbMine = live ? (args.nBlocks ? min((int)args.nBlocks[sMine], nbAll) : nbAll) : 0;
    const long long nsMine = live ? (args.nSamples ? (long long)args.nSamples[sMine] : (long long)nbMine * N) : 0;
    const size_t elem = FLT ? 4 : 2;
    float2 *A = scrA + (size_t)blockIdx.x * 3 * N * 32;
    float  *R = scrR + (size_t)blockIdx.x * 2 * N * 32;
    const int nChunks = N / DET_CHUNK;
    float envH = 0.0f, envB = 0.0f, envK = 0.0f;                     // TransientFilter[3] of this lane's stream
    unsigned phase = 0;                                               // parity of the ring's mbarriers: one flip per DET_STAGES chunks
    if (warp == DET_EW + 2) for (int i = 0; i < 32; i++) segs[lane][i] = 0.0f;
    if (threadIdx.x < 3 * DET_STAGES)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(det_smem_u32(&mbar[threadIdx.x])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    for (int j = 0; j < nbAll + 3; j++) {
        if (warp < DET_EW) {
            // ---- tap energies of block j (:60-71): sample n of the window is centred on (j-1)*N + H + n ----
            if (j < nbAll) {
                float2 *dst = A + (size_t)(j % 3) * N * 32;
                const long long c0 = (long long)(j - 1) * N + H;
                const size_t strideB = (size_t)args.inStride * elem;
                const unsigned char *pcm0 = reinterpret_cast<const unsigned char *>(args.pcm) + (size_t)s0 * strideB;
                for (int n0 = warp * 32; n0 < N; n0 += DET_EW * 32) {
                    // lane = sample n0 + lane of stream s0 + ss.  All loads of a group of streams are issued before the first use;
                    // the two neighbours of a sample come from the adjacent lanes (lanes 0 / 31 load theirs).
                    constexpr int G = 8;
                    const long long t = c0 + n0 + lane;
                    const bool edge = lane == 0 || lane == 31;
                    const long long tEdge = lane == 0 ? t - 1 : t + 1;
                    // streams whose samples t-1 .. t+1 of the whole tile lie inside the stream need no bounds checks
                    const unsigned inside = __ballot_sync(0xffffffffu, c0 + n0 - 1 >= 0 && c0 + n0 + 32 < nsMine);
                    for (int g0 = 0; g0 < 32; g0 += G) {
                        if (!FLT && NCH == 2 && ((inside >> g0) & ((1u << G) - 1u)) == ((1u << G) - 1u)) {
                            // PCM16 stereo: M/S and the taps are exact in f32 (sums of a few 16-bit integers scaled by a power of
                            // two), so they are formed in integers: M = (l + r) * 2^-16, hp = -M0 + 2 M1 - M2, bp = M2 - M0 give
                            // the very floats of the reference's f32 expressions; only the squares, their sum and the root round.
                            uint32_t mid[G], side[G];
#pragma unroll
                            for (int q = 0; q < G; q++) {
                                const uint32_t *p32 = reinterpret_cast<const uint32_t *>(pcm0 + (size_t)(g0 + q) * strideB);
                                mid[q] = __ldg(p32 + t);
                                side[q] = edge ? __ldg(p32 + tEdge) : 0u;
                            }
#pragma unroll
                            for (int q = 0; q < G; q++) {
                                const int l1 = (int)(int16_t)(mid[q] & 0xffffu), r1 = (int)mid[q] >> 16;
                                const int ls = (int)(int16_t)(side[q] & 0xffffu), rs = (int)side[q] >> 16;
                                const int m1 = l1 + r1, d1 = l1 - r1, ms = ls + rs, ds = ls - rs;
                                int m0 = __shfl_up_sync(0xffffffffu, m1, 1), d0 = __shfl_up_sync(0xffffffffu, d1, 1);
                                int m2 = __shfl_down_sync(0xffffffffu, m1, 1), d2 = __shfl_down_sync(0xffffffffu, d1, 1);
                                if (lane == 0) { m0 = ms; d0 = ds; }
                                if (lane == 31) { m2 = ms; d2 = ds; }
                                const float hpM = (float)(2 * m1 - m0 - m2) * 0x1.0p-16f, bpM = (float)(m2 - m0) * 0x1.0p-16f;
                                const float hpS = (float)(2 * d1 - d0 - d2) * 0x1.0p-16f, bpS = (float)(d2 - d0) * 0x1.0p-16f;
                                const float ehp = hpM * hpM + hpS * hpS, ebp = bpM * bpM + bpS * bpS;
                                tile[warp][g0 + q][lane] = make_float2(sqrtf(ehp), sqrtf(ebp));
                            }
                            continue;
                        }
                        DetRaw<NCH, FLT> mid[G], side[G];
#pragma unroll
                        for (int q = 0; q < G; q++) {
                            const long long ns = __shfl_sync(0xffffffffu, nsMine, g0 + q);
                            const unsigned char *pcm = pcm0 + (size_t)(g0 + q) * strideB;
                            mid[q] = det_load_raw<NCH, FLT>(pcm, t, ns);
                            side[q] = det_load_raw<NCH, FLT>(pcm, tEdge, ns, edge);
                        }
#pragma unroll
                        for (int q = 0; q < G; q++) {
                            float v0[NCH], v1[NCH], v2[NCH], vs[NCH];
                            det_cvt_ms<NCH, FLT>(mid[q], v1);
                            det_cvt_ms<NCH, FLT>(side[q], vs);
                            float ehp = 0.0f, ebp = 0.0f;
#pragma unroll
                            for (int c = 0; c < NCH; c++) {
                                v0[c] = __shfl_up_sync(0xffffffffu, v1[c], 1);
                                v2[c] = __shfl_down_sync(0xffffffffu, v1[c], 1);
                                if (lane == 0) v0[c] = vs[c];
                                if (lane == 31) v2[c] = vs[c];
                                const float hp = (-v0[c] + 2.0f * v1[c]) - v2[c], bp = -v0[c] + v2[c];
                                ehp += hp * hp; ebp += bp * bp;
                            }
                            tile[warp][g0 + q][lane] = make_float2(sqrtf(ehp), sqrtf(ebp));
                        }
                    }
                    __syncwarp();
#pragma unroll 8
                    for (int k = 0; k < 32; k++) __stcg(&dst[(size_t)(n0 + k) * 32 + lane], tile[warp][lane][k]);
                    __syncwarp();
                }
            }
        } else if (warp == DET_EW) {
            // ---- post-masking envelopes of block j-1, forwards, in place (:73-88) ----
            const int b = j - 1;
            if (b >= 0 && b < nbAll) {
                float2 *P = A + (size_t)(b % 3) * N * 32;
                const float cH = tb.c1PostHp, cB = tb.c1PostBp;
                unsigned long long *mb = mbar;
                if (lane == 0)
                    for (int k = 0; k < DET_STAGES; k++) det_fetch(ringP[k], P + (size_t)k * DET_CHUNK * 32, sizeof(float2) * DET_CHUNK * 32, &mb[k]);
                for (int k = 0; k < nChunks; k++) {
                    const int slot = k % DET_STAGES;
                    det_mbar_wait(&mb[slot], phase);
                    float2 x[DET_CHUNK];
#pragma unroll
                    for (int q = 0; q < DET_CHUNK; q++) x[q] = ringP[slot][q][lane];
                    __syncwarp();
                    if (lane == 0 && k + DET_STAGES < nChunks)
                        det_fetch(ringP[slot], P + (size_t)(k + DET_STAGES) * DET_CHUNK * 32, sizeof(float2) * DET_CHUNK * 32, &mb[slot]);
#pragma unroll
                    for (int q = 0; q < DET_CHUNK; q++) {
                        const float dh = x[q].x - envH; envH += dh * cH;
                        const float db = x[q].y - envB; envB += db * cB;
                        __stcg(&P[(size_t)(k * DET_CHUNK + q) * 32 + lane], make_float2(envH, envB));
                    }
                    if (slot == DET_STAGES - 1) phase ^= 1u;
                }
            }
        } else if (warp == DET_EW + 1) {
            // ---- pre-masking envelopes of block j-2, backwards from the post-masking end values; error energy (:90-103) ----
            const int b = j - 2;
            if (b >= 0 && b < nbAll) {
                const float2 *Q = A + (size_t)(b % 3) * N * 32;
                float *E = R + (size_t)(b & 1) * N * 32;
                const float cH = tb.c1PreHp, cB = tb.c1PreBp;
                unsigned long long *mb = mbar + DET_STAGES;
                if (lane == 0)
                    for (int k = 0; k < DET_STAGES; k++)
                        det_fetch(ringQ[k], Q + (size_t)(nChunks - 1 - k) * DET_CHUNK * 32, sizeof(float2) * DET_CHUNK * 32, &mb[k]);
                float sH = 0.0f, sB = 0.0f;
                for (int k = 0; k < nChunks; k++) {                  // chunk nChunks-1-k, samples in descending order
                    const int slot = k % DET_STAGES, ck = nChunks - 1 - k;
                    det_mbar_wait(&mb[slot], phase);
                    float2 x[DET_CHUNK];
#pragma unroll
                    for (int q = 0; q < DET_CHUNK; q++) x[q] = ringQ[slot][DET_CHUNK - 1 - q][lane];
                    __syncwarp();
                    if (lane == 0 && k + DET_STAGES < nChunks)
                        det_fetch(ringQ[slot], Q + (size_t)(ck - DET_STAGES) * DET_CHUNK * 32, sizeof(float2) * DET_CHUNK * 32, &mb[slot]);
                    if (k == 0) { sH = x[0].x; sB = x[0].y; }         // the post-masking state at the end of the block
#pragma unroll
                    for (int q = 0; q < DET_CHUNK; q++) {
                        const float dh = x[q].x - sH; sH += dh * cH;
                        const float db = x[q].y - sB; sB += db * cB;
                        const float a = dh * sB, bb = db * sH;
                        __stcg(&E[(size_t)(ck * DET_CHUNK + DET_CHUNK - 1 - q) * 32 + lane], a * a + bb * bb);
                    }
                    if (slot == DET_STAGES - 1) phase ^= 1u;
                }
            }
        } else {
            // ---- block-masking envelope of block j-3 over the error energy, segment sums, decision (:105-239) ----
            const int b = j - 3;
            if (b >= 0 && b < nbAll) {
                const float *E = R + (size_t)(b & 1) * N * 32;
                const float cK = tb.c1Blk;
                const int chunksPerSeg = nChunks >> 3;
                float *seg = segs[lane];
                unsigned long long *mb = mbar + 2 * DET_STAGES;
                if (lane == 0)
                    for (int k = 0; k < DET_STAGES; k++) det_fetch(ringE[k], E + (size_t)k * DET_CHUNK * 32, sizeof(float) * DET_CHUNK * 32, &mb[k]);
                int k = 0;
                for (int sg = 0; sg < 8; sg++) {
                    float acc = 0.0f;
                    for (int m = 0; m < chunksPerSeg; m++, k++) {
                        const int slot = k % DET_STAGES;
                        det_mbar_wait(&mb[slot], phase);
                        float x[DET_CHUNK];
#pragma unroll
                        for (int q = 0; q < DET_CHUNK; q++) x[q] = ringE[slot][q][lane];
                        __syncwarp();
                        if (lane == 0 && k + DET_STAGES < nChunks)
                            det_fetch(ringE[slot], E + (size_t)(k + DET_STAGES) * DET_CHUNK * 32, sizeof(float) * DET_CHUNK * 32, &mb[slot]);
#pragma unroll
                        for (int q = 0; q < DET_CHUNK; q++) { const float d = x[q] - envK; envK += d * cK; acc += envK; }
                        if (slot == DET_STAGES - 1) phase ^= 1u;
                    }
                    seg[2 * sg] = seg[16 + 2 * sg]; seg[2 * sg + 1] = seg[16 + 2 * sg + 1];
                    seg[16 + 2 * sg] = acc; seg[16 + 2 * sg + 1] = (float)(N >> 3);     // the count is one 1.0f per sample (:126)
                }
                const int decision = window_decide(seg, tb.log2N);
                if (b < nbMine) wcAll[(size_t)sMine * wcStride + b] = decision;
            }
        }
        // what this step wrote to A / R is read by the bulk copies (async proxy) of the next step
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
    }
}
}  // namespace

#ifdef ULCB_PHASE_TIMING
extern "C" int ULCB_DebugPhaseTimes(unsigned long long *out13) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(out13, g_ulcbPhase, sizeof(unsigned long long) * 32);
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_ulcbPhase, z, sizeof z);
    return (int)e;
}
#endif

namespace {
__global__ void ulcb_avg_complexity_kernel(const float *__restrict__ cx, uint64_t metaStride, const uint32_t *__restrict__ nBlocks, uint32_t nBlocksAll,
                                           int nStreams, float *__restrict__ avg) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nStreams) return;
    const uint32_t nb = nBlocks ? nBlocks[s] : nBlocksAll;
    double sum = 0.0;                                                   // ComplexitySum += Encoder.BlockComplexity, block order (:164)
    for (uint32_t b = 0; b < nb; b++) sum += (double)cx[(size_t)s * metaStride + b];
    const double mean = nb ? sum / (double)nb : 0.0;                    // :177
    avg[s] = (float)(rint(mean * 1e5) / 1e5);                           // printed with %.5f, typed back, read with %f (:41,:187)
}
}  // namespace

cudaError_t ulcb_avg_complexity_launch(const float *dComplexity, uint64_t metaStride, const uint32_t *dNBlocks, uint32_t nBlocksAll,
                                       int nStreams, float *dAvg, cudaStream_t stream) {
    ulcb_avg_complexity_kernel<<<(nStreams + 127) / 128, 128, 0, stream>>>(dComplexity, metaStride, dNBlocks, nBlocksAll, nStreams, dAvg);
    return cudaGetLastError();
}

bool ulcb_encode_uses_global_ws(int N, int nChan) { return (size_t)nChan * N > 4096 || nChan > 2; }
int  ulcb_encode_threads(int N, int nChan) { return (nChan * N >= 4096) ? 512 : 256; }
size_t ulcb_encode_ws_floats(int N, int nChan) { return (size_t)8 * nChan * N + 64 + (size_t)6 * N; }   // + the detector's 3 block buffers
size_t ulcb_encode_smem_bytes(int N, int nChan, int nthr, bool globalWs, bool pre) {
    const size_t small = nChan == 1 ? sizeof(EncSmallT<4>) : nChan == 2 ? sizeof(EncSmallT<8>) : sizeof(EncSmallT<4 * ULCB_ENC_MAXCH>);
    const size_t band = band_scratch_bytes((nChan == 1 ? 4 : nChan == 2 ? 8 : 4 * ULCB_ENC_MAXCH) + 4);
    size_t s = (small + 15) & ~(size_t)15;
    if (pre) {                                                          // mirrors the carve-up in the kernel
        size_t u = band;
        if (u < 32 * (size_t)nthr) u = 32 * (size_t)nthr;
        if (!globalWs && u < 3 * (size_t)nChan * N) u = 3 * (size_t)nChan * N;
        u = (u + 15) & ~(size_t)15;
        return s + u + sizeof(float) * N + 16 + (globalWs ? 0 : sizeof(float) * (size_t)4 * nChan * N) + 16;
    }
    s += band + 32 * (size_t)nthr + sizeof(float) * N + (globalWs ? 0 : sizeof(float2) * N * 3) + sizeof(float) * (nChan < 2 ? 2 : nChan) * (N + 4);
    if (!globalWs) s += sizeof(float) * ((size_t)7 * nChan * N + 64);
    return s + 16;
}

template <int NCH, int NTHR, bool PRE>
static cudaError_t launch_e(const UlcbEncodeArgs &args, int gridBlocks, cudaStream_t stream) {
    const bool gws = args.workspace != nullptr;
    void (*k)(UlcbEncodeArgs) = gws ? ulcb_encode_kernel<NCH, NTHR, true, PRE> : ulcb_encode_kernel<NCH, NTHR, false, PRE>;
    size_t smem = ulcb_encode_smem_bytes(args.tb.N, NCH ? NCH : args.nChan, NTHR, gws, PRE);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (PRE) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k<<<gridBlocks, PRE ? NTHR : NTHR + 32, smem, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t ulcb_encode_launch(const UlcbEncodeArgs &args, int nChan, int gridBlocks, cudaStream_t stream) {
    int nthr = ulcb_encode_threads(args.tb.N, nChan);
    if (args.wcAll) {                                                   // window controls precomputed: stereo / mono, shared-memory shapes
        if (args.state || nChan > 2) return cudaErrorInvalidValue;
        if (args.encThreads == 256 || args.encThreads == 512) nthr = args.encThreads;
        if (nChan == 2) return nthr == 512 ? launch_e<2, 512, true>(args, gridBlocks, stream) : launch_e<2, 256, true>(args, gridBlocks, stream);
        return nthr == 512 ? launch_e<1, 512, true>(args, gridBlocks, stream) : launch_e<1, 256, true>(args, gridBlocks, stream);
    }
    if (nChan == 2) return nthr == 512 ? launch_e<2, 512, false>(args, gridBlocks, stream) : launch_e<2, 256, false>(args, gridBlocks, stream);
    if (nChan == 1) return nthr == 512 ? launch_e<1, 512, false>(args, gridBlocks, stream) : launch_e<1, 256, false>(args, gridBlocks, stream);
    // 3 .. ULCB_ENC_MAXCH channels: one instantiation with the channel count at run time, work buffers in the global workspace
    if (nChan <= ULCB_ENC_MAXCH && nChan == args.nChan && args.workspace)
        return nthr == 512 ? launch_e<0, 512, false>(args, gridBlocks, stream) : launch_e<0, 256, false>(args, gridBlocks, stream);
    return cudaErrorInvalidValue;
}

// ---- detector pre-pass ----
bool ulcb_detect_supported(int N, int nChan) { return (nChan == 1 || nChan == 2) && N >= 256; }
// with the window controls precomputed the encode kernel keeps X / T, the line buffers and one scratch region in shared memory
// when they fit; larger blocks keep X / T in the global workspace
bool ulcb_encode_pre_uses_global_ws(int N, int nChan) { return ulcb_encode_smem_bytes(N, nChan, ulcb_encode_threads(N, nChan), false, true) > 227 * 1024; }
// The pre-pass runs in waves of one CTA (32 streams) per SM: a launch is bound by the serial walk over the blocks, not by how many
// of the SMs it fills, and the scratch (65 KB per stream at N = 2048) stays bounded however many streams a call brings.
static int det_wave_ctas(int nStreams, int smCount) { const int c = (nStreams + 31) / 32; return c < smCount ? c : smCount; }
static size_t det_a_bytes(int N, int ctas) { return sizeof(float2) * (size_t)ctas * 3 * N * 32; }
size_t ulcb_detect_scratch_bytes(int N, int nStreams, int smCount) {
    const int ctas = det_wave_ctas(nStreams, smCount);
    return det_a_bytes(N, ctas) + sizeof(float) * (size_t)ctas * 2 * N * 32;
}
cudaError_t ulcb_detect_launch(const UlcbEncodeArgs &args, int nChan, int smCount, void *scratch, int32_t *wcAll, uint64_t wcStride, cudaStream_t stream, int *launches) {
    if (!ulcb_detect_supported(args.tb.N, nChan) || args.state) return cudaErrorInvalidValue;
    void (*k)(UlcbEncodeArgs, int, float2 *, float *, int32_t *, uint64_t) =
        nChan == 2 ? (args.pcmFloat ? ulcb_detect_kernel<2, true> : ulcb_detect_kernel<2, false>)
                   : (args.pcmFloat ? ulcb_detect_kernel<1, true> : ulcb_detect_kernel<1, false>);
    const size_t smem = (sizeof(float2) * 2 + sizeof(float)) * DET_STAGES * DET_CHUNK * 32 + sizeof(float2) * DET_EW * 32 * 33 +
                        sizeof(float) * 32 * 33 + sizeof(unsigned long long) * 3 * DET_STAGES + 128;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int ctas = det_wave_ctas(args.nStreams, smCount);
    float2 *scrA = reinterpret_cast<float2 *>(scratch);
    float *scrR = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(scratch) + det_a_bytes(args.tb.N, ctas));
    for (int base = 0; base < args.nStreams; base += ctas * 32) {
        const int left = (args.nStreams - base + 31) / 32;
        k<<<left < ctas ? left : ctas, DET_THREADS, smem, stream>>>(args, base, scrA, scrR, wcAll, wcStride);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}
