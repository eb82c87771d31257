// ulcb_decode.cu - batched ulc-codec block decoder for sm_100a.
//
// One CTA walks one stream front to back (libulc/ulcDecoder.c:198-302 is inherently sequential per
// stream: the half-block lap, LastSubBlockSize, the noise RNG position and - because the container
// has no block index, tools/ulc_Helper.h:10-20 - even the byte offset of block k+1 are only known
// after block k has been parsed).  Thousands of streams give thousands of independent CTAs.
//
// Per block:
//   1. PARSE (replaces Block_Decode_DecodeSubBlockCoefs, ulcDecoder.c:99-197).  The nybble grammar is
//      serial, so it is parsed speculatively: the stream is cut into 8-nybble segments; every thread
//      computes, for each of the 8 possible token-start offsets in its segment, where that token
//      chain leaves the segment (0..4 nybbles into the next one), how many coefficients and noise
//      draws it emits, the last quantiser it sets and whether it hits a stop code.  One warp then
//      resolves the real chain per sub-block (a scan of 5-entry maps + prefix sums), and all threads
//      emit their segment's tokens straight into the coefficient buffer.
//   2. NOISE.  The reference draws one xorshift32 per noise coefficient from a single serial
//      sequence (ulcDecoder.c:75-81); only the MSB is used, as a cumulative sign toggle.  The sequence
//      is the same for every stream (seed 1234567), so its running parity is a precomputed bit table
//      in L2 and a run's signs are parity[g+i+1] ^ parity[g] - fully parallel, bit-exact.
//   3. IMDCT (Fourier_IMDCT_Template.h:22-94) per channel/sub-block: FFT-based DCT-IV (ulcb_dct.cuh),
//      sine-window un-fold against the pending tail, window-switching delay line (ulcDecoder.c:248-272)
//      kept as ONE time-ordered pending buffer per channel.
//   4. M/S undo, interleave, float or PCM16 out (ulcDecoder.c:281-297, tools/WavIO_Helper.c:57-63).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "ulcb_dct.cuh"
#include "ulcb_decode.h"

#ifdef ULCB_PHASE_TIMING
__device__ unsigned long long g_ulcbDecPhase[16];
#define DTK(i) do { if (tid == 0) { long long t_ = clock64(); phaseAcc[i] += t_ - tPrev; tPrev = t_; } } while (0)
extern "C" int ULCB_DebugDecodePhaseTimes(unsigned long long *out16) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_ulcbDecPhase, sizeof(unsigned long long) * 16);
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_ulcbDecPhase, z, sizeof z);
    return (int)e;
}
#else
#define DTK(i) do { } while (0)
#endif

namespace {

constexpr int PAGE_WORDS = 256;                // one parse page = 256 32-bit words = 2048 nybbles
constexpr int PAGE_NYB   = PAGE_WORDS * 8;
constexpr int INFO_ROW   = PAGE_WORDS + 32;    // padded: resolver lane l reads segments 8l..8l+7 conflict-free

__device__ __forceinline__ int info_col(int s) { return s + (s >> 3); }

// token-chain summary of one (segment, start offset)
//   bits 0-2 exit (nybbles spilled into next segment), bit 3 stop, bits 4-17 coefs, bits 18-31 draws
__device__ __forceinline__ uint32_t pack_info(int exitp, int stop, int coefs, int draws) {
    return (uint32_t)exitp | ((uint32_t)stop << 3) | ((uint32_t)coefs << 4) | ((uint32_t)draws << 18);
}

struct NoiseRun { uint32_t start_n; float amp; uint32_t draw; float decay; };   // start | n<<16

// Decimation layouts, one nybble per sub-block, LSB first: bits0-2 shift, bit3 transient
// (libulc/ulcHelper.h:24-46, FormatSpecs.md:35-51)
__constant__ uint16_t kLayout[16] = {
    0x0000, 0x0008, 0x0019, 0x0091, 0x012A, 0x01A2, 0x02A1, 0x0A21,
    0x123B, 0x12B3, 0x13B2, 0x1B32, 0x23B1, 0x2B31, 0x3B21, 0xB321,
};

struct Token { int len, coefs, draws, stop, q; };   // q: -1 none, else quantiser index; stop: 0 none,1 zeros,2 noise

// Classify the token that starts at the low nybble of x (x holds >= 5 nybbles).
__device__ __forceinline__ Token classify(uint32_t x) {
    Token t;
    int v0 = x & 15, v1 = (x >> 4) & 15, v2 = (x >> 8) & 15, v3 = (x >> 12) & 15;
    t.len = 1; t.coefs = 1; t.draws = 0; t.stop = 0; t.q = -1;
    int b12 = (v1 << 4) | v2;
    if (v0 == 0)      { t.len = 2; t.coefs = v1 + 1; }
    else if (v0 == 1) { t.len = 3; t.coefs = b12 + 33; }
    else if (v0 == 8) { t.len = 4; t.coefs = ((b12 << 1) | (v3 & 1)) + 16; t.draws = t.coefs; }
    else if (v0 == 15) {
        t.coefs = 0;
        if (v1 < 14)       { t.len = 2; t.q = v1; }
        else if (v1 == 14) { t.len = 3; if (v2 == 15) t.stop = 1; else t.q = 14 + v2; }
        else               { t.len = 5; t.stop = 2; }
    }
    return t;
}

__device__ __forceinline__ float quant_scale(int q) {      // ulcDecoder.c:96-98
    return 0x1.0p-31f * (float)((1u << 26) >> (q & 31));
}

template <int NCH>
struct Smem {
    float    *buf;       // [NCH][N]   coefficients -> DCT output -> planar PCM
    float    *pend;      // [NCH][N/2] pending samples in time order (delay line + aliased tail)
    uint32_t *page;      // [PAGE_WORDS + 4]
    uint32_t *info;      // [8][INFO_ROW]
    uint8_t  *lq;        // [8][INFO_ROW]
    uint16_t *pmap;      // [INFO_ROW] 5 x 3-bit exits for entry offsets 0..4
    uint32_t *emitA;     // [PAGE_WORDS] coefStart | phase<<16 | q<<20 | valid<<31
    uint32_t *emitB;     // [PAGE_WORDS] drawStart (relative to block draw base)
    NoiseRun *runs;      // [maxRuns]
    int      *ctl;       // control scalars
    int      *scan;      // [SCAN_INTS] scratch of the chain resolver
};
enum { C_NRUNS = 0, C_SEND, C_ENDPOS, C_ENDKIND, C_ENDCOEFS, C_ENDDRAWS, C_ENDQ, C_ENDZYX, C_OK,
       C_EXITPH, C_SUMCOEFS, C_SUMDRAWS, C_LASTQ, C_COUNT };

// ------------------------------------------------------------------------------------------------
// Page load + speculative per-segment DP.
// ------------------------------------------------------------------------------------------------
template <int NCH>
__device__ void load_page(const Smem<NCH> &sm, const uint32_t *abase, long long pageIdx, long long nWordsValid) {
    const int tid = threadIdx.x;
    for (int w = tid; w < PAGE_WORDS + 1; w += blockDim.x) {
        long long gw = pageIdx * PAGE_WORDS + w;
        sm.page[w] = (gw < nWordsValid) ? __ldg(abase + gw) : 0u;
    }
    __syncthreads();
    if (tid < PAGE_WORDS) {
        uint64_t W = (uint64_t)sm.page[tid] | ((uint64_t)sm.page[tid + 1] << 32);
        uint32_t inf[8 + 5];
        int      lqv[8 + 5];
#pragma unroll
        for (int o = 7; o >= 0; o--) {
            Token t = classify((uint32_t)(W >> (4 * o)));
            int nxt = o + t.len;
            uint32_t r; int l;
            int tq = (t.q >= 0) ? (32 | t.q) : 0;
            if (t.stop) { r = pack_info(0, 1, 0, 0); l = 0; }
            else if (nxt >= 8) { r = pack_info(nxt - 8, 0, t.coefs, t.draws); l = tq; }
            else {
                uint32_t nx = 0; int nl = 0;             // select inf[o+len]; indices are static after unrolling
#pragma unroll
                for (int d = 1; d <= 5; d++)
                    if (o + d < 8 && t.len == d) { nx = inf[o + d]; nl = lqv[o + d]; }
                r = nx + (((uint32_t)t.coefs << 4) | ((uint32_t)t.draws << 18));
                l = nl ? nl : tq;
            }
            inf[o] = r; lqv[o] = l;
        }
        int col = info_col(tid);
        uint32_t m = 0;
#pragma unroll
        for (int o = 0; o < 8; o++) {
            sm.info[o * INFO_ROW + col] = inf[o];
            sm.lq[o * INFO_ROW + col]   = (uint8_t)lqv[o];
            if (o < 5) m |= (((inf[o] & 8) ? 7u : (inf[o] & 7u)) << (3 * o));
        }
        sm.pmap[col] = (uint16_t)m;
    }
    __syncthreads();
}

// Read one nybble at page-relative index i (valid for i < PAGE_NYB + 8)
template <int NCH>
__device__ __forceinline__ int page_nyb(const Smem<NCH> &sm, int i) {
    return (sm.page[i >> 3] >> ((i & 7) * 4)) & 15;
}
template <int NCH>
__device__ __forceinline__ uint32_t page_window(const Smem<NCH> &sm, int i) {  // 5+ nybbles starting at i
    uint64_t W = (uint64_t)sm.page[i >> 3] | ((uint64_t)sm.page[(i >> 3) + 1] << 32);
    return (uint32_t)(W >> ((i & 7) * 4));
}

// ------------------------------------------------------------------------------------------------
// Resolve the token chain that starts at page-relative nybble tp with `remain` coefficients still to come,
// quantiser qcur.  Every thread owns one 8-nybble segment of the page.  Which phase (token-start offset 0..4)
// each segment is entered at is a prefix composition of the segments' 5-state transition maps - an associative
// operator, so it is a CTA-wide scan; coefficient / noise-draw counts and "last quantiser set" are then ordinary
// scans of the per-segment values looked up at that phase.  Writes emit records for the segments on the chain
// and the control block (end segment/position/kind, or the carry-out to the next page).
// ------------------------------------------------------------------------------------------------
enum { MAP_IDENT = 0 | (1 << 3) | (2 << 6) | (3 << 9) | (4 << 12) | (5 << 15), SCAN_INTS = 136 };

__device__ __forceinline__ uint32_t map_compose(uint32_t f, uint32_t g) {      // apply f, then g; state 5 = stopped
    uint32_t h = 0;
#pragma unroll
    for (int e = 0; e < 6; e++) h |= ((g >> (3 * ((f >> (3 * e)) & 7))) & 7) << (3 * e);
    return h;
}

template <int NCH>
__device__ void resolve_chain(const Smem<NCH> &sm, int tp, int remain, int qcur, int coefBase, int drawBase) {
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = tid, s0 = tp >> 3, e0 = tp & 7;
    const bool has = s < PAGE_WORDS;
    const int col = info_col(has ? s : 0);
    int *scr = sm.scan;
    constexpr int SEGW = PAGE_WORDS / 32;                          // warps that own segments

    // ---- entry phases: scan of transition maps ----
    uint32_t m = MAP_IDENT;
    if (has && s >= s0) {
        if (s == s0) {
            uint32_t w = sm.info[e0 * INFO_ROW + col];
            uint32_t x = (w & 8) ? 5u : (w & 7u);
            m = x * 0x9249u;                                       // the same exit whatever the (irrelevant) input
        } else {
            uint32_t pm = sm.pmap[col];
            m = 5u << 15;
#pragma unroll
            for (int e = 0; e < 5; e++) { uint32_t v = (pm >> (3 * e)) & 7u; m |= (v == 7u ? 5u : v) << (3 * e); }
        }
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, m, d); if (lane >= d) m = map_compose(t, m); }
    if (lane == 31 && warp < SEGW) scr[warp] = (int)m;
    if (tid == 0) scr[132] = 1 << 30;
    __syncthreads();
    uint32_t pre = MAP_IDENT;
    for (int wv = 0; wv < warp && wv < SEGW; wv++) pre = map_compose(pre, (uint32_t)scr[wv]);
    const uint32_t minc = map_compose(pre, m);
    uint32_t mexc = __shfl_up_sync(FULL, minc, 1);
    if (lane == 0) mexc = pre;
    const int ph = (s == s0) ? e0 : (int)(mexc & 7u);
    const bool on = has && s >= s0 && (s == s0 || ph != 5);

    // ---- counts at the true phase: coefficients, noise draws, last quantiser ----
    int c = 0, d = 0, l = 0, stop = 0;
    if (on) {
        uint32_t w = sm.info[ph * INFO_ROW + col];
        c = (w >> 4) & 0x3fff; d = (int)(w >> 18); l = sm.lq[ph * INFO_ROW + col]; stop = (w >> 3) & 1;
    }
    int ic = c, id = d, il = l;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        int tc = __shfl_up_sync(FULL, ic, k), td = __shfl_up_sync(FULL, id, k), tl = __shfl_up_sync(FULL, il, k);
        if (lane >= k) { ic += tc; id += td; if (!il) il = tl; }
    }
    if (lane == 31 && warp < SEGW) { scr[32 + warp] = ic; scr[64 + warp] = id; scr[96 + warp] = il; }
    __syncthreads();
    int pc = 0, pd = 0, pl = 0;
    for (int wv = 0; wv < warp && wv < SEGW; wv++) { pc += scr[32 + wv]; pd += scr[64 + wv]; if (scr[96 + wv]) pl = scr[96 + wv]; }
    ic += pc; id += pd; if (!il) il = pl;
    const int ec = ic - c, ed = id - d;                            // exclusive sums
    int el = __shfl_up_sync(FULL, il, 1);                          // last quantiser set before this segment
    if (lane == 0) el = pl;

    // ---- end segment: first on-chain segment with a stop code or that completes the sub-block ----
    if (on && (stop || ic >= remain)) atomicMin(&scr[132], s);
    __syncthreads();
    const int endSeg = scr[132];

    if (has) {
        uint32_t a = 0, b = 0;
        if (on && s <= endSeg) {
            int q = el ? (el & 31) : qcur;
            a = (uint32_t)(coefBase + ec) | ((uint32_t)ph << 16) | ((uint32_t)q << 20) | 0x80000000u;
            b = (uint32_t)(drawBase + ed);
        }
        sm.emitA[s] = a; sm.emitB[s] = b;
    }
    if (endSeg < (1 << 30)) {
        if (s == endSeg) {                                         // the owner walks the end segment token by token
            int q = el ? (el & 31) : qcur;
            int cc = ec, dd = ed;
            int p = endSeg * 8 + ph;
            int kind = -1, zyx = 0, ok = 1;
            for (;;) {
                uint32_t x = page_window(sm, p);
                Token t = classify(x);
                if (t.stop) {
                    kind = t.stop;
                    zyx = (x >> 8) & 0xfff;
                    p += t.len; break;
                }
                if (t.q >= 0) q = t.q;
                if (cc + t.coefs > remain) { ok = 0; break; }     // run overshoot: corrupt (ulcDecoder.c:127,139,154)
                cc += t.coefs; dd += t.draws; p += t.len;
                if (cc == remain) { kind = 0; break; }
            }
            sm.ctl[C_SEND] = endSeg; sm.ctl[C_ENDPOS] = p; sm.ctl[C_ENDKIND] = kind;
            sm.ctl[C_ENDCOEFS] = cc; sm.ctl[C_ENDDRAWS] = dd; sm.ctl[C_ENDQ] = q; sm.ctl[C_ENDZYX] = zyx;
            sm.ctl[C_OK] = ok;
        }
    } else if (s == PAGE_WORDS - 1) {
        sm.ctl[C_SEND] = -1;                                       // chain runs off this page
        sm.ctl[C_EXITPH] = (int)(minc & 7u); sm.ctl[C_SUMCOEFS] = ic; sm.ctl[C_SUMDRAWS] = id;
        sm.ctl[C_LASTQ] = il ? (il & 31) : qcur;
        sm.ctl[C_OK] = 1;
    }
}

// Emit the tokens of the chain resolved by resolve_chain into buf / the noise-run list.
template <int NCH>
__device__ void emit_chain(const Smem<NCH> &sm, int tp, int endSeg, int endPos, int maxRuns, uint32_t drawOrigin) {
    const int s = threadIdx.x;
    if (s >= PAGE_WORDS) return;
    uint32_t a = sm.emitA[s];
    if (!(a & 0x80000000u)) return;
    int idx = a & 0xffff, ph = (a >> 16) & 7, q = (a >> 20) & 31;
    if (s == (tp >> 3)) ph = tp & 7;
    uint32_t draw = sm.emitB[s];
    int p = s * 8 + ph;
    const int lim = (s == endSeg) ? endPos : (s + 1) * 8;
    float scale = quant_scale(q);
    while (p < lim) {
        uint32_t x = page_window(sm, p);
        Token t = classify(x);
        if (t.stop) break;
        if (t.q >= 0) { scale = quant_scale(t.q); }
        else if (t.draws) {
            int lv = (int)((x >> 13) & 7) + 1;                     // (X>>1)+1, X = 4th nybble
            int r = atomicAdd(&sm.ctl[C_NRUNS], 1);
            if (r < maxRuns) {
                NoiseRun nr;
                nr.start_n = (uint32_t)idx | ((uint32_t)t.coefs << 16);
                nr.amp = (float)(lv * lv) * scale * 0.25f;
                nr.draw = drawOrigin + draw; nr.decay = 1.0f;
                sm.runs[r] = nr;
            }
            draw += t.draws;
        } else if (t.coefs == 1 && (x & 15) >= 2) {
            int v = (int)(x & 15);
            v = (v ^ 8) - 8;
            v = v < 0 ? -(v * v) : v * v;
            sm.buf[idx] = (float)v * scale;
        }
        idx += t.coefs;
        p += t.len;
    }
}

// ------------------------------------------------------------------------------------------------
// NCHT = channels per stream, fixed at compile time for mono and stereo; 0 = taken from the arguments (3 and more channels:
// M/S pairs (0,1), (2,3), ..., an odd last channel stays unpaired - ulcDecoder.c:281-289).
template <int NCHT, bool PCM16, int NTHR>
__global__ void __launch_bounds__(NTHR, (NTHR <= 256 ? 3 : 1))
ulcb_decode_kernel(UlcbDecodeArgs args) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const UlcbTables &tb = args.tb;
    const int NCH = NCHT ? NCHT : args.nChan;
    const int N = tb.N, H = N >> 1;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int maxRuns = NCH * N / 16 + 8 * NCH;
    constexpr int MAXPAIRS = 4;                  // host picks NTHR >= NCH*N/16, so NCH*N/4 pairs fit 4 per thread

    Smem<NCHT> sm;
    {
        unsigned char *p = smraw;
        sm.buf  = reinterpret_cast<float *>(p);     p += sizeof(float) * NCH * N;
        sm.pend = reinterpret_cast<float *>(p);     p += sizeof(float) * NCH * H;
        sm.page = reinterpret_cast<uint32_t *>(p);  p += sizeof(uint32_t) * (PAGE_WORDS + 4);
        sm.info = reinterpret_cast<uint32_t *>(p);  p += sizeof(uint32_t) * 8 * INFO_ROW;
        sm.emitA = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * PAGE_WORDS;
        sm.emitB = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * PAGE_WORDS;
        sm.runs = reinterpret_cast<NoiseRun *>(p);  p += sizeof(NoiseRun) * maxRuns;
        sm.ctl  = reinterpret_cast<int *>(p);       p += sizeof(int) * 16;
        sm.scan = reinterpret_cast<int *>(p);       p += sizeof(int) * 136;
        sm.pmap = reinterpret_cast<uint16_t *>(p);  p += sizeof(uint16_t) * INFO_ROW;
        sm.lq   = reinterpret_cast<uint8_t *>(p);
    }

    for (int stream = blockIdx.x; stream < args.nStreams; stream += gridDim.x) {
        const uint8_t *src = args.bytes + args.streamOffs[stream];
        const long long srcLen = (long long)(args.streamOffs[stream + 1] - args.streamOffs[stream]);
        const int nBlocks = args.nBlocks ? args.nBlocks[stream] : args.nBlocksAll;
        const uintptr_t mis = reinterpret_cast<uintptr_t>(src) & 3;
        const uint32_t *abase = reinterpret_cast<const uint32_t *>(src - mis);
        const long long nWordsValid = ((long long)mis + srcLen + 3) >> 2;
        const long long endNyb = 2 * ((long long)mis + srcLen);      // aligned-space nybble index of stream end
        UlcbStreamState *st = args.state ? args.state + stream : nullptr;

        // ---- stream state ----
        long long pos = 2 * (long long)mis;      // aligned-space nybble index of the next block
        int lastSub = 0;
        uint32_t drawBase = 0;                   // noise draws consumed so far in this stream
        int deadAt = -1;
        if (st && st->valid) {
            lastSub = st->lastSubBlockSize; drawBase = st->draws;
            for (int i = tid; i < NCH * H; i += nthr) sm.pend[i] = args.statePend[(size_t)stream * NCH * H + i];
        } else {
            for (int i = tid; i < NCH * H; i += nthr) sm.pend[i] = 0.0f;
        }
        long long curPage = -1;
        __syncthreads();
#ifdef ULCB_PHASE_TIMING
        long long phaseAcc[8] = {0}; long long tPrev = clock64();
#endif

        for (int blk = 0; blk < nBlocks; blk++) {
            void *outBlk = PCM16
                ? (void *)(reinterpret_cast<int16_t *>(args.out) + ((size_t)stream * args.outStride + (size_t)blk * N * NCH))
                : (void *)(reinterpret_cast<float *>(args.out) + ((size_t)stream * args.outStride + (size_t)blk * N * NCH));
            if (deadAt < 0 && pos >= endNyb) deadAt = blk;           // stream exhausted
            bool live = (deadAt < 0);
            int wc = 0x10;
            const long long blkStart = pos;
            uint32_t rawBits = 0;

            if (live) {
                // ---------------- parse ----------------
                for (int i = tid; i < NCH * N / 4; i += nthr) reinterpret_cast<float4 *>(sm.buf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tid == 0) sm.ctl[C_NRUNS] = 0;
                auto ensure = [&](long long q) {           // make q's page current (uniform across the CTA)
                    long long pg = q / PAGE_NYB;
                    if (pg != curPage) { curPage = pg; load_page<NCHT>(sm, abase, pg, nWordsValid); }
                };
                auto read_nyb = [&](long long q) -> int { ensure(q); return page_nyb<NCHT>(sm, (int)(q - curPage * PAGE_NYB)); };
                __syncthreads();
                int h0 = read_nyb(pos++);
                if (h0 & 8) wc = h0 | (read_nyb(pos++) << 4); else wc = h0 | 0x10;
                const unsigned layout = kLayout[wc >> 4];
                int coefBase = 0;
                uint32_t blkDraws = 0;
                bool ok = true;
                for (int c = 0; c < NCH && ok; c++) {
                    for (unsigned pat = layout;; ) {
                        const int S = N >> (pat & 7);
                        // sub-block header: bare quantiser (ulcDecoder.c:103-109)
                        int q = read_nyb(pos++);
                        bool silent = false;
                        if (q == 15) q = 30;                           // invalid lead-in; x86 shift semantics -> scale 0
                        else if (q == 14) { q += read_nyb(pos++); if (q == 29) silent = true; }
                        if (!silent) {
                            int remain = S, sub0 = coefBase;
                            for (;;) {
                                DTK(0);
                                ensure(pos);
                                int tp = (int)(pos - curPage * PAGE_NYB);
                                __syncthreads();
                                DTK(5);
                                resolve_chain<NCHT>(sm, tp, remain, q, sub0 + (S - remain), (int)blkDraws);
                                __syncthreads();
                                DTK(6);
                                int sEnd = sm.ctl[C_SEND];
#ifdef ULCB_DEBUG_PARSE
                                if (tid == 0 && args.dbgCoefs && blk <= 2)
                                    printf("blk %d ch %d S %d page %lld tp %d remain %d q %d base %d | sEnd %d endpos %d kind %d cdone %d ddone %d ok %d exit %d sumc %d\n",
                                           blk, c, S, curPage, tp, remain, q, sub0 + (S - remain), sEnd, sm.ctl[C_ENDPOS], sm.ctl[C_ENDKIND],
                                           sm.ctl[C_ENDCOEFS], sm.ctl[C_ENDDRAWS], sm.ctl[C_OK], sm.ctl[C_EXITPH], sm.ctl[C_SUMCOEFS]);
#endif
                                if (!sm.ctl[C_OK]) { ok = false; break; }
                                emit_chain<NCHT>(sm, tp, sEnd < 0 ? PAGE_WORDS : sEnd, sm.ctl[C_ENDPOS], maxRuns, drawBase);
                                DTK(7);
                                if (sEnd >= 0) {
                                    const int kind = sm.ctl[C_ENDKIND];
                                    const int cdone = sm.ctl[C_ENDCOEFS];      // coefficients consumed by this page's part of the chain
                                    pos = curPage * PAGE_NYB + sm.ctl[C_ENDPOS];
                                    blkDraws += (uint32_t)sm.ctl[C_ENDDRAWS];
                                    if (kind == 2) {                    // Fh,Fh,Z,Y,X: decaying noise to the end (ulcDecoder.c:172-186)
                                        const int left = remain - cdone;
                                        const int zyx = sm.ctl[C_ENDZYX];
                                        const int lv = (zyx & 15) + 1;
                                        const int nn = (((zyx >> 4) & 15) << 4) | ((zyx >> 8) & 15);
                                        if (tid == 0 && left > 0) {
                                            int r = atomicAdd(&sm.ctl[C_NRUNS], 1);
                                            NoiseRun nr;
                                            nr.start_n = (uint32_t)(sub0 + (S - remain) + cdone) | ((uint32_t)left << 16);
                                            nr.amp = (float)(lv * lv) * quant_scale(sm.ctl[C_ENDQ]) * (1.0f / 16);
                                            nr.draw = drawBase + blkDraws;
                                            nr.decay = 1.0f + (float)(nn * nn) * -0x1.0p-19f;
                                            if (r < maxRuns) sm.runs[r] = nr;
                                        }
                                        blkDraws += (uint32_t)left;
                                    }
                                    __syncthreads();
                                    break;
                                }
                                // chain continues on the next page
                                int consumed = sm.ctl[C_SUMCOEFS];
                                remain -= consumed; blkDraws += (uint32_t)sm.ctl[C_SUMDRAWS]; q = sm.ctl[C_LASTQ];
                                int ex = sm.ctl[C_EXITPH];
                                pos = (curPage + 1) * PAGE_NYB + ex;
                                __syncthreads();
                                if (pos >= endNyb + 16) { ok = false; break; }   // ran off the stream
                            }
                        }
                        coefBase += S;
                        pat >>= 4;
                        if (!pat || !ok) break;
                    }
                }
                // a block whose bytes run past the end of the stream (it parsed from the zero padding) is not decodable; neither
                // is one whose noise draws lie beyond the sign table the context was created for (maxBlocksPerStream)
                if (ok && (pos > endNyb || (unsigned long long)drawBase + blkDraws + 128ull > tb.nSignBits)) ok = false;
                if (!ok) { live = false; deadAt = blk; }
                else {
                    drawBase += blkDraws;
                    rawBits = (uint32_t)(pos - blkStart) * 4u;          // what ULC_DecodeBlock returns (ulcDecoder.c:301)
                    pos = (pos + 1) & ~1ll;                              // blocks are byte aligned (ulcEncoder_Encode.c:357-358)
                }
            }
            if (args.blockBits) {
                if (tid == 0) args.blockBits[(size_t)stream * args.bitsStride + blk] = live ? rawBits : 0u;
            }

            if (!live) {
                // corrupt / exhausted stream: silence (the reference tool aborts here, ulcDecodeTool.c:153-157)
                if (PCM16) { for (int i = tid; i < N * NCH / 2; i += nthr) reinterpret_cast<uint32_t *>(outBlk)[i] = 0u; }
                else       { for (int i = tid; i < N * NCH; i += nthr) reinterpret_cast<float *>(outBlk)[i] = 0.0f; }
                continue;
            }

            // ---------------- noise fill ----------------
            __syncthreads();
            DTK(0);
            {
                const int nRuns = min(sm.ctl[C_NRUNS], maxRuns);
                const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
                const uint32_t *__restrict__ par = tb.signParity;
                for (int r = warp; r < nRuns; r += nwarps) {
                    NoiseRun nr = sm.runs[r];
                    int start = nr.start_n & 0xffff, n = nr.start_n >> 16;
                    uint32_t g = nr.draw;
                    uint32_t p0 = (__ldg(&par[g >> 5]) >> (g & 31)) & 1u;
                    // A decaying tail's level is the reference's own running product p *= r in f32 (ulcDecoder.c:178-184), not
                    // amp * r^i: with r = 1 - n^2 2^-19 close to 1 the rounding of p * r has the same sign for thousands of steps
                    // and the sequence drifts 1e-4 away from the closed form.  Every lane runs the chain (lockstep: the cost
                    // of one chain), 32 steps between its samples.
                    const bool decays = nr.decay != 1.0f;
                    float a = nr.amp;
                    if (decays) for (int k = 0; k < lane; k++) a = __fmul_rn(a, nr.decay);
                    for (int i = lane; i < n; i += 32) {
                        uint32_t m = g + (uint32_t)i + 1u;
                        uint32_t sgn = ((__ldg(&par[m >> 5]) >> (m & 31)) & 1u) ^ p0;
                        sm.buf[start + i] = sgn ? -a : a;
                        if (decays) {
#pragma unroll 8
                            for (int k = 0; k < 32; k++) a = __fmul_rn(a, nr.decay);
                        }
                    }
                }
            }
            __syncthreads();
            DTK(1);
            if (args.dbgCoefs) {
                float *dc = args.dbgCoefs + ((size_t)stream * args.nBlocksAll + blk) * NCH * N;
                for (int i = tid; i < NCH * N; i += nthr) dc[i] = sm.buf[i];
            }

            // ---------------- inverse transform, sub-block by sub-block ----------------
            {
                int off = 0, j = 0;
                for (unsigned pat = kLayout[wc >> 4];; j++) {
                    const int sh = pat & 7, S = N >> sh, hS = S >> 1, lgHS = tb.log2N - sh - 1;
                    dct4_inplace<MAXPAIRS>(sm.buf + off, N, NCH, S, sh, tb);
                    DTK(2);

                    int ov = S;
                    if (pat & 8) ov >>= (wc & 7);
                    if (ov > lastSub) ov = lastSub;
                    lastSub = S;
                    const int flat = (S - ov) >> 1, held = H - hS;
                    const float2 *__restrict__ win = tb.win + (ov >> 1);

                    // un-fold (Fourier_IMDCT_Template.h:66-93) against the last S/2 pending samples
                    constexpr int MAXU = MAXPAIRS * 2;
                    float uLo[MAXU], uHi[MAXU], tn[MAXU];
#pragma unroll
                    for (int r = 0; r < MAXU; r++) {
                        int idx = tid + r * nthr;
                        if (idx < NCH * hS) {
                            int c = idx >> lgHS, i = idx & (hS - 1);
                            const float *t = sm.buf + c * N + off;
                            float a = sm.pend[c * H + held + i], b = t[hS + i];
                            if (i < flat) { uLo[r] = a; uHi[r] = b; }
                            else {
                                float2 w = __ldg(&win[i - flat]);
                                uLo[r] = w.x * a - w.y * b;
                                uHi[r] = w.y * a + w.x * b;
                            }
                            tn[r] = t[hS - 1 - i];
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int r = 0; r < MAXU; r++) {
                        int idx = tid + r * nthr;
                        if (idx < NCH * hS) {
                            int c = idx >> lgHS, i = idx & (hS - 1);
                            float *t = sm.buf + c * N + off;
                            t[i] = uLo[r]; t[S - 1 - i] = uHi[r];
                            sm.pend[c * H + held + i] = tn[r];
                        }
                    }
                    __syncthreads();

                    if (S != N) {
                        // Delay line (ulcDecoder.c:248-272): [pend[0..held) ++ u[0..S)] rotated left by S,
                        // done in place with three reversals.
                        const int tot = held + S;
                        auto at = [&](int c, int p) -> float & { return (p < held) ? sm.pend[c * H + p] : sm.buf[c * N + off + (p - held)]; };
                        auto reverse = [&](int lo, int hi) {     // reverse [lo, hi)
                            int np = (hi - lo) >> 1;
                            for (int idx = tid; idx < NCH * np; idx += nthr) {
                                int c = idx / np, k = idx - c * np;
                                float &x = at(c, lo + k), &y = at(c, hi - 1 - k);
                                float tmp = x; x = y; y = tmp;
                            }
                            __syncthreads();
                        };
                        reverse(0, S); reverse(S, tot); reverse(0, tot);
                    }
                    off += S;
                    pat >>= 4;
                    if (!pat) break;
                }
            }

            DTK(3);
            // ---------------- M/S undo + interleave + store ----------------
            if (NCH == 2) {
                for (int i = tid; i < N / 4; i += nthr) {
                    float4 m = reinterpret_cast<const float4 *>(sm.buf)[i];
                    float4 s = reinterpret_cast<const float4 *>(sm.buf + N)[i];
                    float l0 = m.x + s.x, r0 = m.x - s.x, l1 = m.y + s.y, r1 = m.y - s.y;
                    float l2 = m.z + s.z, r2 = m.z - s.z, l3 = m.w + s.w, r3 = m.w - s.w;
                    if (PCM16) {
                        auto cv = [](float v) -> uint32_t {
                            v = fminf(fmaxf(v * 32768.0f, -32768.0f), 32767.0f);
                            return (uint32_t)(uint16_t)(int16_t)__float2int_rn(v);
                        };
                        uint4 o;
                        o.x = cv(l0) | (cv(r0) << 16); o.y = cv(l1) | (cv(r1) << 16);
                        o.z = cv(l2) | (cv(r2) << 16); o.w = cv(l3) | (cv(r3) << 16);
                        reinterpret_cast<uint4 *>(outBlk)[i] = o;
                    } else {
                        float4 *o = reinterpret_cast<float4 *>(outBlk) + 2 * i;
                        o[0] = make_float4(l0, r0, l1, r1);
                        o[1] = make_float4(l2, r2, l3, r3);
                    }
                }
            } else {
                // any channel count: output sample i = n * NCH + c
                for (int i = tid; i < N * NCH; i += nthr) {
                    const int n = i / NCH, c = i - n * NCH;
                    float v = sm.buf[c * N + n];
                    if ((c | 1) < NCH) {                                 // paired channel: undo M/S
                        const float m = sm.buf[(c & ~1) * N + n], sd = sm.buf[(c | 1) * N + n];
                        v = (c & 1) ? m - sd : m + sd;
                    }
                    if (PCM16) {
                        v = fminf(fmaxf(v * 32768.0f, -32768.0f), 32767.0f);
                        reinterpret_cast<int16_t *>(outBlk)[i] = (int16_t)__float2int_rn(v);
                    } else reinterpret_cast<float *>(outBlk)[i] = v;
                }
            }
            __syncthreads();
            DTK(4);
        }
#ifdef ULCB_PHASE_TIMING
        if (tid == 0) { for (int i = 0; i < 8; i++) atomicAdd(&g_ulcbDecPhase[i], (unsigned long long)phaseAcc[i]); atomicAdd(&g_ulcbDecPhase[8], (unsigned long long)nBlocks); }
#endif

        // ---- save stream state / status ----
        if (st) {
            for (int i = tid; i < NCH * H; i += nthr) args.statePend[(size_t)stream * NCH * H + i] = sm.pend[i];
            if (tid == 0) { st->valid = 1; st->lastSubBlockSize = lastSub; st->draws = drawBase; }
        }
        if (tid == 0) {
            if (args.status) args.status[stream] = deadAt;               // -1 = every block decoded
            if (args.bytesConsumed) args.bytesConsumed[stream] = (uint64_t)((pos - 2 * (long long)mis) >> 1);
        }
        __syncthreads();
    }
}

}  // namespace

size_t ulcb_decode_smem_bytes(int N, int nChan) {
    size_t maxRuns = (size_t)nChan * N / 16 + 8 * nChan;
    return sizeof(float) * nChan * N + sizeof(float) * nChan * (N / 2) + sizeof(uint32_t) * (PAGE_WORDS + 4) +
           sizeof(uint32_t) * 8 * INFO_ROW + 2 * sizeof(uint32_t) * PAGE_WORDS + sizeof(NoiseRun) * maxRuns +
           sizeof(int) * (16 + 136) + sizeof(uint16_t) * INFO_ROW + 8 * INFO_ROW + 16;
}

template <int NCH, int NTHR>
static cudaError_t launch_t(const UlcbDecodeArgs &args, int pcm16, int gridBlocks, size_t smem, cudaStream_t stream) {
    void (*k)(UlcbDecodeArgs) = pcm16 ? ulcb_decode_kernel<NCH, true, NTHR> : ulcb_decode_kernel<NCH, false, NTHR>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    k<<<gridBlocks, NTHR, smem, stream>>>(args);
    return cudaGetLastError();
}

int ulcb_decode_threads(int N, int nChan) {
    int t = 256;
    while (t < nChan * N / 16) t <<= 1;          // >= nChan*N/16 (4 pairs per thread in the DCT), one of 256 / 512 / 1024 / (too many)
    return t;
}

cudaError_t ulcb_decode_launch(const UlcbDecodeArgs &args, int nChan, int pcm16, int gridBlocks, cudaStream_t stream) {
    const int N = args.tb.N;
    static const bool forceGeneric = getenv("ULCB_DECODE_GENERIC") != nullptr;
    if (!forceGeneric && ulcb_decode_ws_supported(args, nChan)) {
        int dev = 0, sms = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        return ulcb_decode_ws_launch(args, pcm16, sms, stream);
    }
    size_t smem = ulcb_decode_smem_bytes(N, nChan);
    int nthr = ulcb_decode_threads(N, nChan);
    if (nChan == 2) {
        if (nthr == 256)  return launch_t<2, 256>(args, pcm16, gridBlocks, smem, stream);
        if (nthr == 512)  return launch_t<2, 512>(args, pcm16, gridBlocks, smem, stream);
        if (nthr == 1024) return launch_t<2, 1024>(args, pcm16, gridBlocks, smem, stream);
    } else if (nChan == 1) {
        if (nthr == 256)  return launch_t<1, 256>(args, pcm16, gridBlocks, smem, stream);
        if (nthr == 512)  return launch_t<1, 512>(args, pcm16, gridBlocks, smem, stream);
        if (nthr == 1024) return launch_t<1, 1024>(args, pcm16, gridBlocks, smem, stream);
    } else if (nChan == args.nChan) {
        if (nthr == 256)  return launch_t<0, 256>(args, pcm16, gridBlocks, smem, stream);
        if (nthr == 512)  return launch_t<0, 512>(args, pcm16, gridBlocks, smem, stream);
        if (nthr == 1024) return launch_t<0, 1024>(args, pcm16, gridBlocks, smem, stream);
    }
    return cudaErrorInvalidValue;
}
