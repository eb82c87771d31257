// ulcb_decode_bp.cu - block-parallel batched decoder (BlockSize 2048, stereo): the benchmark path of configs[1].
//
// libulc/ulcDecoder.c:198-302 decodes a stream strictly block after block, but every dependency between blocks is
// either known from a block-size index or a cheap scan:
//
//   byte offset of block k+1     a block-size index: the one this library's encoder emits (dBlockBytes), the caller's,
//                                or - for foreign .ulc payloads, which carry none (tools/ulc_Helper.h:10-20) - the
//                                token-length walk of bp_walk_kernel (what tools/ulcDecodeTool.c:153-164 does by decoding)
//   position in the noise RNG    per-block draw counts (parse) -> exclusive scan per stream -> sign-parity table
//   LastSubBlockSize             the previous block's window-control nybble
//   half-block lap               a function of the previous block's coefficients ONLY (the un-fold keeps the first S/2
//                                DCT outputs, Fourier_IMDCT_Template.h:66-93; a window-switched block's delay line holds
//                                nothing older than the block itself), so a run of consecutive blocks can start anywhere
//                                after re-doing ONE block
//
// Pipeline (per wave of streams, all on the context's stream):
//   bp_index_kernel / bp_walk_kernel   block byte offsets per stream
//   bp_parse_kernel      ONE LANE PER BLOCK walks the nybble grammar (ulcDecoder.c:99-197) sequentially - about 1000 tokens of
//                        ~50 instructions instead of the 8-phase speculative resolver - and writes the block's content in
//                        sparse form to an HBM arena: one 32-bit word per coded coefficient (position | bf16 value: every
//                        value the format can express is n * 2^-e with n <= 64, exact in bf16), one 16-byte record per noise
//                        run / decaying tail, plus a 16-byte descriptor (bits consumed, window control, draw count)
//   bp_scan_kernel       per stream: exclusive scan of the draw counts, first undecodable block, status / blockBits out
//   bp_transform_kernel  one 64-thread CTA (Mid warp + Side warp) per (stream, run of consecutive blocks): the block's words
//                        arrive by TMA bulk copy (cp.async.bulk + mbarrier) while the previous block is transformed, are
//                        scattered into the FFT tile, noise runs are expanded eight at a time (4 lanes each) from the
//                        sign-parity table, then DCT-IV as 32 x 32 register FFT, un-fold against the lane-private pending
//                        tail, M/S undo, PCM16 / float stores.
//
// Traffic per block: bytes in (~0.74 KB) + sparse words written and read back (~2 x 3 KB) + PCM out (8 KB).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "ulcb_dct.cuh"
#include "ulcb_decode.h"

namespace {

constexpr int N = 2048, H = N / 2, NCH = 2;
constexpr int ZROW = 33;                              // padded row of the 32 x 32 complex work tile
constexpr int RECSLOT = 1024;                         // words of the TMA landing slot per warp
constexpr unsigned FULLM = 0xffffffffu;
#ifndef ULCB_BP_PFDIST
#define ULCB_BP_PFDIST 4
#endif
constexpr int PFDIST = ULCB_BP_PFDIST;                // parse: 16-byte vectors the L1 prefetch runs ahead of the loads

// Decimation layouts, one nybble per sub-block, LSB first: bits0-2 shift, bit3 transient
// (libulc/ulcHelper.h:24-46, FormatSpecs.md:35-51)
__constant__ uint16_t kLayoutBp[16] = {
    0x0000, 0x0008, 0x0019, 0x0091, 0x012A, 0x01A2, 0x02A1, 0x0A21,
    0x123B, 0x12B3, 0x13B2, 0x1B32, 0x23B1, 0x2B31, 0x3B21, 0xB321,
};

__device__ __forceinline__ float quant_scale(int q) {      // ulcDecoder.c:96-98
    return 0x1.0p-31f * (float)((1u << 26) >> (q & 31));
}

// arena offset (bytes) of the record region of the block that starts at byte `a` of the wave's payload and has global
// index `gidx` in the wave: 8 bytes per payload byte (one word per nybble bounds the words a block can produce: a coefficient
// costs one nybble and one word, a run four nybbles and four words) plus 32 bytes of slack per block, 16-byte aligned.
__device__ __forceinline__ uint64_t region_start(uint64_t a, uint64_t gidx) { return 16ull * ((a + 1) >> 1) + 32ull * gidx; }
__device__ __forceinline__ uint64_t region_bytes(uint32_t len) { return 16ull * ((len + 1u) >> 1) + 16ull; }

// ================================================================================================
// Sequential token walk of one block by one lane (ulcDecoder.c:99-197, 211-231)
// ================================================================================================
struct ParseOut { uint32_t nyb, wc, draws, nCoef0, nCoefTot, nRun0, nRunTot; };

// One token per loop trip and lane.  32 lanes walk 32 different blocks, so every branch of the loop body that a few percent
// of the tokens take is executed on almost every trip; the body is therefore written branch-free for the four token
// kinds that make up 99.5 % of a stream (coefficient, short zero run, noise run, quantiser change) and branches only
// for the rest (sub-block ends and headers, long zero runs, extended quantisers, stop codes, tails).
template <bool EMIT>
__device__ __forceinline__ bool parse_block(const uint8_t *p, const uint32_t *wend, uint32_t maxNyb, const uint16_t *lut,
                                            uint32_t *up, uint4 *down, ParseOut &o) {
    // Nybble reader.  A lane reads its block front to back, but 32 lanes of a warp read 32 different blocks: a load that
    // misses stalls the whole warp, and with one miss per lane every ~40 tokens that is a stall on nearly every trip.  So the
    // bytes come in 16-byte loads that are issued a full 16 bytes (~20 tokens) before their first word is needed and wait
    // in a register FIFO: `cur` holds the current nybble at bit `sh`, `nxt` the following word, q0..q3 the next `cnt` words,
    // `fly` the 16 bytes in flight.
    const uintptr_t pa = reinterpret_cast<uintptr_t>(p);
    const uint4 *vbase = reinterpret_cast<const uint4 *>(pa & ~(uintptr_t)15);
    const long long vavail = (reinterpret_cast<const char *>(wend) - reinterpret_cast<const char *>(vbase)) >> 2;   // readable words from vbase
    const uint32_t wlim = vavail <= 0 ? 0u : (uint32_t)min(vavail, 0x7ffffff0ll);
    // The in-flight registers are loaded IN PLACE by inline assembly ("+r"): assigned as a C++ value inside the refill branch
    // they are copied at the loop's join, and that copy waits for the load on the very trip that issued it.
    uint32_t f0 = 0, f1 = 0, f2 = 0, f3 = 0;
    auto ld16 = [&](uint32_t i) {                           // 16 bytes number i from vbase -> f0..f3, zero beyond the readable end
        if (4u * i + 4u <= wlim) {
            asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "+r"(f0), "+r"(f1), "+r"(f2), "+r"(f3) : "l"(vbase + i));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(vbase + i + PFDIST));     // may point past the buffer: a prefetch does not fault
        } else {
            // the last, partial 16 bytes of the buffer: word by word, still in place
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.lt.u32 p, %4, %7;\n\t@p ld.global.nc.u32 %0, [%8];\n\t@!p mov.u32 %0, 0;\n\t"
                "setp.lt.u32 p, %5, %7;\n\t@p ld.global.nc.u32 %1, [%8+4];\n\t@!p mov.u32 %1, 0;\n\t"
                "setp.lt.u32 p, %6, %7;\n\t@p ld.global.nc.u32 %2, [%8+8];\n\t@!p mov.u32 %2, 0;\n\t"
                "mov.u32 %3, 0;\n\t}"
                : "+r"(f0), "+r"(f1), "+r"(f2), "+r"(f3)
                : "r"(4u * i), "r"(4u * i + 1u), "r"(4u * i + 2u), "r"(wlim), "l"(vbase + i));
        }
    };
    ld16(0);
    uint32_t cur = f0, nxt = f1, q0 = f2, q1 = f3, q2 = 0, q3 = 0, cnt = 2, vi = 2;
    ld16(1);
#define BP_ADVANCE()                                                                         \
    do {                                                                                     \
        cur = nxt; nxt = q0; q0 = q1; q1 = q2; q2 = q3; cnt--;                               \
        if (cnt == 0) { q0 = f0; q1 = f1; q2 = f2; q3 = f3; cnt = 4; ld16(vi); vi++; }       \
    } while (0)
    for (uint32_t k0 = (uint32_t)((pa & 15) >> 2); k0; k0--) BP_ADVANCE();
    uint32_t sh = 8u * (uint32_t)(pa & 3);
    int left = (int)maxNyb;                                // nybbles this block may still consume
#define BP_SKIP(k)                                                                           \
    do {                                                                                     \
        const uint32_t k_ = (uint32_t)(k);                                                   \
        sh += 4u * k_; left -= (int)k_;                                                      \
        if (sh >= 32u) { sh -= 32u; BP_ADVANCE(); }                                          \
    } while (0)
    const unsigned lutAddr = (unsigned)__cvta_generic_to_shared(lut);
    auto lutAt = [&](uint32_t i) -> uint32_t {             // one shared-memory load off a register-held base
        unsigned short v;
        asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(lutAddr + 2u * i));
        return (uint32_t)v;
    };
    uint32_t e0 = 0, e1 = 0, e2 = 0, nw = 0;
    // coefficient words leave four at a time (16-byte stores): every store instruction of this kernel touches 32 different
    // sectors, and the L1 store path, shared with the loads, takes them one at a time
#define BP_EMITW(val)                                                                        \
    do {                                                                                     \
        const uint32_t v_ = (val), s_ = nw & 3u;                                             \
        if (s_ == 3u) *reinterpret_cast<uint4 *>(up + (nw - 3)) = make_uint4(e0, e1, e2, v_); \
        e0 = s_ == 0u ? v_ : e0; e1 = s_ == 1u ? v_ : e1; e2 = s_ == 2u ? v_ : e2;           \
        nw++;                                                                                \
    } while (0)

    uint32_t x = __funnelshift_r(cur, nxt, sh);
    uint32_t wc;
    if (x & 8u) { wc = x & 0xffu; BP_SKIP(2); }            // ulcDecoder.c:211-216
    else        { wc = (x & 15u) | 0x10u; BP_SKIP(1); }
    const uint32_t layout = kLayoutBp[wc >> 4];
    uint32_t pat = layout;
    int c = 0, pos = 0, subEnd = 0;
    bool ok = true, done = false;
    int q = 0, eb = -1;                                     // eb = bf16 exponent bias of the quantiser, -1 = scale 0
    uint32_t draws = 0, nRun = 0, nCoef0 = 0, nRun0 = 0;
    // sub-block boundary: next sub-block / channel (unless `first`), then its header - a bare quantiser (ulcDecoder.c:103-112);
    // repeats while the header says "silent" (Eh,Fh).  Sets `done` after the last sub-block of the last channel.
    auto boundary = [&](bool first) {
        for (;;) {
            if (!first) {
                pat >>= 4;
                if (!pat) {
                    if (c == 0) { nCoef0 = nw; nRun0 = nRun; }
                    if (++c == NCH) { done = true; return; }
                    pat = layout; pos = 0;
                }
            }
            first = false;
            subEnd = pos + (N >> (pat & 7));
            const uint32_t hx = __funnelshift_r(cur, nxt, sh);
            const uint32_t h0 = hx & 15u, h1 = (hx >> 4) & 15u;
            int len = 1;
            if (h0 == 15u) q = 30;                          // a lone Fh is an invalid lead-in: the reference's shift by -2 gives scale 0 on x86
            else if (h0 == 14u) { len = 2; if (h1 == 15u) pos = subEnd; else q = 14 + (int)h1; }   // Eh,Fh: silent sub-block
            else q = (int)h0;
            eb = q > 26 ? -1 : (5 + q) << 7;
            BP_SKIP(len);
            if (left < 0) { ok = false; done = true; return; }
            if (pos != subEnd) return;
        }
    };
    boundary(true);
    while (!done) {
        x = __funnelshift_r(cur, nxt, sh);
        const uint32_t v0 = x & 15u, v1 = (x >> 4) & 15u, v2 = (x >> 8) & 15u, v3 = (x >> 12) & 15u;
        const bool isC = !((0x8103u >> v0) & 1u), isZ = v0 == 0u, isN = v0 == 8u, isQ = v0 == 15u && v1 < 14u;
        const int n8 = (int)((v1 << 5) | (v2 << 1) | (v3 & 1u)) + 16;
        int adv = isC ? 1 : (isZ ? (int)v1 + 1 : (isN ? n8 : 0));
        int len = isC ? 1 : (isN ? 4 : 2);
        if (!(isC || isZ || isN || isQ)) {
            // ---- rare tokens ----
            if (v0 == 1u) { len = 3; adv = (int)((v1 << 4) | v2) + 33; }                          // 1h,Yh,Xh: 33..288 zeros
            else if (v1 == 14u) {
                len = 3;
                if (v2 == 15u) adv = subEnd - pos;                                                // Fh,Eh,Fh: zeros to the end
                else { q = 14 + (int)v2; eb = q > 26 ? -1 : (5 + q) << 7; }                     // Fh,Eh,0h..Ch (Dh,Eh: scale 0)
            } else {                                                                              // Fh,Fh,Zh,Yh,Xh: decaying noise to the end
                len = 5; adv = subEnd - pos;
                if (EMIT) *(down - 1 - nRun) = make_uint4((uint32_t)pos, (uint32_t)adv, draws,
                                                          0x80000000u | ((x >> 8) & 0xfffu) | ((uint32_t)q << 12));
                nRun++;
                draws += (uint32_t)adv;
            }
        }
        if (adv > subEnd - pos) { ok = false; break; }     // run overshoot: corrupt (ulcDecoder.c:127,139,154)
        if (isQ) { q = (int)v1; eb = (5 + q) << 7; }       // Fh,0h..Dh
        if (EMIT) {
            // -7..-2, +2..+7: sign * v^2 * Quant (ulcDecoder.c:119-124)
            if (isC && eb >= 0) BP_EMITW(((uint32_t)pos << 16) | (lutAt(v0) - (uint32_t)eb));
            // 8h,Zh,Yh,Xh: 16..527 noise, level ((X>>1)+1)^2 * Quant / 4
            if (isN) *(down - 1 - nRun) = make_uint4((uint32_t)pos, (uint32_t)adv, draws, eb >= 0 ? lutAt(16u + (v3 >> 1)) - (uint32_t)eb : 0u);
        }
        if (isN) { nRun++; draws += (uint32_t)adv; }
        pos += adv;
        BP_SKIP(len);
        if (left < 0) { ok = false; break; }               // ran off the bytes this block may use
        if (pos == subEnd) boundary(false);
    }
    if (EMIT && (nw & 3u)) *reinterpret_cast<uint4 *>(up + (nw & ~3u)) = make_uint4(e0, e1, e2, 0u);
#undef BP_SKIP
#undef BP_ADVANCE
#undef BP_EMITW
    o.nyb = maxNyb - (uint32_t)left; o.wc = wc; o.draws = draws; o.nCoef0 = nCoef0; o.nCoefTot = nw; o.nRun0 = nRun0; o.nRunTot = nRun;
    return ok;
}

// bf16 patterns before the quantiser's exponent (5 + q) is subtracted: [0..15] sign * v^2 for coefficient nybble v (two's
// complement), [16..23] lv^2 / 4 for noise level lv - 1
__device__ __forceinline__ void init_lut(uint16_t *lut) {
    const int t = threadIdx.x;
    if (t < 24) {
        float f;
        if (t < 16) { const int sv = (t ^ 8) - 8; f = (float)(sv < 0 ? -(sv * sv) : sv * sv); }
        else { const int lv = t - 15; f = (float)(lv * lv) * 0.25f; }
        lut[t] = (uint16_t)(__float_as_uint(f) >> 16);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// block offsets from a caller-supplied block-size index: one warp per stream, exclusive scan
// ------------------------------------------------------------------------------------------------
__global__ void bp_index_kernel(UlcbBpArgs a) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= a.nStreams) return;
    const int gs = a.s0 + s;
    const int nb = a.nBlocks ? (int)a.nBlocks[gs] : a.nBlocksAll;
    const uint32_t *bb = a.blockBytes + (size_t)gs * a.blockBytesStride;
    uint32_t *off = a.blkOff + (size_t)s * (a.maxBlk + 1);
    const uint64_t slen = a.streamOffs[gs + 1] - a.streamOffs[gs];
    uint64_t run = 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        uint32_t v = b < nb ? bb[b] : 0u;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t t = __shfl_up_sync(FULLM, inc, d); if (lane >= d) inc += t; }
        uint64_t ex = run + inc - v;
        if (b < nb) off[b] = (uint32_t)(ex < slen ? ex : slen);      // a block that starts at / beyond the end is undecodable
        run += __shfl_sync(FULLM, inc, 31);
    }
    if (lane == 0) off[nb] = (uint32_t)(run < slen ? run : slen);
}

// ---- mbarrier / TMA bulk-copy helpers ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
    unsigned done = 0;
    for (unsigned spins = 0; !done; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();                 // a copy that never lands is a bug, not a reason to hang the GPU
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load(void *dst, const void *src, unsigned bytes, unsigned long long *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// foreign streams (no index): one lane per stream walks the token lengths of the whole stream
// ------------------------------------------------------------------------------------------------
// A .ulc payload carries no block index (tools/ulc_Helper.h:10-20): where block k+1 starts is only known after the tokens of
// block k have been walked (what tools/ulcDecodeTool.c:153-164 does by decoding).  The walk is one dependent chain per stream,
// so its cost is the latency of ONE step times the steps of one stream, however many streams there are.  The step here is a
// table-driven automaton over the token grammar (ulcDecoder.c:113-186): nine states (token start; inside 0h,X / 1h,Y,X /
// 8h,Z,Y,X; after Fh; after Fh,Eh), and shared-memory tables that take a state and THREE nybbles to the next state (NEXT) and to
// the number of coefficients those nybbles cover (ADV; a multi-nybble run is accounted piecewise: 8h,Z,Y,X = (32 Z + 16) + 2 Y +
// (X & 1)).  Eight steps (24 nybbles = three words) are looked up back to back with no check in between, so the chain per step
// is one shared-memory load and one address add; the checks come after.  A group that would reach the end of its sub-block (or
// holds a stop code) commits only its leading steps; the rest is walked one nybble at a time, and a finished sub-block goes
// through the header / window-control code.  Lanes in those paths hold up their warp, so a call with few streams runs with few
// active lanes per warp (and more warps) - the SMs are otherwise empty.
constexpr int WALK_STATES = 9;
// ulcb_tables.h: walkTab = u64 MAP[4096] (for three nybbles: the next state of every entry state k, as 4 k' in bits 4k+2..4k+5),
// u16 ADV[9][4096] (coefficients the three nybbles cover from entry state k | stop flag << 15), then u32 T1[9][16] one-nybble
// steps (4 k' << 16 | TAIL << 12 | ZEROS << 11 | coefficients).  A state is carried as 4 k: the shift that extracts its field.
constexpr uint32_t WALK_MAP_BYTES = 4096 * 8, WALK_ADV_BYTES = WALK_STATES * 4096 * 2, WALK_T1_BYTES = WALK_STATES * 16 * 4;
constexpr uint32_t WALK_TAB_BYTES = WALK_MAP_BYTES + WALK_ADV_BYTES + WALK_T1_BYTES;
constexpr uint32_t WALK_FLAG = 0x800u, WALK_TAIL = 0x1000u;

// Each lane streams its own 1 MB-scale stream, and thousands of lanes doing so with 16-byte loads are bound by address
// translation and DRAM row misses, not by their arithmetic (measured: 160 ms for 4096 streams whatever the step cost).  So the
// bytes come in bulk: every lane owns a two-slot ring in shared memory and has the TMA engine (cp.async.bulk + one mbarrier per
// slot) fetch the next WALK_CHUNK contiguous bytes of its stream while it walks the current ones.
constexpr int WALK_CHUNK = 512, WALK_SLOTW = WALK_CHUNK / 4;          // bytes / words of stream per ring slot ...
constexpr int WALK_COPYW = WALK_SLOTW + 4;                             // ... each copied with the first 16 bytes of the next chunk, so that a
                                                                       // few words read ahead never leave the slot
constexpr int WALK_RINGW = 2 * WALK_COPYW;                             // words per lane
constexpr int WALK_MAXLANES = 96;                                      // lanes per CTA

// the last bytes of the whole input do not make a full bulk copy: word by word, zero beyond the readable end
__device__ __noinline__ void walk_fill_tail(uint32_t *dst, const uint32_t *ws, const uint32_t *wend, unsigned long long *mb) {
    for (int i = 0; i < WALK_COPYW; i++) dst[i] = (ws + i < wend) ? __ldg(ws + i) : 0u;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mb)) : "memory");
}

__global__ void __launch_bounds__(512, 1) bp_walk_kernel(UlcbBpArgs a, int lanes) {
    extern __shared__ __align__(16) uint32_t wtab[];        // ulcb_tables.h: walkTab
    uint32_t *rings = wtab + WALK_TAB_BYTES / 4;                                                   // [WALK_MAXLANES][WALK_RINGW]
    unsigned long long *mbars = reinterpret_cast<unsigned long long *>(rings + WALK_MAXLANES * WALK_RINGW);   // [WALK_MAXLANES][2]
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.tb.walkTab);
        uint4 *dst = reinterpret_cast<uint4 *>(wtab);
        for (int i = threadIdx.x; i < (int)(WALK_TAB_BYTES / 16); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    const int lane = threadIdx.x & 31;
    const int slotIdx = (threadIdx.x >> 5) * lanes + lane;  // this lane's ring
    const bool active = lane < lanes;
    if (active) {
        mbar_init(&mbars[2 * slotIdx], 1); mbar_init(&mbars[2 * slotIdx + 1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (!active) return;
    const int s = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes + lane;
    if (s >= a.nStreams) return;
    const int gs = a.s0 + s;
    const int nb = a.nBlocks ? (int)a.nBlocks[gs] : a.nBlocksAll;
    uint32_t *off = a.blkOff + (size_t)s * (a.maxBlk + 1);
    const uint64_t sbeg = a.streamOffs[gs], slen = a.streamOffs[gs + 1] - sbeg;
    const uint32_t endNyb = slen > 0x3fffffffull ? 0x7ffffffeu : (uint32_t)(2 * slen);

    // ---- the ring: chunk k of the stream (WALK_CHUNK bytes from the 16-byte aligned start) lives in slot k & 1 ----
    const uint8_t *p = a.bytes + sbeg;
    const uintptr_t pa = reinterpret_cast<uintptr_t>(p);
    const char *vbase = reinterpret_cast<const char *>(pa & ~(uintptr_t)15);
    const char *bulkEnd = reinterpret_cast<const char *>(reinterpret_cast<uintptr_t>(a.wend) & ~(uintptr_t)15);   // bulk copies stay below
    uint32_t *ring = rings + slotIdx * WALK_RINGW;
    unsigned long long *mb = &mbars[2 * slotIdx];
    const unsigned ringAddr = smem_u32(ring);
    auto fill = [&](uint32_t k) {
        const char *src = vbase + (size_t)k * WALK_CHUNK;
        uint32_t *dst = ring + (k & 1u) * WALK_COPYW;
        if (src + 4 * WALK_COPYW <= bulkEnd) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&mb[k & 1u], 4 * WALK_COPYW);
            tma_load(dst, src, 4 * WALK_COPYW, &mb[k & 1u]);
        } else walk_fill_tail(dst, reinterpret_cast<const uint32_t *>(src), a.wend, &mb[k & 1u]);
    };
    // Reader state: w0..w3 = the word that holds the current nybble (at bit sh) and the three after it; wpos = the word after w3,
    // rpos = its shared-memory address, inside chunk kc, which ends at word hk.  Reads are sequential: when the reader enters
    // chunk k, chunk k - 1 is done with and its slot takes chunk k + 1.
    uint32_t wpos = (uint32_t)((pa & 15) >> 2), kc = 0, hk = WALK_SLOTW, par = 0;
    fill(0); fill(1);
    mbar_wait(&mb[0], 0); par = 1u;
    unsigned rpos = ringAddr + 4u * wpos;
    auto cross = [&]() {                                    // wpos has reached hk: on to the next chunk
        kc++;
        mbar_wait(&mb[kc & 1u], (par >> (kc & 1u)) & 1u); par ^= 1u << (kc & 1u);
        fill(kc + 1u);
        rpos = ringAddr + 4u * ((kc & 1u) * WALK_COPYW + (wpos - kc * WALK_SLOTW));
        hk += WALK_SLOTW;
    };
    auto rd = [&](unsigned ofs) -> uint32_t {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(rpos + ofs));
        return v;
    };
    uint32_t w0 = rd(0), w1 = rd(4), w2 = rd(8), w3 = rd(12);
    wpos += 4u; rpos += 16u;
    uint32_t sh = 8u * (uint32_t)(pa & 3), nyb = 0;         // nyb = nybbles consumed since the start of the stream
#define WK_SKIP(k)                                                                           \
    do {                                                                                     \
        const uint32_t k_ = (uint32_t)(k);                                                   \
        sh += 4u * k_; nyb += k_;                                                            \
        if (sh >= 32u) { sh -= 32u; if (wpos >= hk) cross(); w0 = w1; w1 = w2; w2 = w3; w3 = rd(0); wpos++; rpos += 4u; } \
    } while (0)

    // start as if a block had just ended: the first trip takes the boundary path, which opens block 0
    const int WN = a.tb.N, WCH = a.nChan;                    // the walk is not tied to the pipeline's block size and channel count
    int b = -1, c = WCH - 1, pos = 0, subEnd = 0;
    uint32_t pat = 0, layout = 0, st = 0;                   // st = 4 * state
    bool dead = false;
    const char *tabB = reinterpret_cast<const char *>(wtab);
    // the group of eight three-nybble steps (24 nybbles = three words) at the reader's position: xs = 2 * (twelve bits of step j),
    // m = their state maps.  Carried from trip to trip: a trip prepares the following group while its own chain runs.
    uint32_t xs[8];
    uint2 m[8];
    auto window = [&](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t (&x)[8], uint2 (&mm)[8]) {
        x[0] = (a0 << 1) & 0x1ffeu;  x[1] = (a0 >> 11) & 0x1ffeu;  x[2] = (__funnelshift_r(a0, a1, 24) << 1) & 0x1ffeu;
        x[3] = (a1 >> 3) & 0x1ffeu;  x[4] = (a1 >> 15) & 0x1ffeu;  x[5] = (__funnelshift_r(a1, a2, 28) << 1) & 0x1ffeu;
        x[6] = (a2 >> 7) & 0x1ffeu;  x[7] = (a2 >> 19) & 0x1ffeu;
#pragma unroll
        for (int j = 0; j < 8; j++) mm[j] = *reinterpret_cast<const uint2 *>(tabB + 4u * x[j]);
    };
    window(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), xs, m);
    for (;;) {
        if (wpos >= hk) cross();
        const uint32_t n4 = rd(0), n5 = rd(4), n6 = rd(8);
        // the chain: two operations per step, no check in between
        uint32_t sj[9];
        sj[0] = st;
#pragma unroll
        for (int j = 0; j < 8; j++) sj[j + 1] = __funnelshift_rc(m[j].x, m[j].y, sj[j]) & 0x3cu;
        uint32_t av[8];
#pragma unroll
        for (int j = 0; j < 8; j++) av[j] = *reinterpret_cast<const uint16_t *>(tabB + WALK_MAP_BYTES + 2048u * sj[j] + xs[j]);
        // the following group, should this one stand
        uint32_t xsN[8];
        uint2 mN[8];
        window(__funnelshift_r(w3, n4, sh), __funnelshift_r(n4, n5, sh), __funnelshift_r(n5, n6, sh), xsN, mN);
        const uint32_t sum = ((av[0] + av[1] + av[2]) + (av[3] + av[4] + av[5])) + (av[6] + av[7]);   // stop flags land in bits 15-18
        if (sum < 0x8000u && pos + (int)sum < subEnd && nyb + 24u <= endNyb) {
            pos += (int)sum; st = sj[8]; nyb += 24u;
            w0 = w3; w1 = n4; w2 = n5; w3 = n6; wpos += 3u; rpos += 12u;
#pragma unroll
            for (int j = 0; j < 8; j++) { xs[j] = xsN[j]; m[j] = mN[j]; }
            continue;
        }
        // A group that holds a stop code or reaches the end of its sub-block: its leading steps stand ...
        {
            int jn = 0;
            bool go = true;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                go = go && av[j] < 0x8000u && pos + (int)av[j] < subEnd && nyb + 3u * (uint32_t)(j + 1) <= endNyb;
                if (go) { pos += (int)av[j]; st = sj[j + 1]; jn++; }
            }
#pragma unroll 1
            for (; jn > 0; jn--) WK_SKIP(3);
        }
        // ... and the rest goes nybble by nybble up to and through the end of the sub-block
        bool fin = false;
        for (;;) {
            if (st == 0u && pos >= subEnd) {
                if (pos > subEnd) { dead = true; fin = true; break; }       // run overshoot: corrupt (ulcDecoder.c:127,139,154)
                for (;;) {
                    pat >>= 4;
                    if (!pat) {
                        if (++c == WCH) {
                            // block end: it ends inside its last byte (ulcEncoder_Encode.c:357-358); next block's window control
                            if (nyb & 1u) WK_SKIP(1);
                            if (nyb > endNyb) { dead = true; fin = true; break; }
                            b++;
                            off[b] = nyb >> 1;
                            if (b == nb) { fin = true; break; }
                            if (nyb >= endNyb) { dead = true; fin = true; break; }
                            const uint32_t wx = __funnelshift_r(w0, w1, sh);
                            uint32_t wc;
                            if (wx & 8u) { wc = wx & 0xffu; WK_SKIP(2); }            // ulcDecoder.c:211-216
                            else         { wc = (wx & 15u) | 0x10u; WK_SKIP(1); }
                            layout = kLayoutBp[wc >> 4];
                            c = 0;
                        }
                        pat = layout; pos = 0;
                    }
                    subEnd = pos + (WN >> (pat & 7));
                    const uint32_t hx = __funnelshift_r(w0, w1, sh);               // sub-block header: a bare quantiser (ulcDecoder.c:103-112)
                    if ((hx & 15u) == 14u) {
                        WK_SKIP(2);
                        if (((hx >> 4) & 15u) == 15u) { pos = subEnd; continue; }   // Eh,Fh: silent sub-block
                    } else WK_SKIP(1);
                    break;
                }
                break;
            }
            if (nyb >= endNyb) { dead = true; fin = true; break; }          // ran off the stream
            const uint32_t x = __funnelshift_r(w0, w1, sh);
            const uint32_t t = wtab[(WALK_MAP_BYTES + WALK_ADV_BYTES) / 4 + 4u * st + (x & 15u)];
            WK_SKIP(1);
            if (t & (WALK_FLAG | WALK_TAIL)) {              // Fh,Eh,Fh: zeros to the end / Fh,Fh,Zh,Yh,Xh: decaying noise to the end
                if (t & WALK_TAIL) WK_SKIP(3);
                pos = subEnd; st = 0u;
            } else { pos += (int)(t & 0x7ffu); st = t >> 16; }
        }
        if (fin) break;
        window(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), xs, m);
    }
#undef WK_SKIP
    // the chunk fetched ahead may still be in flight: it must land before this CTA's shared memory is given away
    mbar_wait(&mb[(kc + 1u) & 1u], (par >> ((kc + 1u) & 1u)) & 1u);
    if (dead) {
        // the rest of the stream cannot be located (README.md:33: no resync): block b, where the walk died, and all after it are empty
        if (b < 0) { b = 0; off[0] = 0; }
        const uint32_t o = off[b];
        for (int k = b + 1; k <= nb; k++) off[k] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// parse: one lane per block
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bp_parse_kernel(UlcbBpArgs a) {
    __shared__ uint16_t lut[24];
    init_lut(lut);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = (int)(gid / a.maxBlk), b = (int)(gid - (long long)s * a.maxBlk);
    if (s >= a.nStreams) return;
    const int gs = a.s0 + s;
    const int nb = a.nBlocks ? (int)a.nBlocks[gs] : a.nBlocksAll;
    if (b >= nb) return;
    const uint32_t *off = a.blkOff + (size_t)s * (a.maxBlk + 1);
    const uint32_t o0 = off[b], o1 = off[b + 1];
    const uint64_t sbeg = a.streamOffs[gs];
    UlcbBpDesc d;
    d.bits = 0; d.wcRuns = 0x10; d.draws = 0; d.nCoef[0] = d.nCoef[1] = 0;
    if (o1 > o0) {
        const uint32_t len = o1 - o0;
        const uint64_t gidx = (uint64_t)s * (a.maxBlk + 1) + b;
        uint8_t *reg = a.arena + region_start(sbeg - a.base0 + o0, gidx);
        ParseOut po;
        const bool ok = parse_block<true>(a.bytes + sbeg + o0, a.wend, 2 * len, lut, reinterpret_cast<uint32_t *>(reg),
                                          reinterpret_cast<uint4 *>(reg + region_bytes(len)), po);
        // the index must agree with the grammar: the block ends inside its last byte (ulcEncoder_Encode.c:357-358)
        if (ok && ((po.nyb + 1) >> 1) == len) {
            d.bits = po.nyb * 4u;                                   // what ULC_DecodeBlock returns (ulcDecoder.c:301)
            d.wcRuns = po.wc | (po.nRun0 << 8) | ((po.nRunTot - po.nRun0) << 20);
            d.draws = po.draws;
            d.nCoef[0] = (uint16_t)po.nCoef0; d.nCoef[1] = (uint16_t)(po.nCoefTot - po.nCoef0);
        }
    }
    a.desc[(size_t)s * a.maxBlk + b] = d;
}

// ------------------------------------------------------------------------------------------------
// scan: one warp per stream - noise-draw offsets, first undecodable block, status and blockBits
// ------------------------------------------------------------------------------------------------
__global__ void bp_scan_kernel(UlcbBpArgs a) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= a.nStreams) return;
    const int gs = a.s0 + s;
    const int nb = a.nBlocks ? (int)a.nBlocks[gs] : a.nBlocksAll;
    UlcbBpDesc *desc = a.desc + (size_t)s * a.maxBlk;
    uint64_t run = 0;
    int firstBad = -1;
    for (int b0 = 0; b0 < nb && firstBad < 0; b0 += 32) {
        const int b = b0 + lane;
        uint32_t dr = 0, bits = 1;
        if (b < nb) { dr = desc[b].draws; bits = desc[b].bits; }
        uint64_t inc = dr;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t t = __shfl_up_sync(FULLM, inc, d); if (lane >= d) inc += t; }
        const uint64_t ex = run + inc - dr;
        // undecodable: grammar / index error, or noise draws beyond the sign table the context was created for
        const bool bad = bits == 0 || ex + dr + 128ull > a.nSignBits;
        const unsigned bal = __ballot_sync(FULLM, bad);
        if (bal) firstBad = b0 + __ffs(bal) - 1;
        if (b < nb && (firstBad < 0 || b < firstBad)) desc[b].draws = (uint32_t)ex;
        run += __shfl_sync(FULLM, inc, 31);
    }
    if (a.blockBits) {
        uint32_t *bo = a.blockBits + (size_t)gs * a.bitsStride;
        for (int b = lane; b < nb; b += 32) bo[b] = (firstBad < 0 || b < firstBad) ? desc[b].bits : 0u;
    }
    if (lane == 0) {
        a.firstBad[s] = firstBad < 0 ? nb : firstBad;
        if (a.status) a.status[gs] = firstBad;                      // -1 = every block decoded
    }
}

// ================================================================================================
// transform
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

// exp(-i pi j / 64) = (rotC(j), -rotS(j)), j < 32: the part of the DCT-IV pre / post twiddle E[l + 32 j] = E[l] * exp(-i pi j / 64) that does
// not depend on the lane - compile-time constants in the unrolled passes
__device__ __forceinline__ constexpr float rotC(int j) {
    constexpr float t[32] = {1.00000000000000000e+00f, 9.98795456205172405e-01f, 9.95184726672196929e-01f, 9.89176509964781014e-01f, 9.80785280403230431e-01f, 9.70031253194543974e-01f, 9.56940335732208824e-01f, 9.41544065183020806e-01f, 9.23879532511286738e-01f, 9.03989293123443338e-01f, 8.81921264348355050e-01f, 8.57728610000272118e-01f, 8.31469612302545236e-01f, 8.03207531480644943e-01f, 7.73010453362736993e-01f, 7.40951125354959106e-01f, 7.07106781186547573e-01f, 6.71558954847018330e-01f, 6.34393284163645488e-01f, 5.95699304492433468e-01f, 5.55570233019602289e-01f, 5.14102744193221661e-01f, 4.71396736825997809e-01f, 4.27555093430282196e-01f, 3.82683432365089837e-01f, 3.36889853392220051e-01f, 2.90284677254462331e-01f, 2.42980179903263982e-01f, 1.95090322016128331e-01f, 1.46730474455361748e-01f, 9.80171403295607702e-02f, 4.90676743274181260e-02f};
    return t[j];
}
__device__ __forceinline__ constexpr float rotS(int j) {
    constexpr float t[32] = {0.00000000000000000e+00f, 4.90676743274180149e-02f, 9.80171403295606036e-02f, 1.46730474455361748e-01f, 1.95090322016128248e-01f, 2.42980179903263871e-01f, 2.90284677254462331e-01f, 3.36889853392220051e-01f, 3.82683432365089782e-01f, 4.27555093430282085e-01f, 4.71396736825997642e-01f, 5.14102744193221661e-01f, 5.55570233019602178e-01f, 5.95699304492433357e-01f, 6.34393284163645488e-01f, 6.71558954847018330e-01f, 7.07106781186547462e-01f, 7.40951125354959106e-01f, 7.73010453362736993e-01f, 8.03207531480644832e-01f, 8.31469612302545236e-01f, 8.57728610000272118e-01f, 8.81921264348354939e-01f, 9.03989293123443338e-01f, 9.23879532511286738e-01f, 9.41544065183020806e-01f, 9.56940335732208935e-01f, 9.70031253194543974e-01f, 9.80785280403230431e-01f, 9.89176509964781014e-01f, 9.95184726672196818e-01f, 9.98795456205172405e-01f};
    return t[j];
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
    // cvt.rni.sat = round half to even, then saturate to [-32768, 32767]: the same value as lrintf of the clamped product
    short a, b;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(a) : "f"(l * 32768.0f));
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(b) : "f"(r * 32768.0f));
    return (uint32_t)(uint16_t)a | ((uint32_t)(uint16_t)b << 16);
}
template <>
__device__ __forceinline__ void store_pairs<true>(void *outBlk, int slotIdx, float l0, float r0, float l1, float r1) {
    __stcs(reinterpret_cast<uint2 *>(outBlk) + slotIdx, make_uint2(pcm16_pair(l0, r0), pcm16_pair(l1, r1)));
}

struct Tail { int start, end; float amp, r; };      // decaying noise to the end of a sub-block (ulcDecoder.c:172-186): level amp, decay factor r

struct WarpSm {
    float2   zb[32 * ZROW];       // FFT tile / coefficient words / short-block scratch / output swap
    uint32_t rec[RECSLOT];        // TMA landing slot: this channel's coefficient words of the next block; in between the noise work
                                  // list and, for a window-switched block, the pending tail in time order
    Tail     tail[8];
    int      nTails, pad[3];
};
struct CtaSm {
    WarpSm w[2];
    unsigned long long mbar[2];
    uint32_t tmemBase, pad[3];
};
constexpr int COEF16_OFS = 4224;  // byte offset of the bf16 coefficient array of a window-switched block inside zb (the
                                  // sub-block scratch occupies the first 4 KB)

// what one warp needs to know about one block
struct BlkInfo {
    uint32_t bits, wc, drawBase;
    int nCoef, nRuns, runIdx0;      // this channel's coefficient words / run records; index of its first run record
    const uint32_t *words;          // this channel's first coefficient word (global)
    const uint4 *runsEnd;           // end of the block's region: run record i sits at runsEnd[-1 - i]
};

__device__ __forceinline__ BlkInfo load_info(const UlcbBpArgs &a, int s, int b, int c, const uint32_t *off, uint64_t sbegRel) {
    const UlcbBpDesc d = a.desc[(size_t)s * a.maxBlk + b];
    BlkInfo bi;
    bi.bits = d.bits; bi.wc = d.wcRuns & 0xffu; bi.drawBase = d.draws;
    const int nRun0 = (d.wcRuns >> 8) & 0xfff, nRun1 = d.wcRuns >> 20;
    bi.nCoef = d.nCoef[c]; bi.nRuns = c ? nRun1 : nRun0; bi.runIdx0 = c ? nRun0 : 0;
    const uint32_t o0 = off[b], len = off[b + 1] - o0;
    const uint8_t *reg = a.arena + region_start(sbegRel + o0, (uint64_t)s * (a.maxBlk + 1) + b);
    bi.words = reinterpret_cast<const uint32_t *>(reg) + (c ? d.nCoef[0] : 0);
    bi.runsEnd = reinterpret_cast<const uint4 *>(reg + region_bytes(len));
    return bi;
}

// Issue the bulk copy of up to RECSLOT words of this channel starting at word `done` (lane 0 only).  The source is
// aligned down to 16 bytes; returns through `skew` how many leading words of the slot belong to something else.
__device__ __forceinline__ int issue_words(WarpSm &w, unsigned long long *mb, const BlkInfo &bi, int done, int lane, int &skew) {
    const uintptr_t src = reinterpret_cast<uintptr_t>(bi.words + done);
    skew = (int)((src >> 2) & 3u);
    int cnt = bi.nCoef - done;
    if (cnt > RECSLOT - skew) cnt = RECSLOT - skew;
    if (cnt > 0 && lane == 0) {
        const unsigned bytes = (unsigned)(((skew + cnt) * 4 + 15) & ~15);
        mbar_expect_tx(mb, bytes);
        tma_load(w.rec, reinterpret_cast<const void *>(src - 4 * (uintptr_t)skew), bytes, mb);
    }
    return cnt;
}

// Noise fill from the run records (ulcDecoder.c:156-160,179-184): the sign of draw g+i of a run that starts at draw g is
// parity[g+i+1] ^ parity[g].  The runs of a channel (a few dozen, 16..527 coefficients each, ~25 typical) are cut into items
// of 16 coefficient words; the warp works through the flat item list eight items at a time, four lanes per item, four
// words per lane, so that long and short runs cost the same per coefficient.  The item list lives in the (idle) TMA slot.
// `dense` = the block's bf16 coefficient array as 32-bit words (two coefficients per word): at the start of the tile for a long block
// (read back column-wise by the pre-twiddle, conflict-free), behind the sub-block scratch for a window-switched one.
struct RunRec { int s, e; uint32_t g0, amp2, flip, pad; };
constexpr int MAXITEMS = 2 * (RECSLOT - 32 * 6) ;   // u16 items behind 32 run records

__device__ __noinline__ void noise_fill(WarpSm &w, uint32_t *dense, const BlkInfo &bi, const uint32_t *__restrict__ par, int lane) {
    const int grp = lane >> 2, j = lane & 3;
    RunRec *rt = reinterpret_cast<RunRec *>(w.rec);
    uint16_t *items = reinterpret_cast<uint16_t *>(w.rec + 32 * 6);
    auto word = [&](int n) -> uint32_t * { return dense + n; };
    for (int r0 = 0; r0 < bi.nRuns; r0 += 32) {
        // ---- one run per lane: record, item count, item list ----
        const int r = r0 + lane;
        uint4 rr = make_uint4(0, 0, 0, 0);
        if (r < bi.nRuns) rr = __ldg(bi.runsEnd - 1 - (bi.runIdx0 + r));
        const int s = (int)rr.x, e = s + (int)rr.y;
        const uint32_t g0 = bi.drawBase + rr.z;
        uint32_t amp = rr.w & 0xffffu;
        if (rr.w & 0x80000000u) {
            // decaying tail: unit amplitude here, level and decay applied when the coefficients are loaded
            amp = 0x3f80u;
            if (rr.y) {
                const int lv = (int)(rr.w & 15u) + 1, nn = (int)((((rr.w >> 4) & 15u) << 4) | ((rr.w >> 8) & 15u));
                const int q = (int)((rr.w >> 12) & 31u);
                Tail t; t.start = s; t.end = e;
                t.amp = (float)(lv * lv) * quant_scale(q) * (1.0f / 16);
                t.r = 1.0f + (float)(nn * nn) * -0x1.0p-19f;
                const int slot = atomicAdd(&w.nTails, 1);
                if (slot < 8) w.tail[slot] = t;
            }
        }
        RunRec me;
        me.s = s; me.e = e; me.g0 = g0; me.amp2 = amp | (amp << 16);
        me.flip = 0u - ((__ldg(&par[g0 >> 5]) >> (g0 & 31)) & 1u); me.pad = 0;
        rt[lane] = me;
        const int nIt = rr.y ? ((((e - 1) >> 1) - (s >> 1)) >> 4) + 1 : 0;      // items of 16 words
        int inc = nIt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULLM, inc, d); if (lane >= d) inc += t; }
        const int total = __shfl_sync(FULLM, inc, 31);
        // rounds of at most MAXITEMS items (one round except for pathological blocks)
        for (int base = 0; base < total; base += MAXITEMS) {
            const int first = inc - nIt;                                         // my first item (global numbering)
            for (int k = max(first, base); k < inc && k < base + MAXITEMS; k++) items[k - base] = (uint16_t)(lane | ((k - first) << 5));
            __syncwarp();
            const int cnt = min(total - base, MAXITEMS);
            for (int it = grp; it < cnt; it += 8) {
                const uint32_t iv = items[it];
                const RunRec rc = rt[iv & 31u];
                const int wq = (rc.s >> 1) + 16 * (int)(iv >> 5) + 4 * j;
                const int p0 = 2 * wq;
                if (p0 < rc.e) {
                    const uint32_t m0 = rc.g0 + (uint32_t)(p0 - rc.s) + 1u;     // parity bit of position p0 (p0 >= s - 1)
                    const uint32_t a0 = __ldg(&par[m0 >> 5]), a1 = __ldg(&par[(m0 >> 5) + 1]);
                    const uint32_t bits = __funnelshift_r(a0, a1, m0 & 31) ^ rc.flip;
                    if (p0 >= rc.s && p0 + 8 <= rc.e) {                          // interior: four whole words
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            *word(wq + k) = rc.amp2 | (((bits >> (2 * k)) & 1u) << 15) | (((bits >> (2 * k + 1)) & 1u) << 31);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int pl = p0 + 2 * k, ph = pl + 1;
                            const bool vl = pl >= rc.s && pl < rc.e, vh = ph < rc.e;   // ph > pl >= s - 1
                            const uint32_t val = rc.amp2 | (((bits >> (2 * k)) & 1u) << 15) | (((bits >> (2 * k + 1)) & 1u) << 31);
                            uint32_t *dst = word(wq + k);
                            if (vl && vh) *dst = val;
                            else if (vl) *reinterpret_cast<uint16_t *>(dst) = (uint16_t)val;
                            else if (vh) *(reinterpret_cast<uint16_t *>(dst) + 1) = (uint16_t)(val >> 16);
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
}

// A decaying tail's level at offset k is the reference's own running product p *= r in f32 (ulcDecoder.c:178-184), not amp * r^k:
// with r = 1 - n^2 2^-19 close to 1 the rounding of p * r keeps its sign for thousands of steps and the sequence drifts up to 1e-4
// away from the closed form.  A lane walks its coefficients in ascending order and advances the chain between them (all lanes in
// lockstep: the cost of one chain per tail, and tails are rare).
struct TailChain {
    int start, end, k;
    float p, r;
    __device__ __forceinline__ void open(const Tail &t) { start = t.start; end = t.end; k = t.start; p = t.amp; r = t.r; }
    __device__ __forceinline__ void none() { start = end = k = 0; p = 1.0f; r = 1.0f; }
    __device__ __forceinline__ float gain(int idx) {         // idx ascending from call to call
        if (idx < start || idx >= end) return 1.0f;
        for (; k < idx; k++) p = __fmul_rn(p, r);
        return p;
    }
};

// The rare way into a long block's first FFT pass: a decaying noise tail (one per channel in a long block; its gains applied here) or
// the coefficient tap.  The words are expanded to float pairs in the tile in ascending order (the tail's running product), then
// pre-twiddled through the tile.  Out of line: it must not cost the common path registers or instruction-cache footprint.
template <bool DBG>
__device__ __noinline__ void expand_rare(WarpSm &w, int nT, float *dbg, int lane) {
    float2 *zb = w.zb;
    {
        // the words sit densely at the start of the tile they are expanded into: all of them into registers first
        uint32_t wds[32];
#pragma unroll
        for (int j = 0; j < 32; j++) wds[j] = reinterpret_cast<const uint32_t *>(w.zb)[lane + 32 * j];
        __syncwarp();
        TailChain tc;
        if (nT) tc.open(w.tail[0]); else tc.none();
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const int n = lane + 32 * j;
            const uint32_t wa = wds[j];
            float ev = __uint_as_float(wa << 16), od = __uint_as_float(wa & 0xffff0000u);
            if (nT) { ev *= tc.gain(2 * n); od *= tc.gain(2 * n + 1); }
            if (DBG) reinterpret_cast<float2 *>(dbg)[n] = make_float2(ev, od);
            zb[j * ZROW + lane] = make_float2(ev, od);
        }
    }
    __syncwarp();
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        const int ja = j, jb = 31 - j;
        const float2 va = zb[ja * ZROW + lane], vb = zb[jb * ZROW + lane];
        const float ea = va.x, oa = va.y, eb = vb.x, ob = vb.y;
        const float ia = __shfl_xor_sync(FULLM, ob, 31);      // v[N-1-2na]
        const float ib = __shfl_xor_sync(FULLM, oa, 31);      // v[N-1-2nb]
        const float ca = cosf(0.049087385212340517f * (float)ja), sa = sinf(0.049087385212340517f * (float)ja);   // pi / 64
        const float cb = cosf(0.049087385212340517f * (float)jb), sb = sinf(0.049087385212340517f * (float)jb);
        zb[ja * ZROW + lane] = make_float2(ea * ca + ia * sa, ia * ca - ea * sa);
        zb[jb * ZROW + lane] = make_float2(eb * cb + ib * sb, ib * cb - eb * sb);
    }
}

// ---- the pending tail lives in Tensor Memory ----
// Between blocks the un-fold needs, per lane, exactly the 32 values it left there one block earlier (slot k = lane + 32 k2 of
// the "k-layout": time sample 2k - N/2 for k >= N/4, N/2 - 1 - 2k below).  That is lane-private state, which is what TMEM's
// 32x32b access shape addresses: lane t of warp w owns TMEM lane 32 (w % 4) + t, column k2.  32 columns per CTA hold both
// channels' tails (warp 0 = lanes 0-31, warp 1 = lanes 32-63) and cost no shared memory and no registers between blocks.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// k-layout (TMEM) -> time order (shared memory) and back, around a window-switched block
__device__ __noinline__ void pend_to_time(uint32_t taddr, float *pend, int lane) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        float v[8];
        tmem_ld8(taddr + 8 * g, v);
#pragma unroll
        for (int t = 0; t < 8; t++) pend[slot_time(lane + 32 * (8 * g + t))] = v[t];
    }
    __syncwarp();
}
__device__ __noinline__ void pend_to_klayout(uint32_t taddr, const float *pend, int lane) {
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; g++) {
        float v[8];
#pragma unroll
        for (int t = 0; t < 8; t++) v[t] = pend[slot_time(lane + 32 * (8 * g + t))];
        tmem_st8(taddr + 8 * g, v);
    }
    tmem_wait_st();
    __syncwarp();
}

#ifndef ULCB_BP_MINB
#define ULCB_BP_MINB 8
#endif

template <bool PCM16, bool DBG>
__global__ void __launch_bounds__(64, ULCB_BP_MINB)
bp_transform_kernel(UlcbBpArgs a) {
    __shared__ __align__(128) CtaSm sm;
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int runsPerStream = (a.maxBlk + a.runLen - 1) / a.runLen;
    const int s = blockIdx.x / runsPerStream, run = blockIdx.x - s * runsPerStream;
    const int gs = a.s0 + s;
    const int nb = a.nBlocks ? (int)a.nBlocks[gs] : a.nBlocksAll;
    const int bFirst = run * a.runLen;
    if (bFirst >= nb) return;
    const int bLast = min(nb, bFirst + a.runLen);
    const int firstBad = a.firstBad[s];
    const uint32_t *off = a.blkOff + (size_t)s * (a.maxBlk + 1);
    const uint64_t sbegRel = a.streamOffs[gs] - a.base0;
    const UlcbTables &tb = a.tb;
    const float2 *__restrict__ T32 = tb.tw32;
    const float2 *__restrict__ win = tb.win;
    const uint32_t *__restrict__ par = tb.signParity;

    WarpSm &w = sm.w[c];
    float2 *zb = w.zb;
    float *pend = reinterpret_cast<float *>(w.rec);     // time-ordered copy, window-switched blocks only
    float *scr = reinterpret_cast<float *>(w.zb);
    uint16_t *coef16 = reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(w.zb) + COEF16_OFS);
    unsigned long long *mb = &sm.mbar[c];
    if (lane == 0) { mbar_init(mb, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (c == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&sm.tmemBase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = sm.tmemBase;
    const uint32_t taddr = tbase + ((uint32_t)(c * 32) << 16);
    {
        const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int g = 0; g < 4; g++) tmem_st8(taddr + 8 * g, z8);
        tmem_wait_st();
    }
    int lastSub = 0;
    unsigned phase = 0;
    __syncwarp();

    // the block before the run is transformed only for the pending tail it leaves (nothing older survives in it)
    int blk = bFirst > 0 ? bFirst - 1 : 0;
    if (blk >= firstBad) blk = bFirst;                 // ... unless the stream is already dead there
    BlkInfo cur{};
    int skew = 0, inflight = 0;
    if (blk < firstBad) {
        cur = load_info(a, s, blk, c, off, sbegRel);
        inflight = issue_words(w, mb, cur, 0, lane, skew);
    }

#pragma unroll 1
    for (; blk < bLast; blk++) {
        const bool warm = blk < bFirst;
        char *outBlk = reinterpret_cast<char *>(a.out) + ((size_t)gs * a.outStride + (size_t)blk * N * 2) * (PCM16 ? 2 : 4);
        if (blk >= firstBad) {
            // corrupt / exhausted stream: silence (the reference tool aborts here, ulcDecodeTool.c:153-157)
#pragma unroll 4
            for (int j = 0; j < 16; j++) store_pairs<PCM16>(outBlk, c * 512 + lane + 32 * j, 0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const int wc = (int)cur.wc;
        // one sub-block of the whole block size: layout 1 (N/1*, what encoders write) and layout 0 (N/1 without the transient flag:
        // "unused", ulcHelper.h:29, but a decoder must take what parses)
        const bool longBlk = (wc >> 4) <= 1;
        float *dbg = DBG ? a.dbgCoefs + (((size_t)gs * a.nBlocksAll + blk) * 2 + c) * N : nullptr;

        // ---- clear the coefficient words (bf16, 4 KB: at the start of the tile for a long block, behind the sub-block scratch for a
        //      window-switched one), then scatter this channel's words into them ----
        uint16_t *cdst = longBlk ? reinterpret_cast<uint16_t *>(w.zb) : coef16;
        {
            uint4 *z = reinterpret_cast<uint4 *>(cdst);
#pragma unroll 4
            for (int i = 0; i < (N * 2) / 16 / 32; i++) z[lane + 32 * i] = make_uint4(0, 0, 0, 0);
            if (lane == 0) w.nTails = 0;
        }
        __syncwarp();
        for (int done = 0; done < cur.nCoef;) {
            if (!inflight) inflight = issue_words(w, mb, cur, done, lane, skew);
            mbar_wait(mb, phase); phase ^= 1u;
#pragma unroll 4
            for (int i = lane; i < inflight; i += 32) {
                const uint32_t rw = w.rec[skew + i];
                cdst[rw >> 16] = (uint16_t)rw;
            }
            __syncwarp();
            done += inflight; inflight = 0;
        }
        // noise runs (the landing slot holds their work list), then - long block - start the next block's copy behind this
        // block's transform; a window-switched block keeps the slot for its time-ordered pending tail until it is done
        noise_fill(w, reinterpret_cast<uint32_t *>(cdst), cur, par, lane);
        const int nT = min(w.nTails, 8);
        const bool haveNext = blk + 1 < bLast && blk + 1 < firstBad;
        BlkInfo nxt{};
        if (haveNext) nxt = load_info(a, s, blk + 1, c, off, sbegRel);
        inflight = 0;
        if (haveNext && longBlk) inflight = issue_words(w, mb, nxt, 0, lane, skew);

        if (longBlk) {
            // ======================= long block: register path =======================
            float xr[32], xi[32];
            if (nT == 0 && !DBG) {
                // z[n] = (v[2n] + i v[N-1-2n]) E[n], n = lane + 32 j: v[2n] is the even half of word n (own column of the tile),
                // v[N-1-2n] the odd half of word 1023-n = column 31-lane, row 31-j.  Straight into the registers of pass 0.
                // The lane-dependent factor E[lane] of the twiddle moves behind pass 0 (it is common to the whole column).
                const uint32_t *own = reinterpret_cast<const uint32_t *>(w.zb) + lane, *mir = reinterpret_cast<const uint32_t *>(w.zb) + 1023 - lane;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const uint32_t wa = own[32 * j], wm = mir[-32 * j];
                    const float ev = __uint_as_float(wa << 16), od = __uint_as_float(wm & 0xffff0000u);
                    xr[j] = ev * rotC(j) + od * rotS(j); xi[j] = od * rotC(j) - ev * rotS(j);
                }
                __syncwarp();
            } else {
                expand_rare<DBG>(w, nT, dbg, lane);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j++) { const float2 v = zb[j * ZROW + lane]; xr[j] = v.x; xi[j] = v.y; }
                __syncwarp();
            }
            // 1024-point FFT as 32 x 32, ONE copy of the 32-point code run for both passes:
            //   pass 0: lane l transforms column l (n = l + 32 j over j), multiplies by G[l][k1] = E[l] W_1024^(l k1) E[k1] - the
            //           inter-pass twiddle with both lane-dependent halves of the DCT-IV twiddles folded in; G[l][k1] = A_l Q_l^k1 is
            //           built from six per-lane values by a product tree - and stores A[l][k1] at zb[k1][l]
            //   pass 1: lane k1 transforms row k1 (over l); y[k] = Z[k] exp(-i pi k2 / 64), k = k1 + 32 k2, stays in registers
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                if (pass) {
#pragma unroll
                    for (int r = 0; r < 32; r++) { const float2 v = zb[lane * ZROW + r]; xr[r] = v.x; xi[r] = v.y; }
                    __syncwarp();
                }
                fft32(xr, xi);                            // register r now holds output index brev5(r)
                if (!pass) {
                    const float2 gA = __ldg(T32 + lane);
                    float2 P[5];
#pragma unroll
                    for (int t = 0; t < 5; t++) P[t] = __ldg(T32 + 32 * (t + 1) + lane);
#pragma unroll
                    for (int b4 = 0; b4 < 2; b4++) {
                        const float2 g4 = b4 ? cmul(gA, P[4]) : gA;
#pragma unroll
                        for (int b3 = 0; b3 < 2; b3++) {
                            const float2 g3 = b3 ? cmul(g4, P[3]) : g4;
#pragma unroll
                            for (int b2 = 0; b2 < 2; b2++) {
                                const float2 g2 = b2 ? cmul(g3, P[2]) : g3;
#pragma unroll
                                for (int b1 = 0; b1 < 2; b1++) {
                                    const float2 g1 = b1 ? cmul(g2, P[1]) : g2;
#pragma unroll
                                    for (int b0 = 0; b0 < 2; b0++) {
                                        const float2 g = b0 ? cmul(g1, P[0]) : g1;
                                        const int k1 = b4 * 16 + b3 * 8 + b2 * 4 + b1 * 2 + b0, r = brev5(k1);
                                        zb[k1 * ZROW + lane] = make_float2(xr[r] * g.x - xi[r] * g.y, xr[r] * g.y + xi[r] * g.x);
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int r = 0; r < 32; r++) {
                        const int k2 = brev5(r);
                        const float xa = xr[r], xb = xi[r];
                        xr[r] = xa * rotC(k2) + xb * rotS(k2); xi[r] = xb * rotC(k2) - xa * rotS(k2);
                    }
                }
            }
            // un-fold (Fourier_IMDCT_Template.h:66-93) against the pending tail, all at the same k = lane + 32 k2; the tail comes
            // from / goes back to TMEM eight columns at a time
            int ov = (wc >> 4) ? N >> (wc & 7) : N;       // layout 1: the long block is its own transient sub-block; layout 0: no overlap scaling
            if (ov > lastSub) ov = lastSub;
            lastSub = N;
            if (ov == N) {
                // steady state (full overlap): sine window over the whole half block, win[N/2 + i]
                const float2 *__restrict__ wlo = win + H + (H - 1 - 2 * lane), *__restrict__ whi = win + H + (2 * lane - H);
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float pa[8], np[8];
                    tmem_ld8(taddr + 8 * g, pa);
#pragma unroll
                    for (int t = 0; t < 8; t++) {
                        const int k2 = 8 * g + t, r = brev5(k2);
                        if (k2 < 16) {                    // i = N/2-1-2k, X[N/2+i] = -Im y, next tail = Re y
                            const float2 wv = __ldg(wlo - 64 * k2);
                            const float b = -xi[r];
                            np[t] = xr[r];
                            zb[k2 * ZROW + lane] = make_float2(wv.x * pa[t] - wv.y * b, wv.y * pa[t] + wv.x * b);
                        } else {                          // i = 2k-N/2, X[N/2+i] = Re y, next tail = -Im y
                            const float2 wv = __ldg(whi + 64 * k2);
                            const float b = xr[r];
                            np[t] = -xi[r];
                            zb[k2 * ZROW + lane] = make_float2(wv.x * pa[t] - wv.y * b, wv.y * pa[t] + wv.x * b);
                        }
                    }
                    tmem_st8(taddr + 8 * g, np);
                }
            } else {
                // first block of a stream / after a short block / transient long block: flat part + narrower window
                const int flat = (N - ov) >> 1;
                const int wofs = (ov >> 1) - flat;
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float pa[8], np[8];
                    tmem_ld8(taddr + 8 * g, pa);
#pragma unroll
                    for (int t = 0; t < 8; t++) {
                        const int k2 = 8 * g + t, r = brev5(k2), k = lane + 32 * k2;
                        const bool up = k2 >= 16;
                        const int i = up ? 2 * k - H : H - 1 - 2 * k;
                        const float b = up ? xr[r] : -xi[r];      // X[N/2 + i]
                        np[t] = up ? -xi[r] : xr[r];              // X[N/2 - 1 - i]
                        float lo, hi;
                        if (i < flat) { lo = pa[t]; hi = b; }
                        else { const float2 wv = __ldg(&win[wofs + i]); lo = wv.x * pa[t] - wv.y * b; hi = wv.y * pa[t] + wv.x * b; }
                        zb[k2 * ZROW + lane] = make_float2(lo, hi);   // time samples i and N-1-i
                    }
                    tmem_st8(taddr + 8 * g, np);
                }
            }
            tmem_wait_st();
            if (!warm) {
                __syncthreads();
                // warp 0 stores the first half of the block (samples i = .x), warp 1 the second (samples N-1-i = .y); M/S undo
                // (ulcDecoder.c:281-289).  Pair (t, t+1) = slots k_e = p + (c ? 0 : 512) and k_o = (c ? 1023 : 511) - p, p = lane + 32 q.
                const float *m0 = reinterpret_cast<const float *>(sm.w[0].zb) + c, *s0 = reinterpret_cast<const float *>(sm.w[1].zb) + c;
#pragma unroll 2
                for (int q = 0; q < 16; q++) {
                    const int ie = ((c ? q : q + 16) * ZROW + lane) * 2, io = ((c ? 31 - q : 15 - q) * ZROW + 31 - lane) * 2;
                    const float me = m0[ie], se = s0[ie], mo = m0[io], so = s0[io];
                    store_pairs<PCM16>(outBlk, c * 512 + lane + 32 * q, me + se, me - se, mo + so, mo - so);
                }
                __syncthreads();
            } else __syncwarp();
        } else {
            // ======================= window-switched block: shared-memory path, time order =======================
            pend_to_time(taddr, pend, lane);
            int offc = 0;
            for (unsigned pat = kLayoutBp[wc >> 4];;) {
                const int sh = pat & 7, S = N >> sh, hS = S >> 1;
                TailChain tc;
                tc.none();
                for (int t = 0; t < nT; t++) if (w.tail[t].start >= offc && w.tail[t].start < offc + S) tc.open(w.tail[t]);   // at most one per sub-block
                for (int i = lane; i < S; i += 32) {
                    float x = __uint_as_float((uint32_t)coef16[offc + i] << 16);
                    if (nT) x *= tc.gain(offc + i);
                    if (DBG) dbg[offc + i] = x;
                    scr[i] = x;
                }
                __syncwarp();
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
                        const float pa = pend[held + i], b = scr[hS + i];
                        if (i < flat) { uLo[r] = pa; uHi[r] = b; }
                        else { const float2 ww = __ldg(&wv[i - flat]); uLo[r] = ww.x * pa - ww.y * b; uHi[r] = ww.y * pa + ww.x * b; }
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
                if (!warm) {
                    __syncthreads();
                    const float *m = reinterpret_cast<const float *>(sm.w[0].zb), *sd = reinterpret_cast<const float *>(sm.w[1].zb);
                    for (int p = lane; p < (hS >> 1); p += 32) {
                        const int t = c * hS + 2 * p;
                        const float m0 = m[t], m1 = m[t + 1], s0 = sd[t], s1 = sd[t + 1];
                        store_pairs<PCM16>(outBlk, (offc + t) >> 1, m0 + s0, m0 - s0, m1 + s1, m1 - s1);
                    }
                    __syncthreads();
                }
                offc += S;
                pat >>= 4;
                if (!pat) break;
            }
            pend_to_klayout(taddr, pend, lane);
            if (haveNext) inflight = issue_words(w, mb, nxt, 0, lane, skew);
        }
        cur = nxt;
    }
    // every tcgen05 access of this CTA is complete (wait::st above): hand the columns back
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (c == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tbase) : "memory");
}

}  // namespace

// Fast path applicability: BlockSize 2048, stereo, fresh streams (no carried per-stream state).
bool ulcb_decode_bp_supported(const UlcbTables &tb, int nChan) { return tb.N == N && nChan == NCH && tb.tw32 != nullptr; }

size_t ulcb_decode_bp_arena_bytes(uint64_t payloadBytes, int nStreams, int maxBlk) {
    return (size_t)(16ull * ((payloadBytes + 1) >> 1) + 32ull * (uint64_t)nStreams * (uint64_t)(maxBlk + 1) + 256ull);
}

// Block byte offsets of every stream of the call (a.s0 = 0, a.nStreams = all, a.blkOff = the call's table): from the caller's
// block-size index, or by walking the token lengths of foreign streams.
cudaError_t ulcb_decode_bp_offsets(const UlcbBpArgs &a, bool haveIndex, int smCount, cudaStream_t stream, int *launches) {
    if (a.nStreams <= 0 || a.maxBlk <= 0) return cudaSuccess;
    if (haveIndex) {
        const int wpc = 4;
        bp_index_kernel<<<(a.nStreams + wpc - 1) / wpc, wpc * 32, 0, stream>>>(a);
    } else {
        // one lane per stream, latency-bound: few streams run with few active lanes per warp (divergence between the lanes of
        // a warp is what the walk pays for), up to 8 warps per SM; many streams fill the warps
        int lanes = 32;
        while (lanes > 1 && (long long)a.nStreams * 2 <= (long long)lanes * smCount * 8) lanes >>= 1;
        if (const char *e = getenv("ULCB_WALK_LANES")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) lanes = v; }
        const long long warpsTotal = ((long long)a.nStreams + lanes - 1) / lanes;
        int warps = (int)((warpsTotal + smCount - 1) / smCount);
        if (warps < 1) warps = 1;
        if (warps > 16) warps = 16;
        if (warps * lanes > WALK_MAXLANES) warps = WALK_MAXLANES / lanes;     // one ring per lane
        const size_t smem = (size_t)WALK_TAB_BYTES + (size_t)WALK_MAXLANES * WALK_RINGW * sizeof(uint32_t) + WALK_MAXLANES * 16;
        cudaError_t e = cudaFuncSetAttribute(bp_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
        if (e != cudaSuccess) return e;
        bp_walk_kernel<<<(unsigned)((warpsTotal + warps - 1) / warps), 32 * warps, smem, stream>>>(a, lanes);
    }
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// Block sizes of every stream of the call from the offsets table (the index a .ulc file lacks: ULCB_IndexStreams).
namespace {
__global__ void bp_sizes_kernel(UlcbBpArgs a, uint32_t *sizes, uint64_t sizesStride, int32_t *status) {
    const int s = blockIdx.x, gs = a.s0 + s;
    const int nb = a.nBlocks ? (int)a.nBlocks[gs] : a.nBlocksAll;
    const uint32_t *off = a.blkOff + (size_t)s * (a.maxBlk + 1);
    __shared__ int firstBad;
    if (threadIdx.x == 0) firstBad = nb;
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const uint32_t len = off[b + 1] - off[b];
        sizes[(size_t)gs * sizesStride + b] = len;
        if (len == 0) atomicMin(&firstBad, b);                     // a located block is never empty (window control + headers)
    }
    __syncthreads();
    if (threadIdx.x == 0 && status) status[gs] = firstBad < nb ? firstBad : -1;
}
}  // namespace

cudaError_t ulcb_decode_bp_sizes(const UlcbBpArgs &a, uint32_t *dSizes, uint64_t sizesStride, int32_t *dStatus, cudaStream_t stream) {
    if (a.nStreams <= 0 || a.maxBlk <= 0) return cudaSuccess;
    bp_sizes_kernel<<<a.nStreams, 128, 0, stream>>>(a, dSizes, sizesStride, dStatus);
    return cudaGetLastError();
}

// One wave: parse -> scan -> transform (a.blkOff points at the wave's first stream).
cudaError_t ulcb_decode_bp_launch(const UlcbBpArgs &a, int pcm16, cudaStream_t stream, int *launches, cudaEvent_t *ev, int phases) {
    if (a.nStreams <= 0 || a.maxBlk <= 0) return cudaSuccess;
    // phases: bit 0 = parse + scan (ev[0..2]), bit 1 = transform (ev[2..3]); the pipelined caller runs them on two streams
    if (phases & 1) {
        const int wpc = 4;                                        // warps per CTA of the per-stream kernels
        const long long nThr = (long long)a.nStreams * a.maxBlk;
        if (ev) cudaEventRecord(ev[0], stream);
        bp_parse_kernel<<<(unsigned)((nThr + 127) / 128), 128, 0, stream>>>(a);
        if (ev) cudaEventRecord(ev[1], stream);
        bp_scan_kernel<<<(a.nStreams + wpc - 1) / wpc, wpc * 32, 0, stream>>>(a);
        if (launches) *launches += 2;
    }
    if (ev && (phases == 1 || phases == 3)) cudaEventRecord(ev[2], stream);
    if (phases & 2) {
        if (ev && phases == 2) cudaEventRecord(ev[2], stream);
        const int runsPerStream = (a.maxBlk + a.runLen - 1) / a.runLen;
        const long long grid = (long long)a.nStreams * runsPerStream;
        if (a.dbgCoefs) {
            if (pcm16) bp_transform_kernel<true, true><<<(unsigned)grid, 64, 0, stream>>>(a);
            else       bp_transform_kernel<false, true><<<(unsigned)grid, 64, 0, stream>>>(a);
        } else {
            if (pcm16) bp_transform_kernel<true, false><<<(unsigned)grid, 64, 0, stream>>>(a);
            else       bp_transform_kernel<false, false><<<(unsigned)grid, 64, 0, stream>>>(a);
        }
        if (ev) cudaEventRecord(ev[3], stream);
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}
