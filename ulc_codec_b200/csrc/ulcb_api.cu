// ulcb_api.cu - C-ABI layer of libulcb200.so: contexts, host-built tables, the batched entry points
// and the libulc drop-in API (include/ulcb200.h).  Host code only orchestrates; all codec work runs in
// the CUDA kernels.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

#include "../../include/ulcb200.h"
#include "ulcb_dct.cuh"
#include "ulcb_decode.h"
#include "ulcb_encode.h"
#include "ulcb_pcm.h"

cudaError_t ulcb_corpus_launch(int16_t *dOut, uint64_t strideElems, int kind, uint32_t firstId, int nStreams, uint64_t nSamples,
                               uint64_t t0, const int16_t *dSine, const int16_t *dBurst, cudaStream_t stream);

// kinds whose spans own both of their events (the decode pipeline's parse / scan / transform spans share a chain of events)
static inline bool own_end_event(int kind) {
    return kind == ULCB_KERNEL_BP_WALK || kind == ULCB_KERNEL_BP_INDEX || kind == ULCB_KERNEL_ENC_DETECT || kind == ULCB_KERNEL_ENC_ENCODE;
}

// chunked decode pipeline defaults (ULCB_SetDecodePipeline): chunks 0 = automatic - chunks of 128 streams with overlap 2 for the
// host-pointer entry points (their transform waits for PCIe: upload, walk and parse of the next chunks hide behind it, 481 -> 435 ms
// per 2048 x 60 s streams at 16 chunks, 949 -> 862 ms per 4096 at 32; a walk costs ~15 ms however few streams it covers and the walks
// run one after the other, so at 64 streams per chunk they take as long as the call), none for device-resident input (there the
// side-stream kernels take shared memory and issue slots from the transform: 128 -> 165 ms at 4 chunks).  profiles/README.md, r02i / r02l.
#define ULCB_BP_CHUNKS_DEFAULT 0
#define ULCB_BP_OVERLAP_DEFAULT 0
#define ULCB_BP_AUTO_HOST_CHUNK_STREAMS 128
#define ULCB_BP_AUTO_HOST_OVERLAP 2

struct ULCB_Context {
    int device = 0, rateHz = 0, nChan = 0, N = 0;
    uint32_t maxBlocks = 0;
    cudaStream_t stream = nullptr;
    int lastCuda = 0;
    uint64_t launches = 0;
    UlcbTables tb{};
    UlcbEncTables etb{};
    int smCount = 148;
    int16_t *corpusLut = nullptr;       // sine[1024] + burst[882]
    float *dAvgCx = nullptr; int avgCap = 0, avgCount = 0;              // two-pass ABR: per-stream AvgComplexity
    void *detScratch = nullptr; size_t detScratchCap = 0;              // detector pre-pass: per-CTA envelope buffers ...
    int32_t *dWcAll = nullptr; size_t wcAllCap = 0;                     // ... and its decisions [stream][nBlocksAll]
    float *dCxScratch = nullptr; size_t cxScratchCap = 0;               // ... and the first pass's block complexities when the caller keeps none
    float *encWs = nullptr; size_t encWsFloats = 0; int encGrid = 0;   // global workspace for large block sizes
    std::vector<void *> owned;          // device allocations freed with the context
    // staging for the host-pointer entry points (grown on demand)
    void *dIn = nullptr;  size_t dInCap = 0;
    void *dOut = nullptr; size_t dOutCap = 0;
    void *dAux = nullptr; size_t dAuxCap = 0;
    void *dSign = nullptr;              // noise-sign parity table (tb.signParity), rebuilt when it has to grow
    // block-parallel decode pipeline (ulcb_decode_bp.cu): wave scratch, grown on demand
    void *bpOff = nullptr;   size_t bpOffCap = 0;
    void *bpDesc = nullptr;  size_t bpDescCap = 0;
    void *bpBad = nullptr;   size_t bpBadCap = 0;
    void *bpArena = nullptr; size_t bpArenaCap = 0;
    size_t bpPlanCap = 0;               // bytes of arena a wave may plan for
    void *hPlan = nullptr;   size_t hPlanCap = 0;   // pinned: stream offsets / block counts read back to plan the waves
    // chunked pipeline of the block-parallel decode (ULCB_SetDecodePipeline): upload + walk (+ parse) of chunk c+1 run on
    // high-priority side streams beside the transform of chunk c
    int bpChunks = 0, bpOverlap = 0;              // chunks 0 = automatic
    cudaStream_t bpFront = nullptr, bpCopy = nullptr;
    std::vector<cudaEvent_t> bpSyncEv;
    // per-kernel timing (ULCB_SetKernelTiming): event pairs recorded around the launches, resolved by ULCB_KernelTimes
    bool timing = false;
    struct TimedSpan { cudaEvent_t a, b; int kind; bool ownB; };
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> evPool;
    double kernMs[ULCB_KERNEL_KINDS] = {}; uint64_t kernCount[ULCB_KERNEL_KINDS] = {};
};

namespace {

cudaEvent_t timing_event(ULCB_Context *ctx) {
    cudaEvent_t e = nullptr;
    if (!ctx->evPool.empty()) { e = ctx->evPool.back(); ctx->evPool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}

#define CK(ctx, expr)                                                        \
    do {                                                                     \
        cudaError_t e_ = (expr);                                             \
        if (e_ != cudaSuccess) { (ctx)->lastCuda = (int)e_; return ULCB_E_CUDA; } \
    } while (0)

// ---- window rotation table: the reference's own f32 recurrence, so the decoder's window VALUES are
// those of the stock scalar build (libfourier/source/Fourier_Helper.h:218-240 minimax polynomials,
// Fourier_IMDCT_Template.h:76-91 rotation).  Plain mul/add in source order, no contraction.
float poly_sin(float x) {
    volatile float x2 = x * x;
    volatile float r = x2 * 0x1.3C8B08p-13f; r = r + -0x1.3237ECp-8f;
    r = x2 * r; r = r + 0x1.4667ACp-4f;
    r = x2 * r; r = r + -0x1.4ABBB8p-1f;
    r = x2 * r; r = r + 0x1.921FB4p0f;
    return x * r;
}
float poly_cos(float x) {
    volatile float x2 = x * x;
    volatile float r = x2 * 0x1.C2EA62p-11f; r = r + -0x1.550810p-6f;
    r = x2 * r; r = r + 0x1.03BDD6p-2f;
    r = x2 * r; r = r + -0x1.3BD3B2p0f;
    r = x2 * r; r = r + 1.0f;
    return r;
}
void build_window(std::vector<float2> &win, int N) {
    win.assign(N, float2{1.0f, 0.0f});
    for (int ov = 1; ov <= N; ov <<= 1) {
        int cnt = ov > 1 ? ov / 2 : 1, base = ov / 2;
        volatile float c = poly_cos(0.5f / ov), s = poly_sin(0.5f / ov);
        const float wc = poly_cos(1.0f / ov), ws = poly_sin(1.0f / ov);
        for (int i = 0; i < cnt; i++) {
            win[base + i] = float2{c, s};
            volatile float a = c, b = s;
            volatile float t0 = wc * a, t1 = ws * b, t2 = ws * a, t3 = wc * b;
            c = t0 - t1;
            s = t2 + t3;
        }
    }
}

// slot of Z[k] after the in-place radix-4 (final radix-2 if log2 M odd) DIF FFT of size M
int fft_slot(int k, int M) {
    int pos = 0;
    while (M >= 4) { pos += (k & 3) * (M >> 2); k >>= 2; M >>= 2; }
    if (M == 2) pos += k & 1;
    return pos;
}

template <typename T>
int upload(ULCB_Context *ctx, const std::vector<T> &h, const T **d) {
    void *p = nullptr;
    CK(ctx, cudaMalloc(&p, h.size() * sizeof(T) + 16));
    ctx->owned.push_back(p);
    CK(ctx, cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *d = reinterpret_cast<const T *>(p);
    return ULCB_OK;
}

// running parity of the xorshift32 MSBs (libulc/ulcDecoder.c:75-81): bit m = parity of draws 0..m-1.  (Re)built for at least
// nBits bits; the drop-in decoder grows it when a stream outlives the table (ULC_DecodeBlock has no length bound).
int build_sign_table(ULCB_Context *ctx, uint64_t nBits) {
    std::vector<uint32_t> par((nBits + 31) / 32 + 4, 0u);
    uint32_t seed = 1234567u, p = 0;
    for (uint64_t m = 1; m < nBits; m++) {
        seed ^= seed << 13; seed ^= seed >> 17; seed ^= seed << 5;
        p ^= seed >> 31;
        par[m >> 5] |= p << (m & 31);
    }
    void *d = nullptr;
    CK(ctx, cudaMalloc(&d, par.size() * sizeof(uint32_t)));
    if (cudaMemcpy(d, par.data(), par.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return ULCB_E_CUDA; }
    if (ctx->dSign) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->dSign); }
    ctx->dSign = d;
    ctx->tb.signParity = reinterpret_cast<const uint32_t *>(d);
    ctx->tb.nSignBits = nBits;
    return ULCB_OK;
}

int build_tables(ULCB_Context *ctx) {
    const int N = ctx->N;
    UlcbTables &tb = ctx->tb;
    tb.N = N;
    tb.log2N = 31 - __builtin_clz(N);
    int rc;
    {
        std::vector<float2> w(N / 2);
        for (int k = 0; k < N / 2; k++) {
            double a = -2.0 * M_PI * k / (N / 2);
            w[k] = float2{(float)cos(a), (float)sin(a)};
        }
        if ((rc = upload(ctx, w, &tb.fftW))) return rc;
    }
    tb.tw32 = nullptr;
    if (N == 2048) {
        // per-lane bases of the folded inter-pass twiddle G[l][k1] = A_l Q_l^k1 (ulcb_decode_bp.cu): [0] = A_l, [1 + t] = Q_l^(2^t)
        std::vector<float2> t(6 * 32);
        for (int l = 0; l < 32; l++) {
            const double aa = -M_PI * (8.0 * l + 2.0) / 16384.0;
            t[l] = float2{(float)cos(aa), (float)sin(aa)};
            for (int b = 0; b < 5; b++) {
                const double qa = -M_PI * (4.0 * l + 1.0) * (double)(1 << b) / 2048.0;
                t[32 * (b + 1) + l] = float2{(float)cos(qa), (float)sin(qa)};
            }
        }
        if ((rc = upload(ctx, t, &tb.tw32))) return rc;
    }
    for (int sh = 0; sh < 4; sh++) {
        int S = N >> sh, M = S / 2;
        std::vector<float2> e(M);
        std::vector<uint16_t> pos(M);
        for (int n = 0; n < M; n++) {
            double a = -M_PI * (8.0 * n + 1.0) / (8.0 * S);
            e[n] = float2{(float)cos(a), (float)sin(a)};
            pos[n] = (uint16_t)fft_slot(n, M);
        }
        if ((rc = upload(ctx, e, &tb.dctE[sh]))) return rc;
        if ((rc = upload(ctx, pos, &tb.fftPos[sh]))) return rc;
    }
    {
        std::vector<float2> win;
        build_window(win, N);
        if ((rc = upload(ctx, win, &tb.win))) return rc;
    }
    {
        // Token-length automaton of the foreign-stream walk (ulcb_decode_bp.cu: bp_walk_kernel), after FormatSpecs.md:53-131 /
        // libulc/ulcDecoder.c:113-186.  States: 0 token start, 1 after 0h, 2 after 1h, 3 after 1h,Yh, 4 after 8h, 5 after 8h,Zh,
        // 6 after 8h,Zh,Yh, 7 after Fh, 8 after Fh,Eh.
        auto step1 = [](int st, int v, int *adv, int *stop) -> int {
            *adv = 0; *stop = 0;
            switch (st) {
            case 0: if (v == 0) return 1; if (v == 1) return 2; if (v == 8) return 4; if (v == 15) return 7; *adv = 1; return 0;
            case 1: *adv = v + 1; return 0;                    // 0h,Xh: 1..16 zeros
            case 2: *adv = 16 * v + 33; return 3;              // 1h,Yh,Xh: 33..288 zeros
            case 3: *adv = v; return 0;
            case 4: *adv = 32 * v + 16; return 5;              // 8h,Zh,Yh,Xh: 16..527 noise
            case 5: *adv = 2 * v; return 6;
            case 6: *adv = v & 1; return 0;
            case 7: if (v < 14) return 0; if (v == 14) return 8; *stop = 2; return 0;   // Fh,0h..Dh / Fh,Eh,.. / Fh,Fh,..
            default: if (v == 15) *stop = 1; return 0;         // Fh,Eh,0h..Eh quantiser / Fh,Eh,Fh stop
            }
        };
        // u64 MAP[4096]: for three nybbles, the next state of every entry state k as 4 k' in bits 4k+2..4k+5;
        // u16 ADV[9][4096]: coefficients the three nybbles cover from entry state k | stop flag << 15;
        // u32 T1[9][16]: one nybble from state k: 4 k' << 16 | Fh,Fh tail << 12 | Fh,Eh,Fh stop << 11 | coefficients
        std::vector<uint32_t> tab(4096 * 2 + 9 * 4096 / 2 + 9 * 16);
        uint64_t *mapT = reinterpret_cast<uint64_t *>(tab.data());
        uint16_t *advT = reinterpret_cast<uint16_t *>(tab.data() + 4096 * 2);
        uint32_t *t1 = tab.data() + 4096 * 2 + 9 * 4096 / 2;
        for (int x = 0; x < 4096; x++) {
            uint64_t mp = 0;
            for (int st = 0; st < 9; st++) {
                int cs = st, total = 0, flag = 0;
                for (int k = 0; k < 3 && !flag; k++) {
                    int adv, stop;
                    cs = step1(cs, (x >> (4 * k)) & 15, &adv, &stop);
                    total += adv;
                    if (stop) flag = 1;
                }
                if (!flag) mp |= (uint64_t)cs << (4 * st + 2);
                advT[(size_t)st * 4096 + x] = (uint16_t)(flag ? 0x8000 : total);
            }
            mapT[x] = mp;
        }
        for (int st = 0; st < 9; st++)
            for (int v = 0; v < 16; v++) {
                int adv, stop;
                const int ns = step1(st, v, &adv, &stop);
                t1[st * 16 + v] = stop ? (stop == 2 ? 0x1000u : 0x800u) : ((uint32_t)(4 * ns) << 16) | (uint32_t)adv;
            }
        if ((rc = upload(ctx, tab, &tb.walkTab))) return rc;
    }
    return build_sign_table(ctx, (uint64_t)ctx->maxBlocks * ctx->nChan * N + 256);
}

// ---- encoder tables (host libm, so the constants are the reference's own) ----
// libulc/ulcHelper.h:96-120
float bark_to_line(float bark, float nyquist, int n) {
    float hz = 600.0f * sinhf(bark * (1.0f / 6.0f));
    return (hz * (float)n / nyquist) - 0.5f;
}
float line_to_bark(int line, float nyquist, int n) {
    float hz = ((float)line + 0.5f) * nyquist / (float)n;
    return 6.0f * asinhf(hz * (1.0f / 600.0f));
}
void band_edges(float b0, float b1, float nyq, int n, int *beg, int *end) {   // ulcEncoder_Psyopt.c:83-92,190-199
    int lo = (int)floorf(bark_to_line(b0, nyq, n)), hi = (int)ceilf(bark_to_line(b1, nyq, n));
    if (lo < 0) lo = 0;
    if (hi < 0) hi = 0;
    if (lo > n - 1) lo = n - 1;
    if (hi > n) hi = n;
    *beg = lo; *end = hi;
}

int build_enc_tables(ULCB_Context *ctx) {
    const int N = ctx->N, rate = ctx->rateHz;
    UlcbEncTables &t = ctx->etb;
    t.N = N; t.log2N = ctx->tb.log2N; t.rateHz = rate;
    t.rot = ctx->tb.win;
    // one-pole coefficients of the transient detector (ulcEncoder_WindowControl.c:73-129), as (1 - k)
    volatile float k;
    k = expf(-0x1.CC845Cp6f / rate);  t.c1PostHp = 1.0f - k;
    k = expf(-0x1.596344p8f / rate);  t.c1PostBp = 1.0f - k;
    k = expf(-0x1.CC845Cp7f / rate);  t.c1PreHp  = 1.0f - k;
    k = expf(-0x1.596344p8f / rate);  t.c1PreBp  = 1.0f - k;
    k = expf(-0x1.1AF110p-6f * N / rate); t.c1Blk = 1.0f - k;
    const float nyq = (float)rate * 0.5f;
    std::vector<int16_t> edges(4 * 2 * ULCB_NBARK * 2);
    std::vector<int16_t> events(4 * 2 * 2 * ULCB_NBARK * 2);
    std::vector<uint8_t> lband;
    std::vector<float> lfrac;
    for (int sh = 0; sh < 4; sh++) {
        const int n = (N >> sh) / 2;
        for (int kind = 0; kind < 2; kind++)
            for (int b = 0; b < ULCB_NBARK; b++) {
                int beg, end;
                if (kind == 0) band_edges((float)b, (float)b + 2.0f, nyq, n, &beg, &end);
                else           band_edges((float)b - 0.75f, (float)b + 0.25f, nyq, n, &beg, &end);
                edges[((sh * 2 + kind) * ULCB_NBARK + b) * 2 + 0] = (int16_t)beg;
                edges[((sh * 2 + kind) * ULCB_NBARK + b) * 2 + 1] = (int16_t)end;
            }
        for (int kind = 0; kind < 2; kind++) {
            // the 50 band edges in walking order: by position; at equal positions begins first (an empty band's begin
            // must be recorded before its end is evaluated), then by band so that levels carry over in band order
            const int16_t *ed = &edges[((sh * 2 + kind) * ULCB_NBARK) * 2];
            std::vector<int> key;
            for (int b = 0; b < ULCB_NBARK; b++) { key.push_back((ed[2 * b] << 8) | (0 << 7) | b); key.push_back((ed[2 * b + 1] << 8) | (1 << 7) | b); }
            std::sort(key.begin(), key.end());
            int16_t *ev = &events[((sh * 2 + kind) * 2 * ULCB_NBARK) * 2];
            for (size_t e = 0; e < key.size(); e++) { ev[2 * e] = (int16_t)(key[e] >> 8); ev[2 * e + 1] = (int16_t)(key[e] & 0xFF); }
        }
        t.lineOff[sh] = (int)lband.size();
        for (int i = 0; i < n; i++) {                                  // ulcEncoder_Psyopt.c:136-152
            float bb = line_to_bark(i, nyq, n);
            int idx = (int)bb;
            lband.push_back((uint8_t)(idx > 255 ? 255 : idx));
            lfrac.push_back(bb - (float)idx);
        }
    }
    int rc;
    if ((rc = upload(ctx, edges, &t.edges))) return rc;
    if ((rc = upload(ctx, events, &t.events))) return rc;
    if ((rc = upload(ctx, lband, &t.lineBand))) return rc;
    if ((rc = upload(ctx, lfrac, &t.lineFrac))) return rc;
    return ULCB_OK;
}

int grow(ULCB_Context *ctx, void **p, size_t *cap, size_t need) {
    if (need <= *cap) return ULCB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CK(ctx, cudaMalloc(p, need));
    *cap = need;
    return ULCB_OK;
}

// Built: BlockSize 256..8192; mono and stereo at every block size, 3..8 channels while nChan * BlockSize <= 16384 (the generic
// decoder keeps a whole block of every channel in shared memory and gives each thread at most 16 coefficients).
bool supported(int nChan, int N) {
    if (N < 256 || nChan < 1) return false;
    if (N <= 8192 && nChan <= 2) return true;
    return nChan <= 8 && N <= 16384 && nChan * N <= 16384;         // 3..8 channels, and BlockSize 16384 in mono (same work set as 8192 stereo)
}
bool valid(int nChan, int N) {
    return nChan >= 1 && nChan <= 255 && N >= 256 && N <= 32768 && (N & (N - 1)) == 0;
}

__global__ void dct4_test_kernel(float *data, int S, int sh, UlcbTables tb) {
    extern __shared__ __align__(16) float sbuf[];
    float *g = data + (size_t)blockIdx.x * S;
    for (int i = threadIdx.x; i < S; i += blockDim.x) sbuf[i] = g[i];
    __syncthreads();
    dct4_inplace<8>(sbuf, S, 1, S, sh, tb);
    for (int i = threadIdx.x; i < S; i += blockDim.x) g[i] = sbuf[i];
}

}  // namespace

extern "C" {

const char *ULCB_ErrorString(int status) {
    switch (status) {
        case ULCB_OK: return "ok";
        case ULCB_E_PARAM: return "bad parameter";
        case ULCB_E_UNSUPPORTED: return "configuration not supported by this build";
        case ULCB_E_NOMEM: return "out of memory";
        case ULCB_E_CUDA: return "CUDA runtime error";
        case ULCB_E_NODEVICE: return "no sm_100 GPU available (there is no CPU fallback)";
    }
    return "unknown";
}

size_t ULCB_MaxBlockBytes(int nChan, int blockSize) { return (size_t)2 * nChan * blockSize + 8; }

int ULCB_ContextCreate(ULCB_Context **out, int device, int rateHz, int nChan, int blockSize, uint32_t maxBlocksPerStream) {
    if (!out) return ULCB_E_PARAM;
    *out = nullptr;
    if (!valid(nChan, blockSize) || rateHz < 1 || maxBlocksPerStream < 1) return ULCB_E_PARAM;
    if (!supported(nChan, blockSize)) return ULCB_E_UNSUPPORTED;
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || device < 0 || device >= nDev) return ULCB_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return ULCB_E_NODEVICE;
    ULCB_Context *ctx = new (std::nothrow) ULCB_Context();
    if (!ctx) return ULCB_E_NOMEM;
    ctx->device = device; ctx->rateHz = rateHz; ctx->nChan = nChan; ctx->N = blockSize; ctx->maxBlocks = maxBlocksPerStream;
    int rc = ULCB_OK;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) rc = ULCB_E_CUDA;
    ctx->smCount = prop.multiProcessorCount;
    ctx->bpChunks = getenv("ULCB_BP_CHUNKS") ? atoi(getenv("ULCB_BP_CHUNKS")) : ULCB_BP_CHUNKS_DEFAULT;
    ctx->bpOverlap = getenv("ULCB_BP_OVERLAP") ? atoi(getenv("ULCB_BP_OVERLAP")) : ULCB_BP_OVERLAP_DEFAULT;
    if (ctx->bpChunks < 0 || ctx->bpChunks > 1024) ctx->bpChunks = 0;
    if (ctx->bpOverlap < 0 || ctx->bpOverlap > 2) ctx->bpOverlap = 0;
    if (rc == ULCB_OK) rc = build_tables(ctx);
    if (rc == ULCB_OK) rc = build_enc_tables(ctx);
    if (rc != ULCB_OK) { ULCB_ContextDestroy(ctx); return rc; }
    *out = ctx;
    return ULCB_OK;
}

void ULCB_ContextDestroy(ULCB_Context *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    for (void *p : ctx->owned) cudaFree(p);
    if (ctx->dIn) cudaFree(ctx->dIn);
    if (ctx->dOut) cudaFree(ctx->dOut);
    if (ctx->dAux) cudaFree(ctx->dAux);
    if (ctx->dSign) cudaFree(ctx->dSign);
    if (ctx->bpOff) cudaFree(ctx->bpOff);
    if (ctx->bpDesc) cudaFree(ctx->bpDesc);
    if (ctx->bpBad) cudaFree(ctx->bpBad);
    if (ctx->bpArena) cudaFree(ctx->bpArena);
    if (ctx->hPlan) cudaFreeHost(ctx->hPlan);
    if (ctx->encWs) cudaFree(ctx->encWs);
    if (ctx->corpusLut) cudaFree(ctx->corpusLut);
    if (ctx->dAvgCx) cudaFree(ctx->dAvgCx);
    if (ctx->dCxScratch) cudaFree(ctx->dCxScratch);
    if (ctx->detScratch) cudaFree(ctx->detScratch);
    if (ctx->dWcAll) cudaFree(ctx->dWcAll);
    for (auto &sp : ctx->spans) { if (sp.a) cudaEventDestroy(sp.a); if (own_end_event(sp.kind) || sp.ownB) cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : ctx->evPool) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->bpSyncEv) cudaEventDestroy(e);
    if (ctx->bpFront) cudaStreamDestroy(ctx->bpFront);
    if (ctx->bpCopy) cudaStreamDestroy(ctx->bpCopy);
    delete ctx;
}

int ULCB_LastCudaError(const ULCB_Context *ctx) { return ctx ? ctx->lastCuda : 0; }
void *ULCB_Stream(ULCB_Context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t ULCB_LaunchCount(const ULCB_Context *ctx) { return ctx ? ctx->launches : 0; }

int ULCB_Sync(ULCB_Context *ctx) {
    if (!ctx) return ULCB_E_PARAM;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ULCB_OK;
}

// Block-parallel pipeline (ulcb_decode_bp.cu).  hOffs / hNBlocks: host copies of the stream offsets / block counts when the
// caller has them; otherwise they are read back from the device (one small copy + synchronise) to plan the waves.
static int decode_bp(ULCB_Context *ctx, int nStreams, const uint8_t *dBytes, const uint64_t *dStreamOffs, const uint32_t *dNBlocks,
                     uint32_t nBlocksAll, const uint32_t *dBlockBytes, uint64_t blockBytesStride, void *dOut, uint64_t outStrideElems,
                     int format, int32_t *dStatus, uint32_t *dBlockBits, uint64_t bitsStride, float *dbgCoefs,
                     const uint64_t *hOffs, const uint32_t *hNBlocks, const uint8_t *hBytes = nullptr) {
    int rc;
    if (!hOffs || (dNBlocks && !hNBlocks)) {
        const size_t need = sizeof(uint64_t) * (nStreams + 1) + sizeof(uint32_t) * nStreams;
        if (ctx->hPlanCap < need) {
            if (ctx->hPlan) cudaFreeHost(ctx->hPlan);
            ctx->hPlan = nullptr; ctx->hPlanCap = 0;
            CK(ctx, cudaMallocHost(&ctx->hPlan, need));
            ctx->hPlanCap = need;
        }
        uint64_t *po = (uint64_t *)ctx->hPlan;
        uint32_t *pn = (uint32_t *)(po + nStreams + 1);
        if (!hOffs) { CK(ctx, cudaMemcpyAsync(po, dStreamOffs, sizeof(uint64_t) * (nStreams + 1), cudaMemcpyDeviceToHost, ctx->stream)); hOffs = po; }
        if (dNBlocks && !hNBlocks) { CK(ctx, cudaMemcpyAsync(pn, dNBlocks, sizeof(uint32_t) * nStreams, cudaMemcpyDeviceToHost, ctx->stream)); hNBlocks = pn; }
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    uint32_t maxBlk = nBlocksAll;
    if (dNBlocks) { maxBlk = 0; for (int s = 0; s < nStreams; s++) if (hNBlocks[s] > maxBlk) maxBlk = hNBlocks[s]; }
    if (maxBlk == 0) return ULCB_OK;
    if (maxBlk > ctx->maxBlocks) return ULCB_E_PARAM;
    if (outStrideElems < (uint64_t)maxBlk * ctx->N * ctx->nChan) return ULCB_E_PARAM;
    for (int s = 0; s < nStreams; s++) if (hOffs[s + 1] < hOffs[s]) return ULCB_E_PARAM;

    // waves: as many streams as the arena cap holds
    static const size_t capEnv = getenv("ULCB_BP_ARENA_MB") ? (size_t)atoll(getenv("ULCB_BP_ARENA_MB")) << 20 : (size_t)32768 << 20;
    // waves never plan for more than half of what was free when the context first decoded (asked once: cudaMemGetInfo is a slow,
    // serialising driver call, and a step that starts with a tiny kernel has nothing queued to hide it behind - it showed as
    // 50 ms gaps in some steps)
    if (ctx->bpPlanCap == 0) {
        size_t fr = 0, tot = 0, cap0 = capEnv;
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) { const size_t lim = (fr + ctx->bpArenaCap) / 2; if (cap0 > lim) cap0 = lim; }
        ctx->bpPlanCap = cap0 ? cap0 : 1;
    }
    const size_t cap = ctx->bpPlanCap;
    static const int runLen = getenv("ULCB_BP_RUNLEN") ? atoi(getenv("ULCB_BP_RUNLEN")) : 64;
    UlcbBpArgs a{};
    a.tb = ctx->tb; a.nSignBits = ctx->tb.nSignBits; a.nChan = ctx->nChan;
    a.nBlocksAll = (int)nBlocksAll; a.maxBlk = (int)maxBlk; a.runLen = runLen < 1 ? 1 : runLen;
    a.bytes = dBytes; a.streamOffs = dStreamOffs; a.nBlocks = dNBlocks;
    a.wend = reinterpret_cast<const uint32_t *>((reinterpret_cast<uintptr_t>(dBytes) + hOffs[nStreams] + 8) & ~(uintptr_t)3);
    a.blockBytes = dBlockBytes; a.blockBytesStride = blockBytesStride;
    a.out = dOut; a.outStride = outStrideElems; a.status = dStatus; a.blockBits = dBlockBits; a.bitsStride = bitsStride;
    a.dbgCoefs = dbgCoefs;
    // ---- chunked pipeline (ULCB_SetDecodePipeline): the streams of the call in nChunks contiguous chunks; the upload (host-pointer
    // entry points) and the walk of chunk c+1 - with overlap 2 its parse and scan too, into the other half of the arena - run on
    // high-priority side streams beside the transform of chunk c.  The walk is one latency-bound lane per stream and the transform
    // leaves half of the issue slots idle (and, storing into a pinned host buffer, waits for PCIe): side by side they cost less
    // than one after the other.  Results do not depend on the chunking: every stream is decoded by the same kernels.
    int nChunks = ctx->bpChunks, overlap = ctx->bpOverlap;
    if (nChunks == 0) {
        nChunks = hBytes ? (nStreams + ULCB_BP_AUTO_HOST_CHUNK_STREAMS / 2) / ULCB_BP_AUTO_HOST_CHUNK_STREAMS : 1;
        if (nChunks > 1024) nChunks = 1024;
        overlap = hBytes ? ULCB_BP_AUTO_HOST_OVERLAP : 0;
    }
    if (overlap == 0) nChunks = 1;
    if (nChunks > nStreams / 64) nChunks = nStreams / 64;
    const int nBuf = overlap >= 2 ? 2 : 1;
    size_t chunkArena = 0;
    std::vector<int> cb;                                           // chunk boundaries
    while (nChunks > 1) {
        cb.assign(nChunks + 1, 0);
        for (int c = 0; c <= nChunks; c++) cb[c] = (int)((long long)nStreams * c / nChunks);
        chunkArena = 0;
        for (int c = 0; c < nChunks; c++) {
            const size_t need = ulcb_decode_bp_arena_bytes(hOffs[cb[c + 1]] - hOffs[cb[c]], cb[c + 1] - cb[c], (int)maxBlk);
            if (need > chunkArena) chunkArena = need;
        }
        chunkArena = (chunkArena + 255) & ~(size_t)255;
        if (chunkArena * nBuf <= cap) break;
        nChunks = nChunks * 2 <= nStreams ? nChunks * 2 : 1;      // smaller chunks until one fits; else the serial waves below
    }
    if (nChunks > 1) {
        if (!ctx->bpFront) {
            int lo = 0, hi = 0;
            CK(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(ctx, cudaStreamCreateWithPriority(&ctx->bpFront, cudaStreamNonBlocking, hi));
            CK(ctx, cudaStreamCreateWithPriority(&ctx->bpCopy, cudaStreamNonBlocking, hi));
        }
        while (ctx->bpSyncEv.size() < (size_t)(3 * nChunks + 1)) {
            cudaEvent_t e = nullptr;
            CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->bpSyncEv.push_back(e);
        }
        int maxN = 0;
        for (int c = 0; c < nChunks; c++) if (cb[c + 1] - cb[c] > maxN) maxN = cb[c + 1] - cb[c];
        const size_t descPer = (sizeof(UlcbBpDesc) * (size_t)maxN * maxBlk + 255) & ~(size_t)255;
        const size_t badPer = (sizeof(int32_t) * (size_t)maxN + 255) & ~(size_t)255;
        if ((rc = grow(ctx, &ctx->bpOff, &ctx->bpOffCap, sizeof(uint32_t) * (size_t)nStreams * (maxBlk + 1)))) return rc;
        if ((rc = grow(ctx, &ctx->bpDesc, &ctx->bpDescCap, descPer * nBuf))) return rc;
        if ((rc = grow(ctx, &ctx->bpBad, &ctx->bpBadCap, badPer * nBuf))) return rc;
        if ((rc = grow(ctx, &ctx->bpArena, &ctx->bpArenaCap, chunkArena * nBuf))) return rc;
        cudaEvent_t *evC = ctx->bpSyncEv.data(), *evF = evC + nChunks, *evT = evF + nChunks, evStart = ctx->bpSyncEv[3 * nChunks];
        cudaStream_t back = ctx->stream, front = ctx->bpFront, copy = ctx->bpCopy;
        CK(ctx, cudaEventRecord(evStart, back));                   // the side streams start behind what the caller queued
        CK(ctx, cudaStreamWaitEvent(front, evStart, 0));
        if (hBytes) CK(ctx, cudaStreamWaitEvent(copy, evStart, 0));
        for (int c = 0; c < nChunks; c++) {
            const int c0 = cb[c], n = cb[c + 1] - c0, buf = c % nBuf;
            if (hBytes) {
                CK(ctx, cudaMemcpyAsync(const_cast<uint8_t *>(dBytes) + hOffs[c0], hBytes + hOffs[c0], hOffs[c0 + n] - hOffs[c0],
                                        cudaMemcpyHostToDevice, copy));
                CK(ctx, cudaEventRecord(evC[c], copy));
                CK(ctx, cudaStreamWaitEvent(front, evC[c], 0));
            }
            a.s0 = c0; a.nStreams = n; a.base0 = hOffs[c0];
            a.blkOff = (uint32_t *)ctx->bpOff + (size_t)c0 * (maxBlk + 1);
            a.desc = (UlcbBpDesc *)((char *)ctx->bpDesc + descPer * buf); a.firstBad = (int32_t *)((char *)ctx->bpBad + badPer * buf);
            a.arena = (uint8_t *)ctx->bpArena + chunkArena * buf;
            int nl = 0;
            cudaEvent_t e0 = nullptr, e1 = nullptr, evA[3] = {nullptr, nullptr, nullptr}, evB[4] = {nullptr, nullptr, nullptr, nullptr};
            if (ctx->timing) {
                e0 = timing_event(ctx); e1 = timing_event(ctx);
                for (auto &e : evA) e = timing_event(ctx);
                evB[2] = timing_event(ctx); evB[3] = timing_event(ctx);
                cudaEventRecord(e0, front);
            }
            CK(ctx, ulcb_decode_bp_offsets(a, dBlockBytes != nullptr, ctx->smCount, front, &nl));
            if (ctx->timing) { cudaEventRecord(e1, front); ctx->spans.push_back({e0, e1, dBlockBytes ? ULCB_KERNEL_BP_INDEX : ULCB_KERNEL_BP_WALK}); }
            if (nBuf == 2) {
                if (c >= 2) CK(ctx, cudaStreamWaitEvent(front, evT[c - 2], 0));   // this half of the arena is free again
                CK(ctx, ulcb_decode_bp_launch(a, format == ULCB_FMT_PCM16, front, &nl, ctx->timing ? evA : nullptr, 1));
            }
            CK(ctx, cudaEventRecord(evF[c], front));
            CK(ctx, cudaStreamWaitEvent(back, evF[c], 0));
            if (nBuf == 1) CK(ctx, ulcb_decode_bp_launch(a, format == ULCB_FMT_PCM16, back, &nl, ctx->timing ? evA : nullptr, 1));
            CK(ctx, ulcb_decode_bp_launch(a, format == ULCB_FMT_PCM16, back, &nl, ctx->timing ? evB : nullptr, 2));
            if (nBuf == 2) CK(ctx, cudaEventRecord(evT[c], back));
            if (ctx->timing) {
                ctx->spans.push_back({evA[0], evA[1], ULCB_KERNEL_BP_PARSE});
                ctx->spans.push_back({evA[1], evA[2], ULCB_KERNEL_BP_SCAN});
                ctx->spans.push_back({evA[2], nullptr, -1});
                ctx->spans.push_back({evB[2], evB[3], ULCB_KERNEL_BP_TRANSFORM, true});
            }
            ctx->launches += (uint64_t)nl;
        }
        return ULCB_OK;
    }
    if (hBytes) CK(ctx, cudaMemcpyAsync(const_cast<uint8_t *>(dBytes) + hOffs[0], hBytes + hOffs[0], hOffs[nStreams] - hOffs[0],
                                        cudaMemcpyHostToDevice, ctx->stream));
    {
        // block byte offsets of the whole call at once: the token-length walk of foreign streams is one lane per stream and
        // bound by latency, not by how many streams it walks
        int rc0;
        if ((rc0 = grow(ctx, &ctx->bpOff, &ctx->bpOffCap, sizeof(uint32_t) * (size_t)nStreams * (maxBlk + 1)))) return rc0;
        a.s0 = 0; a.nStreams = nStreams; a.base0 = hOffs[0]; a.blkOff = (uint32_t *)ctx->bpOff;
        int nl = 0;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (ctx->timing) { e0 = timing_event(ctx); e1 = timing_event(ctx); cudaEventRecord(e0, ctx->stream); }
        CK(ctx, ulcb_decode_bp_offsets(a, dBlockBytes != nullptr, ctx->smCount, ctx->stream, &nl));
        if (ctx->timing) { cudaEventRecord(e1, ctx->stream); ctx->spans.push_back({e0, e1, dBlockBytes ? ULCB_KERNEL_BP_INDEX : ULCB_KERNEL_BP_WALK}); }
        ctx->launches += (uint64_t)nl;
    }
    for (int s0 = 0; s0 < nStreams;) {
        int s1 = s0 + 1;
        while (s1 < nStreams && ulcb_decode_bp_arena_bytes(hOffs[s1 + 1] - hOffs[s0], s1 + 1 - s0, (int)maxBlk) <= cap) s1++;
        const int n = s1 - s0;
        if ((rc = grow(ctx, &ctx->bpDesc, &ctx->bpDescCap, sizeof(UlcbBpDesc) * (size_t)n * maxBlk))) return rc;
        if ((rc = grow(ctx, &ctx->bpBad, &ctx->bpBadCap, sizeof(int32_t) * (size_t)n))) return rc;
        if ((rc = grow(ctx, &ctx->bpArena, &ctx->bpArenaCap, ulcb_decode_bp_arena_bytes(hOffs[s1] - hOffs[s0], n, (int)maxBlk)))) return rc;
        a.s0 = s0; a.nStreams = n; a.base0 = hOffs[s0];
        a.blkOff = (uint32_t *)ctx->bpOff + (size_t)s0 * (maxBlk + 1); a.desc = (UlcbBpDesc *)ctx->bpDesc; a.firstBad = (int32_t *)ctx->bpBad; a.arena = (uint8_t *)ctx->bpArena;
        int nl = 0;
        cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
        if (ctx->timing) for (auto &e : ev) e = timing_event(ctx);
        CK(ctx, ulcb_decode_bp_launch(a, format == ULCB_FMT_PCM16, ctx->stream, &nl, ctx->timing ? ev : nullptr));
        if (ctx->timing) {
            // an event can close one span and open the next only once: the pool gets it back from the span that opened with it
            ctx->spans.push_back({ev[0], ev[1], ULCB_KERNEL_BP_PARSE});
            ctx->spans.push_back({ev[1], ev[2], ULCB_KERNEL_BP_SCAN});
            ctx->spans.push_back({ev[2], ev[3], ULCB_KERNEL_BP_TRANSFORM});
            ctx->spans.push_back({ev[3], nullptr, -1});
        }
        ctx->launches += (uint64_t)nl;
        s0 = s1;
    }
    return ULCB_OK;
}

static bool use_bp(const ULCB_Context *ctx) {
    static const bool forceSerial = getenv("ULCB_DECODE_GENERIC") != nullptr || getenv("ULCB_DECODE_WS") != nullptr;
    return !forceSerial && ulcb_decode_bp_supported(ctx->tb, ctx->nChan);
}

static int decode_device(ULCB_Context *ctx, int nStreams, const uint8_t *dBytes, const uint64_t *dStreamOffs,
                         const uint32_t *dNBlocks, uint32_t nBlocksAll, void *dOut, uint64_t outStrideElems,
                         int format, int32_t *dStatus, uint32_t *dBlockBits, uint64_t bitsStride,
                         UlcbStreamState *state, float *statePend, float *dbgCoefs = nullptr,
                         const uint32_t *dBlockBytes = nullptr, uint64_t blockBytesStride = 0,
                         const uint64_t *hOffs = nullptr, const uint32_t *hNBlocks = nullptr, const uint8_t *hBytes = nullptr) {
    if (!ctx || nStreams < 0 || !dBytes || !dStreamOffs || !dOut) return ULCB_E_PARAM;
    if (format != ULCB_FMT_PCM16 && format != ULCB_FMT_FLOAT32) return ULCB_E_PARAM;
    if (nStreams == 0) return ULCB_OK;
    // the noise-sign table covers maxBlocksPerStream blocks; per-stream counts on the device are bounded inside the kernels
    if (!dNBlocks && !state && nBlocksAll > ctx->maxBlocks) return ULCB_E_PARAM;
    if (dBlockBits && bitsStride < (dNBlocks ? 1u : nBlocksAll)) return ULCB_E_PARAM;
    if (dBlockBytes && blockBytesStride < (dNBlocks ? 1u : nBlocksAll)) return ULCB_E_PARAM;
    CK(ctx, cudaSetDevice(ctx->device));
    // kernel choice: the block-parallel pipeline where it applies (BlockSize 2048 stereo, whole fresh streams); the
    // stream-serial kernels for every other shape, for carried state (drop-in API), and on request (A/B timing, tests)
    if (!state && use_bp(ctx))
        return decode_bp(ctx, nStreams, dBytes, dStreamOffs, dNBlocks, nBlocksAll, dBlockBytes, blockBytesStride, dOut, outStrideElems,
                         format, dStatus, dBlockBits, bitsStride, dbgCoefs, hOffs, hNBlocks, hBytes);
    if (hBytes) CK(ctx, cudaMemcpyAsync(const_cast<uint8_t *>(dBytes) + hOffs[0], hBytes + hOffs[0], hOffs[nStreams] - hOffs[0], cudaMemcpyHostToDevice, ctx->stream));
    UlcbDecodeArgs a{};
    a.tb = ctx->tb;
    a.nChan = ctx->nChan;
    a.nStreams = nStreams; a.nBlocksAll = (int)nBlocksAll;
    a.bytes = dBytes; a.streamOffs = dStreamOffs; a.nBlocks = dNBlocks;
    a.out = dOut; a.outStride = outStrideElems;
    a.status = dStatus; a.bytesConsumed = nullptr; a.blockBits = dBlockBits; a.bitsStride = bitsStride;
    a.state = state; a.statePend = statePend; a.dbgCoefs = dbgCoefs;
    CK(ctx, ulcb_decode_launch(a, ctx->nChan, format == ULCB_FMT_PCM16, nStreams, ctx->stream));
    ctx->launches++;
    return ULCB_OK;
}

int ULCB_DecodeStreamsDevice(ULCB_Context *ctx, int nStreams, const uint8_t *dBytes, const uint64_t *dStreamOffs,
                             const uint32_t *dNBlocks, uint32_t nBlocksAll, void *dOut, uint64_t outStrideElems,
                             int format, int32_t *dStatus, uint32_t *dBlockBits, uint64_t bitsStride) {
    return decode_device(ctx, nStreams, dBytes, dStreamOffs, dNBlocks, nBlocksAll, dOut, outStrideElems, format,
                         dStatus, dBlockBits, bitsStride, nullptr, nullptr);
}

int ULCB_DecodeStreamsIndexedDevice(ULCB_Context *ctx, int nStreams, const uint8_t *dBytes, const uint64_t *dStreamOffs,
                                    const uint32_t *dNBlocks, uint32_t nBlocksAll, const uint32_t *dBlockBytes, uint64_t blockBytesStride,
                                    void *dOut, uint64_t outStrideElems, int format, int32_t *dStatus, uint32_t *dBlockBits,
                                    uint64_t bitsStride) {
    if (!dBlockBytes) return ULCB_E_PARAM;
    return decode_device(ctx, nStreams, dBytes, dStreamOffs, dNBlocks, nBlocksAll, dOut, outStrideElems, format,
                         dStatus, dBlockBits, bitsStride, nullptr, nullptr, nullptr, dBlockBytes, blockBytesStride);
}

static int decode_host(ULCB_Context *ctx, int nStreams, const uint8_t *bytes, const uint64_t *streamOffs,
                       const uint32_t *nBlocks, uint32_t nBlocksAll, const uint32_t *blockBytes, uint64_t blockBytesStride,
                       void *out, uint64_t outStrideElems, int format, int32_t *status, uint32_t *blockBits, uint64_t bitsStride) {
    if (!ctx || nStreams < 0 || !bytes || !streamOffs || !out) return ULCB_E_PARAM;
    if (format != ULCB_FMT_PCM16 && format != ULCB_FMT_FLOAT32) return ULCB_E_PARAM;
    if (nStreams == 0) return ULCB_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t elem = format == ULCB_FMT_PCM16 ? 2 : 4;
    const size_t nBytes = streamOffs[nStreams];
    uint32_t maxBlk = nBlocksAll;
    if (nBlocks) { maxBlk = 0; for (int s = 0; s < nStreams; s++) if (nBlocks[s] > maxBlk) maxBlk = nBlocks[s]; }
    if (outStrideElems < (uint64_t)maxBlk * ctx->N * ctx->nChan) return ULCB_E_PARAM;
    if (maxBlk > ctx->maxBlocks || (blockBits && bitsStride < maxBlk) || (blockBytes && blockBytesStride < maxBlk)) return ULCB_E_PARAM;
    for (int s = 0; s < nStreams; s++) if (streamOffs[s + 1] < streamOffs[s]) return ULCB_E_PARAM;
    const size_t outBytes = (size_t)nStreams * outStrideElems * elem;
    // aux layout: offs | nBlocks | status | bits | block-size index
    const size_t offOffs = 0, offNb = offOffs + sizeof(uint64_t) * (nStreams + 1);
    const size_t offSt = offNb + sizeof(uint32_t) * nStreams, offBits = (offSt + sizeof(int32_t) * nStreams + 15) & ~(size_t)15;
    const size_t offIdx = offBits + (blockBits ? sizeof(uint32_t) * nStreams * bitsStride : 0);
    const size_t auxBytes = offIdx + (blockBytes ? sizeof(uint32_t) * nStreams * blockBytesStride : 0);
    int rc;
    // PCM out is ~11x the bytes in.  When the caller's buffer is page-locked (cudaHostAlloc / cudaHostRegister) the kernel
    // stores straight into it over PCIe, so the transfer overlaps the decode instead of following it; pageable buffers
    // go through the device staging buffer and one D2H copy.
    void *dOutPtr = nullptr;
    {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer &&
            !getenv("ULCB_NO_ZEROCOPY")) dOutPtr = at.devicePointer;
        else (void)cudaGetLastError();
    }
    if ((rc = grow(ctx, &ctx->dIn, &ctx->dInCap, nBytes + 64))) return rc;
    if (!dOutPtr) { if ((rc = grow(ctx, &ctx->dOut, &ctx->dOutCap, outBytes))) return rc; }
    if ((rc = grow(ctx, &ctx->dAux, &ctx->dAuxCap, auxBytes))) return rc;
    char *aux = (char *)ctx->dAux;
    // the payload bytes are uploaded by decode_device: in one copy, or - when the PCM goes straight into the caller's page-locked
    // buffer, so that the transform runs at the PCIe rate - chunk by chunk inside the pipelined block-parallel decode
    const bool pipelined = dOutPtr != nullptr || ctx->bpChunks > 0;
    if (!pipelined) CK(ctx, cudaMemcpyAsync(ctx->dIn, bytes, nBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemsetAsync((char *)ctx->dIn + nBytes, 0, 64, ctx->stream));
    CK(ctx, cudaMemcpyAsync(aux + offOffs, streamOffs, sizeof(uint64_t) * (nStreams + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nBlocks) CK(ctx, cudaMemcpyAsync(aux + offNb, nBlocks, sizeof(uint32_t) * nStreams, cudaMemcpyHostToDevice, ctx->stream));
    if (blockBytes) CK(ctx, cudaMemcpyAsync(aux + offIdx, blockBytes, sizeof(uint32_t) * nStreams * blockBytesStride, cudaMemcpyHostToDevice, ctx->stream));
    rc = decode_device(ctx, nStreams, (const uint8_t *)ctx->dIn, (const uint64_t *)(aux + offOffs),
                       nBlocks ? (const uint32_t *)(aux + offNb) : nullptr, nBlocksAll, dOutPtr ? dOutPtr : ctx->dOut, outStrideElems, format,
                       (int32_t *)(aux + offSt), blockBits ? (uint32_t *)(aux + offBits) : nullptr, bitsStride, nullptr, nullptr, nullptr,
                       blockBytes ? (const uint32_t *)(aux + offIdx) : nullptr, blockBytesStride, streamOffs, nBlocks, pipelined ? bytes : nullptr);
    if (rc) return rc;
    if (!dOutPtr) CK(ctx, cudaMemcpyAsync(out, ctx->dOut, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) CK(ctx, cudaMemcpyAsync(status, aux + offSt, sizeof(int32_t) * nStreams, cudaMemcpyDeviceToHost, ctx->stream));
    if (blockBits) CK(ctx, cudaMemcpyAsync(blockBits, aux + offBits, sizeof(uint32_t) * nStreams * bitsStride, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ULCB_OK;
}

int ULCB_DecodeStreams(ULCB_Context *ctx, int nStreams, const uint8_t *bytes, const uint64_t *streamOffs,
                       const uint32_t *nBlocks, uint32_t nBlocksAll, void *out, uint64_t outStrideElems, int format,
                       int32_t *status, uint32_t *blockBits, uint64_t bitsStride) {
    return decode_host(ctx, nStreams, bytes, streamOffs, nBlocks, nBlocksAll, nullptr, 0, out, outStrideElems, format, status, blockBits, bitsStride);
}

int ULCB_DecodeStreamsIndexed(ULCB_Context *ctx, int nStreams, const uint8_t *bytes, const uint64_t *streamOffs,
                              const uint32_t *nBlocks, uint32_t nBlocksAll, const uint32_t *blockBytes, uint64_t blockBytesStride,
                              void *out, uint64_t outStrideElems, int format, int32_t *status, uint32_t *blockBits, uint64_t bitsStride) {
    if (!blockBytes) return ULCB_E_PARAM;
    return decode_host(ctx, nStreams, bytes, streamOffs, nBlocks, nBlocksAll, blockBytes, blockBytesStride, out, outStrideElems, format,
                       status, blockBits, bitsStride);
}

// The block index a .ulc file lacks: the token-length walk alone (no parse, no transform), any block size and channel count.
static int index_device(ULCB_Context *ctx, int nStreams, const uint8_t *dBytes, const uint64_t *dStreamOffs, const uint64_t *hOffs,
                        const uint32_t *dNBlocks, uint32_t maxBlk, uint32_t nBlocksAll, uint32_t *dBlockBytes, uint64_t stride, int32_t *dStatus) {
    int rc;
    if ((rc = grow(ctx, &ctx->bpOff, &ctx->bpOffCap, sizeof(uint32_t) * (size_t)nStreams * (maxBlk + 1)))) return rc;
    UlcbBpArgs a{};
    a.tb = ctx->tb; a.nChan = ctx->nChan;
    a.nBlocksAll = (int)nBlocksAll; a.maxBlk = (int)maxBlk; a.runLen = 1;
    a.bytes = dBytes; a.streamOffs = dStreamOffs; a.nBlocks = dNBlocks;
    a.wend = reinterpret_cast<const uint32_t *>((reinterpret_cast<uintptr_t>(dBytes) + hOffs[nStreams] + 8) & ~(uintptr_t)3);
    a.s0 = 0; a.nStreams = nStreams; a.base0 = hOffs[0]; a.blkOff = (uint32_t *)ctx->bpOff;
    int nl = 0;
    CK(ctx, ulcb_decode_bp_offsets(a, false, ctx->smCount, ctx->stream, &nl));
    CK(ctx, ulcb_decode_bp_sizes(a, dBlockBytes, stride, dStatus, ctx->stream));
    ctx->launches += (uint64_t)nl + 1;
    return ULCB_OK;
}

int ULCB_IndexStreams(ULCB_Context *ctx, int nStreams, const uint8_t *bytes, const uint64_t *streamOffs, const uint32_t *nBlocks,
                      uint32_t nBlocksAll, uint32_t *blockBytes, uint64_t blockBytesStride, int32_t *status) {
    if (!ctx || nStreams < 0 || !bytes || !streamOffs || !blockBytes) return ULCB_E_PARAM;
    if (nStreams == 0) return ULCB_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    uint32_t maxBlk = nBlocksAll;
    if (nBlocks) { maxBlk = 0; for (int s = 0; s < nStreams; s++) if (nBlocks[s] > maxBlk) maxBlk = nBlocks[s]; }
    if (maxBlk == 0) return ULCB_OK;
    if (blockBytesStride < maxBlk) return ULCB_E_PARAM;
    for (int s = 0; s < nStreams; s++) if (streamOffs[s + 1] < streamOffs[s]) return ULCB_E_PARAM;
    const size_t nBytes = streamOffs[nStreams];
    const size_t offOffs = 0, offNb = offOffs + sizeof(uint64_t) * (nStreams + 1), offSt = offNb + sizeof(uint32_t) * nStreams;
    const size_t offIdx = (offSt + sizeof(int32_t) * nStreams + 15) & ~(size_t)15;
    const size_t auxBytes = offIdx + sizeof(uint32_t) * (size_t)nStreams * blockBytesStride;
    int rc;
    if ((rc = grow(ctx, &ctx->dIn, &ctx->dInCap, nBytes + 64))) return rc;
    if ((rc = grow(ctx, &ctx->dAux, &ctx->dAuxCap, auxBytes))) return rc;
    char *aux = (char *)ctx->dAux;
    CK(ctx, cudaMemcpyAsync(ctx->dIn, bytes, nBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemsetAsync((char *)ctx->dIn + nBytes, 0, 64, ctx->stream));
    CK(ctx, cudaMemcpyAsync(aux + offOffs, streamOffs, sizeof(uint64_t) * (nStreams + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nBlocks) CK(ctx, cudaMemcpyAsync(aux + offNb, nBlocks, sizeof(uint32_t) * nStreams, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemsetAsync(aux + offIdx, 0, sizeof(uint32_t) * (size_t)nStreams * blockBytesStride, ctx->stream));
    rc = index_device(ctx, nStreams, (const uint8_t *)ctx->dIn, (const uint64_t *)(aux + offOffs), streamOffs,
                      nBlocks ? (const uint32_t *)(aux + offNb) : nullptr, maxBlk, nBlocksAll, (uint32_t *)(aux + offIdx), blockBytesStride,
                      (int32_t *)(aux + offSt));
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(blockBytes, aux + offIdx, sizeof(uint32_t) * (size_t)nStreams * blockBytesStride, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) CK(ctx, cudaMemcpyAsync(status, aux + offSt, sizeof(int32_t) * nStreams, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ULCB_OK;
}

int ULCB_SetKernelTiming(ULCB_Context *ctx, int enable) {
    if (!ctx) return ULCB_E_PARAM;
    ctx->timing = enable != 0;
    return ULCB_OK;
}

int ULCB_SetDecodePipeline(ULCB_Context *ctx, int chunks, int overlap) {
    if (!ctx || chunks < 0 || chunks > 1024 || overlap < 0 || overlap > 2) return ULCB_E_PARAM;
    ctx->bpChunks = chunks; ctx->bpOverlap = overlap;
    return ULCB_OK;
}

int ULCB_GetDecodePipeline(const ULCB_Context *ctx, int *chunks, int *overlap) {
    if (!ctx) return ULCB_E_PARAM;
    if (chunks) *chunks = ctx->bpChunks;
    if (overlap) *overlap = ctx->bpOverlap;
    return ULCB_OK;
}

int ULCB_KernelTimes(ULCB_Context *ctx, double *ms, uint64_t *count, int reset) {
    if (!ctx) return ULCB_E_PARAM;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &sp : ctx->spans) {
        if (sp.kind >= 0 && sp.b) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) { ctx->kernMs[sp.kind] += t; ctx->kernCount[sp.kind]++; }
            else (void)cudaGetLastError();
        }
    }
    // every event opens exactly one span entry (the last of a chain opens an entry without an end): back to the pool
    for (auto &sp : ctx->spans) ctx->evPool.push_back(sp.a);
    for (auto &sp : ctx->spans) if (own_end_event(sp.kind) || sp.ownB) ctx->evPool.push_back(sp.b);
    ctx->spans.clear();
    for (int k = 0; k < ULCB_KERNEL_KINDS; k++) {
        if (ms) ms[k] = ctx->kernMs[k];
        if (count) count[k] = ctx->kernCount[k];
        if (reset) { ctx->kernMs[k] = 0; ctx->kernCount[k] = 0; }
    }
    return ULCB_OK;
}

int ULCB_TestParse(ULCB_Context *ctx, const uint8_t *bytes, uint64_t nBytes, uint32_t nBlocks, float *coefs, uint32_t *blockBits) {
    if (!ctx || !bytes || !coefs || !blockBits || nBlocks < 1) return ULCB_E_PARAM;
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t per = (size_t)ctx->nChan * ctx->N;
    const size_t cBytes = sizeof(float) * per * nBlocks, auxBytes = 2 * sizeof(uint64_t) + sizeof(uint32_t) * nBlocks;
    int rc;
    if ((rc = grow(ctx, &ctx->dIn, &ctx->dInCap, nBytes + 64))) return rc;
    if ((rc = grow(ctx, &ctx->dOut, &ctx->dOutCap, 2 * cBytes))) return rc;
    if ((rc = grow(ctx, &ctx->dAux, &ctx->dAuxCap, auxBytes))) return rc;
    uint64_t offs[2] = {0, nBytes};
    CK(ctx, cudaMemcpyAsync(ctx->dIn, bytes, nBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemsetAsync((char *)ctx->dIn + nBytes, 0, 64, ctx->stream));
    CK(ctx, cudaMemcpyAsync(ctx->dAux, offs, sizeof offs, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t *dBits = (uint32_t *)((char *)ctx->dAux + 2 * sizeof(uint64_t));
    float *dCoefs = (float *)((char *)ctx->dOut + cBytes);
    rc = decode_device(ctx, 1, (const uint8_t *)ctx->dIn, (const uint64_t *)ctx->dAux, nullptr, nBlocks, ctx->dOut, per * nBlocks,
                       ULCB_FMT_FLOAT32, nullptr, dBits, nBlocks, nullptr, nullptr, dCoefs);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(coefs, dCoefs, cBytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaMemcpyAsync(blockBits, dBits, sizeof(uint32_t) * nBlocks, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ULCB_OK;
}

int ULCB_TestDCT4(ULCB_Context *ctx, float *data, int size, int count) {
    if (!ctx || !data || count < 1) return ULCB_E_PARAM;
    int sh = -1;
    for (int i = 0; i < 4; i++) if ((ctx->N >> i) == size) sh = i;
    if (sh < 0) return ULCB_E_PARAM;
    CK(ctx, cudaSetDevice(ctx->device));
    size_t bytes = sizeof(float) * (size_t)size * count;
    int rc;
    if ((rc = grow(ctx, &ctx->dOut, &ctx->dOutCap, bytes))) return rc;
    CK(ctx, cudaMemcpyAsync(ctx->dOut, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaFuncSetAttribute(dct4_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * size)));
    dct4_test_kernel<<<count, size / 32 < 32 ? 32 : (size / 32 > 256 ? 256 : size / 32), sizeof(float) * size, ctx->stream>>>((float *)ctx->dOut, size, sh, ctx->tb);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    CK(ctx, cudaMemcpyAsync(data, ctx->dOut, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ULCB_OK;
}

// ------------------------------------------------------------------------------------------------
// libulc drop-in: decoder (reference: include/ulcDecoder.h, libulc/ulcDecoder.c)
// ------------------------------------------------------------------------------------------------
struct DropinDecoder {
    ULCB_Context *ctx;
    UlcbStreamState *dState;
    float *dPend;
    uint8_t *dSrc; float *dDst; uint64_t *dOffs; uint32_t *dBits;
    uint8_t *hSrc; float *hDst; uint32_t *hMeta;      // pinned; hMeta = {bits, pad, pad, pad, UlcbStreamState}
    size_t srcBytes;
    uint64_t calls;                                    // blocks decoded so far: bounds the noise draws consumed
};

int ULC_DecoderState_Init(struct ULC_DecoderState_t *State) {
    State->BufferData = NULL;                                  // ulcDecoder.c:28
    State->TransformBuffer = State->TransformTemp = State->TransformInvLap = NULL;
    int nChan = State->nChan, N = State->BlockSize;
    if (!valid(nChan, N)) return -1;                           // ulcDecoder.c:30-35
    DropinDecoder *d = (DropinDecoder *)calloc(1, sizeof(DropinDecoder));
    if (!d) return -1;
    int dev = 0;
    cudaGetDevice(&dev);
    // the sign table starts at 1024 blocks and grows with the stream (ULC_DecodeBlock)
    if (ULCB_ContextCreate(&d->ctx, dev, 44100, nChan, N, 1u << 10) != ULCB_OK) { free(d); return -1; }
    size_t cap = ULCB_MaxBlockBytes(nChan, N);
    d->srcBytes = cap < 16384 ? cap : 16384;
    bool ok = cudaMalloc(&d->dState, sizeof(UlcbStreamState)) == cudaSuccess &&
              cudaMalloc(&d->dPend, sizeof(float) * nChan * (N / 2)) == cudaSuccess &&
              cudaMalloc(&d->dSrc, d->srcBytes + 64) == cudaSuccess &&
              cudaMalloc(&d->dDst, sizeof(float) * nChan * N) == cudaSuccess &&
              cudaMalloc(&d->dOffs, 2 * sizeof(uint64_t)) == cudaSuccess &&
              cudaMalloc(&d->dBits, sizeof(uint32_t)) == cudaSuccess &&
              cudaMallocHost(&d->hSrc, d->srcBytes) == cudaSuccess &&
              cudaMallocHost(&d->hDst, sizeof(float) * nChan * N) == cudaSuccess &&
              cudaMallocHost(&d->hMeta, 8 * sizeof(uint32_t)) == cudaSuccess;
    if (ok) {
        uint64_t offs[2] = {0, d->srcBytes};
        ok = cudaMemset(d->dState, 0, sizeof(UlcbStreamState)) == cudaSuccess &&
             cudaMemset(d->dSrc, 0, d->srcBytes + 64) == cudaSuccess &&
             cudaMemcpy(d->dOffs, offs, sizeof offs, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    State->BufferData = d;
    if (!ok) { ULC_DecoderState_Destroy(State); return -1; }
    State->LastSubBlockSize = 0;                               // ulcDecoder.c:52
    return 1;
}

void ULC_DecoderState_Destroy(struct ULC_DecoderState_t *State) {
    DropinDecoder *d = (DropinDecoder *)State->BufferData;
    if (!d) return;
    cudaFree(d->dState); cudaFree(d->dPend); cudaFree(d->dSrc); cudaFree(d->dDst); cudaFree(d->dOffs); cudaFree(d->dBits);
    cudaFreeHost(d->hSrc); cudaFreeHost(d->hDst); cudaFreeHost(d->hMeta);
    ULCB_ContextDestroy(d->ctx);
    free(d);
    State->BufferData = NULL;
}

int ULC_DecodeBlock(struct ULC_DecoderState_t *State, float *DstData, const void *SrcBuffer) {
    DropinDecoder *d = (DropinDecoder *)State->BufferData;
    if (!d) return 0;
    ULCB_Context *ctx = d->ctx;
    const int n = State->nChan * State->BlockSize;
    // a stream may be arbitrarily long: keep the noise-sign table ahead of the draws it can have consumed (at most one per
    // coefficient and block so far, libulc/ulcDecoder.c:156-184)
    const uint64_t need = (d->calls + 2) * (uint64_t)n + 256;
    if (need > ctx->tb.nSignBits && build_sign_table(ctx, need > 4 * ctx->tb.nSignBits ? need : 4 * ctx->tb.nSignBits) != ULCB_OK) return 0;
    memcpy(d->hSrc, SrcBuffer, d->srcBytes);
    if (cudaMemcpyAsync(d->dSrc, d->hSrc, d->srcBytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return 0;
    if (decode_device(ctx, 1, d->dSrc, d->dOffs, nullptr, 1, d->dDst, (uint64_t)n, ULCB_FMT_FLOAT32, nullptr, d->dBits, 1,
                      d->dState, d->dPend) != ULCB_OK) return 0;
    cudaMemcpyAsync(d->hDst, d->dDst, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&d->hMeta[0], d->dBits, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&d->hMeta[4], d->dState, sizeof(UlcbStreamState), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;
    memcpy(DstData, d->hDst, sizeof(float) * n);
    UlcbStreamState st;
    memcpy(&st, &d->hMeta[4], sizeof st);
    State->LastSubBlockSize = st.lastSubBlockSize;
    d->calls++;
    return (int)d->hMeta[0];
}

// ------------------------------------------------------------------------------------------------
// Synthetic corpus (measurement input) - device twin of ulc_codec_b200/corpus.py
// ------------------------------------------------------------------------------------------------
int ULCB_ConvertSamplesDevice(ULCB_Context *ctx, int format, int toFloat, const void *dSrc, void *dDst, uint64_t nSamples) {
    if (!ctx || !dSrc || !dDst || format < 0 || format > ULCB_FMT_PCM24) return ULCB_E_PARAM;
    if (nSamples == 0) return ULCB_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    if (toFloat) CK(ctx, ulcb_pcm_to_float(static_cast<float *>(dDst), dSrc, (size_t)nSamples, format, ctx->smCount, ctx->stream));
    else         CK(ctx, ulcb_pcm_from_float(dDst, static_cast<const float *>(dSrc), (size_t)nSamples, format, ctx->smCount, ctx->stream));
    ctx->launches++;
    return ULCB_OK;
}

int ULCB_GenerateCorpusDevice(ULCB_Context *ctx, int kind, uint32_t firstStreamId, int nStreams, uint64_t nSamples, uint64_t t0,
                              const int16_t *sineLut1024, const int16_t *burstLut882, int16_t *dOut, uint64_t strideElems) {
    if (!ctx || !dOut || !sineLut1024 || !burstLut882 || nStreams < 0 || (kind != 0 && kind != 1) || ctx->nChan != 2) return ULCB_E_PARAM;
    if (strideElems < 2 * nSamples || (strideElems & 1)) return ULCB_E_PARAM;
    CK(ctx, cudaSetDevice(ctx->device));
    if (!ctx->corpusLut) CK(ctx, cudaMalloc(&ctx->corpusLut, sizeof(int16_t) * (1024 + 896)));
    CK(ctx, cudaMemcpyAsync(ctx->corpusLut, sineLut1024, sizeof(int16_t) * 1024, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(ctx->corpusLut + 1024, burstLut882, sizeof(int16_t) * 882, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));                 // the LUT arguments may be pageable host memory
    CK(ctx, ulcb_corpus_launch(dOut, strideElems, kind, firstStreamId, nStreams, nSamples, t0, ctx->corpusLut, ctx->corpusLut + 1024, ctx->stream));
    ctx->launches++;
    return ULCB_OK;
}

// ------------------------------------------------------------------------------------------------
// Batched encoder
// ------------------------------------------------------------------------------------------------
uint32_t ULCB_BlocksForSamples(uint64_t nSamples, int blockSize) {       // tools/ulcEncodeTool.c:98
    return (uint32_t)((nSamples + (uint64_t)blockSize - 1) / (uint64_t)blockSize) + 2u;
}

struct EncTaps { float *coefs, *keys, *noise, *mask; int32_t *ranks, *counts; };

static int encode_device(ULCB_Context *ctx, int nStreams, const void *dPcm, int format, const uint64_t *dNSamples,
                         uint64_t inStrideElems, const uint32_t *dNBlocks, uint32_t nBlocksAll, const ULCB_EncodeParams *prm,
                         uint8_t *dOut, uint64_t outStrideBytes, uint32_t *dBlockBytes, float *dComplexity, int32_t *dWindowCtrl,
                         uint64_t metaStride, uint64_t *dTotalBytes, int32_t *dStatus,
                         UlcbEncStreamState *state, float *stateSamples, const EncTaps *taps) {
    if (!ctx || nStreams < 0 || !dPcm || !dOut || !prm) return ULCB_E_PARAM;
    if (format != ULCB_FMT_PCM16 && format != ULCB_FMT_FLOAT32) return ULCB_E_PARAM;
    if (prm->mode != ULCB_MODE_VBR && prm->mode != ULCB_MODE_CBR && prm->mode != ULCB_MODE_ABR && prm->mode != ULCB_MODE_ABR_2PASS) return ULCB_E_PARAM;
    if (prm->mode == ULCB_MODE_ABR_2PASS && (state || taps)) return ULCB_E_PARAM;     // whole streams only
    if (prm->mode == ULCB_MODE_VBR && !(prm->rate > 0.0f)) return ULCB_E_PARAM;
    if (prm->mode == ULCB_MODE_ABR && !(prm->avgComplexity > 0.0f)) return ULCB_E_PARAM;
    if (outStrideBytes == 0) return ULCB_E_PARAM;
    // without per-stream sample counts every stream holds nBlocks*BlockSize sample points: they must lie inside its slot
    if (!dNSamples && !dNBlocks && inStrideElems < (uint64_t)nBlocksAll * ctx->N * ctx->nChan) return ULCB_E_PARAM;
    if ((dBlockBytes || dComplexity || dWindowCtrl) && !dNBlocks && metaStride < nBlocksAll) return ULCB_E_PARAM;
    if (nStreams == 0) return ULCB_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    UlcbEncodeArgs a{};
    a.tb = ctx->etb;
    a.nChan = ctx->nChan;
    a.nStreams = nStreams; a.nBlocksAll = (int)nBlocksAll; a.nBlocks = dNBlocks;
    a.pcm = dPcm; a.pcmFloat = format == ULCB_FMT_FLOAT32; a.inStride = inStrideElems; a.nSamples = dNSamples;
    a.mode = prm->mode;
    a.param = prm->rate;
    if (prm->mode == ULCB_MODE_VBR) {                                   // libulc/ulcEncoder.c:146-148, host libm logf
        volatile float lg = logf(100.0f / prm->rate);
        a.param = 0x1.E4EFB8p3f * lg;   // the reference writes 0x1.E4EFB7p3f (25 bits); gcc rounds it to this float
    }
    a.avgComplexity = prm->avgComplexity;
    a.out = dOut; a.outStride = outStrideBytes; a.blockBytes = dBlockBytes; a.complexity = dComplexity; a.windowCtrl = dWindowCtrl;
    a.metaStride = metaStride; a.totalBytes = dTotalBytes; a.status = dStatus;
    a.state = state; a.stateSamples = stateSamples;
    if (taps) { a.dbgCoefs = taps->coefs; a.dbgKeys = taps->keys; a.dbgNoise = taps->noise; a.dbgMask = taps->mask; a.dbgRanks = taps->ranks; a.dbgCounts = taps->counts; }
#ifdef ULCB_PHASE_TIMING
    if (getenv("ULCB_DBG_FLAGS")) a.dbgFlags = atoi(getenv("ULCB_DBG_FLAGS"));
#endif
    int grid = nStreams;
    // Whole streams from their start (no carried state), every stream nBlocksAll blocks, stereo / mono whose work set fits shared
    // memory without the detector's buffers: the transient detector runs as a pre-pass over all streams and the encode kernel
    // takes its decisions (ULCB_ENC_PRE=0 keeps the detector inside the encode kernel, as calls with carried state do).
    static const bool preOff = getenv("ULCB_ENC_PRE") && atoi(getenv("ULCB_ENC_PRE")) == 0;
    const bool pre = !preOff && !state && !dNBlocks && nBlocksAll > 0 && ulcb_detect_supported(ctx->N, ctx->nChan);
    if (pre ? ulcb_encode_pre_uses_global_ws(ctx->N, ctx->nChan) : ulcb_encode_uses_global_ws(ctx->N, ctx->nChan)) {
        const int cap = ctx->smCount * 2;
        if (grid > cap) grid = cap;
        const size_t per = ulcb_encode_ws_floats(ctx->N, ctx->nChan);
        if (ctx->encGrid < grid || ctx->encWsFloats != per) {
            if (ctx->encWs) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->encWs); ctx->encWs = nullptr; }
            CK(ctx, cudaMalloc(&ctx->encWs, sizeof(float) * per * (size_t)(cap > grid ? cap : grid)));
            ctx->encGrid = cap > grid ? cap : grid; ctx->encWsFloats = per;
        }
        a.workspace = ctx->encWs; a.wsFloats = per;
    }
    if (pre) {
        const size_t needScr = ulcb_detect_scratch_bytes(ctx->N, nStreams, ctx->smCount), needWc = (size_t)nStreams * nBlocksAll;
        if (ctx->detScratchCap < needScr) {
            if (ctx->detScratch) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->detScratch); ctx->detScratch = nullptr; ctx->detScratchCap = 0; }
            CK(ctx, cudaMalloc(&ctx->detScratch, needScr));
            ctx->detScratchCap = needScr;
        }
        if (ctx->wcAllCap < needWc) {
            if (ctx->dWcAll) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->dWcAll); ctx->dWcAll = nullptr; ctx->wcAllCap = 0; }
            CK(ctx, cudaMalloc(&ctx->dWcAll, sizeof(int32_t) * needWc));
            ctx->wcAllCap = needWc;
        }
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        if (ctx->timing) { t0 = timing_event(ctx); t1 = timing_event(ctx); cudaEventRecord(t0, ctx->stream); }
        int nl = 0;
        CK(ctx, ulcb_detect_launch(a, ctx->nChan, ctx->smCount, ctx->detScratch, ctx->dWcAll, nBlocksAll, ctx->stream, &nl));
        if (ctx->timing) { cudaEventRecord(t1, ctx->stream); ctx->spans.push_back({t0, t1, ULCB_KERNEL_ENC_DETECT}); }
        ctx->launches += (uint64_t)nl;
        a.wcAll = ctx->dWcAll; a.wcStride = nBlocksAll;
        if (getenv("ULCB_ENC_THREADS")) a.encThreads = atoi(getenv("ULCB_ENC_THREADS"));
    }
    // The reference adds the Bark-analysis lines one by one in f64 and so does the kernel.  ULCB_ENC_FAST_BANDS=1 switches whole-stream
    // batches to segmented sums (band sums equal to the last bits of an f64, see the kernel): no extra bitstream flips in the
    // conformance run, but no gain either while two CTAs per SM hide the chains - kept for shapes with one CTA per SM.
    static const bool fastEnv = getenv("ULCB_ENC_FAST_BANDS") && atoi(getenv("ULCB_ENC_FAST_BANDS")) != 0;
    a.exactBands = (!pre || !fastEnv) ? 1 : 0;
    if (prm->mode == ULCB_MODE_ABR_2PASS) {
        // pass 1: analysis only -> block complexities; mean per stream on the device; pass 2: ABR with the per-stream means
        if (ctx->avgCap < nStreams) {
            if (ctx->dAvgCx) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->dAvgCx); ctx->dAvgCx = nullptr; }
            CK(ctx, cudaMalloc(&ctx->dAvgCx, sizeof(float) * (size_t)nStreams));
            ctx->avgCap = nStreams;
        }
        float *cxBuf = dComplexity; uint64_t cxStride = metaStride;
        if (!cxBuf) {
            const size_t need = (size_t)nStreams * nBlocksAll;
            if (ctx->cxScratchCap < need) {
                if (ctx->dCxScratch) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->dCxScratch); ctx->dCxScratch = nullptr; }
                CK(ctx, cudaMalloc(&ctx->dCxScratch, sizeof(float) * need));
                ctx->cxScratchCap = need;
            }
            cxBuf = ctx->dCxScratch; cxStride = nBlocksAll;
        }
        UlcbEncodeArgs p1 = a;
        p1.mode = ULCB_ENC_ANALYSE; p1.complexity = cxBuf; p1.metaStride = cxStride;
        p1.blockBytes = nullptr; p1.windowCtrl = nullptr; p1.totalBytes = nullptr; p1.status = nullptr;
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        if (ctx->timing) { t0 = timing_event(ctx); t1 = timing_event(ctx); cudaEventRecord(t0, ctx->stream); }
        CK(ctx, ulcb_encode_launch(p1, ctx->nChan, grid, ctx->stream));
        if (ctx->timing) { cudaEventRecord(t1, ctx->stream); ctx->spans.push_back({t0, t1, ULCB_KERNEL_ENC_ENCODE}); }
        CK(ctx, ulcb_avg_complexity_launch(cxBuf, cxStride, dNBlocks, nBlocksAll, nStreams, ctx->dAvgCx, ctx->stream));
        ctx->launches += 2; ctx->avgCount = nStreams;
        a.mode = ULCB_ENC_ABR; a.avgPerStream = ctx->dAvgCx;
    }
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (ctx->timing) { t0 = timing_event(ctx); t1 = timing_event(ctx); cudaEventRecord(t0, ctx->stream); }
    CK(ctx, ulcb_encode_launch(a, ctx->nChan, grid, ctx->stream));
    if (ctx->timing) { cudaEventRecord(t1, ctx->stream); ctx->spans.push_back({t0, t1, ULCB_KERNEL_ENC_ENCODE}); }
    ctx->launches++;
    return ULCB_OK;
}

int ULCB_GetAvgComplexity(ULCB_Context *ctx, float *avgComplexity, int nStreams) {
    if (!ctx || !avgComplexity || nStreams < 0 || nStreams > ctx->avgCount) return ULCB_E_PARAM;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpyAsync(avgComplexity, ctx->dAvgCx, sizeof(float) * (size_t)nStreams, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ULCB_OK;
}

int ULCB_EncodeStreamsDevice(ULCB_Context *ctx, int nStreams, const void *dPcm, int format, const uint64_t *dNSamples,
                             uint64_t inStrideElems, const uint32_t *dNBlocks, uint32_t nBlocksAll, const ULCB_EncodeParams *params,
                             uint8_t *dOut, uint64_t outStrideBytes, uint32_t *dBlockBytes, float *dComplexity, int32_t *dWindowCtrl,
                             uint64_t metaStride, uint64_t *dTotalBytes, int32_t *dStatus) {
    return encode_device(ctx, nStreams, dPcm, format, dNSamples, inStrideElems, dNBlocks, nBlocksAll, params, dOut, outStrideBytes,
                         dBlockBytes, dComplexity, dWindowCtrl, metaStride, dTotalBytes, dStatus, nullptr, nullptr, nullptr);
}

static int encode_host(ULCB_Context *ctx, int nStreams, const void *pcm, int format, const uint64_t *nSamples,
                       uint64_t inStrideElems, uint32_t nBlocksAll, const ULCB_EncodeParams *params,
                       uint8_t *out, uint64_t outStrideBytes, uint32_t *blockBytes, float *complexity, int32_t *windowCtrl,
                       uint64_t *totalBytes, int32_t *status, const EncTaps *hostTaps) {
    if (!ctx || nStreams < 0 || !pcm || !out || !params || nBlocksAll < 1) return ULCB_E_PARAM;
    if (format != ULCB_FMT_PCM16 && format != ULCB_FMT_FLOAT32) return ULCB_E_PARAM;
    if (nStreams == 0) return ULCB_OK;
    if (nSamples) for (int s = 0; s < nStreams; s++) if (nSamples[s] * (uint64_t)ctx->nChan > inStrideElems) return ULCB_E_PARAM;
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t elem = format == ULCB_FMT_PCM16 ? 2 : 4;
    const size_t inBytes = (size_t)nStreams * inStrideElems * elem, outBytes = (size_t)nStreams * outStrideBytes;
    const size_t nMeta = (size_t)nStreams * nBlocksAll;
    // aux layout: nSamples u64 | totalBytes u64 | blockBytes u32 | complexity f32 | windowCtrl i32 | status i32
    size_t o = 0;
    const size_t oNs = o; o += sizeof(uint64_t) * nStreams;
    const size_t oTot = o; o += sizeof(uint64_t) * nStreams;
    const size_t oBb = o; o += sizeof(uint32_t) * nMeta;
    const size_t oCx = o; o += sizeof(float) * nMeta;
    const size_t oWc = o; o += sizeof(int32_t) * nMeta;
    const size_t oSt = o; o += sizeof(int32_t) * nStreams;
    size_t oTap = (o + 255) & ~(size_t)255;
    const size_t per = (size_t)ctx->nChan * ctx->N;
    EncTaps dt{};
    size_t tapBytes = 0;
    if (hostTaps) tapBytes = nMeta * (sizeof(float) * (3 * per + ctx->N / 2) + sizeof(int32_t) * (per + 2));
    int rc;
    if ((rc = grow(ctx, &ctx->dIn, &ctx->dInCap, inBytes + 64))) return rc;
    if ((rc = grow(ctx, &ctx->dOut, &ctx->dOutCap, outBytes))) return rc;
    if ((rc = grow(ctx, &ctx->dAux, &ctx->dAuxCap, oTap + tapBytes))) return rc;
    char *aux = (char *)ctx->dAux;
    if (hostTaps) {
        char *p = aux + oTap;
        dt.coefs = (float *)p; p += sizeof(float) * nMeta * per;
        dt.keys = (float *)p;  p += sizeof(float) * nMeta * per;
        dt.noise = (float *)p; p += sizeof(float) * nMeta * per;
        dt.mask = (float *)p;  p += sizeof(float) * nMeta * (ctx->N / 2);
        dt.ranks = (int32_t *)p; p += sizeof(int32_t) * nMeta * per;
        dt.counts = (int32_t *)p;
    }
    CK(ctx, cudaMemcpyAsync(ctx->dIn, pcm, inBytes, cudaMemcpyHostToDevice, ctx->stream));
    if (nSamples) CK(ctx, cudaMemcpyAsync(aux + oNs, nSamples, sizeof(uint64_t) * nStreams, cudaMemcpyHostToDevice, ctx->stream));
    rc = encode_device(ctx, nStreams, ctx->dIn, format, nSamples ? (const uint64_t *)(aux + oNs) : nullptr, inStrideElems, nullptr, nBlocksAll,
                       params, (uint8_t *)ctx->dOut, outStrideBytes, (uint32_t *)(aux + oBb), (float *)(aux + oCx), (int32_t *)(aux + oWc),
                       nBlocksAll, (uint64_t *)(aux + oTot), (int32_t *)(aux + oSt), nullptr, nullptr, hostTaps ? &dt : nullptr);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(out, ctx->dOut, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (blockBytes) CK(ctx, cudaMemcpyAsync(blockBytes, aux + oBb, sizeof(uint32_t) * nMeta, cudaMemcpyDeviceToHost, ctx->stream));
    if (complexity) CK(ctx, cudaMemcpyAsync(complexity, aux + oCx, sizeof(float) * nMeta, cudaMemcpyDeviceToHost, ctx->stream));
    if (windowCtrl) CK(ctx, cudaMemcpyAsync(windowCtrl, aux + oWc, sizeof(int32_t) * nMeta, cudaMemcpyDeviceToHost, ctx->stream));
    if (totalBytes) CK(ctx, cudaMemcpyAsync(totalBytes, aux + oTot, sizeof(uint64_t) * nStreams, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) CK(ctx, cudaMemcpyAsync(status, aux + oSt, sizeof(int32_t) * nStreams, cudaMemcpyDeviceToHost, ctx->stream));
    if (hostTaps) {
        if (hostTaps->coefs) CK(ctx, cudaMemcpyAsync(hostTaps->coefs, dt.coefs, sizeof(float) * nMeta * per, cudaMemcpyDeviceToHost, ctx->stream));
        if (hostTaps->keys)  CK(ctx, cudaMemcpyAsync(hostTaps->keys, dt.keys, sizeof(float) * nMeta * per, cudaMemcpyDeviceToHost, ctx->stream));
        if (hostTaps->noise) CK(ctx, cudaMemcpyAsync(hostTaps->noise, dt.noise, sizeof(float) * nMeta * per, cudaMemcpyDeviceToHost, ctx->stream));
        if (hostTaps->mask)  CK(ctx, cudaMemcpyAsync(hostTaps->mask, dt.mask, sizeof(float) * nMeta * (ctx->N / 2), cudaMemcpyDeviceToHost, ctx->stream));
        if (hostTaps->ranks) CK(ctx, cudaMemcpyAsync(hostTaps->ranks, dt.ranks, sizeof(int32_t) * nMeta * per, cudaMemcpyDeviceToHost, ctx->stream));
        if (hostTaps->counts) CK(ctx, cudaMemcpyAsync(hostTaps->counts, dt.counts, sizeof(int32_t) * nMeta * 2, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ULCB_OK;
}

int ULCB_EncodeStreams(ULCB_Context *ctx, int nStreams, const void *pcm, int format, const uint64_t *nSamples,
                       uint64_t inStrideElems, uint32_t nBlocksAll, const ULCB_EncodeParams *params,
                       uint8_t *out, uint64_t outStrideBytes, uint32_t *blockBytes, float *complexity, int32_t *windowCtrl,
                       uint64_t *totalBytes, int32_t *status) {
    return encode_host(ctx, nStreams, pcm, format, nSamples, inStrideElems, nBlocksAll, params, out, outStrideBytes, blockBytes,
                       complexity, windowCtrl, totalBytes, status, nullptr);
}

int ULCB_TestEncode(ULCB_Context *ctx, int nStreams, const void *pcm, int format, const uint64_t *nSamples, uint64_t inStrideElems,
                    uint32_t nBlocksAll, const ULCB_EncodeParams *params, uint8_t *out, uint64_t outStrideBytes,
                    uint32_t *blockBytes, float *complexity, int32_t *windowCtrl,
                    float *coefs, float *keys, float *noise, float *mask, int32_t *ranks, int32_t *counts) {
    EncTaps t{coefs, keys, noise, mask, ranks, counts};
    return encode_host(ctx, nStreams, pcm, format, nSamples, inStrideElems, nBlocksAll, params, out, outStrideBytes, blockBytes,
                       complexity, windowCtrl, nullptr, nullptr, &t);
}

// ------------------------------------------------------------------------------------------------
// libulc drop-in: encoder (reference: include/ulcEncoder.h, libulc/ulcEncoder.c)
// ------------------------------------------------------------------------------------------------
struct DropinEncoder {
    ULCB_Context *ctx;
    UlcbEncStreamState *dState; float *dSamples;
    float *dSrc; uint8_t *dOut; uint32_t *dBytes; float *dCx; int32_t *dWc;
    float *hSrc; uint8_t *hOut; uint32_t *hMeta;       // pinned; hMeta = {bytes, cx bits, wc, nextWc}
    size_t outCap;
};

int ULC_EncoderState_Init(struct ULC_EncoderState_t *State) {
    State->BufferData = NULL;                                  // ulcEncoder.c:27
    State->SampleBuffer = State->TransformBuffer = State->TransformNoise = State->TransformFwdLap = State->TransformTemp = NULL;
    State->TransformIndex = NULL; State->TransientBuffer = NULL;
    const int nChan = State->nChan, N = State->BlockSize;
    if (!valid(nChan, N)) return -1;                           // ulcEncoder.c:29-34
    DropinEncoder *d = (DropinEncoder *)calloc(1, sizeof(DropinEncoder));
    if (!d) return -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (State->RateHz < 1 || ULCB_ContextCreate(&d->ctx, dev, State->RateHz, nChan, N, 1) != ULCB_OK) { free(d); return -1; }
    const size_t cn = (size_t)nChan * N;
    d->outCap = ULCB_MaxBlockBytes(nChan, N);
    bool ok = cudaMalloc(&d->dState, sizeof(UlcbEncStreamState)) == cudaSuccess &&
              cudaMalloc(&d->dSamples, sizeof(float) * 2 * cn) == cudaSuccess &&
              cudaMalloc(&d->dSrc, sizeof(float) * cn) == cudaSuccess &&
              cudaMalloc(&d->dOut, d->outCap) == cudaSuccess &&
              cudaMalloc(&d->dBytes, sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&d->dCx, sizeof(float)) == cudaSuccess &&
              cudaMalloc(&d->dWc, sizeof(int32_t)) == cudaSuccess &&
              cudaMallocHost(&d->hSrc, sizeof(float) * cn) == cudaSuccess &&
              cudaMallocHost(&d->hOut, d->outCap) == cudaSuccess &&
              cudaMallocHost(&d->hMeta, 4 * sizeof(uint32_t)) == cudaSuccess;
    if (ok) ok = cudaMemset(d->dState, 0, sizeof(UlcbEncStreamState)) == cudaSuccess;
    State->BufferData = d;
    if (!ok) { ULC_EncoderState_Destroy(State); return -1; }
    State->WindowCtrl = 0; State->NextWindowCtrl = 0x10;       // ulcEncoder.c:70
    State->BlockComplexity = 0.0f;
    State->TransientFilter[0] = State->TransientFilter[1] = State->TransientFilter[2] = 0.0f;
    State->TransformTemp = (float *)d->hOut;                   // the reference returns a pointer into TransformTemp
    return 1;
}

void ULC_EncoderState_Destroy(struct ULC_EncoderState_t *State) {
    DropinEncoder *d = (DropinEncoder *)State->BufferData;
    if (!d) return;
    cudaFree(d->dState); cudaFree(d->dSamples); cudaFree(d->dSrc); cudaFree(d->dOut); cudaFree(d->dBytes); cudaFree(d->dCx); cudaFree(d->dWc);
    cudaFreeHost(d->hSrc); cudaFreeHost(d->hOut); cudaFreeHost(d->hMeta);
    ULCB_ContextDestroy(d->ctx);
    free(d);
    State->BufferData = NULL;
}

static const void *dropin_encode(struct ULC_EncoderState_t *State, const float *SrcData, int *Size, const ULCB_EncodeParams *prm) {
    DropinEncoder *d = (DropinEncoder *)State->BufferData;
    if (Size) *Size = 0;
    if (!d) return NULL;
    ULCB_Context *ctx = d->ctx;
    const size_t cn = (size_t)State->nChan * State->BlockSize;
    memcpy(d->hSrc, SrcData, sizeof(float) * cn);
    if (cudaMemcpyAsync(d->dSrc, d->hSrc, sizeof(float) * cn, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return NULL;
    if (encode_device(ctx, 1, d->dSrc, ULCB_FMT_FLOAT32, nullptr, cn, nullptr, 1, prm, d->dOut, d->outCap, d->dBytes, d->dCx, d->dWc, 1,
                      nullptr, nullptr, d->dState, d->dSamples, nullptr) != ULCB_OK) return NULL;
    cudaMemcpyAsync(d->hOut, d->dOut, d->outCap, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&d->hMeta[0], d->dBytes, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&d->hMeta[1], d->dCx, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&d->hMeta[2], d->dWc, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&d->hMeta[3], &d->dState->nextWindowCtrl, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return NULL;
    memcpy(&State->BlockComplexity, &d->hMeta[1], sizeof(float));
    State->WindowCtrl = (int)d->hMeta[2];
    State->NextWindowCtrl = (int)d->hMeta[3];
    if (Size) *Size = (int)d->hMeta[0] * 8;                    // bits, already a multiple of 8 (ulcEncoder_Encode.c:357-359)
    return d->hOut;
}

const void *ULC_EncodeBlock_CBR(struct ULC_EncoderState_t *State, const float *SrcData, int *Size, float RateKbps) {
    ULCB_EncodeParams p = {ULCB_MODE_CBR, RateKbps, 0.0f};
    return dropin_encode(State, SrcData, Size, &p);
}
const void *ULC_EncodeBlock_ABR(struct ULC_EncoderState_t *State, const float *SrcData, int *Size, float RateKbps, float AvgComplexity) {
    ULCB_EncodeParams p = {ULCB_MODE_ABR, RateKbps, AvgComplexity};
    return dropin_encode(State, SrcData, Size, &p);
}
const void *ULC_EncodeBlock_VBR(struct ULC_EncoderState_t *State, const float *SrcData, int *Size, float Quality) {
    ULCB_EncodeParams p = {ULCB_MODE_VBR, Quality, 0.0f};
    return dropin_encode(State, SrcData, Size, &p);
}

}  // extern "C"
