// ulcb_encode.h - internal interface between the C-ABI layer (ulcb_api.cu) and the encode kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ulcb_tables.h"

enum { ULCB_NBARK = 25 };

// Host-built tables for the encoder (depend on BlockSize and RateHz only).  Everything that the
// reference computes with sinhf/asinhf/floorf/ceilf/expf (libulc/ulcHelper.h:96-120,
// libulc/ulcEncoder_Psyopt.c:60-250, libulc/ulcEncoder_WindowControl.c:73-129) is evaluated once on
// the host with the host libm so that the constants are the reference's own.
struct UlcbEncTables {
    int N, log2N, rateHz;
    // rotor table shared with the decoder: entry [span/2 + i] = (cos,sin)((i+1/2)/span * pi/2), i < span/2,
    // for every power-of-two span <= N, from the reference's f32 recurrence
    const float2 *rot;
    // transient detector one-pole coefficients, already as (1 - k)
    float c1PostHp, c1PostBp, c1PreHp, c1PreBp, c1Blk;
    // Bark band edges for sub-block shift sh (lines n = (N >> sh) / 2):
    //   edges[((sh*2 + kind)*NBARK + b)*2 + {0,1}] = {beg, end};  kind 0 = noise spectrum [b, b+2),
    //   kind 1 = masking [b-0.75, b+0.25)
    const int16_t *edges;
    // the same edges as a walk list: events[((sh*2 + kind)*2*NBARK + e)*2 + {0,1}] = {line position, band | 0x80 if end},
    // sorted by position (begins before ends at equal positions, then by band)
    const int16_t *events;
    // per-line Bark interpolation for shift sh, line i: index lineOff[sh] + i -> band index (int)bark, frac
    const uint8_t *lineBand;
    const float   *lineFrac;
    int lineOff[4];
};

// Per-stream encoder state carried between launches (what ULC_EncoderState_t carries,
// include/ulcEncoder.h:64-77, in the closed-form layout of this implementation).
struct UlcbEncStreamState {
    int   valid;              // 0 = fresh stream
    int   nextWindowCtrl;     // window control of the newest block held (NextWindowCtrl)
    int   windowCtrl;         // window control used by the last coded block (WindowCtrl)
    int   prevWindowCtrl;     // the one before
    float env[3];             // TransientFilter[3]
    float seg[16];            // TransientBuffer[8..15] {Sum, SumW}
    float complexity;
    uint32_t calls;           // blocks consumed so far
};

enum { ULCB_ENC_VBR = 0, ULCB_ENC_CBR = 1, ULCB_ENC_ABR = 2,
       ULCB_ENC_ANALYSE = 3 };   // window control + MDCT + BlockComplexity only: the first pass of two-pass ABR, writes no bytes

struct UlcbEncodeArgs {
    UlcbEncTables   tb;
    int             nChan;             // channels per stream (the kernels specialised for 1 / 2 ignore it)
    int             nStreams;
    int             nBlocksAll;        // blocks per stream when nBlocks == nullptr
    const uint32_t *nBlocks;           // [nStreams] or nullptr
    const void     *pcm;               // interleaved int16 or float, [stream][sample][chan]
    int             pcmFloat;          // 0 = int16, 1 = float32
    uint64_t        inStride;          // elements between streams
    const uint64_t *nSamples;          // [nStreams] sample points per stream (zero padded beyond), or nullptr = nBlocks*N
    int             mode;              // ULCB_ENC_*
    float           param;             // VBR: target = 0x1.E4EFB7p3f*logf(100/Quality) (host libm); CBR/ABR: RateKbps
    float           avgComplexity;     // ABR
    const float    *avgPerStream;      // ABR: [nStreams] per-stream AvgComplexity (device), overrides avgComplexity; nullable
    uint8_t        *out;               // [stream] arena of outStride bytes, blocks back to back
    uint64_t        outStride;
    uint32_t       *blockBytes;        // [stream][metaStride] bytes per block; nullable
    float          *complexity;        // [stream][metaStride]; nullable
    int32_t        *windowCtrl;        // [stream][metaStride]; nullable
    uint64_t        metaStride;
    uint64_t       *totalBytes;        // [nStreams]; nullable
    int32_t        *status;            // [nStreams] 0 ok, 1 = arena overflow; nullable
    UlcbEncStreamState *state;         // [nStreams] nullable: carry state across calls
    float          *stateSamples;      // [nStreams][2*nChan*N] blocks k-2, k-1 (M/S planar)
    float          *workspace;         // global workspace for large blocks: [gridDim.x][wsFloats]; nullptr = shared memory
    uint64_t        wsFloats;
    const int32_t  *wcAll;             // [stream][wcStride] transient-detector decisions from ulcb_detect_launch; nullptr = the
    uint64_t        wcStride;          //   encode kernel runs the detector itself (one warp beside the workers)
    int             encThreads;        // with wcAll: 256 / 512 worker threads per CTA, 0 = by block size
    int             exactBands;        // 1 = the Bark analysis adds its lines one by one like the reference (LineSum_t); 0 = segmented sums
    // test taps (nullable), [stream][nBlocksAll][nChan*N] / [..][N/2]
    float          *dbgCoefs;
    float          *dbgKeys;
    float          *dbgNoise;
    float          *dbgMask;
    int32_t        *dbgRanks;
    int32_t        *dbgCounts;         // [stream][nBlocksAll][2] = {nNz, nOut}
    int             dbgFlags;          // ULCB_PHASE_TIMING builds only: skip phases (timing experiments)
};

size_t      ulcb_encode_smem_bytes(int N, int nChan, int nthr, bool globalWs, bool pre = false);
// Detector pre-pass (ulcEncoder_WindowControl.c:31-239 for every block of every stream, 32 streams per CTA, one lane per stream,
// waves of one CTA per SM):
// shapes it takes, its scratch in bytes, and the launch; decisions go to wcAll[stream][wcStride].  The encode launch that
// follows with args.wcAll set needs no detector warp.
bool        ulcb_detect_supported(int N, int nChan);
bool        ulcb_encode_pre_uses_global_ws(int N, int nChan);
size_t      ulcb_detect_scratch_bytes(int N, int nStreams, int smCount);
cudaError_t ulcb_detect_launch(const UlcbEncodeArgs &args, int nChan, int smCount, void *scratch, int32_t *wcAll, uint64_t wcStride, cudaStream_t stream,
                               int *launches);
size_t      ulcb_encode_ws_floats(int N, int nChan);
int         ulcb_encode_threads(int N, int nChan);
bool        ulcb_encode_uses_global_ws(int N, int nChan);
// per-stream mean of the block complexities, rounded like the tool's "%.5f" print / "%f" read-back (tools/ulcEncodeTool.c:41,178-188)
cudaError_t ulcb_avg_complexity_launch(const float *dComplexity, uint64_t metaStride, const uint32_t *dNBlocks, uint32_t nBlocksAll,
                                       int nStreams, float *dAvg, cudaStream_t stream);
cudaError_t ulcb_encode_launch(const UlcbEncodeArgs &args, int nChan, int gridBlocks, cudaStream_t stream);
