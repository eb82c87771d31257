// ulcb_decode.h - internal interface between the C-ABI layer (ulcb_api.cu) and the decode kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ulcb_tables.h"

// Per-stream decoder state carried between launches (the fields ULC_DecoderState_t carries,
// include/ulcDecoder.h:27-31, plus the stream's position in the noise sequence).
struct UlcbStreamState {
    int      valid;              // 0 = fresh stream (lap zero, LastSubBlockSize 0, RNG at seed)
    int      lastSubBlockSize;
    uint32_t draws;              // noise draws consumed so far
    uint32_t pad;
};

struct UlcbDecodeArgs {
    UlcbTables      tb;
    int             nChan;             // channels per stream (the kernels specialised for 1 / 2 ignore it)
    int             nStreams;
    int             nBlocksAll;        // used when nBlocks == nullptr
    const uint8_t  *bytes;             // all streams, device
    const uint64_t *streamOffs;        // [nStreams+1] byte offsets into `bytes`
    const uint32_t *nBlocks;           // [nStreams] or nullptr
    void           *out;               // int16 or float, [stream][block][sample][chan]
    uint64_t        outStride;         // elements between streams
    int32_t        *status;            // [nStreams] -1 ok, else index of first undecodable block; nullable
    uint64_t       *bytesConsumed;     // [nStreams] nullable
    uint32_t       *blockBits;         // [nStreams][bitsStride] bits consumed per block (0 = not decoded); nullable
    uint64_t        bitsStride;
    UlcbStreamState *state;            // [nStreams] nullable: carry state across calls
    float          *statePend;         // [nStreams][nChan*N/2]
    float          *dbgCoefs;          // test hook: [nStreams][nBlocksAll][nChan*N] dequantised coefficients; nullable
};

size_t      ulcb_decode_smem_bytes(int N, int nChan);
int         ulcb_decode_threads(int N, int nChan);
cudaError_t ulcb_decode_launch(const UlcbDecodeArgs &args, int nChan, int pcm16, int gridBlocks, cudaStream_t stream);

// Warp-specialised fast path (ulcb_decode_ws.cu): BlockSize 2048, stereo, fresh streams.  ulcb_decode_launch
// routes to it when it applies; ULCB_DECODE_GENERIC=1 in the environment forces the generic kernel (A/B timing).
bool        ulcb_decode_ws_supported(const UlcbDecodeArgs &args, int nChan);
cudaError_t ulcb_decode_ws_launch(const UlcbDecodeArgs &args, int pcm16, int smCount, cudaStream_t stream);

// ---- block-parallel pipeline (ulcb_decode_bp.cu): BlockSize 2048, stereo, fresh streams ----
struct UlcbBpDesc {                   // one per block, written by the parse kernel
    uint32_t bits;                    // bits consumed as ULC_DecodeBlock returns them; 0 = not decodable
    uint32_t wcRuns;                  // window control | run records of channel 0 << 8 | of channel 1 << 20
    uint32_t draws;                   // noise draws of the block; after the scan: draws before the block in its stream
    uint16_t nCoef[2];                // coefficient words per channel
};

struct UlcbBpArgs {
    UlcbTables      tb;
    uint64_t        nSignBits;
    int             nChan;            // channels (the walk takes any count; parse and transform are built for stereo)
    int             s0, nStreams;     // this wave: streams [s0, s0 + nStreams) of the call
    int             nBlocksAll, maxBlk, runLen;
    const uint8_t  *bytes;            // all streams of the call
    const uint32_t *wend;             // first word that must not be read (end of bytes + slack, aligned down)
    const uint64_t *streamOffs;       // [call streams + 1]
    uint64_t        base0;            // streamOffs[s0]
    const uint32_t *nBlocks;          // [call streams] or nullptr
    const uint32_t *blockBytes;       // caller's block-size index [call streams][blockBytesStride], or nullptr (walk)
    uint64_t        blockBytesStride;
    void           *out; uint64_t outStride;
    int32_t        *status; uint32_t *blockBits; uint64_t bitsStride;
    float          *dbgCoefs;
    // wave scratch
    uint32_t       *blkOff;           // [nStreams][maxBlk + 1] byte offset of every block in its stream (first stream = s0)
    UlcbBpDesc     *desc;             // [nStreams][maxBlk]
    int32_t        *firstBad;         // [nStreams] first undecodable block (nBlocks when none)
    uint8_t        *arena;            // sparse block content
};

bool        ulcb_decode_bp_supported(const UlcbTables &tb, int nChan);
size_t      ulcb_decode_bp_arena_bytes(uint64_t payloadBytes, int nStreams, int maxBlk);
cudaError_t ulcb_decode_bp_offsets(const UlcbBpArgs &args, bool haveIndex, int smCount, cudaStream_t stream, int *launches);
// block sizes from the offsets table: sizes[stream][block] = bytes of the block, 0 from the first block that could not be located
cudaError_t ulcb_decode_bp_sizes(const UlcbBpArgs &args, uint32_t *dSizes, uint64_t sizesStride, int32_t *dStatus, cudaStream_t stream);
// ev (nullable): four events recorded before parse, after parse, after scan and after transform
cudaError_t ulcb_decode_bp_launch(const UlcbBpArgs &args, int pcm16, cudaStream_t stream, int *launches, cudaEvent_t *ev = nullptr, int phases = 3);
