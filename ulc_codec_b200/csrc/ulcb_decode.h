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
