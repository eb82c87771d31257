/* ulcb200.h - C ABI of the B200-native batched ulc-codec hot path (libulcb200.so).
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types cross this boundary.
 *
 * Two layers are exported:
 *
 *  1. The reference's own libulc API, same names, same struct layouts, same argument meaning and
 *     error behaviour, so the reference tools (tools/ulcEncodeTool.c, tools/ulcDecodeTool.c) link
 *     against libulcb200.so instead of libulc unchanged.  Each entry point cites the reference
 *     declaration it replaces.
 *
 *  2. The batched-stream variant the north star asks for (ULCB_*): thousands of independent
 *     streams per call, per-stream semantics identical to "fresh ULC_*State_Init followed by
 *     sequential block calls", including the decoder noise RNG restarting at 1234567 per stream.
 *
 * There is no CPU fallback: every entry point fails (negative status / -1) when no sm_100 device
 * or kernel image is available.
 */
#ifndef ULCB200_H
#define ULCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* 1. libulc drop-in (decoder half; encoder half is declared further down when built)          */
/* ------------------------------------------------------------------------------------------ */

/* replaces: struct ULC_DecoderState_t, /root/reference/include/ulcDecoder.h:13-32.
 * Field order and types are identical.  BufferData is owned by the library (device state +
 * pinned staging); the three float pointers are kept for layout compatibility and are NULL. */
struct ULC_DecoderState_t {
	int    nChan;
	int    BlockSize;
	int    LastSubBlockSize;
	void  *BufferData;
	float *TransformBuffer;
	float *TransformTemp;
	float *TransformInvLap;
};

/* replaces: ULC_DecoderState_Init, include/ulcDecoder.h:39, libulc/ulcDecoder.c:26-68.
 * Returns 1 on success, -1 on bad parameters (nChan 1..255, BlockSize a power of two in
 * 256..32768), unsupported configuration for this build (nChan > 2, BlockSize > 8192),
 * allocation failure, or no usable GPU. */
int ULC_DecoderState_Init(struct ULC_DecoderState_t *State);

/* replaces: ULC_DecoderState_Destroy, include/ulcDecoder.h:42, libulc/ulcDecoder.c:71-74.
 * Safe after a failed Init (BufferData is NULL then). */
void ULC_DecoderState_Destroy(struct ULC_DecoderState_t *State);

/* replaces: ULC_DecodeBlock, include/ulcDecoder.h:56, libulc/ulcDecoder.c:198-302.
 * DstData receives BlockSize*nChan interleaved floats; SrcBuffer is read bytewise and must have
 * ULCB_MaxBlockBytes(nChan, BlockSize) readable bytes (capped at 16 KiB - the reference tool's own
 * stream buffer guarantee, tools/ulcDecodeTool.c:79-81).  Returns bits consumed, 0 on a corrupt
 * block.  Unlike the reference, the noise RNG belongs to the state, not to the process. */
int ULC_DecodeBlock(struct ULC_DecoderState_t *State, float *DstData, const void *SrcBuffer);

/* ------------------------------------------------------------------------------------------ */
/* 2. Batched-stream API                                                                       */
/* ------------------------------------------------------------------------------------------ */

typedef struct ULCB_Context ULCB_Context;

enum {
	ULCB_OK            =  0,
	ULCB_E_PARAM       = -1,   /* bad argument */
	ULCB_E_UNSUPPORTED = -2,   /* valid for the reference, not built here yet (nChan > 2, BlockSize > 8192) */
	ULCB_E_NOMEM       = -3,
	ULCB_E_CUDA        = -4,   /* CUDA runtime error; ULCB_LastCudaError() has the code */
	ULCB_E_NODEVICE    = -5,   /* no sm_100 GPU / kernel image: there is no CPU fallback */
};

enum { ULCB_FMT_PCM16 = 0, ULCB_FMT_FLOAT32 = 1 };

/* Create a context bound to CUDA device `device` for streams of (rateHz, nChan, blockSize).
 * maxBlocksPerStream bounds the stream length the noise-sign table is prepared for. */
int  ULCB_ContextCreate(ULCB_Context **ctx, int device, int rateHz, int nChan, int blockSize,
                        uint32_t maxBlocksPerStream);
void ULCB_ContextDestroy(ULCB_Context *ctx);
int  ULCB_LastCudaError(const ULCB_Context *ctx);
const char *ULCB_ErrorString(int status);

/* Upper bound on the bytes one encoded block can occupy (4 nybbles per coefficient). */
size_t ULCB_MaxBlockBytes(int nChan, int blockSize);

/* Batched decode, DEVICE pointers, asynchronous on the context's stream (ULCB_Sync to wait).
 *   dBytes        all streams' raw block bytes (the .ulc payload after the 24-byte header)
 *   dStreamOffs   [nStreams+1] byte offsets of each stream in dBytes; stream s is
 *                 dBytes[dStreamOffs[s] .. dStreamOffs[s+1]).  dBytes needs 8 readable bytes of
 *                 slack after the last stream.
 *   dNBlocks      [nStreams] blocks to decode per stream, or NULL to decode nBlocksAll for all
 *   dOut          int16 (ULCB_FMT_PCM16) or float (ULCB_FMT_FLOAT32), interleaved
 *                 [stream][block][sample][chan]; outStrideElems elements between streams
 *   dStatus       [nStreams] -1 = all blocks decoded, else index of the first undecodable block
 *                 (that block and the rest of the stream are written as silence); may be NULL
 *   dBlockBits    [nStreams][bitsStride] bits consumed per block as ULC_DecodeBlock would return
 *                 them (0 = not decoded); may be NULL
 * No block index is needed: block boundaries are discovered by parsing, as in the reference. */
int ULCB_DecodeStreamsDevice(ULCB_Context *ctx, int nStreams,
                             const uint8_t *dBytes, const uint64_t *dStreamOffs,
                             const uint32_t *dNBlocks, uint32_t nBlocksAll,
                             void *dOut, uint64_t outStrideElems, int format,
                             int32_t *dStatus, uint32_t *dBlockBits, uint64_t bitsStride);

/* Same with HOST pointers: stages the bytes to the device, decodes, copies PCM and status back.
 * Synchronous.  bytes needs no slack. */
int ULCB_DecodeStreams(ULCB_Context *ctx, int nStreams,
                       const uint8_t *bytes, const uint64_t *streamOffs,
                       const uint32_t *nBlocks, uint32_t nBlocksAll,
                       void *out, uint64_t outStrideElems, int format,
                       int32_t *status, uint32_t *blockBits, uint64_t bitsStride);

int ULCB_Sync(ULCB_Context *ctx);

/* The CUDA stream (cudaStream_t, as an opaque pointer) the context launches on, so callers that
 * own device memory (e.g. through torch) can order their work against it. */
void *ULCB_Stream(ULCB_Context *ctx);

/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t ULCB_LaunchCount(const ULCB_Context *ctx);

/* Stage-level test hooks (host pointers; synchronous).  Used by tests/ only. */
int ULCB_TestDCT4(ULCB_Context *ctx, float *data, int size, int count);
/* parse + dequantise + noise-fill only: coefs[nBlocks][nChan*blockSize], blockBits[nBlocks] for ONE stream */
int ULCB_TestParse(ULCB_Context *ctx, const uint8_t *bytes, uint64_t nBytes, uint32_t nBlocks, float *coefs, uint32_t *blockBits);

#ifdef __cplusplus
}
#endif
#endif /* ULCB200_H */
