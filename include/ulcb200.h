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
/* 1. libulc drop-in                                                                           */
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
 * 256..32768), unsupported configuration for this build (see ULCB_E_UNSUPPORTED),
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

/* replaces: struct ULC_TransientData_t / struct ULC_EncoderState_t, /root/reference/include/ulcEncoder.h:44-78.
 * Field order and types are identical (default feature flags: ULC_USE_NOISE_CODING = 1).  BufferData
 * is the library's handle (device state + pinned staging); TransformTemp points at the pinned host
 * copy of the last encoded block (the reference returns a pointer into TransformTemp too); the other
 * buffer pointers are NULL because those buffers live on the GPU.  WindowCtrl, NextWindowCtrl and
 * BlockComplexity are refreshed after every ULC_EncodeBlock_* call (tools/ulcEncodeTool.c:164 reads
 * BlockComplexity). */
struct ULC_TransientData_t { float Sum, SumW; };
struct ULC_EncoderState_t {
	int    RateHz;
	int    nChan;
	int    BlockSize;
	int    WindowCtrl;
	int    NextWindowCtrl;
	float  BlockComplexity;
	float  TransientFilter[3];
	void  *BufferData;
	float *SampleBuffer;
	float *TransformBuffer;
	float *TransformNoise;
	float *TransformFwdLap;
	float *TransformTemp;
	int   *TransformIndex;
	struct ULC_TransientData_t *TransientBuffer;
};

/* replaces: ULC_EncoderState_Init, include/ulcEncoder.h:85, libulc/ulcEncoder.c:25-80.
 * Returns 1 on success, -1 on bad parameters, unsupported configuration, allocation failure or no GPU. */
int ULC_EncoderState_Init(struct ULC_EncoderState_t *State);

/* replaces: ULC_EncoderState_Destroy, include/ulcEncoder.h:88, libulc/ulcEncoder.c:83-88. */
void ULC_EncoderState_Destroy(struct ULC_EncoderState_t *State);

/* replace: ULC_EncodeBlock_CBR/_ABR/_VBR, include/ulcEncoder.h:135-137, libulc/ulcEncoder.c:117-158.
 * SrcData = BlockSize*nChan interleaved floats (libulc/ulcEncoder_BlockTransform.c:96-98).  Returns the
 * encoded bytes of the PREVIOUS block (valid until the next call on this state) and its size in bits
 * (a multiple of 8) in *Size (may be NULL); NULL on a CUDA failure. */
const void *ULC_EncodeBlock_CBR(struct ULC_EncoderState_t *State, const float *SrcData, int *Size, float RateKbps);
const void *ULC_EncodeBlock_ABR(struct ULC_EncoderState_t *State, const float *SrcData, int *Size, float RateKbps, float AvgComplexity);
const void *ULC_EncodeBlock_VBR(struct ULC_EncoderState_t *State, const float *SrcData, int *Size, float Quality);

/* ------------------------------------------------------------------------------------------ */
/* 2. Batched-stream API                                                                       */
/* ------------------------------------------------------------------------------------------ */

typedef struct ULCB_Context ULCB_Context;

enum {
	ULCB_OK            =  0,
	ULCB_E_PARAM       = -1,   /* bad argument */
	ULCB_E_UNSUPPORTED = -2,   /* valid for the reference, not built here: BlockSize 32768, nChan > 8, or nChan * BlockSize > 16384 beyond mono / stereo at BlockSize <= 8192 */
	ULCB_E_NOMEM       = -3,
	ULCB_E_CUDA        = -4,   /* CUDA runtime error; ULCB_LastCudaError() has the code */
	ULCB_E_NODEVICE    = -5,   /* no sm_100 GPU / kernel image: there is no CPU fallback */
};

enum { ULCB_FMT_PCM16 = 0, ULCB_FMT_FLOAT32 = 1,
       ULCB_FMT_PCM8U = 2, ULCB_FMT_PCM24 = 3 /* these two: ULCB_ConvertSamplesDevice only */ };

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
 * No block index is needed: block boundaries are discovered by parsing, as in the reference.
 * (The call reads dStreamOffs / dNBlocks back once to size its scratch, so it synchronises the context's stream briefly
 * before it launches.) */
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

/* Batched decode WITH a block-size index: dBlockBytes[nStreams][blockBytesStride] holds the byte size of every block of
 * every stream - exactly what ULCB_EncodeStreams[Device] returns in dBlockBytes, or what a container reader recorded.  The
 * .ulc payload itself carries no such index (tools/ulc_Helper.h:10-20): the reference finds block k+1 by decoding block k
 * (tools/ulcDecodeTool.c:153-164), and so do the entry points above (a token-length walk per stream on the device).  With
 * the index every (stream, block) is decoded independently from the first instruction on.  An index that disagrees with
 * the bitstream marks the stream undecodable from the first block whose parsed length differs (dStatus).  All other
 * arguments as ULCB_DecodeStreamsDevice / ULCB_DecodeStreams. */
int ULCB_DecodeStreamsIndexedDevice(ULCB_Context *ctx, int nStreams,
                                    const uint8_t *dBytes, const uint64_t *dStreamOffs,
                                    const uint32_t *dNBlocks, uint32_t nBlocksAll,
                                    const uint32_t *dBlockBytes, uint64_t blockBytesStride,
                                    void *dOut, uint64_t outStrideElems, int format,
                                    int32_t *dStatus, uint32_t *dBlockBits, uint64_t bitsStride);
int ULCB_DecodeStreamsIndexed(ULCB_Context *ctx, int nStreams,
                              const uint8_t *bytes, const uint64_t *streamOffs,
                              const uint32_t *nBlocks, uint32_t nBlocksAll,
                              const uint32_t *blockBytes, uint64_t blockBytesStride,
                              void *out, uint64_t outStrideElems, int format,
                              int32_t *status, uint32_t *blockBits, uint64_t bitsStride);

int ULCB_Sync(ULCB_Context *ctx);

/* The block index a .ulc file does not carry (tools/ulc_Helper.h:10-20): the size in bytes of every block of every stream, found
 * by walking the token lengths only - no dequantisation, no transform.  The reference has no counterpart short of decoding the
 * file (tools/ulcDecodeTool.c:153-164 learns each block's size from ULC_DecodeBlock's return value).  HOST pointers, synchronous;
 * any block size and channel count the context supports.
 *   blockBytes    [nStreams][blockBytesStride] out: bytes of each block; 0 from the first block that cannot be located
 *                 (corrupt or truncated stream) on
 *   status        [nStreams] out, nullable: -1 = every block located, else the index of the first that was not */
int ULCB_IndexStreams(ULCB_Context *ctx, int nStreams, const uint8_t *bytes, const uint64_t *streamOffs, const uint32_t *nBlocks,
                      uint32_t nBlocksAll, uint32_t *blockBytes, uint64_t blockBytesStride, int32_t *status);

/* ---- batched encode ---- */
enum { ULCB_MODE_VBR = 0, ULCB_MODE_CBR = 1, ULCB_MODE_ABR = 2,
       /* ABR whose AvgComplexity the library measures itself, per stream, in a first analysis pass on the device (window
        * control + MDCT + BlockComplexity only) - what a user of tools/ulcEncodeTool.c does by hand: encode once, read
        * "Avg complexity = %.5f" (:178-188), encode again with RateKbps,AvgComplexity (:41).  avgComplexity is ignored. */
       ULCB_MODE_ABR_2PASS = 3 };
typedef struct ULCB_EncodeParams {
	int   mode;            /* ULCB_MODE_* */
	float rate;            /* VBR: Quality (0,100]; CBR/ABR: RateKbps */
	float avgComplexity;   /* ABR only (tools/ulcEncodeTool.c:42,158) */
} ULCB_EncodeParams;

/* The per-stream AvgComplexity values the last ULCB_MODE_ABR_2PASS call measured (synchronises the context's stream). */
int ULCB_GetAvgComplexity(ULCB_Context *ctx, float *avgComplexity, int nStreams);

/* Blocks the reference tool writes for a stream of nSamples sample points: ceil(n/BlockSize)+2
 * (coding delay + MDCT delay, tools/ulcEncodeTool.c:98). */
uint32_t ULCB_BlocksForSamples(uint64_t nSamples, int blockSize);

/* Batched encode, DEVICE pointers, asynchronous on the context's stream.
 *   dPcm          interleaved int16 (ULCB_FMT_PCM16, scaled by 2^-15 as tools/WavIO_Helper.c:49-55)
 *                 or float samples, [stream][sample][chan]; inStrideElems elements between streams
 *   dNSamples     [nStreams] valid sample points per stream (input is zero padded beyond, as
 *                 tools/WavIO_Reader.c:142-145), or NULL = nBlocks*BlockSize
 *   dNBlocks      [nStreams] blocks to produce per stream, or NULL to produce nBlocksAll for all.
 *                 Block b of the output describes input block b-1 (libulc's one-block delay).
 *   dOut          per stream an arena of outStrideBytes bytes; the stream's blocks are written back
 *                 to back (exactly the .ulc payload, tools/ulcEncodeTool.c:162-168)
 *   dBlockBytes   [nStreams][metaStride] bytes of each block; dComplexity / dWindowCtrl likewise
 *                 (State->BlockComplexity / State->WindowCtrl after each call); each may be NULL
 *   dTotalBytes   [nStreams] payload bytes; dStatus [nStreams] 0 = ok, 1 = arena too small (stream
 *                 truncated at the first block that did not fit); each may be NULL */
int ULCB_EncodeStreamsDevice(ULCB_Context *ctx, int nStreams, const void *dPcm, int format, const uint64_t *dNSamples,
                             uint64_t inStrideElems, const uint32_t *dNBlocks, uint32_t nBlocksAll, const ULCB_EncodeParams *params,
                             uint8_t *dOut, uint64_t outStrideBytes, uint32_t *dBlockBytes, float *dComplexity, int32_t *dWindowCtrl,
                             uint64_t metaStride, uint64_t *dTotalBytes, int32_t *dStatus);

/* Same with HOST pointers (metaStride = nBlocksAll).  Synchronous. */
int ULCB_EncodeStreams(ULCB_Context *ctx, int nStreams, const void *pcm, int format, const uint64_t *nSamples,
                       uint64_t inStrideElems, uint32_t nBlocksAll, const ULCB_EncodeParams *params,
                       uint8_t *out, uint64_t outStrideBytes, uint32_t *blockBytes, float *complexity, int32_t *windowCtrl,
                       uint64_t *totalBytes, int32_t *status);

/* Measurement input: the integer-only synthetic PCM16 corpus of SURVEY.md 8d, generated on the device.  Stream
 * firstStreamId+i, samples t0 .. t0+nSamples-1, stereo interleaved, written at dOut + i*strideElems.  kind 0 =
 * music-like, 1 = transient-heavy.  The two look-up tables (host pointers) are the ones ulc_codec_b200/corpus.py
 * builds; the output is bit-identical to corpus.generate().  Asynchronous on the context's stream. */
int ULCB_GenerateCorpusDevice(ULCB_Context *ctx, int kind, uint32_t firstStreamId, int nStreams, uint64_t nSamples, uint64_t t0,
                              const int16_t *sineLut1024, const int16_t *burstLut882, int16_t *dOut, uint64_t strideElems);

/* Device-side sample I/O.  replaces: WAV_ConvertToFloat_{PCM8u,PCM16,PCM24} / WAV_ConvertFromFloat_{...},
 * /root/reference/tools/WavIO_Helper.c:31-87 (what tools/ulcEncodeTool.c:153 and tools/ulcDecodeTool.c:166-171 call
 * around every block).  Converts nSamples interleaved samples between `format` (ULCB_FMT_PCM8U: unsigned 8-bit,
 * ULCB_FMT_PCM16, ULCB_FMT_PCM24: packed little-endian 3-byte, ULCB_FMT_FLOAT32) and float32 in +-1.0 with the
 * reference's scale / clamp / round-half-even rules.  toFloat != 0: dSrc holds `format`, dDst float32; toFloat == 0:
 * dSrc holds float32, dDst `format`.  Device pointers; asynchronous on the context's stream. */
int ULCB_ConvertSamplesDevice(ULCB_Context *ctx, int format, int toFloat, const void *dSrc, void *dDst, uint64_t nSamples);

/* The CUDA stream (cudaStream_t, as an opaque pointer) the context launches on, so callers that
 * own device memory (e.g. through torch) can order their work against it. */
void *ULCB_Stream(ULCB_Context *ctx);

/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t ULCB_LaunchCount(const ULCB_Context *ctx);

/* Per-kernel device time of the block-parallel decode pipeline (measurement aid for bench.py's roofline object; the reference has
 * no counterpart).  While enabled, CUDA events are recorded on the context's stream around every launch; ULCB_KernelTimes
 * synchronises the stream and returns, per kernel kind, the accumulated milliseconds and launch counts since the last reset. */
#define ULCB_KERNEL_BP_WALK       0   /* token-length walk of streams without a block-size index */
#define ULCB_KERNEL_BP_INDEX      1   /* block offsets from a caller-supplied index */
#define ULCB_KERNEL_BP_PARSE      2
#define ULCB_KERNEL_BP_SCAN       3
#define ULCB_KERNEL_BP_TRANSFORM  4
#define ULCB_KERNEL_ENC_DETECT    5   /* encoder: transient-detector pre-pass */
#define ULCB_KERNEL_ENC_ENCODE    6   /* encoder: the encode kernel (each pass of a two-pass ABR call) */
#define ULCB_KERNEL_KINDS         7
int ULCB_SetKernelTiming(ULCB_Context *ctx, int enable);
int ULCB_KernelTimes(ULCB_Context *ctx, double *ms, uint64_t *count, int reset);

/* Chunked pipeline of the block-parallel decode (BlockSize 2048 stereo batches): the streams of a call are processed in `chunks`
 * contiguous chunks; overlap 1 = upload (host-pointer entry points) and token-length walk of chunk c+1 on side CUDA streams beside
 * the transform of chunk c, 2 = its parse too (two arena halves), 0 = no pipelining (the whole call at once, `chunks` ignored).
 * chunks = 0 (the default) lets the library choose: pipelined for the host-pointer entry points, whose transform waits for PCIe, not
 * for device-resident input.  The decoded output does not depend on these settings.  Environment: ULCB_BP_CHUNKS / ULCB_BP_OVERLAP. */
int ULCB_SetDecodePipeline(ULCB_Context *ctx, int chunks, int overlap);
int ULCB_GetDecodePipeline(const ULCB_Context *ctx, int *chunks, int *overlap);

/* Stage-level test hooks (host pointers; synchronous).  Used by tests/ only. */
int ULCB_TestDCT4(ULCB_Context *ctx, float *data, int size, int count);
/* parse + dequantise + noise-fill only: coefs[nBlocks][nChan*blockSize], blockBits[nBlocks] for ONE stream */
int ULCB_TestParse(ULCB_Context *ctx, const uint8_t *bytes, uint64_t nBytes, uint32_t nBlocks, float *coefs, uint32_t *blockBits);

/* encode with per-block taps: coefs/keys/noise [nStreams][nBlocksAll][nChan*blockSize], mask [..][blockSize/2],
 * ranks likewise (int32), counts [..][2] = {codeable coefficients, coefficients requested}; any tap may be NULL */
int ULCB_TestEncode(ULCB_Context *ctx, int nStreams, const void *pcm, int format, const uint64_t *nSamples, uint64_t inStrideElems,
                    uint32_t nBlocksAll, const ULCB_EncodeParams *params, uint8_t *out, uint64_t outStrideBytes,
                    uint32_t *blockBytes, float *complexity, int32_t *windowCtrl,
                    float *coefs, float *keys, float *noise, float *mask, int32_t *ranks, int32_t *counts);

#ifdef __cplusplus
}
#endif
#endif /* ULCB200_H */
