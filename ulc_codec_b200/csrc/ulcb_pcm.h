// ulcb_pcm.h - internal interface of the sample-format conversion kernels (ulcb_pcm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

enum { ULCB_PCM_S16 = 0, ULCB_PCM_F32 = 1, ULCB_PCM_U8 = 2, ULCB_PCM_S24 = 3 };   // = ULCB_FMT_* of include/ulcb200.h

cudaError_t ulcb_pcm_to_float(float *dDst, const void *dSrc, size_t n, int fmt, int smCount, cudaStream_t stream);
cudaError_t ulcb_pcm_from_float(void *dDst, const float *dSrc, size_t n, int fmt, int smCount, cudaStream_t stream);
