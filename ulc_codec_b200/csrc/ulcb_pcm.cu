// ulcb_pcm.cu - device-side sample-format conversion.
//
// The reference tools convert WAV sample formats to and from the float blocks libulc consumes and produces with the
// rules of tools/WavIO_Helper.c:31-87.  These kernels apply exactly those rules to whole device buffers so a caller
// can feed ULCB_EncodeStreamsDevice (float input) from, or turn ULCB_DecodeStreamsDevice's float output into, 8-bit
// unsigned, 16-bit, 24-bit packed or float32 PCM without leaving the GPU:
//   to float   : u8: ((int)(int8)b ^ 0x80) * 2^-7 (sic: the reference xors after the promotion, WavIO_Helper.c:35)
//                s16: x * 2^-15        s24: (x << 8 as int32) * 2^-31
//   from float : clamp(x * 2^(bits-1), -2^(bits-1), 2^(bits-1)-1), lrintf (round half even), u8 re-biased by 0x80
// Pure streaming work: 16 bytes per thread and access where the format's element size allows it, grid sized from
// the SM count; HBM-bound (bytes in + bytes out per sample).
#include <cuda_runtime.h>
#include <stdint.h>
#include "ulcb_pcm.h"

namespace {

__device__ __forceinline__ float s24_to_float(uint32_t b0, uint32_t b1, uint32_t b2) {
    const int32_t x = (int32_t)((b0 << 8) | (b1 << 16) | (b2 << 24));                 // WavIO_Helper.c:71-76
    return (float)x * 0x1.0p-31f;
}
__device__ __forceinline__ int round_clamped(float v, float scale, float lo, float hi) {
    return __float2int_rn(fminf(fmaxf(v * scale, lo), hi));                           // Clamp then lrintf
}

__global__ void to_float_kernel(float *__restrict__ dst, const uint8_t *__restrict__ src, size_t n, int fmt) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        float v[4];
        const int m = (int)(n - i < 4 ? n - i : 4);
        if (fmt == ULCB_PCM_U8) {
            uint32_t w = 0;
            if (m == 4 && ((uintptr_t)(src + i) & 3) == 0) w = *reinterpret_cast<const uint32_t *>(src + i);
            else for (int k = 0; k < m; k++) w |= (uint32_t)src[i + k] << (8 * k);
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = (float)((int)(int8_t)((w >> (8 * k)) & 0xff) ^ 0x80) * 0x1.0p-7f;   // int8 promoted, THEN xor: as the reference
        } else if (fmt == ULCB_PCM_S16) {
            const int16_t *s = reinterpret_cast<const int16_t *>(src) + i;
            if (m == 4 && ((uintptr_t)s & 7) == 0) {
                const uint2 w = *reinterpret_cast<const uint2 *>(s);
                v[0] = (float)(int16_t)(w.x & 0xffff) * 0x1.0p-15f; v[1] = (float)(int16_t)(w.x >> 16) * 0x1.0p-15f;
                v[2] = (float)(int16_t)(w.y & 0xffff) * 0x1.0p-15f; v[3] = (float)(int16_t)(w.y >> 16) * 0x1.0p-15f;
            } else for (int k = 0; k < m; k++) v[k] = (float)s[k] * 0x1.0p-15f;
        } else {                                                                       // ULCB_PCM_S24: 12 bytes per 4 samples
            const uint8_t *s = src + 3 * i;
            if (m == 4 && ((uintptr_t)s & 3) == 0) {
                const uint32_t a = reinterpret_cast<const uint32_t *>(s)[0], b = reinterpret_cast<const uint32_t *>(s)[1],
                               c = reinterpret_cast<const uint32_t *>(s)[2];
                v[0] = (float)(int32_t)(a << 8) * 0x1.0p-31f;
                v[1] = (float)(int32_t)(((a >> 24) | (b << 8)) << 8) * 0x1.0p-31f;
                v[2] = (float)(int32_t)(((b >> 16) | (c << 16)) << 8) * 0x1.0p-31f;
                v[3] = (float)(int32_t)(c & 0xffffff00u) * 0x1.0p-31f;
            } else for (int k = 0; k < m; k++) v[k] = s24_to_float(s[3 * k], s[3 * k + 1], s[3 * k + 2]);
        }
        if (m == 4 && ((uintptr_t)(dst + i) & 15) == 0) *reinterpret_cast<float4 *>(dst + i) = make_float4(v[0], v[1], v[2], v[3]);
        else for (int k = 0; k < m; k++) dst[i + k] = v[k];
    }
}

__global__ void from_float_kernel(uint8_t *__restrict__ dst, const float *__restrict__ src, size_t n, int fmt) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        const int m = (int)(n - i < 4 ? n - i : 4);
        if (m == 4 && ((uintptr_t)(src + i) & 15) == 0) { const float4 f = *reinterpret_cast<const float4 *>(src + i); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
        else for (int k = 0; k < m; k++) v[k] = src[i + k];
        if (fmt == ULCB_PCM_U8) {                                                       // WavIO_Helper.c:39-45
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) w |= (uint32_t)((uint8_t)(int8_t)round_clamped(v[k], 0x1.0p+7f, -128.0f, 127.0f) ^ 0x80u) << (8 * k);
            if (m == 4 && ((uintptr_t)(dst + i) & 3) == 0) *reinterpret_cast<uint32_t *>(dst + i) = w;
            else for (int k = 0; k < m; k++) dst[i + k] = (uint8_t)(w >> (8 * k));
        } else if (fmt == ULCB_PCM_S16) {                                               // WavIO_Helper.c:57-63
            uint32_t q[4];
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = (uint32_t)(uint16_t)(int16_t)round_clamped(v[k], 0x1.0p+15f, -32768.0f, 32767.0f);
            int16_t *d = reinterpret_cast<int16_t *>(dst) + i;
            if (m == 4 && ((uintptr_t)d & 7) == 0) *reinterpret_cast<uint2 *>(d) = make_uint2(q[0] | (q[1] << 16), q[2] | (q[3] << 16));
            else for (int k = 0; k < m; k++) d[k] = (int16_t)q[k];
        } else {                                                                        // ULCB_PCM_S24, WavIO_Helper.c:79-87
            uint32_t q[4];
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = (uint32_t)round_clamped(v[k], 0x1.0p+23f, -8388608.0f, 8388607.0f) & 0xffffffu;
            uint8_t *d = dst + 3 * i;
            if (m == 4 && ((uintptr_t)d & 3) == 0) {
                reinterpret_cast<uint32_t *>(d)[0] = q[0] | (q[1] << 24);
                reinterpret_cast<uint32_t *>(d)[1] = (q[1] >> 8) | (q[2] << 16);
                reinterpret_cast<uint32_t *>(d)[2] = (q[2] >> 16) | (q[3] << 8);
            } else for (int k = 0; k < m; k++) { d[3 * k] = (uint8_t)q[k]; d[3 * k + 1] = (uint8_t)(q[k] >> 8); d[3 * k + 2] = (uint8_t)(q[k] >> 16); }
        }
    }
}

int grid_for(size_t n, int smCount) {
    const size_t want = (n / 4 + 255) / 256;
    const size_t cap = (size_t)smCount * 8;                                            // 8 resident CTAs of 256 threads per SM
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

cudaError_t ulcb_pcm_to_float(float *dDst, const void *dSrc, size_t n, int fmt, int smCount, cudaStream_t stream) {
    if (fmt == ULCB_PCM_F32) return cudaMemcpyAsync(dDst, dSrc, n * sizeof(float), cudaMemcpyDeviceToDevice, stream);
    to_float_kernel<<<grid_for(n, smCount), 256, 0, stream>>>(dDst, static_cast<const uint8_t *>(dSrc), n, fmt);
    return cudaGetLastError();
}

cudaError_t ulcb_pcm_from_float(void *dDst, const float *dSrc, size_t n, int fmt, int smCount, cudaStream_t stream) {
    if (fmt == ULCB_PCM_F32) return cudaMemcpyAsync(dDst, dSrc, n * sizeof(float), cudaMemcpyDeviceToDevice, stream);
    from_float_kernel<<<grid_for(n, smCount), 256, 0, stream>>>(static_cast<uint8_t *>(dDst), dSrc, n, fmt);
    return cudaGetLastError();
}
