// ulcb_corpus.cu - integer-only synthetic PCM16 corpus on the device (SURVEY.md 8d "Synthetic inputs").
//
// Bit-identical to the numpy generator in ulc_codec_b200/corpus.py: every sample is a pure function of
// (kind, stream id, sample index, channel) built from 32/64-bit integer arithmetic and two int16 look-up
// tables that the host passes in as data.  It exists so that the large configurations (4096 x 60 s,
// 65536 x 180 s in resident waves) can be synthesised where they are consumed instead of crossing PCIe,
// while the CPU reference is fed the very same samples from the numpy twin.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {          // murmur3 fmix32
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ long long floordiv(long long a, long long b) {   // b > 0, Python semantics
    long long q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}
__device__ __forceinline__ int noise16(uint32_t seed, uint32_t t, int ch) {
    uint32_t ctr = (t * 2u + (uint32_t)ch) * 0x9E3779B1u;
    return (int)(mix32(ctr ^ seed) & 0xFFFFu) - 32768;
}

struct StreamParams { uint32_t inc[2][8], ph0[2][8], amPhase, impOff, burOff, seed; };

__global__ void corpus_kernel(int16_t *out, uint64_t strideElems, int kind, uint32_t firstId, uint64_t nSamples, uint64_t t0,
                              const int16_t *__restrict__ sineLut, const int16_t *__restrict__ burstLut) {
    __shared__ StreamParams sp;
    __shared__ int16_t lut[1024];
    const uint32_t id = firstId + blockIdx.y;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) lut[i] = sineLut[i];
    if (threadIdx.x == 0) {
        const uint32_t seed = (0x9E3779B9u * (id + 1u)) ^ (0x51ED270Bu * (uint32_t)(kind + 1));
        sp.seed = seed;
        for (int p = 0; p < 8; p++) {
            const uint32_t h = mix32(seed ^ (0x1000193u * (uint32_t)(p + 1)));
            const long long base = (long long)(((unsigned long long)(55u << p) << 32) / 44100ull);
            const long long detune = (long long)((h >> 8) % 2001u) - 1000;
            const long long left = base + floordiv(base * detune, 100000);
            const long long right = left + floordiv(left * 3, 1000);
            sp.inc[0][p] = (uint32_t)left; sp.inc[1][p] = (uint32_t)right;
            sp.ph0[0][p] = h * 0x9E3779B1u; sp.ph0[1][p] = h * 0x85EBCA77u;
        }
        sp.amPhase = mix32(seed ^ 0xA5A5A5A5u) % 88200u;
        sp.impOff = mix32(seed ^ 0x0BADC0DEu) % 6300u;
        sp.burOff = mix32(seed ^ 0x00C0FFEEu) % 14700u;
    }
    __syncthreads();
    const int amp[8] = {6000, 4200, 3000, 2100, 1500, 1000, 700, 500};
    int16_t *dst = out + (uint64_t)blockIdx.y * strideElems;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nSamples; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = t0 + i;
        int v[2];
        if (kind == 0) {
            const long long tri = llabs((long long)((t + sp.amPhase) % 88200ull) - 44100);
            const long long gain = 16384 + (16384 * tri) / 44100;
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                long long acc = 0;
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    const uint32_t phase = sp.ph0[ch][p] + sp.inc[ch][p] * (uint32_t)t;     // mod 2^32
                    acc += (long long)amp[p] * lut[phase >> 22];
                }
                long long tone = acc >> 15;
                tone = (tone * gain) >> 15;
                const long long nz = ((long long)noise16(sp.seed, (uint32_t)t, ch) * 2844) >> 15;
                long long s = tone + nz;
                v[ch] = (int)(s < -32768 ? -32768 : (s > 32767 ? 32767 : s));
            }
        } else {
            const uint32_t ti = (uint32_t)((t + sp.impOff) % 6300ull), tb = (uint32_t)((t + sp.burOff) % 14700ull);
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                const long long n16 = noise16(sp.seed, (uint32_t)t, ch);
                long long s = (n16 * 57) >> 15;                                             // -60 dBFS floor
                if (ti < 4) s += (ti & 1) ? -29491 : 29491;
                if (tb < 882) s += (((n16 * (long long)burstLut[tb]) >> 15) * 16384) >> 15;
                v[ch] = (int)(s < -32768 ? -32768 : (s > 32767 ? 32767 : s));
            }
        }
        reinterpret_cast<uint32_t *>(dst)[i] = (uint32_t)(uint16_t)(int16_t)v[0] | ((uint32_t)(uint16_t)(int16_t)v[1] << 16);
    }
}

}  // namespace

cudaError_t ulcb_corpus_launch(int16_t *dOut, uint64_t strideElems, int kind, uint32_t firstId, int nStreams, uint64_t nSamples,
                               uint64_t t0, const int16_t *dSine, const int16_t *dBurst, cudaStream_t stream) {
    if (nStreams <= 0 || nSamples == 0) return cudaSuccess;
    unsigned gx = (unsigned)((nSamples + 256 * 8 - 1) / (256 * 8));
    if (gx > 1024) gx = 1024;
    for (int s = 0; s < nStreams; s += 65535) {
        int ns = nStreams - s < 65535 ? nStreams - s : 65535;
        corpus_kernel<<<dim3(gx, (unsigned)ns), 256, 0, stream>>>(dOut + (uint64_t)s * strideElems, strideElems, kind, firstId + (uint32_t)s,
                                                                  nSamples, t0, dSine, dBurst);
    }
    return cudaGetLastError();
}
