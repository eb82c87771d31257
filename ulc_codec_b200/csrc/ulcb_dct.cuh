// ulcb_dct.cuh - block-cooperative, shared-memory DCT-IV for the decoder.
//
// Computes the unnormalised DCT-IV the reference defines in
// libfourier/source/Fourier_DCT4_Template.h:6-116 (X[k] = sum v[n] cos((n+1/2)(k+1/2)pi/S)),
// but NOT by the reference's recursive radix-2 DCT-II/DCT-IV factorisation: here it is one
// S/2-point complex FFT between a pre- and a post-twiddle,
//     z[n] = (v[2n] + i v[S-1-2n]) E[n],  Z = FFT_{S/2}(z),  y = Z E,
//     X[2k] = Re y[k],  X[S-1-2k] = -Im y[k],          E[n] = exp(-i pi (8n+1)/(8S)),
// run in place in shared memory by the whole CTA for `nT` transforms of equal size at once.
// The FFT is an in-place radix-4 decimation-in-frequency (one radix-2 stage when log2 is odd);
// its digit-reversed output order is undone inside the post-twiddle pass via UlcbTables::fftPos.
#pragma once
#include "ulcb_tables.h"

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place DCT-IV of nT transforms of size S.  Transform t occupies base[t*stride .. t*stride+S).
// sh = log2(N/S) selects the twiddle tables.  All threads of the CTA must call this.
// MAXPAIRS bounds the register staging of the post pass: nT*S/4 <= MAXPAIRS*blockDim.x.
// WARP = true: the calling warp alone owns the transform(s) (32 "threads", __syncwarp between passes).
template <int MAXPAIRS, bool WARP = false>
__device__ void dct4_inplace(float *base, int stride, int nT, int S, int sh, const UlcbTables &tb) {
    const int tid = WARP ? (threadIdx.x & 31) : threadIdx.x, nthr = WARP ? 32 : blockDim.x;
    auto sync = [] { if (WARP) __syncwarp(); else __syncthreads(); };
    const int M = S >> 1;                       // complex FFT size
    const float2 *__restrict__ E = tb.dctE[sh];
    const int lgM = 31 - __clz(M);              // every size is a power of two: index splits are shifts and masks, not divisions

    // ---- pre-twiddle: pairs (n, M-1-n), n < M/2 ----
    const int halfM = M >> 1;
    for (int idx = tid; idx < nT * halfM; idx += nthr) {
        int t = idx >> (lgM - 1), n = idx & (halfM - 1), m = M - 1 - n;
        float2 *z = reinterpret_cast<float2 *>(base + t * stride);
        float2 a = z[n], b = z[m];              // a = (v[2n], v[2n+1]); b = (v[2m], v[2m+1])
        float2 zn = cmul(make_float2(a.x, b.y), __ldg(&E[n]));
        float2 zm = cmul(make_float2(b.x, a.y), __ldg(&E[m]));
        z[n] = zn;
        z[m] = zm;
    }
    sync();

    // ---- radix-4 DIF stages ----
    const int Mmax = tb.N >> 1;
    const float2 *__restrict__ W = tb.fftW;
    int L = M, lgL = lgM;
    const int lgMmax = 31 - __clz(Mmax);
    for (; L >= 4; L >>= 2, lgL -= 2) {
        const int q = L >> 2;                   // quarter length
        const int wstep = 1 << (lgMmax - lgL);
        for (int idx = tid; idx < nT * (M >> 2); idx += nthr) {
            int t = idx >> (lgM - 2), w = idx & ((M >> 2) - 1);
            int g = w >> (lgL - 2), j = w & (q - 1);
            float2 *x = reinterpret_cast<float2 *>(base + t * stride) + g * L + j;
            float2 a0 = x[0], a1 = x[q], a2 = x[2 * q], a3 = x[3 * q];
            float2 b0 = make_float2(a0.x + a2.x, a0.y + a2.y);
            float2 b1 = make_float2(a0.x - a2.x, a0.y - a2.y);
            float2 b2 = make_float2(a1.x + a3.x, a1.y + a3.y);
            float2 b3 = make_float2(a1.y - a3.y, a3.x - a1.x);      // (a1-a3) * (-i)
            float2 y0 = make_float2(b0.x + b2.x, b0.y + b2.y);
            float2 y1 = make_float2(b1.x + b3.x, b1.y + b3.y);
            float2 y2 = make_float2(b0.x - b2.x, b0.y - b2.y);
            float2 y3 = make_float2(b1.x - b3.x, b1.y - b3.y);
            if (q > 1) {
                int k = j * wstep;
                y1 = cmul(y1, __ldg(&W[k]));
                y2 = cmul(y2, __ldg(&W[2 * k]));
                y3 = cmul(y3, __ldg(&W[3 * k]));
            }
            x[0] = y0; x[q] = y1; x[2 * q] = y2; x[3 * q] = y3;
        }
        sync();
    }
    if (L == 2) {                               // odd log2(M): one radix-2 stage on adjacent pairs
        for (int idx = tid; idx < nT * (M >> 1); idx += nthr) {
            int t = idx >> (lgM - 1), g = idx & ((M >> 1) - 1);
            float2 *x = reinterpret_cast<float2 *>(base + t * stride) + 2 * g;
            float2 a0 = x[0], a1 = x[1];
            x[0] = make_float2(a0.x + a1.x, a0.y + a1.y);
            x[1] = make_float2(a0.x - a1.x, a0.y - a1.y);
        }
        sync();
    }

    // ---- post-twiddle + un-permute: pairs (k, M-1-k); staged in registers across a barrier ----
    const uint16_t *__restrict__ P = tb.fftPos[sh];
    float2 o0[MAXPAIRS], o1[MAXPAIRS];
#pragma unroll
    for (int r = 0; r < MAXPAIRS; r++) {
        int idx = tid + r * nthr;
        if (idx < nT * halfM) {
            int t = idx >> (lgM - 1), k = idx & (halfM - 1), m = M - 1 - k;
            const float2 *z = reinterpret_cast<const float2 *>(base + t * stride);
            float2 yk = cmul(z[__ldg(&P[k])], __ldg(&E[k]));
            float2 ym = cmul(z[__ldg(&P[m])], __ldg(&E[m]));
            o0[r] = make_float2(yk.x, -ym.y);   // X[2k], X[2k+1] = X[S-1-2m]
            o1[r] = make_float2(ym.x, -yk.y);   // X[2m], X[2m+1] = X[S-1-2k]
        }
    }
    sync();
#pragma unroll
    for (int r = 0; r < MAXPAIRS; r++) {
        int idx = tid + r * nthr;
        if (idx < nT * halfM) {
            int t = idx >> (lgM - 1), k = idx & (halfM - 1), m = M - 1 - k;
            float2 *z = reinterpret_cast<float2 *>(base + t * stride);
            z[k] = o0[r];
            z[m] = o1[r];
        }
    }
    sync();
}
