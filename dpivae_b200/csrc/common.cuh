// Shared device helpers for the DPI-VAE sm_100a kernels.
//
// Data layout inside a CTA: every activation matrix lives in shared memory FEATURE-MAJOR,
// row f = feature, 64 "units" (rows of the minibatch for the encoder kernels, (row, MC-sample)
// pairs for the decoder kernel) contiguous along the row, leading dimension LDP = 68 floats.
// LDP = 4 (mod 32) makes 8 consecutive rows start in 8 different 16-byte bank groups, so the
// LDS.128 along the unit axis (forward / dgrad) AND the LDS.128 along the reduction axis from
// 8 different rows (wgrad, dgrad weight operand) are both conflict-free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpv {

constexpr int NT = 256;      // threads per CTA
constexpr int TILE = 64;     // units per tile
constexpr int LDP = 68;      // leading dimension of activation rows
constexpr float LOG_2PI = 1.8378770664093453f;
constexpr float LOG_SQRT_2PI = 0.9189385332046727f;

__host__ __device__ inline int pad4(int v) { return (v + 3) & ~3; }

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// torch.nn.functional.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float clampf_(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ------------------------------------------------------------------------------------------
// Weight staging: global nn.Linear layout w[N][K] -> shared Wt[Kp][ldw] (n contiguous), bias[Np].
// The destination must have been zero-filled (pad rows/cols stay zero).
// ------------------------------------------------------------------------------------------
__device__ inline void stage_linear(const float* __restrict__ w, const float* __restrict__ b, int K, int N,
                                    float* __restrict__ Wt, int ldw, float* __restrict__ bias) {
  for (int e = threadIdx.x; e < K * N; e += NT) {
    int n = e / K, k = e - n * K;
    Wt[k * ldw + n] = w[e];
  }
  for (int e = threadIdx.x; e < N; e += NT) bias[e] = b[e];
}

// ------------------------------------------------------------------------------------------
// O[n][u] = act(bias[n] + sum_k Wt[k][n] * A[k][u]),  u in [0,64), n in [0,Np)
// 4x4 register tiles: thread -> (4 units) x (4 outputs); 16 FFMA per 2 LDS.128.
// ------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ void gemm_fwd(const float* __restrict__ Wt, int ldw, const float* __restrict__ bias,
                                         const float* __restrict__ A, float* __restrict__ O, int K, int Np) {
  const int ntiles = 16 * (Np >> 2);
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    const int tm = t & 15, tn = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float bj = bias[4 * tn + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = bj;
    }
    const float* a = A + 4 * tm;
    const float* w = Wt + 4 * tn;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(a + k * LDP);
      const float4 wv = *reinterpret_cast<const float4*>(w + k * ldw);
      const float am[4] = {av.x, av.y, av.z, av.w};
      const float wn[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(wn[j], am[i], acc[j][i]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 o;
      float* op = reinterpret_cast<float*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = acc[j][i];
        if (ACT == ACT_RELU) v = fmaxf(v, 0.0f);
        if (ACT == ACT_TANH) v = tanhf(v);
        op[i] = v;
      }
      *reinterpret_cast<float4*>(O + (4 * tn + j) * LDP + 4 * tm) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Out[k][u] = (sum_n Wt[k][n] * G[n][u]) * act'(Aact[k][u]),  k in [0,Kp), Kp % 4 == 0
// Thread tile: rows {tk + i*KT} (interleaved -> conflict-free weight loads) x 4 units.
// In place (Out == Aact) is safe: a thread reads exactly the elements it overwrites.
// ------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ void gemm_dgrad(const float* __restrict__ Wt, int ldw, const float* __restrict__ G,
                                           const float* Aact, float* Out, int Kp, int Np) {
  const int KT = Kp >> 2;
  const int ntiles = 16 * KT;
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    const int tm = t & 15, tk = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int m = 0; m < 4; ++m) acc[i][m] = 0.0f;
    const float* g = G + 4 * tm;
    for (int n = 0; n < Np; n += 4) {
      float4 gv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) gv[j] = *reinterpret_cast<const float4*>(g + (n + j) * LDP);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 wv = *reinterpret_cast<const float4*>(Wt + (tk + i * KT) * ldw + n);
        const float wj[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][0] = fmaf(wj[j], gv[j].x, acc[i][0]);
          acc[i][1] = fmaf(wj[j], gv[j].y, acc[i][1]);
          acc[i][2] = fmaf(wj[j], gv[j].z, acc[i][2]);
          acc[i][3] = fmaf(wj[j], gv[j].w, acc[i][3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = (tk + i * KT) * LDP + 4 * tm;
      float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (ACT != ACT_NONE) {
        const float4 a = *reinterpret_cast<const float4*>(Aact + row);
        if (ACT == ACT_RELU) {
          o.x = a.x > 0.0f ? o.x : 0.0f; o.y = a.y > 0.0f ? o.y : 0.0f;
          o.z = a.z > 0.0f ? o.z : 0.0f; o.w = a.w > 0.0f ? o.w : 0.0f;
        } else {
          o.x *= (1.0f - a.x * a.x); o.y *= (1.0f - a.y * a.y);
          o.z *= (1.0f - a.z * a.z); o.w *= (1.0f - a.w * a.w);
        }
      }
      *reinterpret_cast<float4*>(Out + row) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------
// dW[n][k] += sum_u A[k][u] * G[n][u] ; db[n] += sum_u G[n][u]     (u over the 64 units)
// dW/db: this CTA's private accumulators in global memory, nn.Linear layout dW[n*K + k].
// Thread tile: k rows {tk + i*KT} x n rows {tn + j*NTn}, LDS.128 along u for both operands.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void gemm_wgrad(const float* __restrict__ A, const float* __restrict__ G,
                                           float* __restrict__ dW, float* __restrict__ db, int K, int N) {
  const int KT = pad4(K) >> 2, NTn = pad4(N) >> 2;
  const int ntiles = KT * NTn;
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    const int tk = t % KT, tn = t / KT;
    float acc[4][4];
    float bs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
    for (int u = 0; u < TILE; u += 4) {
      float4 av[4], gv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(A + (tk + i * KT) * LDP + u);
#pragma unroll
      for (int j = 0; j < 4; ++j) gv[j] = *reinterpret_cast<const float4*>(G + (tn + j * NTn) * LDP + u);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(av[i].x, gv[j].x, acc[i][j]);
          acc[i][j] = fmaf(av[i].y, gv[j].y, acc[i][j]);
          acc[i][j] = fmaf(av[i].z, gv[j].z, acc[i][j]);
          acc[i][j] = fmaf(av[i].w, gv[j].w, acc[i][j]);
        }
      if (tk == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bs[j] += (gv[j].x + gv[j].y) + (gv[j].z + gv[j].w);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = tn + j * NTn;
      if (n < N) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = tk + i * KT;
          if (k < K) dW[n * K + k] += acc[i][j];
        }
        if (tk == 0) db[n] += bs[j];
      }
    }
  }
}

}  // namespace dpv
