// Shared device helpers for the DPI-VAE sm_100a kernels.
//
// Data layout inside a CTA: every activation matrix lives in shared memory FEATURE-MAJOR,
// row f = feature, 64 "units" (rows of the minibatch for the encoder kernels, (row, MC-sample)
// pairs for the decoder kernel) contiguous along the row, leading dimension LDP = 68 floats.
// LDP = 4 (mod 32) makes 8 consecutive rows start in 8 different 16-byte bank groups, so the
// LDS.128 along the unit axis (forward / dgrad) AND the LDS.128 along the reduction axis from
// 8 different rows (wgrad, dgrad weight operand) are both conflict-free.
//
// Every GEMM flavour comes in three register-tile shapes; the dispatcher picks the largest tile
// that still gives all 256 threads a tile, so the skinny layers (K or N of 2..8) do not leave
// most of the CTA idle at the next barrier.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <curand_kernel.h>
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

// One element of torch.cuda's normal_() stream (ATen/native/cuda/DistributionTemplates.h,
// distribution_elementwise_grid_stride_kernel): thread `sub` of a grid of T threads draws 4 normals per
// curand_normal4 call; call number `it` serves elements sub + T * (4 it + comp).  Equivalent to
//   curand_init(seed, sub, offset + 4 it, &st); curand_normal4(&st)[comp]
// but with ONE Philox4x32-10 evaluation and ONE Box-Muller pair instead of the four evaluations curand_init +
// curand4 perform (skipahead_sequence, skipahead, the draw and the look-ahead): counter = (offset / 4 + it, sub),
// key = seed; offsets of torch's generator are multiples of 4, so curand's intra-counter position is 0.
// Bitwise identity with torch is pinned by tests/test_gpu_parity.py::test_philox_reproduces_torch_cuda_stream.
__device__ __forceinline__ float philox_normal_elem(unsigned long long seed, unsigned long long offset, unsigned int T,
                                                    unsigned long long li) {
  unsigned long long sub, q4;
  if ((li >> 32) == 0ull) {   // common case: 32-bit division (a 64-bit divide costs ~100 instructions per element)
    const unsigned int l32 = (unsigned int)li, q32 = l32 / T;
    q4 = q32;
    sub = l32 - q32 * T;
  } else {
    sub = li % T;
    q4 = li / T;
  }
  const unsigned long long n = (offset >> 2) + (q4 >> 2);
  const int comp = (int)(q4 & 3ull);
  const uint4 ctr = make_uint4((unsigned int)n, (unsigned int)(n >> 32), (unsigned int)sub, (unsigned int)(sub >> 32));
  const uint2 key = make_uint2((unsigned int)seed, (unsigned int)(seed >> 32));
  const uint4 r = curand_Philox4x32_10(ctr, key);
  const float2 g = _curand_box_muller(comp < 2 ? r.x : r.z, comp < 2 ? r.y : r.w);
  return (comp & 1) ? g.y : g.x;
}

// Block-wide copy of n floats global -> shared with EIGHT independent loads in flight per thread (a rolled copy loop waits
// a full memory latency, ~0.6 us, per trip: the kernels' set-up copies of 10-16 k weights took 15-20 us that way); 16-byte
// loads when both ends are 16-byte aligned (a quarter of the round trips), scalar loads for the rest.
template <int NT>
__device__ __forceinline__ void copy_g2s_batched(float* __restrict__ dst, const float* __restrict__ src, int n) {
  const int tid = threadIdx.x;
  int done = 0;
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0) {
    const int n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int e0 = 0; e0 < n4; e0 += 8 * NT) {
      float4 t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { const int e = e0 + k * NT + tid; t[k] = e < n4 ? __ldg(s4 + e) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
      for (int k = 0; k < 8; ++k) { const int e = e0 + k * NT + tid; if (e < n4) d4[e] = t[k]; }
    }
    done = n4 << 2;
  }
  for (int e0 = done; e0 < n; e0 += 8 * NT) {
    float t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { const int e = e0 + k * NT + tid; t[k] = e < n ? __ldg(src + e) : 0.0f; }
#pragma unroll
    for (int k = 0; k < 8; ++k) { const int e = e0 + k * NT + tid; if (e < n) dst[e] = t[k]; }
  }
}

// contiguous vector load / store of V floats (V = 1, 2, 4) from shared memory
template <int V>
__device__ __forceinline__ void ldv(const float* p, float* out) {
  if (V == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else if (V == 2) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    out[0] = v.x; out[1] = v.y;
  } else {
    out[0] = *p;
  }
}
template <int V>
__device__ __forceinline__ void stv(float* p, const float* in) {
  if (V == 4) *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
  else if (V == 2) *reinterpret_cast<float2*>(p) = make_float2(in[0], in[1]);
  else *p = in[0];
}

// ------------------------------------------------------------------------------------------
// Weight staging: global nn.Linear layout w[N][K] -> shared Wt[Kp][ldw] (n contiguous), bias[Np].
// The destination must have been zero-filled (pad rows/cols stay zero).
// ------------------------------------------------------------------------------------------
__device__ inline void stage_linear(const float* __restrict__ w, const float* __restrict__ b, int K, int N,
                                    float* __restrict__ Wt, int ldw, float* __restrict__ bias) {
  for (int e = threadIdx.x; e < K * N; e += NT) {
    const int n = e / K, k = e - n * K;
    Wt[k * ldw + n] = w[e];
  }
  for (int e = threadIdx.x; e < N; e += NT) bias[e] = b[e];
}

// ------------------------------------------------------------------------------------------
// O[n][u] = act(bias[n] + sum_k Wt[k][n] * A[k][u]),  u in [0,64), n in [0,Np), Np % 4 == 0
// ------------------------------------------------------------------------------------------
template <int ACT, int TM, int TN>
__device__ __forceinline__ void gemm_fwd_t(const float* __restrict__ Wt, int ldw, const float* __restrict__ bias,
                                           const float* __restrict__ A, float* __restrict__ O, int K, int Np) {
  constexpr int MT = TILE / TM;
  const int ntiles = MT * (Np / TN);
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    const int tm = t % MT, tn = t / MT;
    float acc[TN][TM];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const float bj = bias[TN * tn + j];
#pragma unroll
      for (int i = 0; i < TM; ++i) acc[j][i] = bj;
    }
    const float* a = A + TM * tm;
    const float* w = Wt + TN * tn;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float am[TM], wn[TN];
      ldv<TM>(a + k * LDP, am);
      ldv<TN>(w + k * ldw, wn);
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int i = 0; i < TM; ++i) acc[j][i] = fmaf(wn[j], am[i], acc[j][i]);
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float o[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        float v = acc[j][i];
        if (ACT == ACT_RELU) v = fmaxf(v, 0.0f);
        if (ACT == ACT_TANH) v = tanhf(v);
        o[i] = v;
      }
      stv<TM>(O + (TN * tn + j) * LDP + TM * tm, o);
    }
  }
}

template <int ACT>
__device__ __forceinline__ void gemm_fwd(const float* __restrict__ Wt, int ldw, const float* __restrict__ bias,
                                         const float* __restrict__ A, float* __restrict__ O, int K, int Np) {
  if (Np >= 64) gemm_fwd_t<ACT, 4, 4>(Wt, ldw, bias, A, O, K, Np);
  else if (Np >= 32) gemm_fwd_t<ACT, 2, 4>(Wt, ldw, bias, A, O, K, Np);
  else if (Np >= 16) gemm_fwd_t<ACT, 1, 4>(Wt, ldw, bias, A, O, K, Np);
  else gemm_fwd_t<ACT, 1, 1>(Wt, ldw, bias, A, O, K, Np);
}

// ------------------------------------------------------------------------------------------
// Out[k][u] = (sum_n Wt[k][n] * G[n][u]) * act'(Aact[k][u]),  k in [0,Kp), Kp % 4 == 0, Np % 4 == 0
// Thread tile: rows {tk + i*KT} (interleaved -> conflict-free weight loads) x TM units.
// In place (Out == Aact) is safe: a thread reads exactly the elements it overwrites.
// ------------------------------------------------------------------------------------------
template <int ACT, int TK, int TM>
__device__ __forceinline__ void gemm_dgrad_t(const float* __restrict__ Wt, int ldw, const float* __restrict__ G,
                                             const float* Aact, float* Out, int Kp, int Np) {
  constexpr int MT = TILE / TM;
  const int KT = Kp / TK;
  const int ntiles = MT * KT;
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    const int tm = t % MT, tk = t / MT;
    float acc[TK][TM];
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
      for (int m = 0; m < TM; ++m) acc[i][m] = 0.0f;
    const float* g = G + TM * tm;
#pragma unroll 2
    for (int n = 0; n < Np; n += 4) {
      float gv[4][TM];
#pragma unroll
      for (int j = 0; j < 4; ++j) ldv<TM>(g + (n + j) * LDP, gv[j]);
#pragma unroll
      for (int i = 0; i < TK; ++i) {
        float wj[4];
        ldv<4>(Wt + (tk + i * KT) * ldw + n, wj);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int m = 0; m < TM; ++m) acc[i][m] = fmaf(wj[j], gv[j][m], acc[i][m]);
      }
    }
#pragma unroll
    for (int i = 0; i < TK; ++i) {
      const int row = (tk + i * KT) * LDP + TM * tm;
      float o[TM];
#pragma unroll
      for (int m = 0; m < TM; ++m) o[m] = acc[i][m];
      if (ACT != ACT_NONE) {
        float a[TM];
        ldv<TM>(Aact + row, a);
#pragma unroll
        for (int m = 0; m < TM; ++m) {
          if (ACT == ACT_RELU) o[m] = a[m] > 0.0f ? o[m] : 0.0f;
          else o[m] *= (1.0f - a[m] * a[m]);
        }
      }
      stv<TM>(Out + row, o);
    }
  }
}

template <int ACT>
__device__ __forceinline__ void gemm_dgrad(const float* __restrict__ Wt, int ldw, const float* __restrict__ G,
                                           const float* Aact, float* Out, int Kp, int Np) {
  if (Kp >= 64) gemm_dgrad_t<ACT, 4, 4>(Wt, ldw, G, Aact, Out, Kp, Np);
  else if (Kp >= 32) gemm_dgrad_t<ACT, 2, 4>(Wt, ldw, G, Aact, Out, Kp, Np);
  else if (Kp >= 16) gemm_dgrad_t<ACT, 1, 4>(Wt, ldw, G, Aact, Out, Kp, Np);
  else gemm_dgrad_t<ACT, 1, 1>(Wt, ldw, G, Aact, Out, Kp, Np);
}

// ------------------------------------------------------------------------------------------
// dW[n][k] += sum_u A[k][u] * G[n][u] ; db[n] += sum_u G[n][u]     (u over the 64 units)
// dW/db: this CTA's private accumulators in global memory (L2-resident), nn.Linear layout
// dW[n*K + k].  Each accumulator has exactly one writer thread, so the fire-and-forget
// RED.ADD keeps the sum order deterministic while not stalling on the L2 round trip.
// Thread tile: k rows {tk + i*KT} x n rows {tn + j*NTn}, LDS.128 along u for both operands.
// ------------------------------------------------------------------------------------------
template <int TK, int TN>
__device__ __forceinline__ void gemm_wgrad_t(const float* __restrict__ A, const float* __restrict__ G,
                                             float* __restrict__ dW, float* __restrict__ db, int K, int N) {
  const int KT = pad4(K) / TK, NTn = pad4(N) / TN;
  const int ntiles = KT * NTn;
  // lane -> tile map: a warp covers an 8 (k) x 4 (n) block of tiles when the shape allows, so each
  // LDS.128 touches <= 8 distinct rows (2 shared-memory wavefronts instead of 4 for 32 distinct rows)
  const bool blocked = (KT % 8 == 0) && (NTn % 4 == 0);
  const int nbk = KT >> 3;
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    int tk, tn;
    if (blocked) {
      const int bid = t >> 5, l = t & 31;
      tk = (bid % nbk) * 8 + (l & 7);
      tn = (bid / nbk) * 4 + (l >> 3);
    } else {
      tk = t % KT;
      tn = t / KT;
    }
    float acc[TK][TN];
    float bs[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bs[j] = 0.0f;
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
#pragma unroll 2
    for (int u = 0; u < TILE; u += 4) {
      float av[TK][4], gv[TN][4];
#pragma unroll
      for (int i = 0; i < TK; ++i) ldv<4>(A + (tk + i * KT) * LDP + u, av[i]);
#pragma unroll
      for (int j = 0; j < TN; ++j) ldv<4>(G + (tn + j * NTn) * LDP + u, gv[j]);
#pragma unroll
      for (int i = 0; i < TK; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][j] = fmaf(av[i][q], gv[j][q], acc[i][j]);
      if (tk == 0) {
#pragma unroll
        for (int j = 0; j < TN; ++j) bs[j] += (gv[j][0] + gv[j][1]) + (gv[j][2] + gv[j][3]);
      }
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = tn + j * NTn;
      if (n < N) {
#pragma unroll
        for (int i = 0; i < TK; ++i) {
          const int k = tk + i * KT;
          if (k < K) atomicAdd(&dW[n * K + k], acc[i][j]);
        }
        if (tk == 0) atomicAdd(&db[n], bs[j]);
      }
    }
  }
}

__device__ __forceinline__ void gemm_wgrad(const float* __restrict__ A, const float* __restrict__ G,
                                           float* __restrict__ dW, float* __restrict__ db, int K, int N) {
  const int kn = pad4(K) * pad4(N);
  if (kn >= 4096) gemm_wgrad_t<4, 4>(A, G, dW, db, K, N);
  else if (kn >= 1024) gemm_wgrad_t<2, 2>(A, G, dW, db, K, N);
  else gemm_wgrad_t<1, 1>(A, G, dW, db, K, N);
}

}  // namespace dpv
// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while the kernel before it on the
// stream is still running -- its prologue (shared-memory zeroing, weight staging: inputs that the previous kernel does
// not write) overlaps the previous kernel's tail. pdl_wait() blocks until the previous kernel has completed and its
// writes are visible; EVERY thread block of the dependent kernel must execute it before touching those writes (and
// at least once in any case, so that the dependent cannot complete before its predecessor). The predecessor calls
// pdl_launch_dependents() at its top to release the early launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifdef __CUDACC__
template <typename K, typename PT>
static inline void launch_pdl3(K kernel, dim3 grid, int block, size_t smem, cudaStream_t s, const PT& params) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, params);
}
template <typename K, typename PT>
static inline void launch_pdl(K kernel, int grid, int block, size_t smem, cudaStream_t s, const PT& params) {
  launch_pdl3(kernel, dim3((unsigned)grid), block, smem, s, params);
}
#endif
