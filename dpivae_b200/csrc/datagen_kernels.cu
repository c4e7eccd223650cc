// On-device synthetic data generator = utils/data.py:9-52 `sample_response` for cases whose
// `full_model` is a Tanh-MLP surrogate behind a StandardScaler (all three reference cases, cases/*/__init__.py):
//   z_j ~ Uniform(lo_j, hi_j) per generative factor (utils/priors.py:32-36), x = full_model(z) + N(0, sigma_x^2),
//   c = z[idx_c] + N(0, sigma_c^2), y = z[idx_y] + N(0, sigma_y^2).
// Random numbers: Philox4x32-10 in torch.cuda's element mapping, i.e. bit-identical to the sequence
//   torch.rand(n) (one call per factor) ; torch.randn(n, nd_x) ; torch.randn(n, nd_c) ; torch.randn(n, nd_y)
// drawn from a CUDA generator at the same seed / offset (tests/test_gpu_datagen.py), which is what the reference's
// torch.distributions calls reduce to when they run on the device.
#include <curand_kernel.h>

#include "common.cuh"
#include "kernels.h"

namespace dpv {

namespace {

// torch.cuda uniform_ (ATen/native/cuda/DistributionTemplates.h uniform_kernel): curand_uniform4 -> (0, 1], 1 -> 0
__device__ __forceinline__ float philox_uniform_elem(unsigned long long seed, unsigned long long offset, unsigned int T,
                                                     unsigned long long li) {
  const unsigned long long sub = li % T, q4 = li / T;
  const unsigned long long n = (offset >> 2) + (q4 >> 2);
  const int comp = (int)(q4 & 3ull);
  const uint4 ctr = make_uint4((unsigned int)n, (unsigned int)(n >> 32), (unsigned int)sub, (unsigned int)(sub >> 32));
  const uint2 key = make_uint2((unsigned int)seed, (unsigned int)(seed >> 32));
  const uint4 r = curand_Philox4x32_10(ctr, key);
  const unsigned int v = comp == 0 ? r.x : (comp == 1 ? r.y : (comp == 2 ? r.z : r.w));
  const float u = _curand_uniform(v);
  return u == 1.0f ? 0.0f : u;
}

// z (n, nf) and the standardised surrogate input A0 (n, nf)
__global__ void __launch_bounds__(256) datagen_latents_kernel(const DataGenParams P) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= P.n * P.nf) return;
  const long long i = e / P.nf;
  const int j = (int)(e - i * P.nf);
  const float u = philox_uniform_elem(P.seed, P.off_u[j], P.T_u, (unsigned long long)i);
  // torch.distributions.Uniform.rsample: low + rand * (high - low), each op rounded separately (no FMA contraction)
  const float z = __fadd_rn(P.lo[j], __fmul_rn(u, __fsub_rn(P.hi[j], P.lo[j])));
  P.z[e] = z;
  P.a0[e] = (z - P.in_mean[j]) / P.in_std[j];           // utils/transforms.py:70-73
}

// O[n][N] = act(A[n][K] W[N][K]^T + b): 64 x 64 output tile, 256 threads x (4 x 4), K in steps of 16
template <int ACT>
__global__ void __launch_bounds__(256) mlp_layer_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                        const float* __restrict__ bias, float* __restrict__ O, long long n, int K,
                                                        int N) {
  __shared__ float As[16][68], Ws[16][68];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long row0 = (long long)blockIdx.x * 64;
  const int col0 = blockIdx.y * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = tid; e < 64 * 16; e += 256) {
      const int r = e >> 4, k = e & 15;
      As[k][r] = (row0 + r < n && k0 + k < K) ? A[(row0 + r) * K + k0 + k] : 0.0f;
      Ws[k][r] = (col0 + r < N && k0 + k < K) ? W[(long long)(col0 + r) * K + k0 + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][4 * ty]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][4 * tx]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = row0 + 4 * ty + i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + 4 * tx + j;
      if (c < N) {
        const float v = acc[i][j] + bias[c];
        O[r * N + c] = ACT == ACT_TANH ? tanhf(v) : v;
      }
    }
  }
}

// x += sigma_x eps_x ; c = z[idx_c] + sigma_c eps_c ; y = z[idx_y] + sigma_y eps_y   (utils/data.py:37-46)
__global__ void __launch_bounds__(256) datagen_finish_kernel(const DataGenParams P) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long nx = P.n * P.nd_x, nc = P.n * P.nd_c, ny = P.n * P.nd_y;
  if (e < nx) {
    P.x[e] = __fadd_rn(P.x[e], __fmul_rn(philox_normal_elem(P.seed, P.off_x, P.T_x, (unsigned long long)e), P.sigma_x));
  } else if (e < nx + nc) {
    const long long q = e - nx, i = q / P.nd_c;
    const int j = (int)(q - i * P.nd_c);
    P.c[q] = __fadd_rn(P.z[i * P.nf + P.idx_c[j]], __fmul_rn(philox_normal_elem(P.seed, P.off_c, P.T_c, (unsigned long long)q), P.sigma_c));
  } else if (e < nx + nc + ny) {
    const long long q = e - nx - nc, i = q / P.nd_y;
    const int j = (int)(q - i * P.nd_y);
    P.y[q] = __fadd_rn(P.z[i * P.nf + P.idx_y[j]], __fmul_rn(philox_normal_elem(P.seed, P.off_y, P.T_y, (unsigned long long)q), P.sigma_y));
  }
}

}  // namespace

void launch_datagen_latents(const DataGenParams& p, cudaStream_t s) {
  const long long tot = p.n * p.nf;
  datagen_latents_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(p);
}
void launch_mlp_layer(const float* A, const float* W, const float* b, float* O, long long n, int K, int N, bool tanh_act, cudaStream_t s) {
  const dim3 grid((unsigned)((n + 63) / 64), (unsigned)((N + 63) / 64));
  if (tanh_act) mlp_layer_kernel<ACT_TANH><<<grid, 256, 0, s>>>(A, W, b, O, n, K, N);
  else mlp_layer_kernel<ACT_NONE><<<grid, 256, 0, s>>>(A, W, b, O, n, K, N);
}
void launch_datagen_finish(const DataGenParams& p, cudaStream_t s) {
  const long long tot = p.n * (p.nd_x + p.nd_c + p.nd_y);
  datagen_finish_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(p);
}

}  // namespace dpv
