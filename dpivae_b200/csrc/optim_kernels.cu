// Deterministic cross-CTA gradient reduction, loss-scalar finalisation, clip_grad_norm_ and the
// fused multi-group Adam update over the flat parameter buffer (dpivae.py:335-373,419-436;
// torch.optim.Adam: lerp first moment, L2 weight decay, bias-corrected step, eps 1e-8).
#include "common.cuh"
#include "kernels.h"

namespace dpv {

// grads[i] = sum over the owning kernel's CTAs (fixed order) of part[cta][i]; block 0 also folds
// the 6 per-CTA loss sums into the 8 normalised scalars (dpivae.py:419-426).
// Thread (parameter lane, row group): 32 consecutive parameters per block (coalesced partial rows), the CTA rows
// split over 8 row groups with 4 independent accumulators each, combined through shared memory in a fixed order
// (deterministic; the serial 148..296-load chain per parameter made this kernel latency-bound at 70 us).
__device__ __forceinline__ void adam_update(const AdamParams& P, long long i, float g);
constexpr int RED_P = 32, RED_G = 8;
__global__ void __launch_bounds__(RED_P * RED_G) reduce_kernel(const ReduceParams P) {
  __shared__ float sm[RED_G][RED_P];
  const int lane = threadIdx.x & (RED_P - 1), rg = threadIdx.x / RED_P;
  const long long i = (long long)blockIdx.x * RED_P + lane;
  float s = 0.0f;
  if (P.grads != nullptr && i < P.n_params) {
    const int cls = P.owner[i];
    const float* base = P.part + P.base[cls] * P.part_stride + i;
    const int ncta = P.n_cta[cls];
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    int c = rg;
    for (; c + 3 * RED_G < ncta; c += 4 * RED_G) {
      a0 += base[(long long)c * P.part_stride];
      a1 += base[(long long)(c + RED_G) * P.part_stride];
      a2 += base[(long long)(c + 2 * RED_G) * P.part_stride];
      a3 += base[(long long)(c + 3 * RED_G) * P.part_stride];
    }
    for (; c < ncta; c += RED_G) a0 += base[(long long)c * P.part_stride];
    s = (a0 + a1) + (a2 + a3);
  }
  sm[rg][lane] = s;
  __syncthreads();
  if (rg == 0 && P.grads != nullptr && i < P.n_params) {
    float t = sm[0][lane];
#pragma unroll
    for (int g = 1; g < RED_G; ++g) t += sm[g][lane];
    P.grads[i] = t;
    if (P.fuse_adam) adam_update(P.adam, i, t);
  }
  if (blockIdx.x == 0 && P.scalars != nullptr) {
    // the 6 loss sums: 32 row groups per scalar, then a fixed-order sum of the 32 partials
    __shared__ float ssum[32][8];
    const int k = threadIdx.x & 7, g = threadIdx.x >> 3;
    float t = 0.0f;
    if (k < 6)
      for (int c = g; c < P.n_cta[0]; c += 32) t += P.part[(P.base[0] + c) * P.part_stride + P.n_params + k];
    ssum[g][k] = t;
    __syncthreads();
    if (threadIdx.x < 6) {
      float tot = 0.0f;
      for (int gg = 0; gg < 32; ++gg) tot += ssum[gg][threadIdx.x];
      const int kk = threadIdx.x;
      // order: ELBO, KLx, KLc(0), KLy(0), Rx, Rc, Ry, reg
      if (kk == 0) P.scalars[0] = tot * P.inv_BD;
      else if (kk == 1) { P.scalars[1] = tot * P.inv_B; P.scalars[2] = 0.0f; P.scalars[3] = 0.0f; }
      else P.scalars[kk + 2] = tot * P.inv_B;
      if (P.fuse_adam && P.adam.log != nullptr && P.adam.ss != nullptr) {   // captured-step log row (what adam_kernel copies otherwise)
        float* row = P.adam.log + ((P.adam.ss->step - 1) % P.adam.log_cap) * 9;
        if (kk == 0) row[0] = tot * P.inv_BD;
        else if (kk == 1) { row[1] = tot * P.inv_B; row[2] = 0.0f; row[3] = 0.0f; }
        else row[kk + 2] = tot * P.inv_B;
      }
    }
  }
}

// clip_coef = min(1, max_norm / (||g||_2 + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void __launch_bounds__(1024) gradnorm_kernel(const float* __restrict__ g, long long n, float max_norm,
                                                        float* __restrict__ clip_coef) {
  __shared__ float red[1024];
  float s = 0.0f;
  for (long long i = threadIdx.x; i < n; i += 1024) s = fmaf(g[i], g[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float c = max_norm / (sqrtf(red[0]) + 1e-6f);
    *clip_coef = c < 1.0f ? c : 1.0f;
  }
}

// torch.optim.Adam arithmetic for parameter i with gradient g (weight decay, exp_avg / exp_avg_sq, bias-corrected step);
// in captured-step mode also the log_sigma_x entry of the step's log row
__device__ __forceinline__ void adam_update(const AdamParams& P, long long i, float g) {
  const int gid = P.group[i];
  const float p = P.params[i];
  const float wd = P.wd[gid];
  if (wd != 0.0f) g = fmaf(wd, p, g);
  float m = P.m[i], v = P.v[i];
  m = fmaf(1.0f - P.beta1, g - m, m);              // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(1.0f - P.beta2, g * g, v * P.beta2);    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float bc2_sqrt = P.ss ? P.ss->bc2_sqrt : P.bc2_sqrt;
  const float step_size = P.ss ? P.ss->step_size[gid] : P.step_size[gid];
  const float denom = sqrtf(v) / bc2_sqrt + P.eps;
  const float pn = p - step_size * (m / denom);
  P.params[i] = pn;
  P.m[i] = m;
  P.v[i] = v;
  if (P.log != nullptr && P.ss != nullptr && i == P.lsx_index) P.log[((P.ss->step - 1) % P.log_cap) * 9 + 8] = pn;
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamParams P) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_params) return;
  float g = P.grads[i];
  if (P.clip_coef != nullptr) {
    g *= *P.clip_coef;
    const_cast<float*>(P.grads)[i] = g;
  }
  adam_update(P, i, g);
  // per-step log row of the device-resident loop (dpivae.py:439-451): 8 loss scalars (+ log_sigma_x after the update, above)
  if (P.log != nullptr && P.ss != nullptr && i < 8) P.log[((P.ss->step - 1) % P.log_cap) * 9 + i] = P.scalars[i];
}

// Head of a captured step: advance the device-resident step state and stage the step's minibatch indices.
__global__ void __launch_bounds__(1024) advance_kernel(const AdvanceParams P) {
  __shared__ long long s_step;
  if (threadIdx.x == 0) {
    StepState* S = P.ss;
    const long long step = S->step + 1;
    S->step = step;
    for (int k = 0; k < 4; ++k) S->philox_off[k] += P.philox_inc;
    const double bc1 = 1.0 - pow(0.9, (double)step), bc2 = 1.0 - pow(0.999, (double)step);
    for (int g = 0; g < P.n_groups; ++g) S->step_size[g] = (float)((double)P.lr[g] / bc1);
    S->bc2_sqrt = (float)sqrt(bc2);
    s_step = step;
  }
  __syncthreads();
  if (P.idx_pool != nullptr) {
    const long long* src = P.idx_pool + ((s_step - 1) % P.pool_rows) * P.B;
    for (long long i = threadIdx.x; i < P.B; i += blockDim.x) P.idx_cur[i] = src[i];
  }
}

// FactorizedNN head post-processing (models/encoders.py:121-128): loc = clamp(pre, +-50), sigma = exp(clamp(pre, -7, 3)),
// scale_tril = diag(sigma + 1e-8).  headpre rows [row0, row0 + nz) = mean head, [row0 + nz, row0 + 2 nz) = sigma head.
__global__ void __launch_bounds__(256) prior_post_kernel(const float* __restrict__ headpre, long long B, int row0, int nz, int full,
                                                         float* __restrict__ loc, float* __restrict__ tril) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int i = 0; i < nz; ++i) {
    const float pm = headpre[(long long)(row0 + i) * B + b], ps = headpre[(long long)(row0 + nz + i) * B + b];
    loc[b * nz + i] = fminf(fmaxf(pm, -50.0f), 50.0f);
    const float sg = expf(fminf(fmaxf(ps, -7.0f), 3.0f)) + 1e-8f;
    // FullCovarianceNN (--full_cov_prior): strict lower triangle from the clamped f_cov head (models/encoders.py:38-43)
    for (int j = 0; j < nz; ++j) {
      float v = i == j ? sg : 0.0f;
      if (full && j < i) v = fminf(fmaxf(headpre[(long long)(row0 + 2 * nz + i * nz + j) * B + b], -20.0f), 20.0f);
      tril[(b * nz + i) * nz + j] = v;
    }
  }
}

// GaussianEncoder.sample without output transform (models/encoders.py:73-93): z = loc + L eps,
// log q = -1/2 (nz log 2 pi + |eps|^2) - sum log diag L; eps, z (n, B, nz), dens (n, B), L (B, nz, nz) lower-triangular.
__global__ void __launch_bounds__(256) gaussian_sample_kernel(const float* __restrict__ loc, const float* __restrict__ tril,
                                                              const float* __restrict__ eps, int n, long long B, int nz,
                                                              float* __restrict__ z, float* __restrict__ dens) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)n * B) return;
  const long long b = q % B;
  const float* e = eps + q * nz;
  const float* L = tril + b * nz * nz;
  float ss = 0.0f, hld = 0.0f;
  for (int i = 0; i < nz; ++i) {
    float acc = loc[b * nz + i];
    for (int j = 0; j <= i; ++j) acc = fmaf(L[i * nz + j], e[j], acc);
    z[q * nz + i] = acc;
    ss = fmaf(e[i], e[i], ss);
    hld += logf(L[i * nz + i]);
  }
  dens[q] = -0.5f * ((float)nz * 1.8378770664093453f + ss) - hld;
}

// FP32 FFMA peak micro-benchmark: 16 independent accumulator chains per thread.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, float a, float b, int iters) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;  // never true; keeps the chains alive
}

float ffma_peak_tflops(cudaStream_t st) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* d = nullptr;
  cudaMalloc(&d, 4);
  const int iters = 4096, blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 0.0f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, st);
    ffma_peak_kernel<<<blocks, 256, 0, st>>>(d, 0.999f, 0.001f, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 16 * 8 * (double)iters * 256.0 * blocks;
    const float tf = (float)(flop / (ms * 1e-3) / 1e12);
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

void launch_reduce(const ReduceParams& p, cudaStream_t s) {
  const long long n = p.grads ? p.n_params : 1;
  reduce_kernel<<<(unsigned)((n + RED_P - 1) / RED_P), RED_P * RED_G, 0, s>>>(p);
}
void launch_gradnorm(const float* grads, long long n, float max_norm, float* clip_coef, cudaStream_t s) {
  gradnorm_kernel<<<1, 1024, 0, s>>>(grads, n, max_norm, clip_coef);
}
void launch_prior_post(const float* headpre, long long B, int row0, int nz, int full, float* loc, float* tril, cudaStream_t s) {
  prior_post_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(headpre, B, row0, nz, full, loc, tril);
}
void launch_gaussian_sample(const float* loc, const float* tril, const float* eps, int n, long long B, int nz, float* z,
                            float* dens, cudaStream_t s) {
  const long long tot = (long long)n * B;
  gaussian_sample_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(loc, tril, eps, n, B, nz, z, dens);
}
void launch_advance(const AdvanceParams& p, cudaStream_t s) { advance_kernel<<<1, 1024, 0, s>>>(p); }
void launch_adam(const AdamParams& p, cudaStream_t s) {
  adam_kernel<<<(unsigned)((p.n_params + 255) / 256), 256, 0, s>>>(p);
}

}  // namespace dpv
