// Encoder / prior-net kernels (sm_100a): forward and backward of the one-hidden-layer ReLU MLPs
// whose heads parameterise q(z|x) (FullCovarianceNN, models/encoders.py:7-44) and the conditional
// priors p(zc|c), p(zy|y) (FactorizedNN, models/encoders.py:96-128).  Work unit: 64 minibatch
// rows per tile, weights of ONE unit resident in shared memory while a persistent CTA walks its
// tiles (unit-outer / tile-inner, so each CTA stages every weight matrix exactly once).
//
// Forward writes hidden activations `hid` [H_tot][B] and head pre-activations `headpre`
// [O_tot][B] feature-major (coalesced along rows); the decoder kernel consumes headpre and
// returns gpre = d loss / d headpre, from which the backward kernel forms wgrads (per-CTA
// partials) -- inputs are data, so no dgrad to x is needed.
#include "common.cuh"
#include "kernels.h"

namespace dpv {

struct EncSmem {
  int w0t, b0, w1t, b1, in, hid, out, total;
  int ldw0, ldw1;
};

__host__ __device__ inline EncSmem enc_plan(int K0max, int Hmax, int Omax) {
  EncSmem s;
  const int K0p = pad4(K0max), Op = pad4(Omax);
  s.ldw0 = Hmax + 4;
  s.ldw1 = Op + 4;
  int o = 0;
  s.w0t = o; o += K0p * s.ldw0;
  s.b0 = o; o += Hmax;
  s.w1t = o; o += Hmax * s.ldw1;
  s.b1 = o; o += Op;
  s.in = o; o += K0p * LDP;
  s.hid = o; o += Hmax * LDP;
  s.out = o; o += Op * LDP;
  s.total = o;
  return s;
}

__host__ __device__ inline void enc_max_dims(const EncParams& P, int& K0, int& H, int& O) {
  K0 = H = O = 0;
  for (int u = 0; u < P.n_units; ++u) {
    K0 = P.u[u].K0 > K0 ? P.u[u].K0 : K0;
    H = P.u[u].H > H ? P.u[u].H : H;
    O = P.u[u].O > O ? P.u[u].O : O;
  }
}

// standardised input tile IN[k][r] = (v - mean) / std     (utils/transforms.py:70-73)
__device__ __forceinline__ void load_input_tile(const EncParams& P, const EncUnit& U, long long row0, float* IN) {
  const float* src = U.src == 0 ? P.x : (U.src == 1 ? P.c : P.y);
  const float* mean = U.src == 0 ? P.mean_x : (U.src == 1 ? P.mean_c : P.mean_y);
  const float* sd = U.src == 0 ? P.std_x : (U.src == 1 ? P.std_c : P.std_y);
  const bool raw = (U.src == 0 && P.x_is_standardised);
  const int K0 = U.K0;
  for (int e = threadIdx.x; e < TILE * K0; e += NT) {
    const int r = e / K0, k = e - r * K0;
    const long long lrow = row0 + r;
    float v = 0.0f;
    if (lrow < P.B) {
      const long long drow = P.idx ? P.idx[lrow] : lrow;
      v = src[drow * K0 + k];
      if (!raw) v = (v - mean[k]) / sd[k];
    }
    IN[k * LDP + r] = v;
  }
}

__global__ void __launch_bounds__(NT, 2) enc_fwd_kernel(const __grid_constant__ EncParams P) {
  extern __shared__ __align__(16) float sm[];
  int K0m, Hm, Om;
  enc_max_dims(P, K0m, Hm, Om);
  const EncSmem S = enc_plan(K0m, Hm, Om);
  const long long ntiles = (P.B + TILE - 1) / TILE;
  for (int u = blockIdx.y; u < P.n_units; u += gridDim.y) {   // small batches: one CTA column per unit
    const EncUnit& U = P.u[u];
    if (U.src == 2 && P.y == nullptr) continue;
    __syncthreads();
    for (int e = threadIdx.x; e < S.total; e += NT) sm[e] = 0.0f;
    __syncthreads();
    stage_linear(P.params + U.g_w0, P.params + U.g_b0, U.K0, U.H, sm + S.w0t, S.ldw0, sm + S.b0);
    stage_linear(P.params + U.g_w1, P.params + U.g_b1, U.H, U.O, sm + S.w1t, S.ldw1, sm + S.b1);
    __syncthreads();
    const int Op = pad4(U.O);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = tile * TILE;
      load_input_tile(P, U, row0, sm + S.in);
      __syncthreads();
      gemm_fwd<ACT_RELU>(sm + S.w0t, S.ldw0, sm + S.b0, sm + S.in, sm + S.hid, U.K0, U.H);
      __syncthreads();
      gemm_fwd<ACT_NONE>(sm + S.w1t, S.ldw1, sm + S.b1, sm + S.hid, sm + S.out, U.H, Op);
      if (P.with_hid) {
        for (int e = threadIdx.x; e < TILE * U.H; e += NT) {
          const int r = e & (TILE - 1), f = e >> 6;
          if (row0 + r < P.B) P.hid[(long long)(U.hid_row + f) * P.B + row0 + r] = sm[S.hid + f * LDP + r];
        }
      }
      __syncthreads();
      for (int e = threadIdx.x; e < TILE * U.O; e += NT) {
        const int r = e & (TILE - 1), f = e >> 6;
        if (row0 + r < P.B) P.headpre[(long long)(U.out_row + f) * P.B + row0 + r] = sm[S.out + f * LDP + r];
      }
      __syncthreads();
    }
  }
  pdl_wait();   // no-op unless launched with overlap = true (see launch_enc_fwd in kernels.h)
}

__global__ void __launch_bounds__(NT, 2) enc_bwd_kernel(const __grid_constant__ EncParams P) {
  extern __shared__ __align__(16) float sm[];
  int K0m, Hm, Om;
  enc_max_dims(P, K0m, Hm, Om);
  const EncSmem S = enc_plan(K0m, Hm, Om);
  const long long ntiles = (P.B + TILE - 1) / TILE;
  float* part = P.part + (long long)blockIdx.x * P.part_stride;
  for (int u = blockIdx.y; u < P.n_units; u += gridDim.y) {   // small batches: one CTA column per unit
    const EncUnit& U = P.u[u];
    __syncthreads();
    for (int e = threadIdx.x; e < S.total; e += NT) sm[e] = 0.0f;
    for (long long e = threadIdx.x; e < (long long)U.K0 * U.H; e += NT) part[U.g_w0 + e] = 0.0f;
    for (int e = threadIdx.x; e < U.H; e += NT) part[U.g_b0 + e] = 0.0f;
    for (long long e = threadIdx.x; e < (long long)U.H * U.O; e += NT) part[U.g_w1 + e] = 0.0f;
    for (int e = threadIdx.x; e < U.O; e += NT) part[U.g_b1 + e] = 0.0f;
    __syncthreads();
    // only w1 is needed (dgrad to the hidden layer); inputs are data
    stage_linear(P.params + U.g_w1, P.params + U.g_b1, U.H, U.O, sm + S.w1t, S.ldw1, sm + S.b1);
    __syncthreads();
    const int Op = pad4(U.O);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = tile * TILE;
      load_input_tile(P, U, row0, sm + S.in);
      for (int e = threadIdx.x; e < TILE * U.H; e += NT) {
        const int r = e & (TILE - 1), f = e >> 6;
        sm[S.hid + f * LDP + r] = row0 + r < P.B ? P.hid[(long long)(U.hid_row + f) * P.B + row0 + r] : 0.0f;
      }
      for (int e = threadIdx.x; e < TILE * U.O; e += NT) {
        const int r = e & (TILE - 1), f = e >> 6;
        sm[S.out + f * LDP + r] = row0 + r < P.B ? P.gpre[(long long)(U.out_row + f) * P.B + row0 + r] : 0.0f;
      }
      __syncthreads();
      gemm_wgrad(sm + S.hid, sm + S.out, part + U.g_w1, part + U.g_b1, U.H, U.O);
      __syncthreads();
      gemm_dgrad<ACT_RELU>(sm + S.w1t, S.ldw1, sm + S.out, sm + S.hid, sm + S.hid, U.H, Op);
      __syncthreads();
      gemm_wgrad(sm + S.in, sm + S.hid, part + U.g_w0, part + U.g_b0, U.K0, U.H);
      __syncthreads();
    }
  }
  pdl_wait();   // no-op unless launched with overlap = true (see launch_enc_fwd in kernels.h)
}

size_t enc_smem_bytes(const EncParams& p, bool bwd) {
  (void)bwd;
  int K0m, Hm, Om;
  enc_max_dims(p, K0m, Hm, Om);
  return (size_t)enc_plan(K0m, Hm, Om).total * sizeof(float);
}

// Few tiles (small minibatches): the units run in parallel CTA columns (gridDim.y) instead of one after the other
// in every CTA; many tiles: unit-outer loop in each persistent CTA (every weight matrix staged once per CTA).
static dim3 enc_grid(const EncParams& p, int grid) {
  return dim3((unsigned)grid, (unsigned)((long long)grid * p.n_units <= 2LL * 148 ? p.n_units : 1), 1);
}
void launch_enc_fwd(const EncParams& p, int grid, size_t smem, cudaStream_t s, bool overlap) {
  if (overlap) launch_pdl3(enc_fwd_kernel, enc_grid(p, grid), NT, smem, s, p);
  else enc_fwd_kernel<<<enc_grid(p, grid), NT, smem, s>>>(p);
}
void launch_enc_bwd(const EncParams& p, int grid, size_t smem, cudaStream_t s, bool overlap) {
  if (overlap) launch_pdl3(enc_bwd_kernel, enc_grid(p, grid), NT, smem, s, p);
  else enc_bwd_kernel<<<enc_grid(p, grid), NT, smem, s>>>(p);
}

int configure_enc_kernels() {
  int e = (int)cudaFuncSetAttribute(enc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e) return e;
  return (int)cudaFuncSetAttribute(enc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
}

}  // namespace dpv
