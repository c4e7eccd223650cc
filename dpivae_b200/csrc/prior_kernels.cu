// Conditional-prior networks p(zc|c), p(zy|y) (FactorizedNN, models/encoders.py:96-128: in -> 64 -> ReLU -> [mean ; sigma]
// heads) as dedicated streaming kernels.  The nets are tiny (2 -> 64 -> 8 in the bridge case: 640 MACs per row), so the
// generic 64-row-tile GEMM kernels of enc_kernels.cu spent their time on staging and on a 64-float-per-row hidden
// record in HBM; here
//   prior_fwd_kernel: one thread per minibatch row, weights in shared memory (broadcast reads), head pre-activations
//                     in registers, written feature-major into `headpre` (coalesced along the rows);
//   prior_bwd_kernel: one thread per (net, hidden unit) walks the CTA's contiguous row range; the hidden activation
//                     is recomputed (no record), the unit's weight-gradient column lives in registers and is written
//                     once per CTA into the per-CTA partial buffer (fixed order -> deterministic reduction).
#include "common.cuh"
#include "kernels.h"

namespace dpv {

namespace {

constexpr int PNT = 256;    // forward: rows per CTA pass
constexpr int PH = 64;      // hidden units per prior net (reference: FactorizedNN(nz, nd, [64]), dpivae.py:156-157)
constexpr int PO_MAX = 16;  // head outputs per net, padded (2 nz <= 16); the kernels are instantiated for PO = 8 and 16: the bridge /
                            // oscillator nets have 8 outputs, and zero-padded outputs cost a third of the backward's FMAs
constexpr int PK = 4;       // inputs per net, padded (nd_c, nd_y <= 4)
constexpr int PTILE = 128;  // backward: rows staged per pass
constexpr int PRG = 4;      // backward: row groups per CTA (thread = (row group, net, hidden unit); rows r = rg mod PRG)

template <int PO>
struct PriorW {   // one net in shared memory
  float w0[PH][PK];
  float b0[PH];
  float w1t[PH][PO];   // heads transposed: [hidden][output]
  float b1[PO];
};

template <int PO>
__device__ __forceinline__ void stage_prior(const EncParams& P, const EncUnit& U, PriorW<PO>& W, int tid, int nthr) {
  for (int e = tid; e < PH * PK; e += nthr) {
    const int k = e / PK, j = e - k * PK;
    W.w0[k][j] = (k < U.H && j < U.K0) ? P.params[U.g_w0 + (long long)k * U.K0 + j] : 0.0f;
  }
  for (int e = tid; e < PH; e += nthr) W.b0[e] = e < U.H ? P.params[U.g_b0 + e] : 0.0f;
  for (int e = tid; e < PH * PO; e += nthr) {
    const int k = e / PO, o = e - k * PO;
    W.w1t[k][o] = (k < U.H && o < U.O) ? P.params[U.g_w1 + (long long)o * U.H + k] : 0.0f;
  }
  for (int e = tid; e < PO; e += nthr) W.b1[e] = e < U.O ? P.params[U.g_b1 + e] : 0.0f;
}

__device__ __forceinline__ void load_ct(const EncParams& P, const EncUnit& U, long long row, float* ct) {
  const float* src = U.src == 1 ? P.c : P.y;
  const float* mean = U.src == 1 ? P.mean_c : P.mean_y;
  const float* sd = U.src == 1 ? P.std_c : P.std_y;
  const long long drow = P.idx ? P.idx[row] : row;
#pragma unroll
  for (int j = 0; j < PK; ++j) ct[j] = j < U.K0 ? (src[drow * U.K0 + j] - mean[j]) / sd[j] : 0.0f;   // utils/transforms.py:70-73
}

}  // namespace

template <int PO>
__global__ void __launch_bounds__(PNT) prior_fwd_kernel(const __grid_constant__ EncParams P) {
  __shared__ PriorW<PO> W;
  for (int u = 0; u < P.n_units; ++u) {
    const EncUnit& U = P.u[u];
    if (U.src == 2 && P.y == nullptr) continue;
    __syncthreads();
    stage_prior(P, U, W, threadIdx.x, PNT);
    __syncthreads();
    // two rows per thread (row, row + stride): every broadcast weight read from shared memory serves both
    const long long stride = (long long)gridDim.x * PNT;
    for (long long row = (long long)blockIdx.x * PNT + threadIdx.x; row < P.B; row += 2 * stride) {
      const long long row2 = row + stride;
      const bool has2 = row2 < P.B;
      float ct[PK], ct2[PK] = {0.f, 0.f, 0.f, 0.f};
      load_ct(P, U, row, ct);
      if (has2) load_ct(P, U, row2, ct2);
      float out[PO], out2[PO];
#pragma unroll
      for (int o = 0; o < PO; ++o) out[o] = out2[o] = W.b1[o];
#pragma unroll 4
      for (int k = 0; k < PH; ++k) {
        const float4 w = *reinterpret_cast<const float4*>(W.w0[k]);
        const float bk = W.b0[k];
        const float h = fmaxf(fmaf(ct[3], w.w, fmaf(ct[2], w.z, fmaf(ct[1], w.y, fmaf(ct[0], w.x, bk)))), 0.0f);
        const float h2 = fmaxf(fmaf(ct2[3], w.w, fmaf(ct2[2], w.z, fmaf(ct2[1], w.y, fmaf(ct2[0], w.x, bk)))), 0.0f);
        const float4* t = reinterpret_cast<const float4*>(W.w1t[k]);
#pragma unroll
        for (int q4 = 0; q4 < PO / 4; ++q4) {
          const float4 tv = t[q4];
          out[4 * q4] = fmaf(h, tv.x, out[4 * q4]); out[4 * q4 + 1] = fmaf(h, tv.y, out[4 * q4 + 1]);
          out[4 * q4 + 2] = fmaf(h, tv.z, out[4 * q4 + 2]); out[4 * q4 + 3] = fmaf(h, tv.w, out[4 * q4 + 3]);
          out2[4 * q4] = fmaf(h2, tv.x, out2[4 * q4]); out2[4 * q4 + 1] = fmaf(h2, tv.y, out2[4 * q4 + 1]);
          out2[4 * q4 + 2] = fmaf(h2, tv.z, out2[4 * q4 + 2]); out2[4 * q4 + 3] = fmaf(h2, tv.w, out2[4 * q4 + 3]);
        }
      }
#pragma unroll
      for (int o = 0; o < PO; ++o)
        if (o < U.O) {
          P.headpre[(long long)(U.out_row + o) * P.B + row] = out[o];
          if (has2) P.headpre[(long long)(U.out_row + o) * P.B + row2] = out2[o];
        }
    }
  }
  pdl_wait();   // no-op unless launched with overlap = true (see launch_enc_fwd in kernels.h)
}

// blockDim = PRG * 64 * n_units: thread (row group rg, net u, hidden unit k)
template <int PO>
__global__ void __launch_bounds__(PRG * 2 * PH) prior_bwd_kernel(const __grid_constant__ EncParams P) {
  __shared__ __align__(16) float CT[2][PTILE][PK];
  __shared__ __align__(16) float G[2][PTILE][PO];
  const int tid = threadIdx.x, nuk = PH * P.n_units, rg = tid / nuk, uk = tid - rg * nuk, u = uk / PH, k = uk - u * PH;
  const EncUnit& U = P.u[u];
  float* part = P.part + (long long)blockIdx.x * P.part_stride;
  // this thread's weights: first-layer row k, head column k
  float w0[PK], w1[PO];
#pragma unroll
  for (int j = 0; j < PK; ++j) w0[j] = (k < U.H && j < U.K0) ? P.params[U.g_w0 + (long long)k * U.K0 + j] : 0.0f;
  const float b0 = k < U.H ? P.params[U.g_b0 + k] : 0.0f;
#pragma unroll
  for (int o = 0; o < PO; ++o) w1[o] = (k < U.H && o < U.O) ? P.params[U.g_w1 + (long long)o * U.H + k] : 0.0f;
  float a_w1[PO], a_w0[PK], a_b0 = 0.0f, a_b1 = 0.0f;
#pragma unroll
  for (int o = 0; o < PO; ++o) a_w1[o] = 0.0f;
#pragma unroll
  for (int j = 0; j < PK; ++j) a_w0[j] = 0.0f;
  // contiguous row range of this CTA
  const long long per = (P.B + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = min(P.B, r0 + per);
  // staging registers: the next tile's standardised inputs and head gradients are loaded while the current tile is
  // processed (blockDim = PRG * 64 * n_units threads cover the 2 * PTILE input rows and 2 * PO * PTILE gradients)
  constexpr int NG = (2 * PO * PTILE) / (PRG * 2 * PH);   // gradient elements per thread (8)
  float pct[PK] = {0.f, 0.f, 0.f, 0.f}, pg[NG];
  auto fetch = [&](long long t0) {
    const int nr = (int)min((long long)PTILE, r1 - t0);
    if (tid < P.n_units * PTILE) {
      const int uu = tid / PTILE, r = tid - uu * PTILE;
#pragma unroll
      for (int j = 0; j < PK; ++j) pct[j] = 0.0f;
      if (t0 < r1 && r < nr) load_ct(P, P.u[uu], t0 + r, pct);
    }
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      const int e = tid + i * (int)blockDim.x;
      const int r = e % PTILE, o = (e / PTILE) % PO, uu = e / (PTILE * PO);
      pg[i] = (t0 < r1 && uu < P.n_units && r < nr && o < P.u[uu].O) ? P.gpre[(long long)(P.u[uu].out_row + o) * P.B + t0 + r] : 0.0f;
    }
  };
  fetch(r0);
  for (long long t0 = r0; t0 < r1; t0 += PTILE) {
    const int nr = (int)min((long long)PTILE, r1 - t0);
    __syncthreads();
    if (tid < P.n_units * PTILE) {
      const int uu = tid / PTILE, r = tid - uu * PTILE;
#pragma unroll
      for (int j = 0; j < PK; ++j) CT[uu][r][j] = pct[j];
    }
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      const int e = tid + i * (int)blockDim.x;
      const int r = e % PTILE, o = (e / PTILE) % PO, uu = e / (PTILE * PO);
      if (uu < 2) G[uu][r][o] = pg[i];
    }
    __syncthreads();
    fetch(t0 + PTILE);
    for (int r = rg; r < nr; r += PRG) {
      const float4 c4 = *reinterpret_cast<const float4*>(CT[u][r]);
      const float pre = fmaf(c4.w, w0[3], fmaf(c4.z, w0[2], fmaf(c4.y, w0[1], fmaf(c4.x, w0[0], b0))));
      const float h = fmaxf(pre, 0.0f);
      float gq[PO / 4];   // four independent partial sums: no 16-deep dependent FMA chain per row
      const float4* g4 = reinterpret_cast<const float4*>(G[u][r]);
#pragma unroll
      for (int q4 = 0; q4 < PO / 4; ++q4) {
        const float4 g = g4[q4];
        gq[q4] = fmaf(g.w, w1[4 * q4 + 3], fmaf(g.z, w1[4 * q4 + 2], fmaf(g.y, w1[4 * q4 + 1], g.x * w1[4 * q4])));
        a_w1[4 * q4] = fmaf(g.x, h, a_w1[4 * q4]); a_w1[4 * q4 + 1] = fmaf(g.y, h, a_w1[4 * q4 + 1]);
        a_w1[4 * q4 + 2] = fmaf(g.z, h, a_w1[4 * q4 + 2]); a_w1[4 * q4 + 3] = fmaf(g.w, h, a_w1[4 * q4 + 3]);
      }
      float gsum;
      if constexpr (PO == 16) gsum = (gq[0] + gq[1]) + (gq[2] + gq[3]);
      else gsum = gq[0] + gq[1];
      const float gh = pre > 0.0f ? gsum : 0.0f;
      a_w0[0] = fmaf(gh, c4.x, a_w0[0]); a_w0[1] = fmaf(gh, c4.y, a_w0[1]); a_w0[2] = fmaf(gh, c4.z, a_w0[2]); a_w0[3] = fmaf(gh, c4.w, a_w0[3]);
      a_b0 += gh;
      if (k < PO) a_b1 += G[u][r][k];   // head-bias gradient of output k
    }
  }
  // fixed-order sum over the row groups (one group per round through the staging buffer), then one write per CTA
  float* RX = &G[0][0][0];   // [2 * PH][PO + PK + 2] <= 2 * PTILE * PO floats
  constexpr int RS = PO + PK + 2;
  static_assert(2 * PH * RS <= 2 * PTILE * PO, "row-group exchange does not fit the staging buffer");
  for (int g = 1; g < PRG; ++g) {
    __syncthreads();
    if (rg == g) {
      float* mine = RX + (size_t)uk * RS;
#pragma unroll
      for (int o = 0; o < PO; ++o) mine[o] = a_w1[o];
#pragma unroll
      for (int j = 0; j < PK; ++j) mine[PO + j] = a_w0[j];
      mine[PO + PK] = a_b0;
      mine[PO + PK + 1] = a_b1;
    }
    __syncthreads();
    if (rg == 0) {
      const float* o_ = RX + (size_t)uk * RS;
#pragma unroll
      for (int o = 0; o < PO; ++o) a_w1[o] += o_[o];
#pragma unroll
      for (int j = 0; j < PK; ++j) a_w0[j] += o_[PO + j];
      a_b0 += o_[PO + PK];
      a_b1 += o_[PO + PK + 1];
    }
  }
  if (rg == 0) {
    if (k < U.H) {
#pragma unroll
      for (int o = 0; o < PO; ++o)
        if (o < U.O) part[U.g_w1 + (long long)o * U.H + k] = a_w1[o];
#pragma unroll
      for (int j = 0; j < PK; ++j)
        if (j < U.K0) part[U.g_w0 + (long long)k * U.K0 + j] = a_w0[j];
      part[U.g_b0 + k] = a_b0;
    }
    if (k < U.O) part[U.g_b1 + k] = a_b1;
  }
  pdl_wait();   // no-op unless launched with overlap = true (see launch_enc_fwd in kernels.h)
}

bool prior_kernels_support(const EncParams& p) {
  if (p.n_units < 1 || p.n_units > 2) return false;
  for (int u = 0; u < p.n_units; ++u)
    if (p.u[u].H > PH || p.u[u].O > PO_MAX || p.u[u].K0 > PK || p.u[u].src == 0) return false;
  return true;
}

static bool prior_small_heads(const EncParams& p) {
  for (int u = 0; u < p.n_units; ++u)
    if (p.u[u].O > 8) return false;
  return true;
}

void launch_prior_fwd(const EncParams& p, int sm_count, cudaStream_t s, bool overlap) {
  long long g = (p.B + 2 * PNT - 1) / (2 * PNT);   // two rows per thread
  if (g > 4LL * sm_count) g = 4LL * sm_count;
  if (prior_small_heads(p)) {
    if (overlap) launch_pdl(prior_fwd_kernel<8>, (int)(g < 1 ? 1 : g), PNT, 0, s, p);
    else prior_fwd_kernel<8><<<(unsigned)(g < 1 ? 1 : g), PNT, 0, s>>>(p);
  } else {
    if (overlap) launch_pdl(prior_fwd_kernel<16>, (int)(g < 1 ? 1 : g), PNT, 0, s, p);
    else prior_fwd_kernel<16><<<(unsigned)(g < 1 ? 1 : g), PNT, 0, s>>>(p);
  }
}

void launch_prior_bwd(const EncParams& p, int grid, cudaStream_t s, bool overlap) {
  if (prior_small_heads(p)) {
    if (overlap) launch_pdl(prior_bwd_kernel<8>, grid, PRG * PH * p.n_units, 0, s, p);
    else prior_bwd_kernel<8><<<grid, PRG * PH * p.n_units, 0, s>>>(p);
  } else {
    if (overlap) launch_pdl(prior_bwd_kernel<16>, grid, PRG * PH * p.n_units, 0, s, p);
    else prior_bwd_kernel<16><<<grid, PRG * PH * p.n_units, 0, s>>>(p);
  }
}

}  // namespace dpv
