// Per-pair latent kernels of the tensor-core path (sm_100a): everything between the encoder heads and the
// decoders that is scalar work per (row, MC-sample) pair, kept OUT of the tensor-core decoder kernel so that it
// runs at full occupancy instead of on the latency-bound epilogue warps:
//
//   forward  : head post-processing (clamp / exp / tril, models/encoders.py:35-43), reparameterised sample
//              z = loc + L eps (models/encoders.py:73-93), logistic + shift/scale bijector and its log-det
//              (utils/transforms.py:97-150), log q, log p(zx) (utils/priors.py:19-23), log p(zc|c), log p(zy|y)
//              (models/vae.py:200-207) -> per-row KL; writes, per 128-pair tile, the decoder kernel's input
//              RECORD: the latent operand [zd | 1 | physics input] already split into fp16 hi/lo planes in the
//              X8 layout of tc.cuh (the decoder kernel bulk-copies it straight into its operand buffer) plus
//              the raw c / y values per pair.
//   backward : from the decoder kernel's per-pair dL/dz record: gradients w.r.t. loc / L / prior-net heads,
//              reduced over the MC axis, chain rule through clamp / exp -> gpre (input of the encoder backward).
//
// Two implementations of each, same tiling as dec_tc_kernel (one tile = 128 pairs = RB rows x n_mc samples):
//   lat_pair_fwd_kernel / lat_pair_bwd_kernel (second half of this file): thread-per-pair kernels for the compile-time
//     shapes of the reference's six case x model presets at n_mc = 16, with the noise drawn ahead by
//     lat_noise_fill_kernel / lat_noise_fill_cyclic_kernel (one Philox evaluation per four elements) -- the path of
//     the benchmarked configurations;
//   lat_fwd_kernel / lat_bwd_kernel (first half): run-time shapes (any case / preset / MC count), 256 threads per tile,
//     intermediates in shared-memory planes, noise per element, tile-ordered noise record for the backward.
//   lat_encode_kernel: encode-only inference of the fp32 mode (one thread per pair).
#include <cuda_fp16.h>
#include <curand_kernel.h>

#include "common.cuh"
#include "kernels.h"
#include "tc.cuh"

namespace dpv {

namespace {

constexpr int TP = 128;
constexpr int LNT = 256;


// Compile-time shapes.  The kernels below are written against run-time dimensions (any case / preset) through the
// accessor class D; for the shapes of the reference's cases they are ALSO instantiated with the dimensions known to the
// compiler, so that the block / triangular-index loops unroll, the index arithmetic folds and the divisions by n_mc
// become shifts (85 % of the generic kernel's executed instructions were integer / control, ncu instruction mix).
//   MT: 0 = P (three latent blocks), 1 = S (one block), -1 = generic (every accessor reads the descriptor)
constexpr int tri(int n) { return n * (n + 1) / 2; }
template <int MT, int NX, int NC, int NY, int NDC, int NDY, int NDP, int NMC>
struct Shape {
  static constexpr bool K = MT >= 0;
  static constexpr int cZ = NX + NC + NY, cnzd = NC + NY, cnb = MT == 0 ? 3 : 1;
  static constexpr int cnL = MT == 0 ? tri(NX) + tri(NC) + tri(NY) : tri(cZ);
  static constexpr int cNX = NX, cNC = NC, cNY = NY, cndc = NDC, cndy = NDY, cndp = NDP, cn_mc = NMC;
  static constexpr int cn_rowpar = cZ + cnL + 2 * cnzd;   // == n_feat: row parameters and per-pair gradient features share one indexing
#define DPV_SCALAR(name, val) \
  static __device__ __forceinline__ int name(const DecParams& P) { if constexpr (K) return (val); else return P.name; }
  DPV_SCALAR(model_type, MT) DPV_SCALAR(nz_x, NX) DPV_SCALAR(nz_c, NC) DPV_SCALAR(nz_y, NY) DPV_SCALAR(Z, cZ)
  DPV_SCALAR(nd_c, NDC) DPV_SCALAR(nd_y, NDY) DPV_SCALAR(nd_p, NDP) DPV_SCALAR(n_mc, NMC) DPV_SCALAR(RB, 128 / (NMC > 0 ? NMC : 1))
  DPV_SCALAR(nL, cnL) DPV_SCALAR(n_blk, cnb)
  DPV_SCALAR(rp_loc, 0) DPV_SCALAR(rp_L, cZ) DPV_SCALAR(rp_pmu, cZ + cnL) DPV_SCALAR(rp_psig, cZ + cnL + cnzd) DPV_SCALAR(n_rowpar, cZ + cnL + 2 * cnzd)
  DPV_SCALAR(f_loc, 0) DPV_SCALAR(f_L, cZ) DPV_SCALAR(f_pmu, cZ + cnL) DPV_SCALAR(f_psig, cZ + cnL + cnzd) DPV_SCALAR(n_feat, cZ + cnL + 2 * cnzd)
#undef DPV_SCALAR
  static __device__ __forceinline__ int blk_size(const DecParams& P, int b) {
    if constexpr (!K) return P.blk_size[b];
    else if constexpr (MT == 1) return cZ;
    else return b == 0 ? NX : (b == 1 ? NC : NY);
  }
  static __device__ __forceinline__ int blk_start(const DecParams& P, int b) {
    if constexpr (!K) return P.blk_start[b];
    else if constexpr (MT == 1) return 0;
    else return b == 0 ? 0 : (b == 1 ? NX : NX + NC);
  }
  static __device__ __forceinline__ int blk_loff(const DecParams& P, int b) {
    if constexpr (!K) return P.blk_loff[b];
    else if constexpr (MT == 1) return 0;
    else return b == 0 ? 0 : (b == 1 ? tri(NX) : tri(NX) + tri(NC));
  }
  static __device__ __forceinline__ int henc(const DecParams& P, int b) {   // head rows: [mean nz | sigma nz | cov nz*nz] per block
    if constexpr (!K) return P.henc[b];
    else if constexpr (MT == 1) return 0;
    else return b == 0 ? 0 : (b == 1 ? 2 * NX + NX * NX : 2 * NX + NX * NX + 2 * NC + NC * NC);
  }
  static __device__ __forceinline__ int hpri(const DecParams& P, int k) {
    if constexpr (!K) return P.hpri[k];
    else {
      constexpr int e = MT == 1 ? 2 * cZ + cZ * cZ : 2 * NX + NX * NX + 2 * NC + NC * NC + 2 * NY + NY * NY;
      return k == 0 ? e : e + 2 * NC;
    }
  }
  // packed lower-triangular tables stay descriptor reads (their index is a run-time value in the row-parameter loops)
  static __device__ __forceinline__ int L_blk(const DecParams& P, int li) { return P.L_blk[li]; }
  static __device__ __forceinline__ int L_i(const DecParams& P, int li) { return P.L_i[li]; }
  static __device__ __forceinline__ int L_j(const DecParams& P, int li) { return P.L_j[li]; }
};
using ShGeneric = Shape<-1, 0, 0, 0, 0, 0, 0, 0>;

template <class D>
__device__ __forceinline__ int block_of_l(const DecParams& P, int i) {
  int b = 0;
  while (b + 1 < D::n_blk(P) && i >= D::blk_start(P, b + 1)) ++b;
  return b;
}

// shared-memory carve-up (floats), identical for both kernels
struct LatSmem {
  float *ROWPAR, *ROWRAW, *ROWLOG, *EPS, *U, *ZXIN, *ZD, *DZ, *SC, *FEAT, *ROWACC, *ROWMSK, *EPS2, *DZ2, *BAR, *RAWH, *ROWAUX;
};
__host__ __device__ inline int lat_smem_floats(const DecParams& P, bool bwd) {
  const int nzd = P.nz_c + P.nz_y, nzin = P.nz_x + P.nd_p;
  int f = P.n_rowpar * RBMAX + (P.nd_c + P.nd_y) * RBMAX + 5 * RBMAX + P.Z * TP + P.nz_x * TP + nzin * TP + nzd * TP +
          (nzd + P.nz_x) * TP + 4 * TP;
  if (bwd) f += P.n_feat * TP + P.n_feat * RBMAX + P.n_rowpar * RBMAX + P.Z * TP + (nzd + P.nz_x) * TP + 4 + 2 * P.O_tot * RBMAX;
  return f;
}
__device__ inline LatSmem lat_carve(float* sm, const DecParams& P, bool bwd) {
  const int nzd = P.nz_c + P.nz_y, nzin = P.nz_x + P.nd_p;
  LatSmem S;
  S.ROWPAR = sm; sm += P.n_rowpar * RBMAX;
  S.ROWRAW = sm; sm += (P.nd_c + P.nd_y) * RBMAX;
  S.ROWLOG = sm; sm += 5 * RBMAX;
  S.EPS = sm; sm += P.Z * TP;
  S.U = sm; sm += P.nz_x * TP;
  S.ZXIN = sm; sm += nzin * TP;
  S.ZD = sm; sm += nzd * TP;
  S.DZ = sm; sm += (nzd + P.nz_x) * TP;
  S.SC = sm; sm += 4 * TP;
  S.FEAT = sm; sm += bwd ? P.n_feat * TP : 0;
  S.ROWACC = sm; sm += bwd ? P.n_feat * RBMAX : 0;
  S.ROWMSK = sm; sm += bwd ? P.n_rowpar * RBMAX : 0;   // backward: chain-rule factors of the head clamps / exps, same row layout as ROWPAR
  S.EPS2 = sm; sm += bwd ? P.Z * TP : 0;               // backward: second buffers of the prefetched noise / dL/dz records
  S.DZ2 = sm; sm += bwd ? (nzd + P.nz_x) * TP : 0;
  S.BAR = sm; sm += 4;                                 // two mbarriers (16 bytes)
  S.RAWH = sm;                                         // backward: 2 x [O_tot][RBMAX] raw head pre-activations (cp.async)
  return S;
}

// per-row parameters of q(z|x) and of the conditional priors, raw c / y
template <class D, int NTH = LNT>
__device__ __forceinline__ void load_row_params(const DecParams& P, const LatSmem& S, long long row0, int nrows, bool msk = false,
                                                const float* rawh = nullptr) {
  const int tid = threadIdx.x, RB = D::RB(P), nzd = D::nz_c(P) + D::nz_y(P);
  const long long B = P.B;
  // head pre-activation of feature row f for tile row r: from the staged tile (backward, prefetched) or from global
  auto hp = [&](int f, int r, long long lrow) -> float { return rawh ? rawh[f * RBMAX + r] : P.headpre[(long long)f * B + lrow]; };
  for (int e = tid; e < RB * D::Z(P); e += NTH) {
    const int i = e / RB, r = e - i * RB;
    const long long lrow = row0 + min(r, nrows - 1);
    const int b = block_of_l<D>(P, i), il = i - D::blk_start(P, b);
    const float pm = hp(D::henc(P, b) + il, r, lrow);
    S.ROWPAR[(D::rp_loc(P) + i) * RBMAX + r] = clampf_(pm, -50.0f, 50.0f);
    // d clamp / d pre: 1 inside the clamp range (models/encoders.py:35-43); d exp(clamp(ps)) / d ps = exp(ps) inside
    if (msk) S.ROWMSK[(D::rp_loc(P) + i) * RBMAX + r] = (pm >= -50.0f && pm <= 50.0f) ? 1.0f : 0.0f;
  }
  for (int e = tid; e < RB * D::nL(P); e += NTH) {
    const int li = e / RB, r = e - li * RB;
    const long long lrow = row0 + min(r, nrows - 1);
    const int b = D::L_blk(P, li), i = D::L_i(P, li), j = D::L_j(P, li), nzb = D::blk_size(P, b);
    float v, mk;
    if (i == j) {
      const float ps = hp(D::henc(P, b) + nzb + i, r, lrow);
      v = expf(clampf_(ps, -7.0f, 3.0f)) + 1e-8f;
      mk = (ps >= -7.0f && ps <= 3.0f) ? expf(ps) : 0.0f;
    } else {
      const float pc = hp(D::henc(P, b) + 2 * nzb + i * nzb + j, r, lrow);
      v = clampf_(pc, -20.0f, 20.0f);
      mk = (pc >= -20.0f && pc <= 20.0f) ? 1.0f : 0.0f;
    }
    S.ROWPAR[(D::rp_L(P) + li) * RBMAX + r] = v;
    if (msk) S.ROWMSK[(D::rp_L(P) + li) * RBMAX + r] = mk;
  }
  for (int e = tid; e < RB * nzd; e += NTH) {
    const int k = e / RB, r = e - k * RB;
    const long long lrow = row0 + min(r, nrows - 1);
    const int which = k < D::nz_c(P) ? 0 : 1;
    const int kk = which ? k - D::nz_c(P) : k, nzk = which ? D::nz_y(P) : D::nz_c(P);
    float mu = 0.0f, sgm = 1.0f, mkm = 0.0f, mks = 0.0f;
    if (which == 0 || P.y != nullptr) {
      const float pm = hp(D::hpri(P, which) + kk, r, lrow);
      const float ps = hp(D::hpri(P, which) + nzk + kk, r, lrow);
      mu = clampf_(pm, -50.0f, 50.0f);
      sgm = expf(clampf_(ps, -7.0f, 3.0f)) + 1e-8f;
      mkm = (pm >= -50.0f && pm <= 50.0f) ? 1.0f : 0.0f;
      mks = (ps >= -7.0f && ps <= 3.0f) ? expf(ps) : 0.0f;
    }
    S.ROWPAR[(D::rp_pmu(P) + k) * RBMAX + r] = mu;
    S.ROWPAR[(D::rp_psig(P) + k) * RBMAX + r] = sgm;
    if (msk) {
      S.ROWMSK[(D::rp_pmu(P) + k) * RBMAX + r] = mkm;
      S.ROWMSK[(D::rp_psig(P) + k) * RBMAX + r] = mks;
    }
  }
  if (!msk)   // raw c / y are only used by the forward (decoder record)
  for (int e = tid; e < RB * (D::nd_c(P) + D::nd_y(P)); e += NTH) {
    const int j = e / RB, r = e - j * RB;
    const long long lrow = row0 + min(r, nrows - 1);
    const long long drow = P.idx ? P.idx[lrow] : lrow;
    float v = 0.0f;
    if (j < D::nd_c(P)) v = P.c[drow * D::nd_c(P) + j];
    else if (P.y != nullptr) v = P.y[drow * D::nd_y(P) + (j - D::nd_c(P))];
    S.ROWRAW[j * RBMAX + r] = v;
  }
}

// z = loc + L eps for the latent blocks b with (b & 1) == hh; returns this thread's share of
// log q - bijector log-det - log p(zx)
template <class D>
__device__ __forceinline__ float sample_blocks(const DecParams& P, const LatSmem& S, int hh, int p, int prow, bool with_logs,
                                               float& dens_share) {
  float lq_part = 0.0f, ld1 = 0.0f, ld2 = 0.0f, lpx = 0.0f;
  for (int b = hh; b < D::n_blk(P); b += 2) {
    const int s = D::blk_start(P, b), nzb = D::blk_size(P, b);
    float ss = 0.0f;
    for (int i = 0; i < nzb; ++i) {
      float acc = S.ROWPAR[(D::rp_loc(P) + s + i) * RBMAX + prow];
      const int base = D::rp_L(P) + D::blk_loff(P, b) + i * (i + 1) / 2;
      for (int j = 0; j <= i; ++j) acc = fmaf(S.ROWPAR[(base + j) * RBMAX + prow], S.EPS[(s + j) * TP + p], acc);
      const float e = S.EPS[(s + i) * TP + p];
      ss = fmaf(e, e, ss);
      const int gi = s + i;
      if (gi < D::nz_x(P)) {
        const float u = sigmoidf_(acc);
        const float a = P.ub[gi] - P.lb[gi];
        const float zx = fmaf(u, a, P.lb[gi]);
        S.U[gi * TP + p] = u;
        S.ZXIN[gi * TP + p] = zx;
        if (with_logs) {
          ld1 += acc - 2.0f * softplusf_(acc);
          ld2 += logf(fabsf(a));
          if (P.prior_kind[gi] == 0) {
            const bool inside = (zx >= P.prior_a[gi]) && (zx < P.prior_b[gi]);
            lpx += (inside ? 0.0f : -INFINITY) - logf(P.prior_b[gi] - P.prior_a[gi]);
          } else {
            const float d = zx - P.prior_a[gi];
            lpx += -(d * d) / (2.0f * P.prior_b[gi] * P.prior_b[gi]) - logf(P.prior_b[gi]) - LOG_SQRT_2PI;
          }
        }
      } else {
        S.ZD[(gi - D::nz_x(P)) * TP + p] = acc;
      }
    }
    if (with_logs) lq_part += -0.5f * ((float)nzb * LOG_2PI + ss) - S.ROWLOG[b * RBMAX + prow];
  }
  dens_share = lq_part - (ld1 + ld2);
  return dens_share - lpx;
}

}  // namespace

template <class D>
__global__ void __launch_bounds__(LNT) lat_fwd_kernel(const __grid_constant__ DecParams P) {
  extern __shared__ __align__(16) float lsm[];
  pdl_launch_dependents();   // the decoder kernel may stage its weights while these tiles are processed
  if (blockIdx.x == 0 && threadIdx.x == 0 && P.gpre_max != nullptr) *P.gpre_max = 0u;   // batch maximum of |gpre|: lat_bwd accumulates it
  const LatSmem S = lat_carve(lsm, P, false);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hh = warp >> 2, p = 32 * q + lane;
  const int n = D::n_mc(P), RB = D::RB(P), nzd = D::nz_c(P) + D::nz_y(P), nzin = D::nz_x(P) + D::nd_p(P);
  const long long B = P.B;
  const bool mlp = P.phys_kind == 0;
  const int c1 = nzd, cs0 = nzd + 1;
  const long long rb = blockIdx.x;
  const long long row0 = rb * RB;
  const int nrows = (int)min((long long)RB, B - row0);
  const int npairs = nrows * n;

  load_row_params<D>(P, S, row0, nrows);
  // reparameterisation noise (kept for the backward)
  float* epsg = P.epsbuf + (long long)rb * D::Z(P) * TP;
  for (int e = tid; e < TP * D::Z(P); e += LNT) {
    const int pp = e & (TP - 1), i = e >> 7;
    float v = 0.0f;
    if (pp < npairs) {
      const int r = pp / n, m = pp - r * n;
      const unsigned long long grow = (unsigned long long)(P.row_off + (row0 + r) * P.row_stride);
      const int b = block_of_l<D>(P, i), il = i - D::blk_start(P, b), nzb = D::blk_size(P, b);
      const unsigned long long li = ((unsigned long long)m * (unsigned long long)P.Bg + grow) * nzb + il;
      v = P.rng.mode == 0 ? P.rng.eps[b][li] : philox_normal_elem(P.rng.seed, P.rng.ss ? P.rng.ss->philox_off[b] : P.rng.offset[b], P.rng.grid_threads[b], li);
    }
    S.EPS[e] = v;
    epsg[e] = v;
  }
  __syncthreads();
  for (int e = tid; e < RB * (D::n_blk(P) + 2); e += LNT) {
    const int t = e / RB, r = e - t * RB;
    float s = 0.0f;
    if (t < D::n_blk(P)) {
      for (int i = 0; i < D::blk_size(P, t); ++i) s += logf(S.ROWPAR[(D::rp_L(P) + D::blk_loff(P, t) + i * (i + 1) / 2 + i) * RBMAX + r]);
    } else {
      const int k0 = t == D::n_blk(P) ? 0 : D::nz_c(P), k1 = t == D::n_blk(P) ? D::nz_c(P) : nzd;
      for (int k = k0; k < k1; ++k) s += logf(S.ROWPAR[(D::rp_psig(P) + k) * RBMAX + r]);
    }
    S.ROWLOG[t * RBMAX + r] = s;
  }
  __syncthreads();

  const bool pvalid = p < npairs;
  const int prow = (pvalid ? p : npairs - 1) / n;
  const int pm_ = (pvalid ? p : npairs - 1) - prow * n;
  float dens_share;
  const float lq_part = sample_blocks<D>(P, S, hh, p, prow, true, dens_share);
  if (hh == 0)
    for (int j = 0; j < D::nd_p(P); ++j) S.ZXIN[(D::nz_x(P) + j) * TP + p] = S.ROWRAW[P.idx_c_phys[j] * RBMAX + prow];
  S.SC[(2 + hh) * TP + p] = dens_share;
  __syncthreads();
  {
    // conditional prior of side hh: p(zc|c) (hh = 0) or p(zy|y) (hh = 1), diagonal Gaussian
    const int a_nz = hh ? D::nz_y(P) : D::nz_c(P), a_j0 = hh ? D::nz_c(P) : 0;
    float mh = 0.0f;
    for (int k = a_j0; k < a_j0 + a_nz; ++k) {
      const float t = (S.ZD[k * TP + p] - S.ROWPAR[(D::rp_pmu(P) + k) * RBMAX + prow]) / S.ROWPAR[(D::rp_psig(P) + k) * RBMAX + prow];
      mh = fmaf(t, t, mh);
    }
    const float lp = -0.5f * ((float)a_nz * LOG_2PI + mh) - S.ROWLOG[(D::n_blk(P) + hh) * RBMAX + prow];
    S.SC[hh * TP + p] = pvalid ? lq_part - lp : 0.0f;
    if (hh == 0 && pvalid && (P.out.dens || P.out.zx || P.out.zc || P.out.zy)) {
      const long long o = (long long)pm_ * B + row0 + prow;
      if (P.out.dens) P.out.dens[o] = S.SC[2 * TP + p] + S.SC[3 * TP + p];
      if (P.out.zx) for (int k = 0; k < D::nz_x(P); ++k) P.out.zx[o * D::nz_x(P) + k] = S.ZXIN[k * TP + p];
      if (P.out.zc) for (int k = 0; k < D::nz_c(P); ++k) P.out.zc[o * D::nz_c(P) + k] = S.ZD[k * TP + p];
      if (P.out.zy) for (int k = 0; k < D::nz_y(P); ++k) P.out.zy[o * D::nz_y(P) + k] = S.ZD[(D::nz_c(P) + k) * TP + p];
    }
    // decoder-kernel input record: latent operand row, chunk hh: [zd | 1 | physics input | 0] * 2^4, fp16 hi / lo
    // planes (physics input: standardised for the MLP surrogate, raw zx for the closed forms)
    unsigned char* rec = P.rec + (long long)rb * P.rec_stride;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = 8 * hh + i;
      float x = 0.0f;
      if (k < nzd) x = pvalid ? S.ZD[k * TP + p] : 0.0f;
      else if (k == c1) x = 1.0f;
      else if (k >= cs0 && k < cs0 + nzin) {
        const float z = S.ZXIN[(k - cs0) * TP + p];
        x = mlp ? (z - P.phys_in_mean[k - cs0]) / P.phys_in_std[k - cs0] : z;
        if (!pvalid && mlp) x = 0.0f;
      }
      v[i] = x * 16.0f;
    }
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    *reinterpret_cast<uint4*>(rec + (hh * TP + p) * 16) = hi;
    *reinterpret_cast<uint4*>(rec + 4096 + (hh * TP + p) * 16) = lo;
    // raw covariates / labels per pair (rows of 128 floats)
    float* raw = reinterpret_cast<float*>(rec + 8192);
    if (hh == 0) {
      for (int j = 0; j < D::nd_c(P); ++j) raw[j * TP + p] = S.ROWRAW[j * RBMAX + prow];
    } else {
      for (int j = 0; j < D::nd_y(P); ++j) raw[(D::nd_c(P) + j) * TP + p] = S.ROWRAW[(D::nd_c(P) + j) * RBMAX + prow];
    }
  }
  __syncthreads();
  // per-row KL = mean over the MC axis (models/vae.py:207)
  if (tid < nrows) {
    float s = 0.0f;
    for (int m = 0; m < n; ++m) s += S.SC[tid * n + m] + S.SC[TP + tid * n + m];
    P.rowkl[row0 + tid] = s / (float)n;
  }
}

template <class D>
__global__ void __launch_bounds__(LNT) lat_bwd_kernel(const __grid_constant__ DecParams P) {
  extern __shared__ __align__(16) float lsm[];
  pdl_launch_dependents();   // the encoder backward kernel may stage its weights while these tiles are processed
  const LatSmem S0 = lat_carve(lsm, P, true);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hh = warp >> 2, p = 32 * q + lane;
  const int n = D::n_mc(P), RB = D::RB(P), nzd = D::nz_c(P) + D::nz_y(P);
  const long long B = P.B;
  const float wpair = 1.0f / ((float)P.Bg * (float)(P.nd_x + D::nd_c(P) + D::nd_y(P)) * (float)n);
  float gmax = 0.0f;
  // Persistent CTA over tiles; the two per-tile records (noise, dL/dz: 10 KB) of the NEXT tile are bulk-copied
  // (cp.async.bulk + mbarrier) while the current one is processed -- the one-tile-per-CTA version spent half of its
  // time waiting for these loads and for the barrier behind them.
  uint64_t* lbar = reinterpret_cast<uint64_t*>(S0.BAR);
  const uint32_t eps_bytes = (uint32_t)(D::Z(P) * TP * 4), dz_bytes = (uint32_t)((nzd + D::nz_x(P)) * TP * 4);
  if (tid == 0) {
    tc::mbar_init(lbar, 1);
    tc::mbar_init(lbar + 1, 1);
    tc::mbar_fence_init();
  }
  __syncthreads();
  auto prefetch = [&](long long tile, int buf) {
    tc::fence_async_smem();   // generic-proxy reads of this buffer (previous tile) before the async-proxy refill
    tc::mbar_expect_tx(lbar + buf, eps_bytes + dz_bytes);
    tc::bulk_g2s(buf ? S0.EPS2 : S0.EPS, P.epsbuf + tile * D::Z(P) * TP, eps_bytes, lbar + buf);
    tc::bulk_g2s(buf ? S0.DZ2 : S0.DZ, P.dzrec + tile * (nzd + D::nz_x(P)) * TP, dz_bytes, lbar + buf);
  };
  if (tid == 0 && (long long)blockIdx.x < P.n_rowblocks) prefetch(blockIdx.x, 0);
  // raw head pre-activations of a tile ([O_tot] feature rows x RB minibatch rows, 32-byte segments of the feature-major
  // `headpre`): staged one tile ahead with 4-byte cp.async copies, so that the row-parameter pass reads shared memory
  const int n_raw = P.O_tot * RB;
  auto stage_heads = [&](long long tile, int buf) {
    float* dst = S0.RAWH + buf * P.O_tot * RBMAX;
    const long long r0 = tile * RB;
    const int nr = (int)min((long long)RB, B - r0);
    for (int e = tid; e < n_raw; e += LNT) {
      const int f = e / RB, r = e - f * RB;
      const float* src = P.headpre + (long long)f * B + r0 + min(r, nr - 1);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dst + f * RBMAX + r)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((long long)blockIdx.x < P.n_rowblocks) stage_heads(blockIdx.x, 0);
  int it = 0;
  for (long long rb = blockIdx.x; rb < P.n_rowblocks; rb += gridDim.x, ++it) {
  const int buf = it & 1;
  LatSmem S = S0;
  if (buf) { S.EPS = S0.EPS2; S.DZ = S0.DZ2; }
  const long long row0 = rb * RB;
  const int nrows = (int)min((long long)RB, B - row0);
  const int npairs = nrows * n;
  if (tid == 0 && rb + gridDim.x < P.n_rowblocks) prefetch(rb + gridDim.x, buf ^ 1);   // its last readers passed the barrier below
  asm volatile("cp.async.wait_group 0;" ::: "memory");   // this tile's staged heads (own copies), then everybody's
  __syncthreads();
  if (rb + gridDim.x < P.n_rowblocks) stage_heads(rb + gridDim.x, buf ^ 1);
  load_row_params<D>(P, S, row0, nrows, true, S0.RAWH + buf * P.O_tot * RBMAX);
  tc::mbar_wait(lbar + buf, (uint32_t)(it >> 1) & 1u);
  __syncthreads();
  const bool pvalid = p < npairs;
  const int prow = (pvalid ? p : npairs - 1) / n;
  float dens_share;
  sample_blocks<D>(P, S, hh, p, prow, false, dens_share);
  __syncthreads();

  // per-pair gradients w.r.t. loc / L / prior parameters
  {
    const int pp = tid & (TP - 1), prt = tid >> 7;
    const int r = (pp < npairs ? pp : npairs - 1) / n;
    const float bw = pp < npairs ? P.beta_x * wpair : 0.0f;
    const float wv = pp < npairs ? 1.0f : 0.0f;
    for (int k = prt; k < nzd; k += 2) {
      float g = wv * S.DZ[k * TP + pp];
      const float sgm = S.ROWPAR[(D::rp_psig(P) + k) * RBMAX + r];
      const float t = (S.ZD[k * TP + pp] - S.ROWPAR[(D::rp_pmu(P) + k) * RBMAX + r]) / sgm;
      g += bw * t / sgm;
      S.FEAT[(D::f_pmu(P) + k) * TP + pp] = -bw * t / sgm;
      S.FEAT[(D::f_psig(P) + k) * TP + pp] = -bw * (t * t - 1.0f) / sgm;
      S.FEAT[(D::f_loc(P) + D::nz_x(P) + k) * TP + pp] = g;
    }
    for (int i = prt; i < D::nz_x(P); i += 2) {
      float g = wv * S.DZ[(nzd + i) * TP + pp];
      if (P.prior_kind[i] == 1) g += bw * (S.ZXIN[i * TP + pp] - P.prior_a[i]) / (P.prior_b[i] * P.prior_b[i]);
      const float u = S.U[i * TP + pp];
      S.FEAT[(D::f_loc(P) + i) * TP + pp] = g * (P.ub[i] - P.lb[i]) * u * (1.0f - u) + bw * (2.0f * u - 1.0f);
    }
  }
  __syncthreads();
  for (int e = tid; e < TP * D::nL(P); e += LNT) {
    const int pp = e & (TP - 1), li = e >> 7;
    const int r = (pp < npairs ? pp : npairs - 1) / n;
    const int b = D::L_blk(P, li), i = D::L_i(P, li), j = D::L_j(P, li), s = D::blk_start(P, b);
    float v = S.FEAT[(D::f_loc(P) + s + i) * TP + pp] * S.EPS[(s + j) * TP + pp];
    if (i == j) v -= (pp < npairs ? P.beta_x * wpair : 0.0f) / S.ROWPAR[(D::rp_L(P) + li) * RBMAX + r];
    S.FEAT[(D::f_L(P) + li) * TP + pp] = v;
  }
  __syncthreads();
  // reduce over the MC axis (n consecutive pairs per row, fixed order)
  for (int e = tid; e < D::n_feat(P) * RBMAX; e += LNT) {
    const int f = e / RBMAX, r = e - f * RBMAX;
    if (r < nrows) {
      const float* src = S.FEAT + f * TP + r * n;
      float s = 0.0f;
      if ((n & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (int m = 0; m < (n >> 2); ++m) {
          const float4 t = s4[m];
          s += (t.x + t.y) + (t.z + t.w);
        }
      } else {
        for (int m = 0; m < n; ++m) s += src[m];
      }
      S.ROWACC[f * RBMAX + r] = s;
    }
  }
  __syncthreads();
  // gradients w.r.t. the head pre-activations (clamp / exp chain rule, models/encoders.py:35-43): the factors were
  // computed from the pre-activations when the row parameters were loaded -- no global re-read here
  for (int e = tid; e < RB * D::Z(P); e += LNT) {
    const int i = e / RB, r = e - i * RB;
    if (r < nrows) {
      const long long lrow = row0 + r;
      const int b = block_of_l<D>(P, i), il = i - D::blk_start(P, b), nzb = D::blk_size(P, b);
      const long long om = (long long)(D::henc(P, b) + il) * B + lrow;
      const float gm_ = S.ROWMSK[(D::rp_loc(P) + i) * RBMAX + r] * S.ROWACC[(D::f_loc(P) + i) * RBMAX + r];
      P.gpre[om] = gm_;
      gmax = fmaxf(gmax, fabsf(gm_));
      for (int j = 0; j < nzb; ++j) {
        const long long oc = (long long)(D::henc(P, b) + 2 * nzb + il * nzb + j) * B + lrow;
        float g = 0.0f;
        if (j < il) {
          const int li = D::blk_loff(P, b) + il * (il + 1) / 2 + j;
          g = S.ROWMSK[(D::rp_L(P) + li) * RBMAX + r] * S.ROWACC[(D::f_L(P) + li) * RBMAX + r];
        }
        P.gpre[oc] = g;
        gmax = fmaxf(gmax, fabsf(g));
      }
      const long long os = (long long)(D::henc(P, b) + nzb + il) * B + lrow;
      const int ld = D::blk_loff(P, b) + il * (il + 1) / 2 + il;
      const float gs_ = S.ROWMSK[(D::rp_L(P) + ld) * RBMAX + r] * S.ROWACC[(D::f_L(P) + ld) * RBMAX + r];
      P.gpre[os] = gs_;
      gmax = fmaxf(gmax, fabsf(gs_));
    }
  }
  for (int e = tid; e < RB * nzd; e += LNT) {
    const int k = e / RB, r = e - k * RB;
    if (r < nrows) {
      const long long lrow = row0 + r;
      const int which = k < D::nz_c(P) ? 0 : 1;
      const int kk = which ? k - D::nz_c(P) : k, nzk = which ? D::nz_y(P) : D::nz_c(P);
      const long long om = (long long)(D::hpri(P, which) + kk) * B + lrow;
      const long long os = (long long)(D::hpri(P, which) + nzk + kk) * B + lrow;
      P.gpre[om] = S.ROWMSK[(D::rp_pmu(P) + k) * RBMAX + r] * S.ROWACC[(D::f_pmu(P) + k) * RBMAX + r];
      P.gpre[os] = S.ROWMSK[(D::rp_psig(P) + k) * RBMAX + r] * S.ROWACC[(D::f_psig(P) + k) * RBMAX + r];
    }
  }
  __syncthreads();   // every buffer of this tile is free again
  }  // tiles
  // largest |gradient of an encoder-head pre-activation| of the batch: the tensor-core encoder backward scales its
  // fp16 operand from it (order-independent maximum of non-negative floats: deterministic)
  if (P.gpre_max != nullptr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
    if (lane == 0 && gmax > 0.0f && gmax < __int_as_float(0x7f800000)) atomicMax(P.gpre_max, __float_as_uint(gmax));
  }
}

// Encode-only inference (models/vae.py:161-162, 125-151): one thread per (MC sample, row) pair, head pre-activations read
// feature-major (coalesced along the rows), latents and density written in the reference's (n, B, .) layout.  No
// decoder, no priors: the HBM-streaming kernel of BASELINE.json config 5.
template <class D>
__global__ void __launch_bounds__(LNT) lat_encode_kernel(const __grid_constant__ DecParams P) {
  extern __shared__ __align__(16) float lsm[];   // EPS[Z][LNT] | HP[heads used][LNT]
  const int tid = threadIdx.x;
  const long long B = P.B;
  const long long q = (long long)blockIdx.x * LNT + tid;
  if (q >= (long long)P.n_mc * B) return;
  const long long m = q / B, r = q - m * B;
  const unsigned long long grow = (unsigned long long)(P.row_off + r * P.row_stride);
  // All head pre-activations of this row go to shared memory with 4-byte async copies (own column, read back by the
  // same thread in the same order): ~33 independent loads in flight under the Philox work, instead of loads that the
  // output stores (possible aliases for the compiler) keep in program order.
  float* HP = lsm + (size_t)D::Z(P) * LNT;
  {
    int nh = 0;
    auto stage = [&](int row) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(HP + (size_t)(nh++) * LNT + tid)),
                   "l"(P.headpre + (long long)row * B + r) : "memory");
    };
#pragma unroll
    for (int b = 0; b < D::n_blk(P); ++b) {
      const int nzb = D::blk_size(P, b);
      for (int i = 0; i < nzb; ++i) {
        stage(D::henc(P, b) + i);
        for (int j = 0; j < i; ++j) stage(D::henc(P, b) + 2 * nzb + i * nzb + j);
        stage(D::henc(P, b) + nzb + i);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = 0; i < D::Z(P); ++i) {
    const int b = block_of_l<D>(P, i), il = i - D::blk_start(P, b), nzb = D::blk_size(P, b);
    const unsigned long long li = ((unsigned long long)m * (unsigned long long)P.Bg + grow) * nzb + il;
    lsm[i * LNT + tid] = P.rng.mode == 0 ? P.rng.eps[b][li] : philox_normal_elem(P.rng.seed, P.rng.ss ? P.rng.ss->philox_off[b] : P.rng.offset[b], P.rng.grid_threads[b], li);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float dens = 0.0f;
  int nh = 0;
#pragma unroll
  for (int b = 0; b < D::n_blk(P); ++b) {
    const int s = D::blk_start(P, b), nzb = D::blk_size(P, b);
    float ss = 0.0f, hld = 0.0f, ld1 = 0.0f, ld2 = 0.0f;
    for (int i = 0; i < nzb; ++i) {
      float acc = clampf_(HP[(size_t)(nh++) * LNT + tid], -50.0f, 50.0f);
      for (int j = 0; j < i; ++j) {
        const float lij = clampf_(HP[(size_t)(nh++) * LNT + tid], -20.0f, 20.0f);
        acc = fmaf(lij, lsm[(s + j) * LNT + tid], acc);
      }
      const float lii = expf(clampf_(HP[(size_t)(nh++) * LNT + tid], -7.0f, 3.0f)) + 1e-8f;
      const float e = lsm[(s + i) * LNT + tid];
      acc = fmaf(lii, e, acc);
      ss = fmaf(e, e, ss);
      hld += logf(lii);
      const int gi = s + i;
      if (gi < D::nz_x(P)) {
        const float u = sigmoidf_(acc);
        const float a = P.ub[gi] - P.lb[gi];
        ld1 += acc - 2.0f * softplusf_(acc);
        ld2 += logf(fabsf(a));
        if (P.out.zx) P.out.zx[q * D::nz_x(P) + gi] = fmaf(u, a, P.lb[gi]);
      } else if (gi < D::nz_x(P) + D::nz_c(P)) {
        if (P.out.zc) P.out.zc[q * D::nz_c(P) + (gi - D::nz_x(P))] = acc;
      } else {
        if (P.out.zy) P.out.zy[q * D::nz_y(P) + (gi - D::nz_x(P) - D::nz_c(P))] = acc;
      }
    }
    const float lq = -0.5f * ((float)nzb * LOG_2PI + ss) - hld;
    if (b == 0) dens = lq - (ld1 + ld2);
    else dens += lq;
  }
  if (P.out.dens) P.out.dens[q] = dens;
}

namespace {

// =====================================================================================================================
// Thread-per-pair kernels for the compile-time shapes (Shape::K): the same math as lat_fwd_kernel / lat_bwd_kernel above,
// but one thread owns one (row, MC-sample) pair END TO END, with its noise, latents and gradients in registers and every
// block / triangular loop unrolled; only the per-ROW quantities (clamped / exponentiated head outputs, their chain-rule
// factors, log-determinants) are staged in shared memory, once per tile of RB rows.  The kernels above move every
// intermediate through shared-memory planes between strided passes of 256 threads (two threads per pair, run-time index
// arithmetic per element): 3840 / 2880 executed instructions per pair against ~700 here (plus the noise).
//
// Noise: the tile-ordered `epsbuf` record is replaced by a buffer in the LOCAL (m, row, i) order of each noise tensor
// (P.eps_local[b]); on an unsharded Philox call it is filled AHEAD by lat_noise_fill_kernel with torch's own mapping --
// one Philox4x32-10 evaluation + two Box-Muller pairs per FOUR elements -- instead of one evaluation per element (the
// per-element generator was ~half of lat_fwd_kernel's instructions).  Row-sharded calls (elements of one evaluation
// belong to different ranks) and injected noise keep the per-element path; the forward then stores what it drew.
constexpr int PNT = 128;   // threads per CTA = pairs per tile

template <int N>
__device__ __forceinline__ void ldg_vec(const float* __restrict__ src, float* v) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) { const float4 t = __ldg(reinterpret_cast<const float4*>(src) + i); v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) { const float2 t = __ldg(reinterpret_cast<const float2*>(src) + i); v[2 * i] = t.x; v[2 * i + 1] = t.y; }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __ldg(src + i);
  }
}
template <int N>
__device__ __forceinline__ void stg_vec(float* dst, const float* v) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) reinterpret_cast<float2*>(dst)[i] = make_float2(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = v[i];
  }
}

// Row parameters of one tile for the thread-per-pair kernels: same values as load_row_params, but every global load of a
// thread is issued BEFORE the first dependent instruction (the four strided passes of load_row_params each waited a full
// memory latency: 27 % of the forward kernel's stall samples), and the per-row derived quantities the pair threads would
// otherwise recompute 16 times are staged next to them in ROWAUX (same indexing as ROWPAR):
//   MODE 1 (forward):  log L_ii, log sigma_prior, 1 / sigma_prior (in the prior-mean slot), raw c / y rows
//   MODE 2 (backward): 1 / L_ii, 1 / sigma_prior, and the clamp / exp chain-rule factors in ROWMSK
template <class D, int MODE>
__device__ __forceinline__ void pair_stage_rows(const DecParams& P, const LatSmem& S, long long row0, int nrows) {
  constexpr int RB = TP / D::cn_mc, Z = D::cZ, nL = D::cnL, nzd = D::cnzd, NC = D::cNC, NY = D::cNY, ndc = D::cndc, ndy = D::cndy;
  constexpr int rpL = Z, rpM = Z + nL, rpS = Z + nL + nzd;
  constexpr int N1 = (RB * Z + PNT - 1) / PNT, N2 = (RB * nL + PNT - 1) / PNT, N3 = (RB * nzd + PNT - 1) / PNT;
  constexpr int N4 = MODE == 1 ? (RB * (ndc + ndy) + PNT - 1) / PNT : 0;
  const int tid = threadIdx.x;
  const long long B = P.B;
  float a1[N1], a2[N2], a3m[N3], a3s[N3], a4[N4 > 0 ? N4 : 1];
  const bool has_y = P.y != nullptr;
#pragma unroll
  for (int t = 0; t < N1; ++t) {
    const int e = tid + t * PNT, i = e / RB, r = e - i * RB;
    a1[t] = 0.0f;
    if (e < RB * Z) {
      const int b = block_of_l<D>(P, i), il = i - D::blk_start(P, b);
      a1[t] = __ldg(P.headpre + (long long)(D::henc(P, b) + il) * B + row0 + min(r, nrows - 1));
    }
  }
#pragma unroll
  for (int t = 0; t < N2; ++t) {
    const int e = tid + t * PNT, li = e / RB, r = e - li * RB;
    a2[t] = 0.0f;
    if (e < RB * nL) {
      const int b = D::L_blk(P, li), i = D::L_i(P, li), j = D::L_j(P, li), nzb = D::blk_size(P, b);
      const int f = D::henc(P, b) + (i == j ? nzb + i : 2 * nzb + i * nzb + j);
      a2[t] = __ldg(P.headpre + (long long)f * B + row0 + min(r, nrows - 1));
    }
  }
#pragma unroll
  for (int t = 0; t < N3; ++t) {
    const int e = tid + t * PNT, k = e / RB, r = e - k * RB;
    a3m[t] = 0.0f; a3s[t] = 0.0f;
    if (e < RB * nzd) {
      const int which = k < NC ? 0 : 1, kk = which ? k - NC : k, nzk = which ? NY : NC;
      if (which == 0 || has_y) {
        const float* h0 = P.headpre + (long long)(D::hpri(P, which) + kk) * B + row0 + min(r, nrows - 1);
        a3m[t] = __ldg(h0);
        a3s[t] = __ldg(h0 + (long long)nzk * B);
      }
    }
  }
  if constexpr (MODE == 1) {
#pragma unroll
    for (int t = 0; t < N4; ++t) {
      const int e = tid + t * PNT, j = e / RB, r = e - j * RB;
      a4[t] = 0.0f;
      if (e < RB * (ndc + ndy)) {
        const long long lrow = row0 + min(r, nrows - 1);
        const long long drow = P.idx ? P.idx[lrow] : lrow;
        if (j < ndc) a4[t] = __ldg(P.c + drow * ndc + j);
        else if (has_y) a4[t] = __ldg(P.y + drow * ndy + (j - ndc));
      }
    }
  }
  // ---- dependent part ----
#pragma unroll
  for (int t = 0; t < N1; ++t) {
    const int e = tid + t * PNT, i = e / RB, r = e - i * RB;
    if (e < RB * Z) {
      const float pm = a1[t];
      S.ROWPAR[i * RBMAX + r] = clampf_(pm, -50.0f, 50.0f);
      // d clamp / d pre: 1 inside the clamp range (models/encoders.py:35-43)
      if constexpr (MODE == 2) S.ROWMSK[i * RBMAX + r] = (pm >= -50.0f && pm <= 50.0f) ? 1.0f : 0.0f;
    }
  }
#pragma unroll
  for (int t = 0; t < N2; ++t) {
    const int e = tid + t * PNT, li = e / RB, r = e - li * RB;
    if (e < RB * nL) {
      const bool diag = D::L_i(P, li) == D::L_j(P, li);
      const float pre = a2[t];
      float v, mk;
      if (diag) {
        v = expf(clampf_(pre, -7.0f, 3.0f)) + 1e-8f;
        mk = (pre >= -7.0f && pre <= 3.0f) ? expf(pre) : 0.0f;   // d exp(clamp(ps)) / d ps = exp(ps) inside the range
        S.ROWAUX[(rpL + li) * RBMAX + r] = MODE == 1 ? logf(v) : 1.0f / v;
      } else {
        v = clampf_(pre, -20.0f, 20.0f);
        mk = (pre >= -20.0f && pre <= 20.0f) ? 1.0f : 0.0f;
      }
      S.ROWPAR[(rpL + li) * RBMAX + r] = v;
      if constexpr (MODE == 2) S.ROWMSK[(rpL + li) * RBMAX + r] = mk;
    }
  }
#pragma unroll
  for (int t = 0; t < N3; ++t) {
    const int e = tid + t * PNT, k = e / RB, r = e - k * RB;
    if (e < RB * nzd) {
      const int which = k < NC ? 0 : 1;
      float mu = 0.0f, sgm = 1.0f, mkm = 0.0f, mks = 0.0f;
      if (which == 0 || has_y) {
        const float pm = a3m[t], ps = a3s[t];
        mu = clampf_(pm, -50.0f, 50.0f);
        sgm = expf(clampf_(ps, -7.0f, 3.0f)) + 1e-8f;
        mkm = (pm >= -50.0f && pm <= 50.0f) ? 1.0f : 0.0f;
        mks = (ps >= -7.0f && ps <= 3.0f) ? expf(ps) : 0.0f;
      }
      S.ROWPAR[(rpM + k) * RBMAX + r] = mu;
      S.ROWPAR[(rpS + k) * RBMAX + r] = sgm;
      S.ROWAUX[(rpM + k) * RBMAX + r] = 1.0f / sgm;
      if constexpr (MODE == 1) S.ROWAUX[(rpS + k) * RBMAX + r] = logf(sgm);
      if constexpr (MODE == 2) {
        S.ROWMSK[(rpM + k) * RBMAX + r] = mkm;
        S.ROWMSK[(rpS + k) * RBMAX + r] = mks;
      }
    }
  }
  if constexpr (MODE == 1) {
#pragma unroll
    for (int t = 0; t < N4; ++t) {
      const int e = tid + t * PNT, j = e / RB, r = e - j * RB;
      if (e < RB * (ndc + ndy)) S.ROWRAW[j * RBMAX + r] = a4[t];
    }
  }
}

// compile-time block geometry of shape D
template <class D, int b> struct Blk {
  static constexpr int nz = D::cnb == 1 ? D::cZ : (b == 0 ? D::cNX : (b == 1 ? D::cNC : D::cNY));
  static constexpr int s = D::cnb == 1 ? 0 : (b == 0 ? 0 : (b == 1 ? D::cNX : D::cNX + D::cNC));
  static constexpr int loff = D::cnb == 1 ? 0 : (b == 0 ? 0 : (b == 1 ? tri(D::cNX) : tri(D::cNX) + tri(D::cNC)));
};

// noise of latent block b for pair (m, local row lrow): eps[s .. s + nz)
template <class D, int b>
__device__ __forceinline__ void pair_eps_fwd(const DecParams& P, int m, long long lrow, bool store, float* eps) {
  constexpr int nz = Blk<D, b>::nz, s = Blk<D, b>::s;
  float* loc = P.eps_local[b] + ((long long)m * P.B + lrow) * nz;
  if (P.eps_ready) {
    ldg_vec<nz>(loc, eps + s);
  } else {
    const unsigned long long li0 = ((unsigned long long)m * (unsigned long long)P.Bg + (unsigned long long)(P.row_off + lrow * P.row_stride)) * nz;
    const unsigned long long off = P.rng.ss ? P.rng.ss->philox_off[b] : P.rng.offset[b];
#pragma unroll
    for (int i = 0; i < nz; ++i)
      eps[s + i] = P.rng.mode == 0 ? P.rng.eps[b][li0 + i] : philox_normal_elem(P.rng.seed, off, P.rng.grid_threads[b], li0 + i);
    if (store) stg_vec<nz>(loc, eps + s);
  }
}

// z = loc + L eps of block b (+ bijector on the z_x dimensions).  LOGS: also log q, the bijector's log-determinant and the
// per-pair part of log p(z_x) (the parts that are constant per dimension arrive through CONSTS, staged once per tile)
template <class D, int b, bool LOGS>
__device__ __forceinline__ void pair_sample_block(const DecParams& P, const float* ROWPAR, const float* ROWAUX, int prow, const float* eps,
                                                  float* z, float* u, float& lq, float& ld1, float& lpx) {
  constexpr int nz = Blk<D, b>::nz, s = Blk<D, b>::s, loff = Blk<D, b>::loff;
  float ss = 0.0f, sl = 0.0f;
#pragma unroll
  for (int i = 0; i < nz; ++i) {
    float acc = ROWPAR[(s + i) * RBMAX + prow];
#pragma unroll
    for (int j = 0; j <= i; ++j) acc = fmaf(ROWPAR[(D::cZ + loff + i * (i + 1) / 2 + j) * RBMAX + prow], eps[s + j], acc);
    const int gi = s + i;
    if constexpr (LOGS) {
      ss = fmaf(eps[s + i], eps[s + i], ss);
      sl += ROWAUX[(D::cZ + loff + i * (i + 1) / 2 + i) * RBMAX + prow];   // log L_ii (sum in index order: = log|det L| of the block)
    }
    if (gi < D::cNX) {
      const float uu = sigmoidf_(acc);
      const float zx = fmaf(uu, P.ub[gi] - P.lb[gi], P.lb[gi]);
      u[gi < D::cNX ? gi : 0] = uu;
      z[gi] = zx;
      if constexpr (LOGS) {
        ld1 += acc - 2.0f * softplusf_(acc);
        if (P.prior_kind[gi] == 0) {
          const bool inside = (zx >= P.prior_a[gi]) && (zx < P.prior_b[gi]);
          lpx += inside ? 0.0f : -INFINITY;
        } else {
          const float d = zx - P.prior_a[gi];
          lpx += -(d * d) / (2.0f * P.prior_b[gi] * P.prior_b[gi]) - logf(P.prior_b[gi]) - LOG_SQRT_2PI;
        }
      }
    } else {
      z[gi] = acc;
    }
  }
  if constexpr (LOGS) lq += -0.5f * ((float)nz * LOG_2PI + ss) - sl;
}

template <class D>
__global__ void __launch_bounds__(PNT) lat_pair_fwd_kernel(const __grid_constant__ DecParams P) {
  extern __shared__ __align__(16) float lsm[];
  pdl_launch_dependents();   // the decoder kernel may stage its weights while these tiles are processed
  constexpr int n = D::cn_mc, RB = TP / n, Z = D::cZ, NX = D::cNX, nzd = D::cnzd, nb = D::cnb, ndc = D::cndc, ndy = D::cndy;
  constexpr int nzin = NX + D::cndp, c1 = nzd, cs0 = nzd + 1, NF = D::cn_rowpar, rpM = Z + D::cnL, rpS = rpM + nzd;
  static_assert((n & (n - 1)) == 0 && n <= 32 && RB <= RBMAX, "MC count must be a power of two that fills whole warps");
  if (blockIdx.x == 0 && threadIdx.x == 0 && P.gpre_max != nullptr) *P.gpre_max = 0u;   // batch maximum of |gpre|: lat_bwd accumulates it
  LatSmem S;
  S.ROWPAR = lsm;
  S.ROWAUX = S.ROWPAR + NF * RBMAX;
  S.ROWRAW = S.ROWAUX + NF * RBMAX;
  float* CONSTS = S.ROWRAW + (ndc + ndy) * RBMAX;   // [0] sum log|ub - lb|, [1] constant part of log p(z_x), [2..] 1 / physics-input std
  const int p = threadIdx.x;
  const long long B = P.B, rb = blockIdx.x, row0 = rb * RB;
  const int nrows = (int)min((long long)RB, B - row0), npairs = nrows * n;
  const bool pvalid = p < npairs;
  const int pc = pvalid ? p : npairs - 1;
  const int prow = pc / n, pm = pc - prow * n;
  const bool mlp = P.phys_kind == 0;

  // noise first: these loads (or the Philox evaluations) do not depend on the row parameters staged below
  float eps[Z];
  {
    const bool store = pvalid && P.with_grad;
    pair_eps_fwd<D, 0>(P, pm, row0 + prow, store, eps);
    if constexpr (nb > 1) {
      pair_eps_fwd<D, 1>(P, pm, row0 + prow, store, eps);
      pair_eps_fwd<D, 2>(P, pm, row0 + prow, store, eps);
    }
  }
  pair_stage_rows<D, 1>(P, S, row0, nrows);
  if (p == PNT - 1) {   // per-dimension constants of the bijector log-det and of the uniform priors (utils/transforms.py:97-150, utils/priors.py:19-23)
    float c_ld2 = 0.0f, c_lpx = 0.0f;
#pragma unroll
    for (int gi = 0; gi < NX; ++gi) {
      c_ld2 += logf(fabsf(P.ub[gi] - P.lb[gi]));
      if (P.prior_kind[gi] == 0) c_lpx += -logf(P.prior_b[gi] - P.prior_a[gi]);
    }
    CONSTS[0] = c_ld2;
    CONSTS[1] = c_lpx;
  }
  if (p >= PNT - 1 - nzin && p < PNT - 1) CONSTS[2 + (PNT - 2 - p)] = 1.0f / P.phys_in_std[PNT - 2 - p];
  __syncthreads();

  float z[Z], u[NX > 0 ? NX : 1];
  float lq = 0.0f, ld1 = 0.0f, lpx = CONSTS[1];
  pair_sample_block<D, 0, true>(P, S.ROWPAR, S.ROWAUX, prow, eps, z, u, lq, ld1, lpx);
  if constexpr (nb > 1) {
    pair_sample_block<D, 1, true>(P, S.ROWPAR, S.ROWAUX, prow, eps, z, u, lq, ld1, lpx);
    pair_sample_block<D, 2, true>(P, S.ROWPAR, S.ROWAUX, prow, eps, z, u, lq, ld1, lpx);
  }
  const float dens = lq - (ld1 + CONSTS[0]);
  // conditional priors p(zc|c), p(zy|y): diagonal Gaussians (models/vae.py:200-207)
  float kl = dens - lpx;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int a_nz = side ? D::cNY : D::cNC, a_j0 = side ? D::cNC : 0;
    float mh = 0.0f, sl = 0.0f;
#pragma unroll
    for (int k = 0; k < (D::cNC > D::cNY ? D::cNC : D::cNY); ++k)
      if (k < a_nz) {
        const int kk = a_j0 + k;
        const float t = (z[NX + kk] - S.ROWPAR[(rpM + kk) * RBMAX + prow]) * S.ROWAUX[(rpM + kk) * RBMAX + prow];
        mh = fmaf(t, t, mh);
        sl += S.ROWAUX[(rpS + kk) * RBMAX + prow];   // log sigma
      }
    kl -= -0.5f * ((float)a_nz * LOG_2PI + mh) - sl;
  }
  // per-row KL = mean over the MC axis (models/vae.py:207): the n samples of a row are n consecutive lanes
  {
    float s = pvalid ? kl : 0.0f;
#pragma unroll
    for (int off = n >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (pvalid && (p & (n - 1)) == 0) P.rowkl[row0 + prow] = s / (float)n;
  }
  if (pvalid && (P.out.dens || P.out.zx || P.out.zc || P.out.zy)) {
    const long long o = (long long)pm * B + row0 + prow;
    if (P.out.dens) P.out.dens[o] = dens;
    if (P.out.zx) for (int k = 0; k < NX; ++k) P.out.zx[o * NX + k] = z[k];
    if (P.out.zc) for (int k = 0; k < D::cNC; ++k) P.out.zc[o * D::cNC + k] = z[NX + k];
    if (P.out.zy) for (int k = 0; k < D::cNY; ++k) P.out.zy[o * D::cNY + k] = z[NX + D::cNC + k];
  }
  // decoder-kernel input record: latent operand row [zd | 1 | physics input | 0] * 2^4 as fp16 hi / lo planes (physics
  // input: standardised for the MLP surrogate, raw zx for the closed forms), then the raw covariates / labels per pair
  unsigned char* rec = P.rec + rb * P.rec_stride;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = 8 * hh + i;
      float x = 0.0f;
      if (k < nzd) x = pvalid ? z[NX + k] : 0.0f;
      else if (k == c1) x = 1.0f;
      else if (k >= cs0 && k < cs0 + nzin) {
        const int q = k - cs0;
        const float zz = q < NX ? z[q < NX ? q : 0] : S.ROWRAW[P.idx_c_phys[q - NX < 0 ? 0 : q - NX] * RBMAX + prow];
        x = mlp ? (zz - P.phys_in_mean[q]) * CONSTS[2 + q] : zz;
        if (!pvalid && mlp) x = 0.0f;
      }
      v[i] = x * 16.0f;
    }
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    *reinterpret_cast<uint4*>(rec + (hh * TP + p) * 16) = hi;
    *reinterpret_cast<uint4*>(rec + 4096 + (hh * TP + p) * 16) = lo;
  }
  float* raw = reinterpret_cast<float*>(rec + 8192);
#pragma unroll
  for (int j = 0; j < ndc + ndy; ++j) raw[j * TP + p] = S.ROWRAW[j * RBMAX + prow];
}

// feature row (shared index of ROWPAR / ROWMSK / FEAT) behind head-output row o of headpre / gpre, or -1 (the unused upper
// triangle of a covariance head: zero gradient)
template <class D>
__device__ __forceinline__ int pair_src_feature(const DecParams& P, int o) {
  int f = -1;
  for (int b = 0; b < D::cnb; ++b) {
    const int nzb = D::blk_size(P, b), s = D::blk_start(P, b), h0 = D::henc(P, b), loff = D::blk_loff(P, b);
    if (o >= h0 && o < h0 + 2 * nzb + nzb * nzb) {
      const int l = o - h0;
      if (l < nzb) f = D::f_loc(P) + s + l;
      else if (l < 2 * nzb) { const int il = l - nzb; f = D::f_L(P) + loff + il * (il + 1) / 2 + il; }
      else { const int q = l - 2 * nzb, il = q / nzb, j = q - il * nzb; if (j < il) f = D::f_L(P) + loff + il * (il + 1) / 2 + j; }
    }
  }
  for (int which = 0; which < 2; ++which) {
    const int nzk = which ? D::cNY : D::cNC, h0 = D::hpri(P, which), k0 = which ? D::cNC : 0;
    if (o >= h0 && o < h0 + 2 * nzk) { const int l = o - h0; f = l < nzk ? D::f_pmu(P) + k0 + l : D::f_psig(P) + k0 + (l - nzk); }
  }
  return f;
}

// Per-pair gradient features of latent block b, written to shared memory in TWO phases that share one buffer (the full
// [feature][pair] plane -- 25 KB for P bridge, 41 KB for the S presets -- held the kernel at 35 % occupancy):
//   phase A: d/d loc (slot = latent index), d/d prior mean, d/d prior sigma (slots Z.., Z + nzd..); gl[] stays in registers
//   phase B: d/d L entries (slot = packed lower-triangular index)
template <class D, int b>
__device__ __forceinline__ void pair_grad_block_a(const DecParams& P, const float* ROWPAR, const float* ROWAUX, int prow, int p, float bw,
                                                  const float* z, const float* u, const float* dz, float* gl, float* FEAT) {
  constexpr int nz = Blk<D, b>::nz, s = Blk<D, b>::s, NX = D::cNX, nzd = D::cnzd, Z = D::cZ;
  constexpr int f_pmu = Z + D::cnL;
#pragma unroll
  for (int i = 0; i < nz; ++i) {
    const int gi = s + i;
    if (gi < NX) {
      float g = dz[nzd + gi];
      if (P.prior_kind[gi] == 1) g += bw * (z[gi] - P.prior_a[gi]) / (P.prior_b[gi] * P.prior_b[gi]);
      const float uu = u[gi < NX ? gi : 0];
      gl[gi] = g * (P.ub[gi] - P.lb[gi]) * uu * (1.0f - uu) + bw * (2.0f * uu - 1.0f);
    } else {
      const int k = gi - NX;
      const float inv = ROWAUX[(f_pmu + k) * RBMAX + prow];   // 1 / sigma of the conditional prior
      const float t = (z[gi] - ROWPAR[(f_pmu + k) * RBMAX + prow]) * inv;
      const float bti = bw * t * inv;
      gl[gi] = dz[k] + bti;
      FEAT[(Z + k) * PNT + p] = -bti;
      FEAT[(Z + nzd + k) * PNT + p] = -bw * (t * t - 1.0f) * inv;
    }
    FEAT[gi * PNT + p] = gl[gi];
  }
}
template <class D, int b>
__device__ __forceinline__ void pair_grad_block_b(const float* ROWAUX, int prow, int p, float bw, const float* eps, const float* gl,
                                                  float* FEAT) {
  constexpr int nz = Blk<D, b>::nz, s = Blk<D, b>::s, loff = Blk<D, b>::loff, Z = D::cZ;
#pragma unroll
  for (int i = 0; i < nz; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const int li = loff + i * (i + 1) / 2 + j;
      float v = gl[s + i] * eps[s + j];
      if (j == i) v -= bw * ROWAUX[(Z + li) * RBMAX + prow];   // 1 / L_ii
      FEAT[li * PNT + p] = v;
    }
  }
}

template <class D>
__global__ void __launch_bounds__(PNT) lat_pair_bwd_kernel(const __grid_constant__ DecParams P) {
  extern __shared__ __align__(16) float lsm[];
  pdl_launch_dependents();   // the encoder backward kernel may stage its weights while these tiles are processed
  constexpr int n = D::cn_mc, RB = TP / n, Z = D::cZ, NX = D::cNX, nb = D::cnb, NF = D::cn_rowpar, NQ = n >> 2;
  constexpr int nL = D::cnL, nzd = D::cnzd, NFA = Z + 2 * nzd, NFB = nL, NFM = NFA > NFB ? NFA : NFB;
  static_assert(n % 4 == 0 && (NQ & (NQ - 1)) == 0, "MC reduction reads float4 groups in a rotated order");
  LatSmem S;
  S.ROWPAR = lsm;
  S.ROWMSK = S.ROWPAR + NF * RBMAX;
  S.ROWAUX = S.ROWMSK + NF * RBMAX;
  S.ROWRAW = nullptr;
  S.FEAT = S.ROWAUX + NF * RBMAX;
  int* SRC = reinterpret_cast<int*>(S.FEAT + NFM * PNT);
  const int p = threadIdx.x;
  const long long B = P.B, rb = blockIdx.x, row0 = rb * RB;
  const int nrows = (int)min((long long)RB, B - row0), npairs = nrows * n;
  const bool pvalid = p < npairs;
  const int pc = pvalid ? p : npairs - 1;
  const int prow = pc / n, pm = pc - prow * n;
  const float wpair = 1.0f / ((float)P.Bg * (float)(P.nd_x + D::cndc + D::cndy) * (float)n);
  const float bw = pvalid ? P.beta_x * wpair : 0.0f;

  // this pair's noise (local order, written by the forward or by the noise pre-pass) and dL/dz (decoder kernel's record)
  float eps[Z], dz[Z];
  ldg_vec<Blk<D, 0>::nz>(P.eps_local[0] + ((long long)pm * B + row0 + prow) * Blk<D, 0>::nz, eps + Blk<D, 0>::s);
  if constexpr (nb > 1) {
    ldg_vec<Blk<D, 1>::nz>(P.eps_local[1] + ((long long)pm * B + row0 + prow) * Blk<D, 1>::nz, eps + Blk<D, 1>::s);
    ldg_vec<Blk<D, 2>::nz>(P.eps_local[2] + ((long long)pm * B + row0 + prow) * Blk<D, 2>::nz, eps + Blk<D, 2>::s);
  }
  {
    const float* DZ = P.dzrec + rb * (long long)Z * TP;
#pragma unroll
    for (int k = 0; k < Z; ++k) dz[k] = pvalid ? __ldg(DZ + k * TP + p) : 0.0f;
  }
  pair_stage_rows<D, 2>(P, S, row0, nrows);
  for (int o = p; o < P.O_tot; o += PNT) SRC[o] = pair_src_feature<D>(P, o);
  __syncthreads();

  // head-output gradients: MC-axis sum of the feature (n consecutive pairs, fixed order) times the clamp / exp chain-rule
  // factor of the head (models/encoders.py:35-43); consecutive threads write consecutive rows of one feature row of gpre.
  // The float4 groups of a row are read in an order rotated by the row index: the 8 rows of a quarter-warp then hit
  // 8 different bank groups (64-byte row stride: 4-way conflicts otherwise); the order is a function of the row only.
  float gmax = 0.0f;
  const int henc_end = D::hpri(P, 0);
  auto reduce_out = [&](bool phase_b) {
    for (int e = p; e < P.O_tot * RB; e += PNT) {
      const int o = e / RB, r = e - o * RB;
      const int f = SRC[o];
      const bool in_b = f >= Z && f < Z + nL;
      if (r < nrows && in_b == phase_b) {
        float g = 0.0f;
        if (f >= 0) {
          const int slot = phase_b ? f - Z : (f < Z ? f : f - nL);
          const float4* s4 = reinterpret_cast<const float4*>(S.FEAT + slot * PNT + r * n);
          const int rot = r >> 1;
          float s = 0.0f;
#pragma unroll
          for (int m = 0; m < NQ; ++m) {
            const float4 t = s4[(m + rot) & (NQ - 1)];
            s += (t.x + t.y) + (t.z + t.w);
          }
          g = S.ROWMSK[f * RBMAX + r] * s;
        }
        P.gpre[(long long)o * B + row0 + r] = g;
        if (o < henc_end) gmax = fmaxf(gmax, fabsf(g));
      }
    }
  };
  float gl[Z];
  {
    float z[Z], u[NX > 0 ? NX : 1];
    float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
    pair_sample_block<D, 0, false>(P, S.ROWPAR, S.ROWAUX, prow, eps, z, u, d0, d1, d2);
    pair_grad_block_a<D, 0>(P, S.ROWPAR, S.ROWAUX, prow, p, bw, z, u, dz, gl, S.FEAT);
    if constexpr (nb > 1) {
      pair_sample_block<D, 1, false>(P, S.ROWPAR, S.ROWAUX, prow, eps, z, u, d0, d1, d2);
      pair_grad_block_a<D, 1>(P, S.ROWPAR, S.ROWAUX, prow, p, bw, z, u, dz, gl, S.FEAT);
      pair_sample_block<D, 2, false>(P, S.ROWPAR, S.ROWAUX, prow, eps, z, u, d0, d1, d2);
      pair_grad_block_a<D, 2>(P, S.ROWPAR, S.ROWAUX, prow, p, bw, z, u, dz, gl, S.FEAT);
    }
  }
  __syncthreads();
  reduce_out(false);   // loc and prior heads (and the zero rows of the unused upper triangles)
  __syncthreads();
  pair_grad_block_b<D, 0>(S.ROWAUX, prow, p, bw, eps, gl, S.FEAT);
  if constexpr (nb > 1) {
    pair_grad_block_b<D, 1>(S.ROWAUX, prow, p, bw, eps, gl, S.FEAT);
    pair_grad_block_b<D, 2>(S.ROWAUX, prow, p, bw, eps, gl, S.FEAT);
  }
  __syncthreads();
  reduce_out(true);    // sigma and covariance heads
  if (P.gpre_max != nullptr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
    if ((p & 31) == 0 && gmax > 0.0f && gmax < __int_as_float(0x7f800000)) atomicMax(P.gpre_max, __float_as_uint(gmax));
  }
}

// Reparameterisation noise of an unsharded Philox call, generated the way torch's normal_ kernel generates it: ONE
// Philox4x32-10 evaluation + two Box-Muller pairs per FOUR elements (generator thread idx, loop iteration j -> elements
// (4 j + k) GT + idx, k = 0..3; the same stream as philox_normal_elem, common.cuh, which spends one evaluation per
// element).  grid = (generator threads / 256, loop iterations, latent blocks); output order == torch's (m, row, i).
__global__ void __launch_bounds__(256) lat_noise_fill_kernel(const __grid_constant__ DecParams P) {
  const int b = blockIdx.z;
  const unsigned int GT = P.rng.grid_threads[b];
  const unsigned long long numel = (unsigned long long)P.n_mc * (unsigned long long)P.Bg * (unsigned long long)P.blk_size[b];
  const unsigned int idx = blockIdx.x * 256u + threadIdx.x;
  const unsigned long long j = blockIdx.y;
  const unsigned long long li0 = (4ull * j) * GT + idx;
  if (idx >= GT || li0 >= numel) return;
  const unsigned long long off = P.rng.ss ? P.rng.ss->philox_off[b] : P.rng.offset[b];
  const unsigned long long nn = (off >> 2) + j;
  const uint4 ctr = make_uint4((unsigned int)nn, (unsigned int)(nn >> 32), idx, 0u);
  const uint2 key = make_uint2((unsigned int)P.rng.seed, (unsigned int)(P.rng.seed >> 32));
  const uint4 r = curand_Philox4x32_10(ctr, key);
  const float2 g0 = _curand_box_muller(r.x, r.y), g1 = _curand_box_muller(r.z, r.w);
  float* out = P.eps_local[b];
  out[li0] = g0.x;
  if (li0 + GT < numel) out[li0 + GT] = g0.y;
  if (li0 + 2ull * GT < numel) out[li0 + 2ull * GT] = g1.x;
  if (li0 + 3ull * GT < numel) out[li0 + 3ull * GT] = g1.y;
}

// The same for a CYCLIC row shard (this rank owns the global rows row_off, row_off + S, ...; S = row_stride).  The four
// elements of one evaluation are GT apart in the flattened (m, row, i) tensor = GT / nz rows apart; with S | GT / nz and
// S | Bg they all belong to the rank that owns generator threads idx with (idx / nz) mod S == row_off, so every rank
// evaluates exactly its own quarter-count of the stream (api.cu checks the divisibility; contiguous row blocks, at most
// ~1.7 GT long per MC sample, would leave almost nothing to share).  Output in the rank's LOCAL (m, row, i) order.
__global__ void __launch_bounds__(256) lat_noise_fill_cyclic_kernel(const __grid_constant__ DecParams P) {
  const int b = blockIdx.z;
  const unsigned int GT = P.rng.grid_threads[b], nz = (unsigned int)P.blk_size[b];
  const unsigned int S = (unsigned int)P.row_stride, r0 = (unsigned int)P.row_off, Bg = (unsigned int)P.Bg;
  const unsigned int t = blockIdx.x * 256u + threadIdx.x;   // this rank's t-th generator thread
  if (t >= P.cyc_nloc[b]) return;
  // (a run-time 32-bit division costs ~30 instructions: the per-launch quotients come from the host, and the two divisors
  // that are powers of two in every shipped shape are shifts)
  const bool nz_p2 = (nz & (nz - 1u)) == 0u, s_p2 = (S & (S - 1u)) == 0u;
  const unsigned int u = nz_p2 ? t >> (__ffs((int)nz) - 1) : t / nz, il = t - u * nz;
  const unsigned int idx = (u * S + r0) * nz + il;
  const unsigned long long numel = (unsigned long long)P.n_mc * (unsigned long long)P.Bg * nz;
  const unsigned long long j = blockIdx.y;
  const unsigned long long li0 = (4ull * j) * GT + idx;
  if (li0 >= numel) return;
  const unsigned long long off = P.rng.ss ? P.rng.ss->philox_off[b] : P.rng.offset[b];
  const unsigned long long nn = (off >> 2) + j;
  const uint4 ctr = make_uint4((unsigned int)nn, (unsigned int)(nn >> 32), idx, 0u);
  const uint2 key = make_uint2((unsigned int)P.rng.seed, (unsigned int)(P.rng.seed >> 32));
  const uint4 r = curand_Philox4x32_10(ctr, key);
  const float2 g0 = _curand_box_muller(r.x, r.y), g1 = _curand_box_muller(r.z, r.w);
  const float vals[4] = {g0.x, g0.y, g1.x, g1.y};
  // (m, local row) of element k = 0 -- one division -- then GT / (nz S) LOCAL rows further per element, wrapping into the
  // next MC sample after B local rows (n_mc * Bg < 2^31: 32-bit arithmetic)
  const unsigned int GTn = P.cyc_gtn[b], step = P.cyc_step[b], Bl = (unsigned int)P.B;
  const unsigned int trow = 4u * (unsigned int)j * GTn + (u * S + r0);   // = li0 / nz (GT is a multiple of nz)
  unsigned int m = trow / Bg;
  const unsigned int gl = trow - m * Bg - r0;
  unsigned int lrow = s_p2 ? gl >> (__ffs((int)S) - 1) : gl / S;
  float* out = P.eps_local[b];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (li0 + (unsigned long long)k * GT < numel) out[((unsigned long long)m * Bl + lrow) * nz + il] = vals[k];
    lrow += step;
    while (lrow >= Bl) { lrow -= Bl; ++m; }
  }
}

template <class D>
static size_t pair_smem_bytes(const DecParams& p, bool bwd) {
  if (bwd) {
    constexpr int nfa = D::cZ + 2 * D::cnzd, nfm = nfa > D::cnL ? nfa : D::cnL;   // feature plane of the larger phase
    return (size_t)(3 * D::cn_rowpar * RBMAX + nfm * PNT + p.O_tot) * sizeof(float);
  }
  return (size_t)(2 * D::cn_rowpar * RBMAX + (D::cndc + D::cndy) * RBMAX + 16) * sizeof(float);
}

}  // namespace

size_t lat_smem_bytes(const DecParams& p, bool bwd) { return (size_t)lat_smem_floats(p, bwd) * sizeof(float); }
// shapes of the reference's cases (cases/*/__init__.py presets): bridge / damped_oscillator / simple_beam, P and S,
// at the training MC count n_mc = 16; everything else runs the generic instantiation
using ShBridgeP = Shape<0, 2, 4, 4, 2, 2, 1, 16>;
using ShBridgeS = Shape<1, 2, 4, 4, 2, 2, 1, 16>;
using ShOscP = Shape<0, 1, 4, 4, 1, 1, 0, 16>;
using ShOscS = Shape<1, 1, 4, 4, 1, 1, 0, 16>;
using ShBeamP = Shape<0, 2, 2, 2, 1, 1, 0, 16>;
using ShBeamS = Shape<1, 2, 2, 2, 1, 1, 0, 16>;

template <int MT, int NX, int NC, int NY, int NDC, int NDY, int NDP, int NMC>
static bool shape_matches(const DecParams& p, Shape<MT, NX, NC, NY, NDC, NDY, NDP, NMC>) {
  return p.model_type == MT && p.nz_x == NX && p.nz_c == NC && p.nz_y == NY && p.nd_c == NDC && p.nd_y == NDY && p.nd_p == NDP &&
         p.n_mc == NMC && p.RB == 128 / NMC;
}
template <class SH>
static void launch_lat_pair(const DecParams& p, long long n_tiles, bool bwd, cudaStream_t s) {
  if constexpr (SH::K) {
    if (p.eps_local[0] != nullptr) {   // thread-per-pair kernels (api.cu hands out the local-order noise buffer when lat_pair_supported)
      if (bwd) lat_pair_bwd_kernel<SH><<<(unsigned)n_tiles, PNT, pair_smem_bytes<SH>(p, true), s>>>(p);
      else lat_pair_fwd_kernel<SH><<<(unsigned)n_tiles, PNT, pair_smem_bytes<SH>(p, false), s>>>(p);
      return;
    }
  }
  if (bwd) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const size_t smem = lat_smem_bytes(p, true);
    long long per_sm = (long long)(227 * 1024) / (long long)(smem + 1024);
    per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
    const long long grid = n_tiles < per_sm * sms ? n_tiles : per_sm * sms;
    lat_bwd_kernel<SH><<<(unsigned)grid, LNT, smem, s>>>(p);
  }
  else lat_fwd_kernel<SH><<<(unsigned)n_tiles, LNT, lat_smem_bytes(p, false), s>>>(p);
}
static void launch_lat(const DecParams& p, long long n_tiles, bool bwd, cudaStream_t s) {
  if (shape_matches(p, ShBridgeP())) launch_lat_pair<ShBridgeP>(p, n_tiles, bwd, s);
  else if (shape_matches(p, ShBridgeS())) launch_lat_pair<ShBridgeS>(p, n_tiles, bwd, s);
  else if (shape_matches(p, ShOscP())) launch_lat_pair<ShOscP>(p, n_tiles, bwd, s);
  else if (shape_matches(p, ShOscS())) launch_lat_pair<ShOscS>(p, n_tiles, bwd, s);
  else if (shape_matches(p, ShBeamP())) launch_lat_pair<ShBeamP>(p, n_tiles, bwd, s);
  else if (shape_matches(p, ShBeamS())) launch_lat_pair<ShBeamS>(p, n_tiles, bwd, s);
  else launch_lat_pair<ShGeneric>(p, n_tiles, bwd, s);
}
template <class SH>
static void launch_lat_encode_t(const DecParams& p, cudaStream_t s) {
  const long long total = (long long)p.n_mc * p.B;
  int nh = 0;
  for (int b = 0; b < p.n_blk; ++b) nh += 2 * p.blk_size[b] + p.blk_size[b] * (p.blk_size[b] - 1) / 2;
  const size_t smem = (size_t)(p.Z + nh) * LNT * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaFuncSetAttribute(lat_encode_kernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set = smem;
  }
  lat_encode_kernel<SH><<<(unsigned)((total + LNT - 1) / LNT), LNT, smem, s>>>(p);
}
// the encode kernel only depends on the latent layout (model type and block sizes), not on n_mc or the data widths
template <int MT, int NX, int NC, int NY, int NDC, int NDY, int NDP, int NMC>
static bool latent_layout_matches(const DecParams& p, Shape<MT, NX, NC, NY, NDC, NDY, NDP, NMC>) {
  return p.model_type == MT && p.nz_x == NX && p.nz_c == NC && p.nz_y == NY;
}
void launch_lat_encode(const DecParams& p, cudaStream_t s) {
  if (latent_layout_matches(p, ShBridgeP())) launch_lat_encode_t<ShBridgeP>(p, s);
  else if (latent_layout_matches(p, ShBridgeS())) launch_lat_encode_t<ShBridgeS>(p, s);
  else if (latent_layout_matches(p, ShOscP())) launch_lat_encode_t<ShOscP>(p, s);
  else if (latent_layout_matches(p, ShOscS())) launch_lat_encode_t<ShOscS>(p, s);
  else if (latent_layout_matches(p, ShBeamP())) launch_lat_encode_t<ShBeamP>(p, s);
  else if (latent_layout_matches(p, ShBeamS())) launch_lat_encode_t<ShBeamS>(p, s);
  else launch_lat_encode_t<ShGeneric>(p, s);
}
// the thread-per-pair kernels cover the compile-time shapes (the MC count and the tile geometry are part of the shape)
bool lat_pair_supported(const DecParams& p) {
  return shape_matches(p, ShBridgeP()) || shape_matches(p, ShBridgeS()) || shape_matches(p, ShOscP()) || shape_matches(p, ShOscS()) ||
         shape_matches(p, ShBeamP()) || shape_matches(p, ShBeamS());
}
void launch_lat_noise_fill(const DecParams& p_in, bool cyclic, cudaStream_t s) {
  DecParams p = p_in;
  if (cyclic)
    for (int b = 0; b < p.n_blk; ++b) {
      const unsigned int GT = p.rng.grid_threads[b], nz = (unsigned int)p.blk_size[b], S = (unsigned int)p.row_stride;
      p.cyc_nloc[b] = GT / S; p.cyc_gtn[b] = GT / nz; p.cyc_step[b] = GT / nz / S;
    }
  unsigned long long gt = 0, iters = 0;
  for (int b = 0; b < p.n_blk; ++b) {
    const unsigned long long GT = p.rng.grid_threads[b], numel = (unsigned long long)p.n_mc * (unsigned long long)p.Bg * (unsigned long long)p.blk_size[b];
    const unsigned long long it = (numel + 4ull * GT - 1) / (4ull * GT);
    gt = GT > gt ? GT : gt;
    iters = it > iters ? it : iters;
  }
  if (cyclic) {
    const unsigned long long mine = gt / (unsigned long long)p.row_stride;   // generator threads of this rank
    lat_noise_fill_cyclic_kernel<<<dim3((unsigned)((mine + 255) / 256), (unsigned)iters, (unsigned)p.n_blk), 256, 0, s>>>(p);
  } else {
    lat_noise_fill_kernel<<<dim3((unsigned)((gt + 255) / 256), (unsigned)iters, (unsigned)p.n_blk), 256, 0, s>>>(p);
  }
}
void launch_lat_fwd(const DecParams& p, long long n_tiles, cudaStream_t s) { launch_lat(p, n_tiles, false, s); }
void launch_lat_bwd(const DecParams& p, long long n_tiles, cudaStream_t s) { launch_lat(p, n_tiles, true, s); }
template <class SH>
static int configure_lat_pair() {
  int e = (int)cudaFuncSetAttribute(lat_fwd_kernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (!e) e = (int)cudaFuncSetAttribute(lat_bwd_kernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  if constexpr (SH::K)
    if (!e) e = (int)cudaFuncSetAttribute(lat_pair_bwd_kernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  return e;
}
int configure_lat_kernels() {
  int e = configure_lat_pair<ShGeneric>();
  if (!e) e = configure_lat_pair<ShBridgeP>();
  if (!e) e = configure_lat_pair<ShBridgeS>();
  if (!e) e = configure_lat_pair<ShOscP>();
  if (!e) e = configure_lat_pair<ShOscS>();
  if (!e) e = configure_lat_pair<ShBeamP>();
  if (!e) e = configure_lat_pair<ShBeamS>();
  return e;
}

}  // namespace dpv
