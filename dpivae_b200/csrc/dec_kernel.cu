// Decoder-side fused kernel of the DPI-VAE step (sm_100a).
//
// One persistent CTA per SM walks "row blocks" (RB minibatch rows x n_mc Monte-Carlo samples) in
// chunks of 64 (row, sample) pairs and, entirely out of shared memory, does per chunk:
//   reparameterised sample + bijector + log q / log prior        (models/encoders.py:73-93,
//                                                                 utils/transforms.py:97-150,
//                                                                 utils/priors.py:19-23)
//   auxiliary decoders c / y + their Gaussian log-likelihoods      (models/decoders.py:36-49)
//   physics decoder (frozen Tanh-MLP surrogate or closed form)      (models/nn.py:67-80,
//                                                                 cases/*/mass_spring.py, simple_beam_model.py)
//   data-driven decoder behind the gradient-reversal layer        (models/decoders.py:79-92)
//   ELBO terms (models/vae.py:177-231) and, when with_grad, the whole backward of
//   ELBO / (B_global * (nd_x+nd_c+nd_y)) (dpivae.py:419,429): decoder wgrads accumulate into this
//   CTA's private partial-gradient vector, latent gradients are reduced over the MC axis and
//   written as gradients w.r.t. the encoder / prior-net head pre-activations (gpre).
#include <curand_kernel.h>

#include "common.cuh"
#include "kernels.h"

namespace dpv {

// scalar rows (one value per pair) inside the SC region
enum { SC_KL = 0, SC_RX, SC_RC, SC_RY, SC_REG, SC_W, SC_LSX, SC_P0, SC_Q0 = SC_P0 + 4, SC_ROWS = SC_Q0 + 4 };

// optional per-phase cycle accounting (thread 0, clock64 after each barrier); see tools/phase_profile.py
enum { PH_SETUP = 0, PH_ROWPAR, PH_EPS, PH_LATENT_FWD, PH_AUX_FWD, PH_AUX_LOSS, PH_AUX_BWD, PH_PHYS_FWD, PH_DATA_FWD,
       PH_XLOSS, PH_DATA_BWD, PH_PHYS_BWD, PH_LATENT_BWD, PH_ROWRED, PH_ROWOUT, PH_COUNT };
#define PHASE(k)                                          \
  do {                                                    \
    if (P.phase != nullptr && tid == 0) {                 \
      const long long _t = clock64();                     \
      PHS[k] += _t - t_last;                              \
      t_last = _t;                                        \
    }                                                     \
  } while (0)


__device__ __forceinline__ void zero_range(float* p, long long cnt) {
  for (long long e = threadIdx.x; e < cnt; e += NT) p[e] = 0.0f;
}

__device__ __forceinline__ void zero_mlp2_part(float* part, const Mlp2S& m) {
  zero_range(part + m.g_w0, (long long)m.K0 * m.H);
  zero_range(part + m.g_b0, m.H);
  zero_range(part + m.g_w1, (long long)m.H * m.O);
  zero_range(part + m.g_b1, m.O);
}

__device__ __forceinline__ void stage_mlp2(float* sm, const float* params, const Mlp2S& m) {
  stage_linear(params + m.g_w0, params + m.g_b0, m.K0, m.H, sm + m.s_w0t, m.ldw0, sm + m.s_b0);
  stage_linear(params + m.g_w1, params + m.g_b1, m.H, m.O, sm + m.s_w1t, m.ldw1, sm + m.s_b1);
}

// (n, B, d) output tensor <- feature-major smem rows, for the valid pairs of this chunk
__device__ __forceinline__ void store_pairs(float* __restrict__ g, const float* __restrict__ rows, int d, long long q0,
                                            long long npairs, long long row0, int n, long long B) {
  if (g == nullptr) return;
  for (int e = threadIdx.x; e < TILE * d; e += NT) {
    const int p = e / d, k = e - p * d;
    const long long q = q0 + p;
    if (q < npairs) {
      const long long r = q / n, m = q - r * n;
      g[(m * B + row0 + r) * d + k] = rows[k * LDP + p];
    }
  }
}

__device__ __forceinline__ int block_of(const DecParams& P, int i) {
  int b = 0;
  while (b + 1 < P.n_blk && i >= P.blk_start[b + 1]) ++b;
  return b;
}

__global__ void __launch_bounds__(NT, 1) dec_kernel(const __grid_constant__ DecParams P) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int n = P.n_mc;
  const int nzd = P.nz_c + P.nz_y;
  const int nzin = P.nz_x + P.nd_p;
  const int ndxp = pad4(P.nd_x);
  const int ocp = pad4(2 * P.nd_c), oyp = pad4(2 * P.nd_y);
  const long long B = P.B;
  float* part = P.part + (long long)blockIdx.x * P.part_stride;

  // ---- one-time: zero smem, stage weights, zero this CTA's gradient partials -----------------
  for (int e = tid; e < P.s_total; e += NT) sm[e] = 0.0f;
  __syncthreads();
  stage_mlp2(sm, P.params, P.fx);
  stage_mlp2(sm, P.params, P.dc);
  stage_mlp2(sm, P.params, P.dy);
  if (P.phys_kind == 0) {
    for (int l = 0; l < P.phys_n_layers; ++l)
      stage_linear(P.frozen + P.pl[l].g_w, P.frozen + P.pl[l].g_b, P.pl[l].K, P.pl[l].N, sm + P.pl[l].s_wt,
                   P.pl[l].ldw, sm + P.pl[l].s_b);
  }
  if (P.with_grad) {
    zero_mlp2_part(part, P.fx);
    zero_mlp2_part(part, P.dc);
    zero_mlp2_part(part, P.dy);
    if (tid == 0) part[P.g_lsx] = 0.0f;
  }
  zero_range(part + P.n_params, NSCAL);
  __syncthreads();

  float* XHP = sm + P.s_XHP;
  float* HD = sm + P.s_HD;
  float* XHD = sm + P.s_XHD;
  float* HC = HD;
  float* HY = HD + 64 * LDP;
  float* EPS = sm + P.s_EPS;
  float* EPSC = sm + P.s_EPSC;
  float* U = sm + P.s_U;
  float* ZXIN = sm + P.s_ZXIN;
  float* S0 = sm + P.s_S0;
  float* ZD = sm + P.s_ZD;
  float* OC = sm + P.s_OC;
  float* OY = sm + P.s_OY;
  float* DZD = sm + P.s_DZD;
  float* DZC = sm + P.s_DZC;
  float* DZY = sm + P.s_DZY;
  float* DZX = sm + P.s_DZX;
  float* SC = sm + P.s_SC;
  float* ROWPAR = sm + P.s_ROWPAR;  // [n_rowpar][RBMAX]
  float* ROWRAW = sm + P.s_ROWRAW;  // [nd_c + nd_y][RBMAX]
  float* ROWX = sm + P.s_ROWX;      // [nd_x][RBMAX] raw x rows of the block
  float* FEAT = sm + P.s_FEAT;      // aliases the activation region (dead by the time it is written)
  // per-row accumulators [n_feat + 5][racc_ld]: single-chunk blocks keep them right behind FEAT in the
  // (dead) activation region, multi-chunk blocks (RB == 1) in a small dedicated region
  const bool multi_chunk = P.n_chunks > 1;
  const int racc_ld = multi_chunk ? 1 : RBMAX;
  float* ROWACC = multi_chunk ? sm + P.s_ROWACC : FEAT + P.n_feat * LDP;
  long long* PHS = reinterpret_cast<long long*>(sm + P.s_PH);
  long long t_last = (P.phase != nullptr && tid == 0) ? clock64() : 0;

  const float lsx = P.params[P.g_lsx];
  const float sx = expf(lsx);
  const float var_x = sx * sx;
  const float wpair = 1.0f / ((float)P.Bg * (float)(P.nd_x + P.nd_c + P.nd_y) * (float)n);
  const int n_acc = P.n_feat + 5;

  float tot[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // thread 0: running sums of the per-row outputs
  const bool n_pow2 = (n & (n - 1)) == 0 && n <= 32;
  PHASE(PH_SETUP);

  for (long long rb = blockIdx.x; rb < P.n_rowblocks; rb += gridDim.x) {
    const long long row0 = rb * P.RB;
    const int nrows = (int)min((long long)P.RB, B - row0);
    const long long npairs = (long long)nrows * n;

    // ---- per-row parameters of q(z|x) and the conditional priors ---------------------------
    for (int e = tid; e < RBMAX * P.Z; e += NT) {
      const int i = e / RBMAX, r = e - i * RBMAX;
      const long long lrow = row0 + min(r, nrows - 1);
      const int b = block_of(P, i), il = i - P.blk_start[b];
      const float pm = P.headpre[(long long)(P.henc[b] + il) * B + lrow];
      ROWPAR[(P.rp_loc + i) * RBMAX + r] = clampf_(pm, -50.0f, 50.0f);
    }
    for (int e = tid; e < RBMAX * P.nL; e += NT) {
      const int li = e / RBMAX, r = e - li * RBMAX;
      const long long lrow = row0 + min(r, nrows - 1);
      const int b = P.L_blk[li], i = P.L_i[li], j = P.L_j[li], nzb = P.blk_size[b];
      float v;
      if (i == j) {
        const float ps = P.headpre[(long long)(P.henc[b] + nzb + i) * B + lrow];
        v = expf(clampf_(ps, -7.0f, 3.0f)) + 1e-8f;
      } else {
        const float pc = P.headpre[(long long)(P.henc[b] + 2 * nzb + i * nzb + j) * B + lrow];
        v = clampf_(pc, -20.0f, 20.0f);
      }
      ROWPAR[(P.rp_L + li) * RBMAX + r] = v;
    }
    for (int e = tid; e < RBMAX * nzd; e += NT) {
      const int k = e / RBMAX, r = e - k * RBMAX;
      const long long lrow = row0 + min(r, nrows - 1);
      const int which = k < P.nz_c ? 0 : 1;
      const int kk = which ? k - P.nz_c : k, nzk = which ? P.nz_y : P.nz_c;
      float mu = 0.0f, sg = 1.0f;
      if (which == 0 || P.y != nullptr) {
        const float pm = P.headpre[(long long)(P.hpri[which] + kk) * B + lrow];
        const float ps = P.headpre[(long long)(P.hpri[which] + nzk + kk) * B + lrow];
        mu = clampf_(pm, -50.0f, 50.0f);
        sg = expf(clampf_(ps, -7.0f, 3.0f)) + 1e-8f;
      }
      ROWPAR[(P.rp_pmu + k) * RBMAX + r] = mu;
      ROWPAR[(P.rp_psig + k) * RBMAX + r] = sg;
    }
    // --full_cov_prior: strict lower triangle of the conditional priors' factors (clamped f_cov head, models/encoders.py:38-43)
    for (int e = tid; e < RBMAX * P.npL; e += NT) {
      const int li = e / RBMAX, r = e - li * RBMAX;
      const long long lrow = row0 + min(r, nrows - 1);
      const int which = li < P.pl_off[1] ? 0 : 1;
      const int nzk = which ? P.nz_y : P.nz_c;
      int l = li - P.pl_off[which], i = 1;
      while (l >= i) { l -= i; ++i; }   // l = j < i
      float v = 0.0f;
      if (which == 0 || P.y != nullptr) v = clampf_(P.headpre[(long long)(P.hpri[which] + 2 * nzk + i * nzk + l) * B + lrow], -20.0f, 20.0f);
      ROWPAR[(P.rp_pL + li) * RBMAX + r] = v;
    }
    for (int e = tid; e < RBMAX * (P.nd_c + P.nd_y); e += NT) {
      const int j = e / RBMAX, r = e - j * RBMAX;
      const long long lrow = row0 + min(r, nrows - 1);
      const long long drow = P.idx ? P.idx[lrow] : lrow;
      float v = 0.0f;
      if (j < P.nd_c) v = P.c != nullptr ? P.c[drow * P.nd_c + j] : 0.0f;   // encode-only calls pass neither c nor y
      else if (P.y != nullptr) v = P.y[drow * P.nd_y + (j - P.nd_c)];
      ROWRAW[j * RBMAX + r] = v;
    }
    for (int e = tid; e < RBMAX * P.nd_x; e += NT) {
      const int r = e / P.nd_x, d = e - r * P.nd_x;
      const long long lrow = row0 + min(r, nrows - 1);
      const long long drow = P.idx ? P.idx[lrow] : lrow;
      ROWX[d * RBMAX + r] = P.x[drow * P.nd_x + d];
    }
    if (multi_chunk)
      for (int e = tid; e < n_acc; e += NT) ROWACC[e] = 0.0f;
    __syncthreads();
    PHASE(PH_ROWPAR);

    for (int chunk = 0; chunk < P.n_chunks; ++chunk) {
      const long long q0 = (long long)chunk * TILE;
      if (q0 >= npairs) break;

      // ---- P1: reparameterisation noise ------------------------------------------------------
      for (int e = tid; e < TILE * P.Z; e += NT) {
        const int p = e & (TILE - 1), i = e >> 6;
        const long long q = q0 + p;
        float v = 0.0f;
        if (q < npairs) {
          const long long r = q / n, m = q - r * n;
          const unsigned long long grow = (unsigned long long)(P.row_off + (row0 + r) * P.row_stride);
          const int b = block_of(P, i), il = i - P.blk_start[b], nzb = P.blk_size[b];
          const unsigned long long li = ((unsigned long long)m * (unsigned long long)P.Bg + grow) * nzb + il;
          v = P.rng.mode == 0 ? P.rng.eps[b][li] : philox_normal_elem(P.rng.seed, P.rng.ss ? P.rng.ss->philox_off[b] : P.rng.offset[b], P.rng.grid_threads[b], li);
        }
        EPS[i * LDP + p] = v;
      }
      if (P.cond) {
        for (int e = tid; e < TILE * P.nz_c; e += NT) {
          const int p = e & (TILE - 1), i = e >> 6;
          const long long q = q0 + p;
          float v = 0.0f;
          if (q < npairs) {
            const long long r = q / n, m = q - r * n;
            const unsigned long long grow = (unsigned long long)(P.row_off + (row0 + r) * P.row_stride);
            const unsigned long long li = ((unsigned long long)m * (unsigned long long)P.Bg + grow) * P.nz_c + i;
            v = P.rng.mode == 0 ? P.rng.eps[3][li] : philox_normal_elem(P.rng.seed, P.rng.ss ? P.rng.ss->philox_off[3] : P.rng.offset[3], P.rng.grid_threads[3], li);
          }
          EPSC[i * LDP + p] = v;
        }
      }
      __syncthreads();
      PHASE(PH_EPS);

      // ---- P2: z = loc + L eps, bijector, log q, log prior (one thread per pair) ---------------
      if (tid < TILE) {
        const int p = tid;
        const long long q = q0 + p;
        const bool valid = q < npairs;
        const long long qq = valid ? q : npairs - 1;
        const int r = (int)(qq / n);
        const int m = (int)(qq - (long long)r * n);
        float dens = 0.0f, ld1 = 0.0f, ld2 = 0.0f, lpx = 0.0f;
        for (int b = 0; b < P.n_blk; ++b) {
          const int s = P.blk_start[b], nzb = P.blk_size[b];
          float ss = 0.0f, hld = 0.0f;
          for (int i = 0; i < nzb; ++i) {
            float acc = ROWPAR[(P.rp_loc + s + i) * RBMAX + r];
            const int base = P.rp_L + P.blk_loff[b] + i * (i + 1) / 2;
            for (int j = 0; j <= i; ++j) acc = fmaf(ROWPAR[(base + j) * RBMAX + r], EPS[(s + j) * LDP + p], acc);
            const float e = EPS[(s + i) * LDP + p];
            ss = fmaf(e, e, ss);
            hld += logf(ROWPAR[(base + i) * RBMAX + r]);
            const int gi = s + i;
            if (gi < P.nz_x) {
              // Logistic(k=1) then ShiftScale(lb, ub)
              const float u = sigmoidf_(acc);
              const float a = P.ub[gi] - P.lb[gi];
              const float zx = fmaf(u, a, P.lb[gi]);
              ld1 += acc - 2.0f * softplusf_(acc);
              ld2 += logf(fabsf(a));
              U[gi * LDP + p] = u;
              ZXIN[gi * LDP + p] = zx;
              if (P.prior_kind[gi] == 0) {
                const bool inside = (zx >= P.prior_a[gi]) && (zx < P.prior_b[gi]);
                lpx += (inside ? 0.0f : -INFINITY) - logf(P.prior_b[gi] - P.prior_a[gi]);
              } else {
                const float d = zx - P.prior_a[gi];
                lpx += -(d * d) / (2.0f * P.prior_b[gi] * P.prior_b[gi]) - logf(P.prior_b[gi]) - LOG_SQRT_2PI;
              }
            } else {
              ZD[(gi - P.nz_x) * LDP + p] = acc;
            }
          }
          const float lq = -0.5f * ((float)nzb * LOG_2PI + ss) - hld;
          if (b == 0) dens = lq - (ld1 + ld2);
          else dens += lq;
        }
        // strict-lower factor entry (i, j), j < i, of conditional prior `which` (zero for the default diagonal priors)
        auto pL = [&](int which, int i, int j) -> float {
          return P.prior_full ? ROWPAR[(P.rp_pL + P.pl_off[which] + i * (i - 1) / 2 + j) * RBMAX + r] : 0.0f;
        };
        if (P.cond) {
          // zc = prior loc + prior scale_tril eps (models/vae.py:165-167, models/encoders.py:84-86)
          for (int k = P.nz_c - 1; k >= 0; --k) {
            float acc = fmaf(ROWPAR[(P.rp_psig + k) * RBMAX + r], EPSC[k * LDP + p], ROWPAR[(P.rp_pmu + k) * RBMAX + r]);
            if (P.prior_full)
              for (int j = 0; j < k; ++j) acc = fmaf(pL(0, k, j), EPSC[j * LDP + p], acc);
            ZD[k * LDP + p] = acc;
          }
        }
        for (int j = 0; j < P.nd_p; ++j) ZXIN[(P.nz_x + j) * LDP + p] = ROWRAW[P.idx_c_phys[j] * RBMAX + r];
        // conditional priors p(zc|c), p(zy|y): MultivariateNormal(loc, scale_tril).log_prob (models/vae.py:202-203) with
        // t = L^-1 (z - loc) by forward substitution (diagonal L for the default FactorizedNN priors)
        float lpc = 0.0f, lpy = 0.0f;
        for (int which = 0; which < 2; ++which) {
          const int k0 = which ? P.nz_c : 0, nzk = which ? P.nz_y : P.nz_c;
          float mh = 0.0f, hl = 0.0f, tv[MAXZ];
          for (int i = 0; i < nzk; ++i) {
            const float sg = ROWPAR[(P.rp_psig + k0 + i) * RBMAX + r];
            float d = ZD[(k0 + i) * LDP + p] - ROWPAR[(P.rp_pmu + k0 + i) * RBMAX + r];
            if (P.prior_full)
              for (int j = 0; j < i; ++j) d = fmaf(-pL(which, i, j), tv[j], d);
            const float t = d / sg;
            tv[i] = t;
            mh = fmaf(t, t, mh);
            hl += logf(sg);
          }
          const float lp = -0.5f * ((float)nzk * LOG_2PI + mh) - hl;
          if (which) lpy = lp; else lpc = lp;
        }
        if (P.zin_x != nullptr && valid) {
          // DPIVAE.decode (models/vae.py:153-158): the caller's latents (n, B, .) replace the sampled ones
          const long long o = (long long)m * B + row0 + r;
          const int nzi = P.nz_x + P.nd_p;
          for (int k = 0; k < nzi; ++k) ZXIN[k * LDP + p] = P.zin_x[o * nzi + k];
          for (int k = 0; k < P.nz_c; ++k) ZD[k * LDP + p] = P.zin_c[o * P.nz_c + k];
          for (int k = 0; k < P.nz_y; ++k) ZD[(P.nz_c + k) * LDP + p] = P.zin_y[o * P.nz_y + k];
        }
        const float klp = dens - ((lpx + lpc) + lpy);
        SC[SC_KL * LDP + p] = valid ? klp : 0.0f;
        SC[SC_W * LDP + p] = valid ? wpair : 0.0f;
        if (valid) {
          const long long o = (long long)m * B + row0 + r;
          if (P.out.dens) P.out.dens[o] = dens;
          if (P.out.zx) for (int k = 0; k < P.nz_x; ++k) P.out.zx[o * P.nz_x + k] = ZXIN[k * LDP + p];
          if (P.out.zc) for (int k = 0; k < P.nz_c; ++k) P.out.zc[o * P.nz_c + k] = ZD[k * LDP + p];
          if (P.out.zy) for (int k = 0; k < P.nz_y; ++k) P.out.zy[o * P.nz_y + k] = ZD[(P.nz_c + k) * LDP + p];
        }
      }
      __syncthreads();
      PHASE(PH_LATENT_FWD);
      if (P.latent_only) continue;  // dpivae_encode: transform_inputs -> encode only

      // ---- auxiliary decoders c, y: forward, log-likelihood, backward ----------------------------
      gemm_fwd<ACT_RELU>(sm + P.dc.s_w0t, P.dc.ldw0, sm + P.dc.s_b0, ZD, HC, P.nz_c, 64);
      gemm_fwd<ACT_RELU>(sm + P.dy.s_w0t, P.dy.ldw0, sm + P.dy.s_b0, ZD + P.nz_c * LDP, HY, P.nz_y, 64);
      __syncthreads();
      gemm_fwd<ACT_NONE>(sm + P.dc.s_w1t, P.dc.ldw1, sm + P.dc.s_b1, HC, OC, 64, ocp);
      gemm_fwd<ACT_NONE>(sm + P.dy.s_w1t, P.dy.ldw1, sm + P.dy.s_b1, HY, OY, 64, oyp);
      __syncthreads();
      store_pairs(P.out.ch, OC, P.nd_c, q0, npairs, row0, n, B);
      store_pairs(P.out.lsc, OC + P.nd_c * LDP, P.nd_c, q0, npairs, row0, n, B);
      store_pairs(P.out.yh, OY, P.nd_y, q0, npairs, row0, n, B);
      store_pairs(P.out.lsy, OY + P.nd_y * LDP, P.nd_y, q0, npairs, row0, n, B);
      if (P.out.ch || P.out.lsc || P.out.yh || P.out.lsy) __syncthreads();
      PHASE(PH_AUX_FWD);
      if (tid < 2 * TILE) {
        const int which = tid >> 6, p = tid & (TILE - 1);
        const long long q = q0 + p;
        const bool valid = q < npairs;
        const int r = (int)((valid ? q : npairs - 1) / n);
        const int nd = which ? P.nd_y : P.nd_c;
        float* O = which ? OY : OC;
        const float aw = (which ? P.alpha_y : P.alpha_c) * SC[SC_W * LDP + p];
        float R = 0.0f;
        if (which == 0 || P.y != nullptr) {
          for (int j = 0; j < nd; ++j) {
            const float mean = O[j * LDP + p], ls = O[(nd + j) * LDP + p];
            const float val = ROWRAW[((which ? P.nd_c : 0) + j) * RBMAX + r];
            const float es = expf(ls), var = es * es, d = val - mean;
            R += -(d * d) / (2.0f * var) - ls - LOG_SQRT_2PI;
            if (P.with_grad) {
              O[j * LDP + p] = -aw * d / var;
              O[(nd + j) * LDP + p] = -aw * (d * d / var - 1.0f);
            }
          }
        }
        SC[(which ? SC_RY : SC_RC) * LDP + p] = valid ? R : 0.0f;
      }
      __syncthreads();
      PHASE(PH_AUX_LOSS);
      if (P.with_grad) {
        gemm_wgrad(HC, OC, part + P.dc.g_w1, part + P.dc.g_b1, 64, 2 * P.nd_c);
        gemm_wgrad(HY, OY, part + P.dy.g_w1, part + P.dy.g_b1, 64, 2 * P.nd_y);
        __syncthreads();
        gemm_dgrad<ACT_RELU>(sm + P.dc.s_w1t, P.dc.ldw1, OC, HC, HC, 64, ocp);
        gemm_dgrad<ACT_RELU>(sm + P.dy.s_w1t, P.dy.ldw1, OY, HY, HY, 64, oyp);
        __syncthreads();
        gemm_wgrad(ZD, HC, part + P.dc.g_w0, part + P.dc.g_b0, P.nz_c, 64);
        gemm_wgrad(ZD + P.nz_c * LDP, HY, part + P.dy.g_w0, part + P.dy.g_b0, P.nz_y, 64);
        gemm_dgrad<ACT_NONE>(sm + P.dc.s_w0t, P.dc.ldw0, HC, nullptr, DZC, pad4(P.nz_c), 64);
        gemm_dgrad<ACT_NONE>(sm + P.dy.s_w0t, P.dy.ldw0, HY, nullptr, DZY, pad4(P.nz_y), 64);
        __syncthreads();
      }
      PHASE(PH_AUX_BWD);

      // ---- physics decoder forward -----------------------------------------------------------------
      if (P.phys_kind == 0) {
        for (int e = tid; e < TILE * nzin; e += NT) {
          const int p = e & (TILE - 1), k = e >> 6;
          S0[k * LDP + p] = (ZXIN[k * LDP + p] - P.phys_in_mean[k]) / P.phys_in_std[k];
        }
        __syncthreads();
        const int nl = P.phys_n_layers;
        for (int l = 0; l < nl; ++l) {
          const float* in = l == 0 ? S0 : sm + P.s_A[l - 1];
          float* out = l == nl - 1 ? XHP : sm + P.s_A[l];
          if (l < nl - 1)
            gemm_fwd<ACT_TANH>(sm + P.pl[l].s_wt, P.pl[l].ldw, sm + P.pl[l].s_b, in, out, P.pl[l].K, pad4(P.pl[l].N));
          else
            gemm_fwd<ACT_NONE>(sm + P.pl[l].s_wt, P.pl[l].ldw, sm + P.pl[l].s_b, in, out, P.pl[l].K, pad4(P.pl[l].N));
          __syncthreads();
        }
      } else if (P.phys_kind == 1) {
        // mass_spring: x(t) = (0/w) sin(w t) + 1 * cos(w t), w = sqrt(1/m)
        for (int e = tid; e < TILE * P.nd_x; e += NT) {
          const int p = e & (TILE - 1), d = e >> 6;
          const float mass = ZXIN[p];
          const float om = sqrtf(1.0f / mass);
          const float bb = 0.0f / om;
          const float ph = om * P.grid[d];
          XHP[d * LDP + p] = bb * sinf(ph) + 1.0f * cosf(ph);
        }
      } else {
        // Euler-Bernoulli point load on x = linspace(0, 1, nd_x); I = 2e-6, L = P = 1
        for (int e = tid; e < TILE * P.nd_x; e += NT) {
          const int p = e & (TILE - 1), d = e >> 6;
          const float E = ZXIN[p] * 1e6f, a = ZXIN[LDP + p], b = 1.0f - a, xg = P.grid[d];
          const float den1 = 6.0f * E * 2e-6f * 1.0f, den2 = 6.0f * E * 2e-6f;
          float w = 1.0f * b * xg * (1.0f - b * b - xg * xg) / den1;
          if (xg > a) {
            const float t = xg - a;
            w += 1.0f * (t * t * t) / den2;
          }
          XHP[d * LDP + p] = -1000.0f * w;
        }
      }
      __syncthreads();
      PHASE(PH_PHYS_FWD);
      // ---- data-driven decoder forward (behind the GRL: identity) ------------------------------------
      gemm_fwd<ACT_RELU>(sm + P.fx.s_w0t, P.fx.ldw0, sm + P.fx.s_b0, ZD, HD, nzd, 128);
      __syncthreads();
      gemm_fwd<ACT_NONE>(sm + P.fx.s_w1t, P.fx.ldw1, sm + P.fx.s_b1, HD, XHD, 128, ndxp);
      __syncthreads();
      store_pairs(P.out.xh_p, XHP, P.nd_x, q0, npairs, row0, n, B);
      store_pairs(P.out.xh_d, XHD, P.nd_x, q0, npairs, row0, n, B);
      if (P.out.xh_p || P.out.xh_d) __syncthreads();
      PHASE(PH_DATA_FWD);

      // ---- Gaussian log-likelihood of x (raw) and its gradient ---------------------------------------
      {
        const int p = tid & (TILE - 1), prt = tid >> 6;
        const long long q = q0 + p;
        const bool valid = q < npairs;
        const int r = (int)((valid ? q : npairs - 1) / n);
        const float* xrow = ROWX + r;
        const float w = SC[SC_W * LDP + p];
        const float gx = -(P.alpha_x * w) / var_x;
        const float inv_l2 = P.has_lambda_x ? 1.0f / (P.lambda_x * P.lambda_x) : 0.0f;
        const float log_l = P.has_lambda_x ? logf(P.lambda_x) : 0.0f;
        float ssq = 0.0f, sreg = 0.0f;
        for (int d = prt; d < P.nd_x; d += 4) {
          const float xp = XHP[d * LDP + p], xd = XHD[d * LDP + p];
          const float res = xrow[d * RBMAX] - (xp + xd);
          ssq = fmaf(res, res, ssq);
          if (P.has_lambda_x) sreg += -(xd * xd) * 0.5f * inv_l2 - log_l - LOG_SQRT_2PI;
          if (P.with_grad) {
            const float g = gx * res;
            XHP[d * LDP + p] = g;
            XHD[d * LDP + p] = g + w * xd * inv_l2;
          }
        }
        SC[(SC_P0 + prt) * LDP + p] = ssq;
        SC[(SC_Q0 + prt) * LDP + p] = sreg;
      }
      __syncthreads();
      if (tid < TILE) {
        const int p = tid;
        const bool valid = q0 + p < npairs;
        const float S = ((SC[(SC_P0 + 0) * LDP + p] + SC[(SC_P0 + 1) * LDP + p]) + SC[(SC_P0 + 2) * LDP + p]) + SC[(SC_P0 + 3) * LDP + p];
        const float Q = ((SC[(SC_Q0 + 0) * LDP + p] + SC[(SC_Q0 + 1) * LDP + p]) + SC[(SC_Q0 + 2) * LDP + p]) + SC[(SC_Q0 + 3) * LDP + p];
        SC[SC_RX * LDP + p] = valid ? (-S / (2.0f * var_x) - (float)P.nd_x * (lsx + LOG_SQRT_2PI)) : 0.0f;
        SC[SC_REG * LDP + p] = valid ? Q : 0.0f;
        if (P.with_grad) {
          // d loss / d log_sigma_x: fixed-shape warp tree, one accumulator per warp (SC_LSX row, slots 0/1)
          float lv = -(P.alpha_x * SC[SC_W * LDP + p]) * (S / var_x - (float)P.nd_x);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) lv += __shfl_xor_sync(0xffffffffu, lv, off);
          if ((tid & 31) == 0) SC[SC_LSX * LDP + (tid >> 5)] += lv;
        }
      }
      __syncthreads();
      PHASE(PH_XLOSS);

      if (P.with_grad) {
        // ---- data-driven decoder backward ----------------------------------------------------------
        gemm_wgrad(HD, XHD, part + P.fx.g_w1, part + P.fx.g_b1, 128, P.nd_x);
        __syncthreads();
        gemm_dgrad<ACT_RELU>(sm + P.fx.s_w1t, P.fx.ldw1, XHD, HD, HD, 128, ndxp);
        __syncthreads();
        gemm_wgrad(ZD, HD, part + P.fx.g_w0, part + P.fx.g_b0, nzd, 128);
        gemm_dgrad<ACT_NONE>(sm + P.fx.s_w0t, P.fx.ldw0, HD, nullptr, DZD, pad4(nzd), 128);
        __syncthreads();
        PHASE(PH_DATA_BWD);

        // ---- physics decoder backward (input gradient only; surrogate weights are frozen) -----------
        if (P.phys_kind == 0) {
          const int nl = P.phys_n_layers;
          for (int l = nl - 1; l >= 0; --l) {
            const float* G = l == nl - 1 ? XHP : sm + P.s_A[l];
            if (l > 0)
              gemm_dgrad<ACT_TANH>(sm + P.pl[l].s_wt, P.pl[l].ldw, G, sm + P.s_A[l - 1], sm + P.s_A[l - 1],
                                   pad4(P.pl[l].K), pad4(P.pl[l].N));
            else
              gemm_dgrad<ACT_NONE>(sm + P.pl[l].s_wt, P.pl[l].ldw, G, nullptr, S0, pad4(P.pl[l].K), pad4(P.pl[l].N));
            __syncthreads();
          }
          for (int e = tid; e < TILE * P.nz_x; e += NT) {
            const int p = e & (TILE - 1), k = e >> 6;
            DZX[k * LDP + p] = S0[k * LDP + p] / P.phys_in_std[k];
          }
        } else {
          const int p = tid & (TILE - 1), prt = tid >> 6;
          float s0 = 0.0f, s1 = 0.0f;
          if (P.phys_kind == 1) {
            const float mass = ZXIN[p];
            const float om = sqrtf(1.0f / mass);
            const float dom = -om / (2.0f * mass);  // d sqrt(1/m) / dm
            for (int d = prt; d < P.nd_x; d += 4) {
              const float t = P.grid[d];
              s0 = fmaf(XHP[d * LDP + p], -sinf(om * t) * t * dom, s0);
            }
          } else {
            const float z0 = ZXIN[p], E = z0 * 1e6f, a = ZXIN[LDP + p], b = 1.0f - a;
            const float den = 6.0f * E * 2e-6f;
            for (int d = prt; d < P.nd_x; d += 4) {
              const float xg = P.grid[d];
              float w = b * xg * (1.0f - b * b - xg * xg) / den;
              // d w / d a  (b = 1 - a)
              float dwa = -(xg * (1.0f - b * b - xg * xg) - 2.0f * b * b * xg) / den;
              if (xg > a) {
                const float t = xg - a;
                w += (t * t * t) / den;
                dwa += -3.0f * t * t / den;
              }
              const float g = XHP[d * LDP + p];
              s0 = fmaf(g, 1000.0f * w / z0, s0);  // d(-1000 w)/d z0 = +1000 w / z0
              s1 = fmaf(g, -1000.0f * dwa, s1);
            }
          }
          SC[(SC_P0 + prt) * LDP + p] = s0;
          SC[(SC_Q0 + prt) * LDP + p] = s1;
          __syncthreads();
          if (tid < TILE) {
            DZX[p] = ((SC[(SC_P0 + 0) * LDP + p] + SC[(SC_P0 + 1) * LDP + p]) + SC[(SC_P0 + 2) * LDP + p]) + SC[(SC_P0 + 3) * LDP + p];
            if (P.nz_x > 1)
              DZX[LDP + p] = ((SC[(SC_Q0 + 0) * LDP + p] + SC[(SC_Q0 + 1) * LDP + p]) + SC[(SC_Q0 + 2) * LDP + p]) + SC[(SC_Q0 + 3) * LDP + p];
          }
        }
        __syncthreads();

        PHASE(PH_PHYS_BWD);
        // ---- latent backward: per-pair gradients w.r.t. loc / L / prior parameters --------------------
        {
          const int p = tid & (TILE - 1), prt = tid >> 6;
          const long long q = q0 + p;
          const int r = (int)((q < npairs ? q : npairs - 1) / n);
          const float bw = P.beta_x * SC[SC_W * LDP + p];
          if (!P.prior_full) {
          for (int k = prt; k < nzd; k += 4) {
            float g = -P.lambda_g0 * DZD[k * LDP + p] + (k < P.nz_c ? DZC[k * LDP + p] : DZY[(k - P.nz_c) * LDP + p]);
            const float sg = ROWPAR[(P.rp_psig + k) * RBMAX + r];
            const float t = (ZD[k * LDP + p] - ROWPAR[(P.rp_pmu + k) * RBMAX + r]) / sg;
            g += bw * t / sg;
            FEAT[(P.f_pmu + k) * LDP + p] = -bw * t / sg;
            FEAT[(P.f_psig + k) * LDP + p] = -bw * (t * t - 1.0f) / sg;
            FEAT[(P.f_loc + P.nz_x + k) * LDP + p] = g;
          }
          } else if (prt < 2) {
            // full factor L of prior `which` = prt: t = L^-1 (z - loc), s = L^-T t;  d log p / d z = -s, d / d loc = s,
            // d / d L_ij = s_i t_j - [i == j] / L_ii  (i >= j); the loss carries -bw log p
            const int which = prt, k0 = which ? P.nz_c : 0, nzk = which ? P.nz_y : P.nz_c;
            auto pLb = [&](int i, int j) -> float { return ROWPAR[(P.rp_pL + P.pl_off[which] + i * (i - 1) / 2 + j) * RBMAX + r]; };
            float tv[MAXZ], sv[MAXZ];
            for (int i = 0; i < nzk; ++i) {
              float d = ZD[(k0 + i) * LDP + p] - ROWPAR[(P.rp_pmu + k0 + i) * RBMAX + r];
              for (int j = 0; j < i; ++j) d = fmaf(-pLb(i, j), tv[j], d);
              tv[i] = d / ROWPAR[(P.rp_psig + k0 + i) * RBMAX + r];
            }
            for (int i = nzk - 1; i >= 0; --i) {
              float d = tv[i];
              for (int j = i + 1; j < nzk; ++j) d = fmaf(-pLb(j, i), sv[j], d);
              sv[i] = d / ROWPAR[(P.rp_psig + k0 + i) * RBMAX + r];
            }
            for (int i = 0; i < nzk; ++i) {
              const int k = k0 + i;
              const float sg = ROWPAR[(P.rp_psig + k) * RBMAX + r];
              float g = -P.lambda_g0 * DZD[k * LDP + p] + (which == 0 ? DZC[i * LDP + p] : DZY[i * LDP + p]);
              g += bw * sv[i];
              FEAT[(P.f_pmu + k) * LDP + p] = -bw * sv[i];
              FEAT[(P.f_psig + k) * LDP + p] = -bw * (sv[i] * tv[i] - 1.0f / sg);
              for (int j = 0; j < i; ++j) FEAT[(P.f_pL + P.pl_off[which] + i * (i - 1) / 2 + j) * LDP + p] = -bw * sv[i] * tv[j];
              FEAT[(P.f_loc + P.nz_x + k) * LDP + p] = g;
            }
          }
          for (int i = prt; i < P.nz_x; i += 4) {
            float g = DZX[i * LDP + p];
            if (P.prior_kind[i] == 1) g += bw * (ZXIN[i * LDP + p] - P.prior_a[i]) / (P.prior_b[i] * P.prior_b[i]);
            const float u = U[i * LDP + p];
            FEAT[(P.f_loc + i) * LDP + p] = g * (P.ub[i] - P.lb[i]) * u * (1.0f - u) + bw * (2.0f * u - 1.0f);
          }
        }
        __syncthreads();
        for (int e = tid; e < TILE * P.nL; e += NT) {
          const int p = e & (TILE - 1), li = e >> 6;
          const long long q = q0 + p;
          const int r = (int)((q < npairs ? q : npairs - 1) / n);
          const int b = P.L_blk[li], i = P.L_i[li], j = P.L_j[li], s = P.blk_start[b];
          float v = FEAT[(P.f_loc + s + i) * LDP + p] * EPS[(s + j) * LDP + p];
          if (i == j) v -= P.beta_x * SC[SC_W * LDP + p] / ROWPAR[(P.rp_L + li) * RBMAX + r];
          FEAT[(P.f_L + li) * LDP + p] = v;
        }
        __syncthreads();
        PHASE(PH_LATENT_BWD);
      }

      // ---- reduce the chunk's pairs over the MC axis into per-row accumulators -----------------------
      {
        const int f0 = P.with_grad ? 0 : P.n_feat;
        if (n_pow2 && !multi_chunk) {
          // n | 32: segmented warp-shuffle tree over the n consecutive pairs of each row
          const int lane = tid & 31, warp = tid >> 5;
          for (int it = warp + 2 * f0; it < 2 * n_acc; it += NT / 32) {
            const int f = it >> 1, p = (it & 1) * 32 + lane;
            const float* src = f < P.n_feat ? FEAT + f * LDP : SC + (f - P.n_feat) * LDP;
            float v = src[p];
            for (int off = n >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            const int r = p / n;
            if ((lane & (n - 1)) == 0 && r < nrows) ROWACC[f * racc_ld + r] = v;
          }
        } else {
          for (int e = tid + f0 * RBMAX; e < n_acc * RBMAX; e += NT) {
            const int f = e / RBMAX, r = e - f * RBMAX;
            if (r < nrows) {
              const long long qa = max((long long)r * n, q0);
              const long long qb = min(min((long long)(r + 1) * n, q0 + TILE), npairs);
              if (qb > qa) {
                const float* src = f < P.n_feat ? FEAT + f * LDP : SC + (f - P.n_feat) * LDP;
                float s = 0.0f;
                for (long long q = qa; q < qb; ++q) s += src[q - q0];
                if (multi_chunk) ROWACC[f * racc_ld + r] += s;
                else ROWACC[f * racc_ld + r] = s;
              }
            }
          }
        }
      }
      __syncthreads();
      PHASE(PH_ROWRED);
    }  // chunks
    if (P.latent_only) continue;

    // ---- per-row outputs -----------------------------------------------------------------------------
    if (tid < nrows) {
      const int r = tid;
      const float inv_n = 1.0f / (float)n;
      const float kl = ROWACC[(P.n_feat + SC_KL) * racc_ld + r] * inv_n;
      const float rx = ROWACC[(P.n_feat + SC_RX) * racc_ld + r] * inv_n;
      const float rc = ROWACC[(P.n_feat + SC_RC) * racc_ld + r] * inv_n;
      const float ry = ROWACC[(P.n_feat + SC_RY) * racc_ld + r] * inv_n;
      const float rg = ROWACC[(P.n_feat + SC_REG) * racc_ld + r] * inv_n;
      const float loss = P.beta_x * kl - P.alpha_x * rx - P.alpha_c * rc - P.alpha_y * ry - rg;
      if (P.out.row_loss) {
        float* o = P.out.row_loss + row0 + r;
        o[0] = loss; o[B] = kl; o[2 * B] = rx; o[3 * B] = rc; o[4 * B] = ry; o[5 * B] = rg;
      }
      ROWACC[(P.n_feat + SC_KL) * racc_ld + r] = kl;
      ROWACC[(P.n_feat + SC_RX) * racc_ld + r] = rx;
      ROWACC[(P.n_feat + SC_RC) * racc_ld + r] = rc;
      ROWACC[(P.n_feat + SC_RY) * racc_ld + r] = ry;
      ROWACC[(P.n_feat + SC_REG) * racc_ld + r] = rg;
      SC[SC_P0 * LDP + r] = loss;
    }
    if (P.with_grad) {
      // gradients w.r.t. the head pre-activations (clamp / exp chain rule, models/encoders.py:35-43)
      for (int e = tid; e < RBMAX * P.Z; e += NT) {
        const int i = e / RBMAX, r = e - i * RBMAX;
        if (r < nrows) {
          const long long lrow = row0 + r;
          const int b = block_of(P, i), il = i - P.blk_start[b], nzb = P.blk_size[b];
          const long long om = (long long)(P.henc[b] + il) * B + lrow;
          const float pm = P.headpre[om];
          P.gpre[om] = (pm >= -50.0f && pm <= 50.0f) ? ROWACC[(P.f_loc + i) * racc_ld + r] : 0.0f;
          // f_cov entries of latent row il: strict lower triangle gets the L gradient, the rest zero
          for (int j = 0; j < nzb; ++j) {
            const long long oc = (long long)(P.henc[b] + 2 * nzb + il * nzb + j) * B + lrow;
            float g = 0.0f;
            if (j < il) {
              const float pc = P.headpre[oc];
              const int li = P.blk_loff[b] + il * (il + 1) / 2 + j;
              g = (pc >= -20.0f && pc <= 20.0f) ? ROWACC[(P.f_L + li) * racc_ld + r] : 0.0f;
            }
            P.gpre[oc] = g;
          }
          const long long os = (long long)(P.henc[b] + nzb + il) * B + lrow;
          const float ps = P.headpre[os];
          const int ld = P.blk_loff[b] + il * (il + 1) / 2 + il;
          P.gpre[os] = (ps >= -7.0f && ps <= 3.0f) ? ROWACC[(P.f_L + ld) * racc_ld + r] * expf(ps) : 0.0f;
        }
      }
      for (int e = tid; e < RBMAX * nzd; e += NT) {
        const int k = e / RBMAX, r = e - k * RBMAX;
        if (r < nrows) {
          const long long lrow = row0 + r;
          const int which = k < P.nz_c ? 0 : 1;
          const int kk = which ? k - P.nz_c : k, nzk = which ? P.nz_y : P.nz_c;
          const long long om = (long long)(P.hpri[which] + kk) * B + lrow;
          const long long os = (long long)(P.hpri[which] + nzk + kk) * B + lrow;
          const float pm = P.headpre[om], ps = P.headpre[os];
          P.gpre[om] = (pm >= -50.0f && pm <= 50.0f) ? ROWACC[(P.f_pmu + k) * racc_ld + r] : 0.0f;
          P.gpre[os] = (ps >= -7.0f && ps <= 3.0f) ? ROWACC[(P.f_psig + k) * racc_ld + r] * expf(ps) : 0.0f;
          if (P.prior_full) {
            // f_cov entries of factor row kk: the strict lower triangle gets the L gradient (through the clamp), the rest zero
            for (int j = 0; j < nzk; ++j) {
              const long long oc = (long long)(P.hpri[which] + 2 * nzk + kk * nzk + j) * B + lrow;
              float g = 0.0f;
              if (j < kk) {
                const float pc = P.headpre[oc];
                g = (pc >= -20.0f && pc <= 20.0f) ? ROWACC[(P.f_pL + P.pl_off[which] + kk * (kk - 1) / 2 + j) * racc_ld + r] : 0.0f;
              }
              P.gpre[oc] = g;
            }
          }
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      for (int r = 0; r < nrows; ++r) {
        tot[0] += SC[SC_P0 * LDP + r];
        tot[1] += ROWACC[(P.n_feat + SC_KL) * racc_ld + r];
        tot[2] += ROWACC[(P.n_feat + SC_RX) * racc_ld + r];
        tot[3] += ROWACC[(P.n_feat + SC_RC) * racc_ld + r];
        tot[4] += ROWACC[(P.n_feat + SC_RY) * racc_ld + r];
        tot[5] += ROWACC[(P.n_feat + SC_REG) * racc_ld + r];
      }
    }
    __syncthreads();
    PHASE(PH_ROWOUT);
  }  // row blocks

  if (tid == 0) {
    for (int k = 0; k < 6; ++k) part[P.n_params + k] = tot[k];
    if (P.with_grad) part[P.g_lsx] = SC[SC_LSX * LDP] + SC[SC_LSX * LDP + 1];
    if (P.phase != nullptr)
      for (int k = 0; k < PH_COUNT; ++k) atomicAdd(reinterpret_cast<unsigned long long*>(P.phase) + k, (unsigned long long)PHS[k]);
  }
}

size_t dec_smem_bytes(const DecParams& p) { return (size_t)p.s_total * sizeof(float); }

void launch_dec(const DecParams& p, int grid, cudaStream_t s) {
  dec_kernel<<<grid, NT, dec_smem_bytes(p), s>>>(p);
}

int configure_dec_kernel() {
  return (int)cudaFuncSetAttribute(dec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
}

}  // namespace dpv
