// Encoder MLPs on the tensor cores (sm_100a, tcgen05 / TMEM): the FullCovarianceNN encoders of
// models/encoders.py:7-44 -- hidden layer ReLU(W0 x_t + b0) and the concatenated heads [f_mean ; f_sigma ; f_cov] --
// for all encoder units at once (P: three 64-wide units side by side, S: one 128-wide unit), 128 minibatch rows per
// tile = the M dimension of tcgen05.mma.  Same fp16 hi/lo operand split as the decoder kernel (tc.cuh), so the
// results carry fp32 accuracy.
//
//   enc_tc_fwd_kernel: x tile -> standardise (utils/transforms.py:70-73) -> X8 operand (+ constant-one column that
//       carries b0) -> MMA (N = all hidden units) -> ReLU epilogue writes the hidden activations as PACKED fp16
//       hi/lo planes straight back to tensor memory, from where the head MMA reads its A operand (no shared-memory
//       round trip), and -- for training -- to global memory in the X8 layout the backward kernel bulk-copies ->
//       head MMA (block-diagonal W1) -> head pre-activations `headpre` [O_tot][B].
//   enc_tc_bwd_kernel: gpre tile + hidden record + x tile -> wgrad of the heads, dgrad through the heads with the
//       ReLU mask, wgrad of the first layers; weight gradients accumulate in TMEM across all tiles of the CTA.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc.cuh"
#include "enc_tc_setup.cuh"

namespace dpv {

namespace {

constexpr int TP = 128;

}  // namespace

// NW warps: 8 (two column halves per row) or 16 (four column quarters; pays when the per-tile epilogues are long: K0 = 64
// inputs and head columns in multiples of 32, i.e. the bridge / damped_oscillator P presets)
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) enc_tc_fwd_kernel(const __grid_constant__ EncTcParams P) {
  constexpr int ENTF = NW * 32, CS = NW / 4;   // threads, column splits per row
  extern __shared__ __align__(1024) unsigned char smb[];
  pdl_launch_dependents();   // the prior-net forward kernel is independent of this one and runs alongside it
  float* smf = reinterpret_cast<float*>(smb);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 16 warps: thread (TMEM lane quadrant q, lane, column quarter qc) owns row p = 32 q + lane of the tile and one quarter of
  // the columns in every epilogue (four warps per scheduler hide the TMEM / shared-memory / global latencies of the serial
  // per-tile chain; with two column halves on 8 warps the issue slots were 32 % used)
  const int q = warp & 3, qc = warp >> 2;
  const int p = 32 * q + lane;
  const int K0 = P.K0, KX = P.KX, Hc = P.Hc, Oc = P.Oc;
  const long long B = P.B;
  float* B1 = smf + (P.f_b1 >> 2);
  float* RED = smf + (P.f_red >> 2);
  int* OROW = reinterpret_cast<int*>(smb + P.f_orow);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smb + P.o_bar);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smb + P.o_bar + 8);

  for (int e = tid; e < (P.total >> 4); e += ENTF) reinterpret_cast<uint4*>(smb)[e] = make_uint4(0u, 0u, 0u, 0u);   // total is a multiple of 128
  __syncthreads();
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tc::tmem_alloc(tptr, 512);

  int k_w0, k_w1;
  enc_tc_stage_fwd<ENTF>(P, smb, k_w0, k_w1);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  // one CTA per SM owns all 512 tensor-memory columns: the base is 0, kept as a compile-time constant so that the
  // MMA operands of the warp-wide issue are provably uniform
  if (*tptr != 0u) __trap();
  constexpr uint32_t tb = 0u;
  const uint32_t el = tc::elect_one();
  const uint32_t trow = tb + ((uint32_t)(32 * q) << 16);
  const uint32_t C_H = 0, C_A = (uint32_t)Hc, C_O = (uint32_t)(2 * Hc);   // accumulator, packed hidden operand, heads
  uint32_t phase = 0;
  const float inv0 = exp2f(-(float)(k_w0 + E_X)), inv1 = exp2f(-(float)(k_w1 + E_HID));
  const float s_x = exp2f((float)E_X), s_h = exp2f((float)E_HID);
  unsigned char* pX = smb + P.a_x;
  tc::Op oX, oW0, oW1;
  oX.base = tc::smem_u32(pX); oX.lo_off = P.l_x; oX.R = TP;
  oW0.base = tc::smem_u32(smb + P.w_0); oW0.lo_off = P.l_0; oW0.R = Hc;
  oW1.base = tc::smem_u32(smb + P.w_1); oW1.lo_off = P.l_1; oW1.R = Oc;
  const int hcols = Hc / CS;                    // hidden columns per thread (multiple of 8)
  const int osplit = (Oc % (8 * CS)) == 0 ? CS : 2;    // head columns per thread: multiples of 8 (else halves, on 8 of the warps)
  const int ocols = Oc / osplit;
  const long long ntiles = (B + TP - 1) / TP;

  // The raw x rows of a tile are fetched one tile AHEAD into registers (the gathered global loads were the largest
  // stall of this kernel: nothing else hides their latency inside the serial per-tile chain); the per-column scaler
  // statistics of this thread's 4 x 8 columns live in registers for the whole kernel.
  constexpr int XC = 8 / CS;            // chunks of 8 columns per thread, K0 <= 64
  const int nch = K0 / (8 * CS);
  float4 xa[XC], xb[XC];
  float mr[XC][8], sr[XC][8];           // mean and RECIPROCAL standard deviation (one multiply per element instead of a division)
#pragma unroll
  for (int c = 0; c < XC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = qc * (K0 / CS) + 8 * c + i;
      mr[c][i] = (c < nch) ? P.mean_x[k] : 0.0f;
      sr[c][i] = (c < nch) ? 1.0f / P.std_x[k] : 1.0f;
    }
  auto fetch_x = [&](long long tile) {
    const long long lr = tile * TP + p;
    const bool ok = tile < ntiles && lr < B;
    const long long drow = ok ? (P.idx ? P.idx[lr] : lr) : 0;
    const float* xr = P.x + drow * K0 + qc * (K0 / CS);
#pragma unroll
    for (int c = 0; c < XC; ++c) {
      xa[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      xb[c] = xa[c];
      if (ok && c < nch) {
        xa[c] = __ldg(reinterpret_cast<const float4*>(xr + 8 * c));
        xb[c] = __ldg(reinterpret_cast<const float4*>(xr + 8 * c) + 1);
      }
    }
  };
  fetch_x(blockIdx.x);

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TP;
    const long long lrow = row0 + p;
    const bool valid = lrow < B;
    // ---- x tile: standardise, scale, split -> X8 operand (thread = row x half of the columns) ----
    {
#pragma unroll
      for (int c = 0; c < XC; ++c) {
        if (c < nch) {
          float v[8] = {xa[c].x, xa[c].y, xa[c].z, xa[c].w, xb[c].x, xb[c].y, xb[c].z, xb[c].w};
          const int k0 = qc * (K0 / CS) + 8 * c;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float t = P.x_is_standardised ? v[i] : (v[i] - mr[c][i]) * sr[c][i];
            // |standardised input| > 3750 sigma would overflow the fp16 operand (inf -> NaN gradients): saturate instead
            v[i] = valid ? fminf(fmaxf(t * s_x, -60000.0f), 60000.0f) : 0.0f;
          }
          put8e(pX, P.l_x, TP, k0 >> 3, p, v);
        }
      }
      fetch_x(tile + gridDim.x);   // next tile's rows: in flight under this tile's MMAs and epilogues
      if (qc == 0) {  // constant-one column (bias of the first layers) + zero padding up to KX
        float v[8] = {s_x, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        put8e(pX, P.l_x, TP, K0 >> 3, p, v);
      } else if (qc == 1) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        put8e(pX, P.l_x, TP, (K0 >> 3) + 1, p, v);
      }
    }
    // ---- hidden layers of all units: one MMA series, N = Hc ----
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) {   // warp-wide issue on the uniform datapath (tc.cuh), MMAs predicated on the elected lane
      tc::fence_after_sync();
      tc::issue_fwd_w(el, C_H, oX, oW0, Hc, KX, 0, 3);
      tc::commit_w(el, bar);
    }
    __syncwarp();
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    __syncwarp();
    tc::fence_after_sync();
    // ---- ReLU epilogue: packed fp16 hi / lo planes -> tensor memory (A operand of the heads) [+ global record] ----
    {
      unsigned char* hrec = P.hidrec ? P.hidrec + (long long)tile * P.hid_stride : nullptr;
      for (int c = 0; c < (hcols >> 3); ++c) {
        const int k0 = qc * hcols + 8 * c;
        float v[8];
        tc::tmem_ld8(trow + C_H + k0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fminf(fmaxf(v[i] * inv0, 0.0f) * s_h, 60000.0f);   // saturate instead of overflowing the fp16 operand
        uint4 hi, lo;
        tc::split8(v, hi, lo);
        float ph[4] = {__uint_as_float(hi.x), __uint_as_float(hi.y), __uint_as_float(hi.z), __uint_as_float(hi.w)};
        float pl[4] = {__uint_as_float(lo.x), __uint_as_float(lo.y), __uint_as_float(lo.z), __uint_as_float(lo.w)};
        tc::tmem_st4(trow + C_A + (k0 >> 1), ph);
        tc::tmem_st4(trow + C_A + (Hc >> 1) + (k0 >> 1), pl);
        if (hrec) {
          *reinterpret_cast<uint4*>(hrec + ((size_t)(k0 >> 3) * TP + p) * 16) = hi;
          *reinterpret_cast<uint4*>(hrec + P.hid_lo + ((size_t)(k0 >> 3) * TP + p) * 16) = lo;
        }
      }
    }
    // ---- heads: A from tensor memory ----
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_sync();
      tc::issue_fwd_ts_w(el, C_O, C_A, oW1, Oc, Hc, 0, 3);
      tc::commit_w(el, bar);
    }
    __syncwarp();
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    __syncwarp();
    tc::fence_after_sync();
    // ---- head pre-activations out: headpre[row(o)][B], coalesced along the minibatch rows ----
    for (int c = 0; c < (qc < osplit ? (ocols >> 3) : 0); ++c) {
      const int o0 = qc * ocols + 8 * c;
      float v[8];
      tc::tmem_ld8(trow + C_O + o0, v);
      if (valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = OROW[o0 + i];
          if (r >= 0) P.headpre[(long long)r * B + lrow] = v[i] * inv1 + B1[o0 + i];
        }
      }
    }
    tc::fence_before_sync();
    __syncthreads();   // X operand / TMEM columns are reused by the next tile
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}


// ------------------------------------------------------------------------------------------------------------
// Backward of the encoder units: per 128-row tile
//   G   = gpre tile (gradients w.r.t. the head pre-activations, written by lat_bwd_kernel), scaled by 2^e_g, X8 hi/lo
//   HID = hidden activations of the forward (bulk-copied record, already X8 hi/lo, scale 2^E_HID)
//   wgrad heads : D_W1[chunk][k][o] += sum_r HID[r][k] G[r][o]            (A, B MN-major; hidden in chunks of 128)
//   dgrad heads : GH[r][k] = sum_o G[r][o] W1[o][k], masked by HID > 0    (epilogue overwrites HID in place)
//   wgrad layer0: D_W0[chunk][k][j] += sum_r GH[r][k] X[r][j]             (constant-one column of X -> bias gradient)
// Weight gradients stay in tensor memory across all tiles of the CTA.
// ------------------------------------------------------------------------------------------------------------
// NPASS = 2 (head width Oc in (64, 128], the single-encoder S presets): the head-gradient operand G is staged and consumed in two
// column passes of <= 64 columns through ONE 32 KB operand buffer -- head wgrad per pass into its own tensor-memory
// columns, head dgrad accumulated over the passes (the contraction over the head outputs is split) -- because
// W1 (64 KB) + hidden record (64 KB) + a full-width G (64 KB) + X (40 KB) do not fit in 227 KB.  Per-column constants
// of the standardisation live in shared memory in that instantiation (the registers hold the second pass's prefetch).
template <int NPASS>
__global__ void __launch_bounds__(ENT, 1) enc_tc_bwd_kernel(const __grid_constant__ EncTcParams P) {
  extern __shared__ __align__(1024) unsigned char smb[];
  float* smf = reinterpret_cast<float*>(smb);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = warp & 3, hh = warp >> 2;
  const int p = 32 * q + lane;
  const int K0 = P.K0, KX = P.KX, Hc = P.Hc, Oc = P.Oc;
  const long long B = P.B;
  float* RED = smf + (P.fb_red >> 2);
  int* OROW = reinterpret_cast<int*>(smb + P.fb_orow);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smb + P.ob_bar);
  uint64_t* hbar = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smb + P.ob_bar + 16);
  float* part = P.part + (long long)blockIdx.x * P.part_stride;

  for (int e = tid; e < (P.total_b >> 4); e += ENT) reinterpret_cast<uint4*>(smb)[e] = make_uint4(0u, 0u, 0u, 0u);   // total_b is a multiple of 128
  __syncthreads();
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_init(hbar, 1);
    tc::mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tc::tmem_alloc(tptr, 512);
  const float* prm = P.params;
  auto unit_of_h = [&](int n, int& local) -> int {
    int u = 0;
    while (u + 1 < P.n_units && n >= P.h_off[u + 1]) ++u;
    local = n - P.h_off[u];
    return u;
  };
  auto unit_of_o = [&](int o, int& local) -> int {
    for (int u = 0; u < P.n_units; ++u)
      if (o >= P.o_off[u] && o < P.o_off[u] + P.O[u]) {
        local = o - P.o_off[u];
        return u;
      }
    local = 0;
    return -1;
  };
  // raw head weights of every unit: one coalesced copy into shared memory (see the forward kernel)
  const float* wb[3] = {prm + P.g_w1[0], prm + P.g_w1[P.n_units > 1 ? 1 : 0], prm + P.g_w1[P.n_units > 2 ? 2 : 0]};
  if (P.fb_raw >= 0) {
    float* RAW = smf + (P.fb_raw >> 2);
    int ro = 0;
    for (int u = 0; u < P.n_units; ++u) {
      const int nblk = P.O[u] * P.H[u];
      const float* src = prm + P.g_w1[u];
      copy_g2s_batched<ENT>(RAW + ro, src, nblk);
      wb[u] = RAW + ro;
      ro += nblk;
    }
    __syncthreads();
  }
  float m1 = 0.0f, m_unused = 0.0f;
  for (int u = 0; u < P.n_units; ++u) {
    const int n1 = P.O[u] * P.H[u];
    for (int e = tid; e < n1; e += ENT) m1 = fmaxf(m1, fabsf(wb[u][e]));
  }
  block_max2(m1, m_unused, RED);
  const int k_w1 = scale_exp_e(m1);
  {
    const float s1 = exp2f((float)k_w1);
    for (int u = 0; u < P.n_units; ++u) {
      const int nch1 = P.H[u] >> 3;
      for (int e = tid; e < P.O[u] * nch1; e += ENT) {
        const int lo = e / nch1, ch = e - lo * nch1;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = wb[u][lo * P.H[u] + 8 * ch + i] * s1;
        put8e(smb + P.wb_1, P.lb_1, Oc, (P.h_off[u] >> 3) + ch, P.o_off[u] + lo, v);
      }
    }
  }
  for (int o = tid; o < Oc; o += ENT) {
    int l;
    const int u = unit_of_o(o, l);
    OROW[o] = u >= 0 ? P.out_row[u] + l : -1;
  }
  // zero this CTA's partial gradients of the encoder units
  for (int u = 0; u < P.n_units; ++u) {
    for (long long e = tid; e < (long long)K0 * P.H[u]; e += ENT) part[P.g_w0[u] + e] = 0.0f;
    for (int e = tid; e < P.H[u]; e += ENT) part[P.g_b0[u] + e] = 0.0f;
    for (long long e = tid; e < (long long)P.H[u] * P.O[u]; e += ENT) part[P.g_w1[u] + e] = 0.0f;
    for (int e = tid; e < P.O[u]; e += ENT) part[P.g_b1[u] + e] = 0.0f;
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  // one CTA per SM owns all 512 tensor-memory columns: the base is 0, kept as a compile-time constant so that the
  // MMA operands of the warp-wide issue are provably uniform
  if (*tptr != 0u) __trap();
  constexpr uint32_t tb = 0u;
  const uint32_t el = tc::elect_one();
  const uint32_t trow = tb + ((uint32_t)(32 * q) << 16);
  const int nchunk = Hc > 128 ? 2 : 1;
  const int ch1 = Hc - 128;                       // first hidden column of the second (overlapping) chunk
  const uint32_t C_GH = 0, C_W1 = (uint32_t)Hc, C_W0 = C_W1 + (uint32_t)(nchunk * Oc);
  uint32_t phase = 0, hphase = 0;
  int e_g = P.e_g;   // refined below once lat_bwd's measured maximum is visible
  const float inv1d = exp2f(-(float)k_w1);
  const float s_x = exp2f((float)E_X);
  unsigned char* pX = smb + P.ab_x;
  unsigned char* pG = smb + P.ab_g;
  unsigned char* pH = smb + P.ab_h;
  tc::Op oX, oG, oH, oW1;
  oX.base = tc::smem_u32(pX); oX.lo_off = P.lb_x; oX.R = TP;
  oG.base = tc::smem_u32(pG); oG.lo_off = P.lb_g; oG.R = TP;
  oH.base = tc::smem_u32(pH); oH.lo_off = P.hid_lo; oH.R = TP;
  oW1.base = tc::smem_u32(smb + P.wb_1); oW1.lo_off = P.lb_1; oW1.R = Oc;
  const int hcols = Hc >> 1, fcols = Oc >> 1;
  // G passes: pass s covers head columns [64 s, 64 s + OW_s); thread (row p, half hh) owns OW_s / 2 of them
  const int ow0 = NPASS == 2 ? 64 : Oc, ow1 = NPASS == 2 ? Oc - 64 : 0;
  const int ocols = ow0 >> 1, ocols1 = ow1 >> 1;
  const long long ntiles = (B + TP - 1) / TP;
  float db1[32];
  float db2[NPASS == 2 ? 32 : 1];
#pragma unroll
  for (int i = 0; i < 32; ++i) db1[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < (NPASS == 2 ? 32 : 1); ++i) db2[i] = 0.0f;
  uint32_t wacc = 0;
  const uint32_t hid_bytes = (uint32_t)P.hid_stride;
  if (tid == 0 && (long long)blockIdx.x < ntiles) {
    tc::mbar_expect_tx(hbar, hid_bytes);
    tc::bulk_g2s(pH, P.hidrec + (long long)blockIdx.x * P.hid_stride, hid_bytes, hbar);
  }

  // inputs of a tile (raw x rows, gradients of the head pre-activations) are fetched one tile ahead into registers
  constexpr int XC = 4;
  const int nch = K0 >> 4;
  float4 xa[XC], xb[XC];
  float gq[32];
  float gq2[NPASS == 2 ? 32 : 1];
  float mr[NPASS == 2 ? 1 : XC][8], sr[NPASS == 2 ? 1 : XC][8];
  if constexpr (NPASS == 1) {
#pragma unroll
    for (int c = 0; c < XC; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = hh * (K0 >> 1) + 8 * c + i;
        mr[c][i] = (c < nch) ? P.mean_x[k] : 0.0f;
        sr[c][i] = (c < nch) ? P.std_x[k] : 1.0f;
      }
  }
  // second-pass head gradients (NPASS = 2): consumed late in a tile, so the next tile's are fetched only after that
  auto fetch_g2 = [&](long long tile) {
    if constexpr (NPASS == 2) {
      const long long lr = tile * TP + p;
      const bool ok = tile < ntiles && lr < B;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          gq2[c * 8 + i] = 0.0f;
          if (ok && c < (ocols1 >> 3)) {
            const int r = OROW[64 + hh * ocols1 + 8 * c + i];
            if (r >= 0) gq2[c * 8 + i] = P.gpre[(long long)r * B + lr];
          }
        }
    }
  };
  auto fetch_in = [&](long long tile) {
    const long long lr = tile * TP + p;
    const bool ok = tile < ntiles && lr < B;
    const long long drow = ok ? (P.idx ? P.idx[lr] : lr) : 0;
    const float* xr = P.x + drow * K0 + hh * (K0 >> 1);
#pragma unroll
    for (int c = 0; c < XC; ++c) {
      xa[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      xb[c] = xa[c];
      if (ok && c < nch) {
        xa[c] = __ldg(reinterpret_cast<const float4*>(xr + 8 * c));
        xb[c] = __ldg(reinterpret_cast<const float4*>(xr + 8 * c) + 1);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        gq[c * 8 + i] = 0.0f;
        if (ok && c < (ocols >> 3)) {
          const int r = OROW[hh * ocols + 8 * c + i];
          if (r >= 0) gq[c * 8 + i] = P.gpre[(long long)r * B + lr];
        }
      }
  };
  // the setup above read only the parameters; the head gradients fetched from here on come from lat_bwd (PDL)
  pdl_wait();
  // lat_bwd is complete from here on: the prior-net backward kernel (independent of this one) may run alongside
  pdl_launch_dependents();
  if (P.gpre_max != nullptr) {
    // The static scale (typical head gradients ~ 16 in fp16 units) is kept unless the measured max |gpre| of this batch
    // would leave the fp16 range with it: then the scale drops just far enough that the largest head gradient lands
    // in [2^13, 2^14) -- nothing saturates (two binades of headroom for the dgrad through the head weights); the bulk
    // of the rows loses low-order bits instead of a few rows contributing garbage.
    const uint32_t gb = __ldcg(P.gpre_max);
    if (gb != 0u) e_g = max(-100, min(e_g, 13 - ((int)((gb >> 23) & 0xFFu) - 127)));
  }
  const float sgp = exp2f((float)e_g);
  fetch_in(blockIdx.x);
  fetch_g2(blockIdx.x);

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TP;
    const long long lrow = row0 + p;
    const bool valid = lrow < B;
    // ---- G operand: gpre rows of the encoder heads for this row (prefetched), scaled; running sums for the head-bias
    //      gradients ----
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < (ocols >> 3)) {
        const int o0 = hh * ocols + 8 * c;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // head gradients of outlier rows can exceed the fp16 range after scaling (inf - inf = NaN in the split): saturate
          v[i] = valid ? fminf(fmaxf(gq[c * 8 + i] * sgp, -60000.0f), 60000.0f) : 0.0f;
          db1[c * 8 + i] += v[i];
        }
        put8e(pG, P.lb_g, TP, o0 >> 3, p, v);
      }
    }
    // ---- x tile (as in the forward, prefetched) ----
    {
#pragma unroll
      for (int c = 0; c < XC; ++c) {
        if (c < nch) {
          float v[8] = {xa[c].x, xa[c].y, xa[c].z, xa[c].w, xb[c].x, xb[c].y, xb[c].z, xb[c].w};
          const int k0 = hh * (K0 >> 1) + 8 * c;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float t;
            if constexpr (NPASS == 1) t = P.x_is_standardised ? v[i] : (v[i] - mr[c][i]) / sr[c][i];
            else t = P.x_is_standardised ? v[i] : (v[i] - P.mean_x[k0 + i]) / P.std_x[k0 + i];
            // |standardised input| > 3750 sigma would overflow the fp16 operand (inf -> NaN gradients): saturate instead
            v[i] = valid ? fminf(fmaxf(t * s_x, -60000.0f), 60000.0f) : 0.0f;
          }
          put8e(pX, P.lb_x, TP, k0 >> 3, p, v);
        }
      }
      fetch_in(tile + gridDim.x);   // next tile's x rows and head gradients: in flight under this tile's MMAs
      if (hh == 0) {
        float v[8] = {valid ? s_x : 0.0f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        put8e(pX, P.lb_x, TP, K0 >> 3, p, v);
      } else {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        put8e(pX, P.lb_x, TP, (K0 >> 3) + 1, p, v);
      }
    }
    // ---- hidden record of this tile has landed? ----
    tc::mbar_wait(hbar, hphase);
    hphase ^= 1u;
    __syncwarp();
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_sync();
      for (int ck = 0; ck < nchunk; ++ck) {
        tc::Op a = oH;
        a.base += (uint32_t)(((ck ? ch1 : 0) >> 3) * TP) * 16u;
        tc::issue_wgrad_w(el, C_W1 + (uint32_t)(ck * Oc), a, oG, ow0, wacc, 3);
      }
      tc::issue_dgrad_w(el, C_GH, oG, oW1, ow0, Hc, 0, 3);
      tc::commit_w(el, bar);
    }
    __syncwarp();
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    __syncwarp();
    tc::fence_after_sync();
    if constexpr (NPASS == 2) {
      // second column pass: the MMAs that read the first pass's G planes are complete (waited above)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < (ocols1 >> 3)) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = valid ? fminf(fmaxf(gq2[c * 8 + i] * sgp, -60000.0f), 60000.0f) : 0.0f;
            db2[c * 8 + i] += v[i];
          }
          put8e(pG, P.lb_g, TP, (hh * ocols1 + 8 * c) >> 3, p, v);
        }
      }
      fetch_g2(tile + gridDim.x);   // next tile's second-pass gradients: in flight under the rest of this tile
      tc::fence_async_smem();
      tc::fence_before_sync();
      __syncthreads();
      if (warp == 0) {
        tc::fence_after_sync();
        for (int ck = 0; ck < nchunk; ++ck) {
          tc::Op a = oH;
          a.base += (uint32_t)(((ck ? ch1 : 0) >> 3) * TP) * 16u;
          tc::issue_wgrad_w(el, C_W1 + (uint32_t)(ck * Oc) + 64u, a, oG, ow1, wacc, 3);
        }
        tc::Op w2 = oW1;
        w2.base += 64u * 16u;   // head rows 64.. of every hidden chunk
        tc::issue_dgrad_w(el, C_GH, oG, w2, ow1, Hc, 1, 3);
        tc::commit_w(el, bar);
      }
      __syncwarp();
      tc::mbar_wait(bar, phase);
      phase ^= 1u;
      __syncwarp();
      tc::fence_after_sync();
    }
    // ---- ReLU mask (from the hidden record) on the dgrad result, written over the record in place ----
    for (int c = 0; c < (hcols >> 3); ++c) {
      const int k0 = hh * hcols + 8 * c;
      float g[8];
      tc::tmem_ld8(trow + C_GH + k0, g);
      const uint4 hv = *reinterpret_cast<const uint4*>(pH + ((size_t)(k0 >> 3) * TP + p) * 16);
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t hbits = (hw[i >> 1] >> ((i & 1) * 16)) & 0x7FFFu;   // |hi half| : zero <=> ReLU inactive
        g[i] = hbits != 0u ? fminf(fmaxf(g[i] * inv1d, -60000.0f), 60000.0f) : 0.0f;
      }
      put8e(pH, P.hid_lo, TP, k0 >> 3, p, g);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      tc::fence_after_sync();
      for (int ck = 0; ck < nchunk; ++ck) {
        tc::Op a = oH;
        a.base += (uint32_t)(((ck ? ch1 : 0) >> 3) * TP) * 16u;
        tc::issue_wgrad_w(el, C_W0 + (uint32_t)(ck * KX), a, oX, KX, wacc, 3);
      }
      tc::commit_w(el, bar);
    }
    __syncwarp();
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    __syncwarp();
    tc::fence_after_sync();
    wacc = 1u;
    if (tid == 0 && tile + gridDim.x < ntiles) {
      tc::fence_async_smem();
      tc::mbar_expect_tx(hbar, hid_bytes);
      tc::bulk_g2s(pH, P.hidrec + (tile + gridDim.x) * P.hid_stride, hid_bytes, hbar);
    }
    __syncthreads();
  }

  // ---- flush: weight gradients out of tensor memory (lane = hidden unit within the chunk) ----
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const float sc1 = exp2f(-(float)(E_HID + e_g)), sc0 = exp2f(-(float)(E_X + e_g));
  for (int ck = 0; ck < nchunk; ++ck) {
    const int k = (ck ? ch1 : 0) + p;                        // hidden column held by this TMEM lane
    const bool mine = nchunk == 1 || (ck == 0 ? k < 128 : k >= 128);   // the overlap is taken from chunk 0
    int lk;
    const int uk = unit_of_h(k, lk);
    // heads: columns o of unit uk
    for (int c = 0; c < (fcols >> 3); ++c) {
      const int o0 = hh * fcols + 8 * c;
      float v[8];
      tc::tmem_ld8(trow + C_W1 + (uint32_t)(ck * Oc) + o0, v);
      if (mine) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int o = o0 + i;
          if (o >= P.o_off[uk] && o < P.o_off[uk] + P.O[uk]) part[P.g_w1[uk] + (long long)(o - P.o_off[uk]) * P.H[uk] + lk] = v[i] * sc1;
        }
      }
    }
    // first layers: columns j < K0, bias in column K0
    for (int c = hh; c < (KX >> 3); c += 2) {
      float v[8];
      tc::tmem_ld8(trow + C_W0 + (uint32_t)(ck * KX) + 8 * c, v);
      if (mine) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = 8 * c + i;
          if (j < K0) part[P.g_w0[uk] + (long long)lk * K0 + j] = v[i] * sc0;
          else if (j == K0) part[P.g_b0[uk] + lk] = v[i] * sc0;
        }
      }
    }
  }
  // head-bias gradients: per-thread running sums -> fixed-order sum over the 128 row slots
  __syncthreads();
  float* R0 = reinterpret_cast<float*>(pH);
  constexpr int RS = NPASS == 2 ? 64 : 32;   // scratch floats per thread: [pass][32]
#pragma unroll
  for (int i = 0; i < 32; ++i) R0[tid * RS + i] = db1[i];
  if constexpr (NPASS == 2) {
#pragma unroll
    for (int i = 0; i < 32; ++i) R0[tid * RS + 32 + i] = db2[i];
  }
  __syncthreads();
  if (tid < Oc) {
    const int ps = (NPASS == 2 && tid >= 64) ? 1 : 0, oc = tid - 64 * ps, oh = ps ? ocols1 : ocols;   // pass, column within it, columns per half
    const int h2 = oc / oh, i = oc - h2 * oh;
    int l;
    const int u = unit_of_o(tid, l);
    if (u >= 0 && i < 32) {
      float s = 0.0f;
      for (int j = 0; j < TP; ++j) s += R0[(h2 * TP + j) * RS + 32 * ps + i];
      part[P.g_b1[u] + l] = s * exp2f(-(float)e_g);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

void launch_enc_tc_bwd(const EncTcParams& p, int grid, cudaStream_t s) {
  if (p.Oc > 64) launch_pdl(enc_tc_bwd_kernel<2>, grid, ENT, (size_t)p.total_b, s, p);
  else launch_pdl(enc_tc_bwd_kernel<1>, grid, ENT, (size_t)p.total_b, s, p);
}

void launch_enc_tc_fwd(const EncTcParams& p, int grid, cudaStream_t s) {
  if (p.K0 == 64 && (p.Oc & 31) == 0 && (p.Hc & 31) == 0) enc_tc_fwd_kernel<16><<<grid, 512, p.total, s>>>(p);
  else enc_tc_fwd_kernel<8><<<grid, 256, p.total, s>>>(p);
}
int configure_enc_tc_kernels() {
  int e = (int)cudaFuncSetAttribute(enc_tc_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (!e) e = (int)cudaFuncSetAttribute(enc_tc_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (!e) e = (int)cudaFuncSetAttribute(enc_tc_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (!e) e = (int)cudaFuncSetAttribute(enc_tc_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  return e;
}

}  // namespace dpv
