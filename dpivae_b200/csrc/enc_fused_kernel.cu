// Encode-only inference in ONE warp-specialised kernel (sm_100a, tcgen05 / TMEM): DPIVAE.encode of models/vae.py:125-151
// (transform_inputs -> FullCovarianceNN encoders, models/encoders.py:7-44 -> GaussianEncoder.sample, models/encoders.py:73-93
// -> logistic + shift/scale bijector on z_x, utils/transforms.py:97-150), BASELINE.json config 5.  The two-kernel path
// (enc_tc_fwd_kernel -> headpre in HBM -> lat_encode_kernel) ran its two MMA series and their epilogues strictly in
// sequence per 128-row tile; here the stages of consecutive tiles overlap and nothing but x (in) and z / dens (out)
// touches HBM:
//
//   front warps 0-7  : raw x rows (fetched one tile ahead into registers) -> standardise -> fp16 hi/lo X8 operand;
//                      ReLU epilogue of the hidden layer: TMEM accumulator -> packed fp16 hi/lo A operand in TMEM
//   latent warps 8-15: two sets of four (TMEM lane quadrants) taking alternate tiles: head accumulator -> clamp / exp / tril
//                      -> Philox noise (torch's CUDA normal_ stream) -> z, log q - log|det J| -> global
//   issue warp 16    : tcgen05.mma series, warp-wide issue on the uniform datapath (tc.cuh): L1(t) | head(t) | L1(t+1) ...
//
// TMEM columns: H accumulator [0, Hc) | packed hidden operand [Hc, 2 Hc) | two head accumulators [2 Hc, 2 Hc + 2 Oc).
// Synchronisation: mbarriers only (tcgen05.commit for MMA completion, plain arrives for the warp hand-overs).
#include <cuda_fp16.h>
#include <curand_kernel.h>

#include "common.cuh"
#include "enc_tc_setup.cuh"
#include "kernels.h"
#include "tc.cuh"

namespace dpv {

namespace {

constexpr int TP = 128;
constexpr int F_WARPS = 8, L_WARPS = 8;
constexpr int EFT = (F_WARPS + L_WARPS + 1) * 32;   // 544 threads
constexpr int F_THREADS = F_WARPS * 32, L_SET_THREADS = 128;

// barrier slots (8 bytes each) at the end of the shared-memory plan
enum { B_XFULL = 0, B_HFULL, B_RELU, B_OFULL0, B_OFULL1, B_OFREE0, B_OFREE1, B_COUNT };

// latent layout of a P model: blocks (x, c, y), head rows per block [mean nz | sigma nz | cov nz*nz]
template <int NX, int NC, int NY>
struct PShape {
  static constexpr int nb = 3;
  static constexpr int nz(int b) { return b == 0 ? NX : (b == 1 ? NC : NY); }
  static constexpr int hcol(int b) { return b == 0 ? 0 : (b == 1 ? 2 * NX + NX * NX : 2 * NX + NX * NX + 2 * NC + NC * NC); }
  static constexpr int n_head = 2 * NX + NX * NX + 2 * NC + NC * NC + 2 * NY + NY * NY;
};

template <int NZ>
__device__ __forceinline__ void store_vec(float* dst, const float* v) {
  if constexpr (NZ == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  else if constexpr (NZ == 2) *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
  else {
#pragma unroll
    for (int i = 0; i < NZ; ++i) dst[i] = v[i];
  }
}

template <int NZ>
__device__ __forceinline__ void load_vec(const float* src, float* v) {
  if constexpr (NZ == 4) {
    const float4 t = *reinterpret_cast<const float4*>(src);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (NZ == 2) {
    const float2 t = *reinterpret_cast<const float2*>(src);
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < NZ; ++i) v[i] = src[i];
  }
}

// One latent block of one row: head columns (TMEM) -> loc, packed lower-triangular L (diagonal exp'd), sum log diag
template <int NZ>
struct BlockPar {
  float loc[NZ], L[NZ * (NZ + 1) / 2], hld;
};
template <int NZ>
__device__ __forceinline__ void load_block(uint32_t taddr, const float* B1, float inv1, BlockPar<NZ>& bp) {
  constexpr int NCOL = 2 * NZ + NZ * NZ, NLD = (NCOL + 7) / 8;
  float hv[NLD * 8];
#pragma unroll
  for (int c = 0; c < NLD; ++c) tc::tmem_ld8(taddr + 8 * c, hv + 8 * c);
  bp.hld = 0.0f;
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    bp.loc[i] = clampf_(fmaf(hv[i], inv1, B1[i]), -50.0f, 50.0f);
#pragma unroll
    for (int j = 0; j < i; ++j) bp.L[i * (i + 1) / 2 + j] = clampf_(fmaf(hv[2 * NZ + i * NZ + j], inv1, B1[2 * NZ + i * NZ + j]), -20.0f, 20.0f);
    const float lii = expf(clampf_(fmaf(hv[NZ + i], inv1, B1[NZ + i]), -7.0f, 3.0f)) + 1e-8f;
    bp.L[i * (i + 1) / 2 + i] = lii;
    bp.hld += logf(lii);
  }
}

// z = loc + L eps of one block for MC sample m; returns log q (and, for the z_x block, minus the bijector's log-det)
template <int NZ, bool BIJ>
__device__ __forceinline__ float sample_block(const EncFusedParams& P, int b, const BlockPar<NZ>& bp, unsigned long long m,
                                              unsigned long long grow, float* z) {
  float eps[NZ];
  if (P.eps_local[b] != nullptr) {   // noise_fill_kernel's output, local (m, row) order: one vector load
    load_vec<NZ>(P.eps_local[b] + (m * (unsigned long long)P.q.B + (grow - (unsigned long long)P.row_off)) * NZ, eps);
  } else {
    const unsigned long long li0 = (m * (unsigned long long)P.Bg + grow) * NZ;
    const unsigned long long off = P.rng.ss ? P.rng.ss->philox_off[b] : P.rng.offset[b];
#pragma unroll
    for (int i = 0; i < NZ; ++i)
      eps[i] = P.rng.mode == 0 ? P.rng.eps[b][li0 + i] : philox_normal_elem(P.rng.seed, off, P.rng.grid_threads[b], li0 + i);
  }
  float ss = 0.0f, ld1 = 0.0f, ld2 = 0.0f;   // summation order of lat_encode_kernel (bitwise-equal density)
#pragma unroll
  for (int i = 0; i < NZ; ++i) {
    float acc = bp.loc[i];
#pragma unroll
    for (int j = 0; j <= i; ++j) acc = fmaf(bp.L[i * (i + 1) / 2 + j], eps[j], acc);
    ss = fmaf(eps[i], eps[i], ss);
    if constexpr (BIJ) {
      const float u = sigmoidf_(acc);
      const float a = P.ub[i] - P.lb[i];
      ld1 += acc - 2.0f * softplusf_(acc);
      ld2 += logf(fabsf(a));
      z[i] = fmaf(u, a, P.lb[i]);
    } else {
      z[i] = acc;
    }
  }
  const float lq = -0.5f * ((float)NZ * LOG_2PI + ss) - bp.hld;
  return BIJ ? lq - (ld1 + ld2) : lq;
}

}  // namespace

// Reparameterisation noise of an encode-only call, generated the way torch's normal_ kernel generates it: ONE Philox4x32-10
// evaluation + two Box-Muller pairs per FOUR elements (generator thread idx, loop iteration j -> elements (4 j + k) GT + idx,
// k = 0..3; same stream as philox_normal_elem, common.cuh, which spends one evaluation per element).  Per-element
// generation inside the fused kernel cost ~1600 of its ~2800 instructions per row.  Output in the LOCAL (m, row, i) order
// of this rank's row shard; grid.y = latent block.
__global__ void __launch_bounds__(256) noise_fill_kernel(const __grid_constant__ EncFusedParams P) {
  pdl_launch_dependents();   // the fused kernel's set-up (weight staging) runs under this kernel
  const int b = blockIdx.y;
  const unsigned int GT = P.rng.grid_threads[b];
  const unsigned long long nzb = (unsigned long long)P.nz[b];
  const unsigned long long numel = (unsigned long long)P.n_mc * (unsigned long long)P.Bg * nzb;
  const unsigned long long e = (unsigned long long)blockIdx.x * 256ull + threadIdx.x;
  const unsigned long long j = e / GT;
  const unsigned int idx = (unsigned int)(e - j * GT);
  if ((4ull * j) * GT + idx >= numel) return;
  const unsigned long long off = P.rng.ss ? P.rng.ss->philox_off[b] : P.rng.offset[b];
  const unsigned long long n = (off >> 2) + j;
  const bool whole = P.Bg == P.q.B && P.row_off == 0;   // unsharded: local order == global order
  const unsigned long long per_m = (unsigned long long)P.Bg * nzb;
  unsigned long long dst[4];
  bool any = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned long long li = (4ull * j + k) * GT + idx;
    dst[k] = ~0ull;
    if (li < numel) {
      if (whole) dst[k] = li;
      else {
        const unsigned long long m = li / per_m, rem = li - m * per_m, grow = rem / nzb, i = rem - grow * nzb;
        if (grow >= (unsigned long long)P.row_off && grow < (unsigned long long)(P.row_off + P.q.B))
          dst[k] = (m * (unsigned long long)P.q.B + (grow - (unsigned long long)P.row_off)) * nzb + i;
      }
      any = any || dst[k] != ~0ull;
    }
  }
  if (!any) return;
  const uint4 ctr = make_uint4((unsigned int)n, (unsigned int)(n >> 32), idx, 0u);
  const uint2 key = make_uint2((unsigned int)P.rng.seed, (unsigned int)(P.rng.seed >> 32));
  const uint4 r = curand_Philox4x32_10(ctr, key);
  const float2 g0 = _curand_box_muller(r.x, r.y), g1 = _curand_box_muller(r.z, r.w);
  float* out = P.eps_local[b];
  if (dst[0] != ~0ull) out[dst[0]] = g0.x;
  if (dst[1] != ~0ull) out[dst[1]] = g0.y;
  if (dst[2] != ~0ull) out[dst[2]] = g1.x;
  if (dst[3] != ~0ull) out[dst[3]] = g1.y;
}

template <class SH>
__global__ void __launch_bounds__(EFT, 1) enc_fused_kernel(const __grid_constant__ EncFusedParams P) {
  extern __shared__ __align__(1024) unsigned char smb[];
  const EncTcParams& Q = P.q;
  float* smf = reinterpret_cast<float*>(smb);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K0 = Q.K0, KX = Q.KX, Hc = Q.Hc, Oc = Q.Oc;
  const long long B = Q.B;
  const float* B1 = smf + (Q.f_b1 >> 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smb + P.o_bars);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smb + P.o_bars + 8 * B_COUNT);

  for (int e = tid; e < (Q.total >> 2); e += EFT) smf[e] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    tc::mbar_init(bars + B_XFULL, F_THREADS);
    tc::mbar_init(bars + B_HFULL, 1);
    tc::mbar_init(bars + B_RELU, F_THREADS);
    tc::mbar_init(bars + B_OFULL0, 1);
    tc::mbar_init(bars + B_OFULL1, 1);
    tc::mbar_init(bars + B_OFREE0, L_SET_THREADS);
    tc::mbar_init(bars + B_OFREE1, L_SET_THREADS);
    tc::mbar_fence_init();
  }
  __syncwarp();
  if (warp == F_WARPS + L_WARPS) tc::tmem_alloc(tptr, 512);
  int k_w0, k_w1;
  enc_tc_stage_fwd<EFT>(Q, smb, k_w0, k_w1);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (*tptr != 0u) __trap();   // one CTA per SM owns all 512 columns: base 0 as a compile-time constant (uniform MMA operands)
  pdl_wait();                  // noise_fill_kernel (when it was launched ahead of this kernel) is complete
  const uint32_t C_H = 0, C_A = (uint32_t)Hc, C_O = (uint32_t)(2 * Hc);
  const long long ntiles = (B + TP - 1) / TP;
  const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp < F_WARPS) {
    // =============================== front warps: X operand + ReLU epilogue ===============================
    const int q = warp & 3, hh = warp >> 2;   // TMEM lane quadrant, column half
    const int p = 32 * q + lane;
    const uint32_t trow = (uint32_t)(32 * q) << 16;
    const float inv0h = exp2f(-(float)(k_w0 + E_X) + (float)E_HID);   // accumulator scale -> hidden operand scale in one multiply
    const float s_x = exp2f((float)E_X);
    unsigned char* pX = smb + Q.a_x;
    constexpr int XC = 4;                 // chunks of 8 columns per thread (K0 <= 64, two column halves)
    const int nch = K0 >> 4;              // chunks of this thread's half
    float4 xa[XC], xb[XC];
    auto fetch_x = [&](long long it) {
      const long long lr = (blockIdx.x + it * gridDim.x) * TP + p;
      const bool ok = it < my_tiles && lr < B;
      const long long drow = ok ? (Q.idx ? Q.idx[lr] : lr) : 0;
      const float* xr = Q.x + drow * K0 + hh * (K0 >> 1);
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
    auto stage_x = [&](long long it) {
      const long long lr = (blockIdx.x + it * gridDim.x) * TP + p;
      const bool valid = lr < B;
#pragma unroll
      for (int c = 0; c < XC; ++c) {
        if (c < nch) {
          float v[8] = {xa[c].x, xa[c].y, xa[c].z, xa[c].w, xb[c].x, xb[c].y, xb[c].z, xb[c].w};
          const int k0 = hh * (K0 >> 1) + 8 * c;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float t = Q.x_is_standardised ? v[i] : (v[i] - Q.mean_x[k0 + i]) * P.istd_x[k0 + i];
            v[i] = valid ? fminf(fmaxf(t * s_x, -60000.0f), 60000.0f) : 0.0f;   // saturate instead of overflowing the fp16 operand
          }
          put8e(pX, Q.l_x, TP, k0 >> 3, p, v);
        }
      }
      if (hh == 0) {   // constant-one column (bias of the first layers); the padding chunk after it stays zero
        float v[8] = {s_x, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        put8e(pX, Q.l_x, TP, K0 >> 3, p, v);
      }
    };
    fetch_x(0);
    if (my_tiles > 0) {
      stage_x(0);
      tc::fence_async_smem();
      tc::mbar_arrive(bars + B_XFULL);
    }
    fetch_x(1);
    const int hcols = Hc >> 1;
    for (long long it = 0; it < my_tiles; ++it) {
      tc::mbar_wait(bars + B_HFULL, (uint32_t)(it & 1));   // L1(it) complete: H ready, X buffer free
      if (it > 0) tc::mbar_wait(bars + (((it - 1) & 1) ? B_OFULL1 : B_OFULL0), (uint32_t)(((it - 1) >> 1) & 1));   // head(it-1) complete: A free
      tc::fence_after_sync();
      for (int c = 0; c < (hcols >> 4); ++c) {
        const int k0 = hh * hcols + 16 * c;
        float v[16];
        tc::tmem_ld16(trow + C_H + k0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fminf(fmaxf(v[i] * inv0h, 0.0f), 60000.0f);
        uint4 h0, l0, h1, l1;
        tc::split8(v, h0, l0);
        tc::split8(v + 8, h1, l1);
        const float ph[8] = {__uint_as_float(h0.x), __uint_as_float(h0.y), __uint_as_float(h0.z), __uint_as_float(h0.w),
                             __uint_as_float(h1.x), __uint_as_float(h1.y), __uint_as_float(h1.z), __uint_as_float(h1.w)};
        const float pl[8] = {__uint_as_float(l0.x), __uint_as_float(l0.y), __uint_as_float(l0.z), __uint_as_float(l0.w),
                             __uint_as_float(l1.x), __uint_as_float(l1.y), __uint_as_float(l1.z), __uint_as_float(l1.w)};
        tc::tmem_st8(trow + C_A + (k0 >> 1), ph);
        tc::tmem_st8(trow + C_A + (Hc >> 1) + (k0 >> 1), pl);
      }
      tc::fence_before_sync();
      tc::mbar_arrive(bars + B_RELU);
      if (it + 1 < my_tiles) {
        stage_x(it + 1);
        tc::fence_async_smem();
        tc::mbar_arrive(bars + B_XFULL);
        fetch_x(it + 2);
      }
    }
  } else if (warp < F_WARPS + L_WARPS) {
    // =============================== latent warps: heads -> z, density ===============================
    const int q = warp & 3, set = (warp - F_WARPS) >> 2;
    const int p = 32 * q + lane;
    const uint32_t trow = ((uint32_t)(32 * q) << 16) + C_O + (uint32_t)(set * Oc);
    const float inv1 = exp2f(-(float)(k_w1 + E_HID));
    uint64_t* ofull = bars + (set ? B_OFULL1 : B_OFULL0);
    uint64_t* ofree = bars + (set ? B_OFREE1 : B_OFREE0);
    constexpr int NX = SH::nz(0), NC = SH::nz(1), NY = SH::nz(2);
    for (long long it = set; it < my_tiles; it += 2) {
      tc::mbar_wait(ofull, (uint32_t)((it >> 1) & 1));
      tc::fence_after_sync();
      BlockPar<NX> bx;
      BlockPar<NC> bc;
      BlockPar<NY> by;
      load_block<NX>(trow + SH::hcol(0), B1 + SH::hcol(0), inv1, bx);
      load_block<NC>(trow + SH::hcol(1), B1 + SH::hcol(1), inv1, bc);
      load_block<NY>(trow + SH::hcol(2), B1 + SH::hcol(2), inv1, by);
      tc::fence_before_sync();
      tc::mbar_arrive(ofree);   // every head value of this tile is in registers: the accumulator may be overwritten
      const long long r = (blockIdx.x + it * gridDim.x) * TP + p;
      if (r < B) {
        const unsigned long long grow = (unsigned long long)(P.row_off + r);
        for (int m = 0; m < P.n_mc; ++m) {
          const long long qi = (long long)m * B + r;
          float zx[NX], zc[NC], zy[NY];
          float dens = sample_block<NX, true>(P, 0, bx, (unsigned long long)m, grow, zx);
          dens += sample_block<NC, false>(P, 1, bc, (unsigned long long)m, grow, zc);
          dens += sample_block<NY, false>(P, 2, by, (unsigned long long)m, grow, zy);
          if (P.zx) store_vec<NX>(P.zx + qi * NX, zx);
          if (P.zc) store_vec<NC>(P.zc + qi * NC, zc);
          if (P.zy) store_vec<NY>(P.zy + qi * NY, zy);
          if (P.dens) P.dens[qi] = dens;
        }
      }
    }
  } else {
    // =============================== MMA issue warp ===============================
    const uint32_t el = tc::elect_one();
    tc::Op oX, oW0, oW1;
    oX.base = tc::smem_u32(smb + Q.a_x); oX.lo_off = Q.l_x; oX.R = TP;
    oW0.base = tc::smem_u32(smb + Q.w_0); oW0.lo_off = Q.l_0; oW0.R = Hc;
    oW1.base = tc::smem_u32(smb + Q.w_1); oW1.lo_off = Q.l_1; oW1.R = Oc;
    for (long long it = 0; it < my_tiles; ++it) {
      tc::mbar_wait(bars + B_XFULL, (uint32_t)(it & 1));
      tc::fence_after_sync();
      tc::issue_fwd_w(el, C_H, oX, oW0, Hc, KX, 0, Q.terms);     // H is free: relu(it-1) was waited for before head(it-1)
      tc::commit_w(el, bars + B_HFULL);
      tc::mbar_wait(bars + B_RELU, (uint32_t)(it & 1));
      if (it >= 2) tc::mbar_wait(bars + ((it & 1) ? B_OFREE1 : B_OFREE0), (uint32_t)(((it - 2) >> 1) & 1));
      tc::fence_after_sync();
      tc::issue_fwd_ts_w(el, C_O + (uint32_t)((it & 1) * Oc), C_A, oW1, Oc, Hc, 0, Q.terms);
      tc::commit_w(el, bars + ((it & 1) ? B_OFULL1 : B_OFULL0));
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == F_WARPS + L_WARPS) tc::tmem_dealloc(0u, 512);
}

using PBridge = PShape<2, 4, 4>;
using POsc = PShape<1, 4, 4>;
using PBeam = PShape<2, 2, 2>;

bool enc_fused_supports(const EncFusedParams& p) {
  const EncTcParams& q = p.q;
  if (p.model_type != 0 || q.n_units != 3 || q.K0 > 64 || (q.K0 & 15) || (q.Hc & 31) || 2 * q.Hc + 2 * q.Oc > 512) return false;
  if ((size_t)p.o_bars + 8 * B_COUNT + 8 > 232448) return false;
  int o = 0;
  for (int u = 0; u < 3; ++u) {   // head rows of unit u = head columns of block u, in order
    if (q.out_row[u] != q.o_off[u] || q.o_off[u] != o) return false;
    o += q.O[u];
  }
  auto is = [&](int nx, int nc, int ny) { return p.nz[0] == nx && p.nz[1] == nc && p.nz[2] == ny; };
  return is(2, 4, 4) || is(1, 4, 4) || is(2, 2, 2);
}
// Returns the number of launches (2 when the noise is pre-generated into p.eps_local, else 1).
int launch_enc_fused(const EncFusedParams& p, int grid, cudaStream_t s) {
  const size_t smem = (size_t)p.o_bars + 8 * B_COUNT + 8;
  int launches = 1;
  if (p.eps_local[0] != nullptr) {
    unsigned long long evals = 0;
    for (int b = 0; b < 3; ++b) {
      const unsigned long long GT = p.rng.grid_threads[b], numel = (unsigned long long)p.n_mc * (unsigned long long)p.Bg * (unsigned long long)p.nz[b];
      const unsigned long long ev = (numel + 4ull * GT - 1) / (4ull * GT) * GT;
      evals = ev > evals ? ev : evals;
    }
    noise_fill_kernel<<<dim3((unsigned)((evals + 255) / 256), 3), 256, 0, s>>>(p);
    ++launches;
  }
  if (p.nz[0] == 2 && p.nz[1] == 4) launch_pdl(enc_fused_kernel<PBridge>, grid, EFT, smem, s, p);
  else if (p.nz[0] == 1) launch_pdl(enc_fused_kernel<POsc>, grid, EFT, smem, s, p);
  else launch_pdl(enc_fused_kernel<PBeam>, grid, EFT, smem, s, p);
  return launches;
}
int configure_enc_fused_kernels() {
  int e = (int)cudaFuncSetAttribute(enc_fused_kernel<PBridge>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (!e) e = (int)cudaFuncSetAttribute(enc_fused_kernel<POsc>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (!e) e = (int)cudaFuncSetAttribute(enc_fused_kernel<PBeam>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  return e;
}

}  // namespace dpv
