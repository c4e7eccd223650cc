// Tensor-core (tcgen05 / TMEM) variant of the decoder-side fused kernel of the DPI-VAE step (sm_100a).
//
// Same contract as dec_kernel.cu (models/vae.py:160-231, models/decoders.py, models/nn.py:67-80,
// dpivae.py:419-429), but every GEMM of the decoders -- forward, dgrad and wgrad -- runs on the 5th-gen
// tensor cores:
//   * one persistent CTA per SM walks tiles of 128 (row, MC-sample) pairs = the M dimension of
//     tcgen05.mma.cta_group::1.kind::f16 (M = 128), accumulators in TMEM, operands in shared memory in
//     the dual-orientation X8 layout of tc.cuh (the same buffer is read K-major by forward / dgrad and
//     MN-major by wgrad);
//   * fp32 accuracy: operands are split x = hi + lo into two fp16 planes (power-of-two pre-scaling keeps
//     them in the fp16 normal range) and each GEMM issues hi*hi + lo*hi + hi*lo into one fp32
//     accumulator (terms = 3); terms = 1 is the plain fp16-input mode;
//   * weight gradients of the trainable decoders accumulate in TMEM across ALL tiles of the CTA and are
//     read back once at the end of the kernel (no per-tile reduction traffic); bias gradients ride on a
//     constant-one input column (first layers) or on per-thread running sums (output layers);
//   * 256 threads (8 warps) are the epilogue of the x path: thread (quadrant q, lane, half hh) owns pair p = 32 q + lane
//     and one half of the accumulator columns, applies bias / ReLU / tanh / likelihood gradients and writes the next
//     operand straight back to shared memory; one more warp issues their MMAs;
//   * the auxiliary decoders c and y (models/decoders.py:36-49, 4 -> 64 -> 4 each) run on FOUR MORE WARPS, one tile ahead
//     of the x path and off its critical chain: forward, likelihood and dgrad are register-resident fp32 math (thread =
//     pair); their weight gradients are ONE masked outer-product reduction on the tensor cores:
//         S[k][j][i] = sum_p m[p][k] g[p][j] ze[p][i]      m = ReLU mask (EXACT in fp16: a single operand plane),
//                                                          g = head gradient, ze = [z | 1]
//         dW1[j][k] = sum_i W0e[k][i] S[k][j][i],  dW0e[k][i] = sum_j W1[j][k] S[k][j][i]   (contracted once, at kernel end)
//     (h = m * (W0e ze) is linear in ze given the mask), so no hidden-activation operand is ever written: the operands
//     are the 16 KB mask plane and a 24-column product plane per side, staged in the x-residual gradient buffer while
//     that buffer is idle between two x heads (no extra shared memory), issued by the aux warps themselves as M = 64 MMAs.
#include <cuda_fp16.h>
#include <curand_kernel.h>

#include "common.cuh"
#include "kernels.h"
#include "tc.cuh"

namespace dpv {

namespace {

constexpr int TP = 128;    // pairs per tile
constexpr int TNT = 256;   // threads per CTA
// TMEM column map (fp32 columns, 128 lanes)
// C_AS: aux masked reductions S, two M = 64 accumulators of NAUX columns (side c, side y); C_T (first-layer dgrad of the
// data-driven decoder) shares the columns of C_S, which is idle by the time stage S8 is issued
constexpr int NAUX = 24;   // product columns per side: (head output j < 4) x (input i < 5: z_0..z_3, 1), padded to 3 chunks
enum { C_W1 = 0, C_W0 = 64, C_AS = 80, C_H = 128, C_X = 256, C_S = 320, C_T = 320, C_A0 = 352, C_A1 = 416, C_A2 = 448,
       C_ALLOC = 512 };
// power-of-two operand scales (exponents)
constexpr int E_LAT = 4, E_H = 6, E_T = 8;
// scalar rows (one value per pair)
enum { S_KL = 0, S_RX, S_RC, S_RY, S_KL2, S_W, S_Q0, S_Q1, S_ROWS };
// inverse-scale table
enum { I_FX0 = 0, I_P0, I_P1, I_P2, I_X, I_XD, I_FX0D, I_P2D, I_P1D, I_P0D, I_COUNT };

// optional per-phase cycle accounting (thread 0, clock64); see tools/phase_profile.py
enum { TPH_SETUP = 0, TPH_ROWPAR_EPS, TPH_LATENT, TPH_AUX1, TPH_A0, TPH_AUX2, TPH_A1, TPH_HD, TPH_A2, TPH_XHEAD, TPH_BWD1, TPH_BWD2,
       TPH_BWD3, TPH_BWD4, TPH_LATENT_BWD, TPH_ROWRED, TPH_ROWOUT, TPH_FLUSH, TPH_COUNT };
#define TPHASE(k)                                  \
  do {                                             \
    if (PROF && tid == 0) {                        \
      const long long _t = clock64();              \
      phs[k] += _t - t_last;                       \
      t_last = _t;                                 \
    }                                              \
  } while (0)


// Packed fp32 pairs (sm_100 FFMA2: two IEEE fp32 FMAs per instruction) for the CUDA-core auxiliary decoders: a pair of
// hidden units is processed at once, weights are stored pair-interleaved in shared memory.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float a, float b) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t ffma2(f2_t a, f2_t b, f2_t c) {
  f2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2_t fmul2(f2_t a, f2_t b) {
  f2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// write 8 consecutive columns (chunk) of row `row` of an X8 operand: hi plane + lo plane
__device__ __forceinline__ void put8(unsigned char* plane, uint32_t lo_off, int R, int chunk, int row, const float* v) {
  uint4 hi, lo;
  tc::split8(v, hi, lo);
  unsigned char* dst = plane + ((size_t)chunk * R + row) * 16;
  *reinterpret_cast<uint4*>(dst) = hi;
  *reinterpret_cast<uint4*>(dst + lo_off) = lo;
}

// tanh(x) = 1 - 2 / (1 + 2^(2 x log2 e)): exp2f + fast division, absolute error of a few 1e-7 (the parity
// tests of tests/test_gpu_tc.py run through it); saturates cleanly to +-1 for large |x|
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = exp2f(x * 2.8853900817779268f);
  return 1.0f - __fdividef(2.0f, 1.0f + e);
}

// element (row p, column k) of the latent operand record: hi + lo fp16 planes, undoing the 2^4 operand scale
__device__ __forceinline__ float lat_elem(const unsigned char* rec, int k, int p) {
  const size_t off = ((size_t)(k >> 3) * TP + p) * 16 + (size_t)(k & 7) * 2;
  const float h = __half2float(*reinterpret_cast<const __half*>(rec + off));
  const float l = __half2float(*reinterpret_cast<const __half*>(rec + 4096 + off));
  return (h + l) * 0.0625f;
}

template <int W>
__device__ __forceinline__ void tld(uint32_t taddr, float* v) {
  if (W == 32) tc::tmem_ld32(taddr, v);
  else tc::tmem_ld16(taddr, v);
}

__device__ __forceinline__ tc::Op mkop(const unsigned char* sm, int off, uint32_t lo_off, int R, int col0 = 0) {
  tc::Op o;
  o.base = tc::smem_u32(sm + off) + (uint32_t)((col0 >> 3) * R) * 16u;
  o.lo_off = lo_off;
  o.R = R;
  return o;
}

constexpr int NTHR = TNT + 32;    // 8 epilogue warps + 1 MMA-issue warp: participants of the stage-signal barriers
constexpr int NAUXT = 128;        // 4 auxiliary-decoder warps (thread = pair)
// Four warpgroups: 0, 1 = x-path epilogue (warps 0..7), 2 = auxiliary decoders (warps 8..11), 3 = MMA issue (warp 12; warps
// 13..15 only take part in the CTA-wide barriers).  Registers are re-balanced per warpgroup with setmaxnreg: the launch
// gives every thread 128, the epilogue warpgroups grow to R_MAIN, the other two shrink.
constexpr int NALL = 512;
constexpr int R_MAIN = 176, R_AUX = 104, R_ISSUE = 56;
constexpr int W_ISSUE = 12;       // the MMA-issue warp
// Warp-specialised MMA issue: the 256 epilogue threads only SIGNAL that the operands of stage `sid` are in place
// (non-blocking bar.arrive on a rotating named barrier); the dedicated issue warp waits for the signal, issues the
// MMAs and commits them to an mbarrier.  The epilogue warps never spend issue slots on descriptor arithmetic and never
// wait for the issue itself, only for the results they consume.
// SMEM_OPS: the stage's MMAs read shared-memory operands written by these threads (generic proxy -> async proxy fence);
// stages whose new operands live in tensor memory (TS-mode MMAs) or arrived by bulk copy skip that fence
template <bool SMEM_OPS = true>
__device__ __forceinline__ void stage_signal(uint32_t sid) {
  if (SMEM_OPS) tc::fence_async_smem();
  tc::fence_before_sync();
  asm volatile("bar.arrive %0, %1;" ::"r"(2u + (sid & 7u)), "r"((uint32_t)NTHR) : "memory");
}
__device__ __forceinline__ void issuer_wait(uint32_t sid) {
  asm volatile("bar.sync %0, %1;" ::"r"(2u + (sid & 7u)), "r"((uint32_t)NTHR) : "memory");
  tc::fence_after_sync();
}
// barrier among the 256 epilogue threads only
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"r"((uint32_t)TNT) : "memory"); }
__device__ __forceinline__ void stage_wait(uint64_t* bar, uint32_t& phase) {
  tc::mbar_wait(bar, phase);
  phase ^= 1u;
  __syncwarp();
  tc::fence_after_sync();
}

// write 4 consecutive columns (half a chunk) of an X8 operand row
__device__ __forceinline__ void put4(unsigned char* plane, uint32_t lo_off, int R, int chunk, int row, int half, const float* v) {
  const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(v[0] - f0.x, v[1] - f0.y), l1 = __floats2half2_rn(v[2] - f1.x, v[3] - f1.y);
  unsigned char* dst = plane + ((size_t)chunk * R + row) * 16 + 8 * half;
  *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  *reinterpret_cast<uint2*>(dst + lo_off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

// exponent k with max * 2^k in [256, 512)
__device__ __forceinline__ int scale_exp(float mx) {
  if (!(mx > 0.0f) || !isfinite(mx)) return 0;
  int e;
  frexpf(mx, &e);  // mx = m * 2^e, m in [0.5, 1)
  return 9 - e;
}

// stage a (N x KP) operand whose element (n, k) is get(n, k), scaled by 2^kexp, into an X8 hi/lo plane pair
template <class G>
__device__ void stage_weight(unsigned char* plane, uint32_t lo_off, int N, int KP, int kexp, G get) {
  const float s = exp2f((float)kexp);
  const int nch = KP >> 3;
  for (int e = threadIdx.x; e < nch * N; e += NALL) {
    const int ch = e / N, n = e - ch * N;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = get(n, 8 * ch + i) * s;
    put8(plane, lo_off, N, ch, n, v);
  }
}

// Dense row-major [N][K] weights (already in shared memory, 16-byte aligned rows) -> X8 hi/lo plane pair.  Consecutive threads
// take consecutive rows n (conflict-free 16-byte operand stores) and walk the K chunks in an order ROTATED by the row index,
// so that the two 16-byte loads of a quarter-warp spread over the banks (2-way instead of the 32-way conflict of eight
// scalar loads at a row stride that is a multiple of 32 words: ~10 us of the kernel's set-up).
__device__ void stage_weight_dense(unsigned char* plane, uint32_t lo_off, int N, int K, int kexp, const float* W) {
  const float s = exp2f((float)kexp);
  const int nch = K >> 3;
  for (int e = threadIdx.x; e < nch * N; e += NALL) {
    const int c0 = e / N, n = e - c0 * N;
    const int ch = (c0 + n) % nch;
    const float4 a = *reinterpret_cast<const float4*>(W + (size_t)n * K + 8 * ch);
    const float4 b = *reinterpret_cast<const float4*>(W + (size_t)n * K + 8 * ch + 4);
    const float v[8] = {a.x * s, a.y * s, a.z * s, a.w * s, b.x * s, b.y * s, b.z * s, b.w * s};
    put8(plane, lo_off, N, ch, n, v);
  }
}

}  // namespace

// One-time set-up of a CTA (all 512 threads): zero shared memory, barriers, tensor-memory allocation, operand staging of
// the weights (fp16 hi / lo planes, power-of-two scales from the block maxima), fp32 tables of the auxiliary decoders.
// Deliberately NOT inlined: its register allocation stays apart from the role loops of the kernel (inlined, the values
// live across it were spilled to local memory and reloaded inside the tile loop: +5 % kernel time).
template <int PHYS, int NDX>
static __device__ __noinline__ void dec_tc_setup(const TcParams& T, unsigned char* smb) {
  const DecParams& P = T.d;
  float* smf = reinterpret_cast<float*>(smb);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nzd = P.nz_c + P.nz_y;
  const int nzin = P.nz_x + P.nd_p;
  constexpr int ndx = NDX;
  constexpr bool mlp = PHYS == 0;
  const int d1 = mlp ? P.pl[0].N : 0, d2 = mlp ? P.pl[1].N : 0, d3 = mlp ? P.pl[2].N : 0;
  float* part = P.part + (long long)blockIdx.x * P.part_stride;
  float* INV = smf + (T.f_inv >> 2);
  float* BX = smf + (T.f_bias_x >> 2);
  float* BP1 = smf + (T.f_bias_p1 >> 2);
  float* BP2 = smf + (T.f_bias_p2 >> 2);
  float* AW0 = smf + (T.f_aw0 >> 2);
  float* AB0 = smf + (T.f_ab0 >> 2);
  float* AW1 = smf + (T.f_aw1 >> 2);
  float* AB1 = smf + (T.f_ab1 >> 2);
  float* WP0F = smf + (T.f_wp0f >> 2);
  float* RED = smf + (T.f_red >> 2);
  uint64_t* rbar = reinterpret_cast<uint64_t*>(smb + T.o_bar + 32);
  uint64_t* bar0 = reinterpret_cast<uint64_t*>(smb + T.o_bar);
  uint64_t* bar1 = bar0 + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smb + T.o_bar + 16);
  uint64_t* gfree = reinterpret_cast<uint64_t*>(smb + T.o_bar + 48);
  uint64_t* adone = gfree + 1;
  uint64_t* abar = gfree + 2;
  // ---- one-time: zero smem, barriers, TMEM, weights -------------------------------------------------
  for (int e = tid; e < (T.total >> 4); e += NALL) reinterpret_cast<uint4*>(smb)[e] = make_uint4(0u, 0u, 0u, 0u);   // T.total is a multiple of 128
  __syncthreads();
  if (tid == 0) {
    tc::mbar_init(bar0, 1);
    tc::mbar_init(bar1, 1);
    tc::mbar_init(rbar, 1);
    tc::mbar_init(rbar + 1, 1);
    tc::mbar_init(gfree, TNT);
    tc::mbar_init(adone, P.with_grad ? NAUXT + 1 : NAUXT);   // every aux thread + (training) the commit of its last MMAs
    tc::mbar_init(abar, 1);
    tc::mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tc::tmem_alloc(tptr, C_ALLOC);

  const float* prm = P.params;
  const int c1 = T.c_ones, cs0 = T.c_s0;
  // Raw weights first: the data-driven decoder's [w0 | b0 | w1 | b1] block of the flat buffer and the frozen physics
  // surrogate are copied ONCE with coalesced loads into the (still unused) operand buffers; the scale search and the
  // operand staging read them from shared memory, and the scratch is zeroed again before the first tile.
  float* FX = smf + (T.a_big >> 2);
  const int n_fx = 128 * nzd + 128 + ndx * 128 + ndx;
  const int n_fr = mlp ? (int)(P.pl[3].g_b + ndx) : 0;
  float* FR = FX + ((n_fx + 3) & ~3);
  {
    const float* src = prm + P.fx.g_w0;
    copy_g2s_batched<NALL>(FX, src, n_fx);
    if constexpr (mlp) copy_g2s_batched<NALL>(FR, P.frozen, n_fr);
  }
  __syncthreads();
  const int o_b0 = 128 * nzd, o_w1 = o_b0 + 128, o_b1 = o_w1 + ndx * 128;
  // element getters of the padded first-layer matrices (bias in the constant-one column)
  auto g_fx0 = [&](int nn, int k) -> float {
    if (k < nzd) return FX[nn * nzd + k];
    return k == c1 ? FX[o_b0 + nn] : 0.0f;
  };
  auto g_fx1 = [&](int nn, int k) -> float { return FX[o_w1 + nn * 128 + k]; };
  auto g_p0 = [&](int nn, int k) -> float {
    if (k >= cs0 && k < cs0 + nzin) return FR[P.pl[0].g_w + nn * nzin + (k - cs0)];
    return k == c1 ? FR[P.pl[0].g_b + nn] : 0.0f;
  };
  auto g_p1 = [&](int nn, int k) -> float { return FR[P.pl[1].g_w + nn * d1 + k]; };
  auto g_p2 = [&](int nn, int k) -> float { return FR[P.pl[2].g_w + nn * d2 + k]; };
  auto g_p3 = [&](int nn, int k) -> float { return FR[P.pl[3].g_w + nn * d3 + k]; };

  auto bias_p3 = [&](int e) -> float {
    if constexpr (mlp) return FR[P.pl[3].g_b + e];
    else return 0.0f;
  };
  const int KZ = T.KZ;
  // power-of-two operand scales from the block maxima: linear scans of the raw ranges, ONE block reduction for all six
  int k_fx0, k_x, k_p0 = 0, k_p1 = 0, k_p2 = 0;
  {
    float m[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int e = tid; e < o_w1; e += NALL) m[0] = fmaxf(m[0], fabsf(FX[e]));                      // fx0: w0 and b0
    for (int e = tid; e < ndx * 128; e += NALL) m[1] = fmaxf(m[1], fabsf(FX[o_w1 + e]));          // fx1: w1
    if constexpr (mlp) {
      for (int e = tid; e < d1 * nzin; e += NALL) m[2] = fmaxf(m[2], fabsf(FR[P.pl[0].g_w + e]));
      for (int e = tid; e < d1; e += NALL) m[2] = fmaxf(m[2], fabsf(FR[P.pl[0].g_b + e]));
      for (int e = tid; e < d2 * d1; e += NALL) m[3] = fmaxf(m[3], fabsf(FR[P.pl[1].g_w + e]));
      for (int e = tid; e < d3 * d2; e += NALL) m[4] = fmaxf(m[4], fabsf(FR[P.pl[2].g_w + e]));
      for (int e = tid; e < ndx * d3; e += NALL) m[5] = fmaxf(m[5], fabsf(FR[P.pl[3].g_w + e]));
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], off));
    if (lane == 0)
#pragma unroll
      for (int i = 0; i < 6; ++i) RED[warp * 6 + i] = m[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float r = RED[i];
      for (int w = 1; w < NALL / 32; ++w) r = fmaxf(r, RED[w * 6 + i]);
      m[i] = r;
    }
    k_fx0 = scale_exp(m[0]);
    k_x = scale_exp(m[1]);
    if constexpr (mlp) {
      k_p0 = scale_exp(m[2]); k_p1 = scale_exp(m[3]); k_p2 = scale_exp(m[4]);
      k_x = min(k_x, scale_exp(m[5]));  // fx1 and the last physics layer accumulate into the same TMEM columns
    }
  }
  // dense matrices: vector loads when their rows are 16-byte aligned in the scratch (they are for every shipped case)
  auto dense_ok = [&](const float* W, int K) { return ((reinterpret_cast<uintptr_t>(W) & 15u) == 0) && (K & 7) == 0; };
  stage_weight(smb + T.w_fx0, T.l_fx0, 128, KZ, k_fx0, g_fx0);
  if (dense_ok(FX + o_w1, 128)) stage_weight_dense(smb + T.w_fx1, T.l_fx1, ndx, 128, k_x, FX + o_w1);
  else stage_weight(smb + T.w_fx1, T.l_fx1, ndx, 128, k_x, g_fx1);
  if constexpr (mlp) {
    stage_weight(smb + T.w_p[0], T.l_p[0], d1, KZ, k_p0, g_p0);
    if (dense_ok(FR + P.pl[1].g_w, d1)) stage_weight_dense(smb + T.w_p[1], T.l_p[1], d2, d1, k_p1, FR + P.pl[1].g_w);
    else stage_weight(smb + T.w_p[1], T.l_p[1], d2, d1, k_p1, g_p1);
    if (dense_ok(FR + P.pl[2].g_w, d2)) stage_weight_dense(smb + T.w_p[2], T.l_p[2], d3, d2, k_p2, FR + P.pl[2].g_w);
    else stage_weight(smb + T.w_p[2], T.l_p[2], d3, d2, k_p2, g_p2);
    if (dense_ok(FR + P.pl[3].g_w, d3)) stage_weight_dense(smb + T.w_p[3], T.l_p[3], ndx, d3, k_x, FR + P.pl[3].g_w);
    else stage_weight(smb + T.w_p[3], T.l_p[3], ndx, d3, k_x, g_p3);
  }
  for (int e = tid; e < ndx; e += NALL) BX[e] = FX[o_b1 + e] + bias_p3(e);
  if constexpr (mlp) {
    for (int e = tid; e < d2; e += NALL) BP1[e] = FR[P.pl[1].g_b + e];
    for (int e = tid; e < d3; e += NALL) BP2[e] = FR[P.pl[2].g_b + e];
  }
  if constexpr (mlp) {
    for (int e = tid; e < d1 * 4; e += NALL) {
      const int k = e >> 2, j = e & 3;
      WP0F[e] = j < P.nz_x ? FR[P.pl[0].g_w + k * nzin + j] : 0.0f;
    }
  }
  // auxiliary decoders (fp32, CUDA cores): side 0 = decoder_c, side 1 = decoder_y
  for (int e = tid; e < 2 * 64 * 4; e += NALL) {
    const int side = e >> 8, k = (e >> 2) & 63, j = e & 3;
    const Mlp2S& M = side ? P.dy : P.dc;
    const int nzs = side ? P.nz_y : P.nz_c, nd = side ? P.nd_y : P.nd_c;
    const int ep = ((side * 32 + (k >> 1)) * 4 + j) * 2 + (k & 1);   // pair-interleaved slot of (side, unit k, column j)
    AW0[ep] = j < nzs ? prm[M.g_w0 + (long long)k * nzs + j] : 0.0f;
    const int o = (j & 1) < nd ? ((j >> 1) * nd + (j & 1)) : -1;  // head column j: mean_(j&1) (j < 2) or log_sigma_(j&1)
    AW1[ep] = o >= 0 ? prm[M.g_w1 + (long long)o * 64 + k] : 0.0f;
    if (j == 0) AB0[side * 64 + k] = prm[M.g_b0 + k];
    if (k == 0) AB1[side * 4 + j] = o >= 0 ? prm[M.g_b1 + o] : 0.0f;
  }

  const float lsx = prm[P.g_lsx];
  const float sx = expf(lsx);
  const float var_x = sx * sx;
  // gradient scale of the x residual: sg = 2^round(log2(64 / sigma_x)), kept inside the fp16 range
  int e_g = (int)rintf(6.0f - lsx * 1.4426950408889634f);
  e_g = max(-8, min(e_g, 24));
  const float sg = exp2f((float)e_g);
  if (tid == 0) {
    INV[I_FX0] = exp2f(-(float)(k_fx0 + E_LAT));
    INV[I_P0] = exp2f(-(float)(k_p0 + E_LAT));
    INV[I_P1] = exp2f(-(float)(k_p1 + E_T));
    INV[I_P2] = exp2f(-(float)(k_p2 + E_T));
    INV[I_X] = exp2f(-(float)(k_x + E_H));
    INV[I_XD] = exp2f(-(float)k_x);
    INV[I_FX0D] = exp2f(-(float)k_fx0);
    INV[I_P2D] = exp2f(-(float)k_p2);
    INV[I_P1D] = exp2f(-(float)k_p1);
    INV[I_P0D] = exp2f(-(float)k_p0);
  }
  if (P.with_grad && tid == 0) part[P.g_lsx] = 0.0f;
  for (int e = tid; e < NSCAL; e += NALL) part[P.n_params + e] = 0.0f;
  // the raw-weight scratch goes back to zero: the operand buffers rely on zero padding
  __syncthreads();
  for (int e = tid; e < ((n_fx + 3) & ~3) + n_fr; e += NALL) FX[e] = 0.0f;
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
}

// PHYS: physics decoder kind (0 MLP surrogate, 1 mass_spring, 2 beam); NDX: response length -- compile-time so
// that the epilogues are straight-line code (a taken branch in this large kernel costs an I-cache miss)
template <bool PROF, int PHYS, int NDX>
__global__ void __launch_bounds__(NALL, 1) dec_tc_kernel(const __grid_constant__ TcParams T) {
  const DecParams& P = T.d;
  extern __shared__ __align__(1024) unsigned char smb[];
  float* smf = reinterpret_cast<float*>(smb);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = warp & 3, hh = (warp >> 2) & 1;   // warps 0..7: x-path epilogue, 8..11: auxiliary decoders, 12: MMA issue
  const int p = 32 * q + lane;  // pair (TMEM lane) owned in the epilogues; hh = column half / aux side (0 = c, 1 = y)
  const int n = P.n_mc;
  const int nzd = P.nz_c + P.nz_y;
  const int nzin = P.nz_x + P.nd_p;
  constexpr int ndx = NDX, nxh = NDX >> 1;  // columns of the x head per thread
  const long long B = P.B;
  const int terms = T.terms;
  constexpr bool mlp = PHYS == 0;
  const int d1 = mlp ? P.pl[0].N : 0, d2 = mlp ? P.pl[1].N : 0, d3 = mlp ? P.pl[2].N : 0;
  float* part = P.part + (long long)blockIdx.x * P.part_stride;
  long long t_last = (PROF && tid == 0) ? clock64() : 0;   // phase accounting starts here: setup is phase 0
  long long phs[PROF ? TPH_COUNT : 1];
#pragma unroll
  for (int i = 0; i < (PROF ? TPH_COUNT : 1); ++i) phs[i] = 0;

  float* INV = smf + (T.f_inv >> 2);
  float* BX = smf + (T.f_bias_x >> 2);    // fx1 bias (+ last physics-layer bias)
  float* BP1 = smf + (T.f_bias_p1 >> 2);
  float* BP2 = smf + (T.f_bias_p2 >> 2);
  float* AW0 = smf + (T.f_aw0 >> 2);      // aux first layers, unit pairs interleaved [side][32 pairs][4 inputs][2 units]
  float* AB0 = smf + (T.f_ab0 >> 2);      //                   [side][64]
  float* AW1 = smf + (T.f_aw1 >> 2);      // aux heads [side][32 pairs][4 outputs: mean_0, mean_1, ls_0, ls_1][2 units]
  float* AB1 = smf + (T.f_ab1 >> 2);      //                   [side][4]
  float* WP0F = smf + (T.f_wp0f >> 2);    // physics layer 0 weights w.r.t. the physics latents, fp32 [64][4]
  const float4* WP0F4 = reinterpret_cast<const float4*>(WP0F);
  float* DZA = smf + (T.f_dza >> 2);
  float* GSX = smf + (T.f_w0f >> 2);      // [4][TP] exchange of the physics-latent gradient halves
  float* SC = smf + (T.f_sc >> 2);
  float* SCA = smf + (T.f_sca >> 2);      // [2 tile parities][2 sides][TP]: R_c, R_y per pair, written by the aux warps
  float* RED = smf + (T.f_red >> 2);
  float* R0 = smf + (T.a_big >> 2);      // end-of-kernel reduction scratch, aliases the BIG operand buffer
  unsigned char* RECB = smb + T.a_rec;   // two tile-record buffers (bulk-copied from the latent kernel's output)
  uint64_t* rbar = reinterpret_cast<uint64_t*>(smb + T.o_bar + 32);  // [2] record-arrival barriers
  uint64_t* bar0 = reinterpret_cast<uint64_t*>(smb + T.o_bar);
  uint64_t* bar1 = bar0 + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(smb + T.o_bar + 16);
  // aux warps <-> x path: gfree = the x-residual gradient buffer G is idle (every x-path thread arrives once per tile),
  // adone = the aux results of a tile (R_c / R_y, dL/dz, weight-gradient MMAs) are complete, abar = the aux MMAs of side c
  uint64_t* gfree = reinterpret_cast<uint64_t*>(smb + T.o_bar + 48);
  uint64_t* adone = gfree + 1;
  uint64_t* abar = gfree + 2;

  dec_tc_setup<PHYS, NDX>(T, smb);
  const float* prm = P.params;
  const int c1 = T.c_ones, cs0 = T.c_s0;
  const int KZ = T.KZ;
  const float lsx = prm[P.g_lsx];
  const float sx = expf(lsx);
  const float var_x = sx * sx;
  // gradient scale of the x residual: sg = 2^round(log2(64 / sigma_x)), kept inside the fp16 range
  int e_g = (int)rintf(6.0f - lsx * 1.4426950408889634f);
  e_g = max(-8, min(e_g, 24));
  const float sg = exp2f((float)e_g);
  const uint32_t tb = *tptr;
  const uint32_t trow = tb + ((uint32_t)(32 * q) << 16);
  uint32_t ph0 = 0, ph1 = 0;

  const float wpair = 1.0f / ((float)P.Bg * (float)(P.nd_x + P.nd_c + P.nd_y) * (float)n);
  const float cx = -(P.alpha_x * wpair) / var_x / sg;  // true dL/dxh = cx * g~
  const float awc = P.alpha_c * wpair, awy = P.alpha_y * wpair;
  const float s_lat = exp2f((float)E_LAT), s_h = exp2f((float)E_H), s_t = exp2f((float)E_T);
  const int RB = P.RB;

  // operands
  const tc::Op oBIG = mkop(smb, T.a_big, T.l_big, TP);
  const tc::Op oG = mkop(smb, T.a_g, T.l_g, TP);
  const tc::Op oWFX0 = mkop(smb, T.w_fx0, T.l_fx0, 128), oWFX1 = mkop(smb, T.w_fx1, T.l_fx1, ndx);
  const tc::Op oWP0 = mkop(smb, T.w_p[0], T.l_p[0], d1 ? d1 : 16), oWP1 = mkop(smb, T.w_p[1], T.l_p[1], d2 ? d2 : 16);
  const tc::Op oWP2 = mkop(smb, T.w_p[2], T.l_p[2], d3 ? d3 : 16), oWP3 = mkop(smb, T.w_p[3], T.l_p[3], ndx);
  unsigned char* pBIG = smb + T.a_big;
  unsigned char* pG = smb + T.a_g;

  // per-thread running sums over all tiles (fixed thread <-> column assignment: deterministic)
  float dbx[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) dbx[i] = 0.0f;
  float dbc[4] = {0.f, 0.f, 0.f, 0.f}, dby[4] = {0.f, 0.f, 0.f, 0.f};   // aux warps: head-bias gradient sums (side c, side y)
  float dlsx = 0.0f;
  float tot[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint32_t wacc = 0;  // 0 on the first tile (weight-gradient accumulators start from zero)
  TPHASE(TPH_SETUP);

  // everything above read only the parameters and the targets; the tile records below come from lat_fwd (PDL)
  pdl_wait();

  if (warp >= W_ISSUE) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R_ISSUE));
  }
  if (warp == W_ISSUE) {
    // ================= MMA-issue warp: mirrors the stage sequence of the epilogue warps =================================
    // all 32 lanes run the descriptor arithmetic (uniform datapath); MMAs / commits are predicated on the elected lane.
    // The tensor-memory base is a compile-time 0 here (one CTA per SM owning all 512 columns; checked below) and the
    // operand bases are recomputed from the kernel parameters, so that every MMA operand is provably warp-uniform.
    if (tb != 0u) __trap();
    constexpr uint32_t tbu = 0u;
    const uint32_t el = tc::elect_one();
    int it = 0;
    uint32_t sid = 0, wacc = 0;
    // PROF: issuer-side accounting (lane 0): cycles blocked waiting for stage signals vs cycles spent issuing MMAs
    long long iw = 0, ii = 0, itl = PROF ? clock64() : 0;
#define ISSUER_MARK(acc)                 \
  do {                                   \
    if (PROF && lane == 0) {             \
      const long long _t = clock64();    \
      acc += _t - itl;                   \
      itl = _t;                          \
    }                                    \
  } while (0)
    // first layers of a tile (S0): fx0 -> C_H, physics layer 0 -> C_X, operand = the bulk-copied tile record.  Issued at
    // the top of the first tile and, when a backward follows, EARLY for every later tile: the x path signals as soon as
    // the previous tile's last reader of C_X is through, so these MMAs queue behind that tile's last weight gradients
    // and their results are ready when the next tile starts.
    auto issue_s0 = [&](int it_) {
      const int buf_ = it_ & 1;
      const tc::Op oL = mkop(RECB + (size_t)buf_ * T.rec_buf, 0, 4096u, TP);
      tc::mbar_wait(rbar + buf_, (uint32_t)(it_ >> 1) & 1u);   // the tile record (latent operand) has landed
      __syncwarp();
      issuer_wait(sid++);   // S0
      ISSUER_MARK(iw);
      tc::issue_fwd_w(el, tbu + C_H, oL, oWFX0, 128, KZ, 0, terms);
      if constexpr (mlp) tc::issue_fwd_w(el, tbu + C_X, oL, oWP0, d1, KZ, 0, terms);
      tc::commit_w(el, bar0);
      __syncwarp();
      ISSUER_MARK(ii);
    };
    const bool early_s0 = P.with_grad != 0;
    for (long long rb = blockIdx.x; rb < P.n_rowblocks; rb += gridDim.x, ++it) {
      const int buf = it & 1;
      const unsigned char* rec = RECB + (size_t)buf * T.rec_buf;
      const tc::Op oLAT = mkop(rec, 0, 4096u, TP);
      if (it == 0 || !early_s0) issue_s0(it);
      if constexpr (mlp) {
        issuer_wait(sid++);   // S2: physics layer 1
        ISSUER_MARK(iw);
        tc::issue_fwd_ts_w(el, tbu + C_S, tbu + C_A0, oWP1, d2, d1, 0, terms);
        tc::commit_w(el, bar0);
        __syncwarp();
        ISSUER_MARK(ii);
      }
      issuer_wait(sid++);   // S5a: output layer of the data-driven decoder, as soon as its hidden activations are staged
      ISSUER_MARK(iw);
      tc::issue_fwd_w(el, tbu + C_X, oBIG, oWFX1, ndx, 128, 0, terms);
      if constexpr (!mlp) tc::commit_w(el, bar0);   // (MLP physics: the commit of S5b covers these MMAs too)
      __syncwarp();
      ISSUER_MARK(ii);
      if constexpr (mlp) {
        issuer_wait(sid++);   // S4: physics layer 2 -> first half of C_H (the fx0 accumulator has been consumed)
        ISSUER_MARK(iw);
        tc::issue_fwd_ts_w(el, tbu + C_H, tbu + C_A1, oWP2, d3, d2, 0, terms);
        tc::commit_w(el, bar0);
        __syncwarp();
        ISSUER_MARK(ii);
        issuer_wait(sid++);   // S5b: last physics layer on top of the data-driven decoder's output
        ISSUER_MARK(iw);
        tc::issue_fwd_ts_w(el, tbu + C_X, tbu + C_A2, oWP3, ndx, d3, 1, terms);
        tc::commit_w(el, bar0);
        __syncwarp();
        ISSUER_MARK(ii);
      }
      if (P.with_grad) {
        issuer_wait(sid++);   // S6: the dgrads the next epilogues wait for go first (bar0), the weight gradient of fx1 on bar1
        ISSUER_MARK(iw);
        {
          // MLP physics: the planes of W_p3 sit right behind those of W_fx1 (same row count), C_X right behind C_H: ONE
          // N = 192 dgrad per K step for both (96 cycles instead of 2 x 64: the A operand is fetched once)
          tc::issue_dgrad_w(el, tbu + C_H, oG, oWFX1, ndx, mlp ? 128 + d3 : 128, 0, terms);
          tc::commit_w(el, bar0);
          tc::issue_wgrad_w(el, tbu + C_W1, oBIG, oG, ndx, wacc, terms);
          tc::commit_w(el, bar1);
        }
        __syncwarp();
        ISSUER_MARK(ii);
        if constexpr (mlp) {
          issuer_wait(sid++);   // S7
          ISSUER_MARK(iw);
          {
            tc::issue_dgrad_ts_w(el, tbu + C_S, tbu + C_A2, oWP2, d3, d2, 0, terms);
            tc::commit_w(el, bar0);
          }
          __syncwarp();
          ISSUER_MARK(ii);
        }
        issuer_wait(sid++);   // S8: first-layer dgrad dL/d(zc|zy) = dL/dh . W_fx0 (N = 16) and the weight gradient of fx0
        ISSUER_MARK(iw);
        // the small physics dgrad of stage S9 (its epilogue is next on the critical path) goes ahead of the 48 MMAs of
        // stage S8, whose results are only needed at the end of the tile
        if constexpr (mlp) {
          issuer_wait(sid++);   // S9
          ISSUER_MARK(iw);
          tc::issue_dgrad_ts_w(el, tbu + C_X, tbu + C_A1, oWP1, d2, d1, 0, terms);
          tc::commit_w(el, bar0);
        }
        tc::issue_dgrad_w(el, tbu + C_T, oBIG, oWFX0, 128, KZ, 0, terms);
        tc::issue_wgrad_w(el, tbu + C_W0, oBIG, oLAT, KZ, wacc, terms);
        tc::commit_w(el, bar1);
        __syncwarp();
        ISSUER_MARK(ii);
        if (rb + gridDim.x < P.n_rowblocks) issue_s0(it + 1);
      }
      wacc = 1u;
    }
    if (PROF && lane == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(P.phase) + 20, (unsigned long long)iw);
      atomicAdd(reinterpret_cast<unsigned long long*>(P.phase) + 21, (unsigned long long)ii);
    }
    tc::fence_before_sync();
    __syncthreads();   // (A) every role is through with its last tile (the reduction scratch R0 aliases the BIG operand buffer)
    tc::fence_after_sync();
    tc::fence_before_sync();
    __syncthreads();   // (B) the running sums of every role are in the scratch
  } else if (warp > W_ISSUE) {
    // idle warps of the issue warpgroup: CTA-wide barriers only
    tc::fence_before_sync();
    __syncthreads();   // (A) every role is through with its last tile (the reduction scratch R0 aliases the BIG operand buffer)
    tc::fence_after_sync();
    tc::fence_before_sync();
    __syncthreads();   // (B) the running sums of every role are in the scratch
  } else if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R_AUX));
    // ================= auxiliary-decoder warps (thread = pair): decoder_c then decoder_y of tile `it`, one tile ahead of
    // the x path.  Forward, Gaussian likelihood and dgrad in registers (packed fp32 pairs); the weight-gradient operands
    // (ReLU mask plane, head-gradient x input products) go into the idle G buffer and are reduced over the pairs by
    // M = 64 MMAs issued from here.
    const int ap = tid - TNT;
    const int aw = warp - 8;
    if (tb != 0u) __trap();
    constexpr uint32_t tbu = 0u;
    const uint32_t el = tc::elect_one();
    unsigned char* pM = smb + T.a_g;                  // X8[8 chunks = 64 hidden units][TP pairs], fp16 0 / 1
    unsigned char* pB = smb + T.a_g + 16384;          // X8[3 chunks = NAUX product columns][TP pairs], hi plane; lo plane behind it
    constexpr uint32_t B_LO = 3u * TP * 16u;
    long long xw = 0, xb = 0, xtl = PROF ? clock64() : 0;   // PROF: cycles waiting (records, buffer hand-overs) vs working
#define AUX_MARK(acc)                        \
  do {                                       \
    if (PROF && ap == 0) {                   \
      const long long _t = clock64();        \
      acc += _t - xtl;                       \
      xtl = _t;                              \
    }                                        \
  } while (0)
    uint32_t wacc = 0;
    int it = 0;
    for (long long rb = blockIdx.x; rb < P.n_rowblocks; rb += gridDim.x, ++it) {
      const long long row0 = rb * RB;
      const int npairs = (int)min((long long)RB, B - row0) * n;
      const int buf = it & 1;
      const unsigned char* rec = RECB + (size_t)buf * T.rec_buf;
      AUX_MARK(xb);
      tc::mbar_wait(rbar + buf, (uint32_t)(it >> 1) & 1u);
      __syncwarp();
      AUX_MARK(xw);
      const float* RAW = reinterpret_cast<const float*>(rec + 8192);
      const bool pvalid = ap < npairs;
      float* dza = DZA + (size_t)(it & 1) * nzd * TP;
      float* sca = SCA + (size_t)(it & 1) * 2 * TP;
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        const int a_nz = s ? P.nz_y : P.nz_c, a_j0 = s ? P.nz_c : 0, a_nd = s ? P.nd_y : P.nd_c;
        const ulonglong2* aW0 = reinterpret_cast<const ulonglong2*>(AW0) + s * 64;   // [unit pair][2]: inputs (0,1), (2,3)
        const ulonglong2* aW1 = reinterpret_cast<const ulonglong2*>(AW1) + s * 64;   // [unit pair][2]: outputs (0,1), (2,3)
        const f2_t* aB0 = reinterpret_cast<const f2_t*>(AB0) + s * 32;
        float z4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) z4[j] = (j < a_nz && pvalid) ? lat_elem(rec, a_j0 + j, ap) : 0.0f;
        unsigned long long mkA = 0ull;
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
        {
          // two hidden units per step on packed pairs (FFMA2); the head sums are kept as (even-unit, odd-unit) partials
          const f2_t zz0 = pk2(z4[0], z4[0]), zz1 = pk2(z4[1], z4[1]), zz2 = pk2(z4[2], z4[2]), zz3 = pk2(z4[3], z4[3]);
          f2_t oP0 = 0ull, oP1 = 0ull, oP2 = 0ull, oP3 = 0ull;
#pragma unroll 2
          for (int c = 0; c < 8; ++c) {
            uint32_t m8 = 0u;
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
              const int kp = 4 * c + i2;   // unit pair (2 kp, 2 kp + 1)
              const ulonglong2 wa = aW0[2 * kp], wb = aW0[2 * kp + 1];
              const f2_t pre2 = ffma2(zz3, wb.y, ffma2(zz2, wb.x, ffma2(zz1, wa.y, ffma2(zz0, wa.x, aB0[kp]))));
              float pa, pb;
              upk2(pre2, pa, pb);
              const float ha = fmaxf(pa, 0.0f), hb = fmaxf(pb, 0.0f);
              m8 |= ((pa > 0.0f ? 1u : 0u) | (pb > 0.0f ? 2u : 0u)) << (2 * i2);
              const f2_t h2 = pk2(ha, hb);
              const ulonglong2 ta = aW1[2 * kp], tb2 = aW1[2 * kp + 1];
              oP0 = ffma2(h2, ta.x, oP0); oP1 = ffma2(h2, ta.y, oP1); oP2 = ffma2(h2, tb2.x, oP2); oP3 = ffma2(h2, tb2.y, oP3);
            }
            mkA |= (unsigned long long)m8 << (8 * c);
          }
          float o0, o1, o2, o3;
          {
            float a, b;
            upk2(oP0, a, b); o0 = (a + b) + AB1[s * 4 + 0];
            upk2(oP1, a, b); o1 = (a + b) + AB1[s * 4 + 1];
            upk2(oP2, a, b); o2 = (a + b) + AB1[s * 4 + 2];
            upk2(oP3, a, b); o3 = (a + b) + AB1[s * 4 + 3];
          }
          // Gaussian log-likelihood of the raw covariate / label and its gradient w.r.t. (mean, log sigma)
          float R = 0.0f;
          if (s == 0 || P.y != nullptr) {
            const float om[2] = {o0, o1}, ol[2] = {o2, o3};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (j < a_nd) {
                const float val = RAW[((s ? P.nd_c : 0) + j) * TP + ap];
                const float es = expf(ol[j]), var = es * es, d = val - om[j];
                R += -(d * d) / (2.0f * var) - ol[j] - LOG_SQRT_2PI;
                if (pvalid) {
                  g4[j] = fminf(fmaxf(-d / var, -60000.0f), 60000.0f);
                  g4[2 + j] = fminf(fmaxf(-(d * d / var - 1.0f), -60000.0f), 60000.0f);
                }
              }
            }
          }
          sca[s * TP + ap] = pvalid ? R : 0.0f;
        }
        if (P.with_grad) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            dbc[i] += s ? 0.0f : g4[i];
            dby[i] += s ? g4[i] : 0.0f;
          }
          // dgrad: dL/dz = W0^T (mask * (W1^T g))
          {
            const f2_t gg0 = pk2(g4[0], g4[0]), gg1 = pk2(g4[1], g4[1]), gg2 = pk2(g4[2], g4[2]), gg3 = pk2(g4[3], g4[3]);
            f2_t gP0 = 0ull, gP1 = 0ull, gP2 = 0ull, gP3 = 0ull;   // (even-unit, odd-unit) partials of dL/dz
#pragma unroll 2
            for (int c = 0; c < 8; ++c) {
              const uint32_t m8 = (uint32_t)(mkA >> (8 * c)) & 0xFFu;
#pragma unroll
              for (int i2 = 0; i2 < 4; ++i2) {
                const int kp = 4 * c + i2;
                const ulonglong2 ta = aW1[2 * kp], tb2 = aW1[2 * kp + 1];
                const f2_t gh2 = ffma2(gg3, tb2.y, ffma2(gg2, tb2.x, ffma2(gg1, ta.y, fmul2(gg0, ta.x))));
                float ga, gb;
                upk2(gh2, ga, gb);
                ga = ((m8 >> (2 * i2)) & 1u) ? ga : 0.0f;
                gb = ((m8 >> (2 * i2 + 1)) & 1u) ? gb : 0.0f;
                const f2_t ghm = pk2(ga, gb);
                const ulonglong2 wa = aW0[2 * kp], wb = aW0[2 * kp + 1];
                gP0 = ffma2(ghm, wa.x, gP0); gP1 = ffma2(ghm, wa.y, gP1); gP2 = ffma2(ghm, wb.x, gP2); gP3 = ffma2(ghm, wb.y, gP3);
              }
            }
            float gz[4];
            {
              float a, b;
              upk2(gP0, a, b); gz[0] = a + b;
              upk2(gP1, a, b); gz[1] = a + b;
              upk2(gP2, a, b); gz[2] = a + b;
              upk2(gP3, a, b); gz[3] = a + b;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < a_nz) dza[(a_j0 + j) * TP + ap] = gz[j];
          }
          // the operand buffer: side c waits until the x path is through with G (tile it - 1), side y until the MMAs
          // of side c have read it
          AUX_MARK(xb);
          if (s == 0) {
            if (it > 0) tc::mbar_wait(gfree, (uint32_t)(it - 1) & 1u);
          } else {
            tc::mbar_wait(abar, (uint32_t)it & 1u);
          }
          __syncwarp();
          tc::fence_after_sync();
          AUX_MARK(xw);
          // Per-pair power-of-two balance between the two operands: the mask plane carries 2^e, the products 2^-e, with
          // e chosen from the largest product of the pair so that it lands in [2^12, 2^13): exact (powers of two), keeps
          // both fp16 planes of every product in the normal range whatever the magnitude of the head gradient
          // (|g z| from 2^-2 to 2^27 without loss; beyond that the products saturate at +-60000 * 2^e)
          const float gmx = fmaxf(fmaxf(fabsf(g4[0]), fabsf(g4[1])), fmaxf(fabsf(g4[2]), fabsf(g4[3])));
          const float zmx = fmaxf(fmaxf(fabsf(z4[0]), fabsf(z4[1])), fmaxf(fmaxf(fabsf(z4[2]), fabsf(z4[3])), 1.0f));
          int e_p = (int)((__float_as_uint(gmx * zmx) >> 23) & 0xFFu) - 127 - 12;
          e_p = max(-14, min(e_p, 15));
          const uint32_t hone = (uint32_t)(e_p + 15) << 10;                      // fp16 bits of 2^e_p
          const float s_b = __uint_as_float((uint32_t)(127 - e_p) << 23);       // 2^-e_p
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t m8 = (uint32_t)(mkA >> (8 * c)) & 0xFFu;
            uint32_t wv[4];
#pragma unroll
            for (int w2 = 0; w2 < 4; ++w2) {
              const uint32_t t2 = (m8 >> (2 * w2)) & 3u;
              wv[w2] = ((t2 & 1u) | ((t2 & 2u) << 15)) * hone;   // 2^e_p in the low / high half where the unit is active
            }
            *reinterpret_cast<uint4*>(pM + ((size_t)c * TP + ap) * 16) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          }
          {
            // product columns 5 j + i = g_j * [z_0 .. z_3, 1]_i * 2^-e_p
            float v[NAUX];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float gs = g4[j] * s_b;
#pragma unroll
              for (int i = 0; i < 4; ++i) v[5 * j + i] = fminf(fmaxf(gs * z4[i], -60000.0f), 60000.0f);
              v[5 * j + 4] = fminf(fmaxf(gs, -60000.0f), 60000.0f);
            }
#pragma unroll
            for (int i = 20; i < NAUX; ++i) v[i] = 0.0f;
#pragma unroll
            for (int c = 0; c < NAUX / 8; ++c) put8(pB, B_LO, TP, c, ap, v + 8 * c);
          }
          tc::fence_async_smem();
          tc::fence_before_sync();
          asm volatile("bar.sync 10, %0;" ::"r"((uint32_t)NAUXT) : "memory");
          tc::fence_after_sync();
          if (aw == 0) {
            tc::issue_mask_wgrad64_w(el, tbu + C_AS + NAUX * s, tc::smem_u32(smb + T.a_g), tc::smem_u32(smb + T.a_g + 16384), B_LO, TP,
                                     NAUX, wacc);
            tc::commit_w(el, s == 0 ? abar : adone);
          }
          __syncwarp();
        }
      }
      if (!P.with_grad && it > 0) tc::mbar_wait(gfree, (uint32_t)(it - 1) & 1u);   // never more than one tile ahead of the consumer
      tc::mbar_arrive(adone);
      wacc = 1u;
    }
    AUX_MARK(xb);
    if (PROF && ap == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(P.phase) + 22, (unsigned long long)xw);
      atomicAdd(reinterpret_cast<unsigned long long*>(P.phase) + 23, (unsigned long long)xb);
    }
    tc::fence_before_sync();
    __syncthreads();   // (A) every role is through with its last tile (the reduction scratch R0 aliases the BIG operand buffer)
    tc::fence_after_sync();
    // ---- end of kernel: head-bias sums to the reduction scratch, masked reductions S -> dW0 / db0 / dW1 ----------------
  if (P.with_grad) {
    const int ap = tid - TNT;
    constexpr uint32_t tbu = 0u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        R0[TNT * 32 + (0 * TP + ap) * 4 + i] = dbc[i];
        R0[TNT * 32 + (1 * TP + ap) * 4 + i] = dby[i];
      }
      tc::fence_after_sync();
      {
        // M = 64 accumulator row u lives in tensor-memory lane 32 (u / 16) + u % 16: this warp's quadrant holds units
        // 16 (warp % 4) .. + 15 of each side in its lanes 0 .. 15 (the tensor-memory loads are warp-wide)
        const int qq = warp & 3, kk = 16 * qq + (lane & 15);
        const bool own = lane < 16;
        const uint32_t tl = tbu + ((uint32_t)(32 * qq) << 16);
        for (int s = 0; s < 2; ++s) {
          const Mlp2S& M = s ? P.dy : P.dc;
          const int nzk = s ? P.nz_y : P.nz_c, nd = s ? P.nd_y : P.nd_c;
          const float awt = s ? awy : awc;
          float S[NAUX];
          tc::tmem_ld8(tl + C_AS + NAUX * s, S);
          tc::tmem_ld8(tl + C_AS + NAUX * s + 8, S + 8);
          tc::tmem_ld8(tl + C_AS + NAUX * s + 16, S + 16);
          float w0e[5], w1[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) w0e[i] = i < nzk ? prm[M.g_w0 + (long long)kk * nzk + i] : 0.0f;
          w0e[4] = prm[M.g_b0 + kk];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int o = (j & 1) < nd ? ((j >> 1) * nd + (j & 1)) : -1;
            w1[j] = o >= 0 ? prm[M.g_w1 + (long long)o * 64 + kk] : 0.0f;
            if (o >= 0 && own) {
              float a = 0.0f;
#pragma unroll
              for (int i = 0; i < 5; ++i) a = fmaf(w0e[i], S[5 * j + i], a);
              part[M.g_w1 + (long long)o * 64 + kk] = a * awt;
            }
          }
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            float a = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) a = fmaf(w1[j], S[5 * j + i], a);
            if (i < nzk && own) part[M.g_w0 + (long long)kk * nzk + i] = a * awt;
            if (i == 4 && own) part[M.g_b0 + kk] = a * awt;
          }
        }
      }
    }
    tc::fence_before_sync();
    __syncthreads();   // (B) the running sums of every role are in the scratch
  } else {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R_MAIN));
  const uint32_t rec_bytes = (uint32_t)P.rec_stride;
  if (tid == 0 && (long long)blockIdx.x < P.n_rowblocks) {
    tc::mbar_expect_tx(rbar, rec_bytes);
    tc::bulk_g2s(RECB, P.rec + (long long)blockIdx.x * P.rec_stride, rec_bytes, rbar);
  }
  int it = 0;
  uint32_t sid = 0;
  for (long long rb = blockIdx.x; rb < P.n_rowblocks; rb += gridDim.x, ++it) {
    const long long row0 = rb * RB;
    const int nrows = (int)min((long long)RB, B - row0);
    const int npairs = nrows * n;
    const int buf = it & 1;
    unsigned char* rec = RECB + (size_t)buf * T.rec_buf;
    // prefetch the next tile's record into the other buffer (its previous tile is fully consumed: every MMA that read
    // it was waited for and all threads passed the end-of-tile barrier)
    if (tid == 0 && rb + gridDim.x < P.n_rowblocks) {
      tc::fence_async_smem();
      tc::mbar_expect_tx(rbar + (buf ^ 1), rec_bytes);
      tc::bulk_g2s(RECB + (size_t)(buf ^ 1) * T.rec_buf, P.rec + (rb + gridDim.x) * P.rec_stride, rec_bytes, rbar + (buf ^ 1));
    }
    tc::mbar_wait(rbar + buf, (uint32_t)(it >> 1) & 1u);
    __syncwarp();
    const tc::Op oLAT = mkop(rec, 0, 4096u, TP);   // latent operand [zd | 1 | physics input] as written by lat_fwd_kernel
    const bool pvalid = p < npairs;
    const int prow = (pvalid ? p : npairs - 1) / n;
    // the KL of the row slot this thread owns in the per-row outputs at the end of the tile: loaded now, used then
    const bool pow2 = (n & (n - 1)) == 0 && n <= 32;
    const int rslot = pow2 ? ((tid < TP && (tid & (n - 1)) == 0 && tid / n < nrows) ? tid / n : -1) : (tid < nrows ? tid : -1);
    const float kl_pre = rslot >= 0 ? __ldg(P.rowkl + row0 + rslot) : 0.0f;
    // ---- first layers of the data-driven decoder and of the physics surrogate: issued now, consumed later ----
    if (it == 0 || !P.with_grad) stage_signal<false>(sid++);   // S0 (later tiles of a training step: signalled early, end of the previous tile)
    TPHASE(TPH_LATENT);

    // ================= physics layer 0 -> tanh (bias folded into the constant-one column) =====================
    stage_wait(bar0, ph0);
    if constexpr (mlp) {
      const float inv = INV[I_P0];
#pragma unroll 2
      for (int c = 0; c < 4; ++c) {
        float v[8];
        tc::tmem_ld8(trow + C_X + 32 * hh + 8 * c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = tanh_fast(v[i] * inv) * s_t;
        // tanh outputs stay in tensor memory as packed fp16 hi / lo planes: A operand of the next layer (TS-mode MMA,
        // no shared-memory copy) and the saved activation of the backward
        tc::tmem_put8_packed(trow + C_A0 + 16 * hh + 4 * c, trow + C_A0 + 32 + 16 * hh + 4 * c, v);
      }
      stage_signal<false>(sid++);   // S2
    }
    TPHASE(TPH_A0);
    // ================= hidden layer of the data-driven decoder: ReLU (bias folded), mask in registers ============
    unsigned long long mkH = 0ull;
    {
      const float inv = INV[I_FX0];
#pragma unroll 2
      for (int c = 0; c < 8; ++c) {
        float v[8];
        tc::tmem_ld8(trow + C_H + 64 * hh + 8 * c, v);
        uint32_t m8 = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a = v[i] * inv;
          m8 |= (a > 0.0f ? 1u : 0u) << i;
          v[i] = fmaxf(a, 0.0f) * s_h;
        }
        mkH |= (unsigned long long)m8 << (8 * c);
        put8(pBIG, T.l_big, TP, 8 * hh + c, p, v);
      }
    }
    TPHASE(TPH_HD);
    stage_signal(sid++);   // S5a: the output layer of the data-driven decoder runs under the rest of the physics chain

    if constexpr (mlp) {
      // ================= physics layer 1 -> tanh ===================================================================
      stage_wait(bar0, ph0);
      {
        const float inv = INV[I_P1];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[8];
          tc::tmem_ld8(trow + C_S + 16 * hh + 8 * c, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = tanh_fast(v[i] * inv + BP1[16 * hh + 8 * c + i]) * s_t;
          tc::tmem_put8_packed(trow + C_A1 + 8 * hh + 4 * c, trow + C_A1 + 16 + 8 * hh + 4 * c, v);
        }
      }
      stage_signal<false>(sid++);   // S4
    }
    TPHASE(TPH_A1);
    if constexpr (mlp) {
      // ================= physics layer 2 -> tanh ===================================================================
      stage_wait(bar0, ph0);
      {
        const float inv = INV[I_P2];
#pragma unroll 2
        for (int c = 0; c < 4; ++c) {
          float v[8];
          tc::tmem_ld8(trow + C_H + 32 * hh + 8 * c, v);   // layer 2 accumulates in the first half of C_H (C_X holds the x head)
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = tanh_fast(v[i] * inv + BP2[32 * hh + 8 * c + i]) * s_h;
          tc::tmem_put8_packed(trow + C_A2 + 16 * hh + 4 * c, trow + C_A2 + 32 + 16 * hh + 4 * c, v);
        }
      }
      stage_signal<false>(sid++);   // S5b: last physics layer, accumulated on top of the data-driven output
    }
    TPHASE(TPH_A2);
    // the raw data row is fetched BEFORE waiting for the head MMAs (the global-load latency hides under them)
    float xv[nxh];
    {
      const long long lrow = row0 + prow;
      const long long drow = P.idx ? P.idx[lrow] : lrow;
      const float* xr = P.x + drow * ndx + nxh * hh;
#pragma unroll
      for (int c = 0; c < nxh / 4; ++c) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(xr) + c);
        xv[4 * c] = t4.x; xv[4 * c + 1] = t4.y; xv[4 * c + 2] = t4.z; xv[4 * c + 3] = t4.w;
      }
    }
    stage_wait(bar0, ph0);

    // ---- x head: xh = xh_p + xh_d, Gaussian log-likelihood of raw x, residual gradient -----------------
    {
      const float inv = INV[I_X];
      float ssq = 0.0f;
      float v[nxh];
      tld<nxh>(trow + C_X + nxh * hh, v);
      const float gsc = pvalid ? sg : 0.0f;
      const float zx0 = PHYS != 0 ? lat_elem(rec, cs0, p) : 0.0f, zx1 = PHYS == 2 ? lat_elem(rec, cs0 + 1, p) : 0.0f;
      // per-pair constants of the closed forms, hoisted out of the column loop (one reciprocal per pair instead of two
      // IEEE divisions per column)
      const float om_p = PHYS == 1 ? sqrtf(1.0f / zx0) : 0.0f;
      const float bb_p = PHYS == 1 ? 0.0f / om_p : 0.0f;
      const float bm_b = 1.0f - zx1, bm_c = 1.0f - bm_b * bm_b;
      const float bm_s = PHYS == 2 ? -1000.0f / (6.0f * (zx0 * 1e6f) * 2e-6f) : 0.0f;
#pragma unroll
      for (int i = 0; i < nxh; ++i) {
        float xh = v[i] * inv + BX[nxh * hh + i];
        if constexpr (PHYS != 0) {
          // closed-form physics (cases/damped_oscillator/mass_spring.py:8-28, cases/simple_beam/simple_beam_model.py:4-30)
          const int d = nxh * hh + i;
          if constexpr (PHYS == 1) {
            const float ph = om_p * P.grid[d];
            xh += bb_p * sinf(ph) + 1.0f * cosf(ph);
          } else {
            // w = [b x (L^2 - b^2 - x^2) + (x > a) (x - a)^3] / (6 E I), deflection = -1000 w
            const float xg = P.grid[d];
            const float t = xg - zx1;
            float w = bm_b * xg * (bm_c - xg * xg);
            w += xg > zx1 ? t * t * t : 0.0f;
            xh = fmaf(bm_s, w, xh);
          }
        }
        const float res = xv[i] - xh;
        ssq = fmaf(res, res, ssq);
        v[i] = fminf(fmaxf(gsc * res, -60000.0f), 60000.0f);
        dbx[i] += v[i];
      }
      // the aux warps are through with tile `it` (their weight-gradient MMAs have read the operands staged in G,
      // R_c / R_y / dL/dz of this tile are in shared memory)
      tc::mbar_wait(adone, (uint32_t)it & 1u);
      if (P.with_grad) {
#pragma unroll
        for (int c = 0; c < nxh / 8; ++c) put8(pG, T.l_g, TP, (nxh * hh) / 8 + c, p, v + 8 * c);
      } else {
        tc::mbar_arrive(gfree);   // forward only: nothing of G is in use; the arrive only paces the aux warps
      }
      SC[(S_Q0 + hh) * TP + p] = ssq;
    }
    if (P.with_grad) {
      // ================= backward ============================================================================
      // the dgrads the next epilogues wait for go first (bar0); the weight gradient of fx1 follows on bar1 and is only
      // waited for when its operand buffers are overwritten
      stage_signal(sid++);   // S6
      epi_sync();            // the per-pair sums of squares below are read across threads
    } else {
      epi_sync();
    }
    if (tid < TP) {
      const bool valid = tid < npairs;
      const float S = SC[S_Q0 * TP + tid] + SC[S_Q1 * TP + tid];
      SC[S_RX * TP + tid] = valid ? (-S / (2.0f * var_x) - (float)ndx * (lsx + LOG_SQRT_2PI)) : 0.0f;
      if (P.with_grad && valid) dlsx += -(P.alpha_x * wpair) * (S / var_x - (float)ndx);
    }
    TPHASE(TPH_XHEAD);

    if (P.with_grad) {
      stage_wait(bar0, ph0);   // dgrad fx1 (C_H) and dgrad of the last physics layer (C_X)
      if constexpr (mlp) {
        // d tanh of physics layer 2: the gradient replaces the saved activation in place (this thread's own packed
        // columns) and is the TS-mode A operand of the next dgrad, which then runs under the ReLU-mask epilogue
        const float inv = INV[I_XD];
#pragma unroll 2
        for (int c = 0; c < 4; ++c) {
          float g[8], a[8];
          tc::tmem_ld8(trow + C_X + 32 * hh + 8 * c, g);
          const uint32_t th = trow + C_A2 + 16 * hh + 4 * c, tl = th + 32;
          tc::tmem_get8_packed(th, tl, 1.0f / s_h, a);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] = g[i] * inv * (1.0f - a[i] * a[i]);
          tc::tmem_put8_packed(th, tl, g);
        }
        stage_signal<false>(sid++);   // S7
      }
      stage_wait(bar1, ph1);   // wgrad fx1 done: BIG (hidden activations) may be overwritten
      if constexpr (mlp) tc::mbar_arrive(gfree);   // ... and G is idle until the next x head: the aux warps may stage their operands
      {
        // ReLU mask on dL/dh -> operand of the fx0 weight gradient and of the first-layer dgrad (both MMAs of stage S8)
        const float inv = INV[I_XD];
#pragma unroll 2
        for (int c = 0; c < 8; ++c) {
          float v[8];
          tc::tmem_ld8(trow + C_H + 64 * hh + 8 * c, v);
          const uint32_t m8 = (uint32_t)(mkH >> (8 * c)) & 0xFFu;
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = ((m8 >> i) & 1u) ? v[i] * inv : 0.0f;
          put8(pBIG, T.l_big, TP, 8 * hh + c, p, v);
        }
        stage_signal(sid++);   // S8: dgrad + wgrad fx0 (waited for at the end of the tile)
      }
      TPHASE(TPH_BWD1);
      if constexpr (mlp) {
        stage_wait(bar0, ph0);   // dgrad physics layer 2
        {
          const float inv = INV[I_P2D];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float g[8], a[8];
            tc::tmem_ld8(trow + C_S + 16 * hh + 8 * c, g);
            const uint32_t th = trow + C_A1 + 8 * hh + 4 * c, tl = th + 16;
            tc::tmem_get8_packed(th, tl, 1.0f / s_t, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = g[i] * inv * (1.0f - a[i] * a[i]);
            tc::tmem_put8_packed(th, tl, g);
          }
        }
        stage_signal<false>(sid++);   // S9
        TPHASE(TPH_BWD2);
        stage_wait(bar0, ph0);   // dgrad physics layer 1
        {
          // d tanh of layer 0 and, on the CUDA cores, the dgrad to the physics latents: dL/ds0[j] = sum_k g[k] W_p0[k][j]
          const float inv = INV[I_P1D];
          float gs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
          for (int c = 0; c < 4; ++c) {
            float g[8], a[8];
            tc::tmem_ld8(trow + C_X + 32 * hh + 8 * c, g);
            tc::tmem_get8_packed(trow + C_A0 + 16 * hh + 4 * c, trow + C_A0 + 32 + 16 * hh + 4 * c, 1.0f / s_t, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float gv = g[i] * inv * (1.0f - a[i] * a[i]);
              const float4 w = WP0F4[32 * hh + 8 * c + i];
              gs[0] = fmaf(gv, w.x, gs[0]); gs[1] = fmaf(gv, w.y, gs[1]); gs[2] = fmaf(gv, w.z, gs[2]); gs[3] = fmaf(gv, w.w, gs[3]);
            }
          }
          // halves combined through shared memory (the C_T columns of tensor memory belong to the fx0 dgrad MMA in flight)
          if (hh == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) GSX[k * TP + p] = gs[k];
          }
          epi_sync();
          if (hh == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < P.nz_x)
                P.dzrec[((long long)rb * (nzd + P.nz_x) + nzd + k) * TP + p] = (gs[k] + GSX[k * TP + p]) * cx / P.phys_in_std[k];
          }
        }
        TPHASE(TPH_BWD3);
      } else {
        // closed-form physics backward: d xh_p / d zx contracted with g~ (this thread's half of the x columns)
        float s0 = 0.0f, s1 = 0.0f;
        {
          const unsigned char* gh = pG;
          // per-pair constants hoisted out of the column loop; the residual gradient g~ is read back 8 columns at a time
          const float z0 = lat_elem(rec, cs0, p);
          const float om = PHYS == 1 ? sqrtf(1.0f / z0) : 0.0f, dom = PHYS == 1 ? -om / (2.0f * z0) : 0.0f;
          const float a = PHYS == 2 ? lat_elem(rec, cs0 + 1, p) : 0.0f, b = 1.0f - a, c2 = 1.0f - b * b;
          const float iden = PHYS == 2 ? 1.0f / (6.0f * (z0 * 1e6f) * 2e-6f) : 0.0f;
          float w0 = 0.0f, w1 = 0.0f;   // beam: sum_d g w_d den, sum_d g dwa_d den
#pragma unroll
          for (int c = 0; c < nxh / 8; ++c) {
            const int ch = (nxh * hh) / 8 + c;
            const uint4 hv = *reinterpret_cast<const uint4*>(gh + ((size_t)ch * TP + p) * 16);
            const uint4 lv = *reinterpret_cast<const uint4*>(gh + T.l_g + ((size_t)ch * TP + p) * 16);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int d = 8 * ch + i;
              const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i >> 1]));
              const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[i >> 1]));
              const float g = (i & 1) ? hf.y + lf.y : hf.x + lf.x;
              if constexpr (PHYS == 1) {
                const float t = P.grid[d];
                s0 = fmaf(g, -sinf(om * t) * t * dom, s0);
              } else {
                const float xg = P.grid[d];
                const float u = c2 - xg * xg;
                float w = b * xg * u;
                float dwa = -(xg * u - 2.0f * b * b * xg);
                if (xg > a) {
                  const float t = xg - a;
                  w = fmaf(t * t, t, w);
                  dwa = fmaf(-3.0f * t, t, dwa);
                }
                w0 = fmaf(g, w, w0);
                w1 = fmaf(g, dwa, w1);
              }
            }
          }
          if constexpr (PHYS == 2) {
            s0 = 1000.0f * w0 * iden / z0;
            s1 = -1000.0f * w1 * iden;
          }
        }
        tc::mbar_arrive(gfree);   // closed-form physics: this was the last read of G
        epi_sync();
        SC[(S_Q0 + hh) * TP + p] = s0;
        float* Q2 = RED;  // second component, [2][TP]
        Q2[hh * TP + p] = s1;
        epi_sync();
        if (tid < TP) {
          float* dzx = P.dzrec + ((long long)rb * (nzd + P.nz_x) + nzd) * TP;
          dzx[tid] = (SC[S_Q0 * TP + tid] + SC[S_Q1 * TP + tid]) * cx;
          if (P.nz_x > 1) dzx[TP + tid] = (Q2[tid] + Q2[TP + tid]) * cx;
        }
      }
      // the first layers of the NEXT tile may go now (C_H and C_X have been read for the last time in this tile): they
      // queue behind this tile's last weight gradients and are complete when the next tile starts
      if (rb + gridDim.x < P.n_rowblocks) stage_signal<false>(sid++);   // S0 of tile it + 1
      stage_wait(bar1, ph1);   // dgrad + wgrad fx0 done: BIG and the record buffer are free for the next tile
      if (hh == 0) {
        // total dL/d(zc|zy): reversed + scaled gradient of the data-driven decoder (utils/transforms.py:207-219)
        // plus the auxiliary decoders' gradient
        float g1[8];
        tc::tmem_ld8(trow + C_T, g1);
        const float sc = -P.lambda_g0 * cx * INV[I_FX0D];
        float* dz = P.dzrec + (long long)rb * (nzd + P.nz_x) * TP;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < nzd) dz[k * TP + p] = fmaf(g1[k], sc, (k < P.nz_c ? awc : awy) * DZA[((it & 1) * nzd + k) * TP + p]);
      }
      TPHASE(TPH_BWD4);

    }
    epi_sync();
    TPHASE(TPH_LATENT_BWD);
    // ---- per-row outputs: MC means of the three reconstruction terms + the KL of lat_fwd_kernel ---------------------
    {
      float rx = 0.0f, rc = 0.0f, ry = 0.0f;
      int r = -1;
      if (pow2) {
        // n consecutive pairs of a row sit in n consecutive lanes: segmented butterfly over the warp (fixed order)
        if (tid < TP) {
          rx = SC[S_RX * TP + tid]; rc = SCA[((it & 1) * 2 + 0) * TP + tid]; ry = SCA[((it & 1) * 2 + 1) * TP + tid];
          for (int off = n >> 1; off > 0; off >>= 1) {
            rx += __shfl_xor_sync(0xffffffffu, rx, off);
            rc += __shfl_xor_sync(0xffffffffu, rc, off);
            ry += __shfl_xor_sync(0xffffffffu, ry, off);
          }
          if ((tid & (n - 1)) == 0 && tid / n < nrows) r = tid / n;
        }
      } else if (tid < nrows) {
        r = tid;
        for (int m = 0; m < n; ++m) {
          rx += SC[S_RX * TP + r * n + m];
          rc += SCA[((it & 1) * 2 + 0) * TP + r * n + m];
          ry += SCA[((it & 1) * 2 + 1) * TP + r * n + m];
        }
      }
      if (r >= 0) {
        const float inv_n = 1.0f / (float)n;
        rx *= inv_n; rc *= inv_n; ry *= inv_n;
        const float kl = kl_pre;   // r == rslot
        const float loss = P.beta_x * kl - P.alpha_x * rx - P.alpha_c * rc - P.alpha_y * ry;
        if (P.out.row_loss) {
          float* o = P.out.row_loss + row0 + r;
          o[0] = loss; o[B] = kl; o[2 * B] = rx; o[3 * B] = rc; o[4 * B] = ry; o[5 * B] = 0.0f;
        }
        // running sums of the row slot this thread owns (fixed thread <-> slot map: deterministic), added up over the
        // threads in a fixed order at the end of the kernel
        tot[0] += loss; tot[1] += kl; tot[2] += rx; tot[3] += rc; tot[4] += ry;
      }
    }
    wacc = 1u;
    TPHASE(TPH_ROWOUT);
  }  // tiles

    tc::fence_before_sync();
    __syncthreads();   // (A) every role is through with its last tile (the reduction scratch R0 aliases the BIG operand buffer)
    tc::fence_after_sync();
  // ---- end of kernel: loss sums, weight gradients out of TMEM, bias / log_sigma_x sums ---------------------
#pragma unroll
  for (int k = 0; k < 5; ++k) R0[TNT * 39 + k * TNT + tid] = tot[k];
  if (P.with_grad && !P.latent_only) {
    const int k = p;  // TMEM lane = hidden unit of the 128-wide layers
    // fx1: dW[n][k] (nd_x x 128)
    {
      const float sc = exp2f(-(float)E_H) * cx;
      float v[nxh];
      tld<nxh>(trow + C_W1 + nxh * hh, v);
#pragma unroll
      for (int i = 0; i < nxh; ++i) part[P.fx.g_w1 + (long long)(nxh * hh + i) * 128 + k] = v[i] * sc;
    }
    {
      if (hh == 0) {
        float v0[16];
        tc::tmem_ld16(trow + C_W0, v0);
        // fx0: dW[k][j] (128 x nzd), bias from the constant-one column; the decoder's own weights see the
        // un-reversed gradient (utils/transforms.py:207-219 reverses only d/dz)
        const float sc = exp2f(-(float)E_LAT) * cx;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < nzd) part[P.fx.g_w0 + (long long)k * nzd + j] = v0[j] * sc;
          if (j == c1) part[P.fx.g_b0 + k] = v0[j] * sc;
        }
      }
    }
    // per-thread running sums -> fixed-order sums over the 128 pair slots of each column, two levels deep
    // (4 pair groups of 32 per column, then the 4 partials) so that no thread walks a 128-long dependent chain
#pragma unroll
    for (int i = 0; i < 32; ++i) R0[tid * 32 + i] = dbx[i];
    R0[TNT * 36 + tid] = dlsx;
  }
    tc::fence_before_sync();
    __syncthreads();   // (B) the running sums of every role are in the scratch
  if (tid < 5) {
    // loss scalars of this CTA: loss, KL, R_x, R_c, R_y summed over the row slots in thread order
    const float* src = R0 + TNT * 39 + tid * TNT;
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < TNT; j += 4) { a4[0] += src[j]; a4[1] += src[j + 1]; a4[2] += src[j + 2]; a4[3] += src[j + 3]; }
    part[P.n_params + tid] = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  }
  if (P.with_grad && !P.latent_only) {
    float* R1 = R0 + TNT * 37;   // partials [4 groups][ndx + 8 + 1]
    {
      const int col = tid & 63, grp = tid >> 6;   // 64 columns x 4 pair groups
      if (col < ndx) {
        const int h2 = col / nxh, i = col - h2 * nxh;
        float s = 0.0f;
#pragma unroll 8
        for (int j = 32 * grp; j < 32 * grp + 32; ++j) s += R0[(h2 * TP + j) * 32 + i];
        R1[grp * 80 + col] = s;
      }
      if (col < 8) {   // aux head biases: side = col >> 2, output jj = col & 3
        const int side2 = col >> 2, jj = col & 3;
        float s = 0.0f;
#pragma unroll 8
        for (int j = 32 * grp; j < 32 * grp + 32; ++j) s += R0[TNT * 32 + (side2 * TP + j) * 4 + jj];
        R1[grp * 80 + 64 + col] = s;
      }
      if (col == 8) {
        float s = 0.0f;
#pragma unroll 8
        for (int j = 32 * grp; j < 32 * grp + 32; ++j) s += R0[TNT * 36 + j];
        R1[grp * 80 + 72] = s;
      }
    }
    epi_sync();
    if (tid < ndx) part[P.fx.g_b1 + tid] = ((R1[tid] + R1[80 + tid]) + (R1[160 + tid] + R1[240 + tid])) * cx;
    if (tid >= 64 && tid < 72) {
      const int col = tid - 64, side2 = col >> 2, jj = col & 3;
      const int nd = side2 ? P.nd_y : P.nd_c;
      if ((jj & 1) < nd) {
        const float s = (R1[64 + col] + R1[80 + 64 + col]) + (R1[160 + 64 + col] + R1[240 + 64 + col]);
        part[(side2 ? P.dy.g_b1 : P.dc.g_b1) + (jj >> 1) * nd + (jj & 1)] = s * (side2 ? awy : awc);
      }
    }
    if (tid == 96) part[P.g_lsx] = (R1[72] + R1[80 + 72]) + (R1[160 + 72] + R1[240 + 72]);
  }
  }  // roles
  if (warp < 8) TPHASE(TPH_FLUSH);
  if (PROF && tid == 0)
#pragma unroll
    for (int k = 0; k < (PROF ? TPH_COUNT : 1); ++k) atomicAdd(reinterpret_cast<unsigned long long*>(P.phase) + k, (unsigned long long)phs[k]);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, C_ALLOC);
}

template <bool PROF, int PHYS, int NDX>
static void launch_one(const TcParams& p, int grid, cudaStream_t s) {
  launch_pdl(dec_tc_kernel<PROF, PHYS, NDX>, grid, NALL, (size_t)p.total, s, p);
}

// supported (physics kind, nd_x) pairs: (MLP, 64) bridge, (mass_spring, 64) damped_oscillator, (beam, 32) simple_beam,
// plus the other nd_x of each closed form
void launch_dec_tc(const TcParams& p, int grid, cudaStream_t s) {
  const int ph = p.d.phys_kind, nx = p.d.nd_x;
  if (ph == 0 && nx == 64) {
    if (p.d.phase != nullptr) launch_one<true, 0, 64>(p, grid, s);
    else launch_one<false, 0, 64>(p, grid, s);
  } else if (ph == 1 && nx == 64) launch_one<false, 1, 64>(p, grid, s);
  else if (ph == 2 && nx == 32) {
    if (p.d.phase != nullptr) launch_one<true, 2, 32>(p, grid, s);
    else launch_one<false, 2, 32>(p, grid, s);
  }
}

bool dec_tc_has_variant(int phys_kind, int nd_x) {
  return (phys_kind == 0 && nd_x == 64) || (phys_kind == 1 && nd_x == 64) || (phys_kind == 2 && nd_x == 32);
}

int configure_dec_tc_kernel() {
  int e = (int)cudaFuncSetAttribute(dec_tc_kernel<true, 0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (!e) e = (int)cudaFuncSetAttribute(dec_tc_kernel<false, 0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (!e) e = (int)cudaFuncSetAttribute(dec_tc_kernel<false, 1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (!e) e = (int)cudaFuncSetAttribute(dec_tc_kernel<false, 2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (!e) e = (int)cudaFuncSetAttribute(dec_tc_kernel<true, 2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  return e;
}

}  // namespace dpv
