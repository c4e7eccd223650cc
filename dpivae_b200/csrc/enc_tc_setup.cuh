// Set-up shared by the tensor-core encoder forward kernels (enc_tc_fwd_kernel, enc_fused_kernel): operand-split helpers and
// the one-off staging of every encoder unit's weights as fp16 hi/lo operand planes in shared memory.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc.cuh"

namespace dpv {
namespace {

constexpr int ENT = 256;    // backward kernel: 8 warps
constexpr int E_X = 4;     // operand scale exponents: standardised inputs
constexpr int E_HID = 6;   // hidden activations

__device__ __forceinline__ void put8e(unsigned char* plane, uint32_t lo_off, int R, int chunk, int row, const float* v) {
  uint4 hi, lo;
  tc::split8(v, hi, lo);
  unsigned char* dst = plane + ((size_t)chunk * R + row) * 16;
  *reinterpret_cast<uint4*>(dst) = hi;
  *reinterpret_cast<uint4*>(dst + lo_off) = lo;
}

// block-wide max of two values at once (uses red[2 NW], NW = warps in the block)
template <int NW = ENT / 32>
__device__ __forceinline__ void block_max2(float& a, float& b, float* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, off));
    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, off));
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = a; red[NW + (threadIdx.x >> 5)] = b; }
  __syncthreads();
  a = red[0]; b = red[NW];
#pragma unroll
  for (int w = 1; w < NW; ++w) { a = fmaxf(a, red[w]); b = fmaxf(b, red[NW + w]); }
}
__device__ __forceinline__ int scale_exp_e(float mx) {
  if (!(mx > 0.0f) || !isfinite(mx)) return 0;
  int e;
  frexpf(mx, &e);
  return 9 - e;
}

// Weights of all encoder units -> operand planes (models/encoders.py:7-31: first layers [Hc x KX] with the bias in column
// K0, block-diagonal heads [Oc x Hc]); head biases and output rows -> B1 / OROW.  NT = threads of the calling block (all
// of them call).  Returns the power-of-two operand scale exponents of the two weight operands.
template <int NT>
__device__ __noinline__ void enc_tc_stage_fwd(const EncTcParams& P, unsigned char* smb, int& k_w0_out, int& k_w1_out) {
  constexpr int ENTF = NT;
  float* smf = reinterpret_cast<float*>(smb);
  const int tid = threadIdx.x;
  const int K0 = P.K0, Hc = P.Hc, Oc = P.Oc;
  float* B1 = smf + (P.f_b1 >> 2);
  float* RED = smf + (P.f_red >> 2);
  int* OROW = reinterpret_cast<int*>(smb + P.f_orow);
  const float* prm = P.params;
  // Raw weights first: every unit's parameters are one contiguous block [w0 | b0 | w1 | b1] of the flat buffer; it is
  // copied ONCE with coalesced loads into shared memory (when the plan has room), and the scale search / operand
  // staging below read it from there -- the per-element global loads of the old set-up cost ~90 us per launch.
  const float* ub[3] = {prm + P.g_w0[0], prm + P.g_w0[P.n_units > 1 ? 1 : 0], prm + P.g_w0[P.n_units > 2 ? 2 : 0]};
  if (P.f_raw >= 0) {
    float* RAW = smf + (P.f_raw >> 2);
    int ro = 0;
    for (int u = 0; u < P.n_units; ++u) {
      const int nblk = P.H[u] * K0 + P.H[u] + P.O[u] * P.H[u] + P.O[u];
      const float* src = prm + P.g_w0[u];
      copy_g2s_batched<ENTF>(RAW + ro, src, nblk);
      ub[u] = RAW + ro;
      ro += nblk;
    }
    __syncthreads();
  }
  // hidden unit n of the concatenated first layers: unit u(n), local index; column K0 of the operand carries the bias
  auto unit_of_h = [&](int n, int& local) -> int {
    int u = 0;
    while (u + 1 < P.n_units && n >= P.h_off[u + 1]) ++u;
    local = n - P.h_off[u];
    return u;
  };
  // head row o of the concatenated (block-diagonal) heads
  auto unit_of_o = [&](int o, int& local) -> int {
    for (int u = 0; u < P.n_units; ++u)
      if (o >= P.o_off[u] && o < P.o_off[u] + P.O[u]) {
        local = o - P.o_off[u];
        return u;
      }
    local = 0;
    return -1;
  };
  // power-of-two operand scales from the block maxima: plain linear scans of each unit's [w0 | b0] and [w1] ranges
  float m0 = 0.0f, m1 = 0.0f;
  for (int u = 0; u < P.n_units; ++u) {
    const int n0 = P.H[u] * K0 + P.H[u], n1 = P.O[u] * P.H[u];
    for (int e = tid; e < n0; e += ENTF) m0 = fmaxf(m0, fabsf(ub[u][e]));
    for (int e = tid; e < n1; e += ENTF) m1 = fmaxf(m1, fabsf(ub[u][n0 + e]));
  }
  block_max2<ENTF / 32>(m0, m1, RED);
  const int k_w0 = scale_exp_e(m0), k_w1 = scale_exp_e(m1);
  // operand staging, unit by unit (no per-element unit search / division): first layers = rows h_off[u] + l of the
  // [Hc x KX] operand, 8 consecutive inputs per item, bias in column K0; heads = block-diagonal [Oc x Hc]
  {
    const float s0 = exp2f((float)k_w0), s1 = exp2f((float)k_w1);
    const int nch0 = (K0 >> 3) + 1;   // data chunks + the bias chunk (the remaining padding chunks stay zero)
    for (int u = 0; u < P.n_units; ++u) {
      const float* w0 = ub[u];
      const float* b0 = ub[u] + P.H[u] * K0;
      const float* w1 = b0 + P.H[u];
      for (int e = tid; e < P.H[u] * nch0; e += ENTF) {
        const int l = e / nch0, ch = e - l * nch0;
        float v[8];
        if (ch < (K0 >> 3)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = w0[l * K0 + 8 * ch + i] * s0;
        } else {
          v[0] = b0[l] * s0;
#pragma unroll
          for (int i = 1; i < 8; ++i) v[i] = 0.0f;
        }
        put8e(smb + P.w_0, P.l_0, Hc, ch, P.h_off[u] + l, v);
      }
      const int nch1 = P.H[u] >> 3;
      for (int e = tid; e < P.O[u] * nch1; e += ENTF) {
        const int lo = e / nch1, ch = e - lo * nch1;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = w1[lo * P.H[u] + 8 * ch + i] * s1;
        put8e(smb + P.w_1, P.l_1, Oc, (P.h_off[u] >> 3) + ch, P.o_off[u] + lo, v);
      }
    }
  }
  for (int o = tid; o < Oc; o += ENTF) {
    int l;
    const int u = unit_of_o(o, l);
    B1[o] = u >= 0 ? ub[u][P.H[u] * K0 + P.H[u] + P.O[u] * P.H[u] + l] : 0.0f;
    OROW[o] = u >= 0 ? P.out_row[u] + l : -1;
  }
  k_w0_out = k_w0;
  k_w1_out = k_w1;
}

}  // namespace
}  // namespace dpv
