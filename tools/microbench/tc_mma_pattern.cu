// The MMA issue sequence of enc_fused_kernel's issue warp, alone on the SM (no waits on other warps): per tile
// 3 x [head_u: 12 TS MMAs (N = 64 / 32 / 32) + commit ; L1_u: 15 SS MMAs (N = 64, K = 80) + commit].  Variants:
//   0: the kernel's lambdas (run-time loops)   1: fully unrolled, immediates   2: as 1 but ONE commit per tile
//   3: as 0 + tcgen05.fence::after_thread_sync before every group   4: as 3 + an mbarrier wait on an already completed phase
//   5: as 4 + 20 more warps waiting on a barrier that never completes
#include <cstdio>
#include "tc.cuh"
using namespace dpv;

template <int VAR>
__global__ void __launch_bounds__(768, 1) pat(long long* out, int ntiles, int terms, int ksx, int Hc, int Oc, int rnd) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + 200 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 16);
  const int tid = threadIdx.x;
  for (int e = tid; e < 200 * 1024 / 4; e += 768) { uint32_t h = (uint32_t)e * 2654435761u; h ^= h >> 15; reinterpret_cast<uint32_t*>(smraw)[e] = rnd ? ((h & 0x83ff83ffu) | 0x30003000u) : 0x3c003c00u; }
  if (tid == 0) { for (int i = 0; i < 16; ++i) tc::mbar_init(bars + i, 1); tc::mbar_fence_init(); tc::mbar_arrive(bars + 10); }
  if (tid < 32) tc::tmem_alloc(tptr, 512);
  tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  if (*tptr != 0u) __trap();
  if (tid < 32) {
    const uint32_t el = tc::elect_one();
    constexpr uint32_t HI = 8u | tc::DESC_VERSION_HI;
    const uint32_t xb0 = tc::smem_u32(smraw), w0b = xb0 + 82 * 1024, w1b = w0b + 62 * 1024;
    const uint32_t l_x = 10 * 128 * 16, l_0 = 10 * 192 * 16, l_1 = 24 * 64 * 16;
    const uint32_t C_A = 192, C_O = 384;
    const uint32_t idesc_l1 = tc::make_idesc(128, 64, 0, 0);
    auto issue_l1 = [&](int u, int buf) {
      const uint32_t xbase = xb0 + (uint32_t)buf * 2u * l_x;
      uint32_t acc = 0;
      for (int t = 0; t < terms; ++t) {
        uint32_t alo = (((xbase + (t == 1 ? l_x : 0u)) >> 4) & 0x3FFFu) | (128u << 16);
        uint32_t wlo = (((w0b + (t == 2 ? l_0 : 0u) + (uint32_t)(64 * u) * 16u) >> 4) & 0x3FFFu) | ((uint32_t)Hc << 16);
#pragma unroll 5
        for (int k = 0; k < ksx; ++k) {
          tc::mma_f16_w(el, (uint32_t)(64 * u), tc::pack64(alo, HI), tc::pack64(wlo, HI), idesc_l1, acc);
          acc = 1; alo += 256u; wlo += 2u * (uint32_t)Hc;
        }
      }
    };
    auto issue_head = [&](int u, int obuf) {
      const int n0 = u == 0 ? 0 : (u == 1 ? 8 : 32), N = u == 0 ? Oc : 32;
      const uint32_t idesc = tc::make_idesc(128, N, 0, 0);
      const uint32_t d = C_O + (uint32_t)(obuf * Oc + n0);
      uint32_t acc = u == 0 ? 0u : 1u;
      for (int t = 0; t < terms; ++t) {
        const uint32_t a = C_A + (uint32_t)(32 * u) + (t == 1 ? (uint32_t)(Hc >> 1) : 0u);
        uint32_t wlo = (((w1b + (t == 2 ? l_1 : 0u) + ((uint32_t)(8 * u) * (uint32_t)Oc + (uint32_t)n0) * 16u) >> 4) & 0x3FFFu) | ((uint32_t)Oc << 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc::mma_f16_ts_w(el, d, a + 8u * k, tc::pack64(wlo, HI), idesc, acc);
          acc = 1; wlo += 2u * (uint32_t)Oc;
        }
      }
    };
    const long long t0 = clock64();
    for (int it = 0; it < ntiles; ++it) {
      const int par = it & 1;
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        if (VAR == 0 || VAR >= 3) {
          if (VAR >= 4) tc::mbar_wait(bars + 10, 0);
          if (VAR >= 3) tc::fence_after_sync();
          issue_head(u, par);
          tc::commit_w(el, bars + u);
          if (VAR >= 4) tc::mbar_wait(bars + 10, 0);
          if (VAR >= 3) tc::fence_after_sync();
          issue_l1(u, par ^ 1);
          tc::commit_w(el, bars + 3 + u);
        } else {
          const int n0 = u == 0 ? 0 : (u == 1 ? 8 : 32);
          const uint32_t d = C_O + (uint32_t)(par * 64 + n0);
          if (u == 0) {
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc::mma_f16_ts_w(el, d, C_A + (t == 1 ? 96u : 0u) + 8u * k, tc::pack64((((w1b + (t == 2 ? l_1 : 0u)) >> 4) & 0x3FFFu | (64u << 16)) + 128u * k, HI),
                                 tc::make_idesc(128, 64, 0, 0), (t | k) ? 1u : 0u);
          } else {
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc::mma_f16_ts_w(el, d, C_A + 32u * u + (t == 1 ? 96u : 0u) + 8u * k,
                                 tc::pack64((((w1b + (t == 2 ? l_1 : 0u) + ((uint32_t)(8 * u) * 64u + (uint32_t)n0) * 16u) >> 4) & 0x3FFFu | (64u << 16)) + 128u * k, HI),
                                 tc::make_idesc(128, 32, 0, 0), 1u);
          }
          if (VAR == 1) tc::commit_w(el, bars + u);
          const uint32_t xbase = xb0 + (uint32_t)(par ^ 1) * 2u * l_x;
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int k = 0; k < 5; ++k)
              tc::mma_f16_w(el, (uint32_t)(64 * u), tc::pack64((((xbase + (t == 1 ? l_x : 0u)) >> 4) & 0x3FFFu | (128u << 16)) + 256u * k, HI),
                            tc::pack64((((w0b + (t == 2 ? l_0 : 0u) + (uint32_t)(64 * u) * 16u) >> 4) & 0x3FFFu | (192u << 16)) + 384u * k, HI), idesc_l1, (t | k) ? 1u : 0u);
          if (VAR == 1) tc::commit_w(el, bars + 3 + u);
        }
      }
      if (VAR == 2) tc::commit_w(el, bars + 6);
    }
    tc::commit_w(el, bars + 8);
    const long long t1 = clock64();
    tc::mbar_wait(bars + 8, 0);
    if (tid == 0) { out[0] = t1 - t0; out[1] = clock64() - t0; tc::mbar_arrive(bars + 11); }
  } else if (VAR == 5 && tid >= 128) {
    tc::mbar_wait(bars + 11, 0);
  }
  tc::fence_before_sync(); __syncthreads();
  if (tid < 32) tc::tmem_dealloc(0u, 512);
}

template <int V>
static void run(long long* d, const char* name) {
  cudaFuncSetAttribute(pat<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  for (int rep = 0; rep < 2; ++rep) pat<V><<<1, 768, 201 * 1024>>>(d, 64, 3, 5, 192, 64, V == 6 ? 1 : 0);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-44s: issue %.0f cyc/tile, complete %.0f cyc/tile (81 MMAs; pipe floor 45 x 48 + 12 x 32 + 24 x 16 = 2928) [%s]\n", name, h[0] / 64.0, h[1] / 64.0,
         cudaGetErrorString(e));
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<0>(d, "kernel lambdas (run-time loops)");
  run<1>(d, "unrolled, immediates, 6 commits per tile");
  run<2>(d, "unrolled, immediates, 1 commit per tile");
  run<3>(d, "lambdas + fence::after_thread_sync per group");
  run<4>(d, "lambdas + fence + satisfied mbarrier wait");
  run<5>(d, "  ... + 20 warps waiting on an mbarrier");
  run<6>(d, "as 4 with pseudo-random operand data");
  return 0;
}
