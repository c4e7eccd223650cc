// Cycles per tcgen05.mma.kind::f16 (K = 16) issued WARP-WIDE on the uniform datapath (tc.cuh *_w helpers; the earlier
// probes issued from one thread in a divergent branch and measured the 93-cycle waterfall loop, not the pipe):
//   form 0: SS  A K-major (smem), B K-major (smem)     M = 128
//   form 1: TS  A in tensor memory, B K-major (smem)   M = 128
//   form 2: SS  M = 64
//   form 3: TS  M = 64
//   form 4: SS  A MN-major, B MN-major (wgrad form)    M = 128
// Operands in the no-swizzle X8 layout with R = 128 (A) / N (B) rows.  Tensor-memory base is the compile-time 0.
#include <cstdio>
#include "tc.cuh"
using namespace dpv;

template <int form>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int N, int nmma) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + 96 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < 96 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smraw)[e] = 0x3c003c00u;
  if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (tid < 32) tc::tmem_alloc(tptr, 512);
  tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  if (*tptr != 0u) __trap();
  long long t0 = 0, t1 = 0;
  if (tid < 32) {
    const uint32_t el = tc::elect_one();
    const uint32_t abase = tc::smem_u32(smraw), bbase = abase + 32 * 1024;
    const int M = (form == 2 || form == 3) ? 64 : 128;
    const uint32_t idesc = form == 4 ? tc::make_idesc(M, N, 1, 1) : tc::make_idesc(M, N, 0, 0);
    const uint32_t hi = 8u | tc::DESC_VERSION_HI;
    t0 = clock64();
    for (int i = 0; i < nmma; i += 8) {
      uint32_t alo = ((abase >> 4) & 0x3FFFu) | (128u << 16), blo = ((bbase >> 4) & 0x3FFFu) | ((uint32_t)N << 16);
      uint32_t at = 256u;
      if (form == 4) { alo = ((abase >> 4) & 0x3FFFu) | (8u << 16); blo = ((bbase >> 4) & 0x3FFFu) | (8u << 16); }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (form == 1 || form == 3) tc::mma_f16_ts_w(el, 0u, at, tc::pack64(blo, hi), idesc, (i + ks) > 0);
        else if (form == 4) tc::mma_f16_w(el, 0u, tc::pack64(alo, 128u | tc::DESC_VERSION_HI), tc::pack64(blo, 128u | tc::DESC_VERSION_HI), idesc, (i + ks) > 0);
        else tc::mma_f16_w(el, 0u, tc::pack64(alo, hi), tc::pack64(blo, hi), idesc, (i + ks) > 0);
        if (form == 4) { alo += 16u; blo += 16u; } else { alo += 256u; blo += 2u * (uint32_t)N; }
        at += 8u;
      }
    }
    tc::commit_w(el, bar);
    t1 = clock64();
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  if (tid == 0) { out[0] = t1 - t0; out[1] = clock64() - t0; }
  tc::fence_before_sync(); __syncthreads();
  if (tid < 32) tc::tmem_dealloc(0u, 512);
}

template <int F>
static void launch(long long* d, int N) {
  cudaFuncSetAttribute(rate<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  rate<F><<<1, 128, 100 * 1024>>>(d, N, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
  const char* names[] = {"SS  M=128", "TS  M=128", "SS  M=64", "TS  M=64", "SS MN/MN M=128"};
  for (int form = 0; form < 5; ++form)
    for (int N : Ns) {
      if (form == 4 && N > 128) continue;
      for (int rep = 0; rep < 2; ++rep) {
        switch (form) {
          case 0: launch<0>(d, N); break;
          case 1: launch<1>(d, N); break;
          case 2: launch<2>(d, N); break;
          case 3: launch<3>(d, N); break;
          default: launch<4>(d, N); break;
        }
      }
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-16s N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (math floor %d) [%s]\n", names[form], N, h[0] / 512.0, h[1] / 512.0,
             ((form == 2 || form == 3) ? 64 : 128) * N / 256, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
    }
  return 0;
}
