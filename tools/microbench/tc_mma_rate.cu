// Throughput of back-to-back tcgen05.mma.kind::f16 (M=128, K=16, SS operands in the no-swizzle X8 layout) issued by
// one thread: cycles per MMA vs N, and with A/B descriptors walking a 128-column operand (as the decoder kernel does).
#include <cstdio>
#include <vector>
#include "tc.cuh"
using namespace dpv;

__global__ void __launch_bounds__(128, 1) rate(long long* out, int N, int nmma, int walk) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + 96 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < 96 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smraw)[e] = 0x3c003c00u;  // fp16 1.0
  if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (tid < 32) tc::tmem_alloc(tptr, 256);
  tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tb = *tptr;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(128, N, 0, 0);
    const uint32_t abase = tc::smem_u32(smraw), bbase = abase + 64 * 1024;
    uint32_t alo = ((abase >> 4) & 0x3FFF) | (128u << 16), blo = ((bbase >> 4) & 0x3FFF) | ((uint32_t)N << 16);
    const uint32_t hi = 8u | (1u << 14);
    t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const uint32_t ao = walk ? (uint32_t)((i & 7) * 2 * 128) : 0u, bo = walk ? (uint32_t)((i & 7) * 2 * N) : 0u;
      tc::mma_f16(tb, tc::pack64(alo + ao, hi), tc::pack64(blo + bo, hi), idesc, i > 0);
    }
    tc::commit(bar);
    t1 = clock64();
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  if (tid == 0) { out[0] = t1 - t0; out[1] = clock64() - t0; }
  tc::fence_before_sync(); __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tb, 256);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int Ns[] = {16, 32, 64, 128, 192, 256};
  for (int walk = 0; walk < 2; ++walk)
    for (int N : Ns) {
      for (int rep = 0; rep < 2; ++rep) rate<<<1, 128, 100 * 1024>>>(d, N, 512, walk);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("N=%3d walk=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (math floor %.0f, smem-feed floor %.0f) [%s]\n", N, walk,
             h[0] / 512.0, h[1] / 512.0, N / 2.0, (128 + N) * 32 / 128.0, cudaGetErrorString(e));
    }
  return 0;
}
