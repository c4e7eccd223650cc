// Cycles per tcgen05.mma.kind::f16 (M=128, K=16) by operand form, issued back to back by one thread:
//   form 0: A K-major (smem), B K-major (smem)            -- forward
//   form 1: A K-major (smem), B MN-major (smem)           -- dgrad
//   form 2: A MN-major (smem), B MN-major (smem)          -- wgrad
//   form 3: A in tensor memory, B K-major (smem)          -- TS forward
// Operands in the no-swizzle X8 layout of tc.cuh with R = 128 rows.
#include <cstdio>
#include "tc.cuh"
using namespace dpv;

__global__ void __launch_bounds__(128, 1) rate(long long* out, int N, int nmma, int form) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + 96 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < 96 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smraw)[e] = 0x3c003c00u;
  if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (tid < 32) tc::tmem_alloc(tptr, 512);
  tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tb = *tptr;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t abase = tc::smem_u32(smraw), bbase = abase + 48 * 1024;
    const int R = 128;
    uint64_t ad[8], bd[8];
    uint32_t idesc = 0;
    for (int ks = 0; ks < 8; ++ks) {
      if (form == 0 || form == 3) { ad[ks] = tc::desc_kmajor(abase, R, 2 * ks); bd[ks] = tc::desc_kmajor(bbase, N, 2 * ks); idesc = tc::make_idesc(128, N, 0, 0); }
      else if (form == 1) { ad[ks] = tc::desc_kmajor(abase, R, 2 * ks); bd[ks] = tc::desc_mnmajor(bbase, R, 0, 16 * ks); idesc = tc::make_idesc(128, N, 0, 1); }
      else { ad[ks] = tc::desc_mnmajor(abase, R, 0, 16 * ks); bd[ks] = tc::desc_mnmajor(bbase, R, 0, 16 * ks); idesc = tc::make_idesc(128, N, 1, 1); }
    }
    t0 = clock64();
    for (int i = 0; i < nmma; i += 8) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (form == 3) tc::mma_f16_ts(tb, tb + 256 + 8 * ks, bd[ks], idesc, (i + ks) > 0);
        else tc::mma_f16(tb, ad[ks], bd[ks], idesc, (i + ks) > 0);
      }
    }
    tc::commit(bar);
    t1 = clock64();
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  if (tid == 0) { out[0] = t1 - t0; out[1] = clock64() - t0; }
  tc::fence_before_sync(); __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int Ns[] = {16, 32, 64, 128};
  const char* names[] = {"SS K/K (fwd)", "SS K/MN (dgrad)", "SS MN/MN (wgrad)", "TS (A in TMEM)"};
  for (int form = 0; form < 4; ++form)
    for (int N : Ns) {
      for (int rep = 0; rep < 2; ++rep) rate<<<1, 128, 100 * 1024>>>(d, N, 512, form);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-18s N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA [%s]\n", names[form], N, h[0] / 512.0, h[1] / 512.0, cudaGetErrorString(e));
    }
  return 0;
}
