// Does other activity on the SM slow a back-to-back tcgen05.mma series?  One issue warp (warp-wide issue, unrolled, as in
// tc_mma_rate3.cu) next to (a) nothing, (b) 8 warps spinning in mbarrier.try_wait on a barrier that never completes,
// (c) 12 warps streaming 16-byte shared-memory stores + loads, (d) 8 warps doing tcgen05.ld / st on other TMEM columns,
// (e) all of them.  SS form M=128 N=64 and TS form N=64 / N=32.
#include <cstdio>
#include "tc.cuh"
using namespace dpv;

template <int FORM, int N>
__global__ void __launch_bounds__(1024, 1) rate(long long* out, int nmma, int mode) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + 160 * 1024);
  uint64_t* never = bar + 1;
  volatile int* stop = reinterpret_cast<volatile int*>(bar + 2);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 3);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 96 * 1024 / 4; e += 1024) reinterpret_cast<uint32_t*>(smraw)[e] = 0x3c003c00u;
  if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_init(never, 1); *stop = 0; tc::mbar_fence_init(); }
  if (tid < 32) tc::tmem_alloc(tptr, 512);
  tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  if (*tptr != 0u) __trap();
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t el = tc::elect_one();
    const uint32_t abase = tc::smem_u32(smraw), bbase = abase + 32 * 1024;
    const uint32_t idesc = tc::make_idesc(128, N, 0, 0);
    const uint32_t hi = 8u | tc::DESC_VERSION_HI;
    t0 = clock64();
    for (int i = 0; i < nmma; i += 8) {
      uint32_t alo = ((abase >> 4) & 0x3FFFu) | (128u << 16), blo = ((bbase >> 4) & 0x3FFFu) | ((uint32_t)N << 16);
      uint32_t at = 256u;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (FORM == 1) tc::mma_f16_ts_w(el, 0u, at, tc::pack64(blo, hi), idesc, (i + ks) > 0);
        else tc::mma_f16_w(el, 0u, tc::pack64(alo, hi), tc::pack64(blo, hi), idesc, (i + ks) > 0);
        alo += 256u; blo += 2u * (uint32_t)N; at += 8u;
      }
    }
    tc::commit_w(el, bar);
    t1 = clock64();
    tc::mbar_wait(bar, 0);
    if (tid == 0) { out[0] = t1 - t0; out[1] = clock64() - t0; }
    *stop = 1;
    if (tid == 0) tc::mbar_arrive(never);
  } else if (warp >= 4 && warp < 12 && (mode == 1 || mode == 4)) {
    tc::mbar_wait(never, 0);                       // spinning warps
  } else if (warp >= 12 && warp < 24 && (mode == 2 || mode == 4)) {
    uint4* p = reinterpret_cast<uint4*>(smraw + 96 * 1024) + (tid - 384);   // 12 warps x 32 lanes x 16 B = 6 KB, x4 slots
    uint4 v = make_uint4(tid, 1, 2, 3);
    while (!*stop) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { p[k * 384] = v; }
#pragma unroll
      for (int k = 0; k < 4; ++k) { const uint4 w = p[k * 384]; v.x += w.y; }
    }
    if (v.x == 0xdeadbeef) out[3] = 1;
  } else if (warp >= 24 && warp < 32 && (mode == 3 || mode == 4)) {
    const uint32_t trow = ((uint32_t)(32 * (warp & 3)) << 16) + 384u + 16u * ((warp - 24) >> 2);
    uint32_t r[16];
    while (!*stop) {
      tc::tmem_ld16_nowait(trow, r);
      tc::tmem_wait_ld();
      tc::tmem_st8_nowait(trow, r);
      tc::tmem_st8_nowait(trow + 8, r + 8);
      tc::tmem_wait_st();
    }
    if (r[0] == 0xdeadbeef) out[3] = 1;
  }
  tc::fence_before_sync(); __syncthreads();
  if (tid < 32) tc::tmem_dealloc(0u, 512);
}

template <int FORM, int N>
static void run(long long* d, const char* name) {
  cudaFuncSetAttribute(rate<FORM, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const char* modes[] = {"alone", "+8 warps spinning on an mbarrier", "+12 warps streaming shared memory", "+8 warps tcgen05.ld/st", "all"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) rate<FORM, N><<<1, 1024, 170 * 1024>>>(d, 512, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-14s %-36s: issue %.1f cyc/MMA, complete %.1f cyc/MMA [%s]\n", name, modes[mode], h[0] / 512.0, h[1] / 512.0, cudaGetErrorString(e));
    if (e != cudaSuccess) exit(1);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<0, 64>(d, "SS M=128 N=64");
  run<1, 64>(d, "TS M=128 N=64");
  run<1, 32>(d, "TS M=128 N=32");
  return 0;
}
