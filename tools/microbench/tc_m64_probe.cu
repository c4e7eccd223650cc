// Where does tcgen05.mma.cta_group::1.kind::f16 with M = 64 put its accumulator rows in tensor memory, and can the
// accumulator sit at a lane offset of 64?  (Needed for a decoder tile split into two independent 64-pair half-tiles that
// share every TMEM column on disjoint lanes.)  A[r][0] = r + 1, B[n][0] = 1 -> D[r][n] = r + 1.
#include <cstdio>
#include "tc.cuh"
using namespace dpv;

__global__ void __launch_bounds__(128, 1) probe(float* out, int lane_off, int a_row0) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 32 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < 8 * 1024; e += 128) reinterpret_cast<uint32_t*>(sm)[e] = 0u;
  __syncthreads();
  // A: X8 layout, R = 128 rows, 2 chunks (K = 16): element (r, k=0) = r + 1 (fp16 exact up to 2048)
  __half* A = reinterpret_cast<__half*>(sm);
  A[((size_t)0 * 128 + tid) * 8 + 0] = __float2half((float)(tid + 1));
  // B: X8 layout, R = 8 rows (N = 8): element (n, k=0) = 1
  __half* Bm = reinterpret_cast<__half*>(sm + 16 * 1024);
  if (tid < 8) Bm[((size_t)0 * 8 + tid) * 8 + 0] = __float2half(1.0f);
  if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (tid < 32) tc::tmem_alloc(tptr, 32);
  tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tb = *tptr;
  // clear the accumulator columns in all 128 lanes
  float z[8] = {-1.f, -1.f, -1.f, -1.f, -1.f, -1.f, -1.f, -1.f};
  tc::tmem_st8(tb + ((uint32_t)(32 * (tid >> 5)) << 16), z);
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(64, 8, 0, 0);
    const uint64_t ad = tc::make_desc(tc::smem_u32(sm) + (uint32_t)a_row0 * 16u, 128u * 16u, 128u);
    const uint64_t bd = tc::desc_kmajor(tc::smem_u32(sm + 16 * 1024), 8, 0);
    tc::mma_f16(tb + ((uint32_t)lane_off << 16), ad, bd, idesc, 0);
    tc::commit(bar);
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  float v[8];
  tc::tmem_ld8(tb + ((uint32_t)(32 * (tid >> 5)) << 16), v);
  out[tid] = v[0];
  out[128 + tid] = v[7];
  tc::fence_before_sync(); __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tb, 32);
}

int main() {
  float* d; cudaMalloc(&d, 256 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  const int cfg[3][2] = {{0, 0}, {64, 64}, {64, 0}};
  for (auto& c : cfg) {
    cudaMemset(d, 0, 256 * 4);
    probe<<<1, 128, 40 * 1024>>>(d, c[0], c[1]);
    cudaError_t e = cudaDeviceSynchronize();
    float h[256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=64 MMA, D lane offset %d, A rows from %d: [%s]\n  TMEM lane -> D column 0 value (= A row + 1; -1 = untouched):\n", c[0], c[1], cudaGetErrorString(e));
    for (int l = 0; l < 128; ++l) printf("%s%4.0f", (l % 16 == 0) ? "\n   " : " ", h[l]);
    printf("\n");
    if (e != cudaSuccess) break;
  }
  return 0;
}
