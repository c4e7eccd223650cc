// Probe: which shared-memory float does tcgen05 read for element (MN = j, K = k) of an MN-major, no-swizzle
// B operand?  A = identity (K-major, known good), B buffer filled with its own float index.
#include <cstdio>
#include <vector>
#include "tc.cuh"
using namespace dpv;

__global__ void __launch_bounds__(128, 1) probe(float* D, int N, uint32_t lbo, uint32_t sbo, int kstep_bytes, int bmajor, int nsteps) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  float4* A4 = reinterpret_cast<float4*>(smraw);            // identity: 128 rows x 64 cols  -> [16][128]
  float* Bf = reinterpret_cast<float*>(A4 + 16 * 128);       // 2048 floats = their own index
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bf + 2048);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < 16 * 128; e += 128) {
    const int ch = e / 128, r = e - ch * 128;
    float v[4];
    for (int i = 0; i < 4; ++i) v[i] = (r == 4 * ch + i) ? 1.f : 0.f;
    A4[e] = make_float4(v[0], v[1], v[2], v[3]);
  }
  for (int e = tid; e < 2048; e += 128) Bf[e] = (float)e;
  if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (tid < 32) tc::tmem_alloc(tptr, 64);
  tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tbase = *tptr;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(128, N, 0, bmajor);
    uint32_t acc = 0;
    for (int s = 0; s < nsteps; ++s) {
      tc::mma_tf32(tbase, tc::desc_kmajor(tc::smem_u32(A4), 128, 2 * s), tc::make_desc(tc::smem_u32(Bf) + s * kstep_bytes, lbo, sbo), idesc, acc);
      acc = 1;
    }
    tc::commit(bar);
  }
  tc::mbar_wait(bar, 0); tc::fence_after_sync();
  const int warp = tid >> 5, lane = tid & 31;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * N + c0 + i] = v[i];
  }
  tc::fence_before_sync(); __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tbase, 64);
}

static void run(const char* name, int N, uint32_t lbo, uint32_t sbo, int kstep_bytes, int bmajor, int nsteps) {
  float* dD; cudaMalloc(&dD, 128 * N * 4); cudaMemset(dD, 0, 128 * N * 4);
  size_t smem = 16 * 128 * 16 + 2048 * 4 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dD, N, lbo, sbo, kstep_bytes, bmajor, nsteps);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(128 * N); cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  printf("== %s: N=%d lbo=%u sbo=%u kstep=%d bmajor=%d [%s]\n   D[k][j] = float index read for (K=k, MN=j)\n", name, N, lbo, sbo, kstep_bytes, bmajor, cudaGetErrorString(e));
  for (int k = 0; k < 8 * nsteps && k < 20; ++k) {
    printf("   k=%2d:", k);
    for (int j = 0; j < N && j < 20; ++j) printf(" %5.0f", D[k * N + j]);
    printf("\n");
  }
  cudaFree(dD);
}

int main() {
  run("K-major reference", 16, 256, 128, 512, 0, 2);          // R=16 rows: LBO = 16*16
  run("MN-major lbo=128 sbo=256", 16, 128, 256, 128, 1, 2);
  run("MN-major lbo=256 sbo=128", 16, 256, 128, 128, 1, 2);
  run("MN-major lbo=128 sbo=512", 16, 128, 512, 128, 1, 2);
  run("MN-major lbo=512 sbo=128", 16, 512, 128, 128, 1, 2);
  run("MN-major lbo=1024 sbo=2048 N=32", 32, 1024, 2048, 128, 1, 1);
  return 0;
}
