// Known-answer test of the tcgen05 building blocks in dpivae_b200/csrc/tc.cuh: forward / dgrad / wgrad
// orientations of the no-swizzle X8 (fp16) layout, hi/lo split accuracy, accumulation, TMEM st/ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I dpivae_b200/csrc -o tc_gemm_test tools/microbench/tc_gemm_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc.cuh"
using namespace dpv;

// mode 0: D[128][N] = A[128][K] * W[N][K]^T      (X = A: R=128, C=K ; Y = W: R=N, C=K)
// mode 1: D[128][Kin] = G[128][Nout] * W[Nout][Kin]  (X = G: R=128, C=Nout ; Y = W: R=Nout, C=Kin)
// mode 2: D[128][N] = H[Rr][128]^T * G[Rr][N]     (X = H: R=Rr, C=128 ; Y = G: R=Rr, C=N), issued twice (accumulate)
__global__ void __launch_bounds__(128, 1) test_kernel(int mode, int terms, const float* __restrict__ X, int RX, int CX,
                                                       const float* __restrict__ Y, int RY, int CY, float* __restrict__ D, int ND) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  uint4* Xh = reinterpret_cast<uint4*>(smraw);
  uint4* Xl = Xh + (CX / 8) * RX;
  uint4* Yh = Xl + (CX / 8) * RX;
  uint4* Yl = Yh + (CY / 8) * RY;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Yl + (CY / 8) * RY);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < (CX / 8) * RX; e += 128) {
    const int ch = e / RX, r = e - ch * RX;
    tc::split8(X + r * CX + 8 * ch, Xh[e], Xl[e]);
  }
  for (int e = tid; e < (CY / 8) * RY; e += 128) {
    const int ch = e / RY, r = e - ch * RY;
    tc::split8(Y + r * CY + 8 * ch, Yh[e], Yl[e]);
  }
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_fence_init();
  }
  if (tid < 32) tc::tmem_alloc(tptr, 256);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (mode == 3) {
    // stage A (row = this thread's TMEM lane) as packed halves: hi plane then lo plane
    const uint32_t tb0 = *tptr;
    const int row = tid;
    for (int c0 = 0; c0 < CX; c0 += 16) {
      float vh[8], vl[8];
      for (int j = 0; j < 8; ++j) {
        const float a0 = X[row * CX + c0 + 2 * j], a1 = X[row * CX + c0 + 2 * j + 1];
        const __half2 h = __floats2half2_rn(a0, a1);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
        vh[j] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
        vl[j] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
      }
      tc::tmem_st8(tb0 + ((uint32_t)((tid >> 5) * 32) << 16) + 64 + c0 / 2, vh);
      tc::tmem_st8(tb0 + ((uint32_t)((tid >> 5) * 32) << 16) + 64 + CX / 2 + c0 / 2, vl);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tptr;
  const tc::Op xo{tc::smem_u32(Xh), (uint32_t)((CX / 8) * RX * 16), RX}, yo{tc::smem_u32(Yh), (uint32_t)((CY / 8) * RY * 16), RY};
  if (tid == 0) {
    if (mode == 0) tc::issue_fwd(tbase, xo, yo, RY, CX, 0, terms);
    if (mode == 1) tc::issue_dgrad(tbase, xo, yo, RY, CY, 0, terms);
    if (mode == 3) {
      // A from tensor memory (columns 64.. : hi plane K/2 columns, then lo plane), W K-major in shared memory
      const uint32_t idesc = tc::make_idesc(128, RY, 0, 0);
      uint32_t acc = 0;
      for (int t = 0; t < terms; ++t)
        for (int k = 0; k < CX; k += 16) {
          const uint32_t a_t = tbase + 64 + (t == 1 ? CX / 2 : 0) + k / 2;
          tc::mma_f16_ts(tbase, a_t, tc::desc_kmajor(yo.base + (t == 2 ? yo.lo_off : 0u), RY, k >> 3), idesc, acc);
          acc = 1;
        }
    }
    if (mode == 2) {
      tc::issue_wgrad(tbase, xo, yo, CY, 0, terms);
      tc::issue_wgrad(tbase, xo, yo, CY, 1, terms);  // second pass accumulates: result = 2x
    }
    tc::commit(bar);
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < ND; c0 += 16) {
    float v[16];
    tc::tmem_ld16(trow + c0, v);
    for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * ND + c0 + i] = v[i];
  }
  {  // TMEM st/ld round trip on columns 128..159
    float w[32], u[32];
    for (int i = 0; i < 32; ++i) w[i] = (float)(tid * 100 + i);
    tc::tmem_st32(trow + 128, w);
    tc::tmem_ld32(trow + 128, u);
    bool ok = true;
    for (int i = 0; i < 32; ++i) ok = ok && (u[i] == w[i]);
    if (!ok) D[0] = NAN;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tbase, 256);
}

static void run(const char* name, int mode, int terms, int RX, int CX, int RY, int CY, int ND, float sx = 1.f, float sy = 1.f) {
  std::vector<float> X((size_t)RX * CX), Y((size_t)RY * CY), D((size_t)128 * ND, -7.f);
  std::vector<double> ref((size_t)128 * ND, 0.0);
  for (auto& v : X) v = sx * ((float)rand() / RAND_MAX - 0.5f);
  for (auto& v : Y) v = sy * ((float)rand() / RAND_MAX - 0.5f);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < ND; ++n) {
      double s = 0;
      if (mode == 0 || mode == 3) for (int k = 0; k < CX; ++k) s += (double)X[m * CX + k] * Y[n * CY + k];
      if (mode == 1) for (int k = 0; k < CX; ++k) s += (double)X[m * CX + k] * Y[k * CY + n];
      if (mode == 2) { for (int r = 0; r < RX; ++r) s += (double)X[r * CX + m] * Y[r * CY + n]; s *= 2; }
      ref[m * ND + n] = s;
    }
  float *dX, *dY, *dD;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dY, Y.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  size_t smem = (size_t)(RX * CX + RY * CY) * 4 + 64;
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  test_kernel<<<1, 128, smem>>>(mode, terms, dX, RX, CX, dY, RY, CY, dD, ND);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double num = 0, den = 0, maxerr = 0;
  for (size_t i = 0; i < D.size(); ++i) {
    double d = (double)D[i] - ref[i];
    num += d * d; den += ref[i] * ref[i];
    if (!(fabs(d) <= maxerr)) maxerr = fabs(d);
  }
  const double rel = sqrt(num / den);
  const double tol = terms == 3 ? 2e-6 : 2e-3;
  printf("%-30s mode %d terms %d: rel-L2 %.3e max|err| %.3e %s  [%s]\n", name, mode, terms, rel, maxerr, rel < tol ? "OK" : "MISMATCH",
         cudaGetErrorString(e));
  cudaFree(dX); cudaFree(dY); cudaFree(dD);
}

int main() {
  srand(1);
  for (int terms = 1; terms <= 3; terms += 2) {
    run("fwd   M128 N64 K64", 0, terms, 128, 64, 64, 64, 64);
    run("fwd   M128 N128 K16", 0, terms, 128, 16, 128, 16, 128);
    run("fwd   M128 N16 K128", 0, terms, 128, 128, 16, 128, 16);
    run("fwd   M128 N32 K64", 0, terms, 128, 64, 32, 64, 32);
    run("dgrad M128 Nout64 Kin128", 1, terms, 128, 64, 64, 128, 128);
    run("dgrad M128 Nout128 Kin16", 1, terms, 128, 128, 128, 16, 16);
    run("dgrad M128 Nout64 Kin32", 1, terms, 128, 64, 64, 32, 32);
    run("wgrad R128 M128 N64", 2, terms, 128, 128, 128, 64, 64);
    run("wgrad R128 M128 N16", 2, terms, 128, 128, 128, 16, 16);
    run("fwd A-in-TMEM N64 K64", 3, terms, 128, 64, 64, 64, 64);
    run("fwd A-in-TMEM N32 K128", 3, terms, 128, 128, 32, 128, 32);
  }
  // small-magnitude operands (precision of the lo halves near the fp16 subnormal range)
  run("fwd small x (1e-2) K64", 0, 3, 128, 64, 64, 64, 64, 1e-2f, 1.f);
  run("fwd small x (1e-3) K64", 0, 3, 128, 64, 64, 64, 64, 1e-3f, 1.f);
  run("fwd scaled w (x512) K64", 0, 3, 128, 64, 64, 64, 64, 1.f, 512.f);
  return 0;
}
