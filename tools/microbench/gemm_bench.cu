// Micro-benchmark of the shared-memory GEMM building blocks (cycles per call, one CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I dpivae_b200/csrc -o gemm_bench tools/microbench/gemm_bench.cu
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace dpv;

template <int WHICH>
__global__ void __launch_bounds__(NT, 1) bench_kernel(long long* out, float* sink, float* scratch, int K, int N, int iters) {
  extern __shared__ __align__(16) float sm[];
  const int ldw = pad4(N) + 4;
  float* Wt = sm;                       // [128][132] max
  float* bias = Wt + 128 * 132;
  float* A = bias + 128;                // [128][LDP]
  float* O = A + 128 * LDP;             // [128][LDP]
  for (int e = threadIdx.x; e < 128 * 132 + 128 + 256 * LDP; e += NT) sm[e] = 0.001f * (e % 97);
  __syncthreads();
  float* part = scratch + (size_t)blockIdx.x * 128 * 128;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (WHICH == 0) gemm_fwd<ACT_NONE>(Wt, ldw, bias, A, O, K, pad4(N));
    if (WHICH == 1) gemm_fwd<ACT_TANH>(Wt, ldw, bias, A, O, K, pad4(N));
    if (WHICH == 2) gemm_dgrad<ACT_RELU>(Wt, ldw, O, A, A, pad4(K), pad4(N));
    if (WHICH == 3) gemm_wgrad(A, O, part, part + 128 * 127, K, N);
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
  if (O[threadIdx.x] == 123.456f) sink[0] = A[threadIdx.x];
}

template <int WHICH>
void run(const char* name, int K, int N, double macs) {
  long long* d_out; float* d_sink; float* d_scr;
  cudaMalloc(&d_out, 148 * 8); cudaMalloc(&d_sink, 4); cudaMalloc(&d_scr, (size_t)148 * 128 * 128 * 4);
  cudaMemset(d_scr, 0, (size_t)148 * 128 * 128 * 4);
  size_t smem = (128 * 132 + 128 + 256 * LDP) * 4;
  cudaFuncSetAttribute(bench_kernel<WHICH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench_kernel<WHICH><<<148, NT, smem>>>(d_out, d_sink, d_scr, K, N, 50);
  bench_kernel<WHICH><<<148, NT, smem>>>(d_out, d_sink, d_scr, K, N, 200);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
  double ideal = macs * 64 / 128.0;
  printf("%-22s K=%3d N=%3d  %7lld cyc/call  ideal %6.0f  eff %4.1f%%  (%s)\n", name, K, N, h[0], ideal, 100.0 * ideal / h[0],
         cudaGetErrorString(e));
  cudaFree(d_out); cudaFree(d_sink); cudaFree(d_scr);
}

int main() {
  const int shapes[][2] = {{64, 64}, {128, 64}, {8, 128}, {64, 32}, {32, 64}, {3, 64}, {64, 4}, {4, 64}, {64, 128}};
  for (auto& s : shapes) {
    int K = s[0], N = s[1];
    run<0>("fwd", K, N, (double)K * N);
    run<1>("fwd+tanh", K, N, (double)K * N);
    run<2>("dgrad+relu", K, N, (double)K * N);
    run<3>("wgrad", K, N, (double)K * N);
  }
  return 0;
}
