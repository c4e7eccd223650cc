// FFMA issue-rate of register outer-product tiles (no memory traffic): how close does a TMxTN
// outer product get to 128 FFMA/clk/SM with 8 warps per SM?
#include <cstdio>
#include <vector>
template <int TM, int TN>
__global__ void __launch_bounds__(256, 1) k(long long* out, float* sink, int iters, float seed) {
  float acc[TN][TM], a[TM], w[TN];
  for (int i = 0; i < TM; ++i) a[i] = seed + threadIdx.x * 0.001f + i;
  for (int j = 0; j < TN; ++j) w[j] = seed * 0.5f + j;
  for (int j = 0; j < TN; ++j) for (int i = 0; i < TM; ++i) acc[j][i] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int i = 0; i < TM; ++i) acc[j][i] = fmaf(w[j], a[i], acc[j][i]);
      // perturb operands so the compiler cannot hoist; 1 op per operand per 'k-step'
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] += 1.0f;
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] += 1.0f;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < TN; ++j) for (int i = 0; i < TM; ++i) s += acc[j][i];
  if (s == 1.2345f) sink[0] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
template <int TM, int TN>
void run() {
  long long* d; float* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  k<TM, TN><<<148, 256>>>(d, s, iters, 1.0f);
  k<TM, TN><<<148, 256>>>(d, s, iters, 1.0f);
  cudaDeviceSynchronize();
  std::vector<long long> h(148); cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
  double ffma = (double)iters * 4 * TM * TN * 8;  // warp-instr per SM
  double other = (double)iters * 4 * (TM + TN) * 8;
  printf("tile %dx%d: %lld cycles; FFMA warp-instr/clk/SM = %.2f (peak 4); incl. FADD %.2f\n", TM, TN, h[0], ffma / h[0], (ffma + other) / h[0]);
}
int main() { run<4, 4>(); run<8, 4>(); run<8, 8>(); run<2, 4>(); run<1, 4>(); return 0; }
