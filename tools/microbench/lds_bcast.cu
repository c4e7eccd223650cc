// How many shared-memory cycles does a warp-wide LDS.128 / LDS.64 / LDS.32 cost when lanes share addresses?
#include <cstdio>
#include <vector>
template <int V, int DISTINCT>   // V floats per lane; DISTINCT = number of distinct V-float chunks per warp
__global__ void __launch_bounds__(256, 1) k(long long* out, float* sink, int iters) {
  __shared__ __align__(16) float sm[8192];
  for (int e = threadIdx.x; e < 8192; e += 256) sm[e] = e;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = lane % DISTINCT;                 // lanes share chunks
  const float* base = sm + warp * 512 + chunk * V;   // contiguous distinct chunks
  float acc[4][4] = {};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float* p = base + ((it + r) & 3) * 128;
      unsigned addr = (unsigned)__cvta_generic_to_shared(p);
      if (V == 4) { float x, y, z, w; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr)); acc[r & 3][0] += x; acc[r & 3][1] += y; acc[r & 3][2] += z; acc[r & 3][3] += w; }
      else if (V == 2) { float x, y; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr)); acc[r & 3][0] += x; acc[r & 3][1] += y; }
      else { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr)); acc[r & 3][0] += x; }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  float tot = 0.f; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) tot += acc[i][j];
  if (tot == 1.2345f) sink[0] = tot;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
template <int V, int D>
void run() {
  long long* d; float* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  k<V, D><<<148, 256>>>(d, s, iters); k<V, D><<<148, 256>>>(d, s, iters);
  cudaDeviceSynchronize();
  std::vector<long long> h(148); cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
  double n = (double)iters * 16 * 8;  // warp-level LDS per SM
  printf("LDS.%-3d distinct chunks/warp %2d (%4d B): %.2f cycles per warp-LDS\n", V * 32, D, D * V * 4, h[0] / n);
}
int main() {
  run<4, 1>(); run<4, 2>(); run<4, 4>(); run<4, 8>(); run<4, 16>(); run<4, 32>();
  run<2, 1>(); run<2, 16>(); run<2, 32>(); run<1, 1>(); run<1, 32>();
  return 0;
}
