// Post-processing of encode / sample outputs on the device: the callers immediately downstream of
// the hot path -- evaluate_model (dpivae.py:527-559: MC mean of y_sample, then utils/metrics.py:11-32 R2 / MSE / MAE) and
// disentanglement_metric with the linear regressor (dpivae.py:618-703: least squares of every generative factor on each
// latent group, test-set R2).  All accumulations in fp64; the k <= 8 normal equations are solved by one thread.
#include "common.cuh"
#include "kernels.h"

namespace dpv {

namespace {

constexpr int MT = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < MT / 32; ++w) s += red[w];
  return s;   // valid in thread 0
}

// out[b][j] = mean over the leading (MC) axis of v[m][b][j]
__global__ void __launch_bounds__(MT) mc_mean_kernel(const float* __restrict__ v, int n, long long BD, float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * MT + threadIdx.x;
  if (e >= BD) return;
  float s = 0.0f;
  for (int m = 0; m < n; ++m) s += v[(long long)m * BD + e];
  out[e] = s / (float)n;
}

// per output column j: acc[j] = {sum y, sum y^2, sum (y - p)^2, sum |y - p|}
__global__ void __launch_bounds__(MT) metric_sums_kernel(const float* __restrict__ y, const float* __restrict__ p, long long N, int d,
                                                         double* __restrict__ acc) {
  __shared__ double red[MT / 32];
  for (int j = 0; j < d; ++j) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * MT + threadIdx.x; i < N; i += (long long)gridDim.x * MT) {
      const double yy = y[i * d + j], r = yy - (double)p[i * d + j];
      s[0] += yy; s[1] += yy * yy; s[2] += r * r; s[3] += fabs(r);
    }
    for (int t = 0; t < 4; ++t) {
      const double tot = block_sum(s[t], red);
      if (threadIdx.x == 0) atomicAdd(acc + 4 * j + t, tot);
    }
  }
}

// sklearn r2_score / mean_squared_error / mean_absolute_error with multioutput="raw_values" (utils/metrics.py:29-31):
// out = [R2_0 .. R2_{d-1} | MSE_0 .. | MAE_0 ..], one value per output column
__global__ void metric_final_kernel(const double* __restrict__ acc, long long N, int d, float* __restrict__ out3d) {
  for (int j = 0; j < d; ++j) {
    const double sy = acc[4 * j], syy = acc[4 * j + 1], sr = acc[4 * j + 2], sa = acc[4 * j + 3];
    const double sst = syy - sy * sy / (double)N;
    out3d[j] = (float)(sst != 0.0 ? 1.0 - sr / sst : (sr == 0.0 ? 1.0 : 0.0));
    out3d[d + j] = (float)(sr / (double)N);
    out3d[2 * d + j] = (float)(sa / (double)N);
  }
}

// normal equations of y ~ [1, x_0 .. x_{k-1}]: acc = upper triangle of A^T A ((k+1)(k+2)/2 values) followed by A^T y (k+1)
__global__ void __launch_bounds__(MT) normal_eq_kernel(const float* __restrict__ X, const float* __restrict__ y, long long ldy,
                                                       long long N, int k, double* __restrict__ acc) {
  __shared__ double red[MT / 32];
  const int K1 = k + 1;
  double s[9 * 10 / 2 + 9];
  const int ns = K1 * (K1 + 1) / 2 + K1;
  for (int t = 0; t < ns; ++t) s[t] = 0.0;
  for (long long i = (long long)blockIdx.x * MT + threadIdx.x; i < N; i += (long long)gridDim.x * MT) {
    double a[9];
    a[0] = 1.0;
    for (int j = 0; j < k; ++j) a[j + 1] = X[i * k + j];
    const double yy = y[i * ldy];
    int t = 0;
    for (int r = 0; r < K1; ++r)
      for (int c = r; c < K1; ++c) s[t++] += a[r] * a[c];
    for (int r = 0; r < K1; ++r) s[t++] += a[r] * yy;
  }
  for (int t = 0; t < ns; ++t) {
    const double tot = block_sum(s[t], red);
    if (threadIdx.x == 0) atomicAdd(acc + t, tot);
  }
}

// one thread: Cholesky solve of the (k+1) x (k+1) system; coefficients (intercept first) to coef[0..k]
__global__ void normal_solve_kernel(const double* __restrict__ acc, int k, double* __restrict__ coef) {
  const int K1 = k + 1;
  double A[9][9], b[9];
  int t = 0;
  for (int r = 0; r < K1; ++r)
    for (int c = r; c < K1; ++c) { A[r][c] = acc[t]; A[c][r] = acc[t]; ++t; }
  for (int r = 0; r < K1; ++r) b[r] = acc[t++];
  // column scaling (Jacobi preconditioning) keeps the factorisation well conditioned for latents of very different scales
  double sc[9];
  for (int r = 0; r < K1; ++r) sc[r] = A[r][r] > 0.0 ? 1.0 / sqrt(A[r][r]) : 1.0;
  for (int r = 0; r < K1; ++r) {
    b[r] *= sc[r];
    for (int c = 0; c < K1; ++c) A[r][c] *= sc[r] * sc[c];
  }
  double L[9][9];
  for (int i = 0; i < K1; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i][j];
      for (int m = 0; m < j; ++m) s -= L[i][m] * L[j][m];
      if (i == j) L[i][i] = s > 1e-300 ? sqrt(s) : 1e-150;   // rank-deficient column: coefficient ~ 0 after back-substitution
      else L[i][j] = s / L[j][j];
    }
  double z[9];
  for (int i = 0; i < K1; ++i) {
    double s = b[i];
    for (int m = 0; m < i; ++m) s -= L[i][m] * z[m];
    z[i] = s / L[i][i];
  }
  for (int i = K1 - 1; i >= 0; --i) {
    double s = z[i];
    for (int m = i + 1; m < K1; ++m) s -= L[m][i] * coef[m];
    coef[i] = s / L[i][i];
  }
  for (int i = 0; i < K1; ++i) coef[i] *= sc[i];
}

// test-set sums for R2 of the fitted model: acc = {sum y, sum y^2, sum (y - yhat)^2}
__global__ void __launch_bounds__(MT) linreg_score_kernel(const float* __restrict__ X, const float* __restrict__ y, long long ldy,
                                                          long long N, int k, const double* __restrict__ coef, double* __restrict__ acc) {
  __shared__ double red[MT / 32];
  double s[3] = {0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * MT + threadIdx.x; i < N; i += (long long)gridDim.x * MT) {
    double p = coef[0];
    for (int j = 0; j < k; ++j) p += coef[j + 1] * (double)X[i * k + j];
    const double yy = y[i * ldy], r = yy - p;
    s[0] += yy; s[1] += yy * yy; s[2] += r * r;
  }
  for (int t = 0; t < 3; ++t) {
    const double tot = block_sum(s[t], red);
    if (threadIdx.x == 0) atomicAdd(acc + t, tot);
  }
}

__global__ void linreg_r2_final_kernel(const double* __restrict__ acc, long long N, float* __restrict__ r2) {
  const double sst = acc[1] - acc[0] * acc[0] / (double)N;
  *r2 = (float)(sst != 0.0 ? 1.0 - acc[2] / sst : (acc[2] == 0.0 ? 1.0 : 0.0));
}

int grid_for(long long N) {
  long long g = (N + MT - 1) / MT;
  return (int)(g < 1 ? 1 : (g > 1184 ? 1184 : g));
}

}  // namespace

void launch_mc_mean(const float* v, int n, long long BD, float* out, cudaStream_t s) {
  mc_mean_kernel<<<(unsigned)((BD + MT - 1) / MT), MT, 0, s>>>(v, n, BD, out);
}

// scratch: 4 * d doubles
void launch_regression_metrics(const float* y, const float* p, long long N, int d, double* scratch, float* out3, cudaStream_t s) {
  cudaMemsetAsync(scratch, 0, sizeof(double) * 4 * d, s);
  metric_sums_kernel<<<grid_for(N), MT, 0, s>>>(y, p, N, d, scratch);
  metric_final_kernel<<<1, 1, 0, s>>>(scratch, N, d, out3);
}

// scratch: 80 doubles (normal equations 54 + coefficients 9 + score sums 3)
void launch_linreg_r2(const float* Xtr, const float* ytr, long long ldy_tr, long long Ntr, const float* Xte, const float* yte,
                      long long ldy_te, long long Nte, int k, double* scratch, float* r2, cudaStream_t s) {
  cudaMemsetAsync(scratch, 0, sizeof(double) * 80, s);
  normal_eq_kernel<<<grid_for(Ntr), MT, 0, s>>>(Xtr, ytr, ldy_tr, Ntr, k, scratch);
  normal_solve_kernel<<<1, 1, 0, s>>>(scratch, k, scratch + 56);
  linreg_score_kernel<<<grid_for(Nte), MT, 0, s>>>(Xte, yte, ldy_te, Nte, k, scratch + 56, scratch + 72);
  linreg_r2_final_kernel<<<1, 1, 0, s>>>(scratch + 72, Nte, r2);
}

}  // namespace dpv
