// Video retrieval maths (reference: classifier.py:963-983):
//   centre each feature set by its own mean, L2-normalise rows, sim = test @ train^T, top-k per test row.
// The reference does this in fp32 through cuBLAS/MKL, whose summation order is unknowable, so its top-k is
// only defined up to fp32 near-ties. Here the whole chain runs in fp64 from the fp32 features (centre,
// normalise, similarity) and the selection breaks exact ties towards the lowest index: the result is
// reproducible bit for bit and equals torch.topk on the float64 evaluation of the same formula.
#include "host_common.h"

namespace dv {

// mean[c] = (1/n) sum_r x[r][c]   (x fp32 [n][d], mean fp64 [d]); one block per 32 columns
__global__ void colmean_f64_kernel(const float* __restrict__ x, double* __restrict__ mean, int n, int d) {
  __shared__ double part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (c < d)
    for (int r = threadIdx.y; r < n; r += blockDim.y) s += (double)x[(long long)r * d + c];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < d) {
    double t = 0.0;
    for (int j = 0; j < (int)blockDim.y; ++j) t += part[j][threadIdx.x];
    mean[c] = t / (double)n;
  }
}

// y[r] = (x[r] - mean) / max(||x[r] - mean||, eps)   (one warp per row, fp64 out)
__global__ void center_normalize_f64_kernel(const float* __restrict__ x, const double* __restrict__ mean,
                                            double* __restrict__ y, int n, int d) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double ss = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double v = (double)x[(long long)row * d + i] - mean[i];
    ss += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const double inv = 1.0 / fmax(sqrt(ss), 1e-12);
  for (int i = lane; i < d; i += 32)
    y[(long long)row * d + i] = ((double)x[(long long)row * d + i] - mean[i]) * inv;
}

// C[m][n] = sum_k A[m][k] * B[n][k]  (fp64, both row-major with K contiguous); optional fp32 copy
__global__ void __launch_bounds__(256)
dgemm_nt_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                float* __restrict__ C32, int M, int N, int K) {
  __shared__ double As[16][64 + 1];
  __shared__ double Bs[16][64 + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int k = e % 16, r = e / 16;
      As[k][r] = (m0 + r < M && k0 + k < K) ? A[(long long)(m0 + r) * K + k0 + k] : 0.0;
      Bs[k][r] = (n0 + r < N && k0 + k < K) ? B[(long long)(n0 + r) * K + k0 + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        C[(long long)m * N + n] = acc[i][j];
        if (C32) C32[(long long)m * N + n] = (float)acc[i][j];
      }
    }
}

// One block per row: k rounds of arg-max (largest value, ties -> lowest index) over a smem copy.
__global__ void __launch_bounds__(256)
topk_rows_f64_kernel(const double* __restrict__ S, long long* __restrict__ idx, int N, int k) {
  extern __shared__ double row[];
  __shared__ double bv[8];
  __shared__ int bi[8];
  const double* src = S + (long long)blockIdx.x * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) row[i] = src[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < k; ++r) {
    double best = -INFINITY;
    int besti = 0x7fffffff;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double v = row[i];
      if (v > best || (v == best && i < besti)) { best = v; besti = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (lane == 0) { bv[warp] = best; bi[warp] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (bv[w] > bv[0] || (bv[w] == bv[0] && bi[w] < bi[0])) { bv[0] = bv[w]; bi[0] = bi[w]; }
      idx[(long long)blockIdx.x * k + r] = bi[0];
      row[bi[0]] = -INFINITY;
    }
    __syncthreads();
  }
}

int retrieval_prepare(const float* x, double* mean, double* y, int n, int d, cudaStream_t st) {
  colmean_f64_kernel<<<ceil_div(d, 32), dim3(32, 8), 0, st>>>(x, mean, n, d);
  DV_LAUNCH_OK();
  center_normalize_f64_kernel<<<ceil_div(n * 32, 256), 256, 0, st>>>(x, mean, y, n, d);
  DV_LAUNCH_OK();
  return kOk;
}

int retrieval_sim_topk(const double* test, const double* train, double* sim, float* sim32, long long* idx,
                       int nt, int ntr, int d, int k, cudaStream_t st) {
  dgemm_nt_kernel<<<dim3(ceil_div(ntr, 64), ceil_div(nt, 64)), 256, 0, st>>>(test, train, sim, sim32, nt, ntr, d);
  DV_LAUNCH_OK();
  const size_t smem = (size_t)ntr * sizeof(double);
  if (smem > 220 * 1024) return fail(kUnsupported, "retrieval: %d gallery items exceed the shared-memory row buffer", ntr);
  static bool attr = false;
  if (!attr) {
    DV_CUDA_OK(cudaFuncSetAttribute(topk_rows_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  topk_rows_f64_kernel<<<nt, 256, smem, st>>>(sim, idx, ntr, k);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
