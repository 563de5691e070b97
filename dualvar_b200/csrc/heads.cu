// fp32 pieces after the global average pool: projection-head linears, L2 normalisation and the
// small dense products of the objectives. Everything here stays fp32 so that embeddings, logits
// (cosine / 0.07) and the log-sum-exp keep full precision (SURVEY.md §7 hard part 4).
//
// Replaces on the reference path: the two 1x1x1 nn.Conv3d + ReLU of the heads
// (model/simclr.py:168-180, model/moco.py:283-308), F.normalize (model/simclr.py:359,367,393;
// model/moco.py:498,502,519,521,557) and the matmul/bmm/einsum of the losses
// (model/simclr.py:202,297; model/moco.py:413-414,429-430). SURVEY.md K11, K12.
#include "host_common.h"

namespace dv {

// C[M,N] = alpha * op(A) * op(B) + beta * C (+ bias[n]) (relu). Row-major; op(A) is MxK, op(B) is KxN.
// ta: A stored [K][M] (lda) else [M][K]; tb: B stored [N][K] (ldb) else [K][N].
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda, int ta,
             const float* __restrict__ B, int ldb, int tb, float beta, float* __restrict__ C, int ldc,
             const float* __restrict__ bias, int relu) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);  // (BM/TM) x (BN/TN) = 256 threads, each TM x TN outputs
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int e = tid; e < BM * BK; e += 256) {
      int m, k;
      if (ta) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = ta ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk];
      As[k][m] = v;
    }
    for (int e = tid; e < BN * BK; e += 256) {
      int n, k;
      if (tb) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = tb ? B[(long long)gn * ldb + gk] : B[(long long)gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[gn];
      if (beta != 0.f) v = fmaf(beta, C[(long long)gm * ldc + gn], v);
      if (relu) v = fmaxf(v, 0.f);
      C[(long long)gm * ldc + gn] = v;
    }
  }
}

// out[n] = beta*out[n] + sum_m X[m][n]   (bias gradient)
__global__ void colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int M, int N, int ld,
                              float beta) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += X[(long long)m * ld + n];
  out[n] = (beta != 0.f ? beta * out[n] : 0.f) + s;
}

// dx = dy * (y > 0)
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                float* __restrict__ dx, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// One warp per row of length d: y = x / max(||x||, eps); inv_norm saved for backward.
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                  float* __restrict__ inv_norm, long long rows, int d, float eps) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * d;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) ss = fmaf(xr[i], xr[i], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  for (int i = lane; i < d; i += 32) y[row * d + i] = xr[i] * inv;
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
}

// dx = (dy - y * <y, dy>) * inv_norm
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                  const float* __restrict__ inv_norm, float* __restrict__ dx,
                                  long long rows, int d) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float dot = 0.f;
  for (int i = lane; i < d; i += 32) dot = fmaf(y[row * d + i], dy[row * d + i], dot);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  const float inv = inv_norm[row];
  for (int i = lane; i < d; i += 32) dx[row * d + i] = (dy[row * d + i] - y[row * d + i] * dot) * inv;
}

int sgemm(int ta, int tb, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
          int ldb, float beta, float* C, int ldc, const float* bias, int relu, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return kOk;
  // small head / loss GEMMs: 32x32 tiles so that a 192x512 output still spreads over ~100 CTAs
  if ((long long)ceil_div(N, 64) * ceil_div(M, 64) < 2 * sm_count()) {
    dim3 grid(ceil_div(N, 32), ceil_div(M, 32));
    sgemm_kernel<32, 32, 16, 2, 2><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, ta, B, ldb, tb, beta, C, ldc,
                                                            bias, relu);
  } else {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
    sgemm_kernel<64, 64, 16, 4, 4><<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, ta, B, ldb, tb, beta, C, ldc,
                                                            bias, relu);
  }
  DV_LAUNCH_OK();
  return kOk;
}

int colsum(const float* X, float* out, int M, int N, int ld, float beta, cudaStream_t stream) {
  colsum_kernel<<<ceil_div(N, 128), 128, 0, stream>>>(X, out, M, N, ld, beta);
  DV_LAUNCH_OK();
  return kOk;
}

int relu_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t stream) {
  int grid = (int)std::min<long long>(ceil_div_ll(n, 256), 1184);
  relu_bwd_kernel<<<grid, 256, 0, stream>>>(dy, y, dx, n);
  DV_LAUNCH_OK();
  return kOk;
}

int l2norm_fwd(const float* x, float* y, float* inv_norm, long long rows, int d, float eps,
               cudaStream_t stream) {
  const int threads = 256;
  l2norm_fwd_kernel<<<(unsigned)ceil_div_ll(rows * 32, threads), threads, 0, stream>>>(x, y, inv_norm,
                                                                                       rows, d, eps);
  DV_LAUNCH_OK();
  return kOk;
}

int l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, long long rows, int d,
               cudaStream_t stream) {
  const int threads = 256;
  l2norm_bwd_kernel<<<(unsigned)ceil_div_ll(rows * 32, threads), threads, 0, stream>>>(dy, y, inv_norm,
                                                                                       dx, rows, d);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
