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

// The same product for the SMALL head / loss GEMMs (a few hundred rows and columns, K <= 512): 32 x 32 tiles so that the
// output still spreads over ~100 CTAs, K tiles of 32 held in two shared-memory buffers with the next tile's global loads
// issued BEFORE the current tile's arithmetic (register prefetch). The one-buffer kernel above waited a full global-load
// latency per K tile - 36 us per launch, 0.8 ms per step for the 22 head launches - this one overlaps it.
__global__ void __launch_bounds__(256)
sgemm_small_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda, int ta,
                   const float* __restrict__ B, int ldb, int tb, float beta, float* __restrict__ C, int ldc,
                   const float* __restrict__ bias, int relu) {
  constexpr int BM = 32, BN = 32, BK = 32;
  __shared__ float As[2][BK][BM + 1];
  __shared__ float Bs[2][BK][BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, each a 2 x 2 block of outputs
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // element e = tid + 256 * i of a 32 x 32 tile: the index that is contiguous in memory varies fastest across threads
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      int m, k;
      if (ta) { m = e & 31; k = e >> 5; } else { k = e & 31; m = e >> 5; }
      const int gm = m0 + m, gka = k0 + k;
      ra[i] = (gm < M && gka < K) ? (ta ? A[(long long)gka * lda + gm] : A[(long long)gm * lda + gka]) : 0.f;
      int n;
      if (tb) { k = e & 31; n = e >> 5; } else { n = e & 31; k = e >> 5; }
      const int gn = n0 + n, gkb = k0 + k;
      rb[i] = (gn < N && gkb < K) ? (tb ? B[(long long)gn * ldb + gkb] : B[(long long)gkb * ldb + gn]) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      int m, k;
      if (ta) { m = e & 31; k = e >> 5; } else { k = e & 31; m = e >> 5; }
      As[buf][k][m] = ra[i];
      int n;
      if (tb) { k = e & 31; n = e >> 5; } else { n = e & 31; k = e >> 5; }
      Bs[buf][k][n] = rb[i];
    }
  };
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const int nk = (K + BK - 1) / BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float a0 = As[buf][k][ty * 2], a1 = As[buf][k][ty * 2 + 1];
      const float b0 = Bs[buf][k][tx * 2], b1 = Bs[buf][k][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    if (kt + 1 < nk) {
      stash(buf ^ 1);      // the other buffer: its last readers passed the barrier at the end of the previous iteration
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int gm = m0 + ty * 2 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int gn = n0 + tx * 2 + j;
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
    sgemm_small_kernel<<<grid, 256, 0, stream>>>(M, N, K, alpha, A, lda, ta, B, ldb, tb, beta, C, ldc, bias, relu);
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
