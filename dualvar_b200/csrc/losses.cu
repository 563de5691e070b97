// Contrastive objectives in fp32: NT-Xent / tc / InfoNCE row kernel, shuffle-rank, small helpers.
//
// Replaces on the reference path: the boolean-mask gathers + CrossEntropyLoss
// of model/simclr.py:198-221 and :299-329, model/moco.py:416-418,427-432; the bmm + mask + softplus of
// the shuffle-rank loss (model/simclr.py:242-263, model/moco.py:451-472); the scatter that re-aligns
// shuffled segments (model/simclr.py:389-392, model/moco.py:562-565); calc_topk_accuracy's top-1/top-5
// (utils/utils.py:75-92). The similarity matrices themselves come from sgemm (heads.cu).
#include "host_common.h"

namespace dv {

// One CTA per row r of the similarity matrix S [R][C] (unscaled dot products).
//   self_col[r] : column to drop (-1: none)        pos_col[r] : the positive column
// Writes logits[r] = [S[r][pos], S[r][c] for c ascending, c != self, c != pos] / T   (reference order),
// accumulates loss_sum += logsumexp(logits[r]) - logits[r][0], top-1 / top-5 hits of the positive,
// and overwrites S[r][:] with dLoss/dS = (softmax - onehot) * grad_scale / T (0 in the self column),
// where grad_scale = 1/R for the mean-reduced cross-entropy.
__global__ void __launch_bounds__(256)
contrast_rows_kernel(float* __restrict__ S, float* __restrict__ logits, const int* __restrict__ self_col,
                     const int* __restrict__ pos_col, int C, int ld_s, int ld_logits, float inv_T,
                     float grad_scale, float* __restrict__ loss_sum, int* __restrict__ hits) {
  const int r = blockIdx.x;
  float* row = S + (long long)r * ld_s;
  const int self = self_col ? self_col[r] : -1;
  const int pos = pos_col[r];
  __shared__ float red[32];
  __shared__ int redi[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // __fmul_rn: every pass must see the SAME rounded logit (a contracted fma(row, inv_T, -mx) would leave the
  // product's rounding residual in exp(z - mx) of the maximum itself: 5e-7 absolute, 3e-4 of a 1e-3 loss)
  const float zp = __fmul_rn(row[pos], inv_T);
  // pass 1: max and how many negatives beat the positive
  float mx = -INFINITY;
  int above = 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (c == self) continue;
    const float z = __fmul_rn(row[c], inv_T);
    mx = fmaxf(mx, z);
    if (c != pos && z > zp) ++above;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    above += __shfl_xor_sync(0xffffffffu, above, o);
  }
  if (lane == 0) { red[warp] = mx; redi[warp] = above; }
  __syncthreads();
  mx = red[0];
  above = redi[0];
  for (int w = 1; w < nwarps; ++w) { mx = fmaxf(mx, red[w]); above += redi[w]; }
  __syncthreads();
  // pass 2: sum of exp
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (c == self) continue;
    se += __expf(__fmul_rn(row[c], inv_T) - mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  if (lane == 0) red[warp] = se;
  __syncthreads();
  se = 0.f;
  for (int w = 0; w < nwarps; ++w) se += red[w];
  const float inv_se = 1.f / se;
  if (threadIdx.x == 0) {
    atomicAdd(loss_sum, (mx - zp) + logf(se));   // not lse - zp: exact when the positive is the row maximum
    if (hits) {
      if (above < 1) atomicAdd(&hits[0], 1);
      if (above < 5) atomicAdd(&hits[1], 1);
    }
  }
  // pass 3: logits in reference order + gradient in place
  float* lrow = logits ? logits + (long long)r * ld_logits : nullptr;
  const float gs = grad_scale * inv_T;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float z = __fmul_rn(row[c], inv_T);
    if (c == self) { row[c] = 0.f; continue; }
    if (lrow) {
      int j;
      if (c == pos) j = 0;
      else j = 1 + c - (self >= 0 && c > self ? 1 : 0) - (c > pos ? 1 : 0);
      lrow[j] = z;
    }
    const float pr = __expf(z - mx) * inv_se;
    row[c] = (pr - (c == pos ? 1.f : 0.f)) * gs;
  }
}

// Shuffle-rank loss, one CTA (64 threads) per sample. a, b: [B][s][e] unit vectors of the two "views"
// (rows of the Gram matrix are [a_0..a_{s-1}, b_0..b_{s-1}]). Writes margin logits, accumulates the
// loss and writes da, db (gradient of weight * mean over B*2s*(2s-2) entries of softplus(min(z, clip))),
// z = (second - highest)/theta. clip_max <= 0 disables the clamp (MoCo variant).
constexpr int kRankMaxRows = 8;
__global__ void __launch_bounds__(64)
rank_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ da,
                 float* __restrict__ db, float* __restrict__ logits, float* __restrict__ loss_sum,
                 int* __restrict__ hits, int s, int e, float inv_theta, float clip_max, float weight,
                 float inv_count) {
  const int smp = blockIdx.x;
  const int n = 2 * s;
  __shared__ float G[kRankMaxRows][kRankMaxRows];
  __shared__ float dG[kRankMaxRows][kRankMaxRows];
  extern __shared__ float x[];  // [n][e]
  for (int i = threadIdx.x; i < n * e; i += blockDim.x) {
    const int rrow = i / e, k = i - rrow * e;
    x[i] = rrow < s ? a[((long long)smp * s + rrow) * e + k] : b[((long long)smp * s + (rrow - s)) * e + k];
  }
  for (int i = threadIdx.x; i < kRankMaxRows * kRankMaxRows; i += blockDim.x) (&dG[0][0])[i] = 0.f;
  __syncthreads();
  for (int pq = threadIdx.x; pq < n * n; pq += blockDim.x) {
    const int i = pq / n, j = pq - i * n;
    float d = 0.f;
    for (int k = 0; k < e; ++k) d = fmaf(x[i * e + k], x[j * e + k], d);
    G[i][j] = d;
  }
  __syncthreads();
  if (threadIdx.x < n) {
    const int i = threadIdx.x;
    const int p = (i + s) % n;
    const float hi = G[i][p];
    float lsum = 0.f, dhi = 0.f;
    int col = 1, beaten = 0;
    float* lrow = logits ? logits + ((long long)smp * n + i) * (n - 1) : nullptr;
    if (lrow) lrow[0] = hi;
    for (int j = 0; j < n; ++j) {
      if (j == i || j == p) continue;
      const float z = (G[i][j] - hi) * inv_theta;
      const bool clipped = clip_max > 0.f && z > clip_max;
      const float zc = clipped ? clip_max : z;
      lsum += log1pf(__expf(zc));
      const float dz = clipped ? 0.f : 1.f / (1.f + __expf(-zc));
      const float gsc = weight * inv_count * inv_theta * dz;
      dG[i][j] += gsc;
      dhi -= gsc;
      if (lrow) lrow[col++] = G[i][j];
      if (G[i][j] > hi) ++beaten;
    }
    dG[i][p] += dhi;
    atomicAdd(loss_sum, lsum * weight * inv_count);
    if (hits && beaten == 0) atomicAdd(hits, 1);
  }
  __syncthreads();
  // dx_i = sum_j (dG[i][j] + dG[j][i]) x_j
  for (int i2 = threadIdx.x; i2 < n * e; i2 += blockDim.x) {
    const int i = i2 / e, k = i2 - i * e;
    float g = 0.f;
    for (int j = 0; j < n; ++j) g = fmaf(dG[i][j] + dG[j][i], x[j * e + k], g);
    if (i < s) da[((long long)smp * s + i) * e + k] = g;
    else db[((long long)smp * s + (i - s)) * e + k] = g;
  }
}

// out[b][perm[b][j]][:] = in[b][j][:]  (forward = scatter; inverse=1 gives the gather used in backward)
__global__ void permute_segments_kernel(const float* __restrict__ in, float* __restrict__ out,
                                        const int* __restrict__ perm, int B, int s, int e, int inverse) {
  const long long total = (long long)B * s * e;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % e);
    const int j = (int)((i / e) % s);
    const int bb = (int)(i / ((long long)e * s));
    const int pj = perm[bb * s + j];
    if (!inverse) out[((long long)bb * s + pj) * e + k] = in[i];
    else out[i] = in[((long long)bb * s + pj) * e + k];
  }
}

// out[r][:] = scale * sum_j in[r][j][:]   (segment mean with scale = 1/s; backward: broadcast)
__global__ void segment_sum_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows,
                                   int s, int e, float scale) {
  const long long total = rows * e;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / e;
    const int k = (int)(i - r * e);
    float acc = 0.f;
    for (int j = 0; j < s; ++j) acc += in[(r * s + j) * e + k];
    out[i] = acc * scale;
  }
}

__global__ void segment_bcast_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows,
                                     int s, int e, float scale, float beta) {
  const long long total = rows * s * e;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % e);
    const long long r = i / ((long long)s * e);
    const float v = in[r * e + k] * scale;
    out[i] = (beta != 0.f) ? fmaf(beta, out[i], v) : v;
  }
}

// rowdot: out[r*ld_out] = <a[r], b[r]>   (MoCo positive logit, model/moco.py:429)
__global__ void rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b,
                              float* __restrict__ out, int rows, int d, int ld_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s = fmaf(a[(long long)row * d + i], b[(long long)row * d + i], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[(long long)row * ld_out] = s;
}

// y[i] = beta*y[i] + alpha[row]*x[i]  with alpha read from a strided column (dq += dS[:,0] * k)
__global__ void row_axpy_kernel(const float* __restrict__ alpha, int ld_alpha, const float* __restrict__ x,
                                float* __restrict__ y, int rows, int d, float beta) {
  const long long total = (long long)rows * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d;
    const float v = alpha[r * ld_alpha] * x[i];
    y[i] = (beta != 0.f) ? fmaf(beta, y[i], v) : v;
  }
}

static int small_grid(long long total) {
  long long g = ceil_div_ll(total, 256);
  if (g > 1184) g = 1184;
  if (g < 1) g = 1;
  return (int)g;
}

int contrast_rows(float* S, float* logits, const int* self_col, const int* pos_col, int R, int C,
                  int ld_s, int ld_logits, float inv_T, float grad_scale, float* loss_sum, int* hits,
                  cudaStream_t stream) {
  contrast_rows_kernel<<<R, 256, 0, stream>>>(S, logits, self_col, pos_col, C, ld_s, ld_logits, inv_T,
                                              grad_scale, loss_sum, hits);
  DV_LAUNCH_OK();
  return kOk;
}

int rank_loss(const float* a, const float* b, float* da, float* db, float* logits, float* loss_sum,
              int* hits, int B, int s, int e, float theta, float clip_max, float weight,
              cudaStream_t stream) {
  DV_REQUIRE(2 * s <= kRankMaxRows, "rank loss supports n_series <= %d", kRankMaxRows / 2);
  const float inv_count = 1.f / ((float)B * 2 * s * (2 * s - 2));
  rank_loss_kernel<<<B, 64, (size_t)2 * s * e * sizeof(float), stream>>>(
      a, b, da, db, logits, loss_sum, hits, s, e, 1.f / theta, clip_max, weight, inv_count);
  DV_LAUNCH_OK();
  return kOk;
}

int permute_segments(const float* in, float* out, const int* perm, int B, int s, int e, int inverse,
                     cudaStream_t stream) {
  permute_segments_kernel<<<small_grid((long long)B * s * e), 256, 0, stream>>>(in, out, perm, B, s, e,
                                                                                inverse);
  DV_LAUNCH_OK();
  return kOk;
}

int segment_sum(const float* in, float* out, long long rows, int s, int e, float scale, cudaStream_t stream) {
  segment_sum_kernel<<<small_grid(rows * e), 256, 0, stream>>>(in, out, rows, s, e, scale);
  DV_LAUNCH_OK();
  return kOk;
}

int segment_bcast(const float* in, float* out, long long rows, int s, int e, float scale, float beta,
                  cudaStream_t stream) {
  segment_bcast_kernel<<<small_grid(rows * s * e), 256, 0, stream>>>(in, out, rows, s, e, scale, beta);
  DV_LAUNCH_OK();
  return kOk;
}

int rowdot(const float* a, const float* b, float* out, int rows, int d, int ld_out, cudaStream_t stream) {
  rowdot_kernel<<<ceil_div(rows * 32, 256), 256, 0, stream>>>(a, b, out, rows, d, ld_out);
  DV_LAUNCH_OK();
  return kOk;
}

int row_axpy(const float* alpha, int ld_alpha, const float* x, float* y, int rows, int d, float beta,
             cudaStream_t stream) {
  row_axpy_kernel<<<small_grid((long long)rows * d), 256, 0, stream>>>(alpha, ld_alpha, x, y, rows, d, beta);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
