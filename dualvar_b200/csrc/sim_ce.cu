// Similarity GEMM fused with the row-wise log-sum-exp / cross-entropy of the contrastive objectives (fp32).
//
// Replaces on the reference path: `torch.matmul(features, features.T)` + diagonal / positive masks + CrossEntropyLoss
// of NT-Xent and the tc loss (model/simclr.py:198-221, :299-329) and `einsum('nc,ck->nk', q, queue)` + cat + CE of
// MoCo's InfoNCE (model/moco.py:426-438); SURVEY.md K13 / K17.
//
//   sim_ce_fwd    : one launch computes the tile S[r0:r0+32][c0:c0+128] = A[rows] . B[cols]^T on CUDA cores (d <= a few
//                   hundred: the whole product is < 0.3 GFLOP even for MoCo's 64 x 16384 queue logits) and, while the
//                   tile is still in registers, (a) stores S, (b) stores the tile's logits z = S / T at their place in
//                   the reference's column order (positive first, self column dropped, rest ascending) and (c) reduces
//                   every row's tile maximum and sum of exp(z - max) - the online-softmax partial of that
//                   (row, column block).
//   sim_ce_finish : one CTA per row combines the partials (plus columns that did not come from the GEMM: MoCo's
//                   positive q.k in column 0) into the row's log-sum-exp, adds the row's cross-entropy to the loss,
//                   counts top-1 / top-5 (utils/utils.py:75-92) and overwrites S with dLoss/dS for the backward GEMMs.
// Compared with sgemm + contrast_rows (losses.cu) the similarity matrix is written once and read once instead of once
// and three times, and the logits never wait for a second kernel.
#include "host_common.h"

namespace dv {

constexpr int kSimRows = 32;    // rows of S per CTA
constexpr int kSimCols = 128;   // columns of S per CTA
constexpr int kSimK = 32;       // k tile

// a: [R][lda] row features. b: column features, row-major [C][ldb] (b_dmajor = 0) or d-major [d][ldb] (MoCo's queue
// buffer, b_dmajor = 1). The tile's columns are S columns col0 + c (col0 = 1 for MoCo: column 0 holds q.k).
// self_col / pos_col are S-column indices per row (self_col may be NULL or hold -1: no column dropped).
__global__ void __launch_bounds__(256)
sim_ce_fwd_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, int b_dmajor, int R, int C,
                  int d, float* __restrict__ S, int ld_s, int col0, float* __restrict__ logits, int ld_logits,
                  const int* __restrict__ self_col, const int* __restrict__ pos_col, float inv_T,
                  float* __restrict__ partials, int nblk) {
  __shared__ float As[kSimRows][kSimK + 1];
  __shared__ float Bs[kSimCols][kSimK + 1];
  const int r0 = blockIdx.y * kSimRows, c0 = blockIdx.x * kSimCols;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // thread -> rows ty*4 .. ty*4+3, columns tx + 32*j
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < d; k0 += kSimK) {
    // A tile: 32 rows x 32 k (one element per 4 threads' worth: 1024 elements / 256 threads)
    for (int e = threadIdx.x; e < kSimRows * kSimK; e += 256) {
      const int rr = e >> 5, kk = e & 31;
      const int r = r0 + rr, k = k0 + kk;
      As[rr][kk] = (r < R && k < d) ? a[(long long)r * lda + k] : 0.f;
    }
    if (b_dmajor) {
      for (int e = threadIdx.x; e < kSimCols * kSimK; e += 256) {
        const int kk = e >> 7, cc = e & 127;             // consecutive threads -> consecutive columns (contiguous)
        const int c = c0 + cc, k = k0 + kk;
        Bs[cc][kk] = (c < C && k < d) ? b[(long long)k * ldb + c] : 0.f;
      }
    } else {
      for (int e = threadIdx.x; e < kSimCols * kSimK; e += 256) {
        const int cc = e >> 5, kk = e & 31;               // consecutive threads -> consecutive k (contiguous)
        const int c = c0 + cc, k = k0 + kk;
        Bs[cc][kk] = (c < C && k < d) ? b[(long long)c * ldb + k] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kSimK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[ty * 4 + i][kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[tx + 32 * j][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // epilogue: S, logits in reference order, online-softmax partial of (row, this column block)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    const bool row_ok = r < R;          // warp-uniform (a warp owns whole rows)
    const int self = (row_ok && self_col) ? self_col[r] : -1;
    const int pos = row_ok ? pos_col[r] : -1;
    float z[4];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx + 32 * j;
      const int sc = col0 + c;          // column of S
      z[j] = -INFINITY;
      if (row_ok && c < C) {
        S[(long long)r * ld_s + sc] = acc[i][j];
        // __fmul_rn: the finish kernel must see the SAME rounded logit (see contrast_rows_kernel)
        const float zz = __fmul_rn(acc[i][j], inv_T);
        if (sc != self) {
          z[j] = zz;
          mx = fmaxf(mx, zz);
          if (logits) {
            const int idx = (sc == pos) ? 0 : 1 + sc - (self >= 0 && sc > self ? 1 : 0) - (sc > pos ? 1 : 0);
            logits[(long long)r * ld_logits + idx] = zz;
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (z[j] != -INFINITY) se += __expf(z[j] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (row_ok && tx == 0) {
      partials[((long long)r * nblk + blockIdx.x) * 2] = mx;      // -inf when the block held only the self column
      partials[((long long)r * nblk + blockIdx.x) * 2 + 1] = se;
    }
  }
}

// One CTA per row. S row = [extra columns 0..col0-1 | GEMM columns]; Ctot = col0 + C.
__global__ void __launch_bounds__(256)
sim_ce_finish_kernel(float* __restrict__ S, int ld_s, int Ctot, int col0, const float* __restrict__ partials, int nblk,
                     float* __restrict__ logits, int ld_logits, const int* __restrict__ self_col,
                     const int* __restrict__ pos_col, float inv_T, float grad_scale, float* __restrict__ loss_sum,
                     int* __restrict__ hits) {
  const int r = blockIdx.x;
  float* row = S + (long long)r * ld_s;
  const int self = self_col ? self_col[r] : -1;
  const int pos = pos_col[r];
  __shared__ float red[32];
  __shared__ int redi[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float zp = __fmul_rn(row[pos], inv_T);
  // row maximum: the column blocks' maxima and the extra columns
  float mx = -INFINITY;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) mx = fmaxf(mx, partials[((long long)r * nblk + b) * 2]);
  for (int c = threadIdx.x; c < col0; c += blockDim.x)
    if (c != self) mx = fmaxf(mx, __fmul_rn(row[c], inv_T));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < nwarps; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  // sum of exp(z - max): rescaled block sums + the extra columns
  float se = 0.f;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    const float m = partials[((long long)r * nblk + b) * 2];
    if (m != -INFINITY) se += partials[((long long)r * nblk + b) * 2 + 1] * __expf(m - mx);
  }
  for (int c = threadIdx.x; c < col0; c += blockDim.x)
    if (c != self) se += __expf(__fmul_rn(row[c], inv_T) - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  if (lane == 0) red[warp] = se;
  __syncthreads();
  se = 0.f;
  for (int w = 0; w < nwarps; ++w) se += red[w];
  const float inv_se = 1.f / se;
  // the one pass over the row: how many negatives beat the positive, logits of the extra columns, gradient in place
  float* lrow = logits ? logits + (long long)r * ld_logits : nullptr;
  const float gs = grad_scale * inv_T;
  int above = 0;
  for (int c = threadIdx.x; c < Ctot; c += blockDim.x) {
    const float z = __fmul_rn(row[c], inv_T);
    if (c == self) { row[c] = 0.f; continue; }
    if (c != pos && z > zp) ++above;
    if (lrow && c < col0) {
      const int idx = (c == pos) ? 0 : 1 + c - (self >= 0 && c > self ? 1 : 0) - (c > pos ? 1 : 0);
      lrow[idx] = z;
    }
    const float pr = __expf(z - mx) * inv_se;
    row[c] = (pr - (c == pos ? 1.f : 0.f)) * gs;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
  if (lane == 0) redi[warp] = above;
  __syncthreads();
  if (threadIdx.x == 0) {
    above = 0;
    for (int w = 0; w < nwarps; ++w) above += redi[w];
    atomicAdd(loss_sum, (mx - zp) + logf(se));   // not lse - zp: exact when the positive is the row maximum
    if (hits) {
      if (above < 1) atomicAdd(&hits[0], 1);
      if (above < 5) atomicAdd(&hits[1], 1);
    }
  }
}

int sim_ce_blocks(int C) { return ceil_div(C, kSimCols); }

int sim_ce_fwd(const float* a, int lda, const float* b, int ldb, int b_dmajor, int R, int C, int d, float* S, int ld_s,
               int col0, float* logits, int ld_logits, const int* self_col, const int* pos_col, float inv_T,
               float* partials, cudaStream_t stream) {
  const int nblk = sim_ce_blocks(C);
  dim3 grid(nblk, ceil_div(R, kSimRows));
  sim_ce_fwd_kernel<<<grid, 256, 0, stream>>>(a, lda, b, ldb, b_dmajor, R, C, d, S, ld_s, col0, logits, ld_logits,
                                               self_col, pos_col, inv_T, partials, nblk);
  DV_LAUNCH_OK();
  return kOk;
}

int sim_ce_finish(float* S, int ld_s, int R, int C, int col0, const float* partials, float* logits, int ld_logits,
                  const int* self_col, const int* pos_col, float inv_T, float grad_scale, float* loss_sum, int* hits,
                  cudaStream_t stream) {
  sim_ce_finish_kernel<<<R, 256, 0, stream>>>(S, ld_s, col0 + C, col0, partials, sim_ce_blocks(C), logits, ld_logits,
                                              self_col, pos_col, inv_T, grad_scale, loss_sum, hits);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
