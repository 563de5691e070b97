// S3D-G self-gating on bf16 NDHWC branch outputs that live as channel slices of the Inception concat
// tensor (reference: SelfGating, backbone/s3dg.py:68-78 — mean over (T,H,W) -> Linear -> sigmoid ->
// channel scale). The Linear itself is dv_sgemm; these kernels do the reductions and the
// in-place scaling so that the concat (backbone/s3dg.py:130) never needs a copy (K9).
#include <cuda_bf16.h>

#include "host_common.h"

#include <algorithm>

namespace dv {

// 8 bf16 channels of one row as fp32
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// Work split of the two slice reductions: a block owns `chunk` consecutive rows of ONE sample and all channel groups
// (8 channels = one 16-byte load per thread and row); partial sums leave through shared memory and one float atomic
// per (block, channel) - the old one-block-per-(sample, 128 channels) kernels read 2 bytes per thread on < 100 blocks
// and ran at a tenth of the HBM rate (72-83 us per call on 25-50 MB slices).
constexpr int kGateThreads = 256;

// x: rows = N*S rows of `ld` channels; slice [coff, coff+C), C % 8 == 0. out[n][c] += sum_s x[n][s][coff+c] / S (out zeroed)
__global__ void __launch_bounds__(kGateThreads)
slice_mean_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int S, int C, int ld, int coff, int chunk) {
  const int n = blockIdx.y;
  const int groups = C >> 3;
  const int lanes = kGateThreads / groups > 0 ? kGateThreads / groups : 1;     // rows in flight per iteration
  const int s0 = blockIdx.x * chunk, s1 = min(S, s0 + chunk);
  __shared__ float acc[1024];          // [C] (C <= 1024)
  for (int i = threadIdx.x; i < C; i += kGateThreads) acc[i] = 0.f;
  __syncthreads();
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  if (lane < lanes) {                  // (threads beyond lanes * groups idle: 256 need not divide by the group count)
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16* p = x + ((long long)n * S) * ld + coff + g * 8;
    for (int s = s0 + lane; s < s1; s += lanes) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(p + (long long)s * ld)), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&acc[g * 8 + i], a[i]);
  }
  __syncthreads();
  const float inv = 1.f / (float)S;
  for (int i = threadIdx.x; i < C; i += kGateThreads) atomicAdd(&out[(long long)n * C + i], acc[i] * inv);
}

// x[n][s][coff+c] *= w[n][c]  (in place), 8 channels per thread
__global__ void gate_scale_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ w, int S, int C,
                                  int ld, int coff, long long total8) {
  const int groups = C >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long r = i / groups;
    const long long n = r / S;
    uint4* p = reinterpret_cast<uint4*>(x + r * ld + coff + g * 8);
    float v[8];
    unpack8(*p, v);
    const float* wp = w + n * C + g * 8;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k] * wp[2 * k], v[2 * k + 1] * wp[2 * k + 1]);
      o[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *p = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// dw[n][c] += sum_s dout[n][s][coff+c] * z[n][s][c],  z = relu(scale[c]*y + shift[c]) recomputed from y (dw zeroed)
__global__ void __launch_bounds__(kGateThreads)
gate_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ y,
                       const float* __restrict__ ss, float* __restrict__ dw, int S, int C, int Cp, int ld, int coff,
                       int chunk) {
  const int n = blockIdx.y;
  const int groups = C >> 3;
  const int lanes = kGateThreads / groups > 0 ? kGateThreads / groups : 1;
  const int s0 = blockIdx.x * chunk, s1 = min(S, s0 + chunk);
  __shared__ float acc[1024];
  for (int i = threadIdx.x; i < C; i += kGateThreads) acc[i] = 0.f;
  __syncthreads();
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  if (lane < lanes) {
    float sc[8], sh[8], a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = ss[g * 8 + i]; sh[i] = ss[Cp + g * 8 + i]; a[i] = 0.f; }
    const __nv_bfloat16* pd = dout + ((long long)n * S) * ld + coff + g * 8;
    const __nv_bfloat16* py = y + ((long long)n * S) * Cp + g * 8;
    for (int s = s0 + lane; s < s1; s += lanes) {
      float d[8], v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(pd + (long long)s * ld)), d);
      unpack8(__ldg(reinterpret_cast<const uint4*>(py + (long long)s * Cp)), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(d[i], fmaxf(fmaf(v[i], sc[i], sh[i]), 0.f), a[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&acc[g * 8 + i], a[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += kGateThreads) atomicAdd(&dw[(long long)n * C + i], acc[i]);
}

// dz[n][s][c] = w[n][c] * dout[n][s][coff+c] + dmean[n][c] / S   (dense [rows][C], C % 8 == 0), 8 channels per thread
__global__ void gate_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const float* __restrict__ w,
                                      const float* __restrict__ dmean, __nv_bfloat16* __restrict__ dz, int S,
                                      int C, int ld, int coff, long long total8) {
  const float inv = 1.f / (float)S;
  const int groups = C >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long r = i / groups;
    const long long n = r / S;
    float d[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dout + r * ld + coff + g * 8)), d);
    const float* wp = w + n * C + g * 8;
    const float* mp = dmean + n * C + g * 8;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(fmaf(wp[2 * k], d[2 * k], mp[2 * k] * inv),
                                                     fmaf(wp[2 * k + 1], d[2 * k + 1], mp[2 * k + 1] * inv));
      o[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dz + r * C + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = 1.f / (1.f + __expf(-x[i]));
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                   long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = dy[i] * y[i] * (1.f - y[i]);
}

// ---- the gate's fully connected layer, fused with its sigmoid (backbone/s3dg.py:74-77): tiny matrices (N clips x C x C),
// one launch instead of sgemm + sigmoid (forward) and sgemm x2 + sigmoid' + column sum (backward)
// w[n][j] = sigmoid(b[j] + sum_k mean[n][k] * W[j][k]): one warp per (n, j)
__global__ void __launch_bounds__(256)
gate_fc_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ W, const float* __restrict__ b,
                   float* __restrict__ w, int N, int C) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)N * C) return;
  const int n = (int)(wid / C), j = (int)(wid - (long long)n * C);
  const float* m = mean + (long long)n * C;
  const float* wr = W + (long long)j * C;
  float s = 0.f;
  for (int k = lane; k < C; k += 32) s = fmaf(m[k], wr[k], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) w[wid] = 1.f / (1.f + __expf(-(s + b[j])));
}

// block j: dpre[n][j] = dw[n][j] * w[n][j] * (1 - w[n][j]); gb[j] = sum_n dpre[n][j]; gW[j][k] = sum_n dpre[n][j] * mean[n][k]
__global__ void __launch_bounds__(256)
gate_fc_bwd_weight_kernel(const float* __restrict__ dw, const float* __restrict__ w, const float* __restrict__ mean,
                          float* __restrict__ dpre, float* __restrict__ gW, float* __restrict__ gb, int N, int C) {
  const int j = blockIdx.x;
  __shared__ float dp[1024];          // [N] (N <= 1024)
  float part = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float wv = w[(long long)n * C + j];
    const float d = dw[(long long)n * C + j] * wv * (1.f - wv);
    dp[n] = d;
    dpre[(long long)n * C + j] = d;
    part += d;
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    gb[j] = t;
  }
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(dp[n], mean[(long long)n * C + k], s);
    gW[(long long)j * C + k] = s;
  }
}

// block n: dmean[n][k] = sum_j dpre[n][j] * W[j][k]
__global__ void __launch_bounds__(256)
gate_fc_bwd_input_kernel(const float* __restrict__ dpre, const float* __restrict__ W, float* __restrict__ dmean, int C) {
  const int n = blockIdx.x;
  __shared__ float dp[1024];          // [C]
  for (int j = threadIdx.x; j < C; j += blockDim.x) dp[j] = dpre[(long long)n * C + j];
  __syncthreads();
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < C; ++j) s = fmaf(dp[j], W[(long long)j * C + k], s);
    dmean[(long long)n * C + k] = s;
  }
}

static int fgrid(long long total) {
  long long g = ceil_div_ll(total, 256);
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// rows of one sample per block such that the grid has ~4 blocks per SM
static int gate_chunk(int N, int S) {
  const int blocks_per_sample = std::max(1, (4 * sm_count() + N - 1) / N);
  return std::max(8, (S + blocks_per_sample - 1) / blocks_per_sample);
}
static int gate_shape_ok(int C, int ld, int coff) {
  if (C <= 0 || C % 8 != 0 || C > 1024 || ld % 8 != 0 || coff % 8 != 0)
    return fail(kUnsupported, "gating kernels need C %% 8 == 0 (<= 1024), ld %% 8 == 0, coff %% 8 == 0 (C=%d ld=%d coff=%d)", C, ld, coff);
  return kOk;
}
int slice_mean(const void* x, float* out, int N, int S, int C, int ld, int coff, cudaStream_t st) {
  if (int rc = gate_shape_ok(C, ld, coff)) return rc;
  DV_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N * C, st));
  const int chunk = gate_chunk(N, S);
  slice_mean_kernel<<<dim3(ceil_div(S, chunk), N), kGateThreads, 0, st>>>((const __nv_bfloat16*)x, out, S, C, ld, coff, chunk);
  DV_LAUNCH_OK();
  return kOk;
}
int gate_scale(void* x, const float* w, int N, int S, int C, int ld, int coff, cudaStream_t st) {
  if (int rc = gate_shape_ok(C, ld, coff)) return rc;
  const long long total8 = (long long)N * S * (C / 8);
  gate_scale_kernel<<<fgrid(total8), 256, 0, st>>>((__nv_bfloat16*)x, w, S, C, ld, coff, total8);
  DV_LAUNCH_OK();
  return kOk;
}
int gate_bwd_reduce(const void* dout, const void* y, const float* ss, float* dw, int N, int S, int C, int Cp,
                    int ld, int coff, cudaStream_t st) {
  if (int rc = gate_shape_ok(C, ld, coff)) return rc;
  if (Cp != C) return fail(kUnsupported, "gate_bwd_reduce: the branch's raw output must have Cp == C (%d vs %d)", Cp, C);
  DV_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * C, st));
  const int chunk = gate_chunk(N, S);
  gate_bwd_reduce_kernel<<<dim3(ceil_div(S, chunk), N), kGateThreads, 0, st>>>(
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)y, ss, dw, S, C, Cp, ld, coff, chunk);
  DV_LAUNCH_OK();
  return kOk;
}
int gate_bwd_apply(const void* dout, const float* w, const float* dmean, void* dz, int N, int S, int C, int Cp,
                   int ld, int coff, cudaStream_t st) {
  if (int rc = gate_shape_ok(C, ld, coff)) return rc;
  if (Cp != C) return fail(kUnsupported, "gate_bwd_apply: Cp == C expected (%d vs %d)", Cp, C);
  const long long total8 = (long long)N * S * (C / 8);
  gate_bwd_apply_kernel<<<fgrid(total8), 256, 0, st>>>((const __nv_bfloat16*)dout, w, dmean, (__nv_bfloat16*)dz, S,
                                                       C, ld, coff, total8);
  DV_LAUNCH_OK();
  return kOk;
}
int gate_fc_fwd(const float* mean, const float* W, const float* b, float* w, int N, int C, cudaStream_t st) {
  const long long threads = (long long)N * C * 32;
  gate_fc_fwd_kernel<<<(unsigned)ceil_div_ll(threads, 256), 256, 0, st>>>(mean, W, b, w, N, C);
  DV_LAUNCH_OK();
  return kOk;
}
int gate_fc_bwd(const float* dw, const float* w, const float* mean, const float* W, float* dpre, float* gW, float* gb,
                float* dmean, int N, int C, cudaStream_t st) {
  if (N > 1024 || C > 1024) return fail(kUnsupported, "gate_fc_bwd: N, C <= 1024 (N=%d C=%d)", N, C);
  gate_fc_bwd_weight_kernel<<<C, 256, 0, st>>>(dw, w, mean, dpre, gW, gb, N, C);
  DV_LAUNCH_OK();
  gate_fc_bwd_input_kernel<<<N, 256, 0, st>>>(dpre, W, dmean, C);
  DV_LAUNCH_OK();
  return kOk;
}
int sigmoid_fwd(const float* x, float* y, long long n, cudaStream_t st) {
  sigmoid_fwd_kernel<<<fgrid(n), 256, 0, st>>>(x, y, n);
  DV_LAUNCH_OK();
  return kOk;
}
int sigmoid_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t st) {
  sigmoid_bwd_kernel<<<fgrid(n), 256, 0, st>>>(dy, y, dx, n);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
