// S3D-G self-gating on bf16 NDHWC branch outputs that live as channel slices of the Inception concat
// tensor (reference: SelfGating, backbone/s3dg.py:68-78 — mean over (T,H,W) -> Linear -> sigmoid ->
// channel scale). The Linear itself is dv_sgemm; these kernels do the reductions and the
// in-place scaling so that the concat (backbone/s3dg.py:130) never needs a copy (K9).
#include <cuda_bf16.h>

#include "host_common.h"

namespace dv {

// x: rows = N*S rows of `ld` channels; slice [coff, coff+C). out[n][c] = mean_s x[n][s][coff+c]
__global__ void slice_mean_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int S, int C,
                                  int ld, int coff) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float part[8][128];
  float s = 0.f;
  if (c < C) {
    const __nv_bfloat16* p = x + (long long)n * S * ld + coff + c;
    for (int i = threadIdx.y; i < S; i += blockDim.y) s += __bfloat162float(p[(long long)i * ld]);
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int j = 0; j < (int)blockDim.y; ++j) t += part[j][threadIdx.x];
    out[(long long)n * C + c] = t / (float)S;
  }
}

// x[n][s][coff+c] *= w[n][c]  (in place)
__global__ void gate_scale_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ w, int S, int C,
                                  int ld, int coff, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const long long n = r / S;
    __nv_bfloat16* p = x + r * ld + coff + c;
    *p = __float2bfloat16_rn(__bfloat162float(*p) * w[n * C + c]);
  }
}

// dw[n][c] += sum_s dout[n][s][coff+c] * z[n][s][c],  z = relu(scale[c]*y + shift[c]) recomputed from y
__global__ void gate_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ y,
                                       const float* __restrict__ ss, float* __restrict__ dw, int S, int C,
                                       int Cp, int ld, int coff) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float part[8][128];
  float s = 0.f;
  if (c < C) {
    const float sc = ss[c], sh = ss[Cp + c];
    const __nv_bfloat16* pd = dout + (long long)n * S * ld + coff + c;
    const __nv_bfloat16* py = y + (long long)n * S * Cp + c;
    for (int i = threadIdx.y; i < S; i += blockDim.y) {
      const float z = fmaxf(fmaf(__bfloat162float(py[(long long)i * Cp]), sc, sh), 0.f);
      s = fmaf(__bfloat162float(pd[(long long)i * ld]), z, s);
    }
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int j = 0; j < (int)blockDim.y; ++j) t += part[j][threadIdx.x];
    dw[(long long)n * C + c] = t;
  }
}

// dz[n][s][c] = w[n][c] * dout[n][s][coff+c] + dmean[n][c] / S   (dense [rows][Cp], pad channels zero)
__global__ void gate_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const float* __restrict__ w,
                                      const float* __restrict__ dmean, __nv_bfloat16* __restrict__ dz, int S,
                                      int C, int Cp, int ld, int coff, long long total) {
  const float inv = 1.f / (float)S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const long long r = i / Cp;
    const long long n = r / S;
    float v = 0.f;
    if (c < C) v = w[n * C + c] * __bfloat162float(dout[r * ld + coff + c]) + dmean[n * C + c] * inv;
    dz[i] = __float2bfloat16_rn(v);
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

static int fgrid(long long total) {
  long long g = ceil_div_ll(total, 256);
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int slice_mean(const void* x, float* out, int N, int S, int C, int ld, int coff, cudaStream_t st) {
  slice_mean_kernel<<<dim3(ceil_div(C, 128), N), dim3(128, 8), 0, st>>>((const __nv_bfloat16*)x, out, S, C, ld, coff);
  DV_LAUNCH_OK();
  return kOk;
}
int gate_scale(void* x, const float* w, int N, int S, int C, int ld, int coff, cudaStream_t st) {
  const long long total = (long long)N * S * C;
  gate_scale_kernel<<<fgrid(total), 256, 0, st>>>((__nv_bfloat16*)x, w, S, C, ld, coff, total);
  DV_LAUNCH_OK();
  return kOk;
}
int gate_bwd_reduce(const void* dout, const void* y, const float* ss, float* dw, int N, int S, int C, int Cp,
                    int ld, int coff, cudaStream_t st) {
  gate_bwd_reduce_kernel<<<dim3(ceil_div(C, 128), N), dim3(128, 8), 0, st>>>(
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)y, ss, dw, S, C, Cp, ld, coff);
  DV_LAUNCH_OK();
  return kOk;
}
int gate_bwd_apply(const void* dout, const float* w, const float* dmean, void* dz, int N, int S, int C, int Cp,
                   int ld, int coff, cudaStream_t st) {
  const long long total = (long long)N * S * Cp;
  gate_bwd_apply_kernel<<<fgrid(total), 256, 0, st>>>((const __nv_bfloat16*)dout, w, dmean, (__nv_bfloat16*)dz, S,
                                                      C, Cp, ld, coff, total);
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
