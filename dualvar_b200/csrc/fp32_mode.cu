// fp32 mode of the clip-encoder path ("1e-4 mode" of the parity bar; the bf16 mode is the throughput mode).
//
// Activations and their gradients are fp32 NDHWC [N][T][H][W][Cp]. The convolutions still run on the tcgen05
// bf16 kernels (conv_fprop.cu / conv_wgrad.cu): every conv operand is also kept as K bf16 "split planes"
//     x = p0 + p1 + ... + p(K-1) (+ O(2^-8K) |x|),  p_k = bf16_rn(x - p0 - ... - p(k-1))   (exact residuals)
// and a convolution is the sum of the plane products x_i * w_j with i + j < K (K = 2: 3 products, 16 mantissa bits;
// K = 3: 6 products, 24 bits = fp32), each one launch of the bf16 kernel that ADDS its fp32 accumulator tile to the
// fp32 output (ConvTileParams::out_f32) or to the packed fp32 weight gradient.
// This file holds the memory-bound fp32 kernels around them; the elementwise ones write the fp32 tensor and its split
// planes in one pass. One thread owns 4 channels (16-byte fp32 accesses, 8-byte plane stores).
//
// Reference ops replaced: nn.BatchNorm3d (+ReLU, + residual add) forward/backward, nn.MaxPool3d,
// nn.AdaptiveAvgPool3d((1,1,1)) and the autograd adds between them - backbone/r21d.py:56-57,106-122,
// backbone/r3d.py:74-89, backbone/c3d.py:16-46, backbone/resnet_2d3d.py, model/simclr.py:166.
#include <cuda_bf16.h>

#include <algorithm>

#include "host_common.h"

namespace dv {

struct PoolGeom {
  int N, T, H, W, To, Ho, Wo, Cp;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
};

namespace {

constexpr int kMaxPlanes = 3;

int flat_grid32(long long total, int threads) {
  long long g = ceil_div_ll(total, threads);
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// K split planes of 4 consecutive values; plane k lives at planes + k * plane_stride (elements)
__device__ __forceinline__ void split_store4(float4 v, __nv_bfloat16* planes, long long plane_stride, int K,
                                             long long idx) {
  for (int k = 0; k < K; ++k) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a);
    o.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(planes + k * plane_stride + idx) = o;
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    v.x -= fa.x; v.y -= fa.y; v.z -= fb.x; v.w -= fb.y;   // exact: the residual has <= 16 significant bits left
  }
}

// fp32 values -> K fp32 planes holding the bf16-representable parts (the weight pack kernels then round exactly)
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           long long n, int K) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = src[i];
    for (int k = 0; k < K; ++k) {
      const float h = __bfloat162float(__float2bfloat16_rn(v));
      dst[k * n + i] = h;
      v -= h;
    }
  }
}

// ------------------------------------------------------------------ per-channel reductions over rows
// block (32, 8): threadIdx.x = group of 4 channels inside a 128-channel tile, threadIdx.y = row lane.
// kMode 0: sum(y), sum(y*y)   (BatchNorm batch statistics)
// kMode 1: sum(g), sum(g*y), g = (dout [+ dout2]) * relu mask   (BatchNorm backward)
struct RedArgs {
  const float* y;
  const float* dout;
  const float* dout2;
  const float* out;   // mask source 2: out > 0
  const float* ss;    // mask source 1: scale*y + shift > 0
  double* sums;
  long long rows;
  int Cp, relu;
  int o_ld, o_coff;   // dout / out are channel slices [o_coff, o_coff + Cp) of rows with o_ld channels (concat tensors)
};

template <int kMode>
__global__ void __launch_bounds__(256) f32_colreduce_kernel(const RedArgs a) {
  __shared__ double part[8][32][8];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const bool live = c < a.Cp;
  double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
  if (live) {
    float4 fs = make_float4(0.f, 0.f, 0.f, 0.f), fb = fs;
    const bool mask_ss = kMode == 1 && a.relu && a.ss != nullptr;
    if (mask_ss) { fs = ld4(a.ss + c); fb = ld4(a.ss + a.Cp + c); }
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < a.rows; r += (long long)gridDim.y * 8) {
      const long long off = r * a.Cp + c;
      const long long ooff = r * a.o_ld + a.o_coff + c;
      const float4 y = ld4(a.y + off);
      if (kMode == 0) {
        s[0] += y.x; s[1] += y.y; s[2] += y.z; s[3] += y.w;
        q[0] += (double)y.x * y.x; q[1] += (double)y.y * y.y; q[2] += (double)y.z * y.z; q[3] += (double)y.w * y.w;
      } else {
        float4 g = ld4(a.dout + ooff);
        if (a.dout2 != nullptr) {
          const float4 e = ld4(a.dout2 + off);
          g.x += e.x; g.y += e.y; g.z += e.z; g.w += e.w;
        }
        if (a.relu) {
          if (mask_ss) {
            g.x = fmaf(y.x, fs.x, fb.x) > 0.f ? g.x : 0.f;
            g.y = fmaf(y.y, fs.y, fb.y) > 0.f ? g.y : 0.f;
            g.z = fmaf(y.z, fs.z, fb.z) > 0.f ? g.z : 0.f;
            g.w = fmaf(y.w, fs.w, fb.w) > 0.f ? g.w : 0.f;
          } else {
            const float4 o = ld4(a.out + ooff);
            g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
            g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
          }
        }
        s[0] += g.x; s[1] += g.y; s[2] += g.z; s[3] += g.w;
        q[0] += (double)g.x * y.x; q[1] += (double)g.y * y.y; q[2] += (double)g.z * y.z; q[3] += (double)g.w * y.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    part[threadIdx.y][threadIdx.x][j] = s[j];
    part[threadIdx.y][threadIdx.x][4 + j] = q[j];
  }
  __syncthreads();
  // 256 threads = 32 channel groups x 8 values
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int gx = tid >> 3, j = tid & 7;
  const int cc = (blockIdx.x * 32 + gx) * 4 + (j & 3);
  if (cc < a.Cp) {
    double t = 0.0;
#pragma unroll
    for (int ry = 0; ry < 8; ++ry) t += part[ry][gx][j];
    atomicAdd(&a.sums[(j < 4 ? 0 : a.Cp) + cc], t);
  }
}

int launch_colreduce(int mode, const RedArgs& a, cudaStream_t stream) {
  const int G = a.Cp / 4;
  dim3 block(32, 8);
  const int gx = ceil_div(G, 32);
  long long gy = ceil_div_ll(a.rows, 8 * 4);   // >= 4 rows per thread
  const long long cap = std::max(1, sm_count() * 8 / gx);
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  dim3 grid(gx, (unsigned)gy);
  if (mode == 0) f32_colreduce_kernel<0><<<grid, block, 0, stream>>>(a);
  else f32_colreduce_kernel<1><<<grid, block, 0, stream>>>(a);
  DV_LAUNCH_OK();
  return kOk;
}

// ------------------------------------------------------------------ BatchNorm apply (+ second branch, residual, ReLU)
__global__ void __launch_bounds__(256) f32_bn_apply_kernel(const float* __restrict__ y1, const float* __restrict__ ss1,
                                                           const float* __restrict__ y2, const float* __restrict__ ss2,
                                                           const float* __restrict__ res, float* __restrict__ out,
                                                           __nv_bfloat16* __restrict__ planes, long long plane_stride,
                                                           int K, long long total4, int Cp, int out_ld, int out_coff,
                                                           int relu) {
  const int G = Cp >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % G) * 4;
    const long long o = (i / G) * out_ld + out_coff + c;   // destination: a channel slice of a (concat) tensor
    const float4 y = ld4(y1 + i * 4), s = ld4(ss1 + c), b = ld4(ss1 + Cp + c);
    float4 v = make_float4(fmaf(y.x, s.x, b.x), fmaf(y.y, s.y, b.y), fmaf(y.z, s.z, b.z), fmaf(y.w, s.w, b.w));
    if (y2 != nullptr) {
      const float4 yy = ld4(y2 + i * 4), s2 = ld4(ss2 + c), b2 = ld4(ss2 + Cp + c);
      v.x += fmaf(yy.x, s2.x, b2.x); v.y += fmaf(yy.y, s2.y, b2.y);
      v.z += fmaf(yy.z, s2.z, b2.z); v.w += fmaf(yy.w, s2.w, b2.w);
    }
    if (res != nullptr) {
      const float4 r = ld4(res + i * 4);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    st4(out + o, v);
    if (planes != nullptr) split_store4(v, planes, plane_stride, K, o);
  }
}

// dy = A*g + B*y + C with g = (dout [+ dout2]) * relu mask; dy goes out as split planes (it only feeds dgrad / wgrad),
// g optionally as fp32 (gradient of the residual branch)
__global__ void __launch_bounds__(256) f32_bn_bwd_apply_kernel(const RedArgs a, const float* __restrict__ coef,
                                                               __nv_bfloat16* __restrict__ dy_planes,
                                                               long long plane_stride, int K, float* __restrict__ g_out,
                                                               long long total4) {
  const int G = a.Cp >> 2;
  const bool mask_ss = a.relu && a.ss != nullptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % G) * 4;
    const float4 y = ld4(a.y + i * 4);
    const long long ooff = (i / G) * a.o_ld + a.o_coff + c;
    float4 g = ld4(a.dout + ooff);
    if (a.dout2 != nullptr) {
      const float4 e = ld4(a.dout2 + i * 4);
      g.x += e.x; g.y += e.y; g.z += e.z; g.w += e.w;
    }
    if (a.relu) {
      if (mask_ss) {
        const float4 fs = ld4(a.ss + c), fb = ld4(a.ss + a.Cp + c);
        g.x = fmaf(y.x, fs.x, fb.x) > 0.f ? g.x : 0.f;
        g.y = fmaf(y.y, fs.y, fb.y) > 0.f ? g.y : 0.f;
        g.z = fmaf(y.z, fs.z, fb.z) > 0.f ? g.z : 0.f;
        g.w = fmaf(y.w, fs.w, fb.w) > 0.f ? g.w : 0.f;
      } else {
        const float4 o = ld4(a.out + ooff);
        g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
        g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
      }
    }
    if (g_out != nullptr) st4(g_out + i * 4, g);
    const float4 A = ld4(coef + c), B = ld4(coef + a.Cp + c), C = ld4(coef + 2 * a.Cp + c);
    const float4 d = make_float4(fmaf(A.x, g.x, fmaf(B.x, y.x, C.x)), fmaf(A.y, g.y, fmaf(B.y, y.y, C.y)),
                                 fmaf(A.z, g.z, fmaf(B.z, y.z, C.z)), fmaf(A.w, g.w, fmaf(B.w, y.w, C.w)));
    split_store4(d, dy_planes, plane_stride, K, i * 4);
  }
}

__global__ void __launch_bounds__(256) f32_add_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      float* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = ld4(x + i * 4), b = ld4(y + i * 4);
    st4(out + i * 4, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
  }
}

__global__ void __launch_bounds__(256) f32_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ planes,
                                                        long long plane_stride, int K, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    split_store4(ld4(x + i * 4), planes, plane_stride, K, i * 4);
}

// ------------------------------------------------------------------ S3D-G self-gating on fp32 concat slices
// (backbone/s3dg.py:68-78; the bf16 versions are in gating.cu). x: rows of `ld` channels, slice [coff, coff + C).
__global__ void f32_slice_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int S, int C, int ld,
                                      int coff) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float part[8][128];
  float s = 0.f;
  if (c < C) {
    const float* p = x + (long long)n * S * ld + coff + c;
    for (int i = threadIdx.y; i < S; i += blockDim.y) s += p[(long long)i * ld];
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int j = 0; j < (int)blockDim.y; ++j) t += part[j][threadIdx.x];
    out[(long long)n * C + c] = t / (float)S;
  }
}

// x[n][s][coff+c] *= w[n][c] in place, split planes of the slice rewritten
__global__ void __launch_bounds__(256) f32_gate_scale_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ planes,
                                                             long long plane_stride, int K, const float* __restrict__ w,
                                                             int S, int C, int ld, int coff, long long total4) {
  const int G = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % G) * 4;
    const long long r = i / G;
    const long long n = r / S;
    const long long o = r * ld + coff + c;
    float4 v = ld4(x + o);
    const float4 g = ld4(w + n * C + c);
    v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    st4(x + o, v);
    split_store4(v, planes, plane_stride, K, o);
  }
}

// dw[n][c] = sum_s dout[n][s][coff+c] * z[n][s][c],  z = relu(scale[c]*y + shift[c]) recomputed from y
__global__ void f32_gate_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                           const float* __restrict__ ss, float* __restrict__ dw, int S, int C, int Cp,
                                           int ld, int coff) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float part[8][128];
  float s = 0.f;
  if (c < C) {
    const float sc = ss[c], sh = ss[Cp + c];
    const float* pd = dout + (long long)n * S * ld + coff + c;
    const float* py = y + (long long)n * S * Cp + c;
    for (int i = threadIdx.y; i < S; i += blockDim.y) {
      const float z = fmaxf(fmaf(py[(long long)i * Cp], sc, sh), 0.f);
      s = fmaf(pd[(long long)i * ld], z, s);
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
__global__ void __launch_bounds__(256) f32_gate_bwd_apply_kernel(const float* __restrict__ dout,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ dmean, float* __restrict__ dz,
                                                                 int S, int C, int Cp, int ld, int coff, long long total) {
  const float inv = 1.f / (float)S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const long long r = i / Cp;
    const long long n = r / S;
    dz[i] = c < C ? w[n * C + c] * dout[r * ld + coff + c] + dmean[n * C + c] * inv : 0.f;
  }
}

// ------------------------------------------------------------------ pooling
// x: [N][S][Cp] fp32 -> out: [N][ld_out] fp32 (first C entries), mean over S
__global__ void f32_avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int S, int C, int Cp,
                                       int ld_out) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float part[8][128];
  float s = 0.f;
  if (c < Cp) {
    const float* p = x + (long long)n * S * Cp + c;
    for (int i = threadIdx.y; i < S; i += blockDim.y) s += p[(long long)i * Cp];
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int j = 0; j < (int)blockDim.y; ++j) t += part[j][threadIdx.x];
    out[(long long)n * ld_out + c] = t / (float)S;
  }
}

__global__ void __launch_bounds__(256) f32_avgpool_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx,
                                                              int S, int C, int Cp, int ld_out, long long total) {
  const float inv = 1.f / (float)S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const long long n = i / ((long long)S * Cp);
    dx[i] = (c < C) ? dout[n * ld_out + c] * inv : 0.f;
  }
}

// MaxPool3d, one thread = one output position x 4 channels; argmax byte = window offset of the FIRST maximum
// (ATen's tie rule, as in pool.cu)
__global__ void __launch_bounds__(256) f32_maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              uint8_t* __restrict__ idx,
                                                              __nv_bfloat16* __restrict__ planes, long long plane_stride,
                                                              int K, const PoolGeom g, long long total) {
  const int G = g.Cp >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long r = i / G;
    const int wo = (int)(r % g.Wo); r /= g.Wo;
    const int ho = (int)(r % g.Ho); r /= g.Ho;
    const int to = (int)(r % g.To);
    const long long n = r / g.To;
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    uint32_t am[4] = {0u, 0u, 0u, 0u};
    for (int a = 0; a < g.kt; ++a) {
      const int t = to * g.st - g.pt + a;
      if (t < 0 || t >= g.T) continue;
      for (int b = 0; b < g.kh; ++b) {
        const int h = ho * g.sh - g.ph + b;
        if (h < 0 || h >= g.H) continue;
        for (int c = 0; c < g.kw; ++c) {
          const int w = wo * g.sw - g.pw + c;
          if (w < 0 || w >= g.W) continue;
          const float4 f = ld4(x + ((((n * g.T + t) * g.H + h) * g.W + w) * (long long)g.Cp + cg * 4));
          const uint32_t code = (uint32_t)((a * g.kh + b) * g.kw + c);
          if (f.x > m[0]) { m[0] = f.x; am[0] = code; }
          if (f.y > m[1]) { m[1] = f.y; am[1] = code; }
          if (f.z > m[2]) { m[2] = f.z; am[2] = code; }
          if (f.w > m[3]) { m[3] = f.w; am[3] = code; }
        }
      }
    }
    const float4 v = make_float4(m[0], m[1], m[2], m[3]);
    st4(y + i * 4, v);
    if (planes != nullptr) split_store4(v, planes, plane_stride, K, i * 4);
    if (idx != nullptr)
      *reinterpret_cast<uint32_t*>(idx + i * 4) = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
  }
}

// gather backward: one thread = one INPUT position x 4 channels
__global__ void __launch_bounds__(256) f32_maxpool_bwd_kernel(const uint8_t* __restrict__ idx,
                                                              const float* __restrict__ dy, float* __restrict__ dx,
                                                              const PoolGeom g, long long total) {
  const int G = g.Cp >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long r = i / G;
    const int w = (int)(r % g.W); r /= g.W;
    const int h = (int)(r % g.H); r /= g.H;
    const int t = (int)(r % g.T);
    const long long n = r / g.T;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int to_lo = max(0, (t + g.pt - g.kt + g.st) / g.st), to_hi = min(g.To - 1, (t + g.pt) / g.st);
    const int ho_lo = max(0, (h + g.ph - g.kh + g.sh) / g.sh), ho_hi = min(g.Ho - 1, (h + g.ph) / g.sh);
    const int wo_lo = max(0, (w + g.pw - g.kw + g.sw) / g.sw), wo_hi = min(g.Wo - 1, (w + g.pw) / g.sw);
    for (int to = to_lo; to <= to_hi; ++to)
      for (int ho = ho_lo; ho <= ho_hi; ++ho)
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
          const uint32_t code = (uint32_t)(((t - (to * g.st - g.pt)) * g.kh + (h - (ho * g.sh - g.ph))) * g.kw +
                                           (w - (wo * g.sw - g.pw)));
          const long long oo = ((((n * g.To + to) * g.Ho + ho) * g.Wo + wo) * (long long)g.Cp + cg * 4);
          const uint32_t k = *reinterpret_cast<const uint32_t*>(idx + oo);
          const float4 d = ld4(dy + oo);
          if ((k & 0xffu) == code) acc[0] += d.x;
          if (((k >> 8) & 0xffu) == code) acc[1] += d.y;
          if (((k >> 16) & 0xffu) == code) acc[2] += d.z;
          if ((k >> 24) == code) acc[3] += d.w;
        }
    st4(dx + i * 4, make_float4(acc[0], acc[1], acc[2], acc[3]));
  }
}

// ------------------------------------------------------------------ module-boundary layouts
// y: fp32 [N][S][Cp] -> x: fp32 [N][C][S]
__global__ void __launch_bounds__(256) f32_ndhwc_to_ncdhw_kernel(const float* __restrict__ y, float* __restrict__ x,
                                                                 int C, int Cp, long long S, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = i % S;
    const int c = (int)((i / S) % C);
    const long long n = i / (S * C);
    x[i] = y[(n * S + s) * Cp + c];
  }
}

// x: fp32 [N][C][S] -> y: fp32 [N][S][Cp] (pad channels zero)
__global__ void __launch_bounds__(256) f32_ncdhw_to_ndhwc_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                 int C, int Cp, long long S, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const long long s = (i / Cp) % S;
    const long long n = i / (S * Cp);
    y[i] = c < C ? x[(n * C + c) * S + s] : 0.f;
  }
}

}  // namespace

// ------------------------------------------------------------------ host entry points (called from api.cu)
int f32_split_planes(const float* src, float* dst, long long n, int K, cudaStream_t stream) {
  split_planes_kernel<<<flat_grid32(n, 256), 256, 0, stream>>>(src, dst, n, K);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_colstats(const float* y, double* stats, long long rows, int Cp, cudaStream_t stream) {
  RedArgs a = {};
  a.y = y; a.sums = stats; a.rows = rows; a.Cp = Cp; a.o_ld = Cp;
  return launch_colreduce(0, a, stream);
}

int f32_bn_apply(const float* y1, const float* ss1, const float* y2, const float* ss2, const float* res, float* out,
                 void* planes, long long plane_stride, int K, long long rows, int Cp, int out_ld, int out_coff, int relu,
                 cudaStream_t stream) {
  const long long total4 = rows * (Cp / 4);
  f32_bn_apply_kernel<<<flat_grid32(total4, 256), 256, 0, stream>>>(y1, ss1, y2, ss2, res, out, (__nv_bfloat16*)planes,
                                                                    plane_stride, K, total4, Cp, out_ld, out_coff, relu);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_bn_bwd_reduce(const float* dout, const float* dout2, const float* out, const float* y, const float* ss,
                      double* sums, long long rows, int Cp, int o_ld, int o_coff, int relu, cudaStream_t stream) {
  RedArgs a = {};
  a.y = y; a.dout = dout; a.dout2 = dout2; a.out = out; a.ss = ss; a.sums = sums; a.rows = rows; a.Cp = Cp;
  a.relu = relu; a.o_ld = o_ld; a.o_coff = o_coff;
  return launch_colreduce(1, a, stream);
}

int f32_bn_bwd_apply(const float* dout, const float* dout2, const float* out, const float* y, const float* ss,
                     const float* coef, void* dy_planes, long long plane_stride, int K, float* g_out, long long rows,
                     int Cp, int o_ld, int o_coff, int relu, cudaStream_t stream) {
  RedArgs a = {};
  a.y = y; a.dout = dout; a.dout2 = dout2; a.out = out; a.ss = ss; a.rows = rows; a.Cp = Cp; a.relu = relu;
  a.o_ld = o_ld; a.o_coff = o_coff;
  const long long total4 = rows * (Cp / 4);
  f32_bn_bwd_apply_kernel<<<flat_grid32(total4, 256), 256, 0, stream>>>(a, coef, (__nv_bfloat16*)dy_planes,
                                                                        plane_stride, K, g_out, total4);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_add(const float* x, const float* y, float* out, long long n, cudaStream_t stream) {
  f32_add_kernel<<<flat_grid32(n / 4, 256), 256, 0, stream>>>(x, y, out, n / 4);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_split(const float* x, void* planes, long long plane_stride, int K, long long n, cudaStream_t stream) {
  f32_split_kernel<<<flat_grid32(n / 4, 256), 256, 0, stream>>>(x, (__nv_bfloat16*)planes, plane_stride, K, n / 4);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_avgpool_fwd(const float* x, float* out, int N, int S, int C, int Cp, int ld_out, cudaStream_t stream) {
  dim3 block(128, 8);
  dim3 grid(ceil_div(Cp, 128), N);
  f32_avgpool_fwd_kernel<<<grid, block, 0, stream>>>(x, out, S, C, Cp, ld_out);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_avgpool_bwd(const float* dout, float* dx, int N, int S, int C, int Cp, int ld_out, cudaStream_t stream) {
  const long long total = (long long)N * S * Cp;
  f32_avgpool_bwd_kernel<<<flat_grid32(total, 256), 256, 0, stream>>>(dout, dx, S, C, Cp, ld_out, total);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_maxpool_fwd(const float* x, float* y, uint8_t* idx, void* planes, long long plane_stride, int K,
                    const PoolGeom& g, cudaStream_t stream) {
  if (idx != nullptr && g.kt * g.kh * g.kw > 255) return fail(kUnsupported, "max-pool window too large for 1-byte argmax");
  const long long total = (long long)g.N * g.To * g.Ho * g.Wo * (g.Cp / 4);
  f32_maxpool_fwd_kernel<<<flat_grid32(total, 256), 256, 0, stream>>>(x, y, idx, (__nv_bfloat16*)planes, plane_stride,
                                                                      K, g, total);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_maxpool_bwd(const uint8_t* idx, const float* dy, float* dx, const PoolGeom& g, cudaStream_t stream) {
  const long long total = (long long)g.N * g.T * g.H * g.W * (g.Cp / 4);
  f32_maxpool_bwd_kernel<<<flat_grid32(total, 256), 256, 0, stream>>>(idx, dy, dx, g, total);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_ndhwc_to_ncdhw(const float* y, float* x, int N, int C, int Cp, long long S, cudaStream_t stream) {
  const long long total = (long long)N * C * S;
  f32_ndhwc_to_ncdhw_kernel<<<flat_grid32(total, 256), 256, 0, stream>>>(y, x, C, Cp, S, total);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_ncdhw_to_ndhwc(const float* x, float* y, int N, int C, int Cp, long long S, cudaStream_t stream) {
  const long long total = (long long)N * S * Cp;
  f32_ncdhw_to_ndhwc_kernel<<<flat_grid32(total, 256), 256, 0, stream>>>(x, y, C, Cp, S, total);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_slice_mean(const float* x, float* out, int N, int S, int C, int ld, int coff, cudaStream_t stream) {
  f32_slice_mean_kernel<<<dim3(ceil_div(C, 128), N), dim3(128, 8), 0, stream>>>(x, out, S, C, ld, coff);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_gate_scale(float* x, void* planes, long long plane_stride, int K, const float* w, int N, int S, int C, int ld,
                   int coff, cudaStream_t stream) {
  const long long total4 = (long long)N * S * (C / 4);
  f32_gate_scale_kernel<<<flat_grid32(total4, 256), 256, 0, stream>>>(x, (__nv_bfloat16*)planes, plane_stride, K, w, S,
                                                                      C, ld, coff, total4);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_gate_bwd_reduce(const float* dout, const float* y, const float* ss, float* dw, int N, int S, int C, int Cp,
                        int ld, int coff, cudaStream_t stream) {
  f32_gate_bwd_reduce_kernel<<<dim3(ceil_div(C, 128), N), dim3(128, 8), 0, stream>>>(dout, y, ss, dw, S, C, Cp, ld,
                                                                                      coff);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_gate_bwd_apply(const float* dout, const float* w, const float* dmean, float* dz, int N, int S, int C, int Cp,
                       int ld, int coff, cudaStream_t stream) {
  const long long total = (long long)N * S * Cp;
  f32_gate_bwd_apply_kernel<<<flat_grid32(total, 256), 256, 0, stream>>>(dout, w, dmean, dz, S, C, Cp, ld, coff, total);
  DV_LAUNCH_OK();
  return kOk;
}

int f32_max_planes() { return kMaxPlanes; }

}  // namespace dv
