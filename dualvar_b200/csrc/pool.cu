// Pooling on bf16 NDHWC activations.
//  - global average pool to fp32 features and its backward
//    (reference: nn.AdaptiveAvgPool3d((1,1,1)) in model/simclr.py:166, model/moco.py:281,
//     F.adaptive_avg_pool3d in model/classifier.py:66)
//  - MaxPool3d forward/backward with arbitrary window/stride/padding
//    (reference: backbone/c3d.py:18-39, backbone/s3dg.py:105,151,162,173,190)
//  - ingest: fp32 NCDHW clip views -> bf16 NDHWC (3 -> 8 channels) with the optional per-channel
//    Normalize and the DualVar segment shuffle folded into the address computation
//    (reference: pretrain.py:386-389 + utils/transforms.py:57-63; model/simclr.py:378-383; K16, K22)
#include <cuda_bf16.h>

#include "host_common.h"

namespace dv {

// x: [N][S][Cp] bf16 -> out: [N][ld_out] fp32 (first C entries), mean over S.
__global__ void avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                   int S, int C, int Cp, int ld_out) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;  // channel; blockDim.y splits S
  __shared__ float part[8][128];
  float s = 0.f;
  if (c < Cp) {
    const __nv_bfloat16* p = x + (long long)n * S * Cp + c;
    for (int i = threadIdx.y; i < S; i += blockDim.y) s += __bfloat162float(p[(long long)i * Cp]);
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int j = 0; j < (int)blockDim.y; ++j) t += part[j][threadIdx.x];
    out[(long long)n * ld_out + c] = t / (float)S;
  }
}

// dx[n][s][c] = dout[n][c] / S (pad channels zero)
__global__ void avgpool_bwd_kernel(const float* __restrict__ dout, __nv_bfloat16* __restrict__ dx,
                                   int S, int C, int Cp, int ld_out, long long total) {
  const float inv = 1.f / (float)S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const long long n = i / ((long long)S * Cp);
    const float v = (c < C) ? dout[n * ld_out + c] * inv : 0.f;
    dx[i] = __float2bfloat16_rn(v);
  }
}

struct PoolGeom {
  int N, T, H, W, To, Ho, Wo, Cp;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
};

// One thread = one output position x 8 channels. Optionally records the argmax as the window offset
// (a*kh + b)*kw + c of the FIRST maximum in (t,h,w) scan order (the element ATen routes the gradient to), one byte
// per element, which turns backward into a plain gather.
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   uint8_t* __restrict__ idx, const PoolGeom g, long long total) {
  const int G = g.Cp >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long r = i / G;
    const int wo = (int)(r % g.Wo); r /= g.Wo;
    const int ho = (int)(r % g.Ho); r /= g.Ho;
    const int to = (int)(r % g.To);
    const long long n = r / g.To;
    float m[8];
    uint32_t am[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; am[j] = 0u; }
    for (int a = 0; a < g.kt; ++a) {
      const int t = to * g.st - g.pt + a;
      if (t < 0 || t >= g.T) continue;
      for (int b = 0; b < g.kh; ++b) {
        const int h = ho * g.sh - g.ph + b;
        if (h < 0 || h >= g.H) continue;
        for (int c = 0; c < g.kw; ++c) {
          const int w = wo * g.sw - g.pw + c;
          if (w < 0 || w >= g.W) continue;
          const uint4 u = *reinterpret_cast<const uint4*>(
              x + ((((n * g.T + t) * g.H + h) * g.W + w) * (long long)g.Cp + cg * 8));
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
          const uint32_t code = (uint32_t)((a * g.kh + b) * g.kw + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(hh[j]);
            if (f.x > m[2 * j]) { m[2 * j] = f.x; am[2 * j] = code; }              // strict >: first maximum wins
            if (f.y > m[2 * j + 1]) { m[2 * j + 1] = f.y; am[2 * j + 1] = code; }
          }
        }
      }
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(m[2 * j], m[2 * j + 1]);
    *reinterpret_cast<uint4*>(y + i * 8) = o;
    if (idx != nullptr) {
      uint2 k;
      k.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
      k.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
      *reinterpret_cast<uint2*>(idx + i * 8) = k;
    }
  }
}

// The same pass for a compile-time window (the 3x3x3 / 1x3x3 / 2x2x2 pools of S3D-G, backbone/s3dg.py:105,151-170): the
// window loops unroll, and maximum + argmax are kept PACKED - two bf16 values / two 16-bit window offsets per register,
// updated with one mask compare (__hgt2_mask: strict >, so the first maximum wins as above) and two bit selects per
// channel pair instead of two compares and four selects per channel. The general kernel was instruction-bound: 103 us
// per call on 64 clips x 16^3 x 192 channels, 20 % of an S3D-G step together with its backward.
template <int KT, int KH, int KW>
__global__ void __launch_bounds__(256)
maxpool_fwd_win_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx,
                       const PoolGeom g, long long total) {
  const int G = g.Cp >> 3;
  const long long sW = g.Cp, sH = (long long)g.W * g.Cp, sT = (long long)g.H * sH;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long r = i / G;
    const int wo = (int)(r % g.Wo); r /= g.Wo;
    const int ho = (int)(r % g.Ho); r /= g.Ho;
    const int to = (int)(r % g.To);
    const long long n = r / g.To;
    const int t0 = to * g.st - g.pt, h0 = ho * g.sh - g.ph, w0 = wo * g.sw - g.pw;
    const __nv_bfloat16* base = x + n * g.T * sT + cg * 8;
    uint32_t m[4], am[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { m[j] = 0xff80ff80u; am[j] = 0u; }      // (-inf, -inf), offset 0
#pragma unroll
    for (int a = 0; a < KT; ++a) {
      const int t = t0 + a;
      if (t < 0 || t >= g.T) continue;
#pragma unroll
      for (int b = 0; b < KH; ++b) {
        const int h = h0 + b;
        if (h < 0 || h >= g.H) continue;
#pragma unroll
        for (int c = 0; c < KW; ++c) {
          const int w = w0 + c;
          if (w < 0 || w >= g.W) continue;
          const uint4 u = *reinterpret_cast<const uint4*>(base + t * sT + h * sH + w * sW);
          const uint32_t v[4] = {u.x, u.y, u.z, u.w};
          const uint32_t code = (uint32_t)((a * KH + b) * KW + c) * 0x00010001u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t gt = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&v[j]),
                                            *reinterpret_cast<const __nv_bfloat162*>(&m[j]));
            m[j] = (v[j] & gt) | (m[j] & ~gt);
            am[j] = (code & gt) | (am[j] & ~gt);
          }
        }
      }
    }
    *reinterpret_cast<uint4*>(y + i * 8) = make_uint4(m[0], m[1], m[2], m[3]);
    if (idx != nullptr) {
      uint2 k;
      k.x = (am[0] & 0xffu) | ((am[0] >> 8) & 0xff00u) | ((am[1] & 0xffu) << 16) | ((am[1] & 0xff0000u) << 8);
      k.y = (am[2] & 0xffu) | ((am[2] >> 8) & 0xff00u) | ((am[3] & 0xffu) << 16) | ((am[3] & 0xff0000u) << 8);
      *reinterpret_cast<uint2*>(idx + i * 8) = k;
    }
  }
}

// Gather form (no atomics): one thread = one INPUT position x 8 channels; it sums dy of every
// output window whose first-max element is this position.
__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                                   const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                   const PoolGeom g, long long total) {
  const int G = g.Cp >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long r = i / G;
    const int w = (int)(r % g.W); r /= g.W;
    const int h = (int)(r % g.H); r /= g.H;
    const int t = (int)(r % g.T);
    const long long n = r / g.T;
    const uint4 ux = *reinterpret_cast<const uint4*>(x + i * 8);
    const __nv_bfloat16* xv = reinterpret_cast<const __nv_bfloat16*>(&ux);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // output windows containing (t,h,w): to in [ceil((t+pt-kt+1)/st), floor((t+pt)/st)]
    const int to_lo = max(0, (t + g.pt - g.kt + g.st) / g.st), to_hi = min(g.To - 1, (t + g.pt) / g.st);
    const int ho_lo = max(0, (h + g.ph - g.kh + g.sh) / g.sh), ho_hi = min(g.Ho - 1, (h + g.ph) / g.sh);
    const int wo_lo = max(0, (w + g.pw - g.kw + g.sw) / g.sw), wo_hi = min(g.Wo - 1, (w + g.pw) / g.sw);
    for (int to = to_lo; to <= to_hi; ++to)
      for (int ho = ho_lo; ho <= ho_hi; ++ho)
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
          const long long oo = ((((n * g.To + to) * g.Ho + ho) * g.Wo + wo) * (long long)g.Cp + cg * 8);
          const uint4 uy = *reinterpret_cast<const uint4*>(y + oo);
          const uint4 ud = *reinterpret_cast<const uint4*>(dy + oo);
          const __nv_bfloat16* yv = reinterpret_cast<const __nv_bfloat16*>(&uy);
          const __nv_bfloat16* dv_ = reinterpret_cast<const __nv_bfloat16*>(&ud);
          // is this position the FIRST element equal to the max in the window's scan order?
          unsigned eq = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (__bfloat162float(xv[j]) == __bfloat162float(yv[j])) eq |= 1u << j;
          if (eq == 0) continue;
          // knock out channels where an earlier window element already equals the max
          for (int a = 0; a < g.kt && eq; ++a) {
            const int tt = to * g.st - g.pt + a;
            if (tt < 0 || tt >= g.T) continue;
            for (int b = 0; b < g.kh && eq; ++b) {
              const int hh = ho * g.sh - g.ph + b;
              if (hh < 0 || hh >= g.H) continue;
              for (int c = 0; c < g.kw && eq; ++c) {
                const int ww = wo * g.sw - g.pw + c;
                if (ww < 0 || ww >= g.W) continue;
                const bool earlier = (tt < t) || (tt == t && (hh < h || (hh == h && ww < w)));
                if (!earlier) continue;
                const uint4 ue = *reinterpret_cast<const uint4*>(
                    x + ((((n * g.T + tt) * g.H + hh) * g.W + ww) * (long long)g.Cp + cg * 8));
                const __nv_bfloat16* ev = reinterpret_cast<const __nv_bfloat16*>(&ue);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (__bfloat162float(ev[j]) == __bfloat162float(yv[j])) eq &= ~(1u << j);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (eq & (1u << j)) acc[j] += __bfloat162float(dv_[j]);
        }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<uint4*>(dx + i * 8) = o;
  }
}

// Backward with the recorded argmax: one thread = one INPUT position x 8 channels; for every window that contains the
// position, 8 index bytes tell which channels routed their gradient here. (kt*kh*kw) x 24 bytes per thread instead of
// re-scanning each window for earlier ties.
__global__ void maxpool_bwd_idx_kernel(const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ dy,
                                       __nv_bfloat16* __restrict__ dx, const PoolGeom g, long long total) {
  const int G = g.Cp >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long r = i / G;
    const int w = (int)(r % g.W); r /= g.W;
    const int h = (int)(r % g.H); r /= g.H;
    const int t = (int)(r % g.T);
    const long long n = r / g.T;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int to_lo = max(0, (t + g.pt - g.kt + g.st) / g.st), to_hi = min(g.To - 1, (t + g.pt) / g.st);
    const int ho_lo = max(0, (h + g.ph - g.kh + g.sh) / g.sh), ho_hi = min(g.Ho - 1, (h + g.ph) / g.sh);
    const int wo_lo = max(0, (w + g.pw - g.kw + g.sw) / g.sw), wo_hi = min(g.Wo - 1, (w + g.pw) / g.sw);
    for (int to = to_lo; to <= to_hi; ++to)
      for (int ho = ho_lo; ho <= ho_hi; ++ho)
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
          // my offset inside window (to,ho,wo)
          const uint32_t code = (uint32_t)(((t - (to * g.st - g.pt)) * g.kh + (h - (ho * g.sh - g.ph))) * g.kw +
                                           (w - (wo * g.sw - g.pw)));
          const long long oo = ((((n * g.To + to) * g.Ho + ho) * g.Wo + wo) * (long long)g.Cp + cg * 8);
          const uint2 k = *reinterpret_cast<const uint2*>(idx + oo);
          const uint32_t pat = code * 0x01010101u;
          // bytes equal to code -> zero bytes in k ^ pat
          const uint32_t x0 = k.x ^ pat, x1 = k.y ^ pat;
          const uint32_t z0 = (x0 - 0x01010101u) & ~x0 & 0x80808080u, z1 = (x1 - 0x01010101u) & ~x1 & 0x80808080u;
          if ((z0 | z1) == 0u) continue;
          const uint4 ud = *reinterpret_cast<const uint4*>(dy + oo);
          const __nv_bfloat16* dv_ = reinterpret_cast<const __nv_bfloat16*>(&ud);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (((x0 >> (8 * j)) & 0xffu) == 0u) acc[j] += __bfloat162float(dv_[j]);
            if (((x1 >> (8 * j)) & 0xffu) == 0u) acc[4 + j] += __bfloat162float(dv_[4 + j]);
          }
        }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<uint4*>(dx + i * 8) = o;
  }
}

// The recorded-argmax backward for a compile-time window: the windows that contain an input position are enumerated by
// the position's offset (a, b, c) INSIDE the window - which is the very code the forward pass recorded - so the loops
// unroll and every window costs one 8-byte index load and a byte compare against a constant. The strides are
// compile-time too: window origin = (position + pad - offset) / stride where that division is exact.
template <int KT, int KH, int KW, int ST, int SH, int SW>
__global__ void __launch_bounds__(256)
maxpool_bwd_idx_win_kernel(const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ dy,
                           __nv_bfloat16* __restrict__ dx, const PoolGeom g, long long total) {
  const int G = g.Cp >> 3;
  const long long oW = g.Cp, oH = (long long)g.Wo * g.Cp, oT = (long long)g.Ho * oH;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long r = i / G;
    const int w = (int)(r % g.W); r /= g.W;
    const int h = (int)(r % g.H); r /= g.H;
    const int t = (int)(r % g.T);
    const long long n = r / g.T;
    const long long obase = n * g.To * oT + cg * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < KT; ++a) {
      const int tt = t + g.pt - a;
      const int to = tt / ST;
      if (tt < 0 || to * ST != tt || to >= g.To) continue;
#pragma unroll
      for (int b = 0; b < KH; ++b) {
        const int hh = h + g.ph - b;
        const int ho = hh / SH;
        if (hh < 0 || ho * SH != hh || ho >= g.Ho) continue;
#pragma unroll
        for (int c = 0; c < KW; ++c) {
          const int ww = w + g.pw - c;
          const int wo = ww / SW;
          if (ww < 0 || wo * SW != ww || wo >= g.Wo) continue;
          const long long oo = obase + to * oT + ho * oH + wo * oW;
          const uint2 k = *reinterpret_cast<const uint2*>(idx + oo);
          const uint32_t pat = (uint32_t)((a * KH + b) * KW + c) * 0x01010101u;
          // bytes equal to my offset -> zero bytes in k ^ pat
          const uint32_t x0 = k.x ^ pat, x1 = k.y ^ pat;
          const uint32_t z0 = (x0 - 0x01010101u) & ~x0 & 0x80808080u, z1 = (x1 - 0x01010101u) & ~x1 & 0x80808080u;
          if ((z0 | z1) == 0u) continue;
          const uint4 ud = *reinterpret_cast<const uint4*>(dy + oo);
          const __nv_bfloat16* dv_ = reinterpret_cast<const __nv_bfloat16*>(&ud);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (((x0 >> (8 * j)) & 0xffu) == 0u) acc[j] += __bfloat162float(dv_[j]);
            if (((x1 >> (8 * j)) & 0xffu) == 0u) acc[4 + j] += __bfloat162float(dv_[4 + j]);
          }
        }
      }
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<uint4*>(dx + i * 8) = o;
  }
}

// ---------------------------------------------------------------------------------- ingest
// src: fp32, element (b, v, c, t, h, w) at  b*sb + v*sv + c*sc + t*st + h*W + w  (strides in elements)
// dst: bf16 [B*nv][T][H][W][8] (clip n = b*nv + j reads view `view + j` of sample b); channel c < C gets (x - mean[c]) * inv_std[c], others 0.
// perm: optional int32 [B][n_series]; output segment j of sample b reads source segment perm[b][j].
// source element -> float: fp32 frames as they are, uint8 frames as transforms.ToTensor does (x / 255 in fp32)
__device__ __forceinline__ float ingest_ld(const float* p) { return *p; }
__device__ __forceinline__ float ingest_ld(const uint8_t* p) { return __fdiv_rn((float)*p, 255.f); }
__device__ __forceinline__ float2 ingest_ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ingest_ld2(const uint8_t* p) {
  const uchar2 u = *reinterpret_cast<const uchar2*>(p);
  return make_float2(__fdiv_rn((float)u.x, 255.f), __fdiv_rn((float)u.y, 255.f));
}

struct IngestArgs {
  const void* src;
  __nv_bfloat16* dst;
  const int* perm;
  long long sb, sv, sc, st;
  int B, C, T, H, W, view, n_series, nv;
  float mean[4], inv_std[4];
  // fp32 mode: n_planes bf16 split planes of the normalised value (plane k = bf16 of what planes < k left over),
  // plane k at dst + k * plane_stride elements; the bf16 mode writes one plane
  int n_planes;
  long long plane_stride;
};

// the part of v that bf16 did not capture (exact in fp32)
__device__ __forceinline__ float ingest_residual(float v) { return v - __bfloat162float(__float2bfloat16_rn(v)); }

template <typename Src>
__global__ void __launch_bounds__(256) ingest_kernel(const IngestArgs a) {
  const long long HW = (long long)a.H * a.W;
  const long long total = (long long)a.B * a.nv * a.T * HW;
  const int seg_len = a.n_series > 0 ? a.T / a.n_series : a.T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long hw = i % HW;
    const int t = (int)((i / HW) % a.T);
    const int n = (int)(i / (HW * a.T));   // output clip index = b * nv + (view - view0)
    const int b = n / a.nv;
    const int vw = a.view + (n - b * a.nv);
    int ts = t;
    if (a.perm) {
      const int seg = t / seg_len;
      ts = a.perm[n * a.n_series + seg] * seg_len + (t - seg * seg_len);
    }
    const Src* p = static_cast<const Src*>(a.src) + b * a.sb + vw * a.sv + ts * a.st + hw;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < a.C) v[c] = (ingest_ld(p + c * a.sc) - a.mean[c]) * a.inv_std[c];
    for (int k = 0; k < a.n_planes; ++k) {
      uint4 o;
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
      oh[0] = __floats2bfloat162_rn(v[0], v[1]);
      oh[1] = __floats2bfloat162_rn(v[2], v[3]);
      oh[2] = __floats2bfloat162_rn(0.f, 0.f);
      oh[3] = oh[2];
      *reinterpret_cast<uint4*>(a.dst + k * a.plane_stride + i * 8) = o;
      if (k + 1 < a.n_planes) {
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = ingest_residual(v[c]);
      }
    }
  }
}


// Space-to-depth ingest for the stride-2 7x7 stems: dst bf16 [B*nv][T][H/2][W/2+3][16],
// dst[n][t][hs][ws+2][(rh*2+rw)*4 + c] = norm(src(b, v, c, t, 2*hs+rh, 2*ws+rw)); two zero columns on the
// left and one on the right so that every 4-position window the stem reads is in bounds.
template <typename Src>
__global__ void __launch_bounds__(256) ingest_s2d_kernel(const IngestArgs a) {
  const int H2 = a.H >> 1, W2 = a.W >> 1, W2p = W2 + 3;
  const long long per_t = (long long)H2 * W2p;
  const long long total = (long long)a.B * a.nv * a.T * per_t;
  const int seg_len = a.n_series > 0 ? a.T / a.n_series : a.T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wsp = (int)(i % W2p);
    const int hs = (int)((i / W2p) % H2);
    const int t = (int)((i / per_t) % a.T);
    const int n = (int)(i / (per_t * a.T));
    uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
    const int ws = wsp - 2;
    float v[2][2][4] = {};
    if (ws >= 0 && ws < W2) {
      const int b = n / a.nv;
      const int vw = a.view + (n - b * a.nv);
      int ts = t;
      if (a.perm) {
        const int seg = t / seg_len;
        ts = a.perm[n * a.n_series + seg] * seg_len + (t - seg * seg_len);
      }
      const Src* p = static_cast<const Src*>(a.src) + b * a.sb + vw * a.sv + ts * a.st + (long long)(2 * hs) * a.W + 2 * ws;
#pragma unroll
      for (int rh = 0; rh < 2; ++rh)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float2 f = make_float2(0.f, 0.f);
          if (c < a.C) {
            f = ingest_ld2(p + c * a.sc + rh * a.W);
            f.x = (f.x - a.mean[c]) * a.inv_std[c];
            f.y = (f.y - a.mean[c]) * a.inv_std[c];
          }
          v[rh][0][c] = f.x;
          v[rh][1][c] = f.y;
        }
    }
    for (int k = 0; k < a.n_planes; ++k) {
      __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
      __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
      h0[0] = __floats2bfloat162_rn(v[0][0][0], v[0][0][1]); h0[1] = __floats2bfloat162_rn(v[0][0][2], v[0][0][3]);
      h0[2] = __floats2bfloat162_rn(v[0][1][0], v[0][1][1]); h0[3] = __floats2bfloat162_rn(v[0][1][2], v[0][1][3]);
      h1[0] = __floats2bfloat162_rn(v[1][0][0], v[1][0][1]); h1[1] = __floats2bfloat162_rn(v[1][0][2], v[1][0][3]);
      h1[2] = __floats2bfloat162_rn(v[1][1][0], v[1][1][1]); h1[3] = __floats2bfloat162_rn(v[1][1][2], v[1][1][3]);
      uint4* d = reinterpret_cast<uint4*>(a.dst + k * a.plane_stride + i * 16);
      d[0] = o0;
      d[1] = o1;
      if (k + 1 < a.n_planes) {
#pragma unroll
        for (int rh = 0; rh < 2; ++rh)
#pragma unroll
          for (int rw = 0; rw < 2; ++rw)
#pragma unroll
            for (int c = 0; c < 4; ++c) v[rh][rw][c] = ingest_residual(v[rh][rw][c]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------- host
static int flat_grid(long long total, int threads) {
  long long g = ceil_div_ll(total, threads);
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int avgpool_fwd(const void* x, float* out, int N, int S, int C, int Cp, int ld_out, cudaStream_t stream) {
  dim3 block(128, 8);
  dim3 grid(ceil_div(Cp, 128), N);
  avgpool_fwd_kernel<<<grid, block, 0, stream>>>((const __nv_bfloat16*)x, out, S, C, Cp, ld_out);
  DV_LAUNCH_OK();
  return kOk;
}

int avgpool_bwd(const float* dout, void* dx, int N, int S, int C, int Cp, int ld_out, cudaStream_t stream) {
  const long long total = (long long)N * S * Cp;
  avgpool_bwd_kernel<<<flat_grid(total, 256), 256, 0, stream>>>(dout, (__nv_bfloat16*)dx, S, C, Cp,
                                                                ld_out, total);
  DV_LAUNCH_OK();
  return kOk;
}

int maxpool_fwd(const void* x, void* y, uint8_t* idx, const PoolGeom& g, cudaStream_t stream) {
  const long long total = (long long)g.N * g.To * g.Ho * g.Wo * (g.Cp / 8);
  if (idx != nullptr && g.kt * g.kh * g.kw > 255) return fail(kUnsupported, "max-pool window too large for 1-byte argmax");
#define DV_POOL_WIN(KT, KH, KW)                                                                               \
  if (g.kt == KT && g.kh == KH && g.kw == KW) {                                                               \
    maxpool_fwd_win_kernel<KT, KH, KW><<<flat_grid(total, 256), 256, 0, stream>>>(                            \
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, idx, g, total);                                           \
    DV_LAUNCH_OK();                                                                                           \
    return kOk;                                                                                               \
  }
  DV_POOL_WIN(3, 3, 3)
  DV_POOL_WIN(1, 3, 3)
  DV_POOL_WIN(2, 2, 2)
  DV_POOL_WIN(1, 2, 2)
#undef DV_POOL_WIN
  maxpool_fwd_kernel<<<flat_grid(total, 256), 256, 0, stream>>>((const __nv_bfloat16*)x,
                                                                (__nv_bfloat16*)y, idx, g, total);
  DV_LAUNCH_OK();
  return kOk;
}

int maxpool_bwd(const void* x, const void* y, const void* dy, void* dx, const PoolGeom& g,
                cudaStream_t stream) {
  const long long total = (long long)g.N * g.T * g.H * g.W * (g.Cp / 8);
  maxpool_bwd_kernel<<<flat_grid(total, 256), 256, 0, stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx,
      g, total);
  DV_LAUNCH_OK();
  return kOk;
}

int maxpool_bwd_idx(const uint8_t* idx, const void* dy, void* dx, const PoolGeom& g, cudaStream_t stream) {
  const long long total = (long long)g.N * g.T * g.H * g.W * (g.Cp / 8);
#define DV_POOL_WIN(KT, KH, KW, ST, SH, SW)                                                                   \
  if (g.kt == KT && g.kh == KH && g.kw == KW && g.st == ST && g.sh == SH && g.sw == SW) {                     \
    maxpool_bwd_idx_win_kernel<KT, KH, KW, ST, SH, SW><<<flat_grid(total, 256), 256, 0, stream>>>(            \
        idx, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, g, total);                                         \
    DV_LAUNCH_OK();                                                                                           \
    return kOk;                                                                                               \
  }
  DV_POOL_WIN(3, 3, 3, 1, 1, 1)
  DV_POOL_WIN(3, 3, 3, 2, 2, 2)
  DV_POOL_WIN(1, 3, 3, 1, 2, 2)
  DV_POOL_WIN(2, 2, 2, 2, 2, 2)
  DV_POOL_WIN(1, 2, 2, 1, 2, 2)
#undef DV_POOL_WIN
  maxpool_bwd_idx_kernel<<<flat_grid(total, 256), 256, 0, stream>>>(idx, (const __nv_bfloat16*)dy,
                                                                    (__nv_bfloat16*)dx, g, total);
  DV_LAUNCH_OK();
  return kOk;
}

int ingest(const void* src, int src_u8, void* dst, const int* perm, long long sb, long long sv, long long sc,
           long long st, int B, int C, int T, int H, int W, int view, int nv, int n_series,
           const float* mean, const float* stdv, int s2d, cudaStream_t stream, int n_planes, long long plane_stride) {
  IngestArgs a;
  a.n_planes = n_planes; a.plane_stride = plane_stride;
  a.src = src; a.dst = (__nv_bfloat16*)dst; a.perm = perm;
  a.sb = sb; a.sv = sv; a.sc = sc; a.st = st;
  a.B = B; a.C = C; a.T = T; a.H = H; a.W = W; a.view = view; a.n_series = n_series; a.nv = nv;
  for (int c = 0; c < 4; ++c) {
    a.mean[c] = (mean && c < C) ? mean[c] : 0.f;
    a.inv_std[c] = (stdv && c < C) ? 1.f / stdv[c] : 1.f;
  }
  if (s2d) {
    const long long total2 = (long long)B * nv * T * (H / 2) * (W / 2 + 3);
    if (src_u8) ingest_s2d_kernel<uint8_t><<<flat_grid(total2, 256), 256, 0, stream>>>(a);
    else ingest_s2d_kernel<float><<<flat_grid(total2, 256), 256, 0, stream>>>(a);
    DV_LAUNCH_OK();
    return kOk;
  }
  const long long total = (long long)B * nv * T * H * W;
  if (src_u8) ingest_kernel<uint8_t><<<flat_grid(total, 256), 256, 0, stream>>>(a);
  else ingest_kernel<float><<<flat_grid(total, 256), 256, 0, stream>>>(a);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
