// BatchNorm3d (training and eval mode) around the convolutions, on bf16 NDHWC activations.
//
// Replaces ATen native_batch_norm / SyncBatchNorm + ReLU + residual add on the reference path
// (reference: nn.BatchNorm3d / nn.ReLU / `x + res` in backbone/r21d.py:56-57,106-122,
//  backbone/r3d.py:74-89, backbone/c3d.py:16-46, backbone/s3dg.py:16-27,44-64).
//
// Forward : the conv epilogue already produced per-channel sum / sum-of-squares (double).
//           bn_finalize turns them into (scale, shift) + saved (mean, invstd) and updates the
//           running statistics; bn_apply computes  out = relu?(s1*y1+b1 [+ s2*y2+b2] [+ res]).
// Backward: bn_bwd_reduce  -> per-channel  sum(g), sum(g*y)   with g = dout * (out > 0)
//           bn_bwd_finalize-> dgamma, dbeta and the coefficients of dy = A*g + B*y + C
//           bn_bwd_apply   -> dy (bf16) and optionally g itself (gradient of the residual branch).
// All passes are HBM-bound: 16-byte vector accesses, one thread owns 8 consecutive channels so the
// per-channel coefficients sit in registers; a CTA covers (channel groups) x (rows) so that a warp
// reads contiguous memory.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <algorithm>

#include "bn_math.cuh"
#include "host_common.h"

namespace dv {

struct Vec8 {
  float v[8];
};

__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  Vec8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}

// streaming 16-byte accesses: every element of these passes is touched once
__device__ __forceinline__ uint4 ldcs16(const __nv_bfloat16* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ Vec8 unpack8(const uint4& u) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  Vec8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void stcs8(__nv_bfloat16* p, const Vec8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ void store8(__nv_bfloat16* p, const Vec8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ Vec8 loadf8(const float* p) {
  Vec8 r;
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// ------------------------------------------------------------------------------ finalize
// stats: [2][Cp] double (sum, sumsq) over `count` elements per channel.
// ss: [2][Cp] float (scale, shift) for bn_apply; saved: [2][Cp] float (mean, invstd).
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ ss,
                                   float* __restrict__ saved, int C, int Cp, double count, float eps,
                                   float momentum, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  bn_finalize_channel(c, training ? stats[c] : 0.0, training ? stats[Cp + c] : 0.0, gamma, beta, running_mean,
                      running_var, ss, saved, C, Cp, count, eps, momentum, training);
}

// ------------------------------------------------------------------------------ apply
struct ApplyArgs {
  const __nv_bfloat16* y1; const float* ss1;
  const __nv_bfloat16* y2; const float* ss2;   // optional second normalised input
  const __nv_bfloat16* res;                    // optional plain residual
  __nv_bfloat16* out;
  long long rows;
  int Cp;       // channels (padded) of y1/y2/res rows
  int out_ld;   // channel stride of out rows (>= Cp; concat destination)
  int out_coff; // channel offset inside out rows
  int relu;
};

// kRows independent rows per thread and iteration: all their 16-byte loads are issued before the first use, which
// is what keeps enough bytes in flight per SM to reach the HBM rate (tests/diag/bn_bw.py).
//
// kFlat (dense tensors: every row is exactly Cp channels wide): the tensor is ONE array of 16-byte vectors, thread t of
// block b owns the vectors b*256 + t + j*stride with stride = gridDim.x*256 a multiple of the vectors per row (flat_grid),
// so its channel group - hence its coefficients - never changes, every warp is full and every warp access is a 512-byte
// aligned run. The (channel group) x (rows) blocks of the general form have partial warps and runs that straddle
// 128-byte lines whenever Cp/8 is not a power of two (144, 88, 232 ... channels: 5.4 instead of 6.1 TB/s,
// tests/diag/bn_bw.py).
template <bool kY2, bool kRes, int kRows, bool kFlat>
__global__ void __launch_bounds__(256) bn_apply_kernel(const ApplyArgs a) {
  const int G = a.Cp >> 3;
  const long long first = kFlat ? (long long)blockIdx.x * 256 + threadIdx.x : (long long)blockIdx.x * blockDim.y + threadIdx.y;
  const long long limit = kFlat ? a.rows * G : a.rows;
  const long long stride = kFlat ? (long long)gridDim.x * 256 : (long long)gridDim.x * blockDim.y;
  for (int cg = kFlat ? (int)(first % G) : (int)threadIdx.x; cg < G; cg += kFlat ? G : (int)blockDim.x) {
    const Vec8 s1 = loadf8(a.ss1 + cg * 8), b1 = loadf8(a.ss1 + a.Cp + cg * 8);
    Vec8 s2, b2;
    if (kY2) { s2 = loadf8(a.ss2 + cg * 8); b2 = loadf8(a.ss2 + a.Cp + cg * 8); }
    for (long long r0 = first; r0 < limit; r0 += kRows * stride) {
      uint4 u1[kRows], u2[kRows], u3[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long r = r0 + k * stride;
        if (r < limit) {
          const long long off = kFlat ? r * 8 : r * a.Cp + cg * 8;
          u1[k] = ldcs16(a.y1 + off);
          if (kY2) u2[k] = ldcs16(a.y2 + off);
          if (kRes) u3[k] = ldcs16(a.res + off);
        }
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long r = r0 + k * stride;
        if (r < limit) {
          Vec8 x = unpack8(u1[k]);
#pragma unroll
          for (int i = 0; i < 8; ++i) x.v[i] = fmaf(x.v[i], s1.v[i], b1.v[i]);
          if (kY2) {
            const Vec8 t = unpack8(u2[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) x.v[i] += fmaf(t.v[i], s2.v[i], b2.v[i]);
          }
          if (kRes) {
            const Vec8 t = unpack8(u3[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) x.v[i] += t.v[i];
          }
          if (a.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x.v[i] = fmaxf(x.v[i], 0.f);
          }
          stcs8(a.out + (kFlat ? r * 8 : r * a.out_ld + a.out_coff + cg * 8), x);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------ backward reduce
// sums: [2][Cp] double: sum(g), sum(g*y);  g = dout * (out > 0) if relu else dout.
// dout/out rows may live inside a wider (concat) tensor: ld / channel offset given.
struct BwdArgs {
  const __nv_bfloat16* dout; const __nv_bfloat16* out; const __nv_bfloat16* y;
  const __nv_bfloat16* dout2;  // optional second gradient contribution (dense [rows][Cp]); g uses dout + dout2
  long long rows;
  int Cp;
  int o_ld, o_coff;   // layout of dout / out rows
  int relu;
  double* sums;       // reduce
  const float* ss;    // optional [2][Cp] forward scale/shift: ReLU mask = (scale*y + shift > 0), `out` not read
  const float* coef;  // apply: [3][Cp] A, B, C
  __nv_bfloat16* dy;  // apply
  __nv_bfloat16* g_out;  // apply: optional masked gradient (residual branch), dense [rows][Cp]
};

// kMask: 0 = no ReLU, 1 = mask recomputed from (scale, shift, y), 2 = mask read from `out`
template <bool kD2, int kMask, int kRows>
__global__ void __launch_bounds__(256) bn_bwd_reduce_flat_kernel(const BwdArgs a) {
  // dense tensors, flat vector indexing (see bn_apply_kernel): a thread's channel group is fixed; its 16 partial sums go
  // to shared memory and are added per channel in a FIXED order (the threads of a channel group are t0, t0 + G, ...),
  // so a block's contribution is reproducible - only the order of the double atomics across blocks varies, as before
  __shared__ float red[256][17];
  const int G = a.Cp >> 3;
  const long long first = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long limit = a.rows * G;
  const long long stride = (long long)gridDim.x * 256;
  const int cg = (int)(first % G);
  float sg[8], sgy[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sg[i] = 0.f; sgy[i] = 0.f; }
  Vec8 fs, fb;
  if (kMask == 1) { fs = loadf8(a.ss + cg * 8); fb = loadf8(a.ss + a.Cp + cg * 8); }
  for (long long v0 = first; v0 < limit; v0 += kRows * stride) {
    uint4 ud[kRows], uy[kRows], ue[kRows], uo[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const long long v = v0 + k * stride;
      if (v < limit) {
        ud[k] = ldcs16(a.dout + v * 8);
        uy[k] = ldcs16(a.y + v * 8);
        if (kD2) ue[k] = ldcs16(a.dout2 + v * 8);
        if (kMask == 2) uo[k] = ldcs16(a.out + v * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const long long v = v0 + k * stride;
      if (v < limit) {
        Vec8 d = unpack8(ud[k]);
        const Vec8 yv = unpack8(uy[k]);
        if (kD2) {
          const Vec8 e = unpack8(ue[k]);
#pragma unroll
          for (int i = 0; i < 8; ++i) d.v[i] += e.v[i];
        }
        if (kMask == 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) d.v[i] = fmaf(yv.v[i], fs.v[i], fb.v[i]) > 0.f ? d.v[i] : 0.f;
        } else if (kMask == 2) {
          const Vec8 o = unpack8(uo[k]);
#pragma unroll
          for (int i = 0; i < 8; ++i) d.v[i] = o.v[i] > 0.f ? d.v[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { sg[i] += d.v[i]; sgy[i] = fmaf(d.v[i], yv.v[i], sgy[i]); }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x][i] = sg[i];
    red[threadIdx.x][8 + i] = sgy[i];
  }
  __syncthreads();
  const int shift = (int)(((long long)blockIdx.x * 256) % G);     // channel group of thread 0
  for (int o = threadIdx.x; o < 2 * a.Cp; o += 256) {
    const int which = o >= a.Cp ? 1 : 0;
    const int c = o - which * a.Cp;
    int t = (c >> 3) - shift;
    if (t < 0) t += G;
    float sum = 0.f;
    for (; t < 256; t += G) sum += red[t][which * 8 + (c & 7)];
    atomicAdd(&a.sums[o], (double)sum);
  }
}

template <bool kD2, int kMask, int kRows>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BwdArgs a) {
  extern __shared__ float red[];  // [blockDim.y][2][8*blockDim.x]
  const int G = a.Cp >> 3;
  const int GT = blockDim.x;
  for (int cg0 = 0; cg0 < G; cg0 += GT) {
    const int cg = cg0 + threadIdx.x;
    float sg[8], sgy[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sg[i] = 0.f; sgy[i] = 0.f; }
    if (cg < G) {
      Vec8 fs, fb;
      if (kMask == 1) { fs = loadf8(a.ss + cg * 8); fb = loadf8(a.ss + a.Cp + cg * 8); }
      const long long stride = (long long)gridDim.x * blockDim.y;
      for (long long r0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; r0 < a.rows; r0 += kRows * stride) {
        uint4 ud[kRows], uy[kRows], ue[kRows], uo[kRows];
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          const long long r = r0 + k * stride;
          if (r < a.rows) {
            const long long oo = r * a.o_ld + a.o_coff + cg * 8, yo = r * a.Cp + cg * 8;
            ud[k] = ldcs16(a.dout + oo);
            uy[k] = ldcs16(a.y + yo);
            if (kD2) ue[k] = ldcs16(a.dout2 + yo);
            if (kMask == 2) uo[k] = ldcs16(a.out + oo);
          }
        }
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          const long long r = r0 + k * stride;
          if (r < a.rows) {
            Vec8 d = unpack8(ud[k]);
            const Vec8 yv = unpack8(uy[k]);
            if (kD2) {
              const Vec8 e = unpack8(ue[k]);
#pragma unroll
              for (int i = 0; i < 8; ++i) d.v[i] += e.v[i];
            }
            if (kMask == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) d.v[i] = fmaf(yv.v[i], fs.v[i], fb.v[i]) > 0.f ? d.v[i] : 0.f;
            } else if (kMask == 2) {
              const Vec8 o = unpack8(uo[k]);
#pragma unroll
              for (int i = 0; i < 8; ++i) d.v[i] = o.v[i] > 0.f ? d.v[i] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { sg[i] += d.v[i]; sgy[i] = fmaf(d.v[i], yv.v[i], sgy[i]); }
          }
        }
      }
    }
    float* mine = red + (size_t)threadIdx.y * 16 * GT;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mine[threadIdx.x * 8 + i] = sg[i];
      mine[8 * GT + threadIdx.x * 8 + i] = sgy[i];
    }
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int j = tid; j < 16 * GT; j += blockDim.x * blockDim.y) {
      float s = 0.f;
      for (int ry = 0; ry < (int)blockDim.y; ++ry) s += red[(size_t)ry * 16 * GT + j];
      const int which = j / (8 * GT);
      const int c = cg0 * 8 + (j - which * 8 * GT);
      if (c < a.Cp) atomicAdd(&a.sums[which * a.Cp + c], (double)s);
    }
    __syncthreads();
  }
}

// sums_local: this rank's sums (parameter gradients); sums_global + count_global: over the whole
// (cross-replica) batch, used for the dy coefficients. saved: mean/invstd from forward.
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums_local,
                                       const double* __restrict__ sums_global,
                                       const float* __restrict__ gamma, const float* __restrict__ saved,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef, int C, int Cp, double count_global,
                                       float grad_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  const bool real = c < C;
  bn_bwd_finalize_channel(c, real ? sums_local[c] : 0.0, real ? sums_local[Cp + c] : 0.0, real ? sums_global[c] : 0.0,
                          real ? sums_global[Cp + c] : 0.0, gamma, saved, dgamma, dbeta, coef, C, Cp, count_global,
                          grad_beta);
}

template <bool kD2, int kMask, bool kGout, int kRows, bool kFlat>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BwdArgs a) {
  const int G = a.Cp >> 3;
  const long long first = kFlat ? (long long)blockIdx.x * 256 + threadIdx.x : (long long)blockIdx.x * blockDim.y + threadIdx.y;
  const long long limit = kFlat ? a.rows * G : a.rows;
  const long long stride = kFlat ? (long long)gridDim.x * 256 : (long long)gridDim.x * blockDim.y;
  for (int cg = kFlat ? (int)(first % G) : (int)threadIdx.x; cg < G; cg += kFlat ? G : (int)blockDim.x) {
    const Vec8 A = loadf8(a.coef + cg * 8), B = loadf8(a.coef + a.Cp + cg * 8),
               C = loadf8(a.coef + 2 * a.Cp + cg * 8);
    Vec8 fs, fb;
    if (kMask == 1) { fs = loadf8(a.ss + cg * 8); fb = loadf8(a.ss + a.Cp + cg * 8); }
    for (long long r0 = first; r0 < limit; r0 += kRows * stride) {
      uint4 ud[kRows], uy[kRows], ue[kRows], uo[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long r = r0 + k * stride;
        if (r < limit) {
          const long long oo = kFlat ? r * 8 : r * a.o_ld + a.o_coff + cg * 8, yo = kFlat ? r * 8 : r * a.Cp + cg * 8;
          ud[k] = ldcs16(a.dout + oo);
          uy[k] = ldcs16(a.y + yo);
          if (kD2) ue[k] = ldcs16(a.dout2 + yo);
          if (kMask == 2) uo[k] = ldcs16(a.out + oo);
        }
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long r = r0 + k * stride;
        if (r < limit) {
          const long long off = kFlat ? r * 8 : r * a.Cp + cg * 8;
          Vec8 d = unpack8(ud[k]);
          const Vec8 yv = unpack8(uy[k]);
          if (kD2) {
            const Vec8 e = unpack8(ue[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) d.v[i] += e.v[i];
          }
          if (kMask == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) d.v[i] = fmaf(yv.v[i], fs.v[i], fb.v[i]) > 0.f ? d.v[i] : 0.f;
          } else if (kMask == 2) {
            const Vec8 o = unpack8(uo[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) d.v[i] = o.v[i] > 0.f ? d.v[i] : 0.f;
          }
          if (kGout) store8(a.g_out + off, d);
          Vec8 res;
#pragma unroll
          for (int i = 0; i < 8; ++i) res.v[i] = fmaf(A.v[i], d.v[i], fmaf(B.v[i], yv.v[i], C.v[i]));
          store8(a.dy + off, res);
        }
      }
    }
  }
}

// out = a + b (bf16), gradient accumulation where two consumers meet (reference: autograd add).
__global__ void __launch_bounds__(256) add_bf16_kernel(const __nv_bfloat16* __restrict__ a,
                                                       const __nv_bfloat16* __restrict__ b,
                                                       __nv_bfloat16* __restrict__ out, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    Vec8 x = load8(a + i * 8);
    const Vec8 y = load8(b + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) x.v[j] += y.v[j];
    store8(out + i * 8, x);
  }
}

// ------------------------------------------------------------------------------ host
static void row_block(int Cp, long long rows, dim3* block, int* grid) {
  const int G = Cp / 8;
  const int gt = G < 256 ? G : 256;
  int ry = 256 / gt;
  if (ry < 1) ry = 1;
  *block = dim3(gt, ry);
  long long want = ceil_div_ll(rows, ry * 4);  // ~4 rows per thread minimum
  const long long cap = (long long)sm_count() * 8;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  *grid = (int)want;
}

// Grid of the flat (dense) kernels: 256-thread blocks whose count is a multiple of G / gcd(G, 32), so that the stride
// gridDim * 256 of a thread's vectors is a multiple of the G vectors of a row (fixed channel group per thread) and of
// 32 (aligned warp runs). DV_BN_FLAT=0 keeps the (channel group) x (rows) blocks everywhere.
static int g_bn_flat = -1;
static bool flat_grid(int Cp, long long rows, int* grid) {
  if (g_bn_flat < 0) {
    const char* e = getenv("DV_BN_FLAT");
    g_bn_flat = e ? atoi(e) : 1;
  }
  if (!g_bn_flat) return false;
  const int G = Cp / 8;
  int gq = G;
  while (gq % 2 == 0 && G / gq < 32) gq /= 2;      // gq = G / gcd(G, 32)
  const long long nvec = rows * G;
  long long want = ceil_div_ll(nvec, 256 * 4);      // ~4 vectors per thread minimum
  const long long cap = (long long)sm_count() * 8;
  if (want > cap) want = cap;
  want = want / gq * gq;
  if (want < gq) want = gq;
  if (want > 65535LL * 16) return false;
  *grid = (int)want;
  return true;
}

int bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                float* running_var, float* ss, float* saved, int C, int Cp, double count, float eps,
                float momentum, int training, cudaStream_t stream) {
  bn_finalize_kernel<<<ceil_div(Cp, 128), 128, 0, stream>>>(stats, gamma, beta, running_mean,
                                                             running_var, ss, saved, C, Cp, count,
                                                             eps, momentum, training);
  DV_LAUNCH_OK();
  return kOk;
}

int bn_apply(const void* y1, const float* ss1, const void* y2, const float* ss2, const void* res,
             void* out, long long rows, int Cp, int out_ld, int out_coff, int relu,
             cudaStream_t stream) {
  ApplyArgs a;
  a.y1 = (const __nv_bfloat16*)y1; a.ss1 = ss1;
  a.y2 = (const __nv_bfloat16*)y2; a.ss2 = ss2;
  a.res = (const __nv_bfloat16*)res;
  a.out = (__nv_bfloat16*)out;
  a.rows = rows; a.Cp = Cp; a.out_ld = out_ld; a.out_coff = out_coff; a.relu = relu;
  dim3 block; int grid;
  if (out_ld == Cp && out_coff == 0 && flat_grid(Cp, rows, &grid)) {
    if (a.y2 && a.res) bn_apply_kernel<true, true, 2, true><<<grid, 256, 0, stream>>>(a);
    else if (a.y2) bn_apply_kernel<true, false, 2, true><<<grid, 256, 0, stream>>>(a);
    else if (a.res) bn_apply_kernel<false, true, 4, true><<<grid, 256, 0, stream>>>(a);
    else bn_apply_kernel<false, false, 8, true><<<grid, 256, 0, stream>>>(a);
    DV_LAUNCH_OK();
    return kOk;
  }
  row_block(Cp, rows, &block, &grid);
  if (a.y2 && a.res) bn_apply_kernel<true, true, 2, false><<<grid, block, 0, stream>>>(a);
  else if (a.y2) bn_apply_kernel<true, false, 2, false><<<grid, block, 0, stream>>>(a);
  else if (a.res) bn_apply_kernel<false, true, 4, false><<<grid, block, 0, stream>>>(a);
  else bn_apply_kernel<false, false, 8, false><<<grid, block, 0, stream>>>(a);
  DV_LAUNCH_OK();
  return kOk;
}

int bn_bwd_reduce(const void* dout, const void* dout2, const void* out, const void* y, const float* ss,
                  double* sums, long long rows, int Cp, int o_ld, int o_coff, int relu, cudaStream_t stream) {
  BwdArgs a = {};
  a.ss = ss;
  a.dout2 = (const __nv_bfloat16*)dout2;
  a.dout = (const __nv_bfloat16*)dout; a.out = (const __nv_bfloat16*)out;
  a.y = (const __nv_bfloat16*)y; a.rows = rows; a.Cp = Cp; a.o_ld = o_ld; a.o_coff = o_coff;
  a.relu = relu; a.sums = sums;
  dim3 block; int grid;
  const int mask = !relu ? 0 : (ss ? 1 : 2);
  if (o_ld == Cp && o_coff == 0 && flat_grid(Cp, rows, &grid)) {
#define DV_REDF(D2, M, R) bn_bwd_reduce_flat_kernel<D2, M, R><<<grid, 256, 0, stream>>>(a)
    if (a.dout2) {
      if (mask == 0) DV_REDF(true, 0, 2); else if (mask == 1) DV_REDF(true, 1, 2); else DV_REDF(true, 2, 2);
    } else {
      if (mask == 0) DV_REDF(false, 0, 4); else if (mask == 1) DV_REDF(false, 1, 4); else DV_REDF(false, 2, 2);
    }
#undef DV_REDF
    DV_LAUNCH_OK();
    return kOk;
  }
  row_block(Cp, rows, &block, &grid);
  const size_t smem = (size_t)block.y * 16 * block.x * sizeof(float);
#define DV_RED(D2, M, R) bn_bwd_reduce_kernel<D2, M, R><<<grid, block, smem, stream>>>(a)
  if (a.dout2) {
    if (mask == 0) DV_RED(true, 0, 2); else if (mask == 1) DV_RED(true, 1, 2); else DV_RED(true, 2, 2);
  } else {
    if (mask == 0) DV_RED(false, 0, 4); else if (mask == 1) DV_RED(false, 1, 4); else DV_RED(false, 2, 2);
  }
#undef DV_RED
  DV_LAUNCH_OK();
  return kOk;
}

int bn_bwd_finalize(const double* sums_local, const double* sums_global, const float* gamma,
                    const float* saved, float* dgamma, float* dbeta, float* coef, int C, int Cp,
                    double count_global, float grad_beta, cudaStream_t stream) {
  bn_bwd_finalize_kernel<<<ceil_div(Cp, 128), 128, 0, stream>>>(
      sums_local, sums_global, gamma, saved, dgamma, dbeta, coef, C, Cp, count_global, grad_beta);
  DV_LAUNCH_OK();
  return kOk;
}

int bn_bwd_apply(const void* dout, const void* dout2, const void* out, const void* y, const float* ss,
                 const float* coef, void* dy, void* g_out, long long rows, int Cp, int o_ld, int o_coff, int relu,
                 cudaStream_t stream) {
  BwdArgs a = {};
  a.ss = ss;
  a.dout2 = (const __nv_bfloat16*)dout2;
  a.dout = (const __nv_bfloat16*)dout; a.out = (const __nv_bfloat16*)out;
  a.y = (const __nv_bfloat16*)y; a.rows = rows; a.Cp = Cp; a.o_ld = o_ld; a.o_coff = o_coff;
  a.relu = relu; a.coef = coef; a.dy = (__nv_bfloat16*)dy; a.g_out = (__nv_bfloat16*)g_out;
  dim3 block; int grid;
  const int mask = !relu ? 0 : (ss ? 1 : 2);
  const bool flat = o_ld == Cp && o_coff == 0 && flat_grid(Cp, rows, &grid);
  if (flat) block = dim3(256, 1);
  else row_block(Cp, rows, &block, &grid);
#define DV_APP(D2, M, GO, R)                                                            \
  do {                                                                                  \
    if (flat) bn_bwd_apply_kernel<D2, M, GO, R, true><<<grid, block, 0, stream>>>(a);   \
    else bn_bwd_apply_kernel<D2, M, GO, R, false><<<grid, block, 0, stream>>>(a);       \
  } while (0)
#define DV_APP_M(D2, GO, R)                                                               \
  do {                                                                                    \
    if (mask == 0) DV_APP(D2, 0, GO, R); else if (mask == 1) DV_APP(D2, 1, GO, R); else DV_APP(D2, 2, GO, R); \
  } while (0)
  if (a.dout2) { if (a.g_out) DV_APP_M(true, true, 2); else DV_APP_M(true, false, 2); }
  else { if (a.g_out) DV_APP_M(false, true, 2); else DV_APP_M(false, false, 4); }
#undef DV_APP_M
#undef DV_APP
  DV_LAUNCH_OK();
  return kOk;
}

int add_bf16(const void* a, const void* b, void* out, long long n, cudaStream_t stream) {
  const long long n8 = n / 8;
  int grid = (int)std::min<long long>(ceil_div_ll(n8, 256), (long long)sm_count() * 8);
  if (grid < 1) grid = 1;
  add_bf16_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                            (__nv_bfloat16*)out, n8);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
