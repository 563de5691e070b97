// Layout conversion kernels at the module boundary: the reference modules own fp32 parameters in
// torch's (Cout, Cin, kt, kh, kw) order and exchange fp32 NCDHW activations
// (backbone/select_backbone.py:30-31); the kernels work on bf16 NDHWC with padded channels.
#include <cuda_bf16.h>

#include "host_common.h"

namespace dv {

// w: fp32 [Cout][Cin][taps] -> wf: bf16 [Cout_p][taps][Cin_p] and wt: bf16 [Cin_p][taps][Cout_p]
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                    __nv_bfloat16* __restrict__ wt, int Cout, int Cin, int taps,
                                    int Cout_p, int Cin_p) {
  const long long total = (long long)Cout_p * taps * Cin_p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_p);
    const int tap = (int)((i / Cin_p) % taps);
    const int co = (int)(i / ((long long)Cin_p * taps));
    float v = 0.f;
    if (co < Cout && ci < Cin) v = w[((long long)co * Cin + ci) * taps + tap];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    if (wf) wf[i] = h;
    if (wt) wt[((long long)ci * taps + tap) * Cout_p + co] = h;
  }
}

// dwp: fp32 [Cout_p][taps][Cin_p] -> dw: fp32 [Cout][Cin][taps]; dw = beta*dw + dwp
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout,
                                    int Cin, int taps, int Cin_p, float beta) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((long long)taps * Cin));
    const float g = dwp[((long long)co * taps + tap) * Cin_p + ci];
    dw[i] = (beta != 0.f) ? fmaf(beta, dw[i], g) : g;
  }
}

// x: fp32 [N][C][S] (S = T*H*W) -> y: bf16 [N][S][Cp], via a 32x32 smem transpose
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                      int C, int Cp, long long S) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long s0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* xn = x + (long long)n * C * S;
  __nv_bfloat16* yn = y + (long long)n * S * Cp;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const long long s = s0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && s < S) ? xn[(long long)c * S + s] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long s = s0 + j;
    const int c = c0 + threadIdx.x;
    if (s < S && c < Cp) yn[s * Cp + c] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}

// y: bf16 [N][S][Cp] -> x: fp32 [N][C][S]
__global__ void ndhwc_to_ncdhw_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ x,
                                      int C, int Cp, long long S) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long s0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const __nv_bfloat16* yn = y + (long long)n * S * Cp;
  float* xn = x + (long long)n * C * S;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long s = s0 + j;
    const int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (s < S && c < Cp) ? __bfloat162float(yn[s * Cp + c]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const long long s = s0 + threadIdx.x;
    if (c < C && s < S) xn[(long long)c * S + s] = tile[threadIdx.x][j];
  }
}

int pack_weights(const float* w, void* wf, void* wt, int Cout, int Cin, int taps, int Cout_p,
                 int Cin_p, cudaStream_t stream) {
  const long long total = (long long)Cout_p * taps * Cin_p;
  const int threads = 256;
  const int blocks = (int)std::min<long long>(ceil_div_ll(total, threads), 8 * 148);
  pack_weights_kernel<<<blocks, threads, 0, stream>>>(w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wt,
                                                      Cout, Cin, taps, Cout_p, Cin_p);
  DV_LAUNCH_OK();
  return kOk;
}

int unpack_wgrad(const float* dwp, float* dw, int Cout, int Cin, int taps, int Cin_p, float beta,
                 cudaStream_t stream) {
  const long long total = (long long)Cout * Cin * taps;
  const int threads = 256;
  const int blocks = (int)std::min<long long>(ceil_div_ll(total, threads), 8 * 148);
  unpack_wgrad_kernel<<<blocks, threads, 0, stream>>>(dwp, dw, Cout, Cin, taps, Cin_p, beta);
  DV_LAUNCH_OK();
  return kOk;
}

int ncdhw_to_ndhwc(const float* x, void* y, int N, int C, int Cp, long long S, cudaStream_t stream) {
  dim3 block(32, 8);
  dim3 grid((unsigned)ceil_div_ll(S, 32), (unsigned)ceil_div(Cp, 32), (unsigned)N);
  ncdhw_to_ndhwc_kernel<<<grid, block, 0, stream>>>(x, (__nv_bfloat16*)y, C, Cp, S);
  DV_LAUNCH_OK();
  return kOk;
}

int ndhwc_to_ncdhw(const void* y, float* x, int N, int C, int Cp, long long S, cudaStream_t stream) {
  dim3 block(32, 8);
  dim3 grid((unsigned)ceil_div_ll(S, 32), (unsigned)ceil_div(Cp, 32), (unsigned)N);
  ndhwc_to_ncdhw_kernel<<<grid, block, 0, stream>>>((const __nv_bfloat16*)y, x, C, Cp, S);
  DV_LAUNCH_OK();
  return kOk;
}

// Stem weights for the space-to-depth formulation (conv_fprop.cu: conv_stem_fprop_bf16).
// w: fp32 [Cout][Cin<=4][kt][7][7] -> ws: bf16 [Cout_p][kt*4][64],
// ws[co][kt_i*4 + a][b*16 + (rh*2+rw)*4 + c] = w[co][c][kt_i][2a+rh-1][2b+rw-1] (0 outside the 7x7 window).
__global__ void pack_stem_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ ws,
                                         int Cout, int Cin, int kt, int Cout_p) {
  const long long total = (long long)Cout_p * kt * 4 * 64;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % 64);
    const int a = (int)((i / 64) % 4);
    const int kti = (int)((i / 256) % kt);
    const int co = (int)(i / (256LL * kt));
    const int b = k >> 4, rh = (k >> 3) & 1, rw = (k >> 2) & 1, c = k & 3;
    const int kh = 2 * a + rh - 1, kw = 2 * b + rw - 1;
    float v = 0.f;
    if (co < Cout && c < Cin && kh >= 0 && kh < 7 && kw >= 0 && kw < 7)
      v = w[((((long long)co * Cin + c) * kt + kti) * 7 + kh) * 7 + kw];
    ws[i] = __float2bfloat16_rn(v);
  }
}

// dws: fp32 [Cout_p][kt*4][64] -> dw: fp32 [Cout][Cin][kt][7][7]; dw = beta*dw + gathered value
__global__ void unpack_stem_wgrad_kernel(const float* __restrict__ dws, float* __restrict__ dw, int Cout,
                                         int Cin, int kt, float beta) {
  const long long total = (long long)Cout * Cin * kt * 49;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(i % 7);
    const int kh = (int)((i / 7) % 7);
    const int kti = (int)((i / 49) % kt);
    const int c = (int)((i / (49LL * kt)) % Cin);
    const int co = (int)(i / (49LL * kt * Cin));
    const int a = (kh + 1) >> 1, rh = (kh + 1) & 1, b = (kw + 1) >> 1, rw = (kw + 1) & 1;
    const float g = dws[(((long long)co * kt + kti) * 4 + a) * 64 + b * 16 + (rh * 2 + rw) * 4 + c];
    dw[i] = (beta != 0.f) ? fmaf(beta, dw[i], g) : g;
  }
}

int pack_stem_weights(const float* w, void* ws, int Cout, int Cin, int kt, int Cout_p, cudaStream_t stream) {
  const long long total = (long long)Cout_p * kt * 256;
  pack_stem_weights_kernel<<<(int)std::min<long long>(ceil_div_ll(total, 256), 1184), 256, 0, stream>>>(
      w, (__nv_bfloat16*)ws, Cout, Cin, kt, Cout_p);
  DV_LAUNCH_OK();
  return kOk;
}

int unpack_stem_wgrad(const float* dws, float* dw, int Cout, int Cin, int kt, float beta, cudaStream_t stream) {
  const long long total = (long long)Cout * Cin * kt * 49;
  unpack_stem_wgrad_kernel<<<(int)std::min<long long>(ceil_div_ll(total, 256), 1184), 256, 0, stream>>>(
      dws, dw, Cout, Cin, kt, beta);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
