/* dualvar_b200 — C ABI of the B200-native DualVar pretraining hot path.
 *
 * The reference (lzhangbj/DualVar) has no FFI: its boundary for this path is the Python nn.Module
 * API, and every arithmetic op below it is a torch/ATen -> cuDNN/cuBLAS call.
 * Each entry point here replaces one of those library calls; the comment on each names the
 * reference call site (file:line under /root/reference) it stands in for. The Python host
 * package dualvar_b200/ binds these with ctypes and keeps the reference's
 * module signatures and state_dict names.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; dv_last_error() gives the message
 *    for the calling thread. Nothing throws across the boundary.
 *  - all pointers are DEVICE pointers unless the name ends in _host; `stream` is a cudaStream_t.
 *  - "NDHWC bf16" activations: [N][T][H][W][Cp] with Cp = channels rounded up to a multiple of 8,
 *    pad channels are zero. fp32 tensors at the module boundary are torch-contiguous NCDHW.
 *  - packed weights: bf16 [Cout_p][taps][Cin_p] (fprop, "wf") and [Cin_p][taps][Cout_p] (dgrad, "wt"),
 *    taps = kt*kh*kw in torch's (kt, kh, kw) order; packed fp32 weight gradient [Cout_p][taps][Cin_p].
 */
#ifndef DUALVAR_B200_H_
#define DUALVAR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dv_conv_geom {
  int32_t N, T, H, W;      /* input extent (batch of clips, frames, height, width) */
  int32_t Cin, Cout;       /* logical channel counts */
  int32_t Cin_p, Cout_p;   /* padded (multiple of 8) channel counts of the NDHWC buffers */
  int32_t kt, kh, kw;      /* filter size   (nn.Conv3d kernel_size) */
  int32_t st, sh, sw;      /* stride        (nn.Conv3d stride, 1 or 2) */
  int32_t pt, ph, pw;      /* zero padding  (nn.Conv3d padding) */
  int32_t To, Ho, Wo;      /* output extent: floor((X + 2p - k)/s) + 1 */
} dv_conv_geom;

const char* dv_last_error(void);
int dv_version(void);
/* 1 if the current device is compute capability 10.x (sm_100a kernels can run), else 0 */
int dv_device_ok(void);

/* ---- layout / parameter staging -------------------------------------------------------- */
/* nn.Conv3d.weight fp32 (Cout,Cin,kt,kh,kw) -> packed bf16 wf and/or wt (either may be NULL) */
int dv_pack_conv_weight(const float* w, void* wf, void* wt, const dv_conv_geom* g, void* stream);
/* packed fp32 dW -> torch-shaped fp32 grad; grad = beta*grad + dW */
int dv_unpack_conv_wgrad(const float* dw_packed, float* grad, const dv_conv_geom* g, float beta,
                         void* stream);
/* fp32 NCDHW -> bf16 NDHWC (padded) and back; S = T*H*W */
int dv_ncdhw_to_ndhwc_bf16(const float* x, void* y, int N, int C, int Cp, int64_t S, void* stream);
int dv_ndhwc_bf16_to_ncdhw(const void* y, float* x, int N, int C, int Cp, int64_t S, void* stream);

/* ---- convolution (tcgen05 implicit GEMM) -------------------------------------------------
 * Replaces F.conv3d / cuDNN fprop, dgrad, wgrad behind nn.Conv3d at
 * backbone/r21d.py:54,64  backbone/r3d.py:33  backbone/c3d.py:15-44  backbone/s3dg.py:11,39-41. */
/* y = conv(x, w) (+ bias); if bn_stats != NULL also accumulates per-channel sum and sum of squares
 * of the stored outputs into bn_stats[0:Cout_p] and bn_stats[Cout_p:2*Cout_p] (double, caller zeroes). */
int dv_conv3d_fprop_bf16(const void* x, const void* wf, void* y, double* bn_stats,
                         const float* bias_padded, const dv_conv_geom* g, void* stream);
/* dx = conv_transpose(dy, w) */
int dv_conv3d_dgrad_bf16(const void* dy, const void* wt, void* dx, const dv_conv_geom* g,
                         void* stream);
/* dw_packed (fp32 [Cout_p][taps][Cin_p]) = correlation(x, dy); buffer is overwritten */
int dv_conv3d_wgrad_bf16(const void* x, const void* dy, float* dw_packed, const dv_conv_geom* g,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DUALVAR_B200_H_ */
