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
/* number of kernels this library has launched in the calling process */
int64_t dv_launch_count(void);
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
/* Consumer-side BatchNorm: the convolution whose INPUT is z = relu?(scale*y_prev + shift) with y_prev the raw output
 * of the convolution below and ss_prev = [2][Cin_p] scale, shift from dv_bn_finalize (reference: nn.BatchNorm3d +
 * nn.ReLU between the spatial and temporal convs of SpatioTemporalConv, backbone/r21d.py:56-57,67-70). z is never
 * written to HBM: the kernels load the raw tile by TMA and apply the affine + ReLU in shared memory in front of
 * tcgen05.mma (rows the TMA zero-filled - the conv padding - stay zero). Results are bit-identical to
 * dv_bn_apply followed by dv_conv3d_fprop_bf16 / dv_conv3d_wgrad_bf16 on its output. */
int dv_conv3d_fprop_bnrelu_bf16(const void* y_prev, const float* ss_prev, int relu, const void* wf, void* y,
                                double* bn_stats, const float* bias_padded, const dv_conv_geom* g, void* stream);
int dv_conv3d_wgrad_bnrelu_bf16(const void* y_prev, const float* ss_prev, int relu, const void* dy, float* dw_packed,
                                const dv_conv_geom* g, void* stream);
/* dx = conv_transpose(dy, w) */
int dv_conv3d_dgrad_bf16(const void* dy, const void* wt, void* dx, const dv_conv_geom* g,
                         void* stream);
/* dgrad with the BatchNorm-backward reduction of the layer below fused into its epilogue: dx is the
 * gradient of z = relu?(scale*y_prev + shift) (y_prev: bf16 [N][T][H][W][Cin_p], the raw conv output the
 * BatchNorm in front of this conv normalised; ss_prev = [2][Cin_p] scale/shift, NULL = no ReLU). Adds
 * sum(g) and sum(g*y_prev), g = dx * (scale*y_prev + shift > 0), over all positions into
 * sums[0:Cin_p] / sums[Cin_p:2*Cin_p] (double, caller zeroes) - what dv_bn_bwd_reduce would compute
 * from dx in a separate pass (autograd of nn.BatchNorm3d + nn.ReLU, backbone/r21d.py:56-57,68). */
int dv_conv3d_dgrad_bnred_bf16(const void* dy, const void* wt, void* dx, const dv_conv_geom* g,
                               const void* y_prev, const float* ss_prev, double* sums, void* stream);
/* The data gradient of a 3x3 spatial convolution with 64 input channels (the gradient's output columns) with the kh
 * taps STACKED along N: one tcgen05.mma per (kw, K step) over the unshifted 16h x 8w box of dy with N = 192, the epilogue
 * adds the three row-shifted 64-column results (autograd of nn.Conv3d in backbone/r21d.py:54-57, same result as
 * dv_conv3d_dgrad_bf16 up to the fp32 summation order). w_stack: bf16 [192][3][Cout_p], row s*64 + ci and tap kw hold
 * W[co][ci][kh = 2 - s][kw] (a permutation of the transposed pack). y_prev / ss_prev / sums as for
 * dv_conv3d_dgrad_bnred_bf16, or all NULL. dv_conv3d_dgrad_stack_ok: 1 if the geometry is eligible. */
int dv_conv3d_dgrad_stack_ok(const dv_conv_geom* g);
int dv_conv3d_dgrad_stack_bf16(const void* dy, const void* w_stack, void* dx, const dv_conv_geom* g, const void* y_prev,
                               const float* ss_prev, double* sums, void* stream);
/* dw_packed (fp32 [Cout_p][taps][Cin_p]) = correlation(x, dy); buffer is overwritten */
int dv_conv3d_wgrad_bf16(const void* x, const void* dy, float* dw_packed, const dv_conv_geom* g,
                         void* stream);


/* ---- BatchNorm3d (+ReLU, +residual) on bf16 NDHWC ------------------------------------------
 * Replaces nn.BatchNorm3d / SyncBatchNorm + nn.ReLU + `x + res` at backbone/r21d.py:56-57,106-122,
 * backbone/r3d.py:74-89, backbone/c3d.py:16-46, backbone/s3dg.py:16-27,44-64. */
/* stats [2][Cp] double (sum, sumsq over `count` values per channel, from the conv epilogue, already
 * all-reduced across replicas when cross-replica BN is on) -> scale_shift [2][Cp], saved [2][Cp]
 * (mean, invstd); updates running_mean/var (momentum, unbiased var) when training != 0; when
 * training == 0 uses the running statistics instead of stats. */
int dv_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float* scale_shift, float* saved, int C, int Cp, double count,
                   float eps, float momentum, int training, void* stream);
/* out = relu?( ss1.scale*y1 + ss1.shift  [+ ss2.scale*y2 + ss2.shift]  [+ res] ); out rows may sit
 * inside a wider tensor (channel stride out_ld, channel offset out_coff) for concat-free Inception. */
int dv_bn_apply(const void* y1, const float* ss1, const void* y2, const float* ss2, const void* res,
                void* out, int64_t rows, int Cp, int out_ld, int out_coff, int relu, void* stream);
/* sums [2][Cp] double += (sum g, sum g*y), g = dout*(out>0) if relu else dout. When the activation was
 * relu(scale*y+shift) with no residual/second branch pass its scale_shift as mask_ss: the mask is then
 * recomputed from y and `out` is not read (one HBM stream less).
 * dout2 (optional, dense [rows][Cp]) is a second gradient contribution summed on the fly — where two
 * consumers of an activation meet (block input: main path + shortcut) no separate add pass is needed. */
int dv_bn_bwd_reduce(const void* dout, const void* dout2, const void* out, const void* y, const float* mask_ss,
                     double* sums, int64_t rows, int Cp, int o_ld, int o_coff, int relu, void* stream);
/* dgamma/dbeta (= grad_beta*old + local sums) and coef [3][Cp] of dy = A*g + B*y + C from the global sums */
int dv_bn_bwd_finalize(const double* sums_local, const double* sums_global, const float* gamma,
                       const float* saved, float* dgamma, float* dbeta, float* coef, int C, int Cp,
                       double count_global, float grad_beta, void* stream);
/* dy = A*g + B*y + C; g_out (optional) = g, the gradient flowing into the residual branch */
int dv_bn_bwd_apply(const void* dout, const void* dout2, const void* out, const void* y, const float* mask_ss,
                    const float* coef, void* dy, void* g_out, int64_t rows, int Cp, int o_ld, int o_coff, int relu,
                    void* stream);
/* out = a + b over n bf16 elements (n % 8 == 0): gradient accumulation where two consumers meet */
int dv_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);

/* ---- pooling / ingest ----------------------------------------------------------------------- */
/* nn.AdaptiveAvgPool3d((1,1,1)) (model/simclr.py:166): x [N][S][Cp] bf16 -> out [N][ld_out] fp32 */
int dv_avgpool_fwd(const void* x, float* out, int N, int S, int C, int Cp, int ld_out, void* stream);
int dv_avgpool_bwd(const float* dout, void* dx, int N, int S, int C, int Cp, int ld_out, void* stream);
typedef struct dv_pool_geom {
  int32_t N, T, H, W, To, Ho, Wo, Cp;
  int32_t kt, kh, kw, st, sh, sw, pt, ph, pw;
} dv_pool_geom;
/* nn.MaxPool3d (backbone/c3d.py:18-39, backbone/s3dg.py:105,151,162,173,190) */
int dv_maxpool3d_fwd(const void* x, void* y, const dv_pool_geom* g, void* stream);
int dv_maxpool3d_bwd(const void* x, const void* y, const void* dy, void* dx, const dv_pool_geom* g,
                     void* stream);
/* Training forward that also records, per output element, the window offset (a*kh + b)*kw + c of the first maximum
 * (1 byte, [N][To][Ho][Wo][Cp]); backward is then a gather over the windows containing each input position - the
 * tie rule (first maximum in scan order) is the same as dv_maxpool3d_bwd's and ATen's. */
int dv_maxpool3d_fwd_idx(const void* x, void* y, uint8_t* argmax, const dv_pool_geom* g, void* stream);
int dv_maxpool3d_bwd_idx(const uint8_t* argmax, const void* dy, void* dx, const dv_pool_geom* g, void* stream);
/* fp32 clips -> bf16 NDHWC (C<=4 -> 8 channels): element (b,view,c,t,h,w) of src is at
 * b*sb + view*sv + c*sc + t*st + h*W + w. Output clip n = b*nv + j holds view (view + j) of sample b,
 * i.e. nv consecutive views per sample in the reference's block.view(-1, C, T, H, W) order
 * (model/simclr.py:352). Optional Normalize (mean_host/std_host, C floats on the HOST, NULL = identity;
 * utils/transforms.py:57-63 + the view/transpose of pretrain.py:386-389 expressed through the strides)
 * and optional segment shuffle: perm int32 [B*nv][n_series] on the device, output segment j of clip n
 * reads source segment perm[n][j] (model/simclr.py:378-383).
 * s2d != 0 writes the space-to-depth layout [B*nv][T][H/2][W/2+3][16] (channel (rh*2+rw)*4+c, 2 zero
 * columns left, 1 right) consumed by dv_conv3d_stem_*; otherwise [B*nv][T][H][W][8]. */
int dv_ingest_clips(const float* src, void* dst, const int32_t* perm, int64_t sb, int64_t sv, int64_t sc,
                    int64_t st, int B, int C, int T, int H, int W, int view, int nv, int n_series,
                    const float* mean_host, const float* std_host, int s2d, void* stream);
/* Same from uint8 frames (decoded images, 4x fewer bytes over PCIe): every element is first converted as
 * transforms.ToTensor does (x / 255 in fp32, utils/augmentation.py:361-364), then normalised. */
int dv_ingest_clips_u8(const uint8_t* src, void* dst, const int32_t* perm, int64_t sb, int64_t sv, int64_t sc,
                       int64_t st, int B, int C, int T, int H, int W, int view, int nv, int n_series,
                       const float* mean_host, const float* std_host, int s2d, void* stream);

/* ---- fp32 heads and objectives --------------------------------------------------------------- */
/* C = alpha*op(A)*op(B) + beta*C (+bias[n]) (relu). Row-major. ta: A stored [K][M]; tb: B stored [N][K].
 * Heads = 1x1x1 nn.Conv3d (model/simclr.py:168-180); similarity matmuls (model/simclr.py:202,297,
 * model/moco.py:413-414,429-430). */
int dv_sgemm(int ta, int tb, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
             int ldb, float beta, float* C, int ldc, const float* bias, int relu, void* stream);
int dv_colsum(const float* X, float* out, int M, int N, int ld, float beta, void* stream);
int dv_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);
/* F.normalize over the last dim d (model/simclr.py:359,367,393) and its backward */
int dv_l2norm_fwd(const float* x, float* y, float* inv_norm, int64_t rows, int d, float eps, void* stream);
int dv_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int64_t rows, int d,
                  void* stream);
/* Row-wise softmax cross-entropy on a similarity matrix S [R][C] with the reference's column order
 * (model/simclr.py:205-221,308-329; model/moco.py:431-432): writes logits (positive first, self dropped,
 * /T), adds sum_r CE to *loss_sum, counts top-1/top-5 hits (utils/utils.py:75-92) and overwrites S with
 * dLoss/dS * grad_scale. self_col may be NULL (no column dropped). */
int dv_contrast_rows(float* S, float* logits, const int32_t* self_col, const int32_t* pos_col, int R,
                     int C, int ld_s, int ld_logits, float inv_T, float grad_scale, float* loss_sum,
                     int32_t* hits, void* stream);
/* The similarity GEMM FUSED with the row-wise log-sum-exp / cross-entropy (replaces torch.matmul(features, features.T)
 * + masks + CrossEntropyLoss of model/simclr.py:198-221,299-329 and einsum('nc,ck->nk') + cat + CE of
 * model/moco.py:426-438). dv_sim_ce_fwd: S[r][col0 + c] = a[r] . b[c] (b row-major [C][ldb], or d-major [d][ldb] =
 * MoCo's queue buffer when b_dmajor = 1), the logits S/T in the reference's column order (positive first, self column
 * dropped, rest ascending; self_col / pos_col are S-column indices, self_col may be NULL) and per (row, 128-column
 * block) the online-softmax partial (max, sum exp) into partials [R][dv_sim_ce_blocks(C)][2].
 * dv_sim_ce_finish: combines the partials with the col0 leading columns the caller filled itself (MoCo: q.k), adds
 * sum_r CE to *loss_sum, counts top-1 / top-5 (utils/utils.py:75-92), writes the leading columns' logits and
 * overwrites S [R][col0 + C] with dLoss/dS * grad_scale (0 in the self column). */
int dv_sim_ce_blocks(int C);
int dv_sim_ce_fwd(const float* a, int lda, const float* b, int ldb, int b_dmajor, int R, int C, int d, float* S,
                  int ld_s, int col0, float* logits, int ld_logits, const int32_t* self_col, const int32_t* pos_col,
                  float inv_T, float* partials, void* stream);
int dv_sim_ce_finish(float* S, int ld_s, int R, int C, int col0, const float* partials, float* logits, int ld_logits,
                     const int32_t* self_col, const int32_t* pos_col, float inv_T, float grad_scale, float* loss_sum,
                     int32_t* hits, void* stream);
/* Shuffle-rank loss + gradient (model/simclr.py:231-278; clip_max<=0: model/moco.py:440-480) */
int dv_rank_loss(const float* a, const float* b, float* da, float* db, float* logits, float* loss_sum,
                 int32_t* hits, int B, int s, int e, float theta, float clip_max, float weight,
                 void* stream);
/* out[b][perm[b][j]] = in[b][j] (inverse != 0: out[b][j] = in[b][perm[b][j]]) (model/simclr.py:389-392) */
int dv_permute_segments(const float* in, float* out, const int32_t* perm, int B, int s, int e,
                        int inverse, void* stream);
int dv_segment_sum(const float* in, float* out, int64_t rows, int s, int e, float scale, void* stream);
int dv_segment_bcast(const float* in, float* out, int64_t rows, int s, int e, float scale, float beta,
                     void* stream);
int dv_rowdot(const float* a, const float* b, float* out, int rows, int d, int ld_out, void* stream);
int dv_row_axpy(const float* alpha, int ld_alpha, const float* x, float* y, int rows, int d, float beta,
                void* stream);

/* ---- stride-2 7x7 stem (Cin <= 4) on the space-to-depth input -------------------------------------
 * Replaces the first nn.Conv3d of R(2+1)D / R3D / S3D (backbone/r21d.py:227, backbone/r3d.py:139,
 * backbone/s3dg.py:143): geometry g describes the ORIGINAL conv (H, W = frame size, kernel (kt,7,7),
 * stride (1,2,2), padding (pt,3,3)). Packed stem weights: bf16 [Cout_p][kt*4][64]; packed gradient fp32
 * of the same shape. */
int dv_pack_stem_weight(const float* w, void* ws, const dv_conv_geom* g, void* stream);
int dv_unpack_stem_wgrad(const float* dws, float* grad, const dv_conv_geom* g, float beta, void* stream);
int dv_conv3d_stem_fprop_bf16(const void* x_s2d, const void* ws, void* y, double* bn_stats,
                              const float* bias_padded, const dv_conv_geom* g, void* stream);
int dv_conv3d_stem_wgrad_bf16(const void* x_s2d, const void* dy, float* dws, const dv_conv_geom* g, void* stream);

/* ---- MoCo bookkeeping -------------------------------------------------------------------------
 * theta_k = m*theta_k + (1-m)*theta_q for every parameter tensor in one launch (model/moco.py:328-334).
 * chunk_table: device int64 [n_chunks][3] = (k pointer, q pointer, element count <= 8192). */
int dv_moco_momentum_update(const int64_t* chunk_table, int n_chunks, float m, void* stream);
/* queue[:, ptr:ptr+B] = keys^T; queue fp32 (d, K), keys fp32 (B, d); K % B == 0 (model/moco.py:343-351) */
int dv_moco_enqueue(const float* keys, float* queue, int B, int d, int K, int ptr, void* stream);
/* the same with the pointer read on the device from the model's queue_ptr buffer (int64, model/moco.py:345), and
 * queue_ptr = (queue_ptr + batch) % K (model/moco.py:352-353): no host value in the launches, so a MoCo step can be
 * captured as a CUDA graph and the reference's per-step int(queue_ptr) device-to-host sync disappears */
int dv_moco_enqueue_at(const float* keys, float* queue, int B, int d, int K, const int64_t* queue_ptr, void* stream);
int dv_moco_advance_ptr(int64_t* queue_ptr, int batch, int K, void* stream);

/* ---- S3D-G self-gating on channel slices of the Inception concat tensor (backbone/s3dg.py:68-78,130) ----
 * x rows have `ld` channels, the branch occupies [coff, coff+C). */
int dv_slice_mean(const void* x, float* out, int N, int S, int C, int ld, int coff, void* stream);
/* the gate's Linear + sigmoid in one launch: w[n][j] = sigmoid(fc_bias[j] + sum_k mean[n][k] * fc_weight[j][k]); and its
 * backward from dw = dLoss/dw: dpre = dw * w * (1 - w), grad_weight = dpre^T mean, grad_bias = column sums of dpre,
 * dmean = dpre fc_weight (backbone/s3dg.py:74-77; N, C <= 1024) */
int dv_gate_fc_fwd(const float* mean, const float* fc_weight, const float* fc_bias, float* w, int N, int C, void* stream);
int dv_gate_fc_bwd(const float* dw, const float* w, const float* mean, const float* fc_weight, float* dpre, float* grad_weight,
                   float* grad_bias, float* dmean, int N, int C, void* stream);
int dv_gate_scale(void* x, const float* w, int N, int S, int C, int ld, int coff, void* stream);
/* dw[n][c] = sum_s dout * relu(scale*y+shift)  (the un-gated activation is recomputed from y) */
int dv_gate_bwd_reduce(const void* dout, const void* y, const float* ss, float* dw, int N, int S, int C, int Cp,
                       int ld, int coff, void* stream);
/* dz = w*dout + dmean/S, dense [N*S][Cp] */
int dv_gate_bwd_apply(const void* dout, const float* w, const float* dmean, void* dz, int N, int S, int C, int Cp,
                      int ld, int coff, void* stream);
int dv_sigmoid_fwd(const float* x, float* y, int64_t n, void* stream);
int dv_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);

/* ---- retrieval (classifier.py:963-983) ----------------------------------------------------------------
 * out (fp64 [n][d]) = L2-normalised rows of (feat - column mean), everything evaluated in fp64 */
int dv_retrieval_prepare(const float* feat, double* mean, double* out, int n, int d, void* stream);
/* sim (fp64 [n_test][n_train], optional fp32 copy) = test @ train^T; idx (int64 [n_test][k]) = indices of the k
 * largest similarities per row in descending order, exact ties broken towards the lowest index */
int dv_retrieval_sim_topk(const double* test, const double* train, double* sim, float* sim32, int64_t* idx,
                          int n_test, int n_train, int d, int k, void* stream);

/* ---- optimizer ----------------------------------------------------------------
 * optim.SGD(momentum, weight_decay) of pretrain.py:272 over every parameter tensor in one launch:
 * g = grad + wd*p ; buf = first_step ? g : momentum*buf + g ; p -= lr*buf.
 * chunk_table: device int64 [n_chunks][4] = (param ptr, grad ptr, momentum-buffer ptr, count <= 8192). */
int dv_sgd_momentum_step(const int64_t* chunk_table, int n_chunks, float lr, float momentum, float weight_decay,
                         int first_step, void* stream);

/* ---- cross-replica BatchNorm statistics: one-shot all-reduce over NVLink peer memory ---------------
 * Replaces the per-layer collectives of nn.SyncBatchNorm (pretrain.py:244) for vectors of <= 4096 doubles.
 * Every rank allocates a zero-initialised symmetric buffer of dv_allreduce_small_buffer_bytes() bytes that is
 * mapped into all peers (peer_buffers[q] = address of rank q's buffer in THIS process, host array of `world`
 * entries); all ranks call with the same n and seq = 1, 2, 3, ... in the same order - or all with seq = 0, in which
 * case the kernel numbers the calls itself from a counter in its buffer (replayable inside a CUDA graph). In place,
 * sum in rank order (bit-identical on all ranks). */
int64_t dv_allreduce_small_buffer_bytes(void);
int dv_allreduce_small_f64(double* inout, int n, const int64_t* peer_buffers, int rank, int world, int64_t seq,
                           void* stream);

/* dv_bn_finalize (training mode) and dv_bn_bwd_finalize with the cross-replica exchange of their inputs fused in:
 * one launch pushes this rank's statistics to all peers, waits for theirs, sums in rank order and finalises
 * (peer_buffers / rank / world / seq as for dv_allreduce_small_f64; count_global = values per channel over all ranks). */
int dv_bn_finalize_sync(const double* stats, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, float* scale_shift, float* saved, int C, int Cp, double count_global,
                        float eps, float momentum, const int64_t* peer_buffers, int rank, int world, int64_t seq,
                        void* stream);
int dv_bn_bwd_finalize_sync(const double* sums_local, const float* gamma, const float* saved, float* dgamma,
                            float* dbeta, float* coef, int C, int Cp, double count_global, float grad_beta,
                            const int64_t* peer_buffers, int rank, int world, int64_t seq, void* stream);
/* Failure reporting of the peer exchange (NCCL's watchdog / ProcessGroup error state on the reference path):
 * a rank that waits longer than the time-out (default 600 s) for a peer, or whose peers exchanged a different
 * per-rank BatchNorm count, records the fact in a host-mapped word and returns - the context is not destroyed.
 * dv_comm_status returns 0 (ok), 1 (time-out) or 2 (unequal per-rank counts) and the peer / call number involved
 * (NULL = not wanted); clear != 0 resets it. */
int dv_comm_set_timeout(double seconds);
int dv_comm_status(int* peer, int64_t* seq, int clear);


/* ---- fp32 mode ("1e-4 mode") ----------------------------------------------------------------
 * Activations and gradients are fp32 NDHWC [N][T][H][W][Cp]; every conv operand is also kept as n_planes (1..3)
 * bf16 SPLIT PLANES, x = p0 + p1 + p2 with p_k = bf16(x - p0 - .. - p(k-1)), plane k at base + k*plane_stride
 * elements. A convolution is the sum over i + j < n_planes of bf16 products (x plane i) * (w plane j), each ONE call
 * of an *_f32acc / *_acc entry point below: the tcgen05 kernels of the bf16 mode, whose epilogue ADDS the fp32
 * accumulator tile to the fp32 destination (or, for the first product, stores it there) instead of rounding it to bf16.
 * Replaces the same reference calls as the bf16 entry points, for the fp32 (TF32-off) reference path. */
/* fp32 values -> n_planes fp32 tensors [n_planes][n] holding the bf16-representable parts (feed each plane to
 * dv_pack_conv_weight / dv_pack_stem_weight, which then rounds exactly) */
int dv_f32_split_planes(const float* src, float* dst_planes, int64_t n, int n_planes, void* stream);
/* fp32 mode, ALL plane products of a convolution in one launch: y = sum_{i + j < K} conv(x_i, w_j) accumulated in one
 * TMEM accumulator (the products are extra taps, smallest contributions first) and stored to the fp32 output once -
 * instead of one dv_conv3d_fprop_f32acc launch per product (one store + K(K+1)/2 - 1 read-modify-write passes).
 * x_planes / dy_planes: the K bf16 split planes, plane_stride elements apart; wf_all: bf16 [Cout_p][K*taps][Cin_p] and
 * wt_all: bf16 [Cin_p][K*taps][Cout_p] with plane j in tap slots [j*taps, (j+1)*taps). fprop of a strided layer reads
 * every plane through one tensor map per stride-parity class its taps reach (at most 12 maps, a non-zero status beyond);
 * stats (may be NULL): double [2*Cout_p], per-channel sum / sum of squares of y ADDED to it by the epilogue (zero it
 * first) - the fused BatchNorm batch statistics, possible here because the launch holds the finished sum. */
int dv_conv3d_fprop_f32planes(const void* x_planes, int64_t plane_stride, int n_planes, const void* wf_all, float* y,
                              double* stats, const float* bias_padded, const dv_conv_geom* g, void* stream);
int dv_conv3d_dgrad_f32planes(const void* dy_planes, int64_t plane_stride, int n_planes, const void* wt_all, float* dx,
                              const dv_conv_geom* g, void* stream);
/* the weight gradient of the same sum in one launch: dw_packed fp32 [Cout_p][taps][Cin_p] (overwritten) =
 * sum_{i + j < K} wgrad(x_i, dy_j); both operands are CONTIGUOUS stacks [K][N][...] of their bf16 split planes (the
 * kernel walks them as one tensor of K*N clips), every position tile is contracted once per product into the same
 * TMEM accumulators and reduced into dw_packed once. Unpack with dv_unpack_conv_wgrad. */
int dv_conv3d_wgrad_f32planes(const void* x_planes, const void* dy_planes, int n_planes, float* dw_packed,
                              const dv_conv_geom* g, void* stream);
/* y (+)= conv(x_plane, w_plane) (+ bias); y fp32 [N][To][Ho][Wo][Cout_p]. accumulate = 0: the first product of a sum,
 * y is overwritten (no zero fill needed); accumulate = 1: added to y */
int dv_conv3d_fprop_f32acc(const void* x_plane, const void* wf_plane, float* y, const float* bias_padded,
                           const dv_conv_geom* g, int accumulate, void* stream);
/* dx (+)= conv_transpose(dy_plane, w_plane); dx fp32 [N][T][H][W][Cin_p]; accumulate as above (positions that no tap
 * reaches are zeroed by the accumulate = 0 call) */
int dv_conv3d_dgrad_f32acc(const void* dy_plane, const void* wt_plane, float* dx, const dv_conv_geom* g, int accumulate,
                           void* stream);
/* dw_packed += correlation(x_plane, dy_plane) (dv_conv3d_wgrad_bf16 without the zero fill) */
int dv_conv3d_wgrad_bf16_acc(const void* x_plane, const void* dy_plane, float* dw_packed, const dv_conv_geom* g,
                             void* stream);
/* the stride-2 7x7 stem on space-to-depth planes (dv_conv3d_stem_fprop_bf16 / _wgrad_bf16), accumulating */
int dv_conv3d_stem_fprop_f32acc(const void* x_s2d_plane, const void* ws_plane, float* y, const float* bias_padded,
                                const dv_conv_geom* g, int accumulate, void* stream);
int dv_conv3d_stem_wgrad_bf16_acc(const void* x_s2d_plane, const void* dy_plane, float* dws, const dv_conv_geom* g,
                                  void* stream);
/* BatchNorm batch statistics of an fp32 conv output: adds per-channel sum / sum of squares (double, caller zeroes)
 * into stats[0:Cp] / stats[Cp:2Cp]; dv_bn_finalize(_sync) then applies unchanged */
int dv_f32_colstats(const float* y, double* stats, int64_t rows, int Cp, void* stream);
/* out = relu?(ss1*y1 [+ ss2*y2] [+ res]) as fp32 and (out_planes != NULL) as split planes (dv_bn_apply); out / out_planes
 * are the channel slice [out_coff, out_coff + Cp) of rows with out_ld channels (out_ld = Cp, out_coff = 0: dense) */
int dv_f32_bn_apply(const float* y1, const float* ss1, const float* y2, const float* ss2, const float* res, float* out,
                    void* out_planes, int64_t plane_stride, int n_planes, int64_t rows, int Cp, int out_ld, int out_coff,
                    int relu, void* stream);
/* dv_bn_bwd_reduce / dv_bn_bwd_apply on fp32 tensors; dy leaves as split planes (it only feeds dgrad and wgrad),
 * g_out (nullable) is the masked gradient for the residual branch */
int dv_f32_bn_bwd_reduce(const float* dout, const float* dout2, const float* out, const float* y, const float* mask_ss,
                         double* sums, int64_t rows, int Cp, int o_ld, int o_coff, int relu, void* stream);
int dv_f32_bn_bwd_apply(const float* dout, const float* dout2, const float* out, const float* y, const float* mask_ss,
                        const float* coef, void* dy_planes, int64_t plane_stride, int n_planes, float* g_out,
                        int64_t rows, int Cp, int o_ld, int o_coff, int relu, void* stream);
/* S3D-G self-gating on an fp32 concat slice (dv_slice_mean / dv_gate_scale / dv_gate_bwd_*); gate_scale also rewrites
 * the slice's split planes */
int dv_f32_slice_mean(const float* x, float* out, int N, int S, int C, int ld, int coff, void* stream);
int dv_f32_gate_scale(float* x, void* x_planes, int64_t plane_stride, int n_planes, const float* w, int N, int S, int C,
                      int ld, int coff, void* stream);
int dv_f32_gate_bwd_reduce(const float* dout, const float* y, const float* ss, float* dw, int N, int S, int C, int Cp,
                           int ld, int coff, void* stream);
int dv_f32_gate_bwd_apply(const float* dout, const float* w, const float* dmean, float* dz, int N, int S, int C, int Cp,
                          int ld, int coff, void* stream);
int dv_f32_add(const float* a, const float* b, float* out, int64_t n, void* stream);
/* fp32 tensor -> split planes */
int dv_f32_split(const float* x, void* planes, int64_t plane_stride, int n_planes, int64_t n, void* stream);
int dv_f32_avgpool_fwd(const float* x, float* out, int N, int S, int C, int Cp, int ld_out, void* stream);
int dv_f32_avgpool_bwd(const float* dout, float* dx, int N, int S, int C, int Cp, int ld_out, void* stream);
/* nn.MaxPool3d on fp32 NDHWC; argmax (nullable) as in dv_maxpool3d_fwd_idx, y_planes (nullable) = split planes of y */
int dv_f32_maxpool3d_fwd(const float* x, float* y, uint8_t* argmax, void* y_planes, int64_t plane_stride, int n_planes,
                         const dv_pool_geom* g, void* stream);
int dv_f32_maxpool3d_bwd(const uint8_t* argmax, const float* dy, float* dx, const dv_pool_geom* g, void* stream);
int dv_f32_ndhwc_to_ncdhw(const float* y, float* x, int N, int C, int Cp, int64_t S, void* stream);
int dv_f32_ncdhw_to_ndhwc(const float* x, float* y, int N, int C, int Cp, int64_t S, void* stream);
/* dv_ingest_clips / dv_ingest_clips_u8 writing n_planes split planes of the normalised clips */
int dv_ingest_clips_planes(const void* src, int src_is_u8, void* dst_planes, int64_t plane_stride, int n_planes,
                           const int32_t* perm, int64_t sb, int64_t sv, int64_t sc, int64_t st, int B, int C, int T,
                           int H, int W, int view, int nv, int n_series, const float* mean_host, const float* std_host,
                           int s2d, void* stream);

/* ---- frame staging (first stage) ---------------------------------------------------
 * A.Scale((128,171)) (PIL bicubic) + A.RandomCrop(112) of the loader's null_transform (utils/augmentation.py:125-176,
 * pretrain.py:491-497), bit-exact with Pillow's 8-bit resampler. frames: uint8 [B][n_views*T][Hs][Ws][3] decoded
 * frames; tmp: uint8 scratch [B*n_views*T][Hs][scale_w][4] (RGBX, 4-byte aligned); crop_lu: int32 [B][n_views][2] = (left, upper) of each
 * clip's crop in the scaled frame (what RandomCrop draws as h_start, w_start); out: uint8 [B][3][n_views*T][crop_h][crop_w],
 * the layout dv_ingest_clips_u8 reads (ToTensor, Normalize and the NDHWC conversion happen there). */
int dv_frames_scale_crop_u8(const uint8_t* frames, uint8_t* tmp, uint8_t* out, const int32_t* crop_lu, int B, int n_views,
                            int T, int Hs, int Ws, int scale_w, int scale_h, int crop_w, int crop_h, void* stream);
/* A.ColorJitter (utils/augmentation.py:429-660, block = 1) on the planar uint8 clips dv_frames_scale_crop_u8 wrote:
 * ToTensor (x / 255) then, per frame, torchvision's tensor adjust_brightness / contrast / saturation / hue in the order
 * and with the factors of params[frame][12] = {apply, b, 1-b, c, 1-c, s, 1-s, h, op0..op3} (op: 0 brightness, 1 contrast,
 * 2 saturation, 3 hue; drawn on the host in the reference's RNG order). out: float32 [B][3][F][H][W] in [0, 1] - what the
 * reference hands to Normalize; dv_ingest_clips consumes it. One CTA keeps a frame in shared memory (3*H*W*4 B <= 200 KB). */
int dv_frames_color_jitter(const uint8_t* clips_u8, float* out, const float* params, int B, int n_frames_per_sample, int H,
                           int W, void* stream);
/* A.GaussianBlur (utils/augmentation.py:706-721) on float32 planar clips [B][3][F][H][W] in [0, 1]: per frame ToPILImage
 * (x * 255 truncated to uint8) -> PIL's Gaussian (three extended-box passes per direction, BoxBlur.c) -> ToTensor (/ 255).
 * params[frame][4] = {apply, box radius, ww, fw} from dv_frames_gaussian_blur_params_host(sigma) (apply = 0: the frame is
 * copied unchanged - RandomApply skipped the stage). out must not alias clips. One CTA keeps a frame in shared memory. */
int dv_frames_gaussian_blur(const float* clips, float* out, const int32_t* params, int B, int n_frames_per_sample, int H,
                            int W, void* stream);
/* host-only: the fixed-point box parameters Pillow derives from a Gaussian sigma (sigma <= 0: {0,0,0,0}) */
int dv_frames_gaussian_blur_params_host(float sigma, int32_t* params4_host);
/* host-only (no GPU needed): the same stage on ONE float32 CHW frame in host memory, running the line filter the kernel
 * runs (one __host__ __device__ function) - lets the CPU test pin that code against Pillow */
int dv_frames_gaussian_blur_host(const float* frame_host, float* out_host, int H, int W, float sigma);
/* host-only (no GPU needed): the 22-bit fixed-point bicubic table of one axis, out_size rows of
 * [first input index, tap count, taps[ksize]] - what Pillow's precompute_coeffs + normalize_coeffs_8bpc produce */
int dv_frames_axis_table_host(int in_size, int out_size, int32_t* table_host, int capacity, int32_t* ksize_host);

/* ---- JPEG frame decoding (replaces PIL.Image.open of the reference loader, dataset/local_dataset.py:283-286) ----
 * Bit-identical to the IJG / libjpeg-turbo decoder with its default settings (islow IDCT, fancy upsampling), i.e. to Pillow.
 * Baseline sequential (SOF0), 8 bit, 1 or 3 components, 4:4:4 / 4:2:2 / 4:2:0, restart intervals; anything else is refused.
 * Host half (serial bit streams, no GPU needed): dv_jpeg_probe_host reads one file's header into
 * info8 = {width, height, components, hmax, vmax, mcus per row, mcu rows, 0} and the int16 coefficient count per frame;
 * dv_jpeg_huffman_decode_host entropy-decodes n files of that geometry on n_threads host threads into
 * coef_host [n][coef_stride] (quantised DCT coefficients, natural order, component after component; meant to be pinned
 * memory) and qt_host [n][3][64] (quantisation tables per component).
 * Device half: dv_jpeg_idct_rgb_u8 dequantises, runs the inverse DCT (planes_tmp: n * dv_jpeg_plane_bytes(info8) bytes),
 * upsamples the chroma planes and converts to RGB: rgb_out uint8 [n][height][width][3] - what dv_frames_scale_crop_u8 reads. */
int dv_jpeg_probe_host(const uint8_t* data, int64_t len, int32_t* info8, int64_t* coef_count);
int dv_jpeg_huffman_decode_host(const uint8_t* const* files, const int64_t* lens, int n, const int32_t* info8,
                                int16_t* coef_host, int64_t coef_stride, uint16_t* qt_host, int n_threads);
int64_t dv_jpeg_plane_bytes(const int32_t* info8);
int dv_jpeg_idct_rgb_u8(const int16_t* coef, const uint16_t* qt, uint8_t* planes_tmp, uint8_t* rgb_out, int n,
                        const int32_t* info8, int64_t coef_stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DUALVAR_B200_H_ */
