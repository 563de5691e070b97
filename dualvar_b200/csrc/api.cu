// extern "C" entry points declared in include/dualvar_b200.h. Thin argument checking + dispatch;
// the kernels live in the sibling .cu files.
#include "../../include/dualvar_b200.h"

#include "conv_tile.cuh"
#include "host_common.h"

namespace dv {
const std::string& last_error_ref();

int conv_fprop_bf16(const void* x, const void* w_packed, void* y, double* stats, const float* bias,
                    const ConvGeom& c, cudaStream_t stream, int y_f32 = 0, const float* xf_ss = nullptr, int xf_relu = 0);
int conv_dgrad_bf16(const void* dy, const void* w_packed_t, void* dx, const ConvGeom& c,
                    cudaStream_t stream, const BnReduce* red, int dx_f32 = 0);
int conv_fprop_f32planes(const void* x_planes, long long plane_stride, int K, const void* wf_all, float* y,
                         double* stats, const float* bias, const ConvGeom& c, cudaStream_t stream);
int conv_dgrad_f32planes(const void* dy_planes, long long plane_stride, int K, const void* wt_all, float* dx,
                         const ConvGeom& c, cudaStream_t stream);
int conv_wgrad_f32planes(const void* x_planes, const void* dy_planes, int K, float* dw, const ConvGeom& c,
                         cudaStream_t stream);
int conv_dgrad_stack_ok(const ConvGeom& c);
int conv_dgrad_stack_bf16(const void* dy, const void* w_stack, void* dx, const ConvGeom& c, cudaStream_t stream,
                          const BnReduce* red);
int conv_wgrad_bf16(const void* x, const void* dy, float* dw, const ConvGeom& c, cudaStream_t stream,
                    bool accumulate = false, const float* xf_ss = nullptr, int xf_relu = 0);
int pack_weights(const float* w, void* wf, void* wt, int Cout, int Cin, int taps, int Cout_p,
                 int Cin_p, cudaStream_t stream);
int unpack_wgrad(const float* dwp, float* dw, int Cout, int Cin, int taps, int Cin_p, float beta,
                 cudaStream_t stream);
int ncdhw_to_ndhwc(const float* x, void* y, int N, int C, int Cp, long long S, cudaStream_t stream);
int ndhwc_to_ncdhw(const void* y, float* x, int N, int C, int Cp, long long S, cudaStream_t stream);

int bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                float* running_var, float* ss, float* saved, int C, int Cp, double count, float eps,
                float momentum, int training, cudaStream_t stream);
int bn_apply(const void* y1, const float* ss1, const void* y2, const float* ss2, const void* res,
             void* out, long long rows, int Cp, int out_ld, int out_coff, int relu, cudaStream_t stream);
int bn_bwd_reduce(const void* dout, const void* dout2, const void* out, const void* y, const float* ss,
                  double* sums, long long rows, int Cp, int o_ld, int o_coff, int relu, cudaStream_t stream);
int bn_bwd_finalize(const double* sums_local, const double* sums_global, const float* gamma,
                    const float* saved, float* dgamma, float* dbeta, float* coef, int C, int Cp,
                    double count_global, float grad_beta, cudaStream_t stream);
int bn_bwd_apply(const void* dout, const void* dout2, const void* out, const void* y, const float* ss,
                 const float* coef, void* dy, void* g_out, long long rows, int Cp, int o_ld, int o_coff, int relu,
                 cudaStream_t stream);
int add_bf16(const void* a, const void* b, void* out, long long n, cudaStream_t stream);
struct PoolGeom {
  int N, T, H, W, To, Ho, Wo, Cp;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
};
int avgpool_fwd(const void* x, float* out, int N, int S, int C, int Cp, int ld_out, cudaStream_t stream);
int avgpool_bwd(const float* dout, void* dx, int N, int S, int C, int Cp, int ld_out, cudaStream_t stream);
int maxpool_fwd(const void* x, void* y, uint8_t* idx, const PoolGeom& g, cudaStream_t stream);
int maxpool_bwd_idx(const uint8_t* idx, const void* dy, void* dx, const PoolGeom& g, cudaStream_t stream);
int maxpool_bwd(const void* x, const void* y, const void* dy, void* dx, const PoolGeom& g, cudaStream_t stream);
int ingest(const void* src, int src_u8, void* dst, const int* perm, long long sb, long long sv, long long sc,
           long long st, int B, int C, int T, int H, int W, int view, int nv, int n_series,
           const float* mean, const float* stdv, int s2d, cudaStream_t stream, int n_planes = 1,
           long long plane_stride = 0);
int conv_stem_fprop_bf16(const void* x_s2d, const void* w_stem, void* y, double* stats, const float* bias,
                         int N, int T, int H2, int W2, int Cout_p, int kt, int pt, cudaStream_t stream,
                         int y_f32 = 0);
int conv_stem_wgrad_bf16(const void* x_s2d, const void* dy, float* dw, int N, int T, int H2, int W2,
                         int Cout_p, int kt, int pt, cudaStream_t stream, bool accumulate = false);
int frames_scale_crop_u8(const uint8_t* frames, uint8_t* tmp, uint8_t* out, const int* crop_lu, int B, int V, int T,
                         int Hs, int Ws, int scale_w, int scale_h, int crop_w, int crop_h, cudaStream_t stream);
int frames_axis_table_host(int in_size, int out_size, int* tab, int capacity, int* ksize);
void blur_params_host(float sigma, int* out4);
int frames_gaussian_blur_host(const float* frame, float* out, int H, int W, float sigma);
int frames_gaussian_blur(const float* in, float* out, const int* params, int B, int F, int H, int W, cudaStream_t stream);
int frames_color_jitter(const uint8_t* clips, float* out, const float* params, int B, int F, int H, int W,
                        cudaStream_t stream);
// fp32 mode (fp32_mode.cu)
int f32_max_planes();
int f32_split_planes(const float* src, float* dst, long long n, int K, cudaStream_t stream);
int f32_colstats(const float* y, double* stats, long long rows, int Cp, cudaStream_t stream);
int f32_bn_apply(const float* y1, const float* ss1, const float* y2, const float* ss2, const float* res, float* out,
                 void* planes, long long plane_stride, int K, long long rows, int Cp, int out_ld, int out_coff, int relu,
                 cudaStream_t stream);
int f32_bn_bwd_reduce(const float* dout, const float* dout2, const float* out, const float* y, const float* ss,
                      double* sums, long long rows, int Cp, int o_ld, int o_coff, int relu, cudaStream_t stream);
int f32_bn_bwd_apply(const float* dout, const float* dout2, const float* out, const float* y, const float* ss,
                     const float* coef, void* dy_planes, long long plane_stride, int K, float* g_out, long long rows,
                     int Cp, int o_ld, int o_coff, int relu, cudaStream_t stream);
int f32_slice_mean(const float* x, float* out, int N, int S, int C, int ld, int coff, cudaStream_t stream);
int f32_gate_scale(float* x, void* planes, long long plane_stride, int K, const float* w, int N, int S, int C, int ld,
                   int coff, cudaStream_t stream);
int f32_gate_bwd_reduce(const float* dout, const float* y, const float* ss, float* dw, int N, int S, int C, int Cp,
                        int ld, int coff, cudaStream_t stream);
int f32_gate_bwd_apply(const float* dout, const float* w, const float* dmean, float* dz, int N, int S, int C, int Cp,
                       int ld, int coff, cudaStream_t stream);
int f32_add(const float* x, const float* y, float* out, long long n, cudaStream_t stream);
int f32_split(const float* x, void* planes, long long plane_stride, int K, long long n, cudaStream_t stream);
int f32_avgpool_fwd(const float* x, float* out, int N, int S, int C, int Cp, int ld_out, cudaStream_t stream);
int f32_avgpool_bwd(const float* dout, float* dx, int N, int S, int C, int Cp, int ld_out, cudaStream_t stream);
int f32_maxpool_fwd(const float* x, float* y, uint8_t* idx, void* planes, long long plane_stride, int K,
                    const PoolGeom& g, cudaStream_t stream);
int f32_maxpool_bwd(const uint8_t* idx, const float* dy, float* dx, const PoolGeom& g, cudaStream_t stream);
int f32_ndhwc_to_ncdhw(const float* y, float* x, int N, int C, int Cp, long long S, cudaStream_t stream);
int f32_ncdhw_to_ndhwc(const float* x, float* y, int N, int C, int Cp, long long S, cudaStream_t stream);
int pack_stem_weights(const float* w, void* ws, int Cout, int Cin, int kt, int Cout_p, cudaStream_t stream);
int unpack_stem_wgrad(const float* dws, float* dw, int Cout, int Cin, int kt, float beta, cudaStream_t stream);
int sgemm(int ta, int tb, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
          int ldb, float beta, float* C, int ldc, const float* bias, int relu, cudaStream_t stream);
int colsum(const float* X, float* out, int M, int N, int ld, float beta, cudaStream_t stream);
int relu_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t stream);
int l2norm_fwd(const float* x, float* y, float* inv_norm, long long rows, int d, float eps, cudaStream_t stream);
int l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, long long rows, int d,
               cudaStream_t stream);
int sim_ce_blocks(int C);
int sim_ce_fwd(const float* a, int lda, const float* b, int ldb, int b_dmajor, int R, int C, int d, float* S, int ld_s,
               int col0, float* logits, int ld_logits, const int* self_col, const int* pos_col, float inv_T,
               float* partials, cudaStream_t stream);
int sim_ce_finish(float* S, int ld_s, int R, int C, int col0, const float* partials, float* logits, int ld_logits,
                  const int* self_col, const int* pos_col, float inv_T, float grad_scale, float* loss_sum, int* hits,
                  cudaStream_t stream);
int contrast_rows(float* S, float* logits, const int* self_col, const int* pos_col, int R, int C,
                  int ld_s, int ld_logits, float inv_T, float grad_scale, float* loss_sum, int* hits,
                  cudaStream_t stream);
int rank_loss(const float* a, const float* b, float* da, float* db, float* logits, float* loss_sum,
              int* hits, int B, int s, int e, float theta, float clip_max, float weight, cudaStream_t stream);
int permute_segments(const float* in, float* out, const int* perm, int B, int s, int e, int inverse,
                     cudaStream_t stream);
int segment_sum(const float* in, float* out, long long rows, int s, int e, float scale, cudaStream_t stream);
int segment_bcast(const float* in, float* out, long long rows, int s, int e, float scale, float beta,
                  cudaStream_t stream);
int rowdot(const float* a, const float* b, float* out, int rows, int d, int ld_out, cudaStream_t stream);
int row_axpy(const float* alpha, int ld_alpha, const float* x, float* y, int rows, int d, float beta,
             cudaStream_t stream);
long long small_allreduce_buffer_bytes();
int bn_finalize_sync(const double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float* ss, float* saved, int C, int Cp, double count_global, float eps, float momentum,
                     const long long* peer_ptrs, int rank, int world, long long seq, cudaStream_t stream);
int bn_bwd_finalize_sync(const double* sums_local, const float* gamma, const float* saved, float* dgamma, float* dbeta,
                         float* coef, int C, int Cp, double count_global, float grad_beta, const long long* peer_ptrs,
                         int rank, int world, long long seq, cudaStream_t stream);
int gate_fc_fwd(const float* mean, const float* W, const float* b, float* w, int N, int C, cudaStream_t st);
int gate_fc_bwd(const float* dw, const float* w, const float* mean, const float* W, float* dpre, float* gW, float* gb,
                float* dmean, int N, int C, cudaStream_t st);
int jpeg_probe_host(const uint8_t* data, long long len, int* info8, long long* coef_count);
int jpeg_huffman_decode_host(const uint8_t* const* files, const long long* lens, int n, const int* info8, int16_t* coef,
                             long long coef_stride, uint16_t* qt, int n_threads);
long long jpeg_plane_bytes(const int* info8);
int jpeg_idct_rgb_u8(const int16_t* coef, const uint16_t* qt, uint8_t* planes, uint8_t* rgb, int n, const int* info8,
                     long long coef_stride, cudaStream_t stream);
void comm_set_timeout(double seconds);
int comm_status(int* peer, long long* seq, int clear);
int small_allreduce_f64(double* inout, int n, const long long* peer_ptrs, int rank, int world, long long seq,
                        cudaStream_t stream);
int sgd_momentum_step(const long long* table, int n_chunks, float lr, float mu, float wd, int first,
                      cudaStream_t stream);
int retrieval_prepare(const float* x, double* mean, double* y, int n, int d, cudaStream_t st);
int retrieval_sim_topk(const double* test, const double* train, double* sim, float* sim32, long long* idx,
                       int nt, int ntr, int d, int k, cudaStream_t st);
int slice_mean(const void* x, float* out, int N, int S, int C, int ld, int coff, cudaStream_t st);
int gate_scale(void* x, const float* w, int N, int S, int C, int ld, int coff, cudaStream_t st);
int gate_bwd_reduce(const void* dout, const void* y, const float* ss, float* dw, int N, int S, int C, int Cp,
                    int ld, int coff, cudaStream_t st);
int gate_bwd_apply(const void* dout, const float* w, const float* dmean, void* dz, int N, int S, int C, int Cp,
                   int ld, int coff, cudaStream_t st);
int sigmoid_fwd(const float* x, float* y, long long n, cudaStream_t st);
int sigmoid_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t st);
int momentum_update(const long long* table, int n_chunks, float m, cudaStream_t stream);
int enqueue(const float* keys, float* queue, int B, int d, int K, int ptr, const long long* ptr_dev, cudaStream_t stream);
int advance_queue_ptr(long long* ptr, int batch, int K, cudaStream_t stream);
}  // namespace dv

using namespace dv;

static int check_geom(const dv_conv_geom* g) {
  DV_REQUIRE(g != nullptr, "geometry is NULL");
  DV_REQUIRE(g->N > 0 && g->T > 0 && g->H > 0 && g->W > 0, "empty input extent");
  DV_REQUIRE(g->Cin > 0 && g->Cout > 0, "empty channels");
  DV_REQUIRE(g->Cin_p % 8 == 0 && g->Cout_p % 8 == 0 && g->Cin_p >= g->Cin && g->Cout_p >= g->Cout,
             "padded channel counts must be multiples of 8 and >= logical counts");
  DV_REQUIRE(g->st >= 1 && g->st <= 2 && g->sh >= 1 && g->sh <= 2 && g->sw >= 1 && g->sw <= 2,
             "stride must be 1 or 2");
  DV_REQUIRE(g->To == (g->T + 2 * g->pt - g->kt) / g->st + 1 &&
                 g->Ho == (g->H + 2 * g->ph - g->kh) / g->sh + 1 &&
                 g->Wo == (g->W + 2 * g->pw - g->kw) / g->sw + 1,
             "output extent does not match floor((X+2p-k)/s)+1");
  DV_REQUIRE(g->To > 0 && g->Ho > 0 && g->Wo > 0, "empty output extent");
  return kOk;
}

template <class G>
static G to_geom(const dv_conv_geom* g) {
  G c;
  c.N = g->N; c.T = g->T; c.H = g->H; c.W = g->W; c.Cin_p = g->Cin_p;
  c.To = g->To; c.Ho = g->Ho; c.Wo = g->Wo; c.Cout_p = g->Cout_p;
  c.kt = g->kt; c.kh = g->kh; c.kw = g->kw;
  c.st = g->st; c.sh = g->sh; c.sw = g->sw;
  c.pt = g->pt; c.ph = g->ph; c.pw = g->pw;
  return c;
}

extern "C" {

const char* dv_last_error(void) { return last_error_ref().c_str(); }
int dv_version(void) { return 1; }
int64_t dv_launch_count(void) { return (int64_t)::dv::launch_counter(); }

int dv_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int dv_pack_conv_weight(const float* w, void* wf, void* wt, const dv_conv_geom* g, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(w != nullptr && (wf != nullptr || wt != nullptr), "NULL weight pointers");
  return pack_weights(w, wf, wt, g->Cout, g->Cin, g->kt * g->kh * g->kw, g->Cout_p, g->Cin_p,
                      (cudaStream_t)stream);
}

int dv_unpack_conv_wgrad(const float* dw_packed, float* grad, const dv_conv_geom* g, float beta,
                         void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(dw_packed != nullptr && grad != nullptr, "NULL gradient pointers");
  return unpack_wgrad(dw_packed, grad, g->Cout, g->Cin, g->kt * g->kh * g->kw, g->Cin_p, beta,
                      (cudaStream_t)stream);
}

int dv_ncdhw_to_ndhwc_bf16(const float* x, void* y, int N, int C, int Cp, int64_t S, void* stream) {
  DV_REQUIRE(x && y && N > 0 && C > 0 && Cp >= C && S > 0, "bad arguments");
  return ncdhw_to_ndhwc(x, y, N, C, Cp, S, (cudaStream_t)stream);
}

int dv_ndhwc_bf16_to_ncdhw(const void* y, float* x, int N, int C, int Cp, int64_t S, void* stream) {
  DV_REQUIRE(x && y && N > 0 && C > 0 && Cp >= C && S > 0, "bad arguments");
  return ndhwc_to_ncdhw(y, x, N, C, Cp, S, (cudaStream_t)stream);
}

int dv_conv3d_fprop_bf16(const void* x, const void* wf, void* y, double* bn_stats,
                         const float* bias_padded, const dv_conv_geom* g, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(x && wf && y, "NULL tensor pointer");
  return conv_fprop_bf16(x, wf, y, bn_stats, bias_padded, to_geom<ConvGeom>(g), (cudaStream_t)stream);
}

int dv_conv3d_fprop_bnrelu_bf16(const void* y_prev, const float* ss_prev, int relu, const void* wf, void* y,
                                double* bn_stats, const float* bias_padded, const dv_conv_geom* g, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(y_prev && ss_prev && wf && y, "NULL tensor pointer");
  return conv_fprop_bf16(y_prev, wf, y, bn_stats, bias_padded, to_geom<ConvGeom>(g), (cudaStream_t)stream, 0, ss_prev,
                         relu ? 1 : 0);
}

int dv_conv3d_wgrad_bnrelu_bf16(const void* y_prev, const float* ss_prev, int relu, const void* dy, float* dw_packed,
                                const dv_conv_geom* g, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(y_prev && ss_prev && dy && dw_packed, "NULL tensor pointer");
  return conv_wgrad_bf16(y_prev, dy, dw_packed, to_geom<ConvGeom>(g), (cudaStream_t)stream, false, ss_prev, relu ? 1 : 0);
}

int dv_conv3d_dgrad_bf16(const void* dy, const void* wt, void* dx, const dv_conv_geom* g,
                         void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(dy && wt && dx, "NULL tensor pointer");
  return conv_dgrad_bf16(dy, wt, dx, to_geom<ConvGeom>(g), (cudaStream_t)stream, nullptr);
}

int dv_conv3d_dgrad_bnred_bf16(const void* dy, const void* wt, void* dx, const dv_conv_geom* g,
                               const void* y_prev, const float* ss_prev, double* sums, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(dy && wt && dx && y_prev && sums, "NULL tensor pointer");
  const BnReduce red = {y_prev, ss_prev, sums};
  return conv_dgrad_bf16(dy, wt, dx, to_geom<ConvGeom>(g), (cudaStream_t)stream, &red);
}

int dv_conv3d_dgrad_stack_ok(const dv_conv_geom* g) {
  if (check_geom(g)) return 0;
  return conv_dgrad_stack_ok(to_geom<ConvGeom>(g));
}
int dv_conv3d_dgrad_stack_bf16(const void* dy, const void* w_stack, void* dx, const dv_conv_geom* g, const void* y_prev,
                               const float* ss_prev, double* sums, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(dy && w_stack && dx, "NULL tensor pointer");
  DV_REQUIRE((y_prev == nullptr) == (sums == nullptr), "y_prev and sums go together");
  const BnReduce red = {y_prev, ss_prev, sums};
  return conv_dgrad_stack_bf16(dy, w_stack, dx, to_geom<ConvGeom>(g), (cudaStream_t)stream, sums ? &red : nullptr);
}

int dv_conv3d_wgrad_bf16(const void* x, const void* dy, float* dw_packed, const dv_conv_geom* g,
                         void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(x && dy && dw_packed, "NULL tensor pointer");
  return conv_wgrad_bf16(x, dy, dw_packed, to_geom<ConvGeom>(g), (cudaStream_t)stream);
}

#define ST ((cudaStream_t)stream)
int dv_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float* scale_shift, float* saved, int C, int Cp, double count,
                   float eps, float momentum, int training, void* stream) {
  DV_REQUIRE(gamma && beta && scale_shift && C > 0 && Cp >= C && Cp % 8 == 0, "bad bn_finalize arguments");
  DV_REQUIRE(training ? (stats != nullptr && count > 0) : (running_mean && running_var),
             "bn_finalize: training needs stats, eval needs running statistics");
  return bn_finalize(stats, gamma, beta, running_mean, running_var, scale_shift, saved, C, Cp, count, eps,
                     momentum, training, ST);
}
int dv_bn_apply(const void* y1, const float* ss1, const void* y2, const float* ss2, const void* res,
                void* out, int64_t rows, int Cp, int out_ld, int out_coff, int relu, void* stream) {
  DV_REQUIRE(y1 && ss1 && out && rows > 0 && Cp % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0,
             "bad bn_apply arguments");
  DV_REQUIRE((y2 == nullptr) == (ss2 == nullptr), "bn_apply: y2 and ss2 go together");
  return bn_apply(y1, ss1, y2, ss2, res, out, rows, Cp, out_ld, out_coff, relu, ST);
}
int dv_bn_bwd_reduce(const void* dout, const void* dout2, const void* out, const void* y, const float* mask_ss,
                     double* sums, int64_t rows, int Cp, int o_ld, int o_coff, int relu, void* stream) {
  DV_REQUIRE(dout && y && sums && (!relu || out || mask_ss) && rows > 0 && Cp % 8 == 0,
             "bad bn_bwd_reduce arguments");
  return bn_bwd_reduce(dout, dout2, out, y, mask_ss, sums, rows, Cp, o_ld, o_coff, relu, ST);
}
int dv_bn_bwd_finalize(const double* sums_local, const double* sums_global, const float* gamma,
                       const float* saved, float* dgamma, float* dbeta, float* coef, int C, int Cp,
                       double count_global, float grad_beta, void* stream) {
  DV_REQUIRE(sums_local && sums_global && gamma && saved && coef && count_global > 0, "bad bn_bwd_finalize arguments");
  return bn_bwd_finalize(sums_local, sums_global, gamma, saved, dgamma, dbeta, coef, C, Cp, count_global,
                         grad_beta, ST);
}
int dv_bn_bwd_apply(const void* dout, const void* dout2, const void* out, const void* y, const float* mask_ss,
                    const float* coef, void* dy, void* g_out, int64_t rows, int Cp, int o_ld, int o_coff, int relu,
                    void* stream) {
  DV_REQUIRE(dout && y && coef && dy && (!relu || out || mask_ss) && rows > 0 && Cp % 8 == 0,
             "bad bn_bwd_apply arguments");
  return bn_bwd_apply(dout, dout2, out, y, mask_ss, coef, dy, g_out, rows, Cp, o_ld, o_coff, relu, ST);
}
int dv_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  DV_REQUIRE(a && b && out && n > 0 && n % 8 == 0, "bad add_bf16 arguments");
  return add_bf16(a, b, out, n, ST);
}
int dv_avgpool_fwd(const void* x, float* out, int N, int S, int C, int Cp, int ld_out, void* stream) {
  DV_REQUIRE(x && out && N > 0 && S > 0 && C > 0 && Cp >= C && ld_out >= C, "bad avgpool arguments");
  return avgpool_fwd(x, out, N, S, C, Cp, ld_out, ST);
}
int dv_avgpool_bwd(const float* dout, void* dx, int N, int S, int C, int Cp, int ld_out, void* stream) {
  DV_REQUIRE(dout && dx && N > 0 && S > 0 && C > 0 && Cp >= C && ld_out >= C, "bad avgpool arguments");
  return avgpool_bwd(dout, dx, N, S, C, Cp, ld_out, ST);
}
static PoolGeom to_pool(const dv_pool_geom* g) {
  PoolGeom p;
  p.N = g->N; p.T = g->T; p.H = g->H; p.W = g->W; p.To = g->To; p.Ho = g->Ho; p.Wo = g->Wo; p.Cp = g->Cp;
  p.kt = g->kt; p.kh = g->kh; p.kw = g->kw; p.st = g->st; p.sh = g->sh; p.sw = g->sw;
  p.pt = g->pt; p.ph = g->ph; p.pw = g->pw;
  return p;
}
int dv_maxpool3d_fwd(const void* x, void* y, const dv_pool_geom* g, void* stream) {
  DV_REQUIRE(x && y && g && g->Cp % 8 == 0, "bad maxpool arguments");
  return maxpool_fwd(x, y, nullptr, to_pool(g), ST);
}
int dv_maxpool3d_fwd_idx(const void* x, void* y, uint8_t* argmax, const dv_pool_geom* g, void* stream) {
  DV_REQUIRE(x && y && argmax && g && g->Cp % 8 == 0, "bad maxpool arguments");
  return maxpool_fwd(x, y, argmax, to_pool(g), ST);
}
int dv_maxpool3d_bwd_idx(const uint8_t* argmax, const void* dy, void* dx, const dv_pool_geom* g, void* stream) {
  DV_REQUIRE(argmax && dy && dx && g && g->Cp % 8 == 0, "bad maxpool arguments");
  return maxpool_bwd_idx(argmax, dy, dx, to_pool(g), ST);
}
int dv_maxpool3d_bwd(const void* x, const void* y, const void* dy, void* dx, const dv_pool_geom* g,
                     void* stream) {
  DV_REQUIRE(x && y && dy && dx && g && g->Cp % 8 == 0, "bad maxpool arguments");
  return maxpool_bwd(x, y, dy, dx, to_pool(g), ST);
}
int dv_ingest_clips(const float* src, void* dst, const int32_t* perm, int64_t sb, int64_t sv, int64_t sc,
                    int64_t st, int B, int C, int T, int H, int W, int view, int nv, int n_series,
                    const float* mean_host, const float* std_host, int s2d, void* stream) {
  DV_REQUIRE(!s2d || (H % 2 == 0 && W % 2 == 0), "space-to-depth ingest needs even H and W");
  DV_REQUIRE(src && dst && B > 0 && C > 0 && C <= 4 && T > 0 && H > 0 && W > 0 && nv > 0, "bad ingest arguments");
  DV_REQUIRE(perm == nullptr || (n_series > 0 && T % n_series == 0), "ingest: T must divide into n_series segments");
  return ingest(src, 0, dst, perm, sb, sv, sc, st, B, C, T, H, W, view, nv, n_series, mean_host, std_host, s2d, ST);
}
int dv_ingest_clips_u8(const uint8_t* src, void* dst, const int32_t* perm, int64_t sb, int64_t sv, int64_t sc,
                       int64_t st, int B, int C, int T, int H, int W, int view, int nv, int n_series,
                       const float* mean_host, const float* std_host, int s2d, void* stream) {
  DV_REQUIRE(!s2d || (H % 2 == 0 && W % 2 == 0), "space-to-depth ingest needs even H and W");
  DV_REQUIRE(src && dst && B > 0 && C > 0 && C <= 4 && T > 0 && H > 0 && W > 0 && nv > 0, "bad ingest arguments");
  DV_REQUIRE(perm == nullptr || (n_series > 0 && T % n_series == 0), "ingest: T must divide into n_series segments");
  DV_REQUIRE(!s2d || (sb % 2 == 0 && sv % 2 == 0 && sc % 2 == 0 && st % 2 == 0), "uint8 space-to-depth ingest needs even strides");
  return ingest(src, 1, dst, perm, sb, sv, sc, st, B, C, T, H, W, view, nv, n_series, mean_host, std_host, s2d, ST);
}
int dv_sgemm(int ta, int tb, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
             int ldb, float beta, float* C, int ldc, const float* bias, int relu, void* stream) {
  DV_REQUIRE(A && B && C && K > 0, "bad sgemm arguments");
  return sgemm(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, relu, ST);
}
int dv_colsum(const float* X, float* out, int M, int N, int ld, float beta, void* stream) {
  DV_REQUIRE(X && out && M > 0 && N > 0, "bad colsum arguments");
  return colsum(X, out, M, N, ld, beta, ST);
}
int dv_relu_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  DV_REQUIRE(dy && y && dx && n > 0, "bad relu_bwd arguments");
  return relu_bwd(dy, y, dx, n, ST);
}
int dv_l2norm_fwd(const float* x, float* y, float* inv_norm, int64_t rows, int d, float eps, void* stream) {
  DV_REQUIRE(x && y && rows > 0 && d > 0, "bad l2norm arguments");
  return l2norm_fwd(x, y, inv_norm, rows, d, eps, ST);
}
int dv_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int64_t rows, int d,
                  void* stream) {
  DV_REQUIRE(dy && y && inv_norm && dx && rows > 0 && d > 0, "bad l2norm arguments");
  return l2norm_bwd(dy, y, inv_norm, dx, rows, d, ST);
}
int dv_contrast_rows(float* S, float* logits, const int32_t* self_col, const int32_t* pos_col, int R,
                     int C, int ld_s, int ld_logits, float inv_T, float grad_scale, float* loss_sum,
                     int32_t* hits, void* stream) {
  DV_REQUIRE(S && pos_col && loss_sum && R > 0 && C > 1, "bad contrast_rows arguments");
  return contrast_rows(S, logits, self_col, pos_col, R, C, ld_s, ld_logits, inv_T, grad_scale, loss_sum,
                       hits, ST);
}
int dv_sim_ce_blocks(int C) { return sim_ce_blocks(C); }
int dv_sim_ce_fwd(const float* a, int lda, const float* b, int ldb, int b_dmajor, int R, int C, int d, float* S,
                  int ld_s, int col0, float* logits, int ld_logits, const int32_t* self_col, const int32_t* pos_col,
                  float inv_T, float* partials, void* stream) {
  DV_REQUIRE(a && b && S && pos_col && partials && R > 0 && C > 0 && d > 0 && col0 >= 0 && ld_s >= col0 + C,
             "bad sim_ce_fwd arguments");
  return sim_ce_fwd(a, lda, b, ldb, b_dmajor, R, C, d, S, ld_s, col0, logits, ld_logits, self_col, pos_col, inv_T,
                    partials, ST);
}
int dv_sim_ce_finish(float* S, int ld_s, int R, int C, int col0, const float* partials, float* logits, int ld_logits,
                     const int32_t* self_col, const int32_t* pos_col, float inv_T, float grad_scale, float* loss_sum,
                     int32_t* hits, void* stream) {
  DV_REQUIRE(S && pos_col && partials && loss_sum && R > 0 && C > 0 && col0 >= 0 && col0 + C > 1,
             "bad sim_ce_finish arguments");
  return sim_ce_finish(S, ld_s, R, C, col0, partials, logits, ld_logits, self_col, pos_col, inv_T, grad_scale,
                       loss_sum, hits, ST);
}
int dv_rank_loss(const float* a, const float* b, float* da, float* db, float* logits, float* loss_sum,
                 int32_t* hits, int B, int s, int e, float theta, float clip_max, float weight,
                 void* stream) {
  DV_REQUIRE(a && b && da && db && loss_sum && B > 0 && s > 1 && e > 0 && theta > 0, "bad rank_loss arguments");
  return rank_loss(a, b, da, db, logits, loss_sum, hits, B, s, e, theta, clip_max, weight, ST);
}
int dv_permute_segments(const float* in, float* out, const int32_t* perm, int B, int s, int e,
                        int inverse, void* stream) {
  DV_REQUIRE(in && out && perm && in != out, "bad permute_segments arguments");
  return permute_segments(in, out, perm, B, s, e, inverse, ST);
}
int dv_segment_sum(const float* in, float* out, int64_t rows, int s, int e, float scale, void* stream) {
  DV_REQUIRE(in && out && rows > 0, "bad segment_sum arguments");
  return segment_sum(in, out, rows, s, e, scale, ST);
}
int dv_segment_bcast(const float* in, float* out, int64_t rows, int s, int e, float scale, float beta,
                     void* stream) {
  DV_REQUIRE(in && out && rows > 0, "bad segment_bcast arguments");
  return segment_bcast(in, out, rows, s, e, scale, beta, ST);
}
int dv_rowdot(const float* a, const float* b, float* out, int rows, int d, int ld_out, void* stream) {
  DV_REQUIRE(a && b && out && rows > 0, "bad rowdot arguments");
  return rowdot(a, b, out, rows, d, ld_out, ST);
}
int dv_row_axpy(const float* alpha, int ld_alpha, const float* x, float* y, int rows, int d, float beta,
                void* stream) {
  DV_REQUIRE(alpha && x && y && rows > 0, "bad row_axpy arguments");
  return row_axpy(alpha, ld_alpha, x, y, rows, d, beta, ST);
}

static int check_stem(const dv_conv_geom* g) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(g->Cin <= 4 && g->kh == 7 && g->kw == 7 && g->sh == 2 && g->sw == 2 && g->st == 1 && g->ph == 3 &&
                 g->pw == 3 && g->H % 2 == 0 && g->W % 2 == 0,
             "stem path needs Cin<=4, 7x7 window, stride (1,2,2), padding (p,3,3), even H and W");
  return kOk;
}
int dv_pack_stem_weight(const float* w, void* ws, const dv_conv_geom* g, void* stream) {
  if (int rc = check_stem(g)) return rc;
  DV_REQUIRE(w && ws, "NULL weight pointers");
  return pack_stem_weights(w, ws, g->Cout, g->Cin, g->kt, g->Cout_p, ST);
}
int dv_unpack_stem_wgrad(const float* dws, float* grad, const dv_conv_geom* g, float beta, void* stream) {
  if (int rc = check_stem(g)) return rc;
  DV_REQUIRE(dws && grad, "NULL gradient pointers");
  return unpack_stem_wgrad(dws, grad, g->Cout, g->Cin, g->kt, beta, ST);
}
int dv_conv3d_stem_fprop_bf16(const void* x_s2d, const void* ws, void* y, double* bn_stats,
                              const float* bias_padded, const dv_conv_geom* g, void* stream) {
  if (int rc = check_stem(g)) return rc;
  DV_REQUIRE(x_s2d && ws && y, "NULL tensor pointer");
  return conv_stem_fprop_bf16(x_s2d, ws, y, bn_stats, bias_padded, g->N, g->T, g->H / 2, g->W / 2, g->Cout_p,
                              g->kt, g->pt, ST);
}
int dv_conv3d_stem_wgrad_bf16(const void* x_s2d, const void* dy, float* dws, const dv_conv_geom* g, void* stream) {
  if (int rc = check_stem(g)) return rc;
  DV_REQUIRE(x_s2d && dy && dws, "NULL tensor pointer");
  return conv_stem_wgrad_bf16(x_s2d, dy, dws, g->N, g->T, g->H / 2, g->W / 2, g->Cout_p, g->kt, g->pt, ST);
}

/* ---- fp32 mode ------------------------------------------------------------------------------------------ */
static int check_planes(int n_planes, int64_t plane_stride, int64_t plane_elems) {
  DV_REQUIRE(n_planes >= 1 && n_planes <= f32_max_planes(), "n_planes must be 1..%d", f32_max_planes());
  DV_REQUIRE(n_planes == 1 || (plane_stride >= plane_elems && plane_stride % 8 == 0),
             "plane_stride must cover one plane and keep 16-byte alignment");
  return kOk;
}
int dv_f32_split_planes(const float* src, float* dst_planes, int64_t n, int n_planes, void* stream) {
  DV_REQUIRE(src && dst_planes && n > 0 && n_planes >= 1 && n_planes <= f32_max_planes(), "bad f32_split_planes arguments");
  return f32_split_planes(src, dst_planes, n, n_planes, ST);
}
int dv_conv3d_fprop_f32acc(const void* x_plane, const void* wf_plane, float* y, const float* bias_padded,
                           const dv_conv_geom* g, int accumulate, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(x_plane && wf_plane && y, "NULL tensor pointer");
  return conv_fprop_bf16(x_plane, wf_plane, y, nullptr, bias_padded, to_geom<ConvGeom>(g), ST, accumulate ? 1 : 2);
}
int dv_conv3d_dgrad_f32acc(const void* dy_plane, const void* wt_plane, float* dx, const dv_conv_geom* g, int accumulate,
                           void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(dy_plane && wt_plane && dx, "NULL tensor pointer");
  return conv_dgrad_bf16(dy_plane, wt_plane, dx, to_geom<ConvGeom>(g), ST, nullptr, accumulate ? 1 : 2);
}
int dv_conv3d_fprop_f32planes(const void* x_planes, int64_t plane_stride, int n_planes, const void* wf_all, float* y,
                              double* stats, const float* bias_padded, const dv_conv_geom* g, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(x_planes && wf_all && y && plane_stride > 0, "bad fprop_f32planes arguments");
  return conv_fprop_f32planes(x_planes, plane_stride, n_planes, wf_all, y, stats, bias_padded, to_geom<ConvGeom>(g), ST);
}
int dv_conv3d_wgrad_f32planes(const void* x_planes, const void* dy_planes, int n_planes, float* dw_packed,
                              const dv_conv_geom* g, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(x_planes && dy_planes && dw_packed, "NULL tensor pointer");
  return conv_wgrad_f32planes(x_planes, dy_planes, n_planes, dw_packed, to_geom<ConvGeom>(g), ST);
}
int dv_conv3d_dgrad_f32planes(const void* dy_planes, int64_t plane_stride, int n_planes, const void* wt_all, float* dx,
                              const dv_conv_geom* g, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(dy_planes && wt_all && dx && plane_stride > 0, "bad dgrad_f32planes arguments");
  return conv_dgrad_f32planes(dy_planes, plane_stride, n_planes, wt_all, dx, to_geom<ConvGeom>(g), ST);
}
int dv_conv3d_wgrad_bf16_acc(const void* x_plane, const void* dy_plane, float* dw_packed, const dv_conv_geom* g,
                             void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(x_plane && dy_plane && dw_packed, "NULL tensor pointer");
  return conv_wgrad_bf16(x_plane, dy_plane, dw_packed, to_geom<ConvGeom>(g), ST, true);
}
int dv_conv3d_stem_fprop_f32acc(const void* x_s2d_plane, const void* ws_plane, float* y, const float* bias_padded,
                                const dv_conv_geom* g, int accumulate, void* stream) {
  if (int rc = check_stem(g)) return rc;
  DV_REQUIRE(x_s2d_plane && ws_plane && y, "NULL tensor pointer");
  return conv_stem_fprop_bf16(x_s2d_plane, ws_plane, y, nullptr, bias_padded, g->N, g->T, g->H / 2, g->W / 2, g->Cout_p,
                              g->kt, g->pt, ST, accumulate ? 1 : 2);
}
int dv_conv3d_stem_wgrad_bf16_acc(const void* x_s2d_plane, const void* dy_plane, float* dws, const dv_conv_geom* g,
                                  void* stream) {
  if (int rc = check_stem(g)) return rc;
  DV_REQUIRE(x_s2d_plane && dy_plane && dws, "NULL tensor pointer");
  return conv_stem_wgrad_bf16(x_s2d_plane, dy_plane, dws, g->N, g->T, g->H / 2, g->W / 2, g->Cout_p, g->kt, g->pt, ST,
                              true);
}
int dv_f32_colstats(const float* y, double* stats, int64_t rows, int Cp, void* stream) {
  DV_REQUIRE(y && stats && rows > 0 && Cp > 0 && Cp % 8 == 0, "bad f32_colstats arguments");
  return f32_colstats(y, stats, rows, Cp, ST);
}
int dv_f32_bn_apply(const float* y1, const float* ss1, const float* y2, const float* ss2, const float* res, float* out,
                    void* out_planes, int64_t plane_stride, int n_planes, int64_t rows, int Cp, int out_ld, int out_coff,
                    int relu, void* stream) {
  DV_REQUIRE(y1 && ss1 && out && rows > 0 && Cp > 0 && Cp % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0 &&
                 out_ld >= out_coff + Cp,
             "bad f32_bn_apply arguments");
  DV_REQUIRE((y2 == nullptr) == (ss2 == nullptr), "f32_bn_apply: y2 and ss2 go together");
  if (out_planes != nullptr)
    if (int rc = check_planes(n_planes, plane_stride, rows * out_ld)) return rc;
  return f32_bn_apply(y1, ss1, y2, ss2, res, out, out_planes, plane_stride, n_planes, rows, Cp, out_ld, out_coff, relu,
                      ST);
}
int dv_f32_bn_bwd_reduce(const float* dout, const float* dout2, const float* out, const float* y, const float* mask_ss,
                         double* sums, int64_t rows, int Cp, int o_ld, int o_coff, int relu, void* stream) {
  DV_REQUIRE(dout && y && sums && (!relu || out || mask_ss) && rows > 0 && Cp > 0 && Cp % 8 == 0 && o_ld % 4 == 0 &&
                 o_coff % 4 == 0 && o_ld >= o_coff + Cp,
             "bad f32_bn_bwd_reduce arguments");
  return f32_bn_bwd_reduce(dout, dout2, out, y, mask_ss, sums, rows, Cp, o_ld, o_coff, relu, ST);
}
int dv_f32_bn_bwd_apply(const float* dout, const float* dout2, const float* out, const float* y, const float* mask_ss,
                        const float* coef, void* dy_planes, int64_t plane_stride, int n_planes, float* g_out,
                        int64_t rows, int Cp, int o_ld, int o_coff, int relu, void* stream) {
  DV_REQUIRE(dout && y && coef && dy_planes && (!relu || out || mask_ss) && rows > 0 && Cp > 0 && Cp % 8 == 0 &&
                 o_ld % 4 == 0 && o_coff % 4 == 0 && o_ld >= o_coff + Cp,
             "bad f32_bn_bwd_apply arguments");
  if (int rc = check_planes(n_planes, plane_stride, rows * Cp)) return rc;
  return f32_bn_bwd_apply(dout, dout2, out, y, mask_ss, coef, dy_planes, plane_stride, n_planes, g_out, rows, Cp, o_ld,
                          o_coff, relu, ST);
}
int dv_f32_slice_mean(const float* x, float* out, int N, int S, int C, int ld, int coff, void* stream) {
  DV_REQUIRE(x && out && N > 0 && S > 0 && C > 0 && ld >= coff + C, "bad f32_slice_mean arguments");
  return f32_slice_mean(x, out, N, S, C, ld, coff, ST);
}
int dv_f32_gate_scale(float* x, void* x_planes, int64_t plane_stride, int n_planes, const float* w, int N, int S, int C,
                      int ld, int coff, void* stream) {
  DV_REQUIRE(x && x_planes && w && N > 0 && S > 0 && C > 0 && C % 8 == 0 && ld % 8 == 0 && coff % 8 == 0 &&
                 ld >= coff + C,
             "bad f32_gate_scale arguments");
  if (int rc = check_planes(n_planes, plane_stride, (int64_t)N * S * ld)) return rc;
  return f32_gate_scale(x, x_planes, plane_stride, n_planes, w, N, S, C, ld, coff, ST);
}
int dv_f32_gate_bwd_reduce(const float* dout, const float* y, const float* ss, float* dw, int N, int S, int C, int Cp,
                           int ld, int coff, void* stream) {
  DV_REQUIRE(dout && y && ss && dw && N > 0 && S > 0 && C > 0 && Cp >= C, "bad f32_gate_bwd_reduce arguments");
  return f32_gate_bwd_reduce(dout, y, ss, dw, N, S, C, Cp, ld, coff, ST);
}
int dv_f32_gate_bwd_apply(const float* dout, const float* w, const float* dmean, float* dz, int N, int S, int C, int Cp,
                          int ld, int coff, void* stream) {
  DV_REQUIRE(dout && w && dmean && dz && N > 0 && S > 0 && C > 0 && Cp >= C, "bad f32_gate_bwd_apply arguments");
  return f32_gate_bwd_apply(dout, w, dmean, dz, N, S, C, Cp, ld, coff, ST);
}
int dv_f32_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  DV_REQUIRE(a && b && out && n > 0 && n % 4 == 0, "bad f32_add arguments");
  return f32_add(a, b, out, n, ST);
}
int dv_f32_split(const float* x, void* planes, int64_t plane_stride, int n_planes, int64_t n, void* stream) {
  DV_REQUIRE(x && planes && n > 0 && n % 8 == 0, "bad f32_split arguments");
  if (int rc = check_planes(n_planes, plane_stride, n)) return rc;
  return f32_split(x, planes, plane_stride, n_planes, n, ST);
}
int dv_f32_avgpool_fwd(const float* x, float* out, int N, int S, int C, int Cp, int ld_out, void* stream) {
  DV_REQUIRE(x && out && N > 0 && S > 0 && C > 0 && Cp >= C && ld_out >= C, "bad avgpool arguments");
  return f32_avgpool_fwd(x, out, N, S, C, Cp, ld_out, ST);
}
int dv_f32_avgpool_bwd(const float* dout, float* dx, int N, int S, int C, int Cp, int ld_out, void* stream) {
  DV_REQUIRE(dout && dx && N > 0 && S > 0 && C > 0 && Cp >= C && ld_out >= C, "bad avgpool arguments");
  return f32_avgpool_bwd(dout, dx, N, S, C, Cp, ld_out, ST);
}
int dv_f32_maxpool3d_fwd(const float* x, float* y, uint8_t* argmax, void* y_planes, int64_t plane_stride, int n_planes,
                         const dv_pool_geom* g, void* stream) {
  DV_REQUIRE(x && y && g && g->Cp % 8 == 0, "bad maxpool arguments");
  if (y_planes != nullptr)
    if (int rc = check_planes(n_planes, plane_stride, (int64_t)g->N * g->To * g->Ho * g->Wo * g->Cp)) return rc;
  return f32_maxpool_fwd(x, y, argmax, y_planes, plane_stride, n_planes, to_pool(g), ST);
}
int dv_f32_maxpool3d_bwd(const uint8_t* argmax, const float* dy, float* dx, const dv_pool_geom* g, void* stream) {
  DV_REQUIRE(argmax && dy && dx && g && g->Cp % 8 == 0, "bad maxpool arguments");
  return f32_maxpool_bwd(argmax, dy, dx, to_pool(g), ST);
}
int dv_f32_ndhwc_to_ncdhw(const float* y, float* x, int N, int C, int Cp, int64_t S, void* stream) {
  DV_REQUIRE(x && y && N > 0 && C > 0 && Cp >= C && S > 0, "bad arguments");
  return f32_ndhwc_to_ncdhw(y, x, N, C, Cp, S, ST);
}
int dv_f32_ncdhw_to_ndhwc(const float* x, float* y, int N, int C, int Cp, int64_t S, void* stream) {
  DV_REQUIRE(x && y && N > 0 && C > 0 && Cp >= C && S > 0, "bad arguments");
  return f32_ncdhw_to_ndhwc(x, y, N, C, Cp, S, ST);
}
int dv_ingest_clips_planes(const void* src, int src_is_u8, void* dst_planes, int64_t plane_stride, int n_planes,
                           const int32_t* perm, int64_t sb, int64_t sv, int64_t sc, int64_t st, int B, int C, int T,
                           int H, int W, int view, int nv, int n_series, const float* mean_host, const float* std_host,
                           int s2d, void* stream) {
  DV_REQUIRE(!s2d || (H % 2 == 0 && W % 2 == 0), "space-to-depth ingest needs even H and W");
  DV_REQUIRE(src && dst_planes && B > 0 && C > 0 && C <= 4 && T > 0 && H > 0 && W > 0 && nv > 0, "bad ingest arguments");
  DV_REQUIRE(perm == nullptr || (n_series > 0 && T % n_series == 0), "ingest: T must divide into n_series segments");
  DV_REQUIRE(!(s2d && src_is_u8) || (sb % 2 == 0 && sv % 2 == 0 && sc % 2 == 0 && st % 2 == 0),
             "uint8 space-to-depth ingest needs even strides");
  const int64_t plane = s2d ? (int64_t)B * nv * T * (H / 2) * (W / 2 + 3) * 16 : (int64_t)B * nv * T * H * W * 8;
  if (int rc = check_planes(n_planes, plane_stride, plane)) return rc;
  return ingest(src, src_is_u8 ? 1 : 0, dst_planes, perm, sb, sv, sc, st, B, C, T, H, W, view, nv, n_series, mean_host,
                std_host, s2d, ST, n_planes, plane_stride);
}

int dv_frames_scale_crop_u8(const uint8_t* frames, uint8_t* tmp, uint8_t* out, const int32_t* crop_lu, int B, int n_views,
                            int T, int Hs, int Ws, int scale_w, int scale_h, int crop_w, int crop_h, void* stream) {
  DV_REQUIRE(frames && tmp && out && crop_lu, "NULL pointer");
  DV_REQUIRE(B > 0 && n_views > 0 && T > 0 && Hs > 0 && Ws > 0, "empty frame batch");
  DV_REQUIRE(scale_w > 0 && scale_h > 0 && crop_w > 0 && crop_h > 0 && crop_w <= scale_w && crop_h <= scale_h,
             "crop %dx%d does not fit the scaled frame %dx%d", crop_w, crop_h, scale_w, scale_h);
  return frames_scale_crop_u8(frames, tmp, out, crop_lu, B, n_views, T, Hs, Ws, scale_w, scale_h, crop_w, crop_h, ST);
}

int dv_frames_color_jitter(const uint8_t* clips_u8, float* out, const float* params, int B, int n_frames_per_sample, int H,
                           int W, void* stream) {
  DV_REQUIRE(clips_u8 && out && params, "NULL pointer");
  DV_REQUIRE(B > 0 && n_frames_per_sample > 0 && H > 0 && W > 0, "empty clip batch");
  return frames_color_jitter(clips_u8, out, params, B, n_frames_per_sample, H, W, ST);
}

int dv_frames_gaussian_blur(const float* clips, float* out, const int32_t* params, int B, int n_frames_per_sample, int H,
                            int W, void* stream) {
  DV_REQUIRE(clips && out && params && clips != out, "NULL or aliased pointers");
  DV_REQUIRE(B > 0 && n_frames_per_sample > 0 && H > 0 && W > 0, "empty clip batch");
  return frames_gaussian_blur(clips, out, params, B, n_frames_per_sample, H, W, ST);
}
int dv_frames_gaussian_blur_params_host(float sigma, int32_t* params4_host) {
  DV_REQUIRE(params4_host != nullptr, "NULL pointer");
  blur_params_host(sigma, params4_host);
  return kOk;
}
int dv_frames_gaussian_blur_host(const float* frame_host, float* out_host, int H, int W, float sigma) {
  DV_REQUIRE(frame_host && out_host && H > 0 && W > 0, "bad gaussian_blur_host arguments");
  return frames_gaussian_blur_host(frame_host, out_host, H, W, sigma);
}

int dv_frames_axis_table_host(int in_size, int out_size, int32_t* table_host, int capacity, int32_t* ksize_host) {
  DV_REQUIRE(in_size > 0 && out_size > 0 && table_host && ksize_host && capacity > 0, "bad axis_table arguments");
  return frames_axis_table_host(in_size, out_size, table_host, capacity, ksize_host);
}

int dv_moco_momentum_update(const int64_t* chunk_table, int n_chunks, float m, void* stream) {
  DV_REQUIRE(chunk_table && n_chunks >= 0 && m >= 0.f && m <= 1.f, "bad momentum_update arguments");
  return momentum_update(reinterpret_cast<const long long*>(chunk_table), n_chunks, m, ST);
}
int dv_gate_fc_fwd(const float* mean, const float* fc_weight, const float* fc_bias, float* w, int N, int C, void* stream) {
  DV_REQUIRE(mean && fc_weight && fc_bias && w && N > 0 && C > 0, "bad gate_fc_fwd arguments");
  return gate_fc_fwd(mean, fc_weight, fc_bias, w, N, C, ST);
}
int dv_gate_fc_bwd(const float* dw, const float* w, const float* mean, const float* fc_weight, float* dpre, float* grad_weight,
                   float* grad_bias, float* dmean, int N, int C, void* stream) {
  DV_REQUIRE(dw && w && mean && fc_weight && dpre && grad_weight && grad_bias && dmean && N > 0 && C > 0, "bad gate_fc_bwd arguments");
  return gate_fc_bwd(dw, w, mean, fc_weight, dpre, grad_weight, grad_bias, dmean, N, C, ST);
}
int dv_moco_enqueue_at(const float* keys, float* queue, int B, int d, int K, const int64_t* queue_ptr, void* stream) {
  DV_REQUIRE(keys && queue && queue_ptr && B > 0 && d > 0 && K > 0, "bad enqueue arguments");
  DV_REQUIRE(K % B == 0, "queue size K=%d must be a multiple of the global batch %d", K, B);
  return enqueue(keys, queue, B, d, K, 0, reinterpret_cast<const long long*>(queue_ptr), ST);
}
int dv_moco_advance_ptr(int64_t* queue_ptr, int batch, int K, void* stream) {
  DV_REQUIRE(queue_ptr && batch > 0 && K > 0 && K % batch == 0, "bad advance_ptr arguments");
  return advance_queue_ptr(reinterpret_cast<long long*>(queue_ptr), batch, K, ST);
}
int dv_moco_enqueue(const float* keys, float* queue, int B, int d, int K, int ptr, void* stream) {
  DV_REQUIRE(keys && queue && B > 0 && d > 0 && K > 0, "bad enqueue arguments");
  DV_REQUIRE(K % B == 0, "queue size K=%d must be a multiple of the global batch %d", K, B);
  DV_REQUIRE(ptr >= 0 && ptr + B <= K, "queue pointer %d out of range", ptr);
  return enqueue(keys, queue, B, d, K, ptr, nullptr, ST);
}

int dv_slice_mean(const void* x, float* out, int N, int S, int C, int ld, int coff, void* stream) {
  DV_REQUIRE(x && out && N > 0 && S > 0 && C > 0 && ld >= coff + C, "bad slice_mean arguments");
  return slice_mean(x, out, N, S, C, ld, coff, ST);
}
int dv_gate_scale(void* x, const float* w, int N, int S, int C, int ld, int coff, void* stream) {
  DV_REQUIRE(x && w && N > 0 && S > 0 && C > 0 && ld >= coff + C, "bad gate_scale arguments");
  return gate_scale(x, w, N, S, C, ld, coff, ST);
}
int dv_gate_bwd_reduce(const void* dout, const void* y, const float* ss, float* dw, int N, int S, int C, int Cp,
                       int ld, int coff, void* stream) {
  DV_REQUIRE(dout && y && ss && dw && N > 0 && S > 0 && C > 0 && Cp >= C, "bad gate_bwd_reduce arguments");
  return gate_bwd_reduce(dout, y, ss, dw, N, S, C, Cp, ld, coff, ST);
}
int dv_gate_bwd_apply(const void* dout, const float* w, const float* dmean, void* dz, int N, int S, int C, int Cp,
                      int ld, int coff, void* stream) {
  DV_REQUIRE(dout && w && dmean && dz && N > 0 && S > 0 && C > 0 && Cp >= C, "bad gate_bwd_apply arguments");
  return gate_bwd_apply(dout, w, dmean, dz, N, S, C, Cp, ld, coff, ST);
}
int dv_sigmoid_fwd(const float* x, float* y, int64_t n, void* stream) {
  DV_REQUIRE(x && y && n > 0, "bad sigmoid arguments");
  return sigmoid_fwd(x, y, n, ST);
}
int dv_sigmoid_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  DV_REQUIRE(dy && y && dx && n > 0, "bad sigmoid arguments");
  return sigmoid_bwd(dy, y, dx, n, ST);
}

int dv_retrieval_prepare(const float* feat, double* mean, double* out, int n, int d, void* stream) {
  DV_REQUIRE(feat && mean && out && n > 0 && d > 0, "bad retrieval_prepare arguments");
  return retrieval_prepare(feat, mean, out, n, d, ST);
}
int dv_retrieval_sim_topk(const double* test, const double* train, double* sim, float* sim32, int64_t* idx,
                          int n_test, int n_train, int d, int k, void* stream) {
  DV_REQUIRE(test && train && sim && idx && n_test > 0 && n_train > 0 && d > 0 && k > 0 && k <= n_train,
             "bad retrieval_sim_topk arguments");
  return retrieval_sim_topk(test, train, sim, sim32, reinterpret_cast<long long*>(idx), n_test, n_train, d, k, ST);
}

int dv_sgd_momentum_step(const int64_t* chunk_table, int n_chunks, float lr, float momentum, float weight_decay,
                         int first_step, void* stream) {
  DV_REQUIRE(chunk_table && n_chunks >= 0 && lr >= 0.f, "bad sgd_momentum_step arguments");
  return sgd_momentum_step(reinterpret_cast<const long long*>(chunk_table), n_chunks, lr, momentum, weight_decay,
                           first_step, ST);
}

int64_t dv_allreduce_small_buffer_bytes(void) { return small_allreduce_buffer_bytes(); }

int dv_allreduce_small_f64(double* inout, int n, const int64_t* peer_buffers, int rank, int world, int64_t seq,
                           void* stream) {
  DV_REQUIRE(inout && peer_buffers, "NULL pointer");
  return small_allreduce_f64(inout, n, reinterpret_cast<const long long*>(peer_buffers), rank, world, seq, ST);
}

int dv_bn_finalize_sync(const double* stats, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, float* scale_shift, float* saved, int C, int Cp, double count_global,
                        float eps, float momentum, const int64_t* peer_buffers, int rank, int world, int64_t seq,
                        void* stream) {
  DV_REQUIRE(stats && gamma && beta && scale_shift && saved && peer_buffers && C > 0 && Cp >= C && Cp % 8 == 0,
             "bad bn_finalize_sync arguments");
  return bn_finalize_sync(stats, gamma, beta, running_mean, running_var, scale_shift, saved, C, Cp, count_global, eps,
                          momentum, reinterpret_cast<const long long*>(peer_buffers), rank, world, seq, ST);
}

int dv_bn_bwd_finalize_sync(const double* sums_local, const float* gamma, const float* saved, float* dgamma,
                            float* dbeta, float* coef, int C, int Cp, double count_global, float grad_beta,
                            const int64_t* peer_buffers, int rank, int world, int64_t seq, void* stream) {
  DV_REQUIRE(sums_local && gamma && saved && coef && peer_buffers && count_global > 0, "bad bn_bwd_finalize_sync arguments");
  return bn_bwd_finalize_sync(sums_local, gamma, saved, dgamma, dbeta, coef, C, Cp, count_global, grad_beta,
                              reinterpret_cast<const long long*>(peer_buffers), rank, world, seq, ST);
}

int dv_jpeg_probe_host(const uint8_t* data, int64_t len, int32_t* info8, int64_t* coef_count) {
  DV_REQUIRE(data && len > 0 && info8 && coef_count, "bad jpeg_probe arguments");
  long long cc = 0;
  const int rc = jpeg_probe_host(data, len, info8, &cc);
  *coef_count = cc;
  return rc;
}

int dv_jpeg_huffman_decode_host(const uint8_t* const* files, const int64_t* lens, int n, const int32_t* info8,
                                int16_t* coef_host, int64_t coef_stride, uint16_t* qt_host, int n_threads) {
  DV_REQUIRE(files && lens && n > 0 && info8 && coef_host && qt_host && coef_stride > 0, "bad jpeg_huffman_decode arguments");
  return jpeg_huffman_decode_host(files, reinterpret_cast<const long long*>(lens), n, info8, coef_host, coef_stride, qt_host,
                                  n_threads);
}

int64_t dv_jpeg_plane_bytes(const int32_t* info8) { return info8 ? jpeg_plane_bytes(info8) : 0; }

int dv_jpeg_idct_rgb_u8(const int16_t* coef, const uint16_t* qt, uint8_t* planes_tmp, uint8_t* rgb_out, int n,
                        const int32_t* info8, int64_t coef_stride, void* stream) {
  DV_REQUIRE(coef && qt && planes_tmp && rgb_out && n > 0 && info8 && coef_stride > 0, "bad jpeg_idct_rgb arguments");
  DV_REQUIRE(info8[0] > 0 && info8[1] > 0 && (info8[2] == 1 || info8[2] == 3), "bad jpeg geometry");
  return jpeg_idct_rgb_u8(coef, qt, planes_tmp, rgb_out, n, info8, coef_stride, ST);
}

int dv_comm_set_timeout(double seconds) {
  comm_set_timeout(seconds);
  return 0;
}

int dv_comm_status(int* peer, int64_t* seq, int clear) {
  long long s = 0;
  const int code = comm_status(peer, &s, clear);
  if (seq) *seq = s;
  return code;
}


}  // extern "C"
