// extern "C" entry points declared in include/dualvar_b200.h. Thin argument checking + dispatch;
// the kernels live in the sibling .cu files.
#include "../../include/dualvar_b200.h"

#include "conv_tile.cuh"
#include "host_common.h"

namespace dv {
const std::string& last_error_ref();

int conv_fprop_bf16(const void* x, const void* w_packed, void* y, double* stats, const float* bias,
                    const ConvGeom& c, cudaStream_t stream);
int conv_dgrad_bf16(const void* dy, const void* w_packed_t, void* dx, const ConvGeom& c,
                    cudaStream_t stream);
int conv_wgrad_bf16(const void* x, const void* dy, float* dw, const ConvGeom& c, cudaStream_t stream);
int pack_weights(const float* w, void* wf, void* wt, int Cout, int Cin, int taps, int Cout_p,
                 int Cin_p, cudaStream_t stream);
int unpack_wgrad(const float* dwp, float* dw, int Cout, int Cin, int taps, int Cin_p, float beta,
                 cudaStream_t stream);
int ncdhw_to_ndhwc(const float* x, void* y, int N, int C, int Cp, long long S, cudaStream_t stream);
int ndhwc_to_ncdhw(const void* y, float* x, int N, int C, int Cp, long long S, cudaStream_t stream);
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

int dv_conv3d_dgrad_bf16(const void* dy, const void* wt, void* dx, const dv_conv_geom* g,
                         void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(dy && wt && dx, "NULL tensor pointer");
  return conv_dgrad_bf16(dy, wt, dx, to_geom<ConvGeom>(g), (cudaStream_t)stream);
}

int dv_conv3d_wgrad_bf16(const void* x, const void* dy, float* dw_packed, const dv_conv_geom* g,
                         void* stream) {
  if (int rc = check_geom(g)) return rc;
  DV_REQUIRE(x && dy && dw_packed, "NULL tensor pointer");
  return conv_wgrad_bf16(x, dy, dw_packed, to_geom<ConvGeom>(g), (cudaStream_t)stream);
}

}  // extern "C"
