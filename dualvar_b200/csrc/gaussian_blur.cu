// Gaussian blur of the reference loader (third stage): A.GaussianBlur of utils/augmentation.py:706-721 -
// per clip a sigma in [0.1, 2], every frame goes ToPILImage (x * 255 truncated to uint8) -> PIL ImageFilter.GaussianBlur
// -> ToTensor (/ 255). The arithmetic lives in Pillow (third-party, 12.2.0 in this image; src/libImaging/BoxBlur.c):
// the Gaussian is approximated by three passes of an "extended box" filter per direction - a running uint32 sum over
// 2r+1 pixels weighted ww plus the two pixels beyond it weighted fw, 24-bit fixed point, rounded - horizontally, then on
// the transposed image. Restated here from the published algorithm; the per-line filter is ONE __host__ __device__
// function, so the exact code the kernel runs is also checked on the build box (no GPU) against Pillow through
// dv_frames_gaussian_blur_host.
//
// One CTA owns one frame (3 x H x W bytes in shared memory, two buffers); a thread owns one line (row or column of one
// channel) of a pass. HBM traffic: 4 B read + 4 B written per value, once.
#include <math.h>

#include <algorithm>
#include <vector>

#include "host_common.h"

namespace dv {

namespace {

// BoxBlur.c _gaussian_blur_radius: float variables, double literals (the C expression types are kept)
float gaussian_box_radius(float radius, int passes) {
  float sigma2, L, l, a;
  sigma2 = radius * radius / passes;
  L = sqrt(12.0 * sigma2 + 1.0);
  l = floor((L - 1.0) / 2.0);
  a = (2 * l + 1) * (l * (l + 1) - 3.0 * sigma2);
  a /= 6.0 * (sigma2 - (l + 1) * (l + 1));
  return l + a;
}

struct BoxParams {
  int radius;
  uint32_t ww, fw;
};

// BoxBlur.c ImagingHorizontalBoxBlur: the fixed-point weights of one extended box
BoxParams box_params(float float_radius) {
  BoxParams p;
  p.radius = (int)float_radius;
  p.ww = (uint32_t)((1 << 24) / (float_radius * 2 + 1));
  p.fw = ((1 << 24) - (p.radius * 2 + 1) * p.ww) / 2;
  return p;
}

// BoxBlur.c ImagingLineBoxBlur32 for one channel of one line of n pixels (in / out strided, distinct buffers)
__host__ __device__ inline void box_blur_line(const uint8_t* in, int is, uint8_t* out, int os, int n, int radius,
                                              uint32_t ww, uint32_t fw) {
  const int lastx = n - 1;
  const int edgeA = radius + 1 < n ? radius + 1 : n;
  const int edgeB = n - radius - 1 > 0 ? n - radius - 1 : 0;
  uint32_t acc = (uint32_t)in[0] * (uint32_t)(radius + 1);
  for (int x = 0; x < edgeA - 1; ++x) acc += in[x * is];
  acc += (uint32_t)in[lastx * is] * (uint32_t)(radius - edgeA + 1);
#define DV_BOX_STEP(x, sub, add, left, right)                                                             \
  do {                                                                                                    \
    acc += (uint32_t)in[(add) * is] - (uint32_t)in[(sub) * is];                                           \
    const uint32_t bulk = acc * ww + ((uint32_t)in[(left) * is] + (uint32_t)in[(right) * is]) * fw;       \
    out[(x) * os] = (uint8_t)((bulk + (1u << 23)) >> 24);                                                 \
  } while (0)
  if (edgeA <= edgeB) {
    for (int x = 0; x < edgeA; ++x) DV_BOX_STEP(x, 0, x + radius, 0, x + radius + 1);
    for (int x = edgeA; x < edgeB; ++x) DV_BOX_STEP(x, x - radius - 1, x + radius, x - radius - 1, x + radius + 1);
    for (int x = edgeB; x <= lastx; ++x) DV_BOX_STEP(x, x - radius - 1, lastx, x - radius - 1, lastx);
  } else {
    for (int x = 0; x < edgeB; ++x) DV_BOX_STEP(x, 0, x + radius, 0, x + radius + 1);
    for (int x = edgeB; x < edgeA; ++x) DV_BOX_STEP(x, 0, lastx, 0, lastx);
    for (int x = edgeA; x <= lastx; ++x) DV_BOX_STEP(x, x - radius - 1, lastx, x - radius - 1, lastx);
  }
#undef DV_BOX_STEP
}

constexpr int kPasses = 3;
constexpr int kBlurThreads = 384;

// in / out: float [B][3][F][HW]; params: int32 [B*F][4] = {apply, radius, ww, fw}; blockIdx.x = frame
__global__ void __launch_bounds__(kBlurThreads) gaussian_blur_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                     const int* __restrict__ params, int F, int H, int W,
                                                                     int n_frames) {
  extern __shared__ uint8_t s_buf[];       // two frames of [3][H][W] bytes
  const int HW = H * W;
  uint8_t* a = s_buf;
  uint8_t* bb = s_buf + 3 * HW;
  for (int n = blockIdx.x; n < n_frames; n += gridDim.x) {
    const int b = n / F, f = n - b * F;
    const long long plane = (long long)F * HW;
    const long long base = ((long long)b * 3 * F + f) * HW;
    const int* prm = params + (long long)n * 4;
    if (prm[0] == 0) {                      // clip not blurred: RandomApply skipped the stage, the floats pass through
      for (int i = threadIdx.x; i < 3 * HW; i += kBlurThreads) {
        const int c = i / HW, j = i - c * HW;
        out[base + c * plane + j] = in[base + c * plane + j];
      }
      continue;
    }
    const int radius = prm[1];
    const uint32_t ww = (uint32_t)prm[2], fw = (uint32_t)prm[3];
    // ToPILImage: pic.mul(255).byte() - float32 product, truncated
    for (int i = threadIdx.x; i < 3 * HW; i += kBlurThreads) {
      const int c = i / HW, j = i - c * HW;
      a[i] = (uint8_t)(int)__fmul_rn(in[base + c * plane + j], 255.f);
    }
    __syncthreads();
    uint8_t* src = a;
    uint8_t* dst = bb;
    for (int pass = 0; pass < kPasses; ++pass) {          // along W: one thread per (channel, row)
      for (int l = threadIdx.x; l < 3 * H; l += kBlurThreads)
        box_blur_line(src + l * W, 1, dst + l * W, 1, W, radius, ww, fw);
      __syncthreads();
      uint8_t* t = src; src = dst; dst = t;
    }
    for (int pass = 0; pass < kPasses; ++pass) {          // along H: one thread per (channel, column)
      for (int l = threadIdx.x; l < 3 * W; l += kBlurThreads) {
        const int c = l / W, x = l - c * W;
        box_blur_line(src + c * HW + x, W, dst + c * HW + x, W, H, radius, ww, fw);
      }
      __syncthreads();
      uint8_t* t = src; src = dst; dst = t;
    }
    for (int i = threadIdx.x; i < 3 * HW; i += kBlurThreads) {      // ToTensor: / 255
      const int c = i / HW, j = i - c * HW;
      out[base + c * plane + j] = __fdiv_rn((float)src[i], 255.f);
    }
    __syncthreads();
  }
}

}  // namespace

// host-only: {apply, radius, ww, fw} of a sigma as the library computes it (sigma <= 0: not applied)
void blur_params_host(float sigma, int* out4) {
  if (!(sigma > 0.f)) { out4[0] = out4[1] = out4[2] = out4[3] = 0; return; }
  const float r = gaussian_box_radius(sigma, kPasses);
  const BoxParams p = box_params(r);
  // Pillow skips a direction whose box radius is exactly 0 (ImagingBoxBlur: `if (xradius != 0)`): the frame then only
  // takes the uint8 round trip
  out4[0] = r != 0.f ? 1 : 2;
  out4[1] = p.radius; out4[2] = (int)p.ww; out4[3] = (int)p.fw;
}

// host-only: the whole stage on one CHW float frame with the code the kernel runs (for the CPU parity test)
int frames_gaussian_blur_host(const float* frame, float* out, int H, int W, float sigma) {
  const int HW = H * W;
  int prm[4];
  blur_params_host(sigma, prm);
  if (prm[0] == 0) {                       // stage not applied: the frame passes through, as in the kernel
    for (int i = 0; i < 3 * HW; ++i) out[i] = frame[i];
    return kOk;
  }
  std::vector<uint8_t> a(3 * HW), b(3 * HW);
  for (int i = 0; i < 3 * HW; ++i) a[i] = (uint8_t)(int)(frame[i] * 255.f);
  uint8_t* src = a.data();
  uint8_t* dst = b.data();
  if (prm[0] == 1) {
    for (int pass = 0; pass < kPasses; ++pass) {
      for (int l = 0; l < 3 * H; ++l) box_blur_line(src + l * W, 1, dst + l * W, 1, W, prm[1], (uint32_t)prm[2], (uint32_t)prm[3]);
      std::swap(src, dst);
    }
    for (int pass = 0; pass < kPasses; ++pass) {
      for (int l = 0; l < 3 * W; ++l) {
        const int c = l / W, x = l - c * W;
        box_blur_line(src + c * HW + x, W, dst + c * HW + x, W, H, prm[1], (uint32_t)prm[2], (uint32_t)prm[3]);
      }
      std::swap(src, dst);
    }
  }
  for (int i = 0; i < 3 * HW; ++i) out[i] = (float)src[i] / 255.f;
  return kOk;
}

int frames_gaussian_blur(const float* in, float* out, const int* params, int B, int F, int H, int W, cudaStream_t stream) {
  const size_t smem = (size_t)2 * 3 * H * W;
  if (smem > 200 * 1024) return fail(kUnsupported, "Gaussian blur keeps a frame in shared memory: %d x %d is too large", H, W);
  static bool attr_set = false;
  if (!attr_set) {
    DV_CUDA_OK(cudaFuncSetAttribute(gaussian_blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int n_frames = B * F;
  const int grid = std::min(n_frames, sm_count() * 8);
  gaussian_blur_kernel<<<grid, kBlurThreads, smem, stream>>>(in, out, params, F, H, W, n_frames);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
