// Colour jitter of the reference loader on the GPU (second stage): A.ColorJitter of
// utils/augmentation.py:429-660 as pretrain.py:505 builds it (block = 1) - per frame (or per clip when consistent) a
// random transform made of torchvision's tensor adjust_brightness / adjust_contrast / adjust_saturation / adjust_hue in
// a shuffled order, applied to the ToTensor output (CHW float32 in [0, 1]). The arithmetic lives in torchvision 0.26
// (transforms/_functional_tensor.py: _blend, rgb_to_grayscale, _rgb2hsv, _hsv2rgb) and is followed here operation by
// operation with round-to-nearest intrinsics (no fma contraction), so every pixel op is bit-identical to the CPU result;
// only the mean of adjust_contrast is summed in a different order (double accumulation, <= 1 ulp of the mean).
//
// One CTA owns one frame: the frame lives in shared memory as float32 (3 x 112 x 112 x 4 B = 147 KB) while the up to four
// operations run over it, so HBM sees one uint8 read and one float32 write per pixel whatever the transform is.
// Input: the planar uint8 clips dv_frames_scale_crop_u8 writes; output: float32 in the same planar layout, which the
// fp32 ingest kernel (Normalize + NDHWC / space-to-depth) consumes. Per-frame parameters come from the host, drawn in
// the reference's RNG order (dualvar_b200/frames.py: draw_color_jitter).
#include <algorithm>

#include "host_common.h"

namespace dv {

namespace {

constexpr int kJitterThreads = 512;
constexpr int kParamStride = 12;   // apply, b, 1-b, c, 1-c, s, 1-s, h, op0..op3

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__device__ __forceinline__ float gray(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}

// torchvision adjust_hue on one pixel: _rgb2hsv, h = (h + f) mod 1, _hsv2rgb
__device__ __forceinline__ void hue_pixel(float& r, float& g, float& b, float hf) {
  const float maxc = fmaxf(fmaxf(r, g), b), minc = fminf(fminf(r, g), b);
  const bool eqc = maxc == minc;
  const float cr = __fsub_rn(maxc, minc);
  const float s = __fdiv_rn(cr, eqc ? 1.f : maxc);
  const float dv = eqc ? 1.f : cr;
  const float rc = __fdiv_rn(__fsub_rn(maxc, r), dv), gc = __fdiv_rn(__fsub_rn(maxc, g), dv),
              bc = __fdiv_rn(__fsub_rn(maxc, b), dv);
  const float hr = (maxc == r) ? __fsub_rn(bc, gc) : 0.f;
  const float hg = (maxc == g && maxc != r) ? __fsub_rn(__fadd_rn(2.f, rc), bc) : 0.f;
  const float hb = (maxc != g && maxc != r) ? __fsub_rn(__fadd_rn(4.f, gc), rc) : 0.f;
  float h = __fadd_rn(__fadd_rn(hr, hg), hb);
  h = fmodf(__fadd_rn(__fdiv_rn(h, 6.f), 1.f), 1.f);
  h = fmodf(__fadd_rn(h, hf), 1.f);                  // torch.remainder: the result takes the divisor's sign
  if (h < 0.f) h = __fadd_rn(h, 1.f);
  const float v = maxc;
  const float h6 = __fmul_rn(h, 6.f);
  const float fi = floorf(h6);
  const float fr = __fsub_rn(h6, fi);
  const int i = ((int)fi) % 6;
  const float p = clamp01(__fmul_rn(v, __fsub_rn(1.f, s)));
  const float q = clamp01(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, fr))));
  const float t = clamp01(__fmul_rn(v, __fsub_rn(1.f, __fmul_rn(s, __fsub_rn(1.f, fr)))));
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// in: uint8 [B][3][F][HW], out: float [B][3][F][HW], params: float [B*F][12]; blockIdx.x = frame
__global__ void __launch_bounds__(kJitterThreads) color_jitter_kernel(const uint8_t* __restrict__ in,
                                                                      float* __restrict__ out,
                                                                      const float* __restrict__ params, int F, int HW,
                                                                      int n_frames) {
  extern __shared__ float s_px[];          // [3][HW]
  __shared__ double s_red[kJitterThreads / 32];
  __shared__ float s_mean;
  float* sr = s_px;
  float* sg = s_px + HW;
  float* sb = s_px + 2 * HW;
  for (int n = blockIdx.x; n < n_frames; n += gridDim.x) {
    const int b = n / F, f = n - b * F;
    const long long plane = (long long)F * HW;
    const long long base = ((long long)b * 3 * F + f) * HW;     // channel 0 of this frame; channel c adds c * plane
    for (int i = threadIdx.x; i < HW; i += kJitterThreads) {   // ToTensor: x / 255 in float32
      sr[i] = __fdiv_rn((float)in[base + i], 255.f);
      sg[i] = __fdiv_rn((float)in[base + plane + i], 255.f);
      sb[i] = __fdiv_rn((float)in[base + 2 * plane + i], 255.f);
    }
    __syncthreads();
    const float* prm = params + (long long)n * kParamStride;
    if (prm[0] != 0.f) {
      for (int k = 0; k < 4; ++k) {
        const int op = (int)prm[8 + k];
        if (op == 0) {                                   // brightness: (b * x + (1 - b) * 0).clamp(0, 1)
          const float ratio = prm[1];
          for (int i = threadIdx.x; i < 3 * HW; i += kJitterThreads) s_px[i] = clamp01(__fmul_rn(ratio, s_px[i]));
        } else if (op == 1) {                            // contrast: blend with the mean grey level of the frame
          double part = 0.0;
          for (int i = threadIdx.x; i < HW; i += kJitterThreads) part += (double)gray(sr[i], sg[i], sb[i]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
          __syncthreads();
          if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kJitterThreads / 32; ++w) t += s_red[w];
            s_mean = (float)(t / (double)HW);
          }
          __syncthreads();
          const float ratio = prm[3], add = __fmul_rn(prm[4], s_mean);
          for (int i = threadIdx.x; i < 3 * HW; i += kJitterThreads)
            s_px[i] = clamp01(__fadd_rn(__fmul_rn(ratio, s_px[i]), add));
        } else if (op == 2) {                            // saturation: blend with the pixel's grey level
          const float ratio = prm[5], om = prm[6];
          for (int i = threadIdx.x; i < HW; i += kJitterThreads) {
            const float r = sr[i], g = sg[i], bb = sb[i];
            const float add = __fmul_rn(om, gray(r, g, bb));
            sr[i] = clamp01(__fadd_rn(__fmul_rn(ratio, r), add));
            sg[i] = clamp01(__fadd_rn(__fmul_rn(ratio, g), add));
            sb[i] = clamp01(__fadd_rn(__fmul_rn(ratio, bb), add));
          }
        } else if (op == 3) {                            // hue
          const float hf = prm[7];
          for (int i = threadIdx.x; i < HW; i += kJitterThreads) {
            float r = sr[i], g = sg[i], bb = sb[i];
            hue_pixel(r, g, bb, hf);
            sr[i] = r; sg[i] = g; sb[i] = bb;
          }
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < HW; i += kJitterThreads) {
      out[base + i] = sr[i];
      out[base + plane + i] = sg[i];
      out[base + 2 * plane + i] = sb[i];
    }
    __syncthreads();
  }
}

}  // namespace

int frames_color_jitter(const uint8_t* clips, float* out, const float* params, int B, int F, int H, int W,
                        cudaStream_t stream) {
  const int HW = H * W;
  const size_t smem = (size_t)3 * HW * sizeof(float);
  if (smem > 200 * 1024) return fail(kUnsupported, "colour jitter keeps a frame in shared memory: %d x %d is too large", H, W);
  static bool attr_set = false;
  if (!attr_set) {
    DV_CUDA_OK(cudaFuncSetAttribute(color_jitter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int n_frames = B * F;
  const int grid = std::min(n_frames, sm_count() * 8);
  color_jitter_kernel<<<grid, kJitterThreads, smem, stream>>>(clips, out, params, F, HW, n_frames);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
