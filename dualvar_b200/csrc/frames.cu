// Frame staging in front of the ingest kernel (first stage): decoded uint8 frames ->
// A.Scale((128, 171)) (PIL bicubic) -> A.RandomCrop(112) as the reference's loader does per clip on the host CPU
// (utils/augmentation.py:125-176, the null_transform of pretrain.py:491-497; dataset/local_dataset.py:289-300),
// bit-exact with Pillow's 8-bit resampler (src/libImaging/Resample.c): per output pixel a bicubic window
// (a = -0.5, support 2 * max(scale, 1)) evaluated in double on the host, normalised, converted to 22-bit fixed point,
// accumulated in int32 from 1 << 21, shifted and saturated; horizontal pass first through a uint8 intermediate, then the
// vertical pass - computed only for the cropped window and written in the planar uint8 layout dv_ingest_clips_u8 reads
// (ToTensor's x / 255, Normalize and the NDHWC / space-to-depth conversion happen there).
// Integer byte work, HBM-bound and tiny next to the encoder (1.8 MB of output per 48-frame sample).
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "host_common.h"

namespace dv {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// one axis: per output index [xmin, xmax, kk[0 .. ksize)] (Resample.c precompute_coeffs + normalize_coeffs_8bpc)
std::vector<int> axis_table(int in_size, int out_size, int* ksize_out) {
  double scale, filterscale;
  scale = filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  std::vector<int> tab((size_t)out_size * (ksize + 2), 0);
  std::vector<double> k(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    int* row = &tab[(size_t)xx * (ksize + 2)];
    row[0] = xmin;
    row[1] = xmax;
    for (int x = 0; x < xmax; ++x) {
      double w = k[x];
      if (ww != 0.0) w /= ww;
      row[2 + x] = w < 0 ? (int)(-0.5 + w * (1 << kPrecisionBits)) : (int)(0.5 + w * (1 << kPrecisionBits));
    }
  }
  *ksize_out = ksize;
  return tab;
}

struct DevTable {
  int* ptr;
  int ksize;
};

// coefficient tables live on the device for the life of the process, one per (device, in, out)
int get_table(int in_size, int out_size, DevTable* out) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, DevTable> cache;
  int dev = 0;
  DV_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(dev, in_size, out_size);
  auto it = cache.find(key);
  if (it == cache.end()) {
    DevTable t;
    const std::vector<int> tab = axis_table(in_size, out_size, &t.ksize);
    DV_CUDA_OK(cudaMalloc(&t.ptr, tab.size() * sizeof(int)));
    DV_CUDA_OK(cudaMemcpy(t.ptr, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
    it = cache.emplace(key, t).first;
  }
  *out = it->second;
  return kOk;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)min(max(v, 0), 255);
}

// src [rows][Ws][3] -> tmp [rows][out_w][3]; one thread = one output pixel
__global__ void __launch_bounds__(256) frames_resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ tmp,
                                                              const int* __restrict__ tab, int ksize, long long total,
                                                              int Ws, int out_w) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % out_w);
    const long long row = i / out_w;
    const int* t = tab + (long long)xx * (ksize + 2);
    const int xmin = t[0], xmax = t[1];
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    const uint8_t* p = src + (row * Ws + xmin) * 3;
    for (int x = 0; x < xmax; ++x) {
      const int k = t[2 + x];
      s0 += p[3 * x] * k; s1 += p[3 * x + 1] * k; s2 += p[3 * x + 2] * k;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
  }
}

// tmp [B*F][Hs][out_w][3] -> out [B][3][F][crop_h][crop_w]: vertical pass of the rows / columns inside each clip's crop
__global__ void __launch_bounds__(256) frames_resize_v_crop_kernel(const uint8_t* __restrict__ tmp,
                                                                   uint8_t* __restrict__ out, const int* __restrict__ tab,
                                                                   int ksize, const int* __restrict__ crop_lu, int F,
                                                                   int T, int V, int Hs, int out_w, int out_h, int crop_w,
                                                                   int crop_h, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % crop_w);
    long long r = i / crop_w;
    const int yy = (int)(r % crop_h); r /= crop_h;
    const int f = (int)(r % F);
    const long long b = r / F;
    const int* lu = crop_lu + (b * V + f / T) * 2;
    const int left = min(max(lu[0], 0), out_w - crop_w), upper = min(max(lu[1], 0), out_h - crop_h);
    const int* t = tab + (long long)(upper + yy) * (ksize + 2);
    const int ymin = t[0], ymax = t[1];
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    const uint8_t* p = tmp + (((b * F + f) * Hs + ymin) * out_w + left + xx) * 3;
    const long long rs = (long long)out_w * 3;
    for (int y = 0; y < ymax; ++y) {
      const int k = t[2 + y];
      s0 += p[y * rs] * k; s1 += p[y * rs + 1] * k; s2 += p[y * rs + 2] * k;
    }
    const long long plane = (long long)F * crop_h * crop_w;
    uint8_t* o = out + (b * 3) * plane + ((long long)f * crop_h + yy) * crop_w + xx;
    o[0] = clip8(s0); o[plane] = clip8(s1); o[2 * plane] = clip8(s2);
  }
}

int fgrid(long long total) {
  long long g = ceil_div_ll(total, 256);
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

// host-only: the fixed-point coefficient table of one axis, for checking the host arithmetic against Pillow's
int frames_axis_table_host(int in_size, int out_size, int* tab, int capacity, int* ksize) {
  int ks = 0;
  const std::vector<int> t = axis_table(in_size, out_size, &ks);
  *ksize = ks;
  if ((long long)t.size() > capacity) return fail(kBadArg, "table needs %lld ints, buffer holds %d", (long long)t.size(), capacity);
  for (size_t i = 0; i < t.size(); ++i) tab[i] = t[i];
  return kOk;
}

int frames_scale_crop_u8(const uint8_t* frames, uint8_t* tmp, uint8_t* out, const int* crop_lu, int B, int V, int T,
                         int Hs, int Ws, int scale_w, int scale_h, int crop_w, int crop_h, cudaStream_t stream) {
  DevTable th, tv;
  if (int rc = get_table(Ws, scale_w, &th)) return rc;
  if (int rc = get_table(Hs, scale_h, &tv)) return rc;
  const int F = V * T;
  const long long rows = (long long)B * F * Hs;
  frames_resize_h_kernel<<<fgrid(rows * scale_w), 256, 0, stream>>>(frames, tmp, th.ptr, th.ksize, rows * scale_w, Ws,
                                                                    scale_w);
  DV_LAUNCH_OK();
  const long long total = (long long)B * F * crop_h * crop_w;
  frames_resize_v_crop_kernel<<<fgrid(total), 256, 0, stream>>>(tmp, out, tv.ptr, tv.ksize, crop_lu, F, T, V, Hs, scale_w,
                                                                scale_h, crop_w, crop_h, total);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
