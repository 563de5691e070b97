// Frame staging in front of the ingest kernel (first stage): decoded uint8 frames ->
// A.Scale((128, 171)) (PIL bicubic) -> A.RandomCrop(112) as the reference's loader does per clip on the host CPU
// (utils/augmentation.py:125-176, the null_transform of pretrain.py:491-497; dataset/local_dataset.py:289-300),
// bit-exact with Pillow's 8-bit resampler (src/libImaging/Resample.c): per output pixel a bicubic window
// (a = -0.5, support 2 * max(scale, 1)) evaluated in double on the host, normalised, converted to 22-bit fixed point,
// accumulated in int32 from 1 << 21, shifted and saturated; horizontal pass first through a uint8 intermediate (kept as
// RGBX words so that both passes move aligned 32-bit pixels), then the vertical pass - computed only for the cropped window and written in the planar uint8 layout dv_ingest_clips_u8 reads
// (ToTensor's x / 255, Normalize and the NDHWC / space-to-depth conversion happen there).
// Integer byte work, HBM-bound and tiny next to the encoder (1.8 MB of output per 48-frame sample).
#include <math.h>

#include <algorithm>

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

// Which rows of a frame the vertical pass will read: those inside the windows of the clip's crop rows
// [upper, upper + crop_h). The horizontal pass only computes these (about two thirds of a 240-row frame for a 112-row crop).
struct RowFilter {
  const int* tabv;      // vertical table (NULL: every row is needed)
  const int* crop_lu;
  int ksize_v, Hs, F, T, V, out_h, crop_h;
  // [lo, hi) of frame n (block-uniform: evaluated once per frame, not per row)
  __device__ __forceinline__ void range(int n, int* lo, int* hi) const {
    if (tabv == nullptr) { *lo = 0; *hi = Hs; return; }
    const int b = n / F, f = n - b * F;
    const int upper = min(max(crop_lu[(b * V + f / T) * 2 + 1], 0), out_h - crop_h);
    *lo = tabv[upper * (ksize_v + 2)];
    const int* last = tabv + (upper + crop_h - 1) * (ksize_v + 2);
    *hi = last[0] + last[1];
  }
};

// Horizontal pass: src [rows][Ws][3] -> tmp [rows][out_w] RGBX words. A block owns output columns (thread = column,
// its fixed-point taps live in registers for the block's whole life) and walks over rows: the row is staged in shared
// memory with 4-byte loads, every thread forms its 3 sums from shared bytes, the result leaves as one 32-bit store per
// thread (coalesced). kTaps bounds the unrolled tap loop (8: up-scaling / mild down-scaling, 12: 320 -> 128, 16: down-scaling
// by <= 3.5). The shared row has 64 bytes of slack behind it: windows are read as whole words.
template <int kTaps>
__global__ void __launch_bounds__(256) frames_resize_h_kernel(const uint8_t* __restrict__ src, uint32_t* __restrict__ tmp,
                                                              const int* __restrict__ tab, int ksize, int n_frames,
                                                              int Ws, int out_w, const RowFilter rf) {
  // grid: x = row slices of a frame, y = frames, z = 256-column blocks. All index arithmetic that is not per pixel is
  // block-uniform and done once per frame (the first version spent two thirds of its instructions on it).
  extern __shared__ __align__(16) uint8_t s_row[];
  const int xx = blockIdx.z * blockDim.x + threadIdx.x;
  const bool live = xx < out_w;
  int k[kTaps];
  int xmin = 0, xmax = 0;
  if (live) {
    const int* t = tab + (long long)xx * (ksize + 2);
    xmin = t[0]; xmax = t[1];
#pragma unroll
    for (int x = 0; x < kTaps; ++x) k[x] = x < xmax ? t[2 + x] : 0;
  } else {
#pragma unroll
    for (int x = 0; x < kTaps; ++x) k[x] = 0;
  }
  // kRowsPerIter consecutive rows are staged per iteration (more bytes in flight per block)
  constexpr int kRowsPerIter = 4;
  constexpr int kWords = (3 * kTaps + 3) / 4;
  const int row_bytes = Ws * 3;
  const int row_stride = (row_bytes + 64 + 15) & ~15;     // shared-memory row pitch (slack for whole-word windows)
  const bool words = (row_bytes & 3) == 0;      // rows start 4-byte aligned when the row length is a multiple of 4
  const int boff = xmin * 3;
  const int w_off = boff >> 2, sh = (boff & 3) * 8;
  for (int n = blockIdx.y; n < n_frames; n += gridDim.y) {
    int lo, hi;
    rf.range(n, &lo, &hi);
    const uint8_t* fsrc = src + (long long)n * rf.Hs * row_bytes;
    uint32_t* fdst = tmp + (long long)n * rf.Hs * out_w;
    for (int y0 = lo + kRowsPerIter * blockIdx.x; y0 < hi; y0 += kRowsPerIter * gridDim.x) {
      const int nr = min(kRowsPerIter, hi - y0);
#pragma unroll
      for (int r = 0; r < kRowsPerIter; ++r) {
        if (r >= nr) break;
        const uint8_t* p = fsrc + (long long)(y0 + r) * row_bytes;
        uint8_t* d = s_row + r * row_stride;
        if (words) {
          const uint32_t* pw = reinterpret_cast<const uint32_t*>(p);
          uint32_t* sw = reinterpret_cast<uint32_t*>(d);
          for (int i = threadIdx.x; i < (row_bytes >> 2); i += blockDim.x) sw[i] = pw[i];
        } else {
          for (int i = threadIdx.x; i < row_bytes; i += blockDim.x) d[i] = p[i];
        }
      }
      __syncthreads();
      if (live) {
#pragma unroll 1
        for (int r = 0; r < nr; ++r) {
          // the thread's window (3 * kTaps bytes from byte xmin * 3) as aligned words, realigned with funnel shifts
          int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
          const uint32_t* sw = reinterpret_cast<const uint32_t*>(s_row + r * row_stride) + w_off;
          uint32_t w[kWords + 1];
#pragma unroll
          for (int i = 0; i <= kWords; ++i) w[i] = sw[i];
          uint32_t a[kWords];
#pragma unroll
          for (int i = 0; i < kWords; ++i) a[i] = __funnelshift_r(w[i], w[i + 1], sh);
#pragma unroll
          for (int x = 0; x < kTaps; ++x) {          // taps beyond xmax have k = 0 (their bytes may be the next pixels)
            const int j = 3 * x;
            s0 += (int)__byte_perm(a[j >> 2], 0, 0x4440 + (j & 3)) * k[x];
            s1 += (int)__byte_perm(a[(j + 1) >> 2], 0, 0x4440 + ((j + 1) & 3)) * k[x];
            s2 += (int)__byte_perm(a[(j + 2) >> 2], 0, 0x4440 + ((j + 2) & 3)) * k[x];
          }
          fdst[(y0 + r) * out_w + xx] = (uint32_t)clip8(s0) | ((uint32_t)clip8(s1) << 8) | ((uint32_t)clip8(s2) << 16);
        }
      }
      __syncthreads();
    }
  }
}

// generic fallback (any tap count): one thread = one output pixel, taps read from the table
__global__ void __launch_bounds__(256) frames_resize_h_generic_kernel(const uint8_t* __restrict__ src,
                                                                      uint32_t* __restrict__ tmp,
                                                                      const int* __restrict__ tab, int ksize,
                                                                      long long total, int Ws, int out_w) {
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
    tmp[i] = (uint32_t)clip8(s0) | ((uint32_t)clip8(s1) << 8) | ((uint32_t)clip8(s2) << 16);
  }
}

// Vertical pass + crop: tmp [B*F][Hs][out_w] RGBX -> out [B][3][F][crop_h][crop_w]; only the rows / columns inside each
// clip's crop. One thread = one output pixel: one aligned 32-bit load per tap (coalesced along the row), taps are
// warp-uniform table reads, three planar byte stores (consecutive threads -> consecutive bytes).
template <int kTaps>
__global__ void __launch_bounds__(256) frames_resize_v_crop_kernel(const uint32_t* __restrict__ tmp,
                                                                   uint8_t* __restrict__ out, const int* __restrict__ tab,
                                                                   int ksize, const int* __restrict__ crop_lu, int F,
                                                                   int T, int V, int Hs, int out_w, int out_h, int crop_w,
                                                                   int crop_h, int n_frames) {
  // grid: x = groups of kRowsPerBlock crop rows, y = frames, z = 256-column blocks; thread = crop column. Everything
  // but the pixel loads and multiply-adds is block-uniform (frame decomposition, crop origin, the row's taps).
  // kTaps > 0: the tap loop is unrolled (all row loads of a pixel in flight together); kTaps == 0: any tap count.
  constexpr int kRowsPerBlock = 8;
  const int xx = blockIdx.z * blockDim.x + threadIdx.x;
  if (xx >= crop_w) return;
  const long long plane = (long long)F * crop_h * crop_w;
  for (int n = blockIdx.y; n < n_frames; n += gridDim.y) {
    const int b = n / F, f = n - b * F;
    const int* lu = crop_lu + (b * V + f / T) * 2;
    const int left = min(max(lu[0], 0), out_w - crop_w), upper = min(max(lu[1], 0), out_h - crop_h);
    const uint32_t* fsrc = tmp + (long long)n * Hs * out_w + left + xx;
    uint8_t* fdst = out + (long long)b * 3 * plane + (long long)f * crop_h * crop_w + xx;
    const int y_end = min(crop_h, (int)(blockIdx.x + 1) * kRowsPerBlock);
    for (int yy = blockIdx.x * kRowsPerBlock; yy < y_end; ++yy) {
      const int* t = tab + (upper + yy) * (ksize + 2);
      const int ymin = t[0], ymax = t[1];
      int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
      const uint32_t* q = fsrc + ymin * out_w;
      if (kTaps > 0) {
        uint32_t px[kTaps > 0 ? kTaps : 1];
#pragma unroll
        for (int y = 0; y < kTaps; ++y) px[y] = y < ymax ? q[y * out_w] : 0u;
#pragma unroll
        for (int y = 0; y < kTaps; ++y) {
          const int k = y < ymax ? t[2 + y] : 0;          // uniform load
          s0 += (int)__byte_perm(px[y], 0, 0x4440) * k; s1 += (int)__byte_perm(px[y], 0, 0x4441) * k;
          s2 += (int)__byte_perm(px[y], 0, 0x4442) * k;
        }
      } else {
        for (int y = 0; y < ymax; ++y) {
          const int k = t[2 + y];
          const uint32_t px = q[y * out_w];
          s0 += (int)(px & 0xffu) * k; s1 += (int)((px >> 8) & 0xffu) * k; s2 += (int)((px >> 16) & 0xffu) * k;
        }
      }
      uint8_t* o = fdst + yy * crop_w;
      o[0] = clip8(s0); o[plane] = clip8(s1); o[2 * plane] = clip8(s2);
    }
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
  uint32_t* tmp32 = reinterpret_cast<uint32_t*>(tmp);
  const int smem = 4 * ((Ws * 3 + 64 + 15) & ~15);     // 4 rows per iteration, each with slack for whole-word windows
  const int n_frames = B * F;
  if (th.ksize <= 16 && smem <= 48 * 1024) {
    const int threads = std::min(256, round_up(scale_w, 32));
    RowFilter rf;
    rf.tabv = tv.ptr;
    rf.crop_lu = crop_lu; rf.ksize_v = tv.ksize; rf.Hs = Hs; rf.F = F; rf.T = T; rf.V = V; rf.out_h = scale_h;
    rf.crop_h = crop_h;
    // 8 row slices per frame: ~5 four-row iterations per block for a 112-row crop of a 240-row frame
    dim3 grid(8, std::min(n_frames, 32768), ceil_div(scale_w, threads));
    if (th.ksize <= 8)
      frames_resize_h_kernel<8><<<grid, threads, smem, stream>>>(frames, tmp32, th.ptr, th.ksize, n_frames, Ws, scale_w, rf);
    else if (th.ksize <= 12)
      frames_resize_h_kernel<12><<<grid, threads, smem, stream>>>(frames, tmp32, th.ptr, th.ksize, n_frames, Ws, scale_w, rf);
    else
      frames_resize_h_kernel<16><<<grid, threads, smem, stream>>>(frames, tmp32, th.ptr, th.ksize, n_frames, Ws, scale_w, rf);
  } else {
    frames_resize_h_generic_kernel<<<fgrid(rows * scale_w), 256, 0, stream>>>(frames, tmp32, th.ptr, th.ksize,
                                                                              rows * scale_w, Ws, scale_w);
  }
  DV_LAUNCH_OK();
  const int vthreads = std::min(256, round_up(crop_w, 32));
  dim3 vgrid(ceil_div(crop_h, 8), std::min(n_frames, 32768), ceil_div(crop_w, vthreads));
#define DV_VPASS(K)                                                                                                  \
  frames_resize_v_crop_kernel<K><<<vgrid, vthreads, 0, stream>>>(tmp32, out, tv.ptr, tv.ksize, crop_lu, F, T, V, Hs, \
                                                                 scale_w, scale_h, crop_w, crop_h, n_frames)
  if (tv.ksize <= 8) DV_VPASS(8); else if (tv.ksize <= 16) DV_VPASS(16); else DV_VPASS(0);
#undef DV_VPASS
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
