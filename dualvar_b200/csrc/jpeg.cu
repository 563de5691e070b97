// Baseline JPEG frame decoding for the GPU data path.
//
// Replaces PIL.Image.open(...) of the reference loader (dataset/local_dataset.py:283-286), i.e. the IJG / libjpeg-turbo
// decoder with its default settings: jidctint.c "islow" inverse DCT, jdsample.c "fancy" (triangle filter) chroma
// upsampling, jdcolor.c 16-bit fixed-point YCbCr -> RGB. All three are integer algorithms, restated here so that the
// decoded pixels are bit-identical to Pillow's (oracle/jpeg.py is the numpy restatement pinned against Pillow).
//
// Split of the work: entropy (Huffman) decoding is a serial bit-stream walk - it runs on host threads (this file,
// huffman_decode_host) and yields the quantised DCT coefficients, 2 bytes per sample; dequantisation + IDCT + upsampling +
// colour conversion - all the arithmetic - run on the GPU over a whole batch of frames (kernels below).
//   coefficients per frame: component 0 blocks [bv0][bh0][64] int16 (natural order), then component 1, component 2
//   planes per frame      : component c samples uint8 [bv_c*8][bh_c*8]
//   output                : uint8 [n][H][W][3] RGB - the layout dv_frames_scale_crop_u8 reads
// Supported: SOF0 (baseline sequential), 8 bit, 1 or 3 components, 4:4:4 / 4:2:2 / 4:2:0, restart intervals.
#include "host_common.h"

#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

namespace dv {

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
  bool present = false;
  // 9-bit lookahead: (length << 8) | symbol, 0 = longer code
  uint16_t look[512];
  int maxcode[18];
  int mincode[17];
  int valptr[17];
  uint8_t symbols[256];
};

struct Component {
  int id, h, v, tq, td, ta;
};

struct Header {
  int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0, restart = 0;
  Component comp[3];
  uint16_t qt[4][64];     // natural order
  bool qt_present[4] = {false, false, false, false};
  HuffTable dc[4], ac[4];
  size_t scan_pos = 0;
  long long coef_count = 0;   // int16 elements per frame
};

void build_table(HuffTable& t, const uint8_t* counts, const uint8_t* symbols, int n) {
  t.present = true;
  memcpy(t.symbols, symbols, n);
  memset(t.look, 0, sizeof(t.look));
  int code = 0, k = 0;
  for (int len = 1; len <= 16; ++len) {
    t.valptr[len] = k;
    t.mincode[len] = code;
    for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
      if (len <= 9) {
        const int first = code << (9 - len);
        for (int j = 0; j < (1 << (9 - len)); ++j) t.look[first + j] = (uint16_t)((len << 8) | symbols[k]);
      }
    }
    t.maxcode[len] = counts[len - 1] ? code - 1 : -1;
    code <<= 1;
  }
  t.maxcode[17] = 0x7fffffff;
}

// nullptr on success, else a message
const char* parse_header(const uint8_t* d, size_t len, Header& h) {
  if (len < 4 || d[0] != 0xFF || d[1] != 0xD8) return "not a JPEG (no SOI)";
  size_t pos = 2;
  bool have_frame = false;
  for (;;) {
    if (pos + 4 > len) return "truncated file";
    if (d[pos] != 0xFF) return "marker expected";
    while (pos + 1 < len && d[pos + 1] == 0xFF) ++pos;
    const int m = d[pos + 1];
    pos += 2;
    if (m == 0xD9) return "EOI before SOS";
    if (pos + 2 > len) return "truncated file";
    const size_t seglen = ((size_t)d[pos] << 8) | d[pos + 1];
    if (seglen < 2 || pos + seglen > len) return "bad segment length";
    const uint8_t* seg = d + pos + 2;
    const size_t n = seglen - 2;
    if (m == 0xDB) {
      size_t i = 0;
      while (i < n) {
        const int pq = seg[i] >> 4, tq = seg[i] & 15;
        if (pq != 0) return "16-bit quantisation tables are not baseline";
        if (tq > 3 || i + 65 > n) return "bad DQT";
        for (int k = 0; k < 64; ++k) h.qt[tq][kZigzag[k]] = seg[i + 1 + k];
        h.qt_present[tq] = true;
        i += 65;
      }
    } else if (m == 0xC0) {
      if (n < 6 || seg[0] != 8) return "only 8-bit samples";
      h.height = (seg[1] << 8) | seg[2];
      h.width = (seg[3] << 8) | seg[4];
      h.ncomp = seg[5];
      if (h.ncomp != 1 && h.ncomp != 3) return "1 or 3 components";
      if (n < (size_t)(6 + 3 * h.ncomp)) return "bad SOF0";
      for (int i = 0; i < h.ncomp; ++i) {
        h.comp[i].id = seg[6 + 3 * i];
        h.comp[i].h = seg[7 + 3 * i] >> 4;
        h.comp[i].v = seg[7 + 3 * i] & 15;
        h.comp[i].tq = seg[8 + 3 * i];
        if (h.comp[i].tq > 3) return "bad quantisation table index";
      }
      have_frame = true;
    } else if ((m >= 0xC1 && m <= 0xCF) && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return "only baseline sequential JPEG (SOF0) is supported";
    } else if (m == 0xC4) {
      size_t i = 0;
      while (i < n) {
        if (i + 17 > n) return "bad DHT";
        const int tc = seg[i] >> 4, th = seg[i] & 15;
        if (tc > 1 || th > 3) return "bad DHT";
        int total = 0;
        for (int k = 0; k < 16; ++k) total += seg[i + 1 + k];
        if (total > 256 || i + 17 + total > n) return "bad DHT";
        build_table(tc ? h.ac[th] : h.dc[th], seg + i + 1, seg + i + 17, total);
        i += 17 + total;
      }
    } else if (m == 0xDD) {
      if (n < 2) return "bad DRI";
      h.restart = (seg[0] << 8) | seg[1];
    } else if (m == 0xDA) {
      if (!have_frame) return "SOS before SOF";
      if (n < 1 || seg[0] != h.ncomp || n < (size_t)(1 + 2 * h.ncomp)) return "non-interleaved scans are not supported";
      for (int i = 0; i < h.ncomp; ++i) {
        const int cid = seg[1 + 2 * i], tt = seg[2 + 2 * i];
        for (int c = 0; c < h.ncomp; ++c)
          if (h.comp[c].id == cid) { h.comp[c].td = tt >> 4; h.comp[c].ta = tt & 15; }
      }
      h.scan_pos = pos + seglen;
      break;
    }
    pos += seglen;
  }
  if (h.ncomp == 1) { h.comp[0].h = h.comp[0].v = 1; }
  h.hmax = h.vmax = 1;
  for (int c = 0; c < h.ncomp; ++c) { h.hmax = h.comp[c].h > h.hmax ? h.comp[c].h : h.hmax; h.vmax = h.comp[c].v > h.vmax ? h.comp[c].v : h.vmax; }
  if (h.ncomp == 3) {
    const bool chroma11 = h.comp[1].h == 1 && h.comp[1].v == 1 && h.comp[2].h == 1 && h.comp[2].v == 1;
    const bool ok = chroma11 && h.comp[0].h == h.hmax && h.comp[0].v == h.vmax &&
                    ((h.hmax == 1 && h.vmax == 1) || (h.hmax == 2 && h.vmax == 1) || (h.hmax == 2 && h.vmax == 2));
    if (!ok) return "unsupported sampling factors";
  }
  if (h.width <= 0 || h.height <= 0) return "empty image";
  h.mcux = (h.width + 8 * h.hmax - 1) / (8 * h.hmax);
  h.mcuy = (h.height + 8 * h.vmax - 1) / (8 * h.vmax);
  h.coef_count = 0;
  for (int c = 0; c < h.ncomp; ++c) {
    if (!h.qt_present[h.comp[c].tq]) return "missing quantisation table";
    if (!h.dc[h.comp[c].td].present || !h.ac[h.comp[c].ta].present) return "missing Huffman table";
    h.coef_count += (long long)h.mcux * h.comp[c].h * h.mcuy * h.comp[c].v * 64;
  }
  return nullptr;
}

struct BitReader {
  const uint8_t* d;
  size_t pos, len;
  uint64_t acc = 0;
  int n = 0;
  inline void fill() {
    while (n <= 48) {
      unsigned b = 0;
      if (pos < len) {
        b = d[pos];
        if (b == 0xFF) {
          const unsigned nxt = pos + 1 < len ? d[pos + 1] : 0xD9;
          if (nxt == 0) pos += 2;
          else b = 0;                      // a marker: feed zeros, do not advance
        } else {
          ++pos;
        }
      }
      acc = (acc << 8) | b;
      n += 8;
    }
  }
  inline unsigned peek(int k) { return (unsigned)((acc >> (n - k)) & ((1u << k) - 1)); }
  inline void skip(int k) { n -= k; }
  inline int get(int k) {
    if (k == 0) return 0;
    if (n < k) fill();
    n -= k;
    return (int)((acc >> n) & ((1u << k) - 1));
  }
  void restart() {
    acc = 0; n = 0;
    while (pos + 1 < len && !(d[pos] == 0xFF && d[pos + 1] >= 0xD0 && d[pos + 1] <= 0xD7)) ++pos;
    pos += 2;
  }
};

inline int decode_symbol(BitReader& br, const HuffTable& t) {
  if (br.n < 16) br.fill();
  const unsigned look = t.look[br.peek(9)];
  if (look) {
    br.skip(look >> 8);
    return look & 255;
  }
  int code = (int)br.peek(9);
  br.skip(9);
  for (int len = 10; len <= 16; ++len) {
    code = (code << 1) | br.get(1);
    if (t.maxcode[len] >= 0 && code <= t.maxcode[len] && code >= t.mincode[len])
      return t.symbols[t.valptr[len] + code - t.mincode[len]];
  }
  return -1;
}

inline int extend(int v, int t) { return v >= (1 << (t - 1)) ? v : v - (1 << t) + 1; }

const char* decode_scan(const uint8_t* d, size_t len, const Header& h, int16_t* coef) {
  memset(coef, 0, sizeof(int16_t) * (size_t)h.coef_count);
  int16_t* base[3];
  int bw[3];
  {
    long long off = 0;
    for (int c = 0; c < h.ncomp; ++c) {
      base[c] = coef + off;
      bw[c] = h.mcux * h.comp[c].h;
      off += (long long)bw[c] * h.mcuy * h.comp[c].v * 64;
    }
  }
  BitReader br{d, h.scan_pos, len};
  int pred[3] = {0, 0, 0};
  long long count = 0;
  for (int my = 0; my < h.mcuy; ++my)
    for (int mx = 0; mx < h.mcux; ++mx) {
      if (h.restart && count && count % h.restart == 0) {
        br.restart();
        pred[0] = pred[1] = pred[2] = 0;
      }
      ++count;
      for (int c = 0; c < h.ncomp; ++c) {
        const Component& cp = h.comp[c];
        const HuffTable& dct = h.dc[cp.td];
        const HuffTable& act = h.ac[cp.ta];
        for (int by = 0; by < cp.v; ++by)
          for (int bx = 0; bx < cp.h; ++bx) {
            int16_t* blk = base[c] + ((long long)(my * cp.v + by) * bw[c] + mx * cp.h + bx) * 64;
            const int t = decode_symbol(br, dct);
            if (t < 0 || t > 15) return "bad Huffman code";
            pred[c] += t ? extend(br.get(t), t) : 0;
            blk[0] = (int16_t)pred[c];
            for (int k = 1; k < 64;) {
              const int rs = decode_symbol(br, act);
              if (rs < 0) return "bad Huffman code";
              const int r = rs >> 4, s = rs & 15;
              if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
              }
              k += r;
              if (k > 63) return "coefficient index out of range";
              blk[kZigzag[k]] = (int16_t)extend(br.get(s), s);
              ++k;
            }
          }
      }
    }
  return nullptr;
}

}  // namespace

// info: width, height, ncomp, hmax, vmax, mcux, mcuy, reserved; coef_count = int16 coefficients per frame
int jpeg_probe_host(const uint8_t* data, long long len, int* info8, long long* coef_count) {
  Header h;
  if (const char* e = parse_header(data, (size_t)len, h)) return fail(kUnsupported, "jpeg: %s", e);
  info8[0] = h.width; info8[1] = h.height; info8[2] = h.ncomp; info8[3] = h.hmax; info8[4] = h.vmax;
  info8[5] = h.mcux; info8[6] = h.mcuy; info8[7] = 0;
  *coef_count = h.coef_count;
  return kOk;
}

// Entropy-decode n files (all of the geometry of info8) on n_threads host threads.
// coef: [n][coef_stride] int16 (pinned host memory for the upload), qt: [n][3][64] uint16 (natural order).
int jpeg_huffman_decode_host(const uint8_t* const* files, const long long* lens, int n, const int* info8, int16_t* coef,
                             long long coef_stride, uint16_t* qt, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n) n_threads = n;
  std::atomic<int> next(0), bad(-1);
  std::vector<std::string> msgs((size_t)n_threads);
  auto work = [&](int tid) {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n || bad.load() >= 0) return;
      Header h;
      const char* e = parse_header(files[i], (size_t)lens[i], h);
      if (!e && (h.width != info8[0] || h.height != info8[1] || h.ncomp != info8[2] || h.hmax != info8[3] || h.vmax != info8[4]))
        e = "frames of one batch must share size and chroma sampling";
      if (!e && h.coef_count > coef_stride) e = "coefficient buffer too small";
      if (!e) e = decode_scan(files[i], (size_t)lens[i], h, coef + (long long)i * coef_stride);
      if (e) {
        msgs[tid] = e;
        int expect = -1;
        bad.compare_exchange_strong(expect, i);
        return;
      }
      for (int c = 0; c < 3; ++c)
        memcpy(qt + ((long long)i * 3 + c) * 64, h.qt[h.comp[c < h.ncomp ? c : 0].tq], 64 * sizeof(uint16_t));
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& t : pool) t.join();
  if (bad.load() >= 0) {
    for (auto& m : msgs)
      if (!m.empty()) return fail(kUnsupported, "jpeg: frame %d: %s", bad.load(), m.c_str());
    return fail(kUnsupported, "jpeg: frame %d failed", bad.load());
  }
  return kOk;
}

// ======================================================================================== device
namespace {

// jidctint.c constants (CONST_BITS = 13)
constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270,
              F_0_899976223 = 7373, F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137,
              F_1_961570560 = 16069, F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;

// One 1-D islow pass on 8 values (in place), outputs descaled by `shift` with rounding.
__device__ __forceinline__ void idct8(int (&v)[8], int shift) {
  int z2 = v[2], z3 = v[6];
  int z1 = (z2 + z3) * F_0_541196100;
  const int tmp2 = z1 + z3 * (-F_1_847759065);
  const int tmp3 = z1 + z2 * F_0_765366865;
  z2 = v[0]; z3 = v[4];
  const int tmp0 = (z2 + z3) << 13;
  const int tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int t0 = v[7], t1 = v[5], t2 = v[3], t3 = v[1];
  z1 = t0 + t3; z2 = t1 + t2; z3 = t0 + t2;
  int z4 = t1 + t3;
  const int z5 = (z3 + z4) * F_1_175875602;
  t0 *= F_0_298631336; t1 *= F_2_053119869; t2 *= F_3_072711026; t3 *= F_1_501321110;
  z1 *= -F_0_899976223; z2 *= -F_2_562915447;
  z3 = z3 * (-F_1_961570560) + z5;
  z4 = z4 * (-F_0_390180644) + z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  const int rnd = 1 << (shift - 1);
  v[0] = (tmp10 + t3 + rnd) >> shift; v[7] = (tmp10 - t3 + rnd) >> shift;
  v[1] = (tmp11 + t2 + rnd) >> shift; v[6] = (tmp11 - t2 + rnd) >> shift;
  v[2] = (tmp12 + t1 + rnd) >> shift; v[5] = (tmp12 - t1 + rnd) >> shift;
  v[3] = (tmp13 + t0 + rnd) >> shift; v[4] = (tmp13 - t0 + rnd) >> shift;
}

struct JpegGeom {
  int n, width, height, ncomp, hmax, vmax;
  int bw[3], bh[3];              // blocks per row / column of each component (padded to whole MCUs)
  long long coef_off[3];         // int16 offset of each component inside a frame's coefficients
  long long coef_stride;         // int16 per frame
  long long plane_off[3];        // byte offset of each component's plane inside a frame's planes
  long long plane_stride;        // bytes per frame
  long long blocks_per_frame;
};

// dequantise + islow IDCT + level shift + range limit: 8 threads per 8x8 block (a column in pass 1, a row in pass 2)
__global__ void __launch_bounds__(256)
jpeg_idct_kernel(const int16_t* __restrict__ coef, const uint16_t* __restrict__ qt, uint8_t* __restrict__ planes, JpegGeom g) {
  __shared__ int ws[32][8][9];
  const int lb = threadIdx.x >> 3, k = threadIdx.x & 7;
  const long long blk = (long long)blockIdx.x * 32 + lb;
  const long long total = g.blocks_per_frame * g.n;
  const bool active = blk < total;
  int frame = 0, comp = 0, by = 0, bx = 0;
  if (active) {
    frame = (int)(blk / g.blocks_per_frame);
    long long r = blk - (long long)frame * g.blocks_per_frame;
    while (comp + 1 < g.ncomp && r >= (long long)g.bw[comp] * g.bh[comp]) { r -= (long long)g.bw[comp] * g.bh[comp]; ++comp; }
    by = (int)(r / g.bw[comp]);
    bx = (int)(r - (long long)by * g.bw[comp]);
    const int16_t* c = coef + (long long)frame * g.coef_stride + g.coef_off[comp] + ((long long)by * g.bw[comp] + bx) * 64;
    const uint16_t* q = qt + ((long long)frame * 3 + comp) * 64;
    int v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (int)c[i * 8 + k] * (int)q[i * 8 + k];     // column k
    idct8(v, 13 - 2);                                                               // pass 1, scaled up by 2^PASS1_BITS
#pragma unroll
    for (int i = 0; i < 8; ++i) ws[lb][i][k] = v[i];
  }
  __syncthreads();
  if (active) {
    int v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ws[lb][k][i];                               // row k
    idct8(v, 13 + 2 + 3);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int idx = v[i] & 1023;                          // RANGE_MASK, then the wrap-around range-limit table
      idx = idx < 512 ? idx : idx - 1024;
      const uint32_t s = (uint32_t)min(255, max(0, idx + 128));
      if (i < 4) lo |= s << (8 * i); else hi |= s << (8 * (i - 4));
    }
    uint8_t* dst = planes + (long long)frame * g.plane_stride + g.plane_off[comp] +
                   ((long long)(by * 8 + k) * g.bw[comp] + bx) * 8;
    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
  }
}



// fancy upsampling (jdsample.c) of one chroma sample position + jdcolor.c conversion: one thread per output pixel
__global__ void __launch_bounds__(256)
jpeg_rgb_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ rgb, JpegGeom g) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)g.width * g.height;
  if (idx >= per * g.n) return;
  const int frame = (int)(idx / per);
  const int rem = (int)(idx - (long long)frame * per);
  const int y = rem / g.width, x = rem - y * g.width;
  const uint8_t* P = planes + (long long)frame * g.plane_stride;
  const int yw = g.bw[0] * 8;
  const int Y = P[g.plane_off[0] + (long long)y * yw + x];
  int R, G, B;
  if (g.ncomp == 1) {
    R = G = B = Y;
  } else {
    const int cwp = g.bw[1] * 8;                                        // padded chroma row pitch
    const int cw = (g.width + g.hmax - 1) / g.hmax, ch = (g.height + g.vmax - 1) / g.vmax;   // real chroma size
    int c[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint8_t* C = P + g.plane_off[1 + k];
      if (g.hmax == 1) {
        c[k] = C[(long long)y * cwp + x];
      } else if (g.vmax == 1) {                                         // h2v1_fancy_upsample
        const int xi = x >> 1;
        const int v = C[(long long)y * cwp + xi];
        if (cw == 1 || (x == 0) || (x == 2 * cw - 1)) c[k] = v;
        else if (x & 1) c[k] = (v * 3 + C[(long long)y * cwp + xi + 1] + 2) >> 2;
        else c[k] = (v * 3 + C[(long long)y * cwp + xi - 1] + 1) >> 2;
      } else {                                                          // h2v2_fancy_upsample
        const int yi = y >> 1, xi = x >> 1;
        const int yo = (y & 1) ? min(yi + 1, ch - 1) : max(yi - 1, 0);  // jdmainct.c: edge rows replicated
        const uint8_t* r0 = C + (long long)yi * cwp;
        const uint8_t* r1 = C + (long long)yo * cwp;
        const int cs = r0[xi] * 3 + r1[xi];
        if (x & 1) {
          if (xi == cw - 1) c[k] = (cs * 4 + 7) >> 4;
          else c[k] = (cs * 3 + (r0[xi + 1] * 3 + r1[xi + 1]) + 7) >> 4;
        } else {
          if (xi == 0) c[k] = (cs * 4 + 8) >> 4;
          else c[k] = (cs * 3 + (r0[xi - 1] * 3 + r1[xi - 1]) + 8) >> 4;
        }
      }
    }
    const int cb = c[0] - 128, cr = c[1] - 128;
    R = Y + ((91881 * cr + 32768) >> 16);                               // FIX(1.40200)
    G = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);                 // FIX(0.34414), FIX(0.71414)
    B = Y + ((116130 * cb + 32768) >> 16);                              // FIX(1.77200)
    R = min(255, max(0, R)); G = min(255, max(0, G)); B = min(255, max(0, B));
  }
  uint8_t* o = rgb + idx * 3;
  o[0] = (uint8_t)R; o[1] = (uint8_t)G; o[2] = (uint8_t)B;
}

}  // namespace

long long jpeg_plane_bytes(const int* info8) {
  long long b = 0;
  const int ncomp = info8[2], hmax = info8[3], vmax = info8[4], mcux = info8[5], mcuy = info8[6];
  for (int c = 0; c < ncomp; ++c) b += (long long)mcux * (c == 0 ? hmax : 1) * 8 * mcuy * (c == 0 ? vmax : 1) * 8;
  return b;
}

int jpeg_idct_rgb_u8(const int16_t* coef, const uint16_t* qt, uint8_t* planes, uint8_t* rgb, int n, const int* info8,
                     long long coef_stride, cudaStream_t stream) {
  JpegGeom g;
  g.n = n; g.width = info8[0]; g.height = info8[1]; g.ncomp = info8[2]; g.hmax = info8[3]; g.vmax = info8[4];
  const int mcux = info8[5], mcuy = info8[6];
  long long co = 0, po = 0, nb = 0;
  for (int c = 0; c < 3; ++c) {
    const bool luma = c == 0;
    g.bw[c] = mcux * (luma ? g.hmax : 1);
    g.bh[c] = mcuy * (luma ? g.vmax : 1);
    g.coef_off[c] = co; g.plane_off[c] = po;
    if (c < g.ncomp) {
      co += (long long)g.bw[c] * g.bh[c] * 64;
      po += (long long)g.bw[c] * g.bh[c] * 64;
      nb += (long long)g.bw[c] * g.bh[c];
    }
  }
  g.coef_stride = coef_stride; g.plane_stride = po; g.blocks_per_frame = nb;
  const long long blocks = nb * n;
  jpeg_idct_kernel<<<(unsigned)((blocks + 31) / 32), 256, 0, stream>>>(coef, qt, planes, g);
  DV_LAUNCH_OK();
  const long long px = (long long)g.width * g.height * n;
  jpeg_rgb_kernel<<<(unsigned)((px + 255) / 256), 256, 0, stream>>>(planes, rgb, g);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
