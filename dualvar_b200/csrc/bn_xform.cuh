// BatchNorm normalise + ReLU applied to a convolution OPERAND tile in shared memory ("consumer-side fusion").
//
// Reference: x -> nn.BatchNorm3d -> nn.ReLU -> nn.Conv3d inside SpatioTemporalConv (backbone/r21d.py:56-57,67-70):
// the reference (and the stand-alone dv_bn_apply pass) materialises z = relu(scale * y + shift) in HBM between the two
// convolutions. Here the consumer convolution's TMA loads the RAW tile of y, and these routines turn it into z in
// place - after the TMA landed, before tcgen05.mma reads it - so z never exists in HBM (4 B per element less traffic
// in forward, and the weight gradient of the consumer recomputes z the same way instead of reading it).
//
// A tile is a SWIZZLE_128B box: row r = one input position (128 B = 64 consecutive channels), the 16-byte group g of
// row r sits at physical group g ^ (r & 7). Rows outside the tensor were zero-filled by TMA (= the convolution's zero
// padding, and the halo of partial tiles): they must stay zero, relu(shift) is not zero. Channels past the logical
// count carry scale = shift = 0 in the staged table, so they stay zero as well.
// The arithmetic is dv_bn_apply's (fmaf(y, scale, shift), max with 0, round to nearest even bf16): bit-identical.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace dv {

struct XfBox {
  int rows;             // rows of the box
  int lw;               // log2 of the box width (always a power of two)
  int bh, bt;           // box extents along h and t (bn = rows / (bw * bh * bt))
  int ow, oh, ot, on;   // coordinates of the box origin in the tensor map
  int dw, dh, dt, dn;   // extents of the tensor map (valid coordinates are 0 .. d-1)
};

__device__ __forceinline__ uint32_t bnrelu_word(uint32_t w, float s0, float b0, float s1, float b1, int relu) {
  const float lo = fmaf(__uint_as_float(w << 16), s0, b0);
  const float hi = fmaf(__uint_as_float(w & 0xffff0000u), s1, b1);
  uint32_t d;
  if (relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));   // max(x, 0) and round in one op
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__device__ __forceinline__ uint4 bnrelu_vec(uint4 v, const float (&sc)[8], const float (&sh)[8], int relu) {
  v.x = bnrelu_word(v.x, sc[0], sh[0], sc[1], sh[1], relu);
  v.y = bnrelu_word(v.y, sc[2], sh[2], sc[3], sh[3], relu);
  v.z = bnrelu_word(v.z, sc[4], sh[4], sc[5], sh[5], relu);
  v.w = bnrelu_word(v.w, sc[6], sh[6], sc[7], sh[7], relu);
  return v;
}

// ss_chunk: shared-memory table of THIS 64-channel chunk: scale[64] then shift[64] (fp32).
// Called by `nthreads` threads (a multiple of 64) with tid in [0, nthreads).
// Fast path (the common case): w and h of the box lie inside the tensor and the box holds one n - the valid rows are
// then ONE contiguous range given by the t extent (the halo planes of a temporal filter), so a row needs two compares;
// rows are processed three at a time with their loads issued together (the loop is latency-, not issue-bound).
__device__ __forceinline__ void bnrelu_box_inplace(uint8_t* box, const XfBox& b, const float* ss_chunk, int relu,
                                                   int tid, int nthreads) {
  const int phys = tid & 7;
  int r = tid >> 3;
  const int step = nthreads >> 3;            // multiple of 8: (r & 7) is the same for every row of this thread
  const int logical = phys ^ (r & 7);
  float sc[8], sh[8];
  {
    const float4* t4 = reinterpret_cast<const float4*>(ss_chunk + logical * 8);
    const float4 a0 = t4[0], a1 = t4[1], c0 = t4[16], c1 = t4[17];
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sh[0] = c0.x; sh[1] = c0.y; sh[2] = c0.z; sh[3] = c0.w; sh[4] = c1.x; sh[5] = c1.y; sh[6] = c1.z; sh[7] = c1.w;
  }
  const int bw = 1 << b.lw;
  uint8_t* base = box + phys * 16;
  const int plane = b.bh << b.lw;                       // rows per t plane
  const bool one_n = b.rows == plane * b.bt;
  const bool wh_inside = b.ow >= 0 && b.ow + bw <= b.dw && b.oh >= 0 && b.oh + b.bh <= b.dh;
  if (one_n && wh_inside) {
    if ((unsigned)b.on >= (unsigned)b.dn) return;        // the whole box lies outside (tail tile of a CTA pair)
    const int t_lo = max(0, -b.ot), t_hi = min(b.bt, b.dt - b.ot);
    const int r_lo = t_lo * plane, r_hi = t_hi * plane;
    // first row of this thread inside the range
    if (r < r_lo) r += (r_lo - r + step - 1) / step * step;
    for (; r + 2 * step < r_hi; r += 3 * step) {
      uint4* p0 = reinterpret_cast<uint4*>(base + r * 128);
      uint4* p1 = reinterpret_cast<uint4*>(base + (r + step) * 128);
      uint4* p2 = reinterpret_cast<uint4*>(base + (r + 2 * step) * 128);
      const uint4 v0 = *p0, v1 = *p1, v2 = *p2;
      *p0 = bnrelu_vec(v0, sc, sh, relu);
      *p1 = bnrelu_vec(v1, sc, sh, relu);
      *p2 = bnrelu_vec(v2, sc, sh, relu);
    }
    for (; r < r_hi; r += step) {
      uint4* p0 = reinterpret_cast<uint4*>(base + r * 128);
      *p0 = bnrelu_vec(*p0, sc, sh, relu);
    }
    return;
  }
  // general path: row -> (iw, ih, it, in) kept incrementally, every coordinate checked
  int iw = r & (bw - 1);
  int q = r >> b.lw;
  int ih = q % b.bh; q /= b.bh;
  int it = q % b.bt;
  int in = q / b.bt;
  for (; r < b.rows; r += step) {
    const bool ok = (unsigned)(b.ow + iw) < (unsigned)b.dw && (unsigned)(b.oh + ih) < (unsigned)b.dh &&
                    (unsigned)(b.ot + it) < (unsigned)b.dt && (unsigned)(b.on + in) < (unsigned)b.dn;
    if (ok) {
      uint4* p = reinterpret_cast<uint4*>(base + r * 128);
      *p = bnrelu_vec(*p, sc, sh, relu);
    }
    // (rows outside the tensor: TMA wrote zeros, nothing to do)
    iw += step;
    ih += iw >> b.lw;
    iw &= bw - 1;
    while (ih >= b.bh) { ih -= b.bh; ++it; }
    while (it >= b.bt) { it -= b.bt; ++in; }
  }
}

// Stage the scale / shift table of a BatchNorm into shared memory as [chunk][scale 64 | shift 64], zero beyond Cp.
// ss: global fp32 [2][Cp] (scale, then shift), as dv_bn_finalize writes it.
__device__ __forceinline__ void stage_ss_table(float* table, const float* __restrict__ ss, int Cp, int k_chunks, int tid,
                                               int nthreads) {
  for (int i = tid; i < k_chunks * 128; i += nthreads) {
    const int kc = i >> 7, j = i & 127;
    const int c = kc * 64 + (j & 63);
    table[i] = c < Cp ? ss[(j >> 6) * Cp + c] : 0.f;
  }
}

}  // namespace dv
