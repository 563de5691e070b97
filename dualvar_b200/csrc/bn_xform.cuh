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

// ss_chunk: shared-memory table of THIS 64-channel chunk: scale[64] then shift[64] (fp32).
// Called by `nthreads` threads (a multiple of 64) with tid in [0, nthreads).
__device__ __forceinline__ void bnrelu_box_inplace(uint8_t* box, const XfBox& b, const float* ss_chunk, int relu,
                                                   int tid, int nthreads) {
  const int phys = tid & 7;
  int r = tid >> 3;
  const int step = nthreads >> 3;            // multiple of 8: (r & 7) is the same for every row of this thread
  const int logical = phys ^ (r & 7);
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = ss_chunk[logical * 8 + i];
    sh[i] = ss_chunk[64 + logical * 8 + i];
  }
  const int bw = 1 << b.lw;
  // row -> (iw, ih, it, in) kept incrementally
  int iw = r & (bw - 1);
  int q = r >> b.lw;
  int ih = q % b.bh; q /= b.bh;
  int it = q % b.bt;
  int in = q / b.bt;
  for (; r < b.rows; r += step) {
    uint4* p = reinterpret_cast<uint4*>(box + r * 128 + phys * 16);
    const bool ok = (unsigned)(b.ow + iw) < (unsigned)b.dw && (unsigned)(b.oh + ih) < (unsigned)b.dh &&
                    (unsigned)(b.ot + it) < (unsigned)b.dt && (unsigned)(b.on + in) < (unsigned)b.dn;
    if (ok) {
      uint4 v = *p;
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo = __uint_as_float(w[j] << 16);
        float hi = __uint_as_float(w[j] & 0xffff0000u);
        lo = fmaf(lo, sc[2 * j], sh[2 * j]);
        hi = fmaf(hi, sc[2 * j + 1], sh[2 * j + 1]);
        if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
        w[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      *p = make_uint4(w[0], w[1], w[2], w[3]);
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
