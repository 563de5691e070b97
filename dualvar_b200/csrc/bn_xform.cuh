// BatchNorm normalise + ReLU applied to a convolution OPERAND tile in shared memory ("consumer-side fusion").
//
// Reference: x -> nn.BatchNorm3d -> nn.ReLU -> nn.Conv3d inside SpatioTemporalConv (backbone/r21d.py:56-57,67-70):
// the reference (and the stand-alone dv_bn_apply pass) materialises z = relu(scale * y + shift) in HBM between the two
// convolutions. Here the consumer convolution loads the RAW tile of y itself and these routines turn it into z on the
// way into shared memory, in front of tcgen05.mma - so z never exists in HBM (4 B per element less traffic in forward,
// and the weight gradient of the consumer recomputes z the same way instead of reading it).
//
// A tile is a SWIZZLE_128B box: row r = one input position (128 B = 64 consecutive channels), the 16-byte group g of
// row r sits at physical group g ^ (r & 7). Rows outside the tensor (= the convolution's zero padding, and the halo of
// partial tiles) must be zero: relu(shift) is not zero. Channels past the logical count carry scale = shift = 0 in the
// staged table, so they come out zero as well.
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

// Where a box comes from: an NDHWC view (possibly a stride-parity sub-view) of the raw tensor y.
struct XfSrc {
  const uint8_t* base;          // first element of the view
  long long sw, sh, st, sn;     // BYTE strides of the view along w, h, t, n
  int c_bytes;                  // bytes of one position's channel vector (Cp * 2): 16-byte groups beyond it are zero
};

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// FILL a SWIZZLE_128B box (what a TMA box copy of chunk `kc` would have written) with z = relu?(scale*y + shift):
// global -> registers -> (affine, ReLU, bf16) -> shared memory. Measured on B200: transforming a TMA-landed tile IN shared
// memory costs one more read and write of the tile there, and the tensor core's operand fetch already saturates the
// shared-memory port (128 B/clk/SM) on these layers - the in-place variant made the fused kernels 1.5-2.3x slower.
// Loading through registers writes the tile to shared memory exactly once, like TMA does.
// ss_chunk: shared-memory table of THIS 64-channel chunk: scale[64] then shift[64] (fp32).
// Called by `nthreads` threads (a multiple of 64) with tid in [0, nthreads). Rows outside the view (the convolution's
// zero padding, tile overhang) and channel groups beyond the tensor are written as zeros.
__device__ __forceinline__ void bnrelu_box_fill(uint8_t* box, const XfBox& b, const XfSrc& src, int kc, const float* ss_chunk,
                                                int relu, int tid, int nthreads) {
  const int logical = tid & 7;                // 16-byte channel group of the chunk this thread owns
  int r = tid >> 3;
  const int step = nthreads >> 3;             // multiple of 8: (r & 7), hence the swizzled position, is fixed per thread
  const int phys = logical ^ (r & 7);
  float sc[8], sh[8];
  {
    const float4* t4 = reinterpret_cast<const float4*>(ss_chunk + logical * 8);
    const float4 a0 = t4[0], a1 = t4[1], c0 = t4[16], c1 = t4[17];
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sh[0] = c0.x; sh[1] = c0.y; sh[2] = c0.z; sh[3] = c0.w; sh[4] = c1.x; sh[5] = c1.y; sh[6] = c1.z; sh[7] = c1.w;
  }
  const int cbyte = kc * 128 + logical * 16;
  const bool c_ok = cbyte < src.c_bytes;
  const uint8_t* gbase = src.base + cbyte;
  uint8_t* sbase = box + phys * 16;
  const int bw = 1 << b.lw;
  int iw = r & (bw - 1);
  int q = r >> b.lw;
  int ih = q % b.bh; q /= b.bh;
  int it = q % b.bt;
  int in = q / b.bt;
  constexpr int kBatch = 4;
  while (r < b.rows) {
    uint4 v[kBatch];
    bool ok[kBatch];
    int rr[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      rr[j] = r;
      const int w = b.ow + iw, h = b.oh + ih, t = b.ot + it, n = b.on + in;
      ok[j] = r < b.rows && c_ok && (unsigned)w < (unsigned)b.dw && (unsigned)h < (unsigned)b.dh &&
              (unsigned)t < (unsigned)b.dt && (unsigned)n < (unsigned)b.dn;
      v[j] = make_uint4(0u, 0u, 0u, 0u);
      if (ok[j]) v[j] = ldg_nc_v4(gbase + n * src.sn + t * src.st + h * src.sh + w * src.sw);
      r += step;
      iw += step;
      ih += iw >> b.lw;
      iw &= bw - 1;
      while (ih >= b.bh) { ih -= b.bh; ++it; }
      while (it >= b.bt) { it -= b.bt; ++in; }
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      if (rr[j] < b.rows) {
        if (ok[j]) v[j] = bnrelu_vec(v[j], sc, sh, relu);
        *reinterpret_cast<uint4*>(sbase + rr[j] * 128) = v[j];
      }
    }
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
