// Shared definitions for the tcgen05 implicit-GEMM convolution kernels (conv_fprop.cu, conv_wgrad.cu).
//
// Data layout in HBM (bf16 mode): activations are NDHWC with the channel count padded to a multiple
// of 8 (16 bytes) so every TMA stride is 16-byte aligned; pad channels always hold zeros.
// An im2col row-block is never materialised: an output tile is a (tn x tt x th x tw) = 128-position
// box of the output tensor, and for filter tap (kt,kh,kw) its A operand is the same box of the
// input tensor shifted by the tap offset, fetched by ONE 5-D TMA box copy (out-of-bounds = zero
// padding). Strided convolutions use one tensor map per input parity class so that the box is
// dense in map coordinates.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace dv {

constexpr int kTileM = 128;        // output positions per tile (= TMEM lanes)
constexpr int kChunkK = 64;        // bf16 channels per TMA box row (= 128 B, one swizzle span)
constexpr int kMaxTaps = 168;      // (3,7,7) stem = 147 taps; fp32 mode: 6 plane products x 27 taps = 162 virtual taps
constexpr int kMaxAMaps = 12;      // 2 x 2 x 2 stride parities; fp32 mode: (parity class, split plane) pairs, e.g. 4 x 3
constexpr int kMaxBlockN = 256;

struct Tap {
  int8_t map;    // which a_map (stride parity class)
  int8_t dt, dh, dw;  // A box origin offset in map coordinates (of the group leader = first tap of a group)
  int16_t widx;  // tap index in the packed weight tensor
  int16_t shift_rows;  // row offset of this tap's 128 rows inside its group's (halo) A box, multiple of 8
};

struct TileGeom {
  // tile box (powers of two, product 128) as log2
  int lw, lh, lt, ln;
  // number of tiles along each dim and output extents (for row validity)
  int tiles_w, tiles_h, tiles_t, tiles_n;
  int ext_w, ext_h, ext_t, ext_n;
  int org_h;   // h offset of the output region in the full tensor (input box coordinates), 0 unless the map is split
  int step_h;  // distance of consecutive tiles along h: 1 << lh, or the 14 output rows of a kStack tile (16-row box)
};

struct alignas(64) ConvTileParams {
  CUtensorMap a_map[kMaxAMaps];
  CUtensorMap b_map;     // weights [rows=Cout_p][taps][Cin_p], box (64, 1, block_n)
  CUtensorMap out_map;   // output NDHWC, box (64, tw, th, tt, tn), 128B swizzle
  Tap taps[kMaxTaps];
  uint8_t group_len[kMaxTaps];  // taps are grouped in runs; a group shares ONE A box (halo re-use)
  uint32_t prog[kMaxTaps + 1];  // per tap: A offset inside its group's box | weight-tile offset << 16, in 16-byte units
  TileGeom g;
  int num_taps;
  int num_groups;
  int max_group;      // largest group
  int a_stage_bytes;  // smem bytes reserved per stage for the A box (multiple of 1024)
  int a_tx_bytes;     // bytes one A box transfers (rows * 128)
  int b_resident;     // 1: all weight tiles of this CTA's channel tile stay in smem for the CTA's life
  int k_chunks;       // ceil(Cin_p / 64)
  int k_steps_last;   // UMMA K=16 steps in the last chunk (1..4)
  int n_tiles;        // channel tiles
  int block_n;        // columns per full channel tile (multiple of 16, <= 256; multiple of 64 if n_tiles > 1)
  int last_n;         // MMA N of the last channel tile (multiple of 16)
  int stages;         // smem ring depth
  int split;          // 1: kSplit kernel instance - two MMA-issuing threads, two half-width accumulators summed by the epilogue
  int total_tiles;
  double* stats;      // nullable: [2][stats_ld] per-channel sum and sum of squares (of the stored bf16)
  int stats_ld;
  const float* bias;  // nullable: per output channel, added before rounding
  long long* prof;    // nullable debug buffer: [grid][8] cycle counters per role (see conv_fprop.cu)
  // BatchNorm-backward reduce fused into a dgrad epilogue (nullable): the tile this kernel stores is the
  // gradient dz of z = relu?(scale*y + shift); `stats` then receives sum(g) and sum(g*y) with
  // g = dz * (scale*y + shift > 0) instead of sum / sum of squares. red_y is y in the output view's layout.
  const void* red_y;
  const float* red_ss;        // nullable [2][stats_ld] scale, shift: ReLU mask (NULL = no ReLU)
  long long red_stride[4];    // element strides of the output view along w, h, t, n (always set)
  long long red_bitoff[7];    // element offset contributed by bit k of the tile row index (w bits, then h, t, n)
  // fp32 mode (nullable): the output view is an fp32 NDHWC tensor and the tile is ADDED to it from registers
  // (red.global.add.v4.f32 per row) instead of being rounded to bf16 and stored by TMA; several launches on
  // bf16 split planes of the operands (x = x0 + x1 + x2, w = w0 + w1 + w2) then sum to an fp32-accurate result.
  // out_map, the staging buffers and the BatchNorm statistics are unused in this mode.
  float* out_f32;
  int f32_store;   // 1: the first product of a sum overwrites the destination (plain 16-byte stores), 0: adds to it
  // Consumer-side BatchNorm (bn_xform.cuh, kernel instance kXf): the A operand is the RAW output y of the convolution
  // below; four extra warps turn every A box into z = relu?(scale*y + shift) in shared memory between the TMA and the
  // MMA, so z is never written to HBM (reference: BatchNorm3d + ReLU between the two convs of SpatioTemporalConv).
  const float* xf_ss;        // fp32 [2][xf_cp] scale, shift of the BatchNorm in front of this convolution
  int xf_cp, xf_relu;
  int a_box[4];              // A box extents along w (power of two), h, t, n
  int a_dims[kMaxAMaps][4];  // W, H, T, N extents of every A tensor map (rows outside stay zero = conv padding)
};

// Optional fused BatchNorm-backward reduction request for conv_dgrad_bf16 (see ConvTileParams::red_y).
struct BnReduce {
  const void* y;      // bf16 [N][T][H][W][Cin_p], the raw conv output that BN normalised to produce dgrad's dx tensor
  const float* ss;    // nullable [2][Cin_p] forward scale/shift (ReLU mask); NULL when the activation had no ReLU
  double* sums;       // [2][Cin_p] double, caller zeroes: sum(g), sum(g*y)
};

// Convolution geometry shared by fprop / dgrad / wgrad host code (padded channel counts).
struct ConvGeom {
  int N, T, H, W, Cin_p;   // input
  int To, Ho, Wo, Cout_p;  // output
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
};

}  // namespace dv
