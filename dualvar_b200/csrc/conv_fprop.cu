// tcgen05 / TMEM / TMA implicit-GEMM 3-D convolution: forward (fprop) and data gradient (dgrad).
//
// Replaces the cuDNN calls behind nn.Conv3d forward and its autograd dgrad on the reference path
// (reference: backbone/r21d.py:54,64  backbone/r3d.py:33  backbone/c3d.py:15-44
//  backbone/s3dg.py:11,39-41).
//
// One kernel, "multi-tap tile GEMM":
//   Out[tile(128 positions), n0:n0+N] = sum_taps sum_kchunks  A_tap[128 x 64] * W_tap[N x 64]^T
// warp 0   : TMA producer (A box of the shifted input + weight box per (tap, k-chunk) stage)
// warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (fp32 accumulators in TMEM,
//            two accumulator buffers so the epilogue of tile i overlaps the main loop of tile i+1)
// warps 2-5: epilogue: tcgen05.ld -> (+bias) -> bf16 -> swizzled smem -> TMA store, plus the
//            per-channel sum / sum-of-squares of the stored values for training-mode BatchNorm
//            (reference: nn.BatchNorm3d after every conv, e.g. backbone/r21d.py:56,106,111).
// The grid is persistent (<= one CTA per SM); a CTA keeps one channel tile for its whole life so
// BN partial sums stay in shared memory and are flushed to HBM once per CTA.
#include "conv_tile.cuh"
#include "host_common.h"
#include "ptx.cuh"

#include <cuda_bf16.h>

#include <vector>

namespace dv {

constexpr int kNumThreads = 192;
constexpr int kAStageBytes = kTileM * 128;  // 16 KB
constexpr int kOutBufBytes = kTileM * 128;  // one 64-channel chunk of the output tile
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kSmemBudget = 232448 - 4096;  // 227 KB minus static smem / alignment slack

__global__ void __launch_bounds__(kNumThreads, 1)
conv_tile_kernel(const __grid_constant__ ConvTileParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_sum[kMaxBlockN];
  __shared__ float s_sq[kMaxBlockN];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // carve dynamic smem (1024-byte aligned for the 128B swizzle atoms)
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int b_stage_bytes = p.block_n * 128;
  uint8_t* a_smem = smem;                                   // stages * 16 KB
  uint8_t* b_smem = a_smem + p.stages * kAStageBytes;       // stages * block_n * 128
  uint8_t* o_smem = b_smem + p.stages * b_stage_bytes;      // 2 * 16 KB

  const int n_tile = blockIdx.x % p.n_tiles;  // grid is a multiple of n_tiles
  const int bn_mma = (n_tile == p.n_tiles - 1) ? p.last_n : p.block_n;
  const int num_k_iters = p.num_taps * p.k_chunks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    fence_barrier_init();
  }
  for (int c = threadIdx.x; c < kMaxBlockN; c += kNumThreads) {
    s_sum[c] = 0.f;
    s_sq[c] = 0.f;
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  const TileGeom& g = p.g;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < kMaxAMaps; ++i) tma_prefetch_desc(&p.a_map[i]);
      tma_prefetch_desc(&p.b_map);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = kAStageBytes + b_stage_bytes;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int m_id = tile / p.n_tiles;
        const int wb = m_id % g.tiles_w; m_id /= g.tiles_w;
        const int hb = m_id % g.tiles_h; m_id /= g.tiles_h;
        const int tb = m_id % g.tiles_t; m_id /= g.tiles_t;
        const int nb = m_id;
        const int w0 = wb << g.lw, h0 = hb << g.lh, t0 = tb << g.lt, n0 = nb << g.ln;
        const int bcol = n_tile * p.block_n;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const Tap tp = p.taps[tap];
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_5d(a_smem + stage * kAStageBytes, &p.a_map[tp.map], &full_bar[stage],
                        kc * kChunkK, w0 + tp.dw, h0 + tp.dh, t0 + tp.dt, n0);
            tma_load_3d(b_smem + stage * b_stage_bytes, &p.b_map, &full_bar[stage], kc * kChunkK,
                        tp.widx, bcol);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kTileM, bn_mma, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kMaxBlockN;
        int kiter = 0;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          for (int kc = 0; kc < p.k_chunks; ++kc, ++kiter) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after_sync();
            const uint64_t adesc =
                make_smem_desc(smem_u32(a_smem + stage * kAStageBytes), 16, 1024);
            const uint64_t bdesc =
                make_smem_desc(smem_u32(b_smem + stage * b_stage_bytes), 16, 1024);
            const int ksteps = (kc == p.k_chunks - 1) ? p.k_steps_last : 4;
            for (int k = 0; k < ksteps; ++k) {
              // +32 bytes along K inside the 128B swizzle span = +2 in (addr >> 4) units
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kiter | k) != 0);
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        (void)num_k_iters;
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;     // tile row == TMEM lane
    const int et = threadIdx.x - 64;   // 0..127
    const bool leader = (et == 0);
    const int rw = row & ((1 << g.lw) - 1);
    const int rh = (row >> g.lw) & ((1 << g.lh) - 1);
    const int rt = (row >> (g.lw + g.lh)) & ((1 << g.lt) - 1);
    const int rn = row >> (g.lw + g.lh + g.lt);
    const int nchunks = (bn_mma + 63) >> 6;
    const int bcol = n_tile * p.block_n;
    const bool do_stats = p.stats != nullptr;
    int it = 0;
    uint32_t obuf = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      int m_id = tile / p.n_tiles;
      const int wb = m_id % g.tiles_w; m_id /= g.tiles_w;
      const int hb = m_id % g.tiles_h; m_id /= g.tiles_h;
      const int tb = m_id % g.tiles_t; m_id /= g.tiles_t;
      const int nb = m_id;
      const int w0 = wb << g.lw, h0 = hb << g.lh, t0 = tb << g.lt, n0 = nb << g.ln;
      const bool valid = (w0 + rw < g.ext_w) && (h0 + rh < g.ext_h) && (t0 + rt < g.ext_t) &&
                         (n0 + rn < g.ext_n);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + acc * kMaxBlockN + (static_cast<uint32_t>(q * 32) << 16);
      for (int cc = 0; cc < nchunks; ++cc, obuf ^= 1) {
        const int ncols = min(64, bn_mma - cc * 64);
        uint8_t* ob = o_smem + obuf * kOutBufBytes;
        // the TMA store that last read this buffer was committed two chunks ago
        if (leader) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        uint8_t* orow = ob + row * 128;
        for (int gi = 0; gi < (ncols >> 4); ++gi) {
          uint32_t v[16];
          tmem_ld16(t_addr + cc * 64 + gi * 16, v);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float f0 = __uint_as_float(v[2 * j]);
            float f1 = __uint_as_float(v[2 * j + 1]);
            if (p.bias != nullptr) {
              const int c = bcol + cc * 64 + gi * 16 + 2 * j;
              f0 += (c < p.stats_ld) ? p.bias[c] : 0.f;
              f1 += (c + 1 < p.stats_ld) ? p.bias[c + 1] : 0.f;
            }
            if (!valid) { f0 = 0.f; f1 = 0.f; }
            __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          const int c16a = (2 * gi) ^ (row & 7);
          const int c16b = (2 * gi + 1) ^ (row & 7);
          *reinterpret_cast<uint4*>(orow + c16a * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(orow + c16b * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        if (cc == nchunks - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before_sync();
          mbar_arrive(&tmem_empty_bar[acc]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (leader) {
          tma_store_5d(&p.out_map, ob, bcol + cc * 64, w0, h0, t0, n0);
          tma_store_commit();
        }
        if (do_stats) {
          // column sums over the stored bf16 tile: thread -> (channel pair, row quarter)
          const int word = et & 31;
          const int rq = et >> 5;
          if (word * 2 < ncols) {
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
            for (int r = rq * 32; r < rq * 32 + 32; ++r) {
              const uint32_t off = r * 128 + ((((word >> 2) ^ (r & 7))) << 4) + ((word & 3) << 2);
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(ob + off);
              const float2 f = __bfloat1622float2(h);
              a0 += f.x; a1 += f.y;
              b0 = fmaf(f.x, f.x, b0); b1 = fmaf(f.y, f.y, b1);
            }
            const int c = cc * 64 + word * 2;
            atomicAdd(&s_sum[c], a0); atomicAdd(&s_sum[c + 1], a1);
            atomicAdd(&s_sq[c], b0);  atomicAdd(&s_sq[c + 1], b1);
          }
        }
      }
    }
    if (leader) tma_store_wait_all<0>();
    if (do_stats) {
      named_bar_sync(1, 128);
      for (int c = et; c < bn_mma; c += 128) {
        const int gc = bcol + c;
        if (gc < p.stats_ld) {
          atomicAdd(&p.stats[gc], static_cast<double>(s_sum[c]));
          atomicAdd(&p.stats[p.stats_ld + gc], static_cast<double>(s_sq[c]));
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ======================================================================== host side

struct View5 {
  // NDHWC view, innermost-first: dims {C, W, H, T, N}; strides in elements
  const void* base;
  long long dim[5];
  long long stride[5];
};

static View5 make_ndhwc(const void* ptr, int N, int T, int H, int W, int Cp) {
  View5 v;
  v.base = ptr;
  v.dim[0] = Cp; v.dim[1] = W; v.dim[2] = H; v.dim[3] = T; v.dim[4] = N;
  v.stride[0] = 1;
  v.stride[1] = Cp;
  v.stride[2] = (long long)W * Cp;
  v.stride[3] = (long long)H * W * Cp;
  v.stride[4] = (long long)T * H * W * Cp;
  return v;
}

// sub-sample dim d: start r, step s
static void subsample(View5& v, int d, int r, int s) {
  v.base = static_cast<const uint8_t*>(v.base) + (long long)r * v.stride[d] * 2;
  v.dim[d] = (v.dim[d] - r + s - 1) / s;
  v.stride[d] *= s;
}

static int encode_view(CUtensorMap* m, const View5& v, const uint32_t box[5]) {
  uint64_t dims[5], strides[5];
  for (int i = 0; i < 5; ++i) {
    dims[i] = (uint64_t)v.dim[i];
    strides[i] = (uint64_t)v.stride[i] * 2;
  }
  return encode_tmap(m, v.base, 2, 5, dims, strides, box, /*swizzle128=*/true);
}

static int ilog2(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

// Choose a 128-position tile box (tn,tt,th,tw powers of two) minimising padded volume.
void choose_tile(int N, int T, int H, int W, int* ln, int* lt, int* lh, int* lw) {
  double best = 1e30;
  for (int a = 0; a <= 7; ++a)          // w
    for (int b = 0; a + b <= 7; ++b)    // h
      for (int c = 0; a + b + c <= 7; ++c) {  // t
        const int d = 7 - a - b - c;          // n
        const int tw = 1 << a, th = 1 << b, tt = 1 << c, tn = 1 << d;
        const double vol = (double)round_up(W, tw) * round_up(H, th) * round_up(T, tt) * round_up(N, tn);
        // prefer longer contiguous runs along W on ties (fewer, larger TMA row segments)
        const double cost = vol * (1.0 + 0.02 * (7 - a) + 0.004 * (7 - a - b));
        if (cost < best) { best = cost; *lw = a; *lh = b; *lt = c; *ln = d; }
      }
}

static void pick_block_n(int rows_p, int* n_tiles, int* block_n, int* last_n) {
  if (rows_p <= kMaxBlockN) {
    *n_tiles = 1;
    *block_n = round_up(rows_p, 16);
    *last_n = *block_n;
    return;
  }
  // several channel tiles: tile starts must be multiples of 64 (epilogue stores 64-channel chunks)
  int best_bn = 256; double best_cost = 1e30;
  for (int bn = 128; bn <= 256; bn += 64) {
    const int nt = ceil_div(rows_p, bn);
    const int last = round_up(rows_p - (nt - 1) * bn, 16);
    // cost ~ MMA columns issued + a per-tile overhead for re-streaming A
    const double cost = (double)(nt - 1) * bn + last + 24.0 * nt;
    if (cost < best_cost) { best_cost = cost; best_bn = bn; }
  }
  *block_n = best_bn;
  *n_tiles = ceil_div(rows_p, best_bn);
  *last_n = round_up(rows_p - (*n_tiles - 1) * best_bn, 16);
}

// Launch the multi-tap tile GEMM: out(view) = sum_taps in(view shifted) * w[tap].
//   in_base : NDHWC view the taps index into (possibly different parity sub-views per tap)
//   taps    : filled by the caller together with the a_maps
static int launch_conv_tiles(ConvTileParams& P, const View5& outv, int out_rows_p,
                             const void* w_packed, int w_rows_p, int w_taps, int kin_p,
                             double* stats, const float* bias, cudaStream_t stream) {
  // tile geometry over the output view
  TileGeom& g = P.g;
  choose_tile((int)outv.dim[4], (int)outv.dim[3], (int)outv.dim[2], (int)outv.dim[1], &g.ln, &g.lt,
              &g.lh, &g.lw);
  g.ext_w = (int)outv.dim[1]; g.ext_h = (int)outv.dim[2]; g.ext_t = (int)outv.dim[3]; g.ext_n = (int)outv.dim[4];
  g.tiles_w = ceil_div(g.ext_w, 1 << g.lw);
  g.tiles_h = ceil_div(g.ext_h, 1 << g.lh);
  g.tiles_t = ceil_div(g.ext_t, 1 << g.lt);
  g.tiles_n = ceil_div(g.ext_n, 1 << g.ln);
  pick_block_n(out_rows_p, &P.n_tiles, &P.block_n, &P.last_n);
  P.k_chunks = ceil_div(kin_p, kChunkK);
  P.k_steps_last = ceil_div(kin_p - (P.k_chunks - 1) * kChunkK, 16);
  const long long m_tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_t * g.tiles_n;
  P.total_tiles = (int)(m_tiles * P.n_tiles);
  const int stage_bytes = kAStageBytes + P.block_n * 128;
  P.stages = (kSmemBudget - 1024 - 2 * kOutBufBytes) / stage_bytes;
  if (P.stages > kMaxStages) P.stages = kMaxStages;
  if (P.stages < 2) return fail(kUnsupported, "conv tile: not enough shared memory for 2 stages");
  P.stats = stats;
  P.stats_ld = out_rows_p;
  P.bias = bias;

  // weights: [rows][taps][kin_p] bf16, box (64, 1, block_n)
  {
    uint64_t dims[3] = {(uint64_t)kin_p, (uint64_t)w_taps, (uint64_t)w_rows_p};
    uint64_t strides[3] = {2, (uint64_t)kin_p * 2, (uint64_t)kin_p * w_taps * 2};
    uint32_t box[3] = {kChunkK, 1, (uint32_t)P.block_n};
    int rc = encode_tmap(&P.b_map, w_packed, 2, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    uint32_t box[5] = {kChunkK, 1u << g.lw, 1u << g.lh, 1u << g.lt, 1u << g.ln};
    int rc = encode_view(&P.out_map, outv, box);
    if (rc) return rc;
  }
  const int smem_bytes = 1024 + P.stages * stage_bytes + 2 * kOutBufBytes;
  static bool attr_set = false;
  if (!attr_set) {
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    attr_set = true;
  }
  int grid = sm_count() / P.n_tiles * P.n_tiles;
  if (grid > P.total_tiles) grid = P.total_tiles;  // total_tiles is a multiple of n_tiles
  conv_tile_kernel<<<grid, kNumThreads, smem_bytes, stream>>>(P);
  DV_LAUNCH_OK();
  return kOk;
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

int conv_fprop_bf16(const void* x, const void* w_packed, void* y, double* stats, const float* bias,
                    const ConvGeom& c, cudaStream_t stream) {
  static thread_local ConvTileParams P;
  const View5 inv = make_ndhwc(x, c.N, c.T, c.H, c.W, c.Cin_p);
  const View5 outv = make_ndhwc(y, c.N, c.To, c.Ho, c.Wo, c.Cout_p);
  int ln, lt, lh, lw;
  choose_tile(c.N, c.To, c.Ho, c.Wo, &ln, &lt, &lh, &lw);
  const uint32_t box[5] = {kChunkK, 1u << lw, 1u << lh, 1u << lt, 1u << ln};
  int map_of_parity[8];
  for (int i = 0; i < 8; ++i) map_of_parity[i] = -1;
  int nmaps = 0, ntaps = 0;
  for (int a = 0; a < c.kt; ++a)
    for (int b = 0; b < c.kh; ++b)
      for (int d = 0; d < c.kw; ++d) {
        const int ot = a - c.pt, oh = b - c.ph, ow = d - c.pw;
        const int rt = posmod(ot, c.st), rh = posmod(oh, c.sh), rw = posmod(ow, c.sw);
        if (rt >= c.T || rh >= c.H || rw >= c.W) continue;  // tap never touches real data
        const int key = (rt * 2 + rh) * 2 + rw;
        if (rt > 1 || rh > 1 || rw > 1) return fail(kUnsupported, "conv stride > 2 not supported");
        if (map_of_parity[key] < 0) {
          View5 v = inv;
          subsample(v, 3, rt, c.st);
          subsample(v, 2, rh, c.sh);
          subsample(v, 1, rw, c.sw);
          int rc = encode_view(&P.a_map[nmaps], v, box);
          if (rc) return rc;
          map_of_parity[key] = nmaps++;
        }
        if (ntaps >= kMaxTaps) return fail(kUnsupported, "too many filter taps (%d)", c.kt * c.kh * c.kw);
        Tap& tp = P.taps[ntaps++];
        tp.map = (int8_t)map_of_parity[key];
        tp.dt = (int8_t)floordiv(ot, c.st);
        tp.dh = (int8_t)floordiv(oh, c.sh);
        tp.dw = (int8_t)floordiv(ow, c.sw);
        tp.widx = (int16_t)((a * c.kh + b) * c.kw + d);
        tp.pad_ = 0;
      }
  for (int i = nmaps; i < kMaxAMaps; ++i) P.a_map[i] = P.a_map[0];
  P.num_taps = ntaps;
  if (ntaps == 0) return fail(kBadArg, "convolution has no valid taps");
  return launch_conv_tiles(P, outv, c.Cout_p, w_packed, c.Cout_p, c.kt * c.kh * c.kw, c.Cin_p,
                           stats, bias, stream);
}

// dX = dgrad(dY, W): one launch per stride-parity class of dX positions; each class is a
// stride-1 multi-tap GEMM over dY with the subset of taps that reach it.
// w_packed_t: [Cin_p][taps][Cout_p] bf16 (transposed pack).
int conv_dgrad_bf16(const void* dy, const void* w_packed_t, void* dx, const ConvGeom& c,
                    cudaStream_t stream) {
  static thread_local ConvTileParams P;
  const View5 dyv = make_ndhwc(dy, c.N, c.To, c.Ho, c.Wo, c.Cout_p);
  const View5 dxv = make_ndhwc(dx, c.N, c.T, c.H, c.W, c.Cin_p);
  bool need_zero = false;
  for (int rt = 0; rt < c.st; ++rt)
    for (int rh = 0; rh < c.sh; ++rh)
      for (int rw = 0; rw < c.sw; ++rw) {
        int cnt = 0;
        for (int a = 0; a < c.kt; ++a) if (posmod(rt + c.pt - a, c.st) == 0)
          for (int b = 0; b < c.kh; ++b) if (posmod(rh + c.ph - b, c.sh) == 0)
            for (int d = 0; d < c.kw; ++d) if (posmod(rw + c.pw - d, c.sw) == 0) ++cnt;
        if (cnt == 0) need_zero = true;
      }
  if (need_zero)
    DV_CUDA_OK(cudaMemsetAsync(dx, 0, (size_t)c.N * c.T * c.H * c.W * c.Cin_p * 2, stream));
  for (int rt = 0; rt < c.st; ++rt)
    for (int rh = 0; rh < c.sh; ++rh)
      for (int rw = 0; rw < c.sw; ++rw) {
        if (rt >= c.T || rh >= c.H || rw >= c.W) continue;
        View5 ov = dxv;
        subsample(ov, 3, rt, c.st);
        subsample(ov, 2, rh, c.sh);
        subsample(ov, 1, rw, c.sw);
        int ln, lt, lh, lw;
        choose_tile((int)ov.dim[4], (int)ov.dim[3], (int)ov.dim[2], (int)ov.dim[1], &ln, &lt, &lh, &lw);
        const uint32_t box[5] = {kChunkK, 1u << lw, 1u << lh, 1u << lt, 1u << ln};
        int ntaps = 0;
        for (int a = 0; a < c.kt; ++a) {
          if (posmod(rt + c.pt - a, c.st) != 0) continue;
          for (int b = 0; b < c.kh; ++b) {
            if (posmod(rh + c.ph - b, c.sh) != 0) continue;
            for (int d = 0; d < c.kw; ++d) {
              if (posmod(rw + c.pw - d, c.sw) != 0) continue;
              Tap& tp = P.taps[ntaps++];
              tp.map = 0;
              tp.dt = (int8_t)((rt + c.pt - a) / c.st);
              tp.dh = (int8_t)((rh + c.ph - b) / c.sh);
              tp.dw = (int8_t)((rw + c.pw - d) / c.sw);
              tp.widx = (int16_t)((a * c.kh + b) * c.kw + d);
              tp.pad_ = 0;
            }
          }
        }
        if (ntaps == 0) continue;  // class receives no gradient (zero-filled above)
        int rc = encode_view(&P.a_map[0], dyv, box);
        if (rc) return rc;
        for (int i = 1; i < kMaxAMaps; ++i) P.a_map[i] = P.a_map[0];
        P.num_taps = ntaps;
        rc = launch_conv_tiles(P, ov, c.Cin_p, w_packed_t, c.Cin_p, c.kt * c.kh * c.kw, c.Cout_p,
                               nullptr, nullptr, stream);
        if (rc) return rc;
      }
  return kOk;
}

// Stride-2 7x7 stem (Cin <= 4) as a 4-tap (per temporal tap) stride-1 GEMM on the space-to-depth input.
// x_s2d: bf16 [N][T][H2][W2+3][16], channel = (rh*2+rw)*4 + c, two zero columns left / one right
// (written by the ingest kernel). For output (ho,wo) and row tap a in {-2..1} the A row is the 128-byte
// window x_s2d[n][t][ho+a][wo .. wo+3][0..15] — an OVERLAPPING-window tensor map (W stride 32 B, inner
// extent 128 B), so K = 64 per tap instead of 49 taps of K = 16.
// w_stem: bf16 [Cout_p][kt*4][64] (pack_stem_weights).
static int encode_stem_map(CUtensorMap* m, const void* x, int N, int T, int H2, int W2, const uint32_t box[5]) {
  const long long W2p = W2 + 3;
  uint64_t dims[5] = {64, (uint64_t)W2, (uint64_t)H2, (uint64_t)T, (uint64_t)N};
  uint64_t strides[5] = {2, 32, (uint64_t)W2p * 32, (uint64_t)H2 * W2p * 32, (uint64_t)T * H2 * W2p * 32};
  return encode_tmap(m, x, 2, 5, dims, strides, box, true);
}

int conv_stem_fprop_bf16(const void* x_s2d, const void* w_stem, void* y, double* stats, const float* bias,
                         int N, int T, int H2, int W2, int Cout_p, int kt, int pt, cudaStream_t stream) {
  static thread_local ConvTileParams P;
  const int To = T + 2 * pt - kt + 1;
  const View5 outv = make_ndhwc(y, N, To, H2, W2, Cout_p);
  int ln, lt, lh, lw;
  choose_tile(N, To, H2, W2, &ln, &lt, &lh, &lw);
  const uint32_t box[5] = {kChunkK, 1u << lw, 1u << lh, 1u << lt, 1u << ln};
  int rc = encode_stem_map(&P.a_map[0], x_s2d, N, T, H2, W2, box);
  if (rc) return rc;
  for (int i = 1; i < kMaxAMaps; ++i) P.a_map[i] = P.a_map[0];
  int ntaps = 0;
  for (int a = 0; a < kt; ++a)
    for (int r = 0; r < 4; ++r) {
      Tap& tp = P.taps[ntaps];
      tp.map = 0; tp.dt = (int8_t)(a - pt); tp.dh = (int8_t)(r - 2); tp.dw = 0;
      tp.widx = (int16_t)ntaps; tp.pad_ = 0;
      ++ntaps;
    }
  P.num_taps = ntaps;
  return launch_conv_tiles(P, outv, Cout_p, w_stem, Cout_p, ntaps, 64, stats, bias, stream);
}

}  // namespace dv
