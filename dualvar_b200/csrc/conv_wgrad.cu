// tcgen05 / TMEM / TMA implicit-GEMM 3-D convolution weight gradient (wgrad).
//
// Replaces cuDNN's wgrad behind autograd of nn.Conv3d on the reference path.
//   dW[co][tap][ci] = sum over output positions m of  dY[m][co] * X[m shifted by tap][ci]
// The contraction runs over output positions, which are the slow dimension of both NDHWC operands,
// so both are fed to tcgen05.mma as MN-major SWIZZLE_128B tiles: a 64-position x 64-channel TMA box
// is exactly one such tile (row = position = K index, 128 B of channels = M/N index).
//   A (M side) = X boxes: "units" (tap, 64-channel chunk of Cin); two units form one M=128 MMA
//   B (N side) = dY boxes: N = up to 256 output channels
//   D          = fp32 accumulators in TMEM, one per unit pair, all sharing the same dY tile
// Split-K over position tiles across gridDim.y; partial results are reduced with fp32 red.global
// into a zero-initialised [Cout_p][taps][Cin_p] buffer.
#include "bn_xform.cuh"
#include "conv_tile.cuh"
#include "host_common.h"
#include "ptx.cuh"

#include <stdlib.h>

#include <algorithm>

namespace dv {

constexpr int kWgThreads = 192;
constexpr int kBoxBytes = 64 * 128;  // one 64-position x 64-channel box
constexpr int kWgMaxUnits = 8;
// Not the whole 227 KB: wgrad runs on a side stream next to the BatchNorm passes of the layers below (engine.py), whose
// blocks need ~17 KB of shared memory each to become resident on the same SM.
constexpr int kWgSmemBudget = 232448 - 2048 - 36 * 1024;

struct alignas(64) WgradParams {
  CUtensorMap a_map[kMaxAMaps];
  CUtensorMap dy_map;
  Tap taps[kMaxTaps];
  TileGeom g;  // 64-position tiles
  int num_taps;
  int k_chunks;     // ceil(Cin_p / 64)
  int total_units;  // num_taps * k_chunks
  int units_per_group;
  int n_tiles, block_n, last_n, acc_stride;
  int pos_tiles, tiles_per_split;
  int stages;
  float* dw;  // [Cout_p][taps_total][Cin_p]
  int cin_p, cout_p, taps_total;
  // halo mode (stride 1): X boxes carry a halo along the filter's slow dimension (t for temporal filters, h for
  // spatial ones); the unit (tap, chunk) is the view of its box shifted by whole 8-row groups, so one box feeds
  // kt (or kh) units. Unit groups (= CTAs along x) own whole boxes.
  int halo;
  int x_box_bytes;         // bytes of one halo box (multiple of 1024)
  int n_groups;            // halo mode: unit groups
  int8_t grp_unit0[9], grp_box0[9];   // first unit / first box of each group (+ end sentinel)
  int8_t box_dt[32], box_dh[32], box_dw[32], box_kc[32];   // halo box origin offset and 64-channel chunk
  int unit_off[32];        // start of each unit inside its group's stage, in 16-byte units
  int16_t unit_widx[32];   // weight tap index of each unit
  int16_t unit_kc[32];     // 64-channel chunk of each unit
  // Consumer-side BatchNorm (bn_xform.cuh): X is the RAW output y of the convolution below and the operand of this
  // weight gradient is z = relu?(scale*y + shift), recomputed in shared memory (NULL: X is used as it is).
  const float* xf_ss;      // fp32 [2][cin_p] scale, shift
  int xf_relu;
  int xb_h, xb_t;          // halo mode: box extents along h and t
  int a_dims[kMaxAMaps][4];   // W, H, T, N extents of every X tensor map (row validity = the convolution's zero padding)
  // fp32 mode, all plane products in one launch: X and dY are stacks of bf16 split planes along N ([K*N][T][H][W][C]);
  // every position tile is contracted once per product (x plane i, dy plane j) into the SAME accumulators, so the
  // fp32 reduction into dw happens once instead of once per product. n_prod = 1: the plain weight gradient.
  int n_prod;
  int prod_nx[6], prod_ny[6];   // batch offset (plane * N) of the product's X / dY plane
  // dY sharing: the two unit groups of a layer are the two CTAs of a cluster, walk the position tiles in lockstep and
  // each fetches half of the dY boxes of a stage, multicast into both (see wgrad_launch)
  int cluster;
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ __align__(8) uint64_t xf_bar[8];   // X boxes of a stage transformed (128 arrivals)
  __shared__ uint32_t tmem_base_slot;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform role index
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.x / p.n_tiles;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int unit0 = p.halo ? p.grp_unit0[group] : group * p.units_per_group;
  const int nu = p.halo ? p.grp_unit0[group + 1] - unit0 : min(p.units_per_group, p.total_units - unit0);
  const int box0 = p.halo ? p.grp_box0[group] : 0;
  const int nbox = p.halo ? p.grp_box0[group + 1] - box0 : 0;
  const int npairs = (nu + 1) >> 1;
  const int bn = (n_tile == p.n_tiles - 1) ? p.last_n : p.block_n;
  const int nbx = (bn + 63) >> 6;  // dY boxes
  const int tile_begin = blockIdx.y * p.tiles_per_split;
  const int tile_end = min(p.pos_tiles, tile_begin + p.tiles_per_split);
  if (tile_begin >= tile_end) return;  // uniform per CTA

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const bool xf = p.xf_ss != nullptr;
  float* xf_table = reinterpret_cast<float*>(smem_raw + (base - raw));           // [k_chunks][scale 64 | shift 64]
  uint8_t* smem = smem_raw + (base - raw) + (xf ? ((p.k_chunks * 512 + 1023) & ~1023) : 0);
  const int a_bytes = p.halo ? p.units_per_group * p.x_box_bytes : p.units_per_group * kBoxBytes;   // halo: units_per_group = max boxes per group
  const int b_bytes = ((p.block_n + 63) >> 6) * kBoxBytes;
  const int stage_bytes = a_bytes + b_bytes;

  const bool clus = p.cluster != 0;
  const uint32_t crank = clus ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], clus ? 2 : 1);   // cluster: a stage is free when BOTH CTAs' MMAs have read it
      mbar_init(&xf_bar[i], 128);
    }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_slot, 512);
    tmem_relinquish();
  }
  if (xf) stage_ss_table(xf_table, p.xf_ss, p.cin_p, p.k_chunks, threadIdx.x, kWgThreads);
  tc_fence_before_sync();
  __syncthreads();
  if (clus) cluster_sync_all();     // the peer's barriers are initialised before anything is multicast to them
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  const TileGeom& g = p.g;

  if (warp == 0) {
    // TMA producer: the whole warp runs the loop (uniform registers feed UTMALDG), one elected lane issues
    const bool issuer = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx = p.halo ? (uint32_t)(nbox * p.x_box_bytes + nbx * kBoxBytes)
                               : (uint32_t)((nu + nbx) * kBoxBytes);
    // plane products outermost, smallest contributions first: the tensor core aligns every addend to the running
    // sum and truncates, so the 2^-16-scale products must meet a small accumulator (measured: interleaved 1e-4, this 1e-7)
    for (int pi = 0; pi < p.n_prod; ++pi) {
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int m_id = tile;
        const int wb = m_id % g.tiles_w; m_id /= g.tiles_w;
        const int hb = m_id % g.tiles_h; m_id /= g.tiles_h;
        const int tb = m_id % g.tiles_t; m_id /= g.tiles_t;
        const int nb = m_id;
        const int w0 = wb << g.lw, h0 = hb << g.lh, t0 = tb << g.lt, n0 = nb << g.ln;
        const int nx = n0 + p.prod_nx[pi], ny = n0 + p.prod_ny[pi];
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (issuer) {
          mbar_expect_tx(&full_bar[stage], tx);
          uint8_t* a_s = smem + stage * stage_bytes;
          uint8_t* b_s = a_s + a_bytes;
          for (int j = 0; j < nbx; ++j) {
            if (!clus)
              tma_load_5d(b_s + j * kBoxBytes, &p.dy_map, &full_bar[stage], n_tile * p.block_n + j * 64, w0, h0, t0, ny);
            else if ((uint32_t)(j & 1) == crank)     // this CTA's share of the dY boxes, delivered to both
              tma_load_5d_mc(b_s + j * kBoxBytes, &p.dy_map, &full_bar[stage], n_tile * p.block_n + j * 64, w0, h0, t0, ny,
                             (uint16_t)3);
          }
          if (p.halo) {
            for (int b = 0; b < nbox; ++b)
              tma_load_5d(a_s + b * p.x_box_bytes, &p.a_map[0], &full_bar[stage], p.box_kc[box0 + b] * 64,
                          w0 + p.box_dw[box0 + b], h0 + p.box_dh[box0 + b], t0 + p.box_dt[box0 + b], nx);
          } else {
            for (int i = 0; i < nu; ++i) {
              const int u = unit0 + i;
              const int tap = u / p.k_chunks;
              const int kc = u - tap * p.k_chunks;
              const Tap tp = p.taps[tap];
              tma_load_5d(a_s + i * kBoxBytes, &p.a_map[tp.map], &full_bar[stage], kc * 64, w0 + tp.dw, h0 + tp.dh,
                          t0 + tp.dt, nx);
            }
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    if (clus) {
      // tail: every stage's last release has arrived (own commit and the peer's multicast one) before this CTA goes on
      // to exit - nothing of the peer may still be in flight towards this CTA's barriers
      for (int i = 0; i < p.stages; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: uniform loop state (descriptor words in uniform registers), one elected lane issues
    const bool issuer = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, bn, 1, 1);
    // MN-major SWIZZLE_128B: 8-position (K) groups 1024 B apart (SBO), 64-channel (M/N) groups one box apart (LBO)
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lo_flags = (uint32_t)(kBoxBytes >> 4) << 16;
    // (start-address field: 14 bits of address >> 4 inside the CTA's own 256 KB window - in a cluster launch the shared
    // window address of a CTA carries its rank in higher bits, which must not spill into the descriptor's LBO field)
    const uint32_t base_enc = (smem_u32(smem) & 0x3ffffu) >> 4, stage_enc = (uint32_t)stage_bytes >> 4;
    const uint32_t a_enc = (uint32_t)a_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    const int n_steps = (tile_end - tile_begin) * p.n_prod;     // one pipeline stage per (position tile, plane product)
    for (int step = 0; step < n_steps; ++step) {
      mbar_wait(xf ? &xf_bar[stage] : &full_bar[stage], phase);
      tc_fence_after_sync();
      const uint32_t a_lo = lo_flags | (base_enc + (uint32_t)stage * stage_enc);
      const uint32_t b_lo = a_lo + a_enc;
      if (issuer) {
        for (int pr = 0; pr < npairs; ++pr) {
          uint32_t al;
          if (p.halo) {
            // unit = shifted view of a halo box; the pair's second 64-channel group starts (LBO) wherever the
            // next unit starts (possibly overlapping the first)
            const uint32_t o0 = (uint32_t)p.unit_off[unit0 + 2 * pr];
            const uint32_t o1 = (2 * pr + 1 < nu) ? (uint32_t)p.unit_off[unit0 + 2 * pr + 1] : o0 + 64u;
            al = ((o1 - o0) << 16) | (base_enc + (uint32_t)stage * stage_enc + o0);
          } else {
            al = a_lo + (uint32_t)pr * (2 * kBoxBytes >> 4);
          }
          const uint32_t dt = tmem_base + pr * p.acc_stride;
          // one UMMA K step = 16 positions = 2048 B = +128 in (addr >> 4) units
          umma_bf16_lohi(dt, al, b_lo, desc_hi, idesc, accumulate);
          umma_bf16_lohi(dt, al + 128, b_lo + 128, desc_hi, idesc, 1);
          umma_bf16_lohi(dt, al + 256, b_lo + 256, desc_hi, idesc, 1);
          umma_bf16_lohi(dt, al + 384, b_lo + 384, desc_hi, idesc, 1);
        }
        if (clus) umma_commit_mc(&empty_bar[stage], (uint16_t)3); else umma_commit(&empty_bar[stage]);
      }
      accumulate = 1;
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (issuer) umma_commit(&acc_bar);
    __syncwarp();
  } else {
    if (xf) {
      // ---- operand transform (warps 2-5 are otherwise idle until the accumulators are complete): every X box of a
      // stage becomes z = relu?(scale*y + shift) in place once its TMA landed; the MMA warp waits for xf_bar
      const int tid = threadIdx.x - 64;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int m_id = tile;
        const int wb = m_id % g.tiles_w; m_id /= g.tiles_w;
        const int hb = m_id % g.tiles_h; m_id /= g.tiles_h;
        const int tb = m_id % g.tiles_t; m_id /= g.tiles_t;
        const int nb = m_id;
        const int w0 = wb << g.lw, h0 = hb << g.lh, t0 = tb << g.lt, n0 = nb << g.ln;
        mbar_wait(&full_bar[stage], phase);
        uint8_t* a_s = smem + stage * stage_bytes;
        XfBox b;
        b.lw = g.lw;
        if (p.halo) {
          b.rows = p.x_box_bytes >> 7; b.bh = p.xb_h; b.bt = p.xb_t;
          b.dw = p.a_dims[0][0]; b.dh = p.a_dims[0][1]; b.dt = p.a_dims[0][2]; b.dn = p.a_dims[0][3];
          for (int i = 0; i < nbox; ++i) {
            b.ow = w0 + p.box_dw[box0 + i]; b.oh = h0 + p.box_dh[box0 + i]; b.ot = t0 + p.box_dt[box0 + i]; b.on = n0;
            bnrelu_box_inplace(a_s + i * p.x_box_bytes, b, xf_table + p.box_kc[box0 + i] * 128, p.xf_relu, tid, 128);
          }
        } else {
          b.rows = 64; b.bh = 1 << g.lh; b.bt = 1 << g.lt;
          for (int i = 0; i < nu; ++i) {
            const int u = unit0 + i;
            const int tap = u / p.k_chunks;
            const int kc = u - tap * p.k_chunks;
            const Tap tp = p.taps[tap];
            b.dw = p.a_dims[tp.map][0]; b.dh = p.a_dims[tp.map][1]; b.dt = p.a_dims[tp.map][2]; b.dn = p.a_dims[tp.map][3];
            b.ow = w0 + tp.dw; b.oh = h0 + tp.dh; b.ot = t0 + tp.dt; b.on = n0;
            bnrelu_box_inplace(a_s + i * kBoxBytes, b, xf_table + kc * 128, p.xf_relu, tid, 128);
          }
        }
        fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
        mbar_arrive(&xf_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(&acc_bar, 0);
    tc_fence_after_sync();
    const int co0 = n_tile * p.block_n;
    for (int pr = 0; pr < npairs; ++pr) {
      const int i = pr * 2 + (row >> 6);
      const bool unit_ok = i < nu;
      int tapw = 0, ci = 0;
      if (unit_ok) {
        int kc;
        if (p.halo) {
          tapw = p.unit_widx[unit0 + i];
          kc = p.unit_kc[unit0 + i];
        } else {
          const int u = unit0 + i;
          const int tap = u / p.k_chunks;
          kc = u - tap * p.k_chunks;
          tapw = p.taps[tap].widx;
        }
        ci = kc * 64 + (row & 63);
      }
      const bool row_ok = unit_ok && ci < p.cin_p;
      float* dst = p.dw + (size_t)tapw * p.cin_p + ci;
      const size_t co_stride = (size_t)p.taps_total * p.cin_p;
      const uint32_t t_addr = tmem_base + pr * p.acc_stride + (static_cast<uint32_t>(q * 32) << 16);
      for (int gi = 0; gi < (bn >> 4); ++gi) {
        uint32_t v[16];
        tmem_ld16(t_addr + gi * 16, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = co0 + gi * 16 + j;
            if (co < p.cout_p) atomicAdd(dst + (size_t)co * co_stride, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (clus) cluster_sync_all();     // neither CTA leaves while the other may still address its shared memory
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ host
void choose_tile_log2(int total_log2, int N, int T, int H, int W, int* ln, int* lt, int* lh, int* lw) {
  double best = 1e30;
  for (int a = 0; a <= total_log2; ++a)
    for (int b = 0; a + b <= total_log2; ++b)
      for (int c = 0; a + b + c <= total_log2; ++c) {
        const int d = total_log2 - a - b - c;
        const int tw = 1 << a, th = 1 << b, tt = 1 << c, tn = 1 << d;
        const double vol = (double)round_up(W, tw) * round_up(H, th) * round_up(T, tt) * round_up(N, tn);
        const double cost = vol * (1.0 + 0.02 * (total_log2 - a) + 0.004 * (total_log2 - a - b));
        if (cost < best) { best = cost; *lw = a; *lh = b; *lt = c; *ln = d; }
      }
}

static int floordiv_w(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int posmod_w(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

static int wgrad_launch(WgradParams& P, int ntaps, int cin_p, int cout_p, int taps_total, float* dw,
                        cudaStream_t stream) {
  TileGeom& g = P.g;
  P.num_taps = ntaps;
  P.k_chunks = ceil_div(cin_p, 64);
  P.total_units = ntaps * P.k_chunks;
  // N side (Cout)
  if (cout_p <= 256) {
    P.n_tiles = 1; P.block_n = round_up(cout_p, 16); P.last_n = P.block_n;
  } else {
    // equal channel tiles (288 -> 144 + 144, not 256 + 32: a narrow last tile re-loads every X box for 1/8 of the work)
    P.n_tiles = ceil_div(cout_p, 256);
    P.block_n = round_up(ceil_div(cout_p, P.n_tiles), 16);
    P.last_n = round_up(cout_p - (P.n_tiles - 1) * P.block_n, 16);
  }
  P.acc_stride = round_up(P.block_n, 32);
  const int nbx = ceil_div(P.block_n, 64);
  int max_pairs = 512 / P.acc_stride;
  int upg = max_pairs * 2;
  int stage_bytes;
  if (P.halo) {
    // caller built the groups; units_per_group carries the largest number of boxes of a group (smem sizing)
    int max_boxes = 0;
    for (int gi = 0; gi < P.n_groups; ++gi) max_boxes = std::max(max_boxes, P.grp_box0[gi + 1] - P.grp_box0[gi]);
    upg = max_boxes;
    stage_bytes = max_boxes * P.x_box_bytes + nbx * kBoxBytes;
  } else {
    if (upg > kWgMaxUnits) upg = kWgMaxUnits;
    // keep at least 3 stages in shared memory
    const int budget = kWgSmemBudget - 1024 - (P.xf_ss != nullptr ? round_up(P.k_chunks * 512, 1024) : 0);
    while (upg > 2 && budget / ((upg + nbx) * kBoxBytes) < 3) upg -= 2;
    if (upg > P.total_units) upg = P.total_units;
    stage_bytes = (upg + nbx) * kBoxBytes;
  }
  P.units_per_group = upg;
  const int groups = P.halo ? P.n_groups : ceil_div(P.total_units, upg);
  const int xf_bytes = P.xf_ss != nullptr ? round_up(P.k_chunks * 512, 1024) : 0;
  P.stages = (kWgSmemBudget - 1024 - xf_bytes) / stage_bytes;
  if (P.stages > 8) P.stages = 8;
  if (P.stages < 2) return fail(kUnsupported, "wgrad: stage too large for shared memory");
  P.pos_tiles = g.tiles_w * g.tiles_h * g.tiles_t * g.tiles_n;
  const int items = groups * P.n_tiles;
  // CTAs per SM over the kernel's life ("waves"): more, shorter CTAs let kernels of the high-priority main stream get an SM
  // sooner when this kernel runs on the weight-gradient side stream, at the price of more fp32 reductions (DV_WGRAD_WAVES)
  static int waves = -1;
  if (waves < 0) { const char* e = getenv("DV_WGRAD_WAVES"); waves = e ? atoi(e) : 2; if (waves < 1) waves = 1; }
  int ksplit = (waves * sm_count()) / items;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > P.pos_tiles) ksplit = P.pos_tiles;
  P.tiles_per_split = ceil_div(P.pos_tiles, ksplit);
  ksplit = ceil_div(P.pos_tiles, P.tiles_per_split);
  P.dw = dw;
  P.cin_p = cin_p; P.cout_p = cout_p; P.taps_total = taps_total;

  static bool attr_set = false;
  if (!attr_set) {
    DV_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kWgSmemBudget));
    attr_set = true;
  }
  const int smem_bytes = 1024 + xf_bytes + P.stages * stage_bytes;
  dim3 grid(items, ksplit);
  // Two unit groups of one channel tile (the 64->144 3x3 layers: 2 + 1 halo boxes): launched side by side they drift apart
  // (44 against 34 KB per position tile), the second read of dY misses L2 and the launch moves 1.8x its algorithmic bytes
  // through HBM (profiles/r02d_bench_n1.json) next to the HBM-bound BatchNorm passes. As a cluster of 2 they stay in
  // lockstep and fetch every dY box once for both. Measured (profiles/r02d_wgrad_cluster_multicast.txt): the launch alone
  // is 2-4 % slower (the smaller group waits for the larger one), the step is a tie (913.6 vs 914.4 samples/s) - the HBM
  // relief pays for the lockstep and no more - so it is opt-in: DV_WGRAD_CLUSTER=1 (read per call).
  const char* ce = getenv("DV_WGRAD_CLUSTER");
  const int cluster_env = ce ? atoi(ce) : 0;
  P.cluster = (cluster_env && P.halo && groups == 2 && P.n_tiles == 1 && P.xf_ss == nullptr && P.n_prod == 1) ? 1 : 0;
  if (P.cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kWgThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    DV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_wgrad_kernel, P));
  } else {
    conv_wgrad_kernel<<<grid, kWgThreads, smem_bytes, stream>>>(P);
  }
  DV_LAUNCH_OK();
  return kOk;
}

// fp32 mode (conv_wgrad_f32planes): number of split planes stacked along N in both operands, 1 = plain tensors
static thread_local int t_wg_planes = 1;

static void set_products(WgradParams& P, int K, int N) {
  P.n_prod = 0;
  for (int s = 2 * (K - 1); s >= 0; --s)        // products with i + j < K, smallest contributions first
    for (int i = 0; i < K; ++i) {
      const int j = s - i;
      if (j < 0 || j >= K - i) continue;
      P.prod_nx[P.n_prod] = i * N;
      P.prod_ny[P.n_prod] = j * N;
      ++P.n_prod;
    }
}

int conv_wgrad_bf16(const void* x, const void* dy, float* dw, const ConvGeom& c,
                    cudaStream_t stream, bool accumulate, const float* xf_ss, int xf_relu) {
  static thread_local WgradParams P;
  const int taps_total = c.kt * c.kh * c.kw;
  const int K = t_wg_planes;
  const long long NK = (long long)K * c.N;       // batch extent of the tensor maps (stacked planes)
  if (K > 1 && xf_ss != nullptr) return fail(kBadArg, "merged plane products have no operand transform");
  set_products(P, K, c.N);
  P.xf_ss = xf_ss;
  P.xf_relu = xf_relu;
  for (int i = 0; i < kMaxAMaps; ++i) { P.a_dims[i][0] = c.W; P.a_dims[i][1] = c.H; P.a_dims[i][2] = c.T; P.a_dims[i][3] = c.N; }
  // accumulate (fp32 mode): dw already holds the sum of earlier operand-plane products; the kernel only adds
  if (!accumulate) DV_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)c.Cout_p * taps_total * c.Cin_p, stream));
  P.halo = 0;
  {
    static int halo_env = -1;
    if (halo_env < 0) { const char* e = getenv("DV_CONV_HALO"); halo_env = e ? atoi(e) : 1; }
    const int k_chunks = ceil_div(c.Cin_p, 64);
    const int units = taps_total * k_chunks;
    const int n_tiles_h = ceil_div(c.Cout_p, 256);
    const int bn = round_up(ceil_div(c.Cout_p, n_tiles_h), 16);
    const int acc_stride = round_up(bn, 32);
    const int max_pairs = 512 / acc_stride;
    const bool temporal = c.kt > 1 && c.kh == 1 && c.kw == 1;
    const bool spatial = c.kt == 1 && c.kh > 1;
    // halo taps: the filter dimension that shifts by whole 8-row groups (t planes / h rows of 8 positions);
    // a box feeds `span` units, so a group must be able to hold at least one box worth of units
    const int span = temporal ? c.kt : c.kh;
    if (halo_env && (temporal || spatial) && c.st == 1 && c.sh == 1 && c.sw == 1 && units <= 32 &&
        span <= 2 * max_pairs) {
      TileGeom& g = P.g;
      int x_rows;   // rows (positions) of one halo box
      if (temporal) {
        // tile (tn=1, tt, th, tw) of 64 positions with th*tw a multiple of 8 rows; minimise padded volume x halo
        double best = 1e30; int blw = 3, blh = 0, blt = 3;
        for (int a = 0; a <= 6; ++a)
          for (int b = 0; a + b <= 6; ++b) {
            if (a + b < 3) continue;
            const int cc = 6 - a - b;
            const int tw = 1 << a, th = 1 << b, tt = 1 << cc;
            const double vol = (double)round_up(c.Wo, tw) * round_up(c.Ho, th) * round_up(c.To, tt);
            const double cost = vol * (double)(tt + c.kt - 1) / tt * (1.0 + 0.02 * (6 - a));
            if (cost < best) { best = cost; blw = a; blh = b; blt = cc; }
          }
        g.lw = blw; g.lh = blh; g.lt = blt; g.ln = 0;
        x_rows = ((1 << g.lt) + c.kt - 1) << (g.lh + g.lw);
      } else {
        // tile (1, 1, th, tw) with tw >= 8: an h shift of the filter is a shift by tw rows = whole 8-row groups;
        // one box per (kw, chunk). Pick the shape that pads the map least (halo cost included).
        double best = 1e30; int blw = 3;
        for (int a = 3; a <= 6; ++a) {
          const int tw = 1 << a, th = 64 >> a;
          const double vol = (double)round_up(c.Wo, tw) * round_up(c.Ho, th);
          const double cost = vol * (double)(th + c.kh - 1) / th;
          if (cost < best) { best = cost; blw = a; }
        }
        g.lw = blw; g.lh = 6 - blw; g.lt = 0; g.ln = 0;
        x_rows = ((1 << g.lh) + c.kh - 1) << g.lw;
      }
      const double waste = (double)round_up(c.Wo, 1 << g.lw) * round_up(c.Ho, 1 << g.lh) * round_up(c.To, 1 << g.lt) /
                           ((double)c.Wo * c.Ho * c.To);
      const int x_box_bytes = x_rows * 128;
      const int nbx = ceil_div(bn, 64);
      // boxes: temporal -> one per chunk; spatial -> one per (kw, chunk). Units of a box are consecutive.
      const int n_boxes = temporal ? k_chunks : c.kw * k_chunks;
      // groups: whole boxes, at most 2*max_pairs units each
      int boxes_per_group = std::max(1, (2 * max_pairs) / span);
      const int budget = kWgSmemBudget - 1024 - (xf_ss != nullptr ? round_up(k_chunks * 512, 1024) : 0);
      while (boxes_per_group > 1 && 3 * (boxes_per_group * x_box_bytes + nbx * kBoxBytes) > budget)
        --boxes_per_group;
      const int n_groups = ceil_div(n_boxes, boxes_per_group);
      if ((temporal || waste <= 1.16) && n_boxes <= 32 && n_groups <= 8 &&
          3 * (std::min(boxes_per_group, n_boxes) * x_box_bytes + nbx * kBoxBytes) <= budget) {
        P.halo = 1;
        P.x_box_bytes = x_box_bytes;
        P.xb_h = (1 << g.lh) + (spatial ? c.kh - 1 : 0);
        P.xb_t = (1 << g.lt) + (temporal ? c.kt - 1 : 0);
        g.ext_w = c.Wo; g.ext_h = c.Ho; g.ext_t = c.To; g.ext_n = c.N;
        g.tiles_w = ceil_div(c.Wo, 1 << g.lw);
        g.tiles_h = ceil_div(c.Ho, 1 << g.lh);
        g.tiles_t = ceil_div(c.To, 1 << g.lt);
        g.tiles_n = c.N;
        const uint32_t ybox[5] = {64, 1u << g.lw, 1u << g.lh, 1u << g.lt, 1};
        const uint32_t xbox[5] = {64, 1u << g.lw, (uint32_t)((1 << g.lh) + (spatial ? c.kh - 1 : 0)),
                                  (uint32_t)((1 << g.lt) + (temporal ? c.kt - 1 : 0)), 1};
        uint64_t ydims[5] = {(uint64_t)c.Cout_p, (uint64_t)c.Wo, (uint64_t)c.Ho, (uint64_t)c.To, (uint64_t)NK};
        uint64_t ystr[5] = {2, (uint64_t)c.Cout_p * 2, (uint64_t)c.Wo * c.Cout_p * 2,
                            (uint64_t)c.Ho * c.Wo * c.Cout_p * 2, (uint64_t)c.To * c.Ho * c.Wo * c.Cout_p * 2};
        int rc = encode_tmap(&P.dy_map, dy, 2, 5, ydims, ystr, ybox, true);
        if (rc) return rc;
        uint64_t xdims[5] = {(uint64_t)c.Cin_p, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.T, (uint64_t)NK};
        uint64_t xstrd[5] = {2, (uint64_t)c.Cin_p * 2, (uint64_t)c.W * c.Cin_p * 2,
                             (uint64_t)c.H * c.W * c.Cin_p * 2, (uint64_t)c.T * c.H * c.W * c.Cin_p * 2};
        rc = encode_tmap(&P.a_map[0], x, 2, 5, xdims, xstrd, xbox, true);
        if (rc) return rc;
        for (int i = 1; i < kMaxAMaps; ++i) P.a_map[i] = P.a_map[0];
        // boxes ordered (kw, chunk); units ordered (box, halo tap): start addresses increase inside a group, so
        // a pair's LBO is positive
        // one halo step: temporal = a t plane of th*tw rows, spatial = an h row of tw positions (multiples of 8 rows)
        const int step_bytes = temporal ? ((1 << (g.lh + g.lw)) * 128) : ((1 << g.lw) * 128);
        int u = 0, b = 0;
        for (int d = 0; d < (temporal ? 1 : c.kw); ++d)
          for (int kc = 0; kc < k_chunks; ++kc, ++b) {
            P.box_dt[b] = (int8_t)(temporal ? -c.pt : 0);
            P.box_dh[b] = (int8_t)(spatial ? -c.ph : 0);
            P.box_dw[b] = (int8_t)(spatial ? d - c.pw : 0);
            P.box_kc[b] = (int8_t)kc;
            const int slot = b % boxes_per_group;
            for (int a = 0; a < span; ++a, ++u) {
              P.unit_off[u] = (slot * x_box_bytes + a * step_bytes) >> 4;
              P.unit_widx[u] = (int16_t)(temporal ? a : a * c.kw + d);
              P.unit_kc[u] = (int16_t)kc;
            }
          }
        P.n_groups = n_groups;
        for (int gi = 0; gi <= n_groups; ++gi) {
          const int bb = std::min(gi * boxes_per_group, n_boxes);
          P.grp_box0[gi] = (int8_t)bb;
          P.grp_unit0[gi] = (int8_t)(bb * span);
        }
        return wgrad_launch(P, taps_total, c.Cin_p, c.Cout_p, taps_total, dw, stream);
      }
    }
  }

  TileGeom& g = P.g;
  choose_tile_log2(6, c.N, c.To, c.Ho, c.Wo, &g.ln, &g.lt, &g.lh, &g.lw);
  g.ext_w = c.Wo; g.ext_h = c.Ho; g.ext_t = c.To; g.ext_n = c.N;
  g.tiles_w = ceil_div(c.Wo, 1 << g.lw);
  g.tiles_h = ceil_div(c.Ho, 1 << g.lh);
  g.tiles_t = ceil_div(c.To, 1 << g.lt);
  g.tiles_n = ceil_div(c.N, 1 << g.ln);
  const uint32_t box[5] = {64, 1u << g.lw, 1u << g.lh, 1u << g.lt, 1u << g.ln};

  auto encode5 = [&](CUtensorMap* m, const void* base, const long long dim[5], const long long str[5]) {
    uint64_t dims[5], strides[5];
    for (int i = 0; i < 5; ++i) { dims[i] = (uint64_t)dim[i]; strides[i] = (uint64_t)str[i] * 2; }
    return encode_tmap(m, base, 2, 5, dims, strides, box, true);
  };
  if (K > 1 && g.ln > 0 && c.N % (1 << g.ln) != 0)
    return fail(kUnsupported, "merged plane products: position tiles would straddle two planes");
  const long long xdim[5] = {c.Cin_p, c.W, c.H, c.T, NK};
  const long long xstr[5] = {1, c.Cin_p, (long long)c.W * c.Cin_p, (long long)c.H * c.W * c.Cin_p,
                             (long long)c.T * c.H * c.W * c.Cin_p};
  const long long ydim[5] = {c.Cout_p, c.Wo, c.Ho, c.To, NK};
  const long long ystr[5] = {1, c.Cout_p, (long long)c.Wo * c.Cout_p, (long long)c.Ho * c.Wo * c.Cout_p,
                             (long long)c.To * c.Ho * c.Wo * c.Cout_p};
  int rc = encode5(&P.dy_map, dy, ydim, ystr);
  if (rc) return rc;

  int map_of_parity[8];
  for (int i = 0; i < 8; ++i) map_of_parity[i] = -1;
  int nmaps = 0, ntaps = 0;
  for (int a = 0; a < c.kt; ++a)
    for (int b = 0; b < c.kh; ++b)
      for (int d = 0; d < c.kw; ++d) {
        const int ot = a - c.pt, oh = b - c.ph, ow = d - c.pw;
        const int rt = posmod_w(ot, c.st), rh = posmod_w(oh, c.sh), rw = posmod_w(ow, c.sw);
        if (rt > 1 || rh > 1 || rw > 1) return fail(kUnsupported, "conv stride > 2 not supported");
        if (rt >= c.T || rh >= c.H || rw >= c.W) continue;
        const int key = (rt * 2 + rh) * 2 + rw;
        if (map_of_parity[key] < 0) {
          long long dim[5], str[5];
          for (int i = 0; i < 5; ++i) { dim[i] = xdim[i]; str[i] = xstr[i]; }
          const uint8_t* bp = static_cast<const uint8_t*>(x);
          const int rr[3] = {rw, rh, rt};
          const int ss[3] = {c.sw, c.sh, c.st};
          for (int i = 0; i < 3; ++i) {
            bp += (long long)rr[i] * str[1 + i] * 2;
            dim[1 + i] = (dim[1 + i] - rr[i] + ss[i] - 1) / ss[i];
            str[1 + i] *= ss[i];
          }
          rc = encode5(&P.a_map[nmaps], bp, dim, str);
          if (rc) return rc;
          for (int i = 0; i < 3; ++i) P.a_dims[nmaps][i] = (int)dim[1 + i];
          map_of_parity[key] = nmaps++;
        }
        Tap& tp = P.taps[ntaps++];
        tp.map = (int8_t)map_of_parity[key];
        tp.dt = (int8_t)floordiv_w(ot, c.st);
        tp.dh = (int8_t)floordiv_w(oh, c.sh);
        tp.dw = (int8_t)floordiv_w(ow, c.sw);
        tp.widx = (int16_t)((a * c.kh + b) * c.kw + d);
        tp.shift_rows = 0;
      }
  if (ntaps == 0) return kOk;
  for (int i = nmaps; i < kMaxAMaps; ++i) P.a_map[i] = P.a_map[0];
  return wgrad_launch(P, ntaps, c.Cin_p, c.Cout_p, taps_total, dw, stream);
}

// fp32 mode: dw (fp32, packed) = sum over the plane products of wgrad(x_i, dy_j) in ONE launch. x_planes / dy_planes:
// bf16 [K][N][...] contiguous stacks of the split planes.
int conv_wgrad_f32planes(const void* x_planes, const void* dy_planes, int K, float* dw, const ConvGeom& c,
                         cudaStream_t stream) {
  if (K < 1 || K > 3) return fail(kBadArg, "1..3 split planes");
  t_wg_planes = K;
  int rc = conv_wgrad_bf16(x_planes, dy_planes, dw, c, stream, false, nullptr, 0);
  t_wg_planes = 1;
  if (rc != kUnsupported) return rc;
  // position tiles spanning several clips with a ragged last one: one launch per product, accumulating
  const long long xs = (long long)c.N * c.T * c.H * c.W * c.Cin_p * 2, ys = (long long)c.N * c.To * c.Ho * c.Wo * c.Cout_p * 2;
  bool first = true;
  for (int s = 2 * (K - 1); s >= 0; --s)
    for (int i = 0; i < K; ++i) {
      const int j = s - i;
      if (j < 0 || j >= K - i) continue;
      rc = conv_wgrad_bf16(static_cast<const uint8_t*>(x_planes) + i * xs, static_cast<const uint8_t*>(dy_planes) + j * ys,
                           dw, c, stream, !first, nullptr, 0);
      if (rc) return rc;
      first = false;
    }
  return kOk;
}

int conv_stem_wgrad_bf16(const void* x_s2d, const void* dy, float* dw, int N, int T, int H2, int W2,
                         int Cout_p, int kt, int pt, cudaStream_t stream, bool accumulate) {
  static thread_local WgradParams P;
  P.xf_ss = nullptr;
  set_products(P, 1, N);
  const int To = T + 2 * pt - kt + 1;
  const int taps_total = kt * 4;
  if (!accumulate) DV_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout_p * taps_total * 64, stream));
  TileGeom& g = P.g;
  choose_tile_log2(6, N, To, H2, W2, &g.ln, &g.lt, &g.lh, &g.lw);
  g.ext_w = W2; g.ext_h = H2; g.ext_t = To; g.ext_n = N;
  g.tiles_w = ceil_div(W2, 1 << g.lw);
  g.tiles_h = ceil_div(H2, 1 << g.lh);
  g.tiles_t = ceil_div(To, 1 << g.lt);
  g.tiles_n = ceil_div(N, 1 << g.ln);
  const uint32_t box[5] = {64, 1u << g.lw, 1u << g.lh, 1u << g.lt, 1u << g.ln};
  {
    uint64_t dims[5] = {(uint64_t)Cout_p, (uint64_t)W2, (uint64_t)H2, (uint64_t)To, (uint64_t)N};
    uint64_t strides[5] = {2, (uint64_t)Cout_p * 2, (uint64_t)W2 * Cout_p * 2, (uint64_t)H2 * W2 * Cout_p * 2,
                           (uint64_t)To * H2 * W2 * Cout_p * 2};
    int rc = encode_tmap(&P.dy_map, dy, 2, 5, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const long long W2p = W2 + 3;
    uint64_t dims[5] = {64, (uint64_t)W2, (uint64_t)H2, (uint64_t)T, (uint64_t)N};
    uint64_t strides[5] = {2, 32, (uint64_t)W2p * 32, (uint64_t)H2 * W2p * 32, (uint64_t)T * H2 * W2p * 32};
    int rc = encode_tmap(&P.a_map[0], x_s2d, 2, 5, dims, strides, box, true);
    if (rc) return rc;
  }
  for (int i = 1; i < kMaxAMaps; ++i) P.a_map[i] = P.a_map[0];
  int ntaps = 0;
  for (int a = 0; a < kt; ++a)
    for (int r = 0; r < 4; ++r) {
      Tap& tp = P.taps[ntaps];
      tp.map = 0; tp.dt = (int8_t)(a - pt); tp.dh = (int8_t)(r - 2); tp.dw = 0;
      tp.widx = (int16_t)ntaps; tp.shift_rows = 0;
      ++ntaps;
    }
  return wgrad_launch(P, ntaps, 64, Cout_p, taps_total, dw, stream);
}

}  // namespace dv
