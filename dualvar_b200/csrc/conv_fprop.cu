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
// warps 2-9: epilogue, two groups of four warps taking alternate 64-channel chunks of the tile:
//            tcgen05.ld -> (+bias) -> bf16 -> swizzled smem -> TMA store, plus the
//            per-channel sum / sum-of-squares of the stored values for training-mode BatchNorm
//            (reference: nn.BatchNorm3d after every conv, e.g. backbone/r21d.py:56,106,111).
// The grid is persistent (<= one CTA per SM); a CTA keeps one channel tile for its whole life so
// BN partial sums stay in shared memory and are flushed to HBM once per CTA.
#include "bn_xform.cuh"
#include "conv_tile.cuh"
#include "host_common.h"
#include "ptx.cuh"

#include <cuda_bf16.h>

#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

namespace dv {

#ifdef DV_DIAG
constexpr bool kDiag = true;    // diagnostics build (libdualvar_b200_diag.so): per-role cycle counters available
#else
constexpr bool kDiag = false;   // product build: the counters and their branches are compiled out
#endif

constexpr int kNumThreads = 320;            // warp 0 TMA, warp 1 MMA, warps 2-5 and 6-9: two epilogue groups
constexpr int kXfThreads = 128;             // kXf instance: warps 10-13 transform the A boxes (consumer-side BatchNorm)
constexpr int kIssue2Threads = 32;          // last warp: second MMA issuer of "split issue" launches (N <= 128)
constexpr int kEpiThreads = 256;
constexpr int kAStageBytes = kTileM * 128;  // 16 KB
constexpr int kOutBufBytes = kTileM * 128;  // one 64-channel chunk of the output tile
constexpr int kMaxStages = 8;
constexpr int kOutBufs = 4;                  // output staging buffers: two per epilogue group
constexpr int kTmemCols = 512;
constexpr int kSmemBudget = 232448 - 20480;  // 227 KB minus static smem (stat partials, barriers)

// kPair: the kernel runs as CTA pairs (cluster of 2, cta_group::2): one M=256 MMA covers the two CTAs' 128-position
// tiles, each CTA stages its own A box and HALF of the weight tile's rows, so the B operand fetch per CTA halves
// (tests/diag/mma_rate.py: SS-mode operand fetch saturates at 128 B/clk/SM, N=144 alone needs 121 B/clk). Only the
// leader CTA (cluster rank 0) issues MMAs; its full / accumulator-free barriers collect both CTAs' signals, commits
// are multicast to both CTAs' barriers.
// kF32: fp32-mode instance - the epilogue adds its fp32 rows to ConvTileParams::out_f32 (compiled out of the bf16 one).
// Split issue (ConvTileParams::split, N <= 128): one thread cannot issue N = 64 MMAs faster than ~67 cycles each (12 MMAs +
//      one commit per stage, tests/diag/mma_rate_pair.py) although the tensor pipe needs 32 and the operand fetch 40; two
//      issuing threads reach 43. The last warp of the CTA therefore issues every second MMA of a tile into a SECOND
//      accumulator (TMEM columns +128; a thread's own MMAs stay ordered, the two threads' MMAs never touch the same
//      columns), both warps commit to the stage / accumulator barriers (count 2), the epilogue adds the two halves.
// kStack ("kh taps stacked along N", 64-column 3x3 dgrads): the tile's M rows are an UNSHIFTED 16h x 8w input box and one
//      MMA per (kw, K step) multiplies it with B = [W(kh=2); W(kh=1); W(kh=0)] (N = 192), so the box is read from shared
//      memory once instead of once per kh tap and the MMA is tensor-bound instead of operand-fetch-bound. Accumulator
//      chunk s belongs to the output row one h row (8 lanes) further up per s: the epilogue adds D0[r] + D1[r + 8] +
//      D2[r + 16] (warp shuffles, a small shared-memory exchange at the warp boundaries) and stores 14h x 8w outputs.
// kXf: consumer-side BatchNorm instance (bn_xform.cuh): A boxes land on a CTA-local barrier, warps 10-13 apply
//      relu?(scale*y + shift) in place and signal the (leader's) transform barrier the MMA warp waits for; the fused
//      BatchNorm-backward reduce of the epilogue is compiled out of this instance (forward only).
template <bool kPair, bool kF32 = false, bool kXf = false, bool kSplit = false, bool kStack = false>
__global__ void __launch_bounds__((kXf ? kNumThreads + kXfThreads : kNumThreads) + (kSplit ? kIssue2Threads : 0), 1)
conv_tile_kernel(const __grid_constant__ ConvTileParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t bres_bar;
  __shared__ __align__(8) uint64_t afull_bar[kXf ? kMaxStages : 1];   // kXf: this CTA's A box landed
  __shared__ __align__(8) uint64_t xf_bar[kXf ? kMaxStages : 1];      // kXf (leader): A boxes of the stage transformed
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_part[8][2][kMaxBlockN];   // BN partial sums per epilogue warp: [group * 4 + row quarter][sum|sumsq][channel]

  // warp-uniform role index (shfl makes the uniformity visible to the compiler: loop state of the
  // producer / MMA warps then lives in uniform registers, which is what UTMALDG / UTCHMMA consume)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  constexpr int kIssue2Warp = (kXf ? kNumThreads + kXfThreads : kNumThreads) / 32;

  // carve dynamic smem (1024-byte aligned for the 128B swizzle atoms)
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs)
  const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // CTA (pair) index: owns tiles unit, unit+units, ...
  const int units = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int b_tap_bytes = (kPair ? p.block_n / 2 : p.block_n) * 128;   // one weight tile [block_n (/2)][64] of this CTA
  const int b_stage_bytes = p.b_resident ? 0 : p.max_group * b_tap_bytes;
  const int stage_bytes = p.a_stage_bytes + b_stage_bytes;
  float* xf_table = reinterpret_cast<float*>(smem);         // kXf: [k_chunks][scale 64 | shift 64]
  if (kXf) smem += (p.k_chunks * 512 + 1023) & ~1023;
  uint8_t* res_b = smem;                                    // resident weights: taps * k_chunks tiles
  uint8_t* ring = res_b + (p.b_resident ? p.num_taps * p.k_chunks * b_tap_bytes : 0);
  uint8_t* o_smem = ring + p.stages * stage_bytes;          // 2 * 16 KB output staging

  const int n_tile = unit % p.n_tiles;  // the number of CTAs (pairs) is a multiple of n_tiles
  const int bn_mma = (n_tile == p.n_tiles - 1) ? p.last_n : p.block_n;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kSplit ? 2 : 1);         // one commit per issuing thread
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], kSplit ? 2 : 1);
      mbar_init(&tmem_empty_bar[i], kPair ? 2 * kEpiThreads : kEpiThreads);   // the leader's barrier collects both CTAs' epilogues
    }
    mbar_init(&bres_bar, 1);
    if (kXf)
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&afull_bar[i], 1);
        mbar_init(&xf_bar[i], kPair ? 2 * kXfThreads : kXfThreads);   // the leader's barrier collects both CTAs' transforms
      }
    fence_barrier_init();
  }
  for (int c = threadIdx.x; c < 8 * 2 * kMaxBlockN; c += blockDim.x) (&s_part[0][0][0])[c] = 0.f;
  if (kXf) stage_ss_table(xf_table, p.xf_ss, p.xf_cp, p.k_chunks, threadIdx.x, blockDim.x);
  if (warp == 1) {
    if (kPair) {
      tmem_alloc2(&tmem_base_slot, kTmemCols);
      tmem_relinquish2();
    } else {
      tmem_alloc(&tmem_base_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before_sync();
  if (kPair) cluster_sync_all(); else __syncthreads();   // peer barriers must be initialised before remote signals
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  // tile id -> position-tile index of this CTA (a pair covers two consecutive position tiles)
  auto m_index = [&](int tile) { return kPair ? (tile / p.n_tiles) * 2 + (int)rank : tile / p.n_tiles; };

  const TileGeom& g = p.g;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      const bool issuer = elect_one();
      if (issuer) {
        for (int i = 0; i < kMaxAMaps; ++i) tma_prefetch_desc(&p.a_map[i]);
        tma_prefetch_desc(&p.b_map);
      }
      int stage = 0;
      uint32_t phase = 0;
      long long prof_wait_empty = 0;
      const long long prof_t0 = (kDiag && p.prof) ? clock64() : 0;
      // this CTA's rows of the weight tile: all of them, or its half of the MMA's N rows in pair mode
      const int bcol = n_tile * p.block_n + (kPair ? (int)rank * (bn_mma / 2) : 0);
      // signals of both CTAs' loads go to the leader's barriers
      const uint32_t bres_addr = kPair ? mapa_u32(smem_u32(&bres_bar), 0) : smem_u32(&bres_bar);
      const uint32_t full0_addr = kPair ? mapa_u32(smem_u32(&full_bar[0]), 0) : smem_u32(&full_bar[0]);
      const uint32_t tx_mult = kPair ? 2u : 1u;
      if (p.b_resident && issuer) {
        // weight-stationary: every (tap, k-chunk) weight tile of this channel tile is loaded once
        if (rank == 0) mbar_expect_tx(&bres_bar, tx_mult * p.num_taps * p.k_chunks * b_tap_bytes);
        for (int tap = 0; tap < p.num_taps; ++tap)
          for (int kc = 0; kc < p.k_chunks; ++kc)
            tma_load_3d_to<kPair>(res_b + (tap * p.k_chunks + kc) * b_tap_bytes, &p.b_map, bres_addr, kc * kChunkK,
                                  p.taps[tap].widx, bcol);
      }
      for (int tile = unit; tile < p.total_tiles; tile += units) {
        int m_id = m_index(tile);
        const int wb = m_id % g.tiles_w; m_id /= g.tiles_w;
        const int hb = m_id % g.tiles_h; m_id /= g.tiles_h;
        const int tb = m_id % g.tiles_t; m_id /= g.tiles_t;
        const int nb = m_id;
        const int w0 = wb << g.lw, h0 = hb * g.step_h, t0 = tb << g.lt, n0 = nb << g.ln;
        int gb = 0;
        for (int grp = 0; grp < p.num_groups; ++grp) {
          const int len = p.group_len[grp];
          const Tap lead = p.taps[gb];
          // kXf: the A box signals this CTA's own barrier (its transform warps wait there); only weight tiles count
          // towards the leader's full barrier
          const uint32_t tx_bytes = (kXf ? 0 : p.a_tx_bytes) + (p.b_resident ? 0 : len * b_tap_bytes);
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            const long long c0 = (kDiag && p.prof) ? clock64() : 0;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (kDiag && p.prof) prof_wait_empty += clock64() - c0;
            if (issuer) {
              if (rank == 0 && (!kXf || tx_bytes != 0)) mbar_expect_tx(&full_bar[stage], tx_mult * tx_bytes);
              const uint32_t full_addr = full0_addr + (uint32_t)stage * 8u;
              uint8_t* st = ring + stage * stage_bytes;
              if (kXf) {
                mbar_expect_tx(&afull_bar[stage], p.a_tx_bytes);
                tma_load_5d(st, &p.a_map[lead.map], &afull_bar[stage], kc * kChunkK, w0 + lead.dw,
                            h0 + g.org_h + lead.dh, t0 + lead.dt, n0);
              } else
              tma_load_5d_to<kPair>(st, &p.a_map[lead.map], full_addr, kc * kChunkK, w0 + lead.dw,
                                    h0 + g.org_h + lead.dh, t0 + lead.dt, n0);
              if (!p.b_resident)
                for (int i = 0; i < len; ++i)
                  tma_load_3d_to<kPair>(st + p.a_stage_bytes + i * b_tap_bytes, &p.b_map, full_addr, kc * kChunkK,
                                        p.taps[gb + i].widx, bcol);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          gb += len;
        }
      }
      if (kPair) {
        // producer tail: every stage released (all multicast commits delivered) before this CTA may exit
        for (int i = 0; i < p.stages; ++i) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (kDiag && p.prof && issuer) {
        p.prof[blockIdx.x * 16 + 0] = clock64() - prof_t0;   // producer total
        p.prof[blockIdx.x * 16 + 1] = prof_wait_empty;       // producer waiting for a free stage
      }
    }
  } else if (!kSplit && warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only in pair mode)
    if (rank == 0) {
      const bool issuer = elect_one();
      const uint32_t idesc = make_idesc_bf16(kPair ? 2 * kTileM : kTileM, bn_mma, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long prof_wait_full = 0, prof_wait_acc = 0;
      const long long prof_t0 = (kDiag && p.prof) ? clock64() : 0;
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO=1024 B, version 1, SWIZZLE_128B
      const uint32_t desc_lo_flags = 1u << 16;                           // LBO field (unused for K-major swizzled)
      const uint32_t ring_enc = smem_u32(ring) >> 4, res_enc = smem_u32(res_b) >> 4;
      const uint32_t stage_enc = (uint32_t)stage_bytes >> 4, btb_enc = (uint32_t)b_tap_bytes >> 4;
      const uint32_t a_stage_enc = (uint32_t)p.a_stage_bytes >> 4;
      if (p.b_resident) {
        mbar_wait(&bres_bar, 0);
        tc_fence_after_sync();
      }
      for (int tile = unit; tile < p.total_tiles; tile += units, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const long long ca = (kDiag && p.prof) ? clock64() : 0;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        if (kDiag && p.prof) prof_wait_acc += clock64() - ca;
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kMaxBlockN;
        uint32_t accumulate = 0;
        int gb = 0;
        for (int grp = 0; grp < p.num_groups; ++grp) {
          const int len = p.group_len[grp];
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            const long long cf = (kDiag && p.prof) ? clock64() : 0;
            if (!kXf || !p.b_resident) mbar_wait(&full_bar[stage], phase);   // TMA data (kXf: weight tiles only)
            if (kXf) mbar_wait(&xf_bar[stage], phase);                       // both CTAs' A boxes transformed
            if (kDiag && p.prof) prof_wait_full += clock64() - cf;
            tc_fence_after_sync();
            // descriptor low words (start address >> 4 | LBO field); the high word is a constant
            const uint32_t st_lo = desc_lo_flags | (ring_enc + (uint32_t)stage * stage_enc);
            const uint32_t b_lo = p.b_resident ? (desc_lo_flags | (res_enc + (uint32_t)kc * btb_enc)) : st_lo;
            const int ksteps = (kc == p.k_chunks - 1) ? p.k_steps_last : 4;
            for (int i = 0; i < len; ++i) {
              // the tap's 128 rows start shift_rows rows into the shared (halo) A box; shifts are multiples
              // of 8 rows = 1024 B, so the 128B-swizzle phase of every row is unchanged.
              // (offsets come from the kernel-parameter constant bank: uniform loads, no smem round trip)
              const uint32_t aoff = (uint32_t)p.taps[gb + i].shift_rows * 8u;
              const uint32_t boff = p.b_resident ? (uint32_t)((gb + i) * p.k_chunks) * btb_enc
                                                 : a_stage_enc + (uint32_t)i * btb_enc;
              const uint32_t al = st_lo + aoff, bl = b_lo + boff;
              if (issuer) {
                // +32 bytes along K inside the 128B swizzle span = +2 in (addr >> 4) units
                umma_issue<kPair>(d_tmem, al, bl, desc_hi, idesc, accumulate);
                if (ksteps > 1) umma_issue<kPair>(d_tmem, al + 2, bl + 2, desc_hi, idesc, 1);
                if (ksteps > 2) umma_issue<kPair>(d_tmem, al + 4, bl + 4, desc_hi, idesc, 1);
                if (ksteps > 3) umma_issue<kPair>(d_tmem, al + 6, bl + 6, desc_hi, idesc, 1);
              }
              accumulate = 1;
            }
            if (issuer) umma_commit_to<kPair>(&empty_bar[stage]);
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          gb += len;
        }
        if (issuer) umma_commit_to<kPair>(&tmem_full_bar[acc]);
        __syncwarp();
      }
      if (kDiag && p.prof && issuer) {
        p.prof[blockIdx.x * 16 + 2] = clock64() - prof_t0;   // MMA issuer total
        p.prof[blockIdx.x * 16 + 3] = prof_wait_full;        // waiting for TMA data
        p.prof[blockIdx.x * 16 + 4] = prof_wait_acc;         // waiting for the epilogue to free an accumulator
      }
    }
  } else if (kSplit && (warp == 1 || warp == kIssue2Warp)) {
    // ------------------------------------------------------------------ split issue: two MMA issuers (leader CTA only)
    // One thread cannot issue 64-column MMAs as fast as the tensor pipe retires them (tests/diag/mma_rate_pair.py), so
    // warp 1 and the last warp each issue every second MMA of a tile, into separate accumulators, and each commits
    // to the stage / accumulator barriers (count 2). The loop holds nothing but the MMAs: per tap ONE constant-bank word
    // (ConvTileParams::prog, fetched a tap ahead) with both operand offsets.
    const uint32_t which = warp == 1 ? 0u : 1u;
    if (rank == 0) {
      const bool issuer = elect_one();
      const uint32_t idesc = make_idesc_bf16(kPair ? 2 * kTileM : kTileM, bn_mma, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO=1024 B, version 1, SWIZZLE_128B
      const uint32_t desc_lo_flags = 1u << 16;                           // LBO field (unused for K-major swizzled)
      const uint32_t ring_lo = desc_lo_flags | (smem_u32(ring) >> 4);
      const uint32_t res_lo = desc_lo_flags | (smem_u32(res_b) >> 4);
      const uint32_t stage_enc = (uint32_t)stage_bytes >> 4, btb_enc = (uint32_t)b_tap_bytes >> 4;
      const int k_chunks = p.k_chunks, k_last = p.k_steps_last, num_groups = p.num_groups, stages = p.stages;
      const bool resident = p.b_resident != 0;
      if (resident) {
        mbar_wait(&bres_bar, 0);
        tc_fence_after_sync();
      }
      for (int tile = unit; tile < p.total_tiles; tile += units, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kMaxBlockN + which * (kMaxBlockN / 2);
        uint32_t accumulate = 0;
        int gb = 0;
        for (int grp = 0; grp < num_groups; ++grp) {
          const int len = p.group_len[grp];
          for (int kc = 0; kc < k_chunks; ++kc) {
            uint32_t nxt = p.prog[gb];     // first tap's offsets: in flight during the wait
            if (!kXf || !resident) mbar_wait(&full_bar[stage], phase);      // TMA data (kXf: weight tiles only)
            if (kXf) mbar_wait(&xf_bar[stage], phase);                       // both CTAs' A boxes transformed
            tc_fence_after_sync();
            // descriptor low words (start address >> 4 | LBO field); the high word is a constant. A tap's word of the
            // program holds its A offset inside the (halo) box - shifts are multiples of 8 rows = 1024 B, so the
            // 128B-swizzle phase of every row is unchanged - and its weight tile's offset (resident: from the resident
            // base, chunk kc adds btb; streamed: from the stage base)
            const uint32_t st_lo = ring_lo + (uint32_t)stage * stage_enc;
            const uint32_t b_lo = resident ? res_lo + (uint32_t)kc * btb_enc : st_lo;
            // KS K steps of every tap of the stage (+32 bytes along K inside the swizzle span = +2 in addr >> 4).
            // MMA number i * KS + ks of the stage belongs to thread (i * KS + ks) & 1, i.e. this thread takes the K
            // steps q, q + 2 of a tap with q = parity of (i * KS) ^ which.
            auto run = [&](auto ks_c) {
              constexpr int KS = decltype(ks_c)::value;
              uint32_t q = which;
              for (int i = 0; i < len; ++i) {
                const uint32_t cur = nxt;
                nxt = p.prog[gb + i + 1];
                const uint32_t al = st_lo + (cur & 0xffffu), bl = b_lo + (cur >> 16);
                if (q < (uint32_t)KS) {
                  if (issuer) umma_issue<kPair>(d_tmem, al + 2 * q, bl + 2 * q, desc_hi, idesc, accumulate);
                  accumulate = 1;
                }
                if (KS > 2 && q + 2 < (uint32_t)KS) {
                  if (issuer) umma_issue<kPair>(d_tmem, al + 2 * q + 4, bl + 2 * q + 4, desc_hi, idesc, 1);
                }
                if (KS & 1) q ^= 1u;
              }
            };
            const int ksteps = (kc == k_chunks - 1) ? k_last : 4;
            using std::integral_constant;
            if (ksteps == 4) run(integral_constant<int, 4>{});
            else if (ksteps == 1) run(integral_constant<int, 1>{});
            else if (ksteps == 2) run(integral_constant<int, 2>{});
            else run(integral_constant<int, 3>{});
            if (issuer) umma_commit_to<kPair>(&empty_bar[stage]);
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
          gb += len;
        }
        if (issuer) umma_commit_to<kPair>(&tmem_full_bar[acc]);
        __syncwarp();
      }
    }
  } else if (kXf && warp >= 10 && warp < 14) {
    // ------------------------------------------------------------------ A-operand transform (kXf: warps 10-13)
    const int tid = threadIdx.x - kNumThreads;
    const uint32_t xf0_addr = kPair ? mapa_u32(smem_u32(&xf_bar[0]), 0) : smem_u32(&xf_bar[0]);
    int stage = 0;
    uint32_t phase = 0;
    XfBox b;
    b.rows = p.a_tx_bytes >> 7; b.lw = 31 - __clz(p.a_box[0]); b.bh = p.a_box[1]; b.bt = p.a_box[2];
    for (int tile = unit; tile < p.total_tiles; tile += units) {
      int m_id = m_index(tile);
      const int wb = m_id % g.tiles_w; m_id /= g.tiles_w;
      const int hb = m_id % g.tiles_h; m_id /= g.tiles_h;
      const int tb = m_id % g.tiles_t; m_id /= g.tiles_t;
      const int nb = m_id;
      const int w0 = wb << g.lw, h0 = hb * g.step_h, t0 = tb << g.lt, n0 = nb << g.ln;
      int gb = 0;
      for (int grp = 0; grp < p.num_groups; ++grp) {
        const Tap lead = p.taps[gb];
        b.ow = w0 + lead.dw; b.oh = h0 + g.org_h + lead.dh; b.ot = t0 + lead.dt; b.on = n0;
        b.dw = p.a_dims[lead.map][0]; b.dh = p.a_dims[lead.map][1]; b.dt = p.a_dims[lead.map][2]; b.dn = p.a_dims[lead.map][3];
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&afull_bar[stage], phase);
          bnrelu_box_inplace(ring + stage * stage_bytes, b, xf_table + kc * 128, p.xf_relu, tid, kXfThreads);
          fence_proxy_async_smem();      // generic-proxy writes -> visible to tcgen05.mma's async-proxy reads
          if (kPair) mbar_arrive_cluster(xf0_addr + (uint32_t)stage * 8u); else mbar_arrive(&xf_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        gb += p.group_len[grp];
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ------------------------------------------------------------------ epilogue (2 groups x 128 threads)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;     // tile row == TMEM lane
    const int grp = (warp - 2) >> 2;   // epilogue group: chunks grp, grp + 2, ... of every tile
    const int et = (threadIdx.x - 64) & 127;   // thread index inside the group
    const bool leader = (et == 0);
    uint8_t* const o_grp = o_smem + grp * 2 * kOutBufBytes;
    const int rw = row & ((1 << g.lw) - 1);
    const int rh = (row >> g.lw) & ((1 << g.lh) - 1);
    const int rt = (row >> (g.lw + g.lh)) & ((1 << g.lt) - 1);
    const int rn = row >> (g.lw + g.lh + g.lt);
    const int nchunks = (bn_mma + 63) >> 6;
    const int bcol = n_tile * p.block_n;
    const bool do_stats = p.stats != nullptr;
    const bool red = !kXf && do_stats && p.red_y != nullptr;
    int it = 0;
    uint32_t obuf = 0;
    long long prof_epi_wait = 0, prof_ld = 0, prof_sts = 0, prof_bar = 0, prof_store = 0, prof_stat = 0, prof_yld = 0;
    const bool prof_on = kDiag && p.prof && threadIdx.x == 64;
    const long long prof_t0 = prof_on ? clock64() : 0;
    const uint32_t acc_free0 = kPair ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : smem_u32(&tmem_empty_bar[0]);
    for (int tile = unit; tile < p.total_tiles; tile += units, ++it) {
      int m_id = m_index(tile);
      const int wb = m_id % g.tiles_w; m_id /= g.tiles_w;
      const int hb = m_id % g.tiles_h; m_id /= g.tiles_h;
      const int tb = m_id % g.tiles_t; m_id /= g.tiles_t;
      const int nb = m_id;
      const int w0 = wb << g.lw, h0 = hb * g.step_h, t0 = tb << g.lt, n0 = nb << g.ln;
      const bool valid = (w0 + rw < g.ext_w) && (h0 + rh < g.ext_h) && (t0 + rt < g.ext_t) &&
                         (n0 + rn < g.ext_n);
      if (kStack) {
        // ---- kStack epilogue: out[r] = D0[r] + D1[r + 8] + D2[r + 16] for the 112 rows with rh < 14. Both epilogue groups
        // work on every tile: group g takes the columns [32g, 32g + 32) (two pieces of 16), one shared staging buffer.
        const int sacc = it & 1;
        mbar_wait(&tmem_full_bar[sacc], (it >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t ta = tmem_base + sacc * kMaxBlockN + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* ob = o_smem;                                                                 // staging buffer (both groups)
        float* xb = reinterpret_cast<float*>(o_smem + kOutBufBytes) + grp * (4 * 24 * 16);    // [4 warps][24 rows][16]
        if (threadIdx.x == 64) tma_store_wait_read<0>();                                      // the previous tile's store has read `ob`
        named_bar_sync(3, kEpiThreads);
        const bool keep = rh < 14;
        const bool up1 = lane >= 24 && q < 3, up2 = lane >= 16 && q < 3;
#pragma unroll 1
        for (int pi = 0; pi < 2; ++pi) {
          const int pc = grp * 2 + pi;
          uint32_t d0[16], d1[16], d2[16];
          tmem_ld16(ta + pc * 16, d0);
          tmem_ld16(ta + 64 + pc * 16, d1);
          tmem_ld16(ta + 128 + pc * 16, d2);
          tmem_ld_wait();
          if (pi == 1) {
            tc_fence_before_sync();
            if (kPair) mbar_arrive_cluster(acc_free0 + (uint32_t)sacc * 8u); else mbar_arrive(&tmem_empty_bar[sacc]);
          }
          // rows the warp above needs from this one: D1 of lanes 0-7, D2 of lanes 0-15
          if (lane < 8) {
            float4* dst = reinterpret_cast<float4*>(xb + (q * 24 + lane) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_float4(__uint_as_float(d1[4 * j]), __uint_as_float(d1[4 * j + 1]), __uint_as_float(d1[4 * j + 2]),
                                   __uint_as_float(d1[4 * j + 3]));
          }
          if (lane < 16) {
            float4* dst = reinterpret_cast<float4*>(xb + (q * 24 + 8 + lane) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_float4(__uint_as_float(d2[4 * j]), __uint_as_float(d2[4 * j + 1]), __uint_as_float(d2[4 * j + 2]),
                                   __uint_as_float(d2[4 * j + 3]));
          }
          named_bar_sync(1 + grp, 128);
          float x1v[16], x2v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            x1v[j] = __shfl_down_sync(0xffffffffu, __uint_as_float(d1[j]), 8);
            x2v[j] = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[j]), 16);
          }
          if (up1) {
            const float4* src = reinterpret_cast<const float4*>(xb + ((q + 1) * 24 + (lane - 24)) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 v4 = src[j];
              x1v[4 * j] = v4.x; x1v[4 * j + 1] = v4.y; x1v[4 * j + 2] = v4.z; x1v[4 * j + 3] = v4.w;
            }
          }
          if (up2) {
            const float4* src = reinterpret_cast<const float4*>(xb + ((q + 1) * 24 + 8 + (lane - 16)) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 v4 = src[j];
              x2v[4 * j] = v4.x; x2v[4 * j + 1] = v4.y; x2v[4 * j + 2] = v4.z; x2v[4 * j + 3] = v4.w;
            }
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float s0 = __uint_as_float(d0[j]) + x1v[j] + x2v[j];
            const float s1 = __uint_as_float(d0[j + 1]) + x1v[j + 1] + x2v[j + 1];
            __nv_bfloat162 h2 = __floats2bfloat162_rn(keep ? s0 : 0.f, keep ? s1 : 0.f);
            pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h2);
          }
          uint8_t* orow = ob + row * 128;
          *reinterpret_cast<uint4*>(orow + (((2 * pc) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(orow + (((2 * pc + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          if (pi == 0) named_bar_sync(1 + grp, 128);      // the exchange buffer is re-used by the second piece
        }
        fence_proxy_async_smem();
        named_bar_sync(3, kEpiThreads);
        if (threadIdx.x == 64) {
          tma_store_5d(&p.out_map, ob, 0, w0, h0, t0, n0);
          tma_store_commit();
        }
        if (red) {
          // column sums of g = dx * relu_mask(y) and g * y over the staged tile (rows with rh >= 14 hold zeros): thread ->
          // (channel pair, eighth of the rows)
          const int word = et & 31;
          const int r8 = grp * 4 + (et >> 5);
          const int c = word * 2;
          if (c < p.stats_ld) {
            const __nv_bfloat16* ybase = reinterpret_cast<const __nv_bfloat16*>(p.red_y) + c + w0 * p.red_stride[0] +
                                         h0 * p.red_stride[1] + t0 * p.red_stride[2] + n0 * p.red_stride[3];
            uint32_t yv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int r = r8 * 16 + i;
              yv[i] = ldg_u32_pred(ybase + (r & 7) * p.red_stride[0] + (r >> 3) * p.red_stride[1], (r >> 3) < 14);
            }
            float sc0 = 0.f, sc1 = 0.f, sh0 = 1.f, sh1 = 1.f;   // no ReLU: mask always true
            if (p.red_ss != nullptr) {
              sc0 = p.red_ss[c]; sc1 = p.red_ss[c + 1];
              sh0 = p.red_ss[p.stats_ld + c]; sh1 = p.red_ss[p.stats_ld + c + 1];
            }
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int r = r8 * 16 + i;
              const uint32_t off = r * 128 + ((((word >> 2) ^ (r & 7))) << 4) + ((word & 3) << 2);
              const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ob + off));
              const float2 yy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv[i]));
              const float g0 = fmaf(yy.x, sc0, sh0) > 0.f ? d.x : 0.f;
              const float g1 = fmaf(yy.y, sc1, sh1) > 0.f ? d.y : 0.f;
              a0 += g0; a1 += g1;
              b0 = fmaf(g0, yy.x, b0); b1 = fmaf(g1, yy.y, b1);
            }
            float2* ps = reinterpret_cast<float2*>(&s_part[r8][0][c]);
            float2* pq = reinterpret_cast<float2*>(&s_part[r8][1][c]);
            float2 s2 = *ps, q2 = *pq;
            s2.x += a0; s2.y += a1; q2.x += b0; q2.y += b1;
            *ps = s2; *pq = q2;
          }
        }
        continue;
      }
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long ce = prof_on ? clock64() : 0;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      if (prof_on) prof_epi_wait += clock64() - ce;
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + acc * kMaxBlockN + (static_cast<uint32_t>(q * 32) << 16);
      // the two groups take alternate chunks, swapping roles every tile so that odd chunk counts balance out
      const int first_cc = (grp + it) & 1;
      // last chunk of this tile that this group loads from TMEM (-1: none)
      const int last_cc = (nchunks - 1 - first_cc >= 0) ? first_cc + ((nchunks - 1 - first_cc) & ~1) : -1;
      if (last_cc < 0) {
        tc_fence_before_sync();
        if (kPair) mbar_arrive_cluster(acc_free0 + (uint32_t)acc * 8u); else mbar_arrive(&tmem_empty_bar[acc]);
      }
      for (int cc = first_cc; cc < nchunks; cc += 2) {
        const int ncols = min(64, bn_mma - cc * 64);
        uint8_t* ob = o_grp + obuf * kOutBufBytes;
        obuf ^= 1;
        const long long pc0 = prof_on ? clock64() : 0;
        // fused BN-backward reduce: this thread's 32 (row, channel pair) words of y for the column pass
        // below are requested now, so their latency hides behind the TMEM load / convert / store phase.
        // Row bit k of the tile adds red_bitoff[k] elements (tile dims are powers of two), so every row
        // offset is a sum of kernel-parameter constants: no per-row index arithmetic.
        uint32_t yv[32];
        if (red) {
          const int c = bcol + cc * 64 + (et & 31) * 2;
          const bool col_ok = ((et & 31) * 2 < ncols) && (c < p.stats_ld);
          const int rq = et >> 5;
          const __nv_bfloat16* ybase =
              reinterpret_cast<const __nv_bfloat16*>(p.red_y) + c + w0 * p.red_stride[0] + h0 * p.red_stride[1] +
              t0 * p.red_stride[2] + n0 * p.red_stride[3] + ((rq & 1) ? p.red_bitoff[5] : 0) +
              ((rq & 2) ? p.red_bitoff[6] : 0);
          // rows of this quarter that fall outside the tensor (partial tiles): per-dimension limits on the
          // row's bit fields
          const int lim_w = g.ext_w - w0, lim_h = g.ext_h - h0, lim_t = g.ext_t - t0, lim_n = g.ext_n - n0;
          const bool full = lim_w >= (1 << g.lw) && lim_h >= (1 << g.lh) && lim_t >= (1 << g.lt) &&
                            lim_n >= (1 << g.ln);
          auto row_off = [&](int i) {
            return ((i & 1) ? p.red_bitoff[0] : 0) + ((i & 2) ? p.red_bitoff[1] : 0) +
                   ((i & 4) ? p.red_bitoff[2] : 0) + ((i & 8) ? p.red_bitoff[3] : 0) +
                   ((i & 16) ? p.red_bitoff[4] : 0);
          };
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) yv[i] = ldg_u32_pred(ybase + row_off(i), col_ok);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int r = rq * 32 + i;
              const int ok = (int)col_ok & (int)((r & ((1 << g.lw) - 1)) < lim_w) &
                             (int)(((r >> g.lw) & ((1 << g.lh) - 1)) < lim_h) &
                             (int)(((r >> (g.lw + g.lh)) & ((1 << g.lt) - 1)) < lim_t) &
                             (int)((r >> (g.lw + g.lh + g.lt)) < lim_n);
              yv[i] = ldg_u32_pred(ybase + row_off(i), ok != 0);
            }
          }
        }
        const long long pc1 = prof_on ? clock64() : 0;
        // all TMEM loads of the chunk in flight together, one wait
        uint32_t v[64];
        tmem_ld16(t_addr + cc * 64, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        if (ncols > 16) tmem_ld16(t_addr + cc * 64 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        if (ncols > 32) tmem_ld16(t_addr + cc * 64 + 32, *reinterpret_cast<uint32_t(*)[16]>(&v[32]));
        if (ncols > 48) tmem_ld16(t_addr + cc * 64 + 48, *reinterpret_cast<uint32_t(*)[16]>(&v[48]));
        tmem_ld_wait();
        if (kSplit) {
          // split issue: the second issuing thread's half of the sum sits 128 columns further
#pragma unroll
          for (int part = 0; part < 4; ++part) {
            if (part * 16 < ncols) {
              uint32_t u[16];
              tmem_ld16(t_addr + kMaxBlockN / 2 + cc * 64 + part * 16, u);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                v[part * 16 + j] = __float_as_uint(__uint_as_float(v[part * 16 + j]) + __uint_as_float(u[j]));
            }
          }
        }
        const long long pc2 = prof_on ? clock64() : 0;
        if (p.bias != nullptr) {   // conv bias (C3D): uniform branch, off the common path
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const int c = bcol + cc * 64 + j;
            if (j < ncols && c < p.stats_ld) v[j] = __float_as_uint(__uint_as_float(v[j]) + p.bias[c]);
          }
        }
        if (!valid) {              // rows outside the tensor (partial tiles) store zeros
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = 0u;
        }
        if (cc == last_cc) {
          // the accumulator now lives in registers: hand the TMEM buffer back to the MMA warp
          tc_fence_before_sync();
          if (kPair) mbar_arrive_cluster(acc_free0 + (uint32_t)acc * 8u); else mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (kF32) {
          // fp32 mode: the accumulator row goes straight from registers into the fp32 output (this thread's
          // position, 64 consecutive channels = whole 32-byte sectors), added to what earlier launches left there
          if (valid) {
            float* dst = p.out_f32 + (long long)(w0 + rw) * p.red_stride[0] + (long long)(h0 + rh) * p.red_stride[1] +
                         (long long)(t0 + rt) * p.red_stride[2] + (long long)(n0 + rn) * p.red_stride[3] + bcol + cc * 64;
            if (p.f32_store) {     // first product of the sum: overwrite (no zero fill, no read-modify-write)
#pragma unroll
              for (int j = 0; j < 64; j += 4)
                if (j < ncols && bcol + cc * 64 + j < p.stats_ld)
                  st_global_v4_f32(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 64; j += 4)
                if (j < ncols && bcol + cc * 64 + j < p.stats_ld)
                  red_add_v4_f32(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                 __uint_as_float(v[j + 3]));
            }
          }
          if (do_stats) {
            // merged plane products: the accumulator is the finished sum, so the batch statistics are taken here.
            // Transposing butterfly over the warp's 32 rows: every round halves the values a lane carries and
            // doubles the rows they cover; after 5 rounds lane l holds the sums of columns 2l and 2l + 1.
            const int lane = et & 31;
            float cs[32], cq[32];
            {
              const bool up = (lane & 16) != 0;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float lo = (i < ncols) ? __uint_as_float(v[i]) : 0.f;
                const float hi = (i + 32 < ncols) ? __uint_as_float(v[i + 32]) : 0.f;
                const float keep = up ? hi : lo;
                const float got = __shfl_xor_sync(0xffffffffu, up ? lo : hi, 16);
                cs[i] = keep + got;
                cq[i] = fmaf(keep, keep, got * got);
              }
            }
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) {
              const bool up = (lane & m) != 0;
#pragma unroll
              for (int i = 0; i < 2 * m; ++i) {
                const float ks = up ? cs[i + 2 * m] : cs[i], ss = up ? cs[i] : cs[i + 2 * m];
                const float kq = up ? cq[i + 2 * m] : cq[i], sq = up ? cq[i] : cq[i + 2 * m];
                cs[i] = ks + __shfl_xor_sync(0xffffffffu, ss, m);
                cq[i] = kq + __shfl_xor_sync(0xffffffffu, sq, m);
              }
            }
            if (lane * 2 < ncols) {      // each (epilogue warp, channel) partial has one owner: no atomics
              const int c = cc * 64 + lane * 2;
              float2* ps = reinterpret_cast<float2*>(&s_part[grp * 4 + (et >> 5)][0][c]);
              float2* pq = reinterpret_cast<float2*>(&s_part[grp * 4 + (et >> 5)][1][c]);
              float2 s2 = *ps, q2 = *pq;
              s2.x += cs[0]; s2.y += cs[1]; q2.x += cq[0]; q2.y += cq[1];
              *ps = s2; *pq = q2;
            }
          }
          continue;
        }
        uint8_t* orow = ob + row * 128;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          if (gi * 16 < ncols) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float f0 = __uint_as_float(v[gi * 16 + 2 * j]);
              const float f1 = __uint_as_float(v[gi * 16 + 2 * j + 1]);
              __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
              pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            const int c16a = (2 * gi) ^ (row & 7);
            const int c16b = (2 * gi + 1) ^ (row & 7);
            *reinterpret_cast<uint4*>(orow + c16a * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(orow + c16b * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        fence_proxy_async_smem();
        const long long pc3 = prof_on ? clock64() : 0;
        // the group's previous store (other buffer) has been read before the barrier releases anybody to
        // overwrite that buffer with the next chunk (2 buffers per group)
        if (leader) tma_store_wait_read<0>();
        named_bar_sync(1 + grp, 128);
        const long long pc4 = prof_on ? clock64() : 0;
        if (leader) {
          tma_store_5d(&p.out_map, ob, bcol + cc * 64, w0, h0, t0, n0);
          tma_store_commit();
        }
        const long long pc5 = prof_on ? clock64() : 0;
        if (red) {
          // column sums of g = dz * relu_mask(y) and g * y over the stored bf16 dz tile (same thread ->
          // (channel pair, row quarter) ownership as the forward statistics below)
          const int word = et & 31;
          const int rq = et >> 5;
          const int c = bcol + cc * 64 + word * 2;
          if (word * 2 < ncols && c < p.stats_ld) {
            float sc0 = 0.f, sc1 = 0.f, sh0 = 1.f, sh1 = 1.f;   // no ReLU: mask always true
            if (p.red_ss != nullptr) {
              sc0 = p.red_ss[c]; sc1 = p.red_ss[c + 1];
              sh0 = p.red_ss[p.stats_ld + c]; sh1 = p.red_ss[p.stats_ld + c + 1];
            }
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int r = rq * 32 + i;
              const uint32_t off = r * 128 + ((((word >> 2) ^ (r & 7))) << 4) + ((word & 3) << 2);
              const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ob + off));
              const float2 yy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv[i]));
              const float g0 = fmaf(yy.x, sc0, sh0) > 0.f ? d.x : 0.f;
              const float g1 = fmaf(yy.y, sc1, sh1) > 0.f ? d.y : 0.f;
              a0 += g0; a1 += g1;
              b0 = fmaf(g0, yy.x, b0); b1 = fmaf(g1, yy.y, b1);
            }
            const int lc = cc * 64 + word * 2;
            float2* ps = reinterpret_cast<float2*>(&s_part[grp * 4 + rq][0][lc]);
            float2* pq = reinterpret_cast<float2*>(&s_part[grp * 4 + rq][1][lc]);
            float2 s2 = *ps, q2 = *pq;
            s2.x += a0; s2.y += a1; q2.x += b0; q2.y += b1;
            *ps = s2; *pq = q2;
          }
        } else if (do_stats) {
          // column sums over the stored bf16 tile: thread -> (channel pair, row quarter); every
          // (row quarter, channel) partial is owned by exactly one thread -> no atomics
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
            float2* ps = reinterpret_cast<float2*>(&s_part[grp * 4 + rq][0][c]);
            float2* pq = reinterpret_cast<float2*>(&s_part[grp * 4 + rq][1][c]);
            float2 s2 = *ps, q2 = *pq;
            s2.x += a0; s2.y += a1; q2.x += b0; q2.y += b1;
            *ps = s2; *pq = q2;
          }
        }
        if (prof_on) {
          const long long pc6 = clock64();
          prof_yld += pc1 - pc0; prof_ld += pc2 - pc1; prof_sts += pc3 - pc2; prof_bar += pc4 - pc3;
          prof_store += pc5 - pc4; prof_stat += pc6 - pc5;
        }
      }
    }
    if (prof_on) {
      p.prof[blockIdx.x * 16 + 8] = prof_yld;     // issuing the y loads of the fused BN-backward reduce
      p.prof[blockIdx.x * 16 + 9] = prof_ld;      // tcgen05.ld + wait
      p.prof[blockIdx.x * 16 + 10] = prof_sts;    // convert + st.shared + proxy fence
      p.prof[blockIdx.x * 16 + 11] = prof_bar;    // named barrier (waiting for the other epilogue warps)
      p.prof[blockIdx.x * 16 + 12] = prof_store;  // TMA store issue + wait for the store before last to drain
      p.prof[blockIdx.x * 16 + 13] = prof_stat;   // column-sum pass
      p.prof[blockIdx.x * 16 + 5] = clock64() - prof_t0;     // epilogue total
      p.prof[blockIdx.x * 16 + 6] = prof_epi_wait;           // epilogue waiting for an accumulator
      p.prof[blockIdx.x * 16 + 7] = it;                      // tiles processed by this CTA
    }
    if (leader) tma_store_wait_all<0>();
    if (do_stats) {
      named_bar_sync(3, kEpiThreads);
      for (int c = threadIdx.x - 64; c < bn_mma; c += kEpiThreads) {
        const int gc = bcol + c;
        if (gc < p.stats_ld) {
          double su = 0.0, sq = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) {
            su += (double)s_part[w8][0][c];
            sq += (double)s_part[w8][1][c];
          }
          atomicAdd(&p.stats[gc], su);
          atomicAdd(&p.stats[p.stats_ld + gc], sq);
        }
      }
    }
  }

  tc_fence_before_sync();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    if (kPair) tmem_dealloc2(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ======================================================================== host side

struct View5 {
  // NDHWC view, innermost-first: dims {C, W, H, T, N}; strides in elements
  const void* base;
  long long dim[5];
  long long stride[5];
  int esize;   // element bytes: 2 (bf16) or 4 (fp32 output of the fp32 mode)
  int store;   // fp32 output: 1 = overwrite, 0 = add
};

static View5 make_ndhwc(const void* ptr, int N, int T, int H, int W, int Cp) {
  View5 v;
  v.base = ptr;
  v.esize = 2;
  v.store = 0;
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
  v.base = static_cast<const uint8_t*>(v.base) + (long long)r * v.stride[d] * v.esize;
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

// One filter tap as the host sees it: which input view (stride-parity class) it reads and the offset of
// its box origin relative to the output tile origin, in that view's coordinates.
struct TapSpec {
  int view;
  int ot, oh, ow;
  int widx;
};

using MapEncoder = int (*)(CUtensorMap*, const void* ctx, int view, const uint32_t box[5]);

// fp32 mode, all plane products of a convolution in ONE launch (conv_fprop_f32planes / conv_dgrad_f32planes): the input
// "views" are the bf16 split planes of one tensor (same geometry), not stride-parity classes - tap grouping stays legal.
static thread_local bool t_plane_views = false;
static thread_local int t_plane_taps = 0;      // filter taps per weight plane (tap index / t_plane_taps = weight plane)
static thread_local bool t_f32_split = false;  // set by the merged plane-product entry points: two-issuer instances allowed
static bool f32_split_enabled() {               // DV_F32_SPLIT_ISSUE=0: one issuing thread in the fp32 instances (A/B)
  static int v = -1;
  if (v < 0) { const char* e = getenv("DV_F32_SPLIT_ISSUE"); v = e ? atoi(e) : 1; }
  return v != 0;
}

static long long* g_prof = nullptr;
#ifdef DV_DIAG
void set_conv_profile(long long* p) { g_prof = p; }
#endif
static int g_halo_enabled = 1;      // DV_CONV_HALO=0 disables tap grouping (A/B testing)
static int g_resident_enabled = 1;  // DV_CONV_RESIDENT=0 disables weight-stationary CTAs
static int g_pair_enabled = 1;      // DV_CONV_PAIR=0 disables CTA pairs (cta_group::2)
static int g_split_enabled = 1;     // DV_CONV_SPLIT=0 disables the two-region tiling of maps with H = 8 (mod 16)
static int g_split_issue = 1;       // DV_CONV_SPLIT_ISSUE=0: one MMA-issuing thread for every launch
static void read_env_once() {
  static bool done = false;
  if (done) return;
  done = true;
  if (const char* e = getenv("DV_CONV_HALO")) g_halo_enabled = atoi(e);
  if (const char* e = getenv("DV_CONV_RESIDENT")) g_resident_enabled = atoi(e);
  if (const char* e = getenv("DV_CONV_PAIR")) g_pair_enabled = atoi(e);
  if (const char* e = getenv("DV_CONV_SPLIT")) g_split_enabled = atoi(e);
  if (const char* e = getenv("DV_CONV_SPLIT_ISSUE")) g_split_issue = atoi(e);
}

// out(view) = sum_taps A_view(tap)[box shifted by tap] * W[tap]
//
// Tap grouping (halo re-use): taps that differ only by whole rows-of-8 of the tile share ONE TMA box
// with a halo; each tap's MMA reads its 128 rows through a descriptor shifted by a multiple of 1024 B.
//   temporal-only filters  : one group, box (tw, th, tt + kt - 1), tile th*tw multiple of 8
//   filters with kh > 1    : tile (tt=1, th=16, tw=8), one group per kw, box (8, 16 + kh - 1, kt)
// Everything else (strided layers = several views, small maps) keeps one box per tap.
// rows_lw: log2 of the tile width in "rows" mode (tile = 2^(7-rows_lw) h x 2^rows_lw w); org_h: h offset of the
// output region inside the full tensor (added to the input box coordinates; outv is already the region's view).
static int conv_multi_tap_region(ConvTileParams& P, MapEncoder enc, const void* enc_ctx, int n_views,
                                 const std::vector<TapSpec>& taps_in, const View5& outv, int out_rows_p,
                                 const void* w_packed, int w_rows_p, int w_taps, int kin_p, double* stats,
                                 const float* bias, cudaStream_t stream, bool allow_group,
                                 const BnReduce* red, int rows_lw, int org_h, bool rows_forced = false) {
  read_env_once();
  if (taps_in.empty()) return fail(kBadArg, "convolution has no valid taps");
  if ((int)taps_in.size() > kMaxTaps) return fail(kUnsupported, "too many filter taps (%d)", (int)taps_in.size());
  TileGeom& g = P.g;
  const int eN = (int)outv.dim[4], eT = (int)outv.dim[3], eH = (int)outv.dim[2], eW = (int)outv.dim[1];
  int min_t = 1 << 20, max_t = -(1 << 20), min_h = min_t, max_h = max_t, min_w = min_t, max_w = max_t;
  for (const TapSpec& t : taps_in) {
    min_t = std::min(min_t, t.ot); max_t = std::max(max_t, t.ot);
    min_h = std::min(min_h, t.oh); max_h = std::max(max_h, t.oh);
    min_w = std::min(min_w, t.ow); max_w = std::max(max_w, t.ow);
  }
  const int span_t = max_t - min_t + 1, span_h = max_h - min_h + 1, span_w = max_w - min_w + 1;
  enum { kPlain, kTemporal, kRows } mode = kPlain;
  if (g_halo_enabled && allow_group && (n_views == 1 || t_plane_views) && taps_in.size() > 1) {
    if (span_h == 1 && span_w == 1 && span_t > 1) {
      mode = kTemporal;
    } else if (span_h > 1 && rows_forced) {
      mode = kRows;
    } else if (span_h > 1) {
      // "rows" tiles (th x tw, halo along h, one box per kw) against one box per tap: estimated cycles per unit of
      // output = max(MMA issue/tensor time, L2->SM feed at ~42 B/clk/SM); padded tiles cost both
      // (constants from tests/diag/mma_rate.py and the TMA feed rate measured with conv_roles.py)
      int nb, bn, ln_;
      pick_block_n(out_rows_p, &nb, &bn, &ln_);
      const double mma_per_tile = (double)taps_in.size() * ceil_div(kin_p, 16) * std::max(bn / 2.0, 58.0);
      const double w_bytes = (double)taps_in.size() * bn * kin_p * 2 * 0.5;   // pair: half the rows per CTA
      const bool w_res = nb == 1 && w_bytes <= 112 * 1024;
      auto cost = [&](double tiles, double a_rows) {
        const double feed = (a_rows * kin_p * 2 + (w_res ? 0.0 : w_bytes)) / 42.0;
        return tiles * std::max(mma_per_tile, feed);
      };
      int pl[4];
      choose_tile(eN, eT, eH, eW, &pl[0], &pl[1], &pl[2], &pl[3]);
      const double plain_tiles = (double)ceil_div(eN, 1 << pl[0]) * ceil_div(eT, 1 << pl[1]) * ceil_div(eH, 1 << pl[2]) *
                                 ceil_div(eW, 1 << pl[3]);
      double best = cost(plain_tiles, 128.0 * taps_in.size());
      for (int lw = 3; lw <= 5; ++lw) {
        const int rtw = 1 << lw, rth = 128 >> lw;
        const int box_rows = (rth + span_h - 1) * rtw * span_t;
        if (box_rows * 128 > 64 * 1024) continue;
        const double tiles = (double)eN * eT * ceil_div(eH, rth) * ceil_div(eW, rtw);
        const double c = cost(tiles, (double)span_w * box_rows);
        if (c < best * 0.97) { best = c; mode = kRows; rows_lw = lw; }
      }
    }
  }
  uint32_t abox[5];  // A box (channels, w, h, t, n)
  if (mode == kTemporal) {
    double best = 1e30;
    int blw = 3, blh = 0, blt = 4;
    for (int a = 0; a <= 7; ++a)
      for (int b = 0; a + b <= 7; ++b) {
        if (a + b < 3) continue;               // th*tw must be a multiple of 8 rows
        const int c = 7 - a - b;
        const int tw = 1 << a, th = 1 << b, tt = 1 << c;
        if (tt + span_t - 1 > 256 || (tt + span_t - 1) * th * tw * 128 > 48 * 1024) continue;
        const double vol = (double)round_up(eW, tw) * round_up(eH, th) * round_up(eT, tt);
        const double cost = vol * (double)(tt + span_t - 1) / tt * (1.0 + 0.02 * (7 - a));
        if (cost < best) { best = cost; blw = a; blh = b; blt = c; }
      }
    g.lw = blw; g.lh = blh; g.lt = blt; g.ln = 0;
    abox[0] = kChunkK; abox[1] = 1u << g.lw; abox[2] = 1u << g.lh; abox[3] = (1u << g.lt) + span_t - 1; abox[4] = 1;
  } else if (mode == kRows) {
    g.lw = rows_lw; g.lh = 7 - rows_lw; g.lt = 0; g.ln = 0;
    abox[0] = kChunkK; abox[1] = 1u << g.lw; abox[2] = (1u << g.lh) + span_h - 1; abox[3] = span_t; abox[4] = 1;
  } else {
    choose_tile(eN, eT, eH, eW, &g.ln, &g.lt, &g.lh, &g.lw);
    abox[0] = kChunkK; abox[1] = 1u << g.lw; abox[2] = 1u << g.lh; abox[3] = 1u << g.lt; abox[4] = 1u << g.ln;
  }
  for (int i = 0; i < 4; ++i) P.a_box[i] = (int)abox[1 + i];
  g.ext_w = eW; g.ext_h = eH; g.ext_t = eT; g.ext_n = eN;
  g.org_h = org_h;
  g.tiles_w = ceil_div(eW, 1 << g.lw);
  g.tiles_h = ceil_div(eH, 1 << g.lh);
  g.step_h = 1 << g.lh;
  g.tiles_t = ceil_div(eT, 1 << g.lt);
  g.tiles_n = ceil_div(eN, 1 << g.ln);

  // ---- taps and groups
  int ntaps = 0, ngroups = 0, max_group = 1;
  if (mode == kPlain) {
    for (const TapSpec& t : taps_in) {
      Tap& tp = P.taps[ntaps];
      tp.map = (int8_t)t.view; tp.dt = (int8_t)t.ot; tp.dh = (int8_t)t.oh; tp.dw = (int8_t)t.ow;
      tp.widx = (int16_t)t.widx; tp.shift_rows = 0;
      P.group_len[ngroups++] = 1;
      ++ntaps;
    }
  } else {
    // groups: all taps with the same ow (temporal mode has a single ow); leader origin = (min_t, min_h, ow)
    const int bh = (int)abox[2], bw = (int)abox[1];
    // plane products: one group per (operand plane, weight plane, ow) - the kh / kt taps of ONE product share a box, so a
    // stage holds one box and one product's weight tiles (all weight planes of an operand plane in one group would not
    // leave room for three stages)
    const int n_wplanes = (t_plane_views && t_plane_taps > 0) ? ceil_div(w_taps, t_plane_taps) : 1;
    // (operand plane, weight plane) pairs with the smallest contributions (largest plane indices) first
    for (int rank = n_views + n_wplanes - 2; rank >= 0; --rank)
    for (int view = 0; view < n_views; ++view)
    for (int wp = 0; wp < n_wplanes; ++wp)
    for (int ow = min_w; ow <= max_w; ++ow) {
      if (view + wp != rank) continue;
      int len = 0;
      for (const TapSpec& t : taps_in) {
        if (t.ow != ow || t.view != view || (n_wplanes > 1 && t.widx / t_plane_taps != wp)) continue;
        Tap& tp = P.taps[ntaps++];
        tp.map = (int8_t)view; tp.dt = (int8_t)min_t; tp.dh = (int8_t)min_h; tp.dw = (int8_t)ow;
        tp.widx = (int16_t)t.widx;
        tp.shift_rows = (int16_t)(((t.ot - min_t) * bh + (t.oh - min_h)) * bw);
        ++len;
      }
      if (len == 0) continue;
      if (len > 255) return fail(kUnsupported, "tap group too large");
      P.group_len[ngroups++] = (uint8_t)len;
      max_group = std::max(max_group, len);
    }
  }
  P.num_taps = ntaps; P.num_groups = ngroups; P.max_group = max_group;
  const int a_rows = (int)(abox[1] * abox[2] * abox[3] * abox[4]);
  P.a_tx_bytes = a_rows * 128;
  P.a_stage_bytes = round_up(P.a_tx_bytes, 1024);

  for (int v = 0; v < n_views; ++v) {
    int rc = enc(&P.a_map[v], enc_ctx, v, abox);
    if (rc) return rc;
  }
  for (int i = n_views; i < kMaxAMaps; ++i) P.a_map[i] = P.a_map[0];

  pick_block_n(out_rows_p, &P.n_tiles, &P.block_n, &P.last_n);
  P.k_chunks = ceil_div(kin_p, kChunkK);
  P.k_steps_last = ceil_div(kin_p - (P.k_chunks - 1) * kChunkK, 16);
  const long long m_tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_t * g.tiles_n;
  // CTA pairs (one M=256 MMA over two position tiles, weight rows split between the CTAs) whenever there is
  // enough work to fill the chip with pairs
  const bool pair = g_pair_enabled == 2 || (g_pair_enabled && m_tiles * P.n_tiles >= 2LL * sm_count());   // 2 = force (tests)
  P.total_tiles = (int)((pair ? (m_tiles + 1) / 2 : m_tiles) * P.n_tiles);
  const int b_tap_bytes = (pair ? P.block_n / 2 : P.block_n) * 128;
  const bool xf = P.xf_ss != nullptr;
  const int xf_bytes = xf ? round_up(P.k_chunks * 512, 1024) : 0;      // staged scale / shift table
  const int out_bytes = kOutBufs * kOutBufBytes;
  const int avail = kSmemBudget - 1024 - out_bytes - xf_bytes;
  const int res_bytes = ntaps * P.k_chunks * b_tap_bytes;
  // weight-stationary when the whole filter of this channel tile fits next to >= 3 A stages and the CTA
  // amortises the load over several tiles
  P.b_resident = (g_resident_enabled && P.n_tiles == 1 && res_bytes <= 112 * 1024 &&
                  avail - res_bytes >= 3 * P.a_stage_bytes && m_tiles >= 3 * sm_count()) ? 1 : 0;
  const int stage_bytes = P.a_stage_bytes + (P.b_resident ? 0 : max_group * b_tap_bytes);
  P.stages = (avail - (P.b_resident ? res_bytes : 0)) / stage_bytes;
  if (P.stages > kMaxStages) P.stages = kMaxStages;
  if (P.stages < 3 && mode != kPlain)   // wide layers: a group's weight tiles do not fit -> one box per tap
    return conv_multi_tap_region(P, enc, enc_ctx, n_views, taps_in, outv, out_rows_p, w_packed, w_rows_p, w_taps,
                                 kin_p, stats, bias, stream, false, red, rows_lw, org_h);
  if (P.stages < 2) return fail(kUnsupported, "conv tile: not enough shared memory for 2 stages");
  // split issue: two issuing threads on two half-width accumulators where one thread is the limit - 64-column tiles
  // (one thread: >= 50 cycles per MMA against 32 of tensor time / 40 of operand fetch) with enough MMAs per tile that
  // the MMA loop, not the epilogue, paces the tile (144->64 temporal fprop, 27 MMAs: -20 %; 83->64, 18 MMAs: +10 %).
  // Wider tiles (the 88-channel stem) are epilogue-bound and lose to the second accumulator's loads
  // (DV_CONV_SPLIT_ISSUE=2 extends it to N <= 128 and any MMA count; tests/diag/layer_ab.py).
  const int mma_per_tile = ntaps * ((P.k_chunks - 1) * 4 + P.k_steps_last);
  // (fp32 mode: the merged plane-product launches that store the finished sum; both issuers walk the taps in order,
  // so each half accumulator still meets its smallest contributions first)
  P.split = (g_split_issue && (outv.esize == 2 || (outv.esize == 4 && outv.store && t_f32_split)) && P.n_tiles == 1 &&
             P.block_n <= (g_split_issue == 2 ? kMaxBlockN / 2 : 64) && (g_split_issue == 2 || mma_per_tile >= 24) &&
             P.group_len[0] * (P.k_chunks > 1 ? 4 : P.k_steps_last) >= 2) ? 1 : 0;
  if (P.split) {
    // the split kernel's issue program: per tap, A offset inside its group's box | weight-tile offset << 16 (16-byte units)
    int t = 0;
    for (int grp = 0; grp < ngroups; ++grp)
      for (int i = 0; i < P.group_len[grp]; ++i, ++t) {
        const uint32_t a_off = (uint32_t)P.taps[t].shift_rows * 8u;
        const uint32_t b_off = (uint32_t)(P.b_resident ? t * P.k_chunks * b_tap_bytes : P.a_stage_bytes + i * b_tap_bytes) >> 4;
        if (a_off > 0xffffu || b_off > 0xffffu) return fail(kUnsupported, "conv tile: operand offset out of range");
        P.prog[t] = a_off | (b_off << 16);
      }
    P.prog[ntaps] = 0;
  }
  P.stats = stats;
  P.stats_ld = out_rows_p;
  P.bias = bias;
  P.prof = g_prof;
  P.red_y = nullptr;
  P.red_ss = nullptr;
  for (int i = 0; i < 4; ++i) P.red_stride[i] = outv.stride[i + 1];
  P.out_f32 = outv.esize == 4 ? static_cast<float*>(const_cast<void*>(outv.base)) : nullptr;
  P.f32_store = outv.store;
  if (P.out_f32 != nullptr && (red != nullptr || (stats != nullptr && !outv.store)))
    return fail(kBadArg, "fp32 output: fused statistics only for a launch that stores the finished sum");
  if (xf && (P.out_f32 != nullptr || red != nullptr))
    return fail(kBadArg, "consumer-side BatchNorm is a forward, bf16-output launch");
  if (red != nullptr) {
    // red->y is already offset to the output view's origin (stride-parity class)
    P.stats = red->sums;
    P.red_y = red->y;
    P.red_ss = red->ss;
    const int lg[4] = {g.lw, g.lh, g.lt, g.ln};
    int k = 0;
    for (int d = 0; d < 4; ++d)
      for (int b = 0; b < lg[d]; ++b) P.red_bitoff[k++] = outv.stride[d + 1] << b;
  }

  {  // weights: [rows][taps][kin_p] bf16, box (64, 1, block_n)
    uint64_t dims[3] = {(uint64_t)kin_p, (uint64_t)w_taps, (uint64_t)w_rows_p};
    uint64_t strides[3] = {2, (uint64_t)kin_p * 2, (uint64_t)kin_p * w_taps * 2};
    uint32_t box[3] = {kChunkK, 1, (uint32_t)(pair ? P.block_n / 2 : P.block_n)};
    int rc = encode_tmap(&P.b_map, w_packed, 2, 3, dims, strides, box, true);
    if (rc) return rc;
  }
  if (P.out_f32 == nullptr) {
    uint32_t box[5] = {kChunkK, 1u << g.lw, 1u << g.lh, 1u << g.lt, 1u << g.ln};
    int rc = encode_view(&P.out_map, outv, box);
    if (rc) return rc;
  }
  const int smem_bytes = 1024 + xf_bytes + (P.b_resident ? res_bytes : 0) + P.stages * stage_bytes + out_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget));
    attr_set = true;
  }
  // CTAs (or CTA pairs): one per SM (pair of SMs), a multiple of the channel-tile count
  int units = (pair ? sm_count() / 2 : sm_count()) / P.n_tiles * P.n_tiles;
  if (units > P.total_tiles) units = P.total_tiles;  // total_tiles is a multiple of n_tiles
  const bool f32 = P.out_f32 != nullptr;
  if (!pair) {
    if (xf && P.split) conv_tile_kernel<false, false, true, true><<<units, kNumThreads + kXfThreads + kIssue2Threads, smem_bytes, stream>>>(P);
    else if (xf) conv_tile_kernel<false, false, true><<<units, kNumThreads + kXfThreads, smem_bytes, stream>>>(P);
    else if (f32 && P.split) conv_tile_kernel<false, true, false, true><<<units, kNumThreads + kIssue2Threads, smem_bytes, stream>>>(P);
    else if (f32) conv_tile_kernel<false, true><<<units, kNumThreads, smem_bytes, stream>>>(P);
    else if (P.split) conv_tile_kernel<false, false, false, true><<<units, kNumThreads + kIssue2Threads, smem_bytes, stream>>>(P);
    else conv_tile_kernel<false><<<units, kNumThreads, smem_bytes, stream>>>(P);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * units);
    cfg.blockDim = dim3((xf ? kNumThreads + kXfThreads : kNumThreads) + (P.split ? kIssue2Threads : 0));
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (xf && P.split) DV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tile_kernel<true, false, true, true>, P));
    else if (xf) DV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tile_kernel<true, false, true>, P));
    else if (f32 && P.split) DV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tile_kernel<true, true, false, true>, P));
    else if (f32) DV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tile_kernel<true, true>, P));
    else if (P.split) DV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tile_kernel<true, false, false, true>, P));
    else DV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tile_kernel<true>, P));
  }
  DV_LAUNCH_OK();
  return kOk;
}

// out(view) = sum_taps A_view(tap)[box shifted by tap] * W[tap].
// Spatial filters on maps whose height is 8 (mod 16), e.g. 56: the 16 x 8 "rows" tile would pad the last tile row
// by half (12.5 % of all MMAs at 56 x 56). The map is split instead: rows [0, H-8) with 16 x 8 tiles and the last 8
// rows with 8 x 16 tiles, two launches accumulating into the same BatchNorm statistics.
static int conv_multi_tap(ConvTileParams& P, MapEncoder enc, const void* enc_ctx, int n_views,
                          const std::vector<TapSpec>& taps_in, const View5& outv, int out_rows_p,
                          const void* w_packed, int w_rows_p, int w_taps, int kin_p, double* stats,
                          const float* bias, cudaStream_t stream, bool allow_group = true,
                          const BnReduce* red = nullptr) {
  read_env_once();
  const int eH = (int)outv.dim[2], eW = (int)outv.dim[1];
  bool spatial = false;
  for (const TapSpec& t : taps_in) spatial = spatial || t.oh != taps_in[0].oh;
  if (g_halo_enabled && g_split_enabled && allow_group && (n_views == 1 || t_plane_views) && spatial && eH % 16 == 8 && eH >= 24 &&
      (double)round_up(eW, 8) / eW <= 1.16 && (double)round_up(eW, 16) / eW <= 1.16) {
    const int h_main = eH - 8;
    View5 va = outv, vb = outv;
    va.dim[2] = h_main;
    vb.dim[2] = 8;
    vb.base = static_cast<const uint8_t*>(outv.base) + (long long)h_main * outv.stride[2] * outv.esize;
    BnReduce rb;
    if (red != nullptr) {
      rb = *red;
      rb.y = static_cast<const uint8_t*>(red->y) + (long long)h_main * outv.stride[2] * 2;
    }
    int rc = conv_multi_tap_region(P, enc, enc_ctx, n_views, taps_in, va, out_rows_p, w_packed, w_rows_p, w_taps,
                                   kin_p, stats, bias, stream, true, red, 3, 0, true);
    if (rc) return rc;
    return conv_multi_tap_region(P, enc, enc_ctx, n_views, taps_in, vb, out_rows_p, w_packed, w_rows_p, w_taps,
                                 kin_p, stats, bias, stream, true, red != nullptr ? &rb : nullptr, 4, h_main, true);
  }
  return conv_multi_tap_region(P, enc, enc_ctx, n_views, taps_in, outv, out_rows_p, w_packed, w_rows_p, w_taps, kin_p,
                               stats, bias, stream, allow_group, red, 3, 0);
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

struct ViewSet {
  View5 v[kMaxAMaps];
};
static int encode_from_viewset(CUtensorMap* m, const void* ctx, int view, const uint32_t box[5]) {
  return encode_view(m, static_cast<const ViewSet*>(ctx)->v[view], box);
}

int conv_fprop_bf16(const void* x, const void* w_packed, void* y, double* stats, const float* bias,
                    const ConvGeom& c, cudaStream_t stream, int y_f32, const float* xf_ss, int xf_relu) {
  static thread_local ConvTileParams P;
  // xf_ss != NULL: x is the RAW output of the convolution below and the operand is relu?(scale*x + shift) (bn_xform.cuh)
  P.xf_ss = xf_ss;
  P.xf_cp = c.Cin_p;
  P.xf_relu = xf_relu;
  const View5 inv = make_ndhwc(x, c.N, c.T, c.H, c.W, c.Cin_p);
  View5 outv = make_ndhwc(y, c.N, c.To, c.Ho, c.Wo, c.Cout_p);
  if (y_f32) { outv.esize = 4; outv.store = y_f32 == 2; }   // y is float [N][To][Ho][Wo][Cout_p]: 1 = added to, 2 = overwritten
  ViewSet vs;
  int map_of_parity[8];
  for (int i = 0; i < 8; ++i) map_of_parity[i] = -1;
  int nviews = 0;
  std::vector<TapSpec> taps;
  for (int a = 0; a < c.kt; ++a)
    for (int b = 0; b < c.kh; ++b)
      for (int d = 0; d < c.kw; ++d) {
        const int ot = a - c.pt, oh = b - c.ph, ow = d - c.pw;
        const int rt = posmod(ot, c.st), rh = posmod(oh, c.sh), rw = posmod(ow, c.sw);
        if (rt > 1 || rh > 1 || rw > 1) return fail(kUnsupported, "conv stride > 2 not supported");
        if (rt >= c.T || rh >= c.H || rw >= c.W) continue;  // tap never touches real data
        const int key = (rt * 2 + rh) * 2 + rw;
        if (map_of_parity[key] < 0) {
          View5 v = inv;
          subsample(v, 3, rt, c.st);
          subsample(v, 2, rh, c.sh);
          subsample(v, 1, rw, c.sw);
          vs.v[nviews] = v;
          for (int i = 0; i < 4; ++i) P.a_dims[nviews][i] = (int)v.dim[1 + i];
          map_of_parity[key] = nviews++;
        }
        taps.push_back({map_of_parity[key], floordiv(ot, c.st), floordiv(oh, c.sh), floordiv(ow, c.sw),
                        (a * c.kh + b) * c.kw + d});
      }
  return conv_multi_tap(P, encode_from_viewset, &vs, nviews, taps, outv, c.Cout_p, w_packed, c.Cout_p,
                        c.kt * c.kh * c.kw, c.Cin_p, stats, bias, stream);
}

// dX = dgrad(dY, W): one launch per stride-parity class of dX positions; each class is a
// stride-1 multi-tap GEMM over dY with the subset of taps that reach it.
// w_packed_t: [Cin_p][taps][Cout_p] bf16 (transposed pack).
int conv_dgrad_bf16(const void* dy, const void* w_packed_t, void* dx, const ConvGeom& c,
                    cudaStream_t stream, const BnReduce* red, int dx_f32) {
  static thread_local ConvTileParams P;
  P.xf_ss = nullptr;
  ViewSet vs;
  vs.v[0] = make_ndhwc(dy, c.N, c.To, c.Ho, c.Wo, c.Cout_p);
  View5 dxv = make_ndhwc(dx, c.N, c.T, c.H, c.W, c.Cin_p);
  if (dx_f32) { dxv.esize = 4; dxv.store = dx_f32 == 2; }   // dx is float [N][T][H][W][Cin_p]: 1 = added to, 2 = overwritten
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
  // positions no tap reaches (e.g. 1x1 stride-2 convs) receive no tile: zero them (not in add mode: the caller did)
  if (need_zero && dx_f32 != 1)
    DV_CUDA_OK(cudaMemsetAsync(dx, 0, (size_t)c.N * c.T * c.H * c.W * c.Cin_p * (dx_f32 ? 4 : 2), stream));
  for (int rt = 0; rt < c.st; ++rt)
    for (int rh = 0; rh < c.sh; ++rh)
      for (int rw = 0; rw < c.sw; ++rw) {
        if (rt >= c.T || rh >= c.H || rw >= c.W) continue;
        View5 ov = dxv;
        subsample(ov, 3, rt, c.st);
        subsample(ov, 2, rh, c.sh);
        subsample(ov, 1, rw, c.sw);
        std::vector<TapSpec> taps;
        for (int a = 0; a < c.kt; ++a) {
          if (posmod(rt + c.pt - a, c.st) != 0) continue;
          for (int b = 0; b < c.kh; ++b) {
            if (posmod(rh + c.ph - b, c.sh) != 0) continue;
            for (int d = 0; d < c.kw; ++d) {
              if (posmod(rw + c.pw - d, c.sw) != 0) continue;
              taps.push_back({0, (rt + c.pt - a) / c.st, (rh + c.ph - b) / c.sh, (rw + c.pw - d) / c.sw,
                              (a * c.kh + b) * c.kw + d});
            }
          }
        }
        if (taps.empty()) continue;  // class receives no gradient (zero-filled above)
        BnReduce rv;
        if (red != nullptr) {
          // y has dx's layout: same parity-class origin and strides as the output view
          rv = *red;
          rv.y = static_cast<const uint8_t*>(red->y) +
                 (static_cast<const uint8_t*>(ov.base) - static_cast<const uint8_t*>(dxv.base));
        }
        int rc = conv_multi_tap(P, encode_from_viewset, &vs, 1, taps, ov, c.Cin_p, w_packed_t, c.Cin_p,
                                c.kt * c.kh * c.kw, c.Cout_p, nullptr, nullptr, stream, true,
                                red != nullptr ? &rv : nullptr);
        if (rc) return rc;
      }
  return kOk;
}

// Stride-2 7x7 stem (Cin <= 4) as a 4-tap (per temporal tap) stride-1 GEMM on the space-to-depth input.
// x_s2d: bf16 [N][T][H2][W2+3][16], channel = (rh*2+rw)*4 + c, two zero columns left / one right
// (written by the ingest kernel). For output (ho,wo) and row tap a in {-2..1} the A row is the 128-byte
// window x_s2d[n][t][ho+a][wo .. wo+3][0..15] — an OVERLAPPING-window tensor map (W stride 32 B, inner
// extent 128 B), so K = 64 per tap instead of 49 taps of K = 16. The 4 row taps form one halo group.
// w_stem: bf16 [Cout_p][kt*4][64] (pack_stem_weights).
struct StemCtx {
  const void* x;
  int N, T, H2, W2;
};
static int encode_stem_map(CUtensorMap* m, const void* ctx, int /*view*/, const uint32_t box[5]) {
  const StemCtx* s = static_cast<const StemCtx*>(ctx);
  const long long W2p = s->W2 + 3;
  uint64_t dims[5] = {64, (uint64_t)s->W2, (uint64_t)s->H2, (uint64_t)s->T, (uint64_t)s->N};
  uint64_t strides[5] = {2, 32, (uint64_t)W2p * 32, (uint64_t)s->H2 * W2p * 32,
                         (uint64_t)s->T * s->H2 * W2p * 32};
  return encode_tmap(m, s->x, 2, 5, dims, strides, box, true);
}

// fp32 mode: y (fp32) = sum over the plane products (i, j), i + j < K, of conv(x_i, w_j) in ONE launch - the products are
// extra taps of one accumulator (smallest contributions first), the epilogue stores the fp32 rows once instead of one
// store + five read-modify-write passes (dv_conv3d_fprop_f32acc per product). x_planes: bf16 [K][N][T][H][W][Cin_p];
// wf_all: bf16 [Cout_p][K * taps][Cin_p] (plane j in tap slots [j * taps, (j + 1) * taps)). Stride 1 only.
int conv_fprop_f32planes(const void* x_planes, long long plane_stride, int K, const void* wf_all, float* y,
                         double* stats, const float* bias, const ConvGeom& c, cudaStream_t stream) {
  if (K < 1 || K > 3) return fail(kBadArg, "1..3 split planes");
  static thread_local ConvTileParams P;
  P.xf_ss = nullptr;
  const int taps_total = c.kt * c.kh * c.kw;
  const bool unit_stride = c.st == 1 && c.sh == 1 && c.sw == 1;
  // views: (stride-parity class, plane) - a strided layer reads each plane through one subsampled map per class
  ViewSet vs;
  int view_of[8][3];
  for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) view_of[i][k] = -1;
  int nviews = 0;
  struct Hit { int key, ot, oh, ow, widx; };
  std::vector<Hit> hits;
  for (int a = 0; a < c.kt; ++a)
    for (int b = 0; b < c.kh; ++b)
      for (int d = 0; d < c.kw; ++d) {
        const int ot = a - c.pt, oh = b - c.ph, ow = d - c.pw;
        const int rt = posmod(ot, c.st), rh = posmod(oh, c.sh), rw = posmod(ow, c.sw);
        if (rt > 1 || rh > 1 || rw > 1) return fail(kUnsupported, "conv stride > 2 not supported");
        if (rt >= c.T || rh >= c.H || rw >= c.W) continue;  // tap never touches real data
        const int key = (rt * 2 + rh) * 2 + rw;
        if (view_of[key][0] < 0) {
          if (nviews + K > kMaxAMaps)
            return fail(kUnsupported, "merged plane products: more than %d (parity class, plane) views", kMaxAMaps);
          for (int i = 0; i < K; ++i) {
            View5 v = make_ndhwc(static_cast<const uint8_t*>(x_planes) + (long long)i * plane_stride * 2, c.N, c.T, c.H,
                                 c.W, c.Cin_p);
            subsample(v, 3, rt, c.st);
            subsample(v, 2, rh, c.sh);
            subsample(v, 1, rw, c.sw);
            vs.v[nviews] = v;
            for (int q = 0; q < 4; ++q) P.a_dims[nviews][q] = (int)v.dim[1 + q];
            view_of[key][i] = nviews++;
          }
        }
        hits.push_back({key, floordiv(ot, c.st), floordiv(oh, c.sh), floordiv(ow, c.sw), (a * c.kh + b) * c.kw + d});
      }
  View5 outv = make_ndhwc(y, c.N, c.To, c.Ho, c.Wo, c.Cout_p);
  outv.esize = 4; outv.store = 1;
  std::vector<TapSpec> taps;
  for (int s = 2 * (K - 1); s >= 0; --s)            // i + j = s: smallest contributions first
    for (int i = 0; i < K; ++i) {
      const int j = s - i;
      if (j < 0 || j >= K - i) continue;            // products with i + j < K only
      for (const Hit& h : hits) taps.push_back({view_of[h.key][i], h.ot, h.oh, h.ow, j * taps_total + h.widx});
    }
  // stride 1: view = plane, the taps of one plane product may share halo boxes; strided layers keep one box per tap
  t_plane_views = unit_stride;
  t_plane_taps = taps_total;
  // (no two-issuer instance here: measured on the 144->64 temporal layer it is 9 % slower than one issuer - 1.43 against
  // 1.31 ms at 48 clips - while the 64<-144 data gradient gains 22 %; profiles/r02d_fp32_merged_plane_products.txt)
  const int rc = conv_multi_tap(P, encode_from_viewset, &vs, nviews, taps, outv, c.Cout_p, wf_all, c.Cout_p,
                                K * taps_total, c.Cin_p, stats, bias, stream);
  t_plane_views = false;
  t_f32_split = false;
  return rc;
}

// fp32 mode: dx (fp32) = sum over the plane products of dgrad(dy_i, w_j) in one launch per stride-parity class.
// dy_planes: bf16 [K][N][To][Ho][Wo][Cout_p]; wt_all: bf16 [Cin_p][K * taps][Cout_p].
int conv_dgrad_f32planes(const void* dy_planes, long long plane_stride, int K, const void* wt_all, float* dx,
                         const ConvGeom& c, cudaStream_t stream) {
  if (K < 1 || K > 3) return fail(kBadArg, "1..3 split planes");
  static thread_local ConvTileParams P;
  P.xf_ss = nullptr;
  const int taps_total = c.kt * c.kh * c.kw;
  ViewSet vs;
  for (int i = 0; i < K; ++i)
    vs.v[i] = make_ndhwc(static_cast<const uint8_t*>(dy_planes) + (long long)i * plane_stride * 2, c.N, c.To, c.Ho, c.Wo,
                         c.Cout_p);
  View5 dxv = make_ndhwc(dx, c.N, c.T, c.H, c.W, c.Cin_p);
  dxv.esize = 4; dxv.store = 1;
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
  if (need_zero) DV_CUDA_OK(cudaMemsetAsync(dx, 0, (size_t)c.N * c.T * c.H * c.W * c.Cin_p * 4, stream));
  for (int rt = 0; rt < c.st; ++rt)
    for (int rh = 0; rh < c.sh; ++rh)
      for (int rw = 0; rw < c.sw; ++rw) {
        if (rt >= c.T || rh >= c.H || rw >= c.W) continue;
        View5 ov = dxv;
        subsample(ov, 3, rt, c.st);
        subsample(ov, 2, rh, c.sh);
        subsample(ov, 1, rw, c.sw);
        std::vector<TapSpec> taps;
        for (int s = 2 * (K - 1); s >= 0; --s)
          for (int i = 0; i < K; ++i) {
            const int j = s - i;
            if (j < 0 || j >= K - i) continue;
            for (int a = 0; a < c.kt; ++a) {
              if (posmod(rt + c.pt - a, c.st) != 0) continue;
              for (int b = 0; b < c.kh; ++b) {
                if (posmod(rh + c.ph - b, c.sh) != 0) continue;
                for (int d = 0; d < c.kw; ++d) {
                  if (posmod(rw + c.pw - d, c.sw) != 0) continue;
                  taps.push_back({i, (rt + c.pt - a) / c.st, (rh + c.ph - b) / c.sh, (rw + c.pw - d) / c.sw,
                                  j * taps_total + (a * c.kh + b) * c.kw + d});
                }
              }
            }
          }
        if (taps.empty()) continue;
        for (int i = 0; i < K; ++i)
          for (int d = 0; d < 4; ++d) P.a_dims[i][d] = (int)vs.v[i].dim[1 + d];
        t_plane_views = true;
        t_plane_taps = taps_total;
        t_f32_split = f32_split_enabled();
        const int rc = conv_multi_tap(P, encode_from_viewset, &vs, K, taps, ov, c.Cin_p, wt_all, c.Cin_p, K * taps_total,
                                      c.Cout_p, nullptr, nullptr, stream);
        t_plane_views = false;
        t_f32_split = false;
        if (rc) return rc;
      }
  return kOk;
}

// Eligibility of the kh-stacked data gradient (see kStack): 3x3 spatial filter, stride 1, padding 1, 64 input channels,
// H a multiple of 14 and W of 8, and enough tiles to fill the chip with CTA pairs.
int conv_dgrad_stack_ok(const ConvGeom& c) {
  const long long tiles = (long long)c.N * c.T * (c.H / 14) * (c.W / 8);
  return c.kt == 1 && c.kh == 3 && c.kw == 3 && c.st == 1 && c.sh == 1 && c.sw == 1 && c.pt == 0 && c.ph == 1 && c.pw == 1 &&
         c.Cin_p == 64 && c.H % 14 == 0 && c.W % 8 == 0 && c.To == c.T && c.Ho == c.H && c.Wo == c.W &&
         tiles >= 2LL * sm_count() && ceil_div(c.Cout_p, 64) * 3 * 96 * 128 <= 112 * 1024;
}

// dX = dgrad(dY, W) with the kh taps stacked along N. w_stack: bf16 [192][3][Cout_p]: row s*64 + ci, tap kw holds
// W[co][ci][kh = 2 - s][kw] (the engine permutes the transposed pack).
int conv_dgrad_stack_bf16(const void* dy, const void* w_stack, void* dx, const ConvGeom& c, cudaStream_t stream,
                          const BnReduce* red) {
  if (!conv_dgrad_stack_ok(c)) return fail(kUnsupported, "stacked dgrad: geometry not eligible");
  static thread_local ConvTileParams P;
  P = ConvTileParams{};
  TileGeom& g = P.g;
  g.lw = 3; g.lh = 4; g.lt = 0; g.ln = 0;
  g.step_h = 14;
  g.tiles_w = c.W / 8; g.tiles_h = c.H / 14; g.tiles_t = c.T; g.tiles_n = c.N;
  g.ext_w = c.W; g.ext_h = c.H; g.ext_t = c.T; g.ext_n = c.N;
  g.org_h = 0;
  for (int kw = 0; kw < 3; ++kw) {
    Tap& tp = P.taps[kw];
    tp.map = 0; tp.dt = 0; tp.dh = -1; tp.dw = (int8_t)(1 - kw); tp.widx = (int16_t)kw; tp.shift_rows = 0;
    P.group_len[kw] = 1;
  }
  P.num_taps = 3; P.num_groups = 3; P.max_group = 1;
  P.a_tx_bytes = 128 * 128;
  P.a_stage_bytes = 128 * 128;
  P.k_chunks = ceil_div(c.Cout_p, kChunkK);
  P.k_steps_last = ceil_div(c.Cout_p - (P.k_chunks - 1) * kChunkK, 16);
  P.n_tiles = 1; P.block_n = 192; P.last_n = 192;
  const long long m_tiles = (long long)g.tiles_w * g.tiles_h * g.tiles_t * g.tiles_n;
  P.total_tiles = (int)((m_tiles + 1) / 2);
  const int b_tap_bytes = 96 * 128;
  const int res_bytes = 3 * P.k_chunks * b_tap_bytes;
  const int out_bytes = kOutBufBytes + 12 * 1024;         // one staging buffer + the two groups' exchange buffers
  const int avail = kSmemBudget - 1024 - out_bytes;
  P.b_resident = 1;
  P.stages = (avail - res_bytes) / P.a_stage_bytes;
  if (P.stages > kMaxStages) P.stages = kMaxStages;
  if (P.stages < 3) return fail(kUnsupported, "stacked dgrad: not enough shared memory");
  P.stats = nullptr; P.stats_ld = c.Cin_p; P.bias = nullptr; P.prof = g_prof;   // (g_prof: diagnostics build only)
  P.out_f32 = nullptr; P.xf_ss = nullptr; P.split = 0;
  const View5 dyv = make_ndhwc(dy, c.N, c.T, c.H, c.W, c.Cout_p);
  const View5 dxv = make_ndhwc(dx, c.N, c.T, c.H, c.W, c.Cin_p);
  for (int i = 0; i < 4; ++i) P.red_stride[i] = dxv.stride[i + 1];
  if (red != nullptr) {
    P.stats = red->sums;
    P.red_y = red->y;
    P.red_ss = red->ss;
  }
  {
    const uint32_t abox[5] = {kChunkK, 8, 16, 1, 1};
    int rc = encode_view(&P.a_map[0], dyv, abox);
    if (rc) return rc;
    for (int i = 1; i < kMaxAMaps; ++i) P.a_map[i] = P.a_map[0];
    uint64_t dims[3] = {(uint64_t)c.Cout_p, 3, 192};
    uint64_t strides[3] = {2, (uint64_t)c.Cout_p * 2, (uint64_t)c.Cout_p * 3 * 2};
    uint32_t bbox[3] = {kChunkK, 1, 96};
    rc = encode_tmap(&P.b_map, w_stack, 2, 3, dims, strides, bbox, true);
    if (rc) return rc;
    const uint32_t obox[5] = {kChunkK, 8, 14, 1, 1};
    rc = encode_view(&P.out_map, dxv, obox);
    if (rc) return rc;
  }
  const int smem_bytes = 1024 + res_bytes + P.stages * P.a_stage_bytes + out_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    DV_CUDA_OK(cudaFuncSetAttribute(conv_tile_kernel<true, false, false, false, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    attr_set = true;
  }
  int units = sm_count() / 2;
  if (units > P.total_tiles) units = P.total_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * units);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  DV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tile_kernel<true, false, false, false, true>, P));
  DV_LAUNCH_OK();
  return kOk;
}

int conv_stem_fprop_bf16(const void* x_s2d, const void* w_stem, void* y, double* stats, const float* bias,
                         int N, int T, int H2, int W2, int Cout_p, int kt, int pt, cudaStream_t stream,
                         int y_f32) {
  static thread_local ConvTileParams P;
  P.xf_ss = nullptr;
  const int To = T + 2 * pt - kt + 1;
  View5 outv = make_ndhwc(y, N, To, H2, W2, Cout_p);
  if (y_f32) { outv.esize = 4; outv.store = y_f32 == 2; }
  StemCtx sc = {x_s2d, N, T, H2, W2};
  std::vector<TapSpec> taps;
  for (int a = 0; a < kt; ++a)
    for (int r = 0; r < 4; ++r) taps.push_back({0, a - pt, r - 2, 0, a * 4 + r});
  return conv_multi_tap(P, encode_stem_map, &sc, 1, taps, outv, Cout_p, w_stem, Cout_p, kt * 4, 64, stats,
                        bias, stream);
}

}  // namespace dv
