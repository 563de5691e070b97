// Debug probe (not on the product path): does TMA accept a tensor map whose dim-1 stride is smaller
// than the dim-0 extent (overlapping 128-byte windows)? Used to validate the space-to-depth stem.
#include "../host_common.h"
#include "../ptx.cuh"

namespace dv {

__global__ void probe_overlap_kernel(const __grid_constant__ CUtensorMap m, uint16_t* out, int c1) {
  __shared__ __align__(1024) uint8_t buf[8 * 128];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 8 * 128);
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(buf)),
        "l"(reinterpret_cast<uint64_t>(&m)), "r"(smem_u32(&bar)), "r"(0), "r"(c1)
        : "memory");
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(buf)[i];
}

// src: device buffer of >= 4096 uint16 holding src[i] = i. Loads 8 windows of 64 elements starting at
// window index c1 with window stride 16 elements; writes the raw (swizzled) smem image to out.
int probe_overlap(const void* src, void* out, int c1, cudaStream_t stream) {
  CUtensorMap m;
  uint64_t dims[2] = {64, 200};
  uint64_t strides[2] = {2, 32};
  uint32_t box[2] = {64, 8};
  int rc = encode_tmap(&m, src, 2, 2, dims, strides, box, true);
  if (rc) return rc;
  probe_overlap_kernel<<<1, 128, 0, stream>>>(m, (uint16_t*)out, c1);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv

// ---------------------------------------------------------------------------------------------------
// Debug microbenchmark (tests/diag/mma_rate.py): steady-state cycles per tcgen05.mma (M=128, N=n, K=16,
// bf16, both operands in shared memory, SWIZZLE_128B K-major) issued back to back by one thread per CTA,
// operands cycling through `region` bytes of shared memory. Tells the SS-mode operand-fetch limit apart
// from issue-loop overhead in conv_tile_kernel. mode 1: every MMA re-uses one B tile (weight-stationary).
namespace dv {

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int n_mma, int region, int mode, long long* out) {
  extern __shared__ uint8_t dsm[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ __align__(8) uint64_t ring_bar[8];
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(dsm);
  const uint32_t base = (raw + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < region / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(dsm + (base - raw))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&ring_bar[i], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t flags = 1u << 16;
    const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)n * 128;   // one 64-wide K chunk = 4 K-steps
    const uint32_t pair = a_bytes + (mode == 1 ? 0 : b_bytes);
    const uint32_t b_fixed = base + (uint32_t)region - b_bytes;
    const uint32_t span = (mode == 1 ? (uint32_t)region - b_bytes : (uint32_t)region) / pair * pair;
    uint32_t off = 0;
    const long long t0 = clock64();
    if (mode == 3) {
      // straight-line issue: 16 MMAs per iteration whose descriptors were all computed before the loop - the pure
      // issue rate of tcgen05.mma, without any address arithmetic between two MMAs
      uint32_t al[16], bl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t o = (uint32_t)(j >> 2) * pair;
        al[j] = (flags | ((base + o) >> 4)) + 2u * (j & 3);
        bl[j] = (flags | ((base + o + a_bytes) >> 4)) + 2u * (j & 3);
      }
      for (int i = 0; i < n_mma; i += 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) umma_bf16_lohi(tmem, al[j], bl[j], hi, idesc, (i | j) > 0);
      }
      const long long t1b = clock64();
      umma_commit(&done_bar);
      mbar_wait(&done_bar, 0);
      const long long t2b = clock64();
      out[blockIdx.x * 2] = t1b - t0;
      out[blockIdx.x * 2 + 1] = t2b - t0;
    } else {
    const int period = mode >= 16 ? (mode & 0xff) : 0;   // commit to a rotating barrier every `period` MMAs
    const bool alt_acc = (mode & 0x100) != 0;              // switch accumulator (TMEM columns 0 / 256) at every commit
    int since = 0, slot = 0;
    uint32_t acc_off = 0;
    for (int i = 0; i < n_mma; i += 4) {
      const uint32_t a = flags | ((base + off) >> 4);
      const uint32_t b = flags | ((mode == 1 ? b_fixed : base + off + a_bytes) >> 4);
      umma_bf16_lohi(tmem + acc_off, a, b, hi, idesc, i > 0);
      umma_bf16_lohi(tmem + acc_off, a + 2, b + 2, hi, idesc, 1);
      umma_bf16_lohi(tmem + acc_off, a + 4, b + 4, hi, idesc, 1);
      umma_bf16_lohi(tmem + acc_off, a + 6, b + 6, hi, idesc, 1);
      off += pair;
      if (off >= span) off = 0;
      since += 4;
      if (period && since >= period) {
        umma_commit(&ring_bar[slot]);
        slot = (slot + 1) & 7;
        since = 0;
        if (alt_acc) acc_off ^= 256u;
      }
    }
    const long long t1 = clock64();
    umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;      // issue time
    out[blockIdx.x * 2 + 1] = t2 - t0;  // until the last MMA completed
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 512);
  }
}

// mode 2: the same loop as M=256 MMAs over CTA pairs (cta_group::2): each CTA holds its 128 rows of A and half
// of B's rows; only the leader issues. Sub-modes (mode >> 8 after the low byte 2):
//   0x002  one issuing thread, loop of 4 MMAs per operand pair, no commits
//   0x102  one issuing thread, 12 straight-line MMAs (descriptors formed before them) + one multicast commit to a
//          rotating barrier per 12 - the shape of one 3-tap stage of conv_tile_kernel
//   0x202  TWO issuing threads (warps 0 and 1 of the leader), each the 0x102 loop on half of the MMAs, separate
//          accumulators (TMEM columns 0 / 256) and separate operand regions
//   0x302  two issuing threads, each 6 of the 12 MMAs of every stage (same operand tiles), separate accumulators
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma_rate_pair_kernel(int n, int n_mma, int region, int sub, long long* out) {
  extern __shared__ uint8_t dsm[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ __align__(8) uint64_t ring_bar[8];
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(dsm);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < region / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(dsm + (base - raw))[i] = 0x3c003c00u;
  const int issuers = sub >= 2 ? 2 : 1;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, issuers);
    for (int i = 0; i < 8; ++i) mbar_init(&ring_bar[i], issuers);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc2(&tmem_slot, 512);
    tmem_relinquish2();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < 2) {
    t0 = clock64();
    if (rank == 0 && w < issuers) {
      const uint32_t idesc = make_idesc_bf16(256, n, 0, 0);
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t flags = 1u << 16;
      const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)(n / 2) * 128;
      const uint32_t pair = a_bytes + ((b_bytes + 1023u) & ~1023u);
      if (sub == 0) {
        const uint32_t span = (uint32_t)region / pair * pair;
        uint32_t off = 0;
        for (int i = 0; i < n_mma; i += 4) {
          const uint32_t a = flags | ((base + off) >> 4);
          const uint32_t b = flags | ((base + off + a_bytes) >> 4);
          umma2_bf16_lohi(tmem, a, b, hi, idesc, i > 0);
          umma2_bf16_lohi(tmem, a + 2, b + 2, hi, idesc, 1);
          umma2_bf16_lohi(tmem, a + 4, b + 4, hi, idesc, 1);
          umma2_bf16_lohi(tmem, a + 6, b + 6, hi, idesc, 1);
          off += pair;
          if (off >= span) off = 0;
        }
      } else {
        // a "stage" = one A box of 18 KB with three tap views shifted by 1024 B + three B tiles
        const uint32_t stage = 18 * 1024 + 3 * ((b_bytes + 1023u) & ~1023u);
        const uint32_t half = (uint32_t)region / 2 / stage * stage;
        const uint32_t org = (sub == 2) ? (uint32_t)w * half : 0u;
        const uint32_t span = (sub == 2) ? half : (uint32_t)region / stage * stage;
        const uint32_t acc = tmem + (uint32_t)w * 256u;
        const int per_stage = (sub == 3) ? 6 : 12;
        const int mine = n_mma / issuers;
        uint32_t off = 0;
        int slot = 0;
        for (int i = 0; i < mine; i += per_stage) {
          const uint32_t a = flags | ((base + org + off) >> 4);
          const uint32_t b = flags | ((base + org + off + 18 * 1024) >> 4);
          const uint32_t bt = ((b_bytes + 1023u) & ~1023u) >> 4;
          if (sub == 3) {
            // this thread's half of the stage: K-steps {0,1} (w = 0) or {2,3} (w = 1) of the three taps
            const uint32_t k0 = (uint32_t)w * 4u;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              umma2_bf16_lohi(acc, a + t * 64 + k0, b + t * bt + k0, hi, idesc, (i | t) > 0);
              umma2_bf16_lohi(acc, a + t * 64 + k0 + 2, b + t * bt + k0 + 2, hi, idesc, 1);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              umma2_bf16_lohi(acc, a + t * 64, b + t * bt, hi, idesc, (i | t) > 0);
              umma2_bf16_lohi(acc, a + t * 64 + 2, b + t * bt + 2, hi, idesc, 1);
              umma2_bf16_lohi(acc, a + t * 64 + 4, b + t * bt + 4, hi, idesc, 1);
              umma2_bf16_lohi(acc, a + t * 64 + 6, b + t * bt + 6, hi, idesc, 1);
            }
          }
          umma2_commit_both(&ring_bar[slot]);
          slot = (slot + 1) & 7;
          off += stage;
          if (off >= span) off = 0;
        }
      }
      t1 = clock64();
      umma2_commit_both(&done_bar);
    }
    mbar_wait(&done_bar, 0);
    const long long t2 = clock64();
    if (w == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after_sync();
    tmem_dealloc2(tmem, 512);
  }
}

int mma_rate(int n, int n_mma, int region, int mode, long long* out, int grid, cudaStream_t stream) {
  if ((mode & 0xff) == 2 && mode < 0x400) {
    if (n < 32 || n > 256 || (n & 15)) return fail(kBadArg, "mma_rate: bad N for cta_group::2");
    DV_CUDA_OK(cudaFuncSetAttribute(mma_rate_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    mma_rate_pair_kernel<<<grid & ~1, 128, region + 1024, stream>>>(n, n_mma, region, mode >> 8, out);
    DV_LAUNCH_OK();
    return kOk;
  }
  if (n < 16 || n > 256 || (n & 15) || region < 64 * 1024 || region > 200 * 1024) return fail(kBadArg, "mma_rate: bad arguments");
  DV_CUDA_OK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  mma_rate_kernel<<<grid, 128, region + 1024, stream>>>(n, n_mma, region, mode, out);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
