// One-shot all-reduce (sum) of a small fp64 vector across the GPUs of one node over NVLink peer memory.
//
// Replaces the per-layer statistic exchange of cross-replica BatchNorm on the reference path
// (nn.SyncBatchNorm after nn.SyncBatchNorm.convert_sync_batchnorm, pretrain.py:244: an all_gather of
// mean/invstd/count per BatchNorm layer in forward and an all_reduce of sum_dy/sum_dy_xmu in backward).
// A pretraining step exchanges ~400 vectors of <= 2 KB..15 KB; each is on the critical path (the next layer needs the
// statistics), so the cost is pure latency: NCCL's ~20 us per call is 6-8 ms of an 80 ms step.
//
// Every rank owns a symmetric buffer (same layout on all ranks, mapped into every peer's address space):
//   slots [2 parities][world][kMaxElems] fp64, then flags [2][world] u64, then one u64 call counter.
// Call number `seq` (1, 2, ...; all ranks call in the same order) uses parity seq & 1. The host may number the calls
// itself (seq >= 1) or pass seq = 0: the kernel then takes the next number from the call counter in its own buffer
// (one CTA per call, calls of a channel run in stream order), which is what lets a step that contains these
// exchanges be captured ONCE as a CUDA graph and replayed (graph_step.py) - a host-side number would be baked in.
//   1. push: copy my vector into slot [parity][my rank] of EVERY rank (remote stores over NVLink),
//   2. fence.sys, then release-store flag [parity][my rank] = seq on every rank,
//   3. acquire-spin on my own flags [parity][q] == seq for all q,
//   4. sum my slots in rank order (bit-identical result on all ranks) back into the caller's vector.
// A rank can only enter call seq + 2 after every peer's flag of call seq + 1 arrived, i.e. after every peer left
// call seq - so two parities make slot reuse safe. One CTA. The spin has a time-out (dv_comm_set_timeout, default
// 600 s like NCCL's watchdog: ranks legitimately drift apart by many seconds around checkpoints and loader start-up);
// when it expires the kernel records (code, peer, call) in a host-mapped status word and returns - the CUDA context
// survives and the host raises from dv_comm_status() at its next exchange. With each vector every rank also pushes
// one scalar `tag` (the local per-channel count of cross-replica BatchNorm); ranks whose tags differ are reported the
// same way (nn.SyncBatchNorm weights ranks by their counts; this path requires equal per-rank batches).
#include <cuda_runtime.h>
#include <stdint.h>

#include "bn_math.cuh"
#include "host_common.h"

namespace dv {

constexpr int kArMaxWorld = 8;
constexpr int kArMaxElems = 4096;

struct ArPeers {
  unsigned long long base[kArMaxWorld];   // peer-mapped address of every rank's symmetric buffer
  unsigned long long timeout_ns;          // spin time-out
  int* status;                            // host-mapped [4]: code (0 ok, 1 time-out, 2 tag mismatch), peer, seq lo, seq hi
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void ar_report(const ArPeers& peers, int code, int peer, unsigned long long seq) {
  if (peers.status != nullptr && atomicCAS(peers.status, 0, code) == 0) {   // first error wins
    peers.status[1] = peer;
    peers.status[2] = (int)(seq & 0xffffffffull);
    peers.status[3] = (int)(seq >> 32);
    __threadfence_system();
  }
}

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Steps 1-3 for the calling CTA: push `src[0:n]` to every rank, signal, wait for every rank's signal. Afterwards
// ar_sum(i) returns the sum over ranks (in rank order) of element i.
__device__ __forceinline__ const double* ar_exchange(const double* __restrict__ src, int n, const ArPeers& peers,
                                                     int rank, int world, unsigned long long seq, double tag = 0.0) {
  const size_t slot_elems = (size_t)kArMaxElems;
  const size_t flags_off = 2ull * kArMaxWorld * slot_elems;   // in doubles (= u64 words)
  if (seq == 0ull) {     // device-side numbering (graph-replayable): next value of this channel's call counter
    __shared__ unsigned long long s_seq;
    if (threadIdx.x == 0) {
      unsigned long long* ctr = reinterpret_cast<unsigned long long*>(peers.base[rank]) + flags_off + 2ull * kArMaxWorld;
      s_seq = *ctr + 1ull;
      *ctr = s_seq;
    }
    __syncthreads();
    seq = s_seq;
  }
  const int parity = (int)(seq & 1ull);
  for (int p = 0; p < world; ++p) {
    double* dst = reinterpret_cast<double*>(peers.base[p]) + ((size_t)parity * kArMaxWorld + rank) * slot_elems;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) dst[kArMaxElems - 1] = tag;     // n < kArMaxElems: the last element of a slot is the tag
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    unsigned long long* remote = reinterpret_cast<unsigned long long*>(peers.base[threadIdx.x]) + flags_off +
                                 (size_t)parity * kArMaxWorld + rank;
    st_release_sys_u64(remote, seq);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers.base[rank]) + flags_off +
                                     (size_t)parity * kArMaxWorld + threadIdx.x;
    const unsigned long long t0 = global_timer_ns();
    int polls = 0;
    bool arrived = true;
    while (ld_acquire_sys_u64(mine) != seq) {
      if ((++polls & 1023) == 0 && global_timer_ns() - t0 > peers.timeout_ns) {
        // a peer never arrived: report to the host and give up on this call (its result is undefined);
        // the context stays usable so the host can raise, checkpoint or tear down in order
        ar_report(peers, 1, (int)threadIdx.x, seq);
        arrived = false;
        break;
      }
    }
    if (arrived) {
      const double* slots = reinterpret_cast<const double*>(peers.base[rank]) + (size_t)parity * kArMaxWorld * slot_elems;
      if (__ldcv(slots + (size_t)threadIdx.x * kArMaxElems + kArMaxElems - 1) != tag)
        ar_report(peers, 2, (int)threadIdx.x, seq);
    }
  }
  __syncthreads();
  return reinterpret_cast<const double*>(peers.base[rank]) + (size_t)parity * kArMaxWorld * slot_elems;
}
__device__ __forceinline__ double ar_sum(const double* base, int world, int i) {
  double s = 0.0;
  for (int q = 0; q < world; ++q) s += __ldcv(base + (size_t)q * kArMaxElems + i);
  return s;
}

__global__ void __launch_bounds__(512, 1)
small_allreduce_kernel(double* __restrict__ inout, int n, ArPeers peers, int rank, int world, unsigned long long seq) {
  const double* base = ar_exchange(inout, n, peers, rank, world, seq);
  for (int i = threadIdx.x; i < n; i += blockDim.x) inout[i] = ar_sum(base, world, i);
}

// Cross-replica BatchNorm forward: exchange of the local (sum, sumsq) + finalisation in one launch
// (nn.SyncBatchNorm forward: all_gather of the per-rank statistics, then mean/invstd/running-stat update).
__global__ void __launch_bounds__(512, 1)
bn_finalize_sync_kernel(const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                        float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ ss,
                        float* __restrict__ saved, int C, int Cp, double count_global, float eps, float momentum,
                        ArPeers peers, int rank, int world, unsigned long long seq) {
  const double* base = ar_exchange(stats, 2 * Cp, peers, rank, world, seq, count_global / world);
  for (int c = threadIdx.x; c < Cp; c += blockDim.x)
    bn_finalize_channel(c, ar_sum(base, world, c), ar_sum(base, world, Cp + c), gamma, beta, running_mean, running_var, ss,
                        saved, C, Cp, count_global, eps, momentum, 1);
}

// Cross-replica BatchNorm backward: exchange of (sum g, sum g*y) + dgamma/dbeta (local sums) + dy coefficients
// (global sums) in one launch (nn.SyncBatchNorm backward: all_reduce of sum_dy / sum_dy_xmu).
__global__ void __launch_bounds__(512, 1)
bn_bwd_finalize_sync_kernel(const double* __restrict__ sums_local, const float* __restrict__ gamma,
                            const float* __restrict__ saved, float* __restrict__ dgamma, float* __restrict__ dbeta,
                            float* __restrict__ coef, int C, int Cp, double count_global, float grad_beta,
                            ArPeers peers, int rank, int world, unsigned long long seq) {
  const double* base = ar_exchange(sums_local, 2 * Cp, peers, rank, world, seq, count_global / world);
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    const bool real = c < C;
    bn_bwd_finalize_channel(c, real ? sums_local[c] : 0.0, real ? sums_local[Cp + c] : 0.0,
                            real ? ar_sum(base, world, c) : 0.0, real ? ar_sum(base, world, Cp + c) : 0.0, gamma, saved,
                            dgamma, dbeta, coef, C, Cp, count_global, grad_beta);
  }
}

// ---- host state: time-out and the host-mapped status word
static double g_timeout_s = 600.0;
static int* g_status_host = nullptr;
static int* g_status_dev = nullptr;

static int ensure_status() {
  if (g_status_host != nullptr) return kOk;
  DV_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&g_status_host), 4 * sizeof(int), cudaHostAllocMapped));
  for (int i = 0; i < 4; ++i) g_status_host[i] = 0;
  DV_CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_status_dev), g_status_host, 0));
  return kOk;
}

void comm_set_timeout(double seconds) { g_timeout_s = seconds > 0.0 ? seconds : 600.0; }

// code: 0 ok, 1 = a peer did not arrive within the time-out, 2 = ranks exchanged different per-rank counts
int comm_status(int* peer, long long* seq, int clear) {
  if (g_status_host == nullptr) return 0;
  const int code = *reinterpret_cast<volatile int*>(g_status_host);
  if (code != 0) {
    if (peer) *peer = g_status_host[1];
    if (seq) *seq = (long long)(((unsigned long long)(unsigned)g_status_host[3] << 32) | (unsigned)g_status_host[2]);
    if (clear) for (int i = 0; i < 4; ++i) g_status_host[i] = 0;
  }
  return code;
}

static int fill_peers(ArPeers* peers, const long long* peer_ptrs, int world) {
  for (int i = 0; i < kArMaxWorld; ++i) peers->base[i] = i < world ? (unsigned long long)peer_ptrs[i] : 0ull;
  if (int rc = ensure_status()) return rc;
  peers->timeout_ns = (unsigned long long)(g_timeout_s * 1e9);
  peers->status = g_status_dev;
  return kOk;
}

long long small_allreduce_buffer_bytes() {
  return (long long)(2ll * kArMaxWorld * kArMaxElems + 2ll * kArMaxWorld + 1ll) * 8;   // slots, flags, call counter
}

int small_allreduce_f64(double* inout, int n, const long long* peer_ptrs, int rank, int world, long long seq,
                        cudaStream_t stream) {
  if (n <= 0 || n >= kArMaxElems) return fail(kBadArg, "small_allreduce: n must be in 1..%d", kArMaxElems - 1);
  if (world < 1 || world > kArMaxWorld || rank < 0 || rank >= world) return fail(kBadArg, "small_allreduce: bad rank/world");
  if (seq < 0) return fail(kBadArg, "small_allreduce: seq is 0 (device-side numbering) or starts at 1");
  ArPeers peers;
  if (int rc = fill_peers(&peers, peer_ptrs, world)) return rc;
  small_allreduce_kernel<<<1, 512, 0, stream>>>(inout, n, peers, rank, world, (unsigned long long)seq);
  DV_LAUNCH_OK();
  return kOk;
}

static int ar_peers(ArPeers* peers, int n, const long long* peer_ptrs, int rank, int world, long long seq) {
  if (n <= 0 || n >= kArMaxElems) return fail(kBadArg, "peer all-reduce: n must be in 1..%d", kArMaxElems - 1);
  if (world < 1 || world > kArMaxWorld || rank < 0 || rank >= world) return fail(kBadArg, "peer all-reduce: bad rank/world");
  if (seq < 0) return fail(kBadArg, "peer all-reduce: seq is 0 (device-side numbering) or starts at 1");
  return fill_peers(peers, peer_ptrs, world);
}

int bn_finalize_sync(const double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float* ss, float* saved, int C, int Cp, double count_global, float eps, float momentum,
                     const long long* peer_ptrs, int rank, int world, long long seq, cudaStream_t stream) {
  ArPeers peers;
  if (int rc = ar_peers(&peers, 2 * Cp, peer_ptrs, rank, world, seq)) return rc;
  bn_finalize_sync_kernel<<<1, 512, 0, stream>>>(stats, gamma, beta, running_mean, running_var, ss, saved, C, Cp,
                                                 count_global, eps, momentum, peers, rank, world,
                                                 (unsigned long long)seq);
  DV_LAUNCH_OK();
  return kOk;
}

int bn_bwd_finalize_sync(const double* sums_local, const float* gamma, const float* saved, float* dgamma, float* dbeta,
                         float* coef, int C, int Cp, double count_global, float grad_beta, const long long* peer_ptrs,
                         int rank, int world, long long seq, cudaStream_t stream) {
  ArPeers peers;
  if (int rc = ar_peers(&peers, 2 * Cp, peer_ptrs, rank, world, seq)) return rc;
  bn_bwd_finalize_sync_kernel<<<1, 512, 0, stream>>>(sums_local, gamma, saved, dgamma, dbeta, coef, C, Cp, count_global,
                                                     grad_beta, peers, rank, world, (unsigned long long)seq);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
