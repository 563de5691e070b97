// Debug probe (not on the product path): does TMA accept a tensor map whose dim-1 stride is smaller
// than the dim-0 extent (overlapping 128-byte windows)? Used to validate the space-to-depth stem.
#include "host_common.h"
#include "ptx.cuh"

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
