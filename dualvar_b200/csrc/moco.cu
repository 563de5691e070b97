// MoCo bookkeeping kernels.
//  - momentum update of the key encoder: theta_k = m*theta_k + (1-m)*theta_q over ALL parameter tensors
//    in ONE launch (reference: ~3 ATen kernels per tensor, model/moco.py:328-334)
//  - enqueue: queue[:, ptr:ptr+B] = keys^T (reference: strided transposed copy, model/moco.py:350-351)
#include "host_common.h"

namespace dv {

constexpr int kChunk = 8192;

// table: [n_chunks][3] int64 = (k pointer, q pointer, element count <= kChunk); pointers are chunk starts.
__global__ void __launch_bounds__(256)
momentum_update_kernel(const long long* __restrict__ table, float m) {
  const long long* e = table + (long long)blockIdx.x * 3;
  float* k = reinterpret_cast<float*>(e[0]);
  const float* q = reinterpret_cast<const float*>(e[1]);
  const int n = (int)e[2];
  const float om = 1.f - m;
  const bool vec = ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(q)) & 15) == 0;
  if (vec) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 a = reinterpret_cast<float4*>(k)[i];
      const float4 b = reinterpret_cast<const float4*>(q)[i];
      // same association as the reference: k*m + q*(1-m)
      a.x = a.x * m + b.x * om; a.y = a.y * m + b.y * om; a.z = a.z * m + b.z * om; a.w = a.w * m + b.w * om;
      reinterpret_cast<float4*>(k)[i] = a;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) k[i] = k[i] * m + q[i] * om;
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) k[i] = k[i] * m + q[i] * om;
  }
}

// keys: [B][d] row-major -> queue: [d][K] columns ptr .. ptr+B-1
__global__ void enqueue_kernel(const float* __restrict__ keys, float* __restrict__ queue, int B, int d, int K,
                               int ptr) {
  __shared__ float tile[32][33];
  const int b0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int b = b0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (b < B && c < d) ? keys[(long long)b * d + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, b = b0 + threadIdx.x;
    if (c < d && b < B) queue[(long long)c * K + ptr + b] = tile[threadIdx.x][j];
  }
}

int momentum_update(const long long* table, int n_chunks, float m, cudaStream_t stream) {
  if (n_chunks <= 0) return kOk;
  momentum_update_kernel<<<n_chunks, 256, 0, stream>>>(table, m);
  DV_LAUNCH_OK();
  return kOk;
}

int enqueue(const float* keys, float* queue, int B, int d, int K, int ptr, cudaStream_t stream) {
  dim3 block(32, 8);
  dim3 grid(ceil_div(B, 32), ceil_div(d, 32));
  enqueue_kernel<<<grid, block, 0, stream>>>(keys, queue, B, d, K, ptr);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
