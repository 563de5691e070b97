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

// keys: [B][d] row-major -> queue: [d][K] columns ptr .. ptr+B-1. ptr_dev != NULL: the pointer is read from the
// model's own queue_ptr buffer (int64) - no host value in the launch, so a MoCo step replays as a CUDA graph.
__global__ void enqueue_kernel(const float* __restrict__ keys, float* __restrict__ queue, int B, int d, int K,
                               int ptr, const long long* __restrict__ ptr_dev) {
  __shared__ float tile[32][33];
  if (ptr_dev != nullptr) ptr = (int)(*ptr_dev % K);
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

int enqueue(const float* keys, float* queue, int B, int d, int K, int ptr, const long long* ptr_dev,
            cudaStream_t stream) {
  dim3 block(32, 8);
  dim3 grid(ceil_div(B, 32), ceil_div(d, 32));
  enqueue_kernel<<<grid, block, 0, stream>>>(keys, queue, B, d, K, ptr, ptr_dev);
  DV_LAUNCH_OK();
  return kOk;
}

// queue_ptr = (queue_ptr + batch) % K on the device (model/moco.py:352-353)
__global__ void advance_ptr_kernel(long long* ptr, int batch, int K) { *ptr = (*ptr + batch) % K; }

int advance_queue_ptr(long long* ptr, int batch, int K, cudaStream_t stream) {
  advance_ptr_kernel<<<1, 1, 0, stream>>>(ptr, batch, K);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv

// ---------------------------------------------------------------------------------------------------
// SGD with momentum and weight decay over ALL parameter tensors in one launch (reference: optim.SGD
// with one param-group per tensor, pretrain.py:262-272 — hundreds of tiny foreach launches).
// torch semantics (dampening 0, no nesterov): g = grad + wd*p ; buf = first ? g : mu*buf + g ; p -= lr*buf
// table: [n_chunks][4] int64 = (p pointer, grad pointer, momentum-buffer pointer, element count <= 8192)
namespace dv {

__global__ void __launch_bounds__(256)
sgd_momentum_kernel(const long long* __restrict__ table, float lr, float mu, float wd, int first) {
  const long long* e = table + (long long)blockIdx.x * 4;
  float* p = reinterpret_cast<float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  float* b = reinterpret_cast<float*>(e[2]);
  const int n = (int)e[3];
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  const int n4 = vec ? (n >> 2) : 0;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(b)[i];
    const float gx = fmaf(wd, pv.x, gv.x), gy = fmaf(wd, pv.y, gv.y), gz = fmaf(wd, pv.z, gv.z), gw = fmaf(wd, pv.w, gv.w);
    bv.x = first ? gx : fmaf(mu, bv.x, gx); bv.y = first ? gy : fmaf(mu, bv.y, gy);
    bv.z = first ? gz : fmaf(mu, bv.z, gz); bv.w = first ? gw : fmaf(mu, bv.w, gw);
    pv.x -= lr * bv.x; pv.y -= lr * bv.y; pv.z -= lr * bv.z; pv.w -= lr * bv.w;
    reinterpret_cast<float4*>(b)[i] = bv;
    reinterpret_cast<float4*>(p)[i] = pv;
  }
  for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
    const float gg = fmaf(wd, p[i], g[i]);
    const float bb = first ? gg : fmaf(mu, b[i], gg);
    b[i] = bb;
    p[i] -= lr * bb;
  }
}

int sgd_momentum_step(const long long* table, int n_chunks, float lr, float mu, float wd, int first,
                      cudaStream_t stream) {
  if (n_chunks <= 0) return kOk;
  sgd_momentum_kernel<<<n_chunks, 256, 0, stream>>>(table, lr, mu, wd, first);
  DV_LAUNCH_OK();
  return kOk;
}

}  // namespace dv
