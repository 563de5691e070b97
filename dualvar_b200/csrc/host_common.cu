#include "host_common.h"

#include <cudaTypedefs.h>
#include <stdarg.h>

#include <atomic>
#include <mutex>

namespace dv {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error_ref() { return g_last_error; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

static std::atomic<long long> g_launches{0};
long long launch_counter() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}

static PFN_cuTensorMapEncodeTiled_v12000 resolve_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128,
                bool is_float32) {
  auto fn = resolve_encode();
  if (!fn) return fail(kCudaError, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i];
      if (strides_bytes[i] % 16 != 0)
        return fail(kBadArg, "tensor map stride %d = %llu bytes is not a multiple of 16", i,
                    (unsigned long long)strides_bytes[i]);
    }
    if (box[i] == 0 || box[i] > 256) return fail(kBadArg, "tensor map box[%d]=%u out of range", i, box[i]);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return fail(kBadArg, "tensor map base %p is not 16-byte aligned", base);
  CUtensorMapDataType dt = is_float32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                      : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                         : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  CUresult r = fn(out, dt, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string s = "cuTensorMapEncodeTiled failed code " + std::to_string((int)r) + " rank " +
                    std::to_string(rank) + " dims";
    for (int i = 0; i < rank; ++i) s += " " + std::to_string(dims[i]);
    s += " box";
    for (int i = 0; i < rank; ++i) s += " " + std::to_string(box[i]);
    s += " strides";
    for (int i = 1; i < rank; ++i) s += " " + std::to_string(strides_bytes[i]);
    set_last_error(s);
    return kCudaError;
  }
  return kOk;
}

}  // namespace dv
