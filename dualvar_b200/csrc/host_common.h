// Host-side plumbing shared by every C-ABI entry point: status codes, last-error string,
// CUDA error checks and the TMA tensor-map encoder (resolved from the driver at run time so the
// library links against cudart only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace dv {

enum Status : int { kOk = 0, kBadArg = 1, kCudaError = 2, kUnsupported = 3 };

void set_last_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define DV_CUDA_OK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return ::dv::fail(::dv::kCudaError, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                  \
  } while (0)

#define DV_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return ::dv::fail(::dv::kBadArg, __VA_ARGS__); \
  } while (0)

long long launch_counter();
void count_launch();
#define DV_LAUNCH_OK()                  \
  do {                                  \
    ::dv::count_launch();               \
    DV_CUDA_OK(cudaGetLastError());     \
  } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int ceil_div(int x, int m) { return (x + m - 1) / m; }
inline long long ceil_div_ll(long long x, long long m) { return (x + m - 1) / m; }

int sm_count();

// Encode a tiled tensor map. dims/strides are innermost-first; strides_bytes[0] is implied
// (element size) and ignored, strides_bytes[i>0] must be multiples of 16.
// swizzle128: 128-byte swizzle (inner box must be <= 128 B) else no swizzle.
int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128,
                bool is_float32 = false);

}  // namespace dv
