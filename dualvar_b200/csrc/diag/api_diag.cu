// extern "C" entry points of the diagnostics library (include/dualvar_b200_diag.h); built only into
// libdualvar_b200_diag.so (-DDV_DIAG), never into the product library.
#include "../../../include/dualvar_b200_diag.h"

#include <cuda_runtime.h>

namespace dv {
void set_conv_profile(long long* p);
int probe_overlap(const void* src, void* out, int c1, cudaStream_t stream);
int mma_rate(int n, int n_mma, int region, int mode, long long* out, int grid, cudaStream_t stream);
}  // namespace dv

extern "C" {

int dv_debug_set_conv_profile(int64_t* buf) {
  dv::set_conv_profile(reinterpret_cast<long long*>(buf));
  return 0;
}

int dv_debug_probe_overlap_tmap(const void* src, void* out, int c1, void* stream) {
  return dv::probe_overlap(src, out, c1, static_cast<cudaStream_t>(stream));
}

int dv_debug_mma_rate(int n, int n_mma, int region_bytes, int mode, int64_t* cycles, int grid, void* stream) {
  return dv::mma_rate(n, n_mma, region_bytes, mode, reinterpret_cast<long long*>(cycles), grid,
                      static_cast<cudaStream_t>(stream));
}

}  // extern "C"
