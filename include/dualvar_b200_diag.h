/* Diagnostics-only entry points of libdualvar_b200_diag.so (tests/diag/*.py microbenchmarks and role counters).
 * That library is the product library rebuilt with -DDV_DIAG (conv kernels carry per-role cycle counters) plus
 * dualvar_b200/csrc/diag/*.cu; it exports everything include/dualvar_b200.h declares and the functions below.
 * The product library (libdualvar_b200.so) contains none of this. Select it with DV_LIB_PATH. */
#ifndef DUALVAR_B200_DIAG_H
#define DUALVAR_B200_DIAG_H
#include "dualvar_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* debug (tests/diag only): per-CTA role cycle counters of the next conv_tile_kernel launches are written to
 * buf [148][16] (producer total/wait, MMA total/wait-data/wait-accumulator, epilogue total/wait, tiles, epilogue phases); NULL = off */
int dv_debug_set_conv_profile(int64_t* buf);
/* debug probe (tests only): TMA tensor map with overlapping windows */
int dv_debug_probe_overlap_tmap(const void* src, void* out, int c1, void* stream);
/* debug microbenchmark (tests/diag/mma_rate.py): cycles for n_mma back-to-back tcgen05.mma M=128 N=n K=16 (bf16,
 * shared-memory operands cycling through region_bytes) per CTA -> cycles[grid][2] = (issue, completion) */
int dv_debug_mma_rate(int n, int n_mma, int region_bytes, int mode, int64_t* cycles, int grid, void* stream);

#ifdef __cplusplus
}
#endif
#endif
