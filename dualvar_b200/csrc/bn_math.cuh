// Per-channel BatchNorm finalisation shared by bn.cu (single GPU / NCCL path) and comm.cu (fused with the NVLink
// peer all-reduce of the statistics). Reference: nn.BatchNorm3d / nn.SyncBatchNorm forward statistics and backward
// (backbone/r21d.py:56,106,111 after convert_sync_batchnorm, pretrain.py:244).
#pragma once
#include <cuda_runtime.h>

namespace dv {

// sum / sumsq over `count` values of channel c -> scale/shift (ss), saved mean/invstd, running statistics
__device__ __forceinline__ void bn_finalize_channel(int c, double sum, double sumsq, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float* __restrict__ running_mean,
                                                    float* __restrict__ running_var, float* __restrict__ ss,
                                                    float* __restrict__ saved, int C, int Cp, double count, float eps,
                                                    float momentum, int training) {
  if (c >= C) {
    ss[c] = 0.f; ss[Cp + c] = 0.f;
    if (saved) { saved[c] = 0.f; saved[Cp + c] = 0.f; }
    return;
  }
  float mean, invstd;
  if (training) {
    const double m = sum / count;
    double var = sumsq / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    mean = running_mean[c];
    invstd = rsqrtf(running_var[c] + eps);
  }
  const float sc = gamma[c] * invstd;
  ss[c] = sc;
  ss[Cp + c] = beta[c] - mean * sc;
  if (saved) { saved[c] = mean; saved[Cp + c] = invstd; }
}

// local sums -> dgamma/dbeta of this rank; global sums (+ global count) -> coefficients of dy = A*g + B*y + C
__device__ __forceinline__ void bn_bwd_finalize_channel(int c, double sg_l, double sgy_l, double sg, double sgy,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ saved, float* __restrict__ dgamma,
                                                        float* __restrict__ dbeta, float* __restrict__ coef, int C,
                                                        int Cp, double count_global, float grad_beta) {
  if (c >= C) {
    coef[c] = 0.f; coef[Cp + c] = 0.f; coef[2 * Cp + c] = 0.f;
    return;
  }
  const double mean = saved[c], invstd = saved[Cp + c];
  const double dg_l = (sgy_l - mean * sg_l) * invstd;
  if (dgamma) dgamma[c] = (grad_beta != 0.f ? grad_beta * dgamma[c] : 0.f) + (float)dg_l;
  if (dbeta) dbeta[c] = (grad_beta != 0.f ? grad_beta * dbeta[c] : 0.f) + (float)sg_l;
  const double dg = (sgy - mean * sg) * invstd;
  const double A = (double)gamma[c] * invstd;
  const double B = -A * invstd * dg / count_global;
  const double Cc = A * (-sg / count_global + mean * invstd * dg / count_global);
  coef[c] = (float)A;
  coef[Cp + c] = (float)B;
  coef[2 * Cp + c] = (float)Cc;
}

}  // namespace dv
