// BatchNorm running-statistics recurrence (torch.nn.BatchNorm2d, model.py:114,121,141), shared by
// bn_running_update_kernel / train_tail_kernel (train_kernels.cu) and by the extra CTA of the TXP-CNN
// training forward (txpcnn_fwd.cu), which runs it underneath the convolutions.
#pragma once
#include "sstg_common.cuh"

namespace sstg {

constexpr int BN_TILE = 128;  // scenes staged per pass (128 x 30 doubles = 30 KB of static shared memory)
__device__ __forceinline__ void bn_running_update_body(const float* __restrict__ stats,
                                                       const int32_t* __restrict__ seg_off, int S, double momentum,
                                                       float* __restrict__ bn) {
  // The recurrence is sequential in the scene index, but its inputs are not: 256 threads stage a
  // tile of per-scene statistics (and element counts) in shared memory with coalesced loads, then
  // 30 threads run the recurrence over the tile out of shared memory.
  __shared__ double st[BN_TILE * SSTG_N_BN];  // momentum * (mean | unbiased variance), ready for the recurrence
  __shared__ double fac[BN_TILE];             // momentum * M / (M - 1) of each staged scene
  const int i = threadIdx.x;
  double r = i < SSTG_N_BN ? static_cast<double>(bn[i]) : 0.0;
  for (int s0 = 0; s0 < S; s0 += BN_TILE) {
    const int ns = (S - s0 < BN_TILE) ? (S - s0) : BN_TILE;
    // the unbiased-variance factor M / (M - 1): ONE fp64 division per scene, by one thread per scene (the
    // division is the expensive part; it is neither inside the sequential recurrence nor repeated per channel)
    if (i < ns) {
      const double M = static_cast<double>(SSTG_T_OBS) * (seg_off[s0 + i + 1] - seg_off[s0 + i]);
      fac[i] = momentum * (M / (M - 1.0));
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ns * SSTG_N_BN; j += blockDim.x) {
      const int s = j / SSTG_N_BN, k = j - s * SSTG_N_BN;
      const double v = stats[static_cast<size_t>(s0) * SSTG_N_BN + j];
      st[j] = ((k / 5) & 1) ? fac[s] * v : momentum * v;
    }
    __syncthreads();
    if (i < SSTG_N_BN) {
      const double keep = 1.0 - momentum;
      for (int s = 0; s < ns; ++s)
        r = static_cast<double>(static_cast<float>(st[s * SSTG_N_BN + i] + keep * r));
    }
    __syncthreads();
  }
  if (i < SSTG_N_BN) bn[i] = static_cast<float>(r);
}

}  // namespace sstg
