// Pieces shared by the TXP-CNN forward (txpcnn_fwd.cu) and backward (txpcnn_bwd.cu) kernels:
// tile geometry, the shared-memory header, and the register-tiled 3x3 column convolution.
#pragma once
#include "sstg_common.cuh"

namespace sstg {
namespace txp {

constexpr int X_COLS = SSTG_XTILE_COLS;
constexpr int X_THREADS = 2 * X_COLS;
constexpr int X_ROWS = T_PRED * C;  // 60 rows of the activation tile (row = channel*5 + h)
constexpr int X_WSLOT = 56;         // 6 output channels x 9 taps, padded to a float4 multiple
constexpr int X_W0 = 8 * 2 * X_WSLOT;
constexpr int X_WL = 12 * 2 * X_WSLOT;
constexpr int X_WTOTAL = X_W0 + 4 * X_WL;  // 6272 floats

struct XSmem {
  float w[X_WTOTAL];
  float bias[5][12];
  float slope[4];
  int scol[2 * X_COLS];  // tile column -> shared-memory column
  int goff[2 * X_COLS];  // tile column -> float offset of (row 0, column) relative to the tile's first scene
  int nstr[2 * X_COLS];  // tile column -> N of its scene (row stride in global memory)
};

// ---- packed FP32 (sm_100 FFMA2): d.xy = a.xy * b.xy + d.xy, two IEEE fp32 FMAs in one issue slot
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

// One 3x3 layer for one column with packed FMAs: acc[pr][h] holds output channels
// (half*6 + 2 pr, half*6 + 2 pr + 1) at feature row h.  Weight slot layout (forward only):
// wl[(ci*2 + half)*56 + tap*6 + co6] -- tap-major, so an LDS.128 delivers aligned channel pairs.
// CI input channels starting at ci0 (ci0 > 0: a partial sum over a slice of the input channels).
template <int CI>
__device__ __forceinline__ void conv_column2(const float* __restrict__ act, int Wp, int sc,
                                             const float* __restrict__ wl, int half, float2 (&acc)[3][5],
                                             int ci0 = 0) {
#pragma unroll 2
  for (int ci = ci0; ci < ci0 + CI; ++ci) {
    float2 x[5][3];
    const float* a = act + (ci * 5) * Wp + sc - 1;
#pragma unroll
    for (int h = 0; h < 5; ++h)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float v = a[h * Wp + kw];
        x[h][kw] = make_float2(v, v);
      }
    const float4* w4 = reinterpret_cast<const float4*>(wl + (ci * 2 + half) * X_WSLOT);
    float4 w[X_WSLOT / 4];
#pragma unroll
    for (int i = 0; i < X_WSLOT / 4; ++i) w[i] = w4[i];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
          const int f = (kh * 3 + kw) * 6 + 2 * pr;  // even: an aligned pair inside one float4
          const float4 q = w[f >> 2];
          const float2 wp = (f & 2) ? make_float2(q.z, q.w) : make_float2(q.x, q.y);
#pragma unroll
          for (int h = 0; h < 5; ++h) {
            const int hi = h + kh - 1;
            if (hi >= 0 && hi < 5) ffma2(acc[pr][h], wp, x[hi][kw]);
          }
        }
  }
}

__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

}  // namespace txp
}  // namespace sstg
