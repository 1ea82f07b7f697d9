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

// one 3x3 layer for one column: acc[6][5] += W[half] * patch
template <int CI>
__device__ __forceinline__ void conv_column(const float* __restrict__ act, int Wp, int sc,
                                            const float* __restrict__ wl, int half, float (&acc)[6][5]) {
#pragma unroll 2
  for (int ci = 0; ci < CI; ++ci) {
    float x[5][3];
    const float* a = act + (ci * 5) * Wp + sc - 1;
#pragma unroll
    for (int h = 0; h < 5; ++h) {
      x[h][0] = a[h * Wp + 0];
      x[h][1] = a[h * Wp + 1];
      x[h][2] = a[h * Wp + 2];
    }
    const float4* w4 = reinterpret_cast<const float4*>(wl + (ci * 2 + half) * X_WSLOT);
    float w[X_WSLOT];
#pragma unroll
    for (int i = 0; i < X_WSLOT / 4; ++i) {
      const float4 v = w4[i];
      w[4 * i + 0] = v.x;
      w[4 * i + 1] = v.y;
      w[4 * i + 2] = v.z;
      w[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int co = 0; co < 6; ++co)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float wv = w[co * 9 + kh * 3 + kw];
#pragma unroll
          for (int h = 0; h < 5; ++h) {
            const int hi = h + kh - 1;
            if (hi >= 0 && hi < 5) acc[co][h] = fmaf(wv, x[hi][kw], acc[co][h]);
          }
        }
  }
}

}  // namespace txp
}  // namespace sstg
