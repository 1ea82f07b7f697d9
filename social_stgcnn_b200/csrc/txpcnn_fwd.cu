// Kernel (3): the TXP-CNN time extrapolator of reference model.py:187-195.
//
//   x  = view(h): (5,8,N) -> (8,5,N)      FLAT reinterpretation, not a transpose (model.py:187)
//   x  = PReLU0(conv3x3_0(x))             8 -> 12 "channels" (= time), plane (H=5 feature, W=N pedestrian)
//   x  = PReLUk(conv3x3_k(x)) + x         k = 1,2,3   (range(1, n_txpcnn-1), model.py:191-192)
//   x  = conv3x3_out(x)                   12 -> 12    (model.py:194)
//   out= view(x): (12,5,N) -> (5,12,N)    flat again  (model.py:195)
//
// Because both views are flat, the (5,8,N) st_gcn output block IS the (8,5,N) input
// as 40 rows of N, and the 60 output rows of N ARE the (5,12,N) result: no data movement.
//
// This kernel is FP32-FMA bound (56,160 FMA per pedestrian column against 400 bytes), so
// the design goal is FFMA issue density:
//   * a CTA owns a TILE of whole scenes, up to 128 pedestrian columns (one scene of up to
//     256 if it is larger); tiny real-data scenes share a tile.  In shared memory every
//     scene is separated by one permanently-zero column, which implements the conv's zero
//     padding in the pedestrian axis and stops neighbouring scenes bleeding into each
//     other without a single select in the inner loop.
//   * 256 threads = 2 output-channel halves x 128 columns.  A thread keeps the 6x5
//     accumulators of its column in registers; per input channel it loads the 5x3 input
//     patch (conflict-free LDS) and streams 54 weights as warp-uniform LDS.128 broadcasts:
//     234 FFMA per 29 shared loads.
//   * the five layers run in place in ONE 60-row activation tile (two barriers per layer),
//     which keeps the CTA at ~56 KB of shared memory and 4 CTAs (32 warps) per SM.
//   * persistent grid: weights are re-laid out into shared memory once per CTA.
#include "sstg_common.cuh"

namespace sstg {

namespace {

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

template <int JX>
__global__ void __launch_bounds__(X_THREADS)
txpcnn_fwd_kernel(const float* __restrict__ Hin, const int32_t* __restrict__ seg_off,
                  const int32_t* __restrict__ tile_off, int n_tiles, int Wp, const float* __restrict__ params,
                  float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  XSmem& sm = *reinterpret_cast<XSmem*>(smem_raw);
  float* act = reinterpret_cast<float*>(smem_raw + ((sizeof(XSmem) + 15) & ~size_t(15)));  // [60][Wp]
  const int tid = threadIdx.x;
  const int half = tid / X_COLS;  // warp-uniform
  const int qt = tid - half * X_COLS;

  // ---- weights: params [co][ci][3][3] -> shared [layer][ci][half][co%6][9]
  for (int i = tid; i < X_WTOTAL; i += X_THREADS) {
    int l, r;
    if (i < X_W0) {
      l = 0;
      r = i;
    } else {
      l = 1 + (i - X_W0) / X_WL;
      r = (i - X_W0) % X_WL;
    }
    const int ci = r / (2 * X_WSLOT), r2 = r % (2 * X_WSLOT);
    const int hf = r2 / X_WSLOT, r3 = r2 % X_WSLOT;
    float v = 0.f;
    if (r3 < 54) {
      const int co = hf * 6 + r3 / 9, tap = r3 % 9;
      const int cin = (l == 0) ? 8 : 12;
      const int base = (l == 0) ? P_TP0_W : (l == 4 ? P_TPOUT_W : P_TP1_W + (l - 1) * P_TP_STRIDE);
      v = params[base + (co * cin + ci) * 9 + tap];
    }
    sm.w[i] = v;
  }
  if (tid < 60) {
    const int l = tid / 12, co = tid % 12;
    const int base = (l == 0) ? P_TP0_B : (l == 4 ? P_TPOUT_B : P_TP1_W + (l - 1) * P_TP_STRIDE + 1296);
    sm.bias[l][co] = params[base + co];
  }
  if (tid < 4) sm.slope[tid] = params[P_PRELUS + tid];

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int s_lo = tile_off[tile], s_hi = tile_off[tile + 1];
    const int p_lo = seg_off[s_lo];
    const int cols = seg_off[s_hi] - p_lo;
    __syncthreads();  // previous tile fully written out / weights visible
    // ---- column tables + zero separator columns
    for (int q = tid; q < cols; q += X_THREADS) {
      int j = 0;
      while (seg_off[s_lo + j + 1] - p_lo <= q) ++j;
      const int sb = seg_off[s_lo + j] - p_lo;
      sm.scol[q] = q + j + 1;
      sm.nstr[q] = seg_off[s_lo + j + 1] - p_lo - sb;
      sm.goff[q] = sb * 40 + (q - sb);  // H block of scene j starts at sb*40 (relative), row stride N
    }
    const int nsc = s_hi - s_lo;
    for (int i = tid; i < (nsc + 1) * X_ROWS; i += X_THREADS) {
      const int j = i / X_ROWS, row = i - j * X_ROWS;
      const int col = (j < nsc) ? (seg_off[s_lo + j] - p_lo + j) : (cols + nsc);
      act[row * Wp + col] = 0.f;
    }
    __syncthreads();
    // ---- load the 40 input rows of every scene of the tile (coalesced along the pedestrian axis)
    const float* Hb = Hin + static_cast<size_t>(p_lo) * 40;
    for (int i = tid; i < 40 * cols; i += X_THREADS) {
      const int row = i / cols, q = i - row * cols;
      act[row * Wp + sm.scol[q]] = Hb[sm.goff[q] + row * sm.nstr[q]];
    }
    __syncthreads();

    int sc[JX];
    bool on[JX];
#pragma unroll
    for (int u = 0; u < JX; ++u) {
      const int q = qt + u * X_COLS;
      on[u] = q < cols;
      sc[u] = on[u] ? sm.scol[q] : 1;
    }

#pragma unroll 1
    for (int l = 0; l < 5; ++l) {
      float acc[JX][6][5];
      const float* wl = sm.w + (l == 0 ? 0 : X_W0 + (l - 1) * X_WL);
#pragma unroll
      for (int u = 0; u < JX; ++u) {
#pragma unroll
        for (int co = 0; co < 6; ++co) {
          const float b = sm.bias[l][half * 6 + co];
#pragma unroll
          for (int h = 0; h < 5; ++h) acc[u][co][h] = b;
        }
        if (on[u]) {
          if (l == 0)
            conv_column<8>(act, Wp, sc[u], wl, half, acc[u]);
          else
            conv_column<12>(act, Wp, sc[u], wl, half, acc[u]);
        }
      }
      if (l < 4) {
        const float sl = sm.slope[l];
#pragma unroll
        for (int u = 0; u < JX; ++u)
#pragma unroll
          for (int co = 0; co < 6; ++co)
#pragma unroll
            for (int h = 0; h < 5; ++h) {
              float v = prelu(acc[u][co][h], sl);
              if (l > 0) v += act[((half * 6 + co) * 5 + h) * Wp + sc[u]];  // residual (model.py:192)
              acc[u][co][h] = v;
            }
        __syncthreads();  // everyone has read the old activations
#pragma unroll
        for (int u = 0; u < JX; ++u)
          if (on[u])
#pragma unroll
            for (int co = 0; co < 6; ++co)
#pragma unroll
              for (int h = 0; h < 5; ++h) act[((half * 6 + co) * 5 + h) * Wp + sc[u]] = acc[u][co][h];
        __syncthreads();
      } else {
        float* Ob = out + static_cast<size_t>(p_lo) * 60;
#pragma unroll
        for (int u = 0; u < JX; ++u)
          if (on[u]) {
            const int q = qt + u * X_COLS;
            const int N = sm.nstr[q];
            const int j = sm.scol[q] - 1 - q;               // scene ordinal within the tile
            const int sb = seg_off[s_lo + j] - p_lo;        // first column of that scene
            float* o = Ob + static_cast<size_t>(sb) * 60 + (q - sb);
#pragma unroll
            for (int co = 0; co < 6; ++co)
#pragma unroll
              for (int h = 0; h < 5; ++h) o[((half * 6 + co) * 5 + h) * N] = acc[u][co][h];
          }
      }
    }
  }
}

}  // namespace

int launch_txpcnn_forward(const float* H, const int32_t* seg_off, const int32_t* tile_off, int n_tiles, int max_n,
                          const float* params, float* out, cudaStream_t st) {
  if (n_tiles == 0) return SSTG_OK;
  // widest tile: up to X_COLS columns of up to MAX_SCENES scenes, or one big scene;
  // one zero separator column before every scene and one after the last
  int Wp = X_COLS + SSTG_XTILE_MAX_SCENES + 1;
  if (max_n + 2 > Wp) Wp = max_n + 2;
  Wp = (Wp + 3) & ~3;
  const size_t smem = ((sizeof(XSmem) + 15) & ~size_t(15)) + sizeof(float) * X_ROWS * Wp;
  int dev = 0, sms = 148;
  SSTG_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
  SSTG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "sm count");
  auto launch = [&](auto kern) -> int {
    SSTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                  "txpcnn smem attr");
    int per_sm = 1;
    SSTG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, X_THREADS, smem), "occupancy");
    if (per_sm < 1) per_sm = 1;
    const long long want = static_cast<long long>(sms) * per_sm;
    const int grid = static_cast<int>(n_tiles < want ? n_tiles : want);
    kern<<<grid, X_THREADS, smem, st>>>(H, seg_off, tile_off, n_tiles, Wp, params, out);
    SSTG_CUDA_TRY(cudaGetLastError(), "txpcnn launch");
    return SSTG_OK;
  };
  if (max_n > X_COLS) return launch(txpcnn_fwd_kernel<2>);
  return launch(txpcnn_fwd_kernel<1>);
}

}  // namespace sstg
