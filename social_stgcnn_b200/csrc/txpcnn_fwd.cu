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
// the design goal is FMA-pipe utilisation:
//   * a CTA owns a TILE of whole scenes, up to 128 pedestrian columns (one scene of up to
//     256 if it is larger); tiny real-data scenes share a tile.  In shared memory every
//     scene is separated by one permanently-zero column, which implements the conv's zero
//     padding in the pedestrian axis and stops neighbouring scenes bleeding into each
//     other without a single select in the inner loop.
//   * 256 threads = 2 output-channel halves x 128 columns.  A thread keeps the 6x5
//     accumulators of its column as 3x5 packed float2 (channel pairs) and issues
//     fma.rn.f32x2 (SASS FFMA2: two fp32 FMAs per issue slot); per input channel it loads the
//     5x3 input patch (conflict-free LDS, duplicated into pairs) and streams 54 tap-major
//     weights as warp-uniform LDS.128 broadcasts: 117 FFMA2 per 29 shared loads.
//   * the tile's 40 input rows arrive by 20 cp.async per thread (one wait), and the five layers
//     run in place in ONE 60-row activation tile (two barriers per layer): ~68 KB of shared
//     memory and 128 registers -> 2 CTAs (16 warps) per SM.
//   * persistent grid: weights are re-laid out into shared memory once per CTA.
// Inference on scenes of up to 128 pedestrians uses the Winograd kernel of txpcnn_wino.cu instead;
// this direct kernel serves training (SAVE: it also stores every layer's pre-activation and
// output for the backward pass) and larger scenes.
#include "txpcnn_common.cuh"
#include "bn_update.cuh"

namespace sstg {

namespace {

using namespace txp;

// -DSSTG_PHASE_CLOCKS: thread 0 of every CTA stamps clock64() at the phase boundaries of its first tile
// (profiling aid for the latency-bound small-group case; read back with sstg_debug_phase_clocks_xf).
#ifdef SSTG_PHASE_CLOCKS
__device__ long long g_phase_xf[160 * 16];
#define PHASE_XF(k)                                                                                      \
  do {                                                                                                   \
    if (threadIdx.x == 0 && tile == cta) g_phase_xf[cta * 16 + (k)] = clock64(); \
  } while (0)
#else
#define PHASE_XF(k)
#endif

// SAVE (training): also store every layer's pre-activation u_l = conv_l(x_l) and its output
// x_{l+1} (l = 0..3) for the backward pass, in the same (12,5,N)-rows-per-scene layout as
// `out`, layer l at float offset l * layer_stride.
template <int JX, bool SAVE>
__global__ void __launch_bounds__(X_THREADS)
txpcnn_fwd_kernel(const float* __restrict__ Hin, const int32_t* __restrict__ seg_off,
                  const int32_t* __restrict__ tile_off, int n_tiles, int Wp, const float* __restrict__ params,
                  float* __restrict__ out, float* __restrict__ Usave, float* __restrict__ Xsave,
                  long long layer_stride, const float* __restrict__ bn_stats, int bn_scenes, double bn_momentum,
                  float* __restrict__ bn) {
  // Training only: when `bn` is given, CTA 0 of the grid is not a convolution CTA -- it runs the BatchNorm
  // running-statistics recurrence (sequential over the scenes, ~10 us for a 128-scene group) underneath the
  // convolutions instead of on the critical path at the end of the step.
  const int bn_ctas = (SAVE && bn != nullptr) ? 1 : 0;
  const int conv_ctas = static_cast<int>(gridDim.x) - bn_ctas;
  const int cta = static_cast<int>(blockIdx.x) - bn_ctas;   // index among the convolution CTAs
  if constexpr (SAVE) {
    if (bn != nullptr && blockIdx.x == 0) {
      pdl_wait();  // the statistics come from the st_gcn forward
      pdl_release();
      bn_running_update_body(bn_stats, seg_off, bn_scenes, bn_momentum, bn);
      return;
    }
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  XSmem& sm = *reinterpret_cast<XSmem*>(smem_raw);
  float* act = reinterpret_cast<float*>(smem_raw + ((sizeof(XSmem) + 15) & ~size_t(15)));  // [60][Wp]
  const int tid = threadIdx.x;
  const int half = tid / X_COLS;  // warp-uniform
  const int idx = tid - half * X_COLS;
#ifdef SSTG_PHASE_CLOCKS
  if (tid == 0 && cta < 160) g_phase_xf[cta * 16 + 0] = clock64();
#endif

  // ---- weights: params [co][ci][3][3] -> shared [layer][ci][half][tap][co%6]  (tap-major for FFMA2 pairs);
  //      the loads of a batch are all issued before its first store (one L2 round trip per batch)
  constexpr int W_BATCH = 5;
  for (int i0 = tid; i0 < X_WTOTAL; i0 += W_BATCH * X_THREADS) {
    float v[W_BATCH];
#pragma unroll
    for (int j = 0; j < W_BATCH; ++j) {
      const int i = i0 + j * X_THREADS;
      v[j] = 0.f;
      if (i < X_WTOTAL) {
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
        if (r3 < 54) {
          const int co = hf * 6 + r3 % 6, tap = r3 / 6;
          const int cin = (l == 0) ? 8 : 12;
          const int base = (l == 0) ? P_TP0_W : (l == 4 ? P_TPOUT_W : P_TP1_W + (l - 1) * P_TP_STRIDE);
          v[j] = __ldg(params + base + (co * cin + ci) * 9 + tap);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < W_BATCH; ++j) {
      const int i = i0 + j * X_THREADS;
      if (i < X_WTOTAL) sm.w[i] = v[j];
    }
  }
  if (tid < 60) {
    const int l = tid / 12, co = tid % 12;
    const int base = (l == 0) ? P_TP0_B : (l == 4 ? P_TPOUT_B : P_TP1_W + (l - 1) * P_TP_STRIDE + 1296);
    sm.bias[l][co] = params[base + co];
  }
  if (tid < 4) sm.slope[tid] = params[P_PRELUS + tid];
  pdl_wait();  // everything above read only `params`
  pdl_release();

  for (int tile = cta; tile < n_tiles; tile += conv_ctas) {
    PHASE_XF(1);
    const int s_lo = tile_off[tile], s_hi = tile_off[tile + 1];
    const int p_lo = seg_off[s_lo];
    const int cols = seg_off[s_hi] - p_lo;
    // Training tiles are small (one ~10-pedestrian scene per CTA) and the step is latency-bound: a tile of
    // <= 64 columns spreads each column's convolution over 2 or 4 lanes (slices of the input channels,
    // butterfly-summed so that all lanes of a column hold the same bits).  CTA-uniform.
    const int cgs = (!SAVE || JX != 1) ? 1 : (cols <= X_COLS / 4 ? 4 : (cols <= X_COLS / 2 ? 2 : 1));
    const int qt = cgs == 4 ? (idx >> 2) : (cgs == 2 ? (idx >> 1) : idx);
    const int cg = idx & (cgs - 1);
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
    // thread = (column, row parity): 20 independent 4-byte cp.async in flight per thread, no index division
    const float* Hb = Hin + static_cast<size_t>(p_lo) * 40;
    for (int q = tid & (X_COLS - 1); q < cols; q += X_COLS) {
      const float* src = Hb + sm.goff[q];
      float* dst = act + sm.scol[q];
      const int ns = sm.nstr[q];
#pragma unroll
      for (int row = tid >> 7; row < 40; row += X_THREADS / X_COLS) cp_async4(dst + row * Wp, src + row * ns);
    }
    cp_async_wait_all();
    __syncthreads();
    PHASE_XF(2);

    int sc[JX], gN[JX];
    long long gO[JX];  // float offset of (row 0, this column) in a 60-rows-per-scene buffer
    bool on[JX];
#pragma unroll
    for (int u = 0; u < JX; ++u) {
      const int q = qt + u * X_COLS;
      on[u] = q < cols;
      sc[u] = on[u] ? sm.scol[q] : 1;
      gN[u] = on[u] ? sm.nstr[q] : 1;
      const int j = on[u] ? sm.scol[q] - 1 - q : 0;          // scene ordinal within the tile
      const int sb = seg_off[s_lo + j] - p_lo;               // first column of that scene
      gO[u] = static_cast<long long>(p_lo + sb) * 60 + (q - sb);
    }

#pragma unroll 1
    for (int l = 0; l < 5; ++l) {
      float2 acc[JX][3][5];  // [column][channel pair][feature row]
      const float* wl = sm.w + (l == 0 ? 0 : X_W0 + (l - 1) * X_WL);
#pragma unroll
      for (int u = 0; u < JX; ++u) {
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
          float2 b = make_float2(sm.bias[l][half * 6 + 2 * pr], sm.bias[l][half * 6 + 2 * pr + 1]);
          if (cgs != 1) b = make_float2(0.f, 0.f);  // split columns: the owner of an output adds the bias
#pragma unroll
          for (int h = 0; h < 5; ++h) acc[u][pr][h] = b;
        }
        if (!SAVE || cgs == 1) {
          if (on[u]) {
            if (l == 0)
              conv_column2<8>(act, Wp, sc[u], wl, half, acc[u]);
            else
              conv_column2<12>(act, Wp, sc[u], wl, half, acc[u]);
          }
        } else {
          if (on[u]) {
            if (cgs == 4) {
              if (l == 0)
                conv_column2<2>(act, Wp, sc[u], wl, half, acc[u], cg * 2);
              else
                conv_column2<3>(act, Wp, sc[u], wl, half, acc[u], cg * 3);
            } else {
              if (l == 0)
                conv_column2<4>(act, Wp, sc[u], wl, half, acc[u], cg * 4);
              else
                conv_column2<6>(act, Wp, sc[u], wl, half, acc[u], cg * 6);
            }
          }
        }
      }
      if (SAVE && JX == 1 && cgs != 1) {
        // ---- split column: the 2 / 4 lanes of a column reduce-scatter their partial sums (a lane hands over the
        // half its partner keeps), so each lane ends up OWNING 8 / 4 of the column's 15 (channel pair, row) outputs
        // and does the activation, residual, saves and write-back for those only.
        float2 f[16];
#pragma unroll
        for (int pr = 0; pr < 3; ++pr)
#pragma unroll
          for (int h = 0; h < 5; ++h) f[pr * 5 + h] = acc[0][pr][h];
        f[15] = make_float2(0.f, 0.f);
        const bool up1 = (cg & 1) != 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 send = up1 ? f[j] : f[j + 8], keep = up1 ? f[j + 8] : f[j];
          f[j].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 1);
          f[j].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 1);
        }
        int base = up1 ? 8 : 0, nown = 8;
        if (cgs == 4) {
          const bool up2 = (cg & 2) != 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 send = up2 ? f[j] : f[j + 4], keep = up2 ? f[j + 4] : f[j];
            f[j].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 2);
            f[j].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 2);
          }
          base += up2 ? 4 : 0;
          nown = 4;
        }
        PHASE_XF(3 + 2 * l);
        const float sl = l < 4 ? sm.slope[l] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int sidx = base + j;
          if (j < nown && sidx < 15 && on[0]) {
            const int pr = sidx / 5, h = sidx - pr * 5;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int co = half * 6 + 2 * pr + e, row = co * 5 + h;
              float& a = e ? f[j].y : f[j].x;
              a += sm.bias[l][co];
              if (l < 4) {
                float v = prelu(a, sl);
                if (l > 0) v += act[row * Wp + sc[0]];  // residual (model.py:192)
                const long long g = l * layer_stride + gO[0] + static_cast<long long>(row) * gN[0];
                Usave[g] = a;
                Xsave[g] = v;
                a = v;
              } else {
                out[gO[0] + static_cast<long long>(row) * gN[0]] = a;
              }
            }
          }
        }
        if (l < 4) {
          __syncthreads();  // everyone has read the old activations
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int sidx = base + j;
            if (j < nown && sidx < 15 && on[0]) {
              const int pr = sidx / 5, h = sidx - pr * 5;
              act[((half * 6 + 2 * pr) * 5 + h) * Wp + sc[0]] = f[j].x;
              act[((half * 6 + 2 * pr + 1) * 5 + h) * Wp + sc[0]] = f[j].y;
            }
          }
          __syncthreads();
          PHASE_XF(4 + 2 * l);
        } else {
          PHASE_XF(12);
        }
        continue;
      }
      PHASE_XF(3 + 2 * l);
      if (l < 4) {
        const float sl = sm.slope[l];
#pragma unroll
        for (int u = 0; u < JX; ++u)
#pragma unroll
          for (int pr = 0; pr < 3; ++pr)
#pragma unroll
            for (int h = 0; h < 5; ++h)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                float& a = e ? acc[u][pr][h].y : acc[u][pr][h].x;
                const int row = (half * 6 + 2 * pr + e) * 5 + h;
                float v = prelu(a, sl);
                if (l > 0) v += act[row * Wp + sc[u]];  // residual (model.py:192)
                if (SAVE && on[u] && cg == 0) {
                  const long long g = l * layer_stride + gO[u] + static_cast<long long>(row) * gN[u];
                  Usave[g] = a;
                  Xsave[g] = v;
                }
                a = v;
              }
        __syncthreads();  // everyone has read the old activations
#pragma unroll
        for (int u = 0; u < JX; ++u)
          if (on[u])
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
              for (int h = 0; h < 5; ++h) {
                act[((half * 6 + 2 * pr) * 5 + h) * Wp + sc[u]] = acc[u][pr][h].x;
                act[((half * 6 + 2 * pr + 1) * 5 + h) * Wp + sc[u]] = acc[u][pr][h].y;
              }
        __syncthreads();
        PHASE_XF(4 + 2 * l);
      } else {
#pragma unroll
        for (int u = 0; u < JX; ++u)
          if (on[u] && cg == 0) {
            float* o = out + gO[u];
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
              for (int h = 0; h < 5; ++h) {
                o[((half * 6 + 2 * pr) * 5 + h) * gN[u]] = acc[u][pr][h].x;
                o[((half * 6 + 2 * pr + 1) * 5 + h) * gN[u]] = acc[u][pr][h].y;
              }
          }
        PHASE_XF(12);
      }
    }
  }
}

}  // namespace

#ifdef SSTG_PHASE_CLOCKS
extern "C" int sstg_debug_phase_clocks_xf(long long* host_out, int n) {
  return cudaMemcpyFromSymbol(host_out, g_phase_xf, sizeof(long long) * (n < 160 * 16 ? n : 160 * 16)) == cudaSuccess ? 0 : -3;
}
#endif

int txpcnn_tile_width(int max_n) {
  int Wp = txp::X_COLS + SSTG_XTILE_MAX_SCENES + 1;
  if (max_n + 2 > Wp) Wp = max_n + 2;
  return (Wp + 3) & ~3;
}

int launch_txpcnn_forward(const float* H, const int32_t* seg_off, const int32_t* tile_off, int n_tiles, int max_n,
                          const float* params, float* out, float* Usave, float* Xsave, long long layer_stride,
                          int tile_cols, cudaStream_t st, const float* bn_stats, int bn_scenes, double bn_momentum,
                          float* bn) {
  if (n_tiles == 0) return SSTG_OK;
  const bool save = Usave != nullptr && Xsave != nullptr;
  // inference: the Winograd F(2,3) kernel (txpcnn_wino.cu) -- 64-pair tiles for scenes of up to 128 pedestrians,
  // 128-pair tiles (256-thread CTAs) when the batch holds larger scenes (a scene > 128 columns is a tile of its own)
  if (!save && max_n <= SSTG_MAX_PEDS) {
    const int mode = txpcnn_wino_mode();
    if (mode != 0)
      return launch_txpcnn_forward_wino(H, seg_off, tile_off, n_tiles, params, out, mode,
                                        max_n <= X_COLS ? tile_cols : -(max_n + 4), st);
  }
  // widest tile: up to X_COLS columns of up to MAX_SCENES scenes, or one big scene;
  // one zero separator column before every scene and one after the last
  const int Wp = txpcnn_tile_width(max_n);
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
    if (save) {
      const bool with_bn = bn != nullptr && bn_scenes > 0;
      SSTG_CUDA_TRY(launch_pdl(kern, grid + (with_bn ? 1 : 0), X_THREADS, smem, st, H, seg_off, tile_off, n_tiles, Wp,
                               params, out, Usave, Xsave, layer_stride, bn_stats, bn_scenes, bn_momentum,
                               with_bn ? bn : nullptr),
                    "txpcnn launch");
      return SSTG_OK;
    }
    kern<<<grid, X_THREADS, smem, st>>>(H, seg_off, tile_off, n_tiles, Wp, params, out, Usave, Xsave, layer_stride,
                                        nullptr, 0, 0.0, nullptr);
    SSTG_CUDA_TRY(cudaGetLastError(), "txpcnn launch");
    return SSTG_OK;
  };
  if (max_n > X_COLS) return save ? launch(txpcnn_fwd_kernel<2, true>) : launch(txpcnn_fwd_kernel<2, false>);
  return save ? launch(txpcnn_fwd_kernel<1, true>) : launch(txpcnn_fwd_kernel<1, false>);
}

}  // namespace sstg
