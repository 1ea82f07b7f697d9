// Kernel (3'), inference only: the TXP-CNN (reference model.py:187-195) with the 3x3 convolutions
// evaluated by a 1-D Winograd F(2,3) transform along the pedestrian axis.
//
// For one output row and a PAIR of adjacent output columns (w, w+1) the three taps along the
// pedestrian axis need inputs d0..d3 = x[w-1..w+2]:
//     V = (d0 - d2, d1 + d2, d2 - d1, d1 - d3)            input transform  (4 adds, shared by all channels)
//     U = (g0, (g0+g1+g2)/2, (g0-g1+g2)/2, g2)            filter transform (once per CTA, in shared memory)
//     M_j += U_j V_j  summed over input channels and the 3 vertical taps    (4 FMAs instead of 6)
//     y(w) = M0 + M1 + M2,   y(w+1) = M1 - M2 - M3        output transform (once per layer)
// i.e. 1.5x fewer multiply-adds than the direct form, and the 4-vector (M0..M3) is two natural
// float2 pairs, so every FMA is a packed FFMA2 with NO operand duplication.
//
// A thread owns (channel group, column pair): M for COT = 12/NSPLIT output channels x 5 feature
// rows x 4 lives in registers.  Pairs never straddle scenes; in the shared activation tile every
// scene starts at an ODD column (1 or 2 zero separator columns in front of it), so d0..d3 are two
// aligned 8-byte loads and the zero padding / scene isolation still costs no select.  An odd
// scene's last pair computes a second output that is simply not stored.
// Everything else (whole-scene tiles, five layers in place in one 60-row tile, persistent grid)
// follows txpcnn_fwd.cu, which remains the kernel for training (SAVE) and for scenes > 128.
//
// Filters.  The transformed filters are warp-uniform operands.  CONSTW = true (full grids): they live in a
// __constant__ bank and reach the FMA pipe as UNIFORM-register pairs (`LDCU.64 UR, c[3][..]` ->
// `FFMA2 R, R, UR.F32x2, R`): no LDS.128, no staging registers in the inner loop -- 7.8 -> 7.0 ms per 131,072
// scenes.  For the address to be provably uniform the channel group of a thread must be a compile-time
// constant, so the main loop is instantiated once per group and selected by a warp-uniform branch.  The bank
// is refreshed in stream order before every launch (filter transform -> staging at the head of the output
// buffer -> cudaMemcpyToSymbolAsync): launches with DIFFERENT weights must therefore not overlap on one
// device (the only global state of the library; SSTG_TXP_FILTERS=smem selects the variant without it).
// CONSTW = false (small grids, or that knob): every CTA transforms the filters into its shared memory.
#include <cstdlib>
#include <type_traits>

#include "txpcnn_common.cuh"

namespace sstg {

namespace {

using namespace txp;

#ifndef WINO_NPF
#define WINO_NPF 3  // filter vectors in flight ahead of their use, shared-memory filters (2..6 measured)
#endif
#ifndef WINO_NPF_CONST
#define WINO_NPF_CONST 9   // the same for constant-bank filters: they wait in UNIFORM registers (36 of the warp's 63),
#endif                     // so a deeper ring costs no vector register.  Both must divide 3 * COT (ring phase).
// W_PAIRS (template parameter): column pairs per tile = threads per channel group.  64 (two 64-pedestrian scenes, 128-thread
// CTAs, two per SM) or 32 (64-thread CTAs, four per SM: the same warps per SM, but four independent tiles whose per-layer
// barriers and scalar epilogues drift apart -- sstg_plan_host cuts such tiles when SSTG_XTILE_PAIRS=32).
constexpr int W_U0 = 8 * 3 * 12 * 4;           // layer 0: [ci 8][kh 3][co 12][4]
constexpr int W_UL = 12 * 3 * 12 * 4;          // layers 1..4
constexpr int W_UTOTAL = W_U0 + 4 * W_UL;      // 8064 floats
constexpr int W_WP_MAX = 200;                  // widest tile row: 128 columns + up to 2 separators per scene
constexpr int W_WP_BIG = 264;                  // ... of the 128-pair form: one scene of up to 256 columns

__constant__ float c_U[W_UTOTAL + 4 * 32];  // transformed filters of the launch in flight (CONSTW variant);
                                            // the tail pads the prefetch ring's run past the last filter

// scene placement of one tile inside the shared activation tile
struct WTab {
  int s_first[SSTG_XTILE_MAX_SCENES + 1];  // first (odd) shared-memory column of each scene
  int s_pair[SSTG_XTILE_MAX_SCENES + 1];   // first pair index of each scene
  int s_n[SSTG_XTILE_MAX_SCENES];          // pedestrians of each scene
  int s_sb[SSTG_XTILE_MAX_SCENES];         // first tile column of each scene
  int n_pairs, used_cols;
};
struct WTabPF : WTab {  // the prefetch variant keeps the tile's extent next to its table (computed a tile ahead)
  int nsc, p_lo, cols;
};
__device__ __forceinline__ int tab_p_lo(const WTab&) { return 0; }
__device__ __forceinline__ int tab_p_lo(const WTabPF& t) { return t.p_lo; }

template <bool CONSTW, bool PF>
struct WSmemT {
  float U[CONSTW ? 4 : W_UTOTAL];
  float bias[5][12];
  float slope[4];
  // PF: the table of the tile being computed and of the tile being prefetched
  typename std::conditional<PF, WTabPF, WTab>::type tab[PF ? 2 : 1];
};

// W_WP (even, <= W_WP_MAX) is the row stride of the activation tile: the launcher sizes it for the widest
// tile of the batch, so dense batches (2 scenes of 64 -> 134 columns) fit three CTAs per SM.
// U[l][ci][kh][co][4] = (g0, (g0+g1+g2)/2, (g0-g1+g2)/2, g2) of filter row kh
__device__ __forceinline__ float4 wino_filter(const float* __restrict__ params, int i) {
  int l, r;
  if (i < W_U0 / 4) {
    l = 0;
    r = i;
  } else {
    l = 1 + (i - W_U0 / 4) / (W_UL / 4);
    r = (i - W_U0 / 4) % (W_UL / 4);
  }
  const int co = r % 12, kh = (r / 12) % 3, ci = r / 36;
  const int cin = (l == 0) ? 8 : 12;
  const int base = (l == 0) ? P_TP0_W : (l == 4 ? P_TPOUT_W : P_TP1_W + (l - 1) * P_TP_STRIDE);
  const float* g = params + base + (co * cin + ci) * 9 + kh * 3;
  const float g0 = g[0], g1 = g[1], g2 = g[2];
  return make_float4(g0, 0.5f * (g0 + g1 + g2), 0.5f * (g0 - g1 + g2), g2);
}

__global__ void wino_filter_kernel(const float* __restrict__ params, float* __restrict__ U) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W_UTOTAL / 4) reinterpret_cast<float4*>(U)[i] = wino_filter(params, i);
}

// PF = true (SSTG_TXP_PREFETCH=1, not the default yet): the 40 input rows of the NEXT tile are fetched into a separate
// staging tile while layers 1..4 of the current tile run, so the per-tile prologue -- scene placement, zeroing and one
// global-memory round trip behind three barriers: 12.5 % of the stall samples of the plain kernel
// (profiles/r1am_ncu_full_summary.csv) although only 2 % of its instructions -- leaves the critical path.  Layer 0 then reads
// the staging tile, everything else is unchanged (same arithmetic, same order: bit-identical results).
template <int NSPLIT, bool CONSTW, bool PF, int W_PAIRS = SSTG_XTILE_MAX_PAIRS>
__global__ void __launch_bounds__(NSPLIT * W_PAIRS, 128 / W_PAIRS)
txpcnn_wino_kernel(const float* __restrict__ Hin, const int32_t* __restrict__ seg_off,
                   const int32_t* __restrict__ tile_off, int n_tiles, const float* __restrict__ params,
                   float* __restrict__ out, int W_WP) {
  constexpr int COT = 12 / NSPLIT;
  constexpr int THREADS = NSPLIT * W_PAIRS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using WSmem = WSmemT<CONSTW, PF>;
  WSmem& sm = *reinterpret_cast<WSmem*>(smem_raw);
  float* act = reinterpret_cast<float*>(smem_raw + ((sizeof(WSmem) + 15) & ~size_t(15)));  // [60][W_WP]
  float* stage = act + (PF ? X_ROWS * W_WP : 0);                                           // PF: [40][W_WP] input of the tile
  const int tid = threadIdx.x;
  const int grp = tid / W_PAIRS;  // output-channel group (warp-uniform)
  const int pt = tid - grp * W_PAIRS;
  const int co0 = grp * COT;

  // ---- filter transform into shared memory (the CONSTW variant reads the constant bank instead)
  if constexpr (!CONSTW) {
    for (int i = tid; i < W_UTOTAL / 4; i += THREADS) reinterpret_cast<float4*>(sm.U)[i] = wino_filter(params, i);
  }
  if (tid < 60) {
    const int l = tid / 12, co = tid % 12;
    const int base = (l == 0) ? P_TP0_B : (l == 4 ? P_TPOUT_B : P_TP1_W + (l - 1) * P_TP_STRIDE + 1296);
    sm.bias[l][co] = params[base + co];
  }
  if (tid < 4) sm.slope[tid] = params[P_PRELUS + tid];

  // ---- helpers of the prefetch variant
  // scene placement of tile `tile` (one thread): every scene starts at an odd column
  auto place = [&](int tile, WTabPF& tb) {
    const int s_lo = tile_off[tile], s_hi = tile_off[tile + 1];
    const int nsc = s_hi - s_lo;
    const int p_lo = seg_off[s_lo];
    int c = 1, pr = 0;
    for (int j = 0; j < nsc; ++j) {
      const int sb = seg_off[s_lo + j] - p_lo, n = seg_off[s_lo + j + 1] - p_lo - sb;
      tb.s_first[j] = c;
      tb.s_pair[j] = pr;
      tb.s_n[j] = n;
      tb.s_sb[j] = sb;
      pr += (n + 1) >> 1;
      c += n + 1;
      c += (~c) & 1;
    }
    tb.s_first[nsc] = c;
    tb.s_pair[nsc] = pr;
    tb.n_pairs = pr;
    tb.used_cols = c + 1;
    tb.nsc = nsc;
    tb.p_lo = p_lo;
    tb.cols = seg_off[s_hi] - p_lo;
  };
  // zero column 0 and the two columns behind every scene in `rows` rows of `buf`
  auto zero_gaps = [&](const WTabPF& tb, float* buf, int rows) {
    for (int i = tid; i < (2 * tb.nsc + 1) * rows; i += THREADS) {
      const int row = i % rows, k = i / rows;
      const int col = (k == 0) ? 0 : tb.s_first[(k - 1) >> 1] + tb.s_n[(k - 1) >> 1] + ((k - 1) & 1);
      if (col < W_WP) buf[row * W_WP + col] = 0.f;
    }
  };
  // issue the cp.async loads of the 40 input rows of every scene of the tile into `buf` (no wait)
  auto load_rows = [&](const WTabPF& tb, float* buf) {
    const float* Hb = Hin + static_cast<size_t>(tb.p_lo) * 40;
    for (int q = tid & 127; q < tb.cols; q += 128) {
      int j = 0;
      while (j + 1 < tb.nsc && tb.s_sb[j + 1] <= q) ++j;
      const int n = q - tb.s_sb[j], ns = tb.s_n[j];
      const float* src = Hb + tb.s_sb[j] * 40 + n;
      float* dst = buf + tb.s_first[j] + n;
      for (int row = (THREADS % 128 == 0) ? (tid >> 7) : 0; row < 40; row += (THREADS % 128 == 0) ? THREADS / 128 : 1)
        if (THREADS % 128 == 0 || tid < 128) cp_async4(dst + row * W_WP, src + row * ns);
    }
  };

  int cur = 0;  // PF: index of the current tile's table
  if constexpr (PF) {
    if (static_cast<int>(blockIdx.x) < n_tiles) {  // the CTA's first tile: fetched in the open
      if (tid == 0) place(blockIdx.x, sm.tab[0]);
      __syncthreads();
      zero_gaps(sm.tab[0], stage, 40);
      __syncthreads();
      load_rows(sm.tab[0], stage);
    }
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    auto& tb = sm.tab[PF ? cur : 0];
    const int s_lo = PF ? 0 : tile_off[tile], s_hi = PF ? 0 : tile_off[tile + 1];
    const int nsc = s_hi - s_lo;
    const int p_lo = PF ? tab_p_lo(tb) : seg_off[s_lo];
    const int cols = PF ? 0 : seg_off[s_hi] - p_lo;
    if constexpr (PF) {
      cp_async_wait_all();
      __syncthreads();  // previous tile done; this tile's table and staging rows (prefetched) visible
      zero_gaps(tb, act, X_ROWS);  // ordered before the first stores into `act` by layer 0's barrier
    } else {
      __syncthreads();  // previous tile done; U visible
      if (tid == 0) {   // scene placement: every scene starts at an odd column
        int c = 1, pr = 0;
        for (int j = 0; j < nsc; ++j) {
          const int sb = seg_off[s_lo + j] - p_lo, n = seg_off[s_lo + j + 1] - p_lo - sb;
          tb.s_first[j] = c;
          tb.s_pair[j] = pr;
          tb.s_n[j] = n;
          tb.s_sb[j] = sb;
          pr += (n + 1) >> 1;
          c += n + 1;        // one separator ...
          c += (~c) & 1;     // ... or two, so that the next scene starts odd again
        }
        tb.s_first[nsc] = c;  // one past the last separator pair
        tb.s_pair[nsc] = pr;
        tb.n_pairs = pr;
        tb.used_cols = c + 1;
      }
      __syncthreads();
      // ---- zero column 0 and the two columns behind every scene (the 1- or 2-column gap; when the gap
      // is one column the second store hits the next scene's first column, which the loads below rewrite)
      for (int i = tid; i < (2 * nsc + 1) * X_ROWS; i += THREADS) {
        const int row = i % X_ROWS, k = i / X_ROWS;
        const int col = (k == 0) ? 0 : tb.s_first[(k - 1) >> 1] + tb.s_n[(k - 1) >> 1] + ((k - 1) & 1);
        if (col < W_WP) act[row * W_WP + col] = 0.f;
      }
      __syncthreads();
      // ---- 40 input rows of every scene (cp.async, coalesced along the pedestrian axis)
      {
        const float* Hb = Hin + static_cast<size_t>(p_lo) * 40;
        for (int q = tid & 127; q < cols; q += 128) {
          int j = 0;
          while (j + 1 < nsc && tb.s_sb[j + 1] <= q) ++j;
          const int n = q - tb.s_sb[j], ns = tb.s_n[j];
          const float* src = Hb + tb.s_sb[j] * 40 + n;
          float* dst = act + tb.s_first[j] + n;
          for (int row = (THREADS % 128 == 0) ? (tid >> 7) : 0; row < 40; row += (THREADS % 128 == 0) ? THREADS / 128 : 1)
            if (THREADS % 128 == 0 || tid < 128) cp_async4(dst + row * W_WP, src + row * ns);
        }
        cp_async_wait_all();
      }
      __syncthreads();
    }

    // ---- this thread's pair
    const bool on = pt < tb.n_pairs;
    int sc0 = 1, pn = 1;
    bool v1 = false;
    long long gO = 0;
    if (on) {
      int j = 0;
      while (tb.s_pair[j + 1] <= pt) ++j;
      const int n = 2 * (pt - tb.s_pair[j]);
      sc0 = tb.s_first[j] + n;
      pn = tb.s_n[j];
      v1 = n + 1 < pn;
      gO = static_cast<long long>(p_lo + tb.s_sb[j]) * 60 + n;
    }

#pragma unroll 1
    for (int l = 0; l < 5; ++l) {
      float2 Ma[COT][5], Mb[COT][5];  // (M0, M1), (M2, M3)
#pragma unroll
      for (int c = 0; c < COT; ++c)
#pragma unroll
        for (int h = 0; h < 5; ++h) Ma[c][h] = Mb[c][h] = make_float2(0.f, 0.f);
      auto main_loop = [&](auto grp_c) {
        // Software pipeline (only 8-12 warps per SM hide latency here, so the loads are rotated by hand):
        // the 10 activation loads of input channel ci+1 and the next NPF filter vectors are in flight
        // while the 10 FFMA2 of the current (kh, channel) step issue.
        constexpr int GRP = decltype(grp_c)::value;  // CONSTW: compile-time group => uniform filter addresses
        const float* Ul = (CONSTW ? c_U + GRP * COT * 4 : sm.U + co0 * 4) + (l == 0 ? 0 : W_U0 + (l - 1) * W_UL);
        const int CI = (l == 0) ? 8 : 12;
        constexpr int STEPS = 3 * COT;
        constexpr int NPF = !CONSTW ? WINO_NPF
                            : (STEPS % WINO_NPF_CONST == 0) ? WINO_NPF_CONST
                            : (STEPS % 6 == 0) ? 6 : WINO_NPF;
        static_assert(STEPS % NPF == 0, "the filter ring must keep its phase from one input channel to the next");
        const float* a = ((PF && l == 0) ? stage : act) + sc0 - 1;  // even offset: aligned float2 loads
        float2 d01[5], d23[5];
#pragma unroll
        for (int h = 0; h < 5; ++h) {
          d01[h] = *reinterpret_cast<const float2*>(a + h * W_WP);
          d23[h] = *reinterpret_cast<const float2*>(a + h * W_WP + 2);
        }
        float4 ub[NPF];
#pragma unroll
        for (int i = 0; i < NPF; ++i) ub[i] = *reinterpret_cast<const float4*>(Ul + ((i / COT) * 12 + (i % COT)) * 4);
#pragma unroll 1
        for (int ci = 0; ci < CI; ++ci) {
          float2 Va[5], Vb[5];
#pragma unroll
          for (int h = 0; h < 5; ++h) {
            Va[h] = make_float2(d01[h].x - d23[h].x, d01[h].y + d23[h].x);
            Vb[h] = make_float2(d23[h].x - d01[h].y, d01[h].y - d23[h].y);
          }
          // prefetch the next input channel's activations (the last trip re-reads row block CI-1: harmless)
          const float* an = a + ((ci + 1 < CI ? ci + 1 : ci) * 5) * W_WP;
#pragma unroll
          for (int h = 0; h < 5; ++h) {
            d01[h] = *reinterpret_cast<const float2*>(an + h * W_WP);
            d23[h] = *reinterpret_cast<const float2*>(an + h * W_WP + 2);
          }
          const float* Uc = Ul + ci * (3 * 12 * 4);
#pragma unroll
          for (int st = 0; st < STEPS; ++st) {
            const float4 u = ub[st % NPF];
            const int s2 = st + NPF;  // may run into the next input channel's block (contiguous in U)
            const int off = (s2 < STEPS) ? ((s2 / COT) * 12 + (s2 % COT)) * 4
                                         : (36 + ((s2 - STEPS) / COT) * 12 + ((s2 - STEPS) % COT)) * 4;
            ub[st % NPF] = *reinterpret_cast<const float4*>(Uc + off);
            const int kh = st / COT, c = st % COT;
            const float2 ua = make_float2(u.x, u.y), ubb = make_float2(u.z, u.w);
#pragma unroll
            for (int h = 0; h < 5; ++h) {
              const int hi = h + kh - 1;
              if (hi >= 0 && hi < 5) {
                ffma2(Ma[c][h], ua, Va[hi]);
                ffma2(Mb[c][h], ubb, Vb[hi]);
              }
            }
          }
        }
      };
      if (on) {
        if constexpr (CONSTW) {
          static_assert(NSPLIT <= 4, "channel groups");
          if (grp == 0) main_loop(std::integral_constant<int, 0>{});
          else if (grp == 1) main_loop(std::integral_constant<int, 1>{});
          else if (grp == 2) main_loop(std::integral_constant<int, 2 % NSPLIT>{});
          else main_loop(std::integral_constant<int, 3 % NSPLIT>{});
        } else {
          main_loop(std::integral_constant<int, 0>{});
        }
      }
      // ---- output transform, bias, PReLU, residual
      float y0[COT][5], y1[COT][5];
      const float sl = l < 4 ? sm.slope[l] : 0.f;
#pragma unroll
      for (int c = 0; c < COT; ++c) {
        const float b = sm.bias[l][co0 + c];
#pragma unroll
        for (int h = 0; h < 5; ++h) {
          float a0 = (Ma[c][h].x + Ma[c][h].y) + Mb[c][h].x + b;
          float a1 = (Ma[c][h].y - Mb[c][h].x) - Mb[c][h].y + b;
          if (l < 4) {
            a0 = prelu(a0, sl);
            a1 = prelu(a1, sl);
            if (l > 0) {  // residual (model.py:192)
              const float* r = act + ((co0 + c) * 5 + h) * W_WP + sc0;
              a0 += r[0];
              a1 += r[1];
            }
          }
          y0[c][h] = a0;
          y1[c][h] = a1;
        }
      }
      if (l < 4) {
        if constexpr (PF) {  // the next tile's placement, published by the barrier below
          if (l == 0 && tid == 0 && tile + static_cast<int>(gridDim.x) < n_tiles)
            place(tile + static_cast<int>(gridDim.x), sm.tab[cur ^ 1]);
        }
        __syncthreads();  // everyone has read the old activations
        if constexpr (PF) {  // layer 0 was the last reader of the staging tile
          if (l == 0 && tile + static_cast<int>(gridDim.x) < n_tiles) zero_gaps(sm.tab[cur ^ 1], stage, 40);
        }
        if (on) {
#pragma unroll
          for (int c = 0; c < COT; ++c)
#pragma unroll
            for (int h = 0; h < 5; ++h) {
              float* w = act + ((co0 + c) * 5 + h) * W_WP + sc0;
              w[0] = y0[c][h];
              if (v1) w[1] = y1[c][h];  // an odd scene's last pair must leave the separator column zero
            }
        }
        __syncthreads();
        if constexpr (PF) {  // in flight during layers 1..4; awaited at the top of the next tile
          if (l == 0 && tile + static_cast<int>(gridDim.x) < n_tiles) load_rows(sm.tab[cur ^ 1], stage);
        }
      } else if (on) {
        float* o = out + gO;
#pragma unroll
        for (int c = 0; c < COT; ++c)
#pragma unroll
          for (int h = 0; h < 5; ++h) {
            float* w = o + static_cast<long long>((co0 + c) * 5 + h) * pn;
            w[0] = y0[c][h];
            if (v1) w[1] = y1[c][h];
          }
      }
    }
    if constexpr (PF) cur ^= 1;
  }
}

template <int NSPLIT, bool CONSTW, bool PF = false, int W_PAIRS = SSTG_XTILE_MAX_PAIRS>
int launch_wino_v(const float* H, const int32_t* seg_off, const int32_t* tile_off, int n_tiles, const float* params,
                  float* out, int W_WP, int grid_cap, cudaStream_t st) {
  const size_t smem = ((sizeof(WSmemT<CONSTW, PF>) + 15) & ~size_t(15)) + sizeof(float) * (X_ROWS + (PF ? 40 : 0)) * W_WP;
  auto kern = txpcnn_wino_kernel<NSPLIT, CONSTW, PF, W_PAIRS>;
  grid_cap = (grid_cap / 2) * (128 / W_PAIRS);   // CTAs per SM: 1 / 2 / 4 for 128 / 64 / 32 pairs per tile
  SSTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                "txpcnn_wino smem attr");
  const int grid = n_tiles < grid_cap ? n_tiles : grid_cap;
  if (CONSTW) {
    // refresh the constant bank in stream order; the head of `out` (>= 60 floats per pedestrian, >= n_tiles
    // pedestrians) is the staging buffer -- the kernel below overwrites it afterwards
    wino_filter_kernel<<<(W_UTOTAL / 4 + 255) / 256, 256, 0, st>>>(params, out);
    SSTG_CUDA_TRY(cudaGetLastError(), "wino_filter launch");
    SSTG_CUDA_TRY(cudaMemcpyToSymbolAsync(c_U, out, sizeof(float) * W_UTOTAL, 0, cudaMemcpyDeviceToDevice, st),
                  "txpcnn_wino filters -> constant bank");
  }
  kern<<<grid, NSPLIT * W_PAIRS, smem, st>>>(H, seg_off, tile_off, n_tiles, params, out, W_WP);
  SSTG_CUDA_TRY(cudaGetLastError(), "txpcnn_wino launch");
  return SSTG_OK;
}

// SSTG_TXP_FILTERS = smem: never use the constant bank (the library's only global state, see the header comment)
bool wino_const_allowed() {
  const char* e = getenv("SSTG_TXP_FILTERS");
  return !(e && e[0] == 's');
}
// SSTG_TXP_PREFETCH = 1: the next-tile prefetch variant of the 2-group constant-bank kernel (A/B knob; off by default
// until it has been measured)
bool wino_prefetch() {
  const char* e = getenv("SSTG_TXP_PREFETCH");
  return e && e[0] == '1';
}

template <int NSPLIT>
int launch_wino(const float* H, const int32_t* seg_off, const int32_t* tile_off, int n_tiles, const float* params,
                float* out, int W_WP, cudaStream_t st, bool big = false) {
  int dev = 0, sms = 148;
  SSTG_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
  SSTG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "sm count");
  const int grid_cap = 2 * sms;  // 255 registers x 128 threads: two CTAs per SM
  // the constant-bank variant pays a ~6 us refresh per launch: worth it once the persistent grid is full
  // (then n_tiles * 60 floats of `out` also exceed the 8,064-float staging area by far)
  if constexpr (NSPLIT == 2) {
    if (big) {   // scenes of 129..256 pedestrians: 128 pairs per tile, 256 threads, one CTA per SM
      if (n_tiles >= sms && wino_const_allowed())
        return launch_wino_v<NSPLIT, true, false, 128>(H, seg_off, tile_off, n_tiles, params, out, W_WP, grid_cap, st);
      return launch_wino_v<NSPLIT, false, false, 128>(H, seg_off, tile_off, n_tiles, params, out, W_WP, grid_cap, st);
    }
  }
  if (n_tiles >= grid_cap && wino_const_allowed()) {
    if constexpr (NSPLIT == 2) {
      if (sstg_xtile_pairs() == 32)   // tiles of <= 32 pairs (sstg_plan_host under the same knob): 64-thread CTAs, four per SM
        return launch_wino_v<NSPLIT, true, false, 32>(H, seg_off, tile_off, n_tiles, params, out, W_WP, grid_cap, st);
      if (wino_prefetch())
        return launch_wino_v<NSPLIT, true, true>(H, seg_off, tile_off, n_tiles, params, out, W_WP, grid_cap, st);
    }
    return launch_wino_v<NSPLIT, true>(H, seg_off, tile_off, n_tiles, params, out, W_WP, grid_cap, st);
  }
  return launch_wino_v<NSPLIT, false>(H, seg_off, tile_off, n_tiles, params, out, W_WP, grid_cap, st);
}

}  // namespace

// SSTG_TXP_WINO = 0 (direct kernel), 2 or 4 (Winograd with 2 / 4 channel groups per column pair).
int txpcnn_wino_mode() {
  const char* e = getenv("SSTG_TXP_WINO");
  if (e && (e[0] == '0' || e[0] == '2' || e[0] == '3' || e[0] == '4')) return e[0] - '0';
  return 2;
}

// tile_cols_hint: an upper bound of the shared-memory columns any tile of this batch needs
// (1 + sum over its scenes of N_s + gap), 0 for the worst case of the 64-pair form, or NEGATIVE:
// -(pedestrians of the largest scene + 4) for batches with scenes of 129..256 pedestrians (128-pair form).
int launch_txpcnn_forward_wino(const float* H, const int32_t* seg_off, const int32_t* tile_off, int n_tiles,
                               const float* params, float* out, int mode, int tile_cols_hint, cudaStream_t st) {
  if (n_tiles == 0) return SSTG_OK;
  if (tile_cols_hint < 0) {   // -(columns of the largest scene + 4): the batch holds scenes of more than 128 pedestrians
    int wpb = (-tile_cols_hint + 3) & ~3;
    if (wpb > W_WP_BIG) wpb = W_WP_BIG;
    if (wpb < W_WP_MAX) wpb = W_WP_MAX;     // smaller scenes still share tiles of up to 128 columns + separators
    return launch_wino<2>(H, seg_off, tile_off, n_tiles, params, out, wpb, st, true);
  }
  int wp = (tile_cols_hint > 0 && tile_cols_hint < W_WP_MAX) ? tile_cols_hint : W_WP_MAX;
  wp = (wp + 3) & ~3;
  if (wp > W_WP_MAX) wp = W_WP_MAX;
  if (mode == 4) return launch_wino<4>(H, seg_off, tile_off, n_tiles, params, out, wp, st);
  if (mode == 3) return launch_wino<3>(H, seg_off, tile_off, n_tiles, params, out, wp, st);
  return launch_wino<2>(H, seg_off, tile_off, n_tiles, params, out, wp, st);
}

}  // namespace sstg
