// Training-path kernels, part 2: backward pass of the TXP-CNN (reference model.py:187-195,
// differentiated by autograd in train.py:67).
//
// Forward (saved by txpcnn_fwd_kernel<.., SAVE=true>):
//   x0 = view(H);  u0 = conv0(x0), x1 = PReLU0(u0);  u_k = conv_k(x_k), x_{k+1} = PReLU_k(u_k) + x_k (k=1..3);
//   out = conv_out(x4)
// Backward, per tile of whole scenes (the same tiling as the forward):
//   G4 = conv_out^T(dOut);   for k = 3..1:  du_k = G_{k+1} * PReLU_k'(u_k),  G_k = G_{k+1} + conv_k^T(du_k)
//   du_0 = G_1 * PReLU_0'(u_0),  dH = conv_0^T(du_0)
//   dW_k = du_k (*) x_k (valid correlation over the scene's zero-padded plane), db_k = sum du_k,
//   dslope_k = sum G_{k+1} * u_k over u_k <= 0.
// The transposed convolutions reuse the forward's register-tiled column convolution with
// transposed + flipped weights; a thread keeps G for its own (channel half, column) in
// registers.  Weight gradients: a warp owns (output-channel half, input channel) units,
// lanes stride over the tile's columns with 6x9 register accumulators, one warp reduction
// per unit, and the totals go to a per-CTA shared accumulator that only that warp touches
// -- so the order of every floating-point addition is fixed (no atomics).  Each CTA writes
// its accumulator as one row of `partial`; sstg_grad_finalize adds the rows in order.
//
// Small tiles (the reference's optimiser groups are ~128 scenes of ~10 pedestrians: one scene per CTA,
// the step is LATENCY-bound): a tile of <= 64 columns spreads each column's transposed convolution over
// 2 or 4 lanes (slices of the 12 channels, butterfly-summed so all lanes hold the same bits), gives every
// (output-channel pair, input channel) weight-gradient unit to a trio of lanes that split the columns
// (packed accumulators, two shuffles per value to combine), issues all loads of a global-load phase
// before their first use, and brings the next layer's activations in with cp.async underneath the
// transposed convolution.
#include "txpcnn_common.cuh"

namespace sstg {

namespace {

using namespace txp;

constexpr int B_WT = 5 * X_WL;                     // transposed weights, 12 "input" channels per layer
constexpr int B_GRAD = SSTG_N_PARAMS - P_TP0_W;    // 7421 TXP-CNN parameters
constexpr int B_GRAD_STRIDE = SSTG_TXP_GRAD_STRIDE;

// -DSSTG_PHASE_CLOCKS: thread 0 of every CTA stamps clock64() at the phase boundaries of its FIRST tile
// (profiling aid for the latency-bound small-group case; read back with sstg_debug_phase_clocks).
#ifdef SSTG_PHASE_CLOCKS
__device__ long long g_phase[160 * 40];
#define PHASE(k)                                                                         \
  do {                                                                                   \
    if (threadIdx.x == 0 && tile == static_cast<int>(blockIdx.x)) g_phase[blockIdx.x * 40 + (k)] = clock64(); \
  } while (0)
#else
#define PHASE(k)
#endif

struct BSmem {
  float wT[B_WT];
  float slope[4];
  int scol[2 * X_COLS];
  int nstr[2 * X_COLS];
  int sbase[2 * X_COLS];
  float red[8][4];
};

__device__ __forceinline__ int w_base(int l) {
  return l == 0 ? P_TP0_W : (l == 4 ? P_TPOUT_W : P_TP1_W + (l - 1) * P_TP_STRIDE);
}
__device__ __forceinline__ int b_base(int l) {
  return l == 0 ? P_TP0_B : (l == 4 ? P_TPOUT_B : P_TP1_W + (l - 1) * P_TP_STRIDE + 1296);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dW[hf*6 + co][ci][kh][kw] += sum_{h, c} du[hf*6+co][h][c] * x[ci][h+kh-1][c+kw-1] over the tile's columns.
// Packed FMAs over channel pairs: acc[pr][tap] = (dW[2pr][tap], dW[2pr+1][tap]).
__device__ __forceinline__ void dw_unit(const float* __restrict__ DU, const float* __restrict__ XT, int Wp,
                                        int c_end, int ci, int hf, int cin, int lane, float* __restrict__ gw) {
  float2 acc[3][9];
#pragma unroll
  for (int pr = 0; pr < 3; ++pr)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[pr][k] = make_float2(0.f, 0.f);
  for (int c = 1 + lane; c <= c_end; c += 32) {
    float2 x[5][3];
    const float* xp = XT + (ci * 5) * Wp + c - 1;
#pragma unroll
    for (int h = 0; h < 5; ++h)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float v = xp[h * Wp + kw];
        x[h][kw] = make_float2(v, v);
      }
#pragma unroll
    for (int pr = 0; pr < 3; ++pr) {
      float2 d[5];
#pragma unroll
      for (int h = 0; h < 5; ++h)
        d[h] = make_float2(DU[((hf * 6 + 2 * pr) * 5 + h) * Wp + c], DU[((hf * 6 + 2 * pr + 1) * 5 + h) * Wp + c]);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int h = 0; h < 5; ++h) {
            const int hi = h + kh - 1;
            if (hi >= 0 && hi < 5) ffma2(acc[pr][kh * 3 + kw], d[h], x[hi][kw]);
          }
    }
  }
  // ---- warp reduction of the 54 partial sums (padded to 64) as a reduce-scatter: at every butterfly
  // step a lane hands over the half of its values its partner keeps, so 32+16+8+4+2 = 62 shuffles
  // (instead of 54 x 5) leave lane L with the totals of values 2L and 2L+1, and the lanes then update
  // the accumulator in parallel.
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 0.f;
#pragma unroll
  for (int pr = 0; pr < 3; ++pr)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      v[(pr * 9 + k) * 2] = acc[pr][k].x;
      v[(pr * 9 + k) * 2 + 1] = acc[pr][k].y;
    }
#pragma unroll
  for (int half = 32; half >= 2; half >>= 1) {
    const int o = half >> 1;                       // 16, 8, 4, 2, 1
    const bool up = (lane & o) != 0;               // this lane keeps the upper half
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int idx = 2 * lane + j;                  // value index: (pr*9 + k)*2 + e, 54 of them
    if (idx < 54) {
      const int e = idx & 1, pk = idx >> 1, pr = pk / 9, k = pk - pr * 9;
      gw[((hf * 6 + 2 * pr + e) * cin + ci) * 9 + k] += v[j];
    }
  }
}

// Small tile (<= 64 columns): three adjacent lanes own one (output-channel PAIR, input channel) unit and a
// third of the tile's columns each; a lane walks its columns in ascending order with a sliding 5x3 window
// of x and 9 packed accumulators (18 weight gradients), the three slices are added in slice order through
// two shuffles per value, and lane 0 of the trio updates the CTA accumulator entries only it ever touches.
// 10 trios per warp, 8 warps: 72 units (12 input channels) in one pass.
// The bias gradients db[co] = sum of du over the plane ride along in the trios of input channel 0.
__device__ __forceinline__ void dw_pair_slices(const float* __restrict__ DU, const float* __restrict__ XT, int Wp,
                                               int c_end, int cin, int warp, int lane, float* __restrict__ gw,
                                               float* __restrict__ gb) {
  const int grp = lane / 3, sl = lane - grp * 3;
  const int pu = warp * 10 + grp;
  const bool act = lane < 30 && pu < 6 * cin;
  const int p = act ? pu / cin : 0, ci = act ? pu - p * cin : 0;
  const int per = (c_end + 2) / 3;
  const int c_lo = 1 + sl * per;
  int c_hi = c_lo + per - 1;
  if (c_hi > c_end) c_hi = c_end;
  if (!act) c_hi = 0;
  float2 acc[9], accb = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = make_float2(0.f, 0.f);
  const float* xr = XT + (ci * 5) * Wp;
  const float* d0 = DU + ((2 * p) * 5) * Wp;
  const float* d1 = d0 + 5 * Wp;
  float x0[5], x1[5], x2[5];
#pragma unroll
  for (int h = 0; h < 5; ++h) {
    x0[h] = xr[h * Wp + c_lo - 1];
    x1[h] = xr[h * Wp + c_lo];
  }
#pragma unroll 2
  for (int c = c_lo; c <= c_hi; ++c) {
    float2 d[5];
#pragma unroll
    for (int h = 0; h < 5; ++h) {
      x2[h] = xr[h * Wp + c + 1];
      d[h] = make_float2(d0[h * Wp + c], d1[h * Wp + c]);
    }
#pragma unroll
    for (int h = 0; h < 5; ++h) {
      accb.x += d[h].x;
      accb.y += d[h].y;
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int h = 0; h < 5; ++h) {
        const int hi = h + kh - 1;
        if (hi >= 0 && hi < 5) {
          ffma2(acc[kh * 3 + 0], d[h], make_float2(x0[hi], x0[hi]));
          ffma2(acc[kh * 3 + 1], d[h], make_float2(x1[hi], x1[hi]));
          ffma2(acc[kh * 3 + 2], d[h], make_float2(x2[hi], x2[hi]));
        }
      }
#pragma unroll
    for (int h = 0; h < 5; ++h) {
      x0[h] = x1[h];
      x1[h] = x2[h];
    }
  }
  float* g0 = gw + ((2 * p) * cin + ci) * 9;
  float* g1 = g0 + cin * 9;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float ax = (acc[k].x + __shfl_down_sync(0xffffffffu, acc[k].x, 1)) + __shfl_down_sync(0xffffffffu, acc[k].x, 2);
    const float ay = (acc[k].y + __shfl_down_sync(0xffffffffu, acc[k].y, 1)) + __shfl_down_sync(0xffffffffu, acc[k].y, 2);
    if (act && sl == 0) {
      g0[k] += ax;
      g1[k] += ay;
    }
  }
  {
    const float bx = (accb.x + __shfl_down_sync(0xffffffffu, accb.x, 1)) + __shfl_down_sync(0xffffffffu, accb.x, 2);
    const float by = (accb.y + __shfl_down_sync(0xffffffffu, accb.y, 1)) + __shfl_down_sync(0xffffffffu, accb.y, 2);
    if (act && sl == 0 && ci == 0) {
      gb[2 * p] += bx;
      gb[2 * p + 1] += by;
    }
  }
}

template <int JX>
__global__ void __launch_bounds__(X_THREADS, 1)
txpcnn_bwd_kernel(const float* __restrict__ dOut, const float* __restrict__ Usave, const float* __restrict__ Xsave,
                  long long layer_stride, const float* __restrict__ Hin, const int32_t* __restrict__ seg_off,
                  const int32_t* __restrict__ tile_off, int n_tiles, int Wp, const float* __restrict__ params,
                  float* __restrict__ dH, float* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BSmem& sm = *reinterpret_cast<BSmem*>(smem_raw);
  float* DU = reinterpret_cast<float*>(smem_raw + ((sizeof(BSmem) + 15) & ~size_t(15)));  // [60][Wp]
  float* XT = DU + X_ROWS * Wp;                                                          // [60][Wp]
  float* gacc = XT + X_ROWS * Wp;                                                        // [B_GRAD_STRIDE]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = tid / X_COLS;
  const int idx = tid - half * X_COLS;

  // ---- transposed + flipped weights, tap-major for the packed column convolution:
  //      wT[l][(b*2 + hf)*56 + tap'*6 + a6] = W_l[b][hf*6 + a6][8 - tap']
  //      (loads issued in batches of W_BATCH before the first store: one L2 round trip per batch)
  constexpr int W_BATCH = 9;
  for (int i0 = tid; i0 < B_WT; i0 += W_BATCH * X_THREADS) {
    float v[W_BATCH];
#pragma unroll
    for (int j = 0; j < W_BATCH; ++j) {
      const int i = i0 + j * X_THREADS;
      v[j] = 0.f;
      if (i < B_WT) {
        const int l = i / X_WL, r = i % X_WL;
        const int b = r / (2 * X_WSLOT), r2 = r % (2 * X_WSLOT);
        const int hf = r2 / X_WSLOT, r3 = r2 % X_WSLOT;
        if (r3 < 54) {
          const int a = hf * 6 + r3 % 6, tap = r3 / 6;
          const int cin = (l == 0) ? 8 : 12;
          if (a < cin) v[j] = params[w_base(l) + (b * cin + a) * 9 + (8 - tap)];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < W_BATCH; ++j) {
      const int i = i0 + j * X_THREADS;
      if (i < B_WT) sm.wT[i] = v[j];
    }
  }
  if (tid < 4) sm.slope[tid] = params[P_PRELUS + tid];
  for (int i = tid; i < B_GRAD_STRIDE; i += X_THREADS) gacc[i] = 0.f;
  float dsl[4] = {0.f, 0.f, 0.f, 0.f};
  pdl_wait();  // everything above read only `params`
  pdl_release();

#ifdef SSTG_PHASE_CLOCKS
  if (threadIdx.x == 0) g_phase[blockIdx.x * 40 + 39] = clock64();
#endif
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    PHASE(0);
    const int s_lo = tile_off[tile], s_hi = tile_off[tile + 1];
    const int p_lo = seg_off[s_lo];
    const int cols = seg_off[s_hi] - p_lo;
    const int nsc = s_hi - s_lo;
    const int c_end = cols + nsc - 1;  // last data column of the shared tile
    // lanes per column of the transposed convolution (CTA-uniform): 4 / 2 for small tiles, else 1
    const int cgs = JX != 1 ? 1 : (cols <= X_COLS / 4 ? 4 : (cols <= X_COLS / 2 ? 2 : 1));
    const int qt = cgs == 4 ? (idx >> 2) : (cgs == 2 ? (idx >> 1) : idx);
    const int cg = idx & (cgs - 1);
    // split columns: after the reduce-scatter of the transposed convolution, lane `cg` OWNS entries
    // [own_base, own_base + own_n) of the column's 15 (channel pair, feature row) entries -- it alone keeps
    // their dL/dx, loads their pre-activations and writes their du (entry e <-> pair e / 5, row e % 5).
    const int own_n = cgs == 4 ? 4 : 8;
    const int own_base = ((cg & 1) ? 8 : 0) + ((cgs == 4 && (cg & 2)) ? 4 : 0);
    __syncthreads();
    for (int q = tid; q < cols; q += X_THREADS) {
      int j = 0;
      while (seg_off[s_lo + j + 1] - p_lo <= q) ++j;
      const int sb = seg_off[s_lo + j] - p_lo;
      sm.scol[q] = q + j + 1;
      sm.nstr[q] = seg_off[s_lo + j + 1] - p_lo - sb;
      sm.sbase[q] = sb;
    }
    for (int i = tid; i < (nsc + 1) * X_ROWS; i += X_THREADS) {
      const int j = i / X_ROWS, row = i - j * X_ROWS;
      const int col = (j < nsc) ? (seg_off[s_lo + j] - p_lo + j) : (cols + nsc);
      DU[row * Wp + col] = 0.f;
      XT[row * Wp + col] = 0.f;
    }
    __syncthreads();
    PHASE(1);

    int sc[JX], gN[JX];
    long long gO[JX];
    bool on[JX];
    float2 G[JX][3][5];  // dL/dx of this thread's (channel pair, feature row), both columns
#pragma unroll
    for (int u = 0; u < JX; ++u) {
      const int q = qt + u * X_COLS;
      on[u] = q < cols;
      sc[u] = on[u] ? sm.scol[q] : 1;
      gN[u] = on[u] ? sm.nstr[q] : 1;
      const int sb = on[u] ? sm.sbase[q] : 0;
      gO[u] = static_cast<long long>(p_lo + sb) * 60 + (q - sb);
      // layer 4 (tpcnn_ouput, no activation): du_4 = dOut
      if (JX == 1 && cgs != 1) {  // split column: only the owned entries (kept in the first own_n slots of G)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int e = own_base + j;
          const bool mine = j < own_n && e < 15 && on[u];
          const int pr = e / 5, h = e - pr * 5;
          const int r0 = (half * 6 + 2 * pr) * 5 + h;
          float2& Gj = G[u][j / 5][j % 5];
          Gj.x = mine ? __ldg(dOut + gO[u] + static_cast<long long>(r0) * gN[u]) : 0.f;
          Gj.y = mine ? __ldg(dOut + gO[u] + static_cast<long long>(r0 + 5) * gN[u]) : 0.f;
          if (mine) {
            DU[r0 * Wp + sc[u]] = Gj.x;
            DU[(r0 + 5) * Wp + sc[u]] = Gj.y;
          }
        }
        continue;
      }
#pragma unroll
      for (int co = 0; co < 6; ++co)
#pragma unroll
        for (int h = 0; h < 5; ++h) {
          const int row = (half * 6 + co) * 5 + h;
          const float g = on[u] ? __ldg(dOut + gO[u] + static_cast<long long>(row) * gN[u]) : 0.f;
          (co & 1 ? G[u][co >> 1][h].y : G[u][co >> 1][h].x) = g;
        }
      if (on[u]) {
#pragma unroll
        for (int co = 0; co < 6; ++co)
#pragma unroll
          for (int h = 0; h < 5; ++h)
            DU[((half * 6 + co) * 5 + h) * Wp + sc[u]] = (co & 1) ? G[u][co >> 1][h].y : G[u][co >> 1][h].x;
      }
    }
    // activation planes -> XT: `rows` rows of a [rows][N] block per scene, as 4-byte cp.async: the copies
    // of layer l-1's input run underneath layer l's transposed convolution (which reads DU only)
    // thread = (column q, row phase): a row of the tile takes 32/64/128/256 lanes (the smallest that covers
    // `cols`), so no index division and the per-copy address is one add
    const int lpr_log = cols <= 32 ? 5 : (cols <= 64 ? 6 : (cols <= 128 ? 7 : 8));
    const int fq = tid & ((1 << lpr_log) - 1), frow0 = tid >> lpr_log, frstep = X_THREADS >> lpr_log;
    const bool fon = fq < cols;
    const int fdst = fon ? sm.scol[fq] : 0, fns = fon ? sm.nstr[fq] : 0, fsb = fon ? sm.sbase[fq] : 0;
    auto fill_xt_async = [&](const float* __restrict__ xb, int rows) {
      if (!fon) return;
      const float* src = xb + fsb * rows + (fq - fsb) + frow0 * fns;
      float* dst = XT + frow0 * Wp + fdst;
      for (int row = frow0; row < rows; row += frstep, src += frstep * fns, dst += frstep * Wp) cp_async4(dst, src);
    };
    fill_xt_async(Xsave + 3 * layer_stride + static_cast<long long>(p_lo) * 60, 60);  // x4
    cp_async_wait_all();
    __syncthreads();
    PHASE(2);

#pragma unroll 1
    for (int l = 4; l >= 0; --l) {
      const int cin = (l == 0) ? 8 : 12;
      PHASE(3 + (4 - l) * 6);
      // ---- weight and bias gradients of layer l
      float* gw = gacc + (w_base(l) - P_TP0_W);
      if (cgs > 1) {
        dw_pair_slices(DU, XT, Wp, c_end, cin, warp, lane, gw, gacc + (b_base(l) - P_TP0_W));
      } else {
        for (int unit = warp; unit < 2 * cin; unit += 8) dw_unit(DU, XT, Wp, c_end, unit >> 1, unit & 1, cin, lane, gw);
      }
      for (int co = warp; co < 12 && cgs == 1; co += 8) {
        float sb_ = 0.f;
        for (int c = 1 + lane; c <= c_end; c += 32)
#pragma unroll
          for (int h = 0; h < 5; ++h) sb_ += DU[(co * 5 + h) * Wp + c];
        sb_ = warp_sum(sb_);
        if (lane == 0) gacc[b_base(l) - P_TP0_W + co] += sb_;
      }
      PHASE(4 + (4 - l) * 6);
      if (l > 0) {
        __syncthreads();  // every read of XT (= x_l) is done: the next layer's input may land
        if (l >= 2)
          fill_xt_async(Xsave + (l - 2) * layer_stride + static_cast<long long>(p_lo) * 60, 60);
        else
          fill_xt_async(Hin + static_cast<long long>(p_lo) * 40, 40);
      }
      // ---- transposed convolution: T = conv_l^T(du_l)
      float2 T[JX][3][5];
#pragma unroll
      for (int u = 0; u < JX; ++u) {
#pragma unroll
        for (int pr = 0; pr < 3; ++pr)
#pragma unroll
          for (int h = 0; h < 5; ++h) T[u][pr][h] = make_float2(0.f, 0.f);
        if (cgs == 1) {
          if (on[u]) conv_column2<12>(DU, Wp, sc[u], sm.wT + l * X_WL, half, T[u]);
        } else {
          // a slice of the 12 channels per lane, then a reduce-scatter over the 2 / 4 lanes of the column: at each
          // step a lane hands over the half of the entries its partner keeps; the sums land in T's first slots
          if (on[u]) {
            if (cgs == 4) conv_column2<3>(DU, Wp, sc[u], sm.wT + l * X_WL, half, T[u], cg * 3);
            else conv_column2<6>(DU, Wp, sc[u], sm.wT + l * X_WL, half, T[u], cg * 6);
          }
          float2 f[16];
#pragma unroll
          for (int e = 0; e < 15; ++e) f[e] = T[u][e / 5][e % 5];
          f[15] = make_float2(0.f, 0.f);
          const bool up1 = (cg & 1) != 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 send = up1 ? f[j] : f[j + 8], keep = up1 ? f[j + 8] : f[j];
            f[j].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 1);
            f[j].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 1);
          }
          if (cgs == 4) {
            const bool up2 = (cg & 2) != 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 send = up2 ? f[j] : f[j + 4], keep = up2 ? f[j + 4] : f[j];
              f[j].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 2);
              f[j].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 2);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) T[u][j / 5][j % 5] = f[j];
        }
      }
      PHASE(5 + (4 - l) * 6);
      __syncthreads();  // every read of DU for layer l is done
      PHASE(6 + (4 - l) * 6);
      if (JX == 1 && cgs != 1) {
        // ---- split column, owner side: entry e = own_base + j lives in slot j of T and G
        const int kk = l - 1;
        const float slope_k = l > 0 ? sm.slope[kk] : 0.f;
        float2 uvo[8];
        if (l > 0) {  // every load before its first use
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int e = own_base + j;
            const bool mine = j < own_n && e < 15 && on[0];
            const int pr = e / 5, h = e - pr * 5;
            const float* up = Usave + kk * layer_stride + gO[0] + static_cast<long long>((half * 6 + 2 * pr) * 5 + h) * gN[0];
            uvo[j].x = mine ? __ldg(up) : 1.f;
            uvo[j].y = mine ? __ldg(up + 5 * gN[0]) : 1.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int e = own_base + j;
          const bool mine = j < own_n && e < 15 && on[0];
          const int pr = e / 5, h = e - pr * 5;
          const int r0 = (half * 6 + 2 * pr) * 5 + h;
          const float2 Tv = T[0][j / 5][j % 5];
          if (l == 0) {
            if (mine) {
              const int q = qt;
              const int sb = sm.sbase[q];
              float* o = dH + static_cast<long long>(p_lo + sb) * 40 + (q - sb);
              if (r0 < 40) o[static_cast<long long>(r0) * gN[0]] = Tv.x;
              if (r0 + 5 < 40) o[static_cast<long long>(r0 + 5) * gN[0]] = Tv.y;
            }
            continue;
          }
          float2& Gj = G[0][j / 5][j % 5];
          const float gx = (l == 4) ? Tv.x : Gj.x + Tv.x, gy = (l == 4) ? Tv.y : Gj.y + Tv.y;
          Gj = make_float2(gx, gy);
          if (mine) {
            const bool px = uvo[j].x > 0.f, py = uvo[j].y > 0.f;
            if (!px) dsl[kk] = fmaf(gx, uvo[j].x, dsl[kk]);
            if (!py) dsl[kk] = fmaf(gy, uvo[j].y, dsl[kk]);
            DU[r0 * Wp + sc[0]] = px ? gx : slope_k * gx;
            DU[(r0 + 5) * Wp + sc[0]] = py ? gy : slope_k * gy;
          }
        }
        if (l == 0) break;
        PHASE(7 + (4 - l) * 6);
        cp_async_wait_all();
        PHASE(8 + (4 - l) * 6);
        __syncthreads();
        continue;
      }
      if (l == 0) {
#pragma unroll
        for (int u = 0; u < JX; ++u)
          if (on[u] && cg == 0) {
            const int q = qt + u * X_COLS;
            const int sb = sm.sbase[q];
            float* o = dH + static_cast<long long>(p_lo + sb) * 40 + (q - sb);
#pragma unroll
            for (int co = 0; co < 6; ++co)
#pragma unroll
              for (int h = 0; h < 5; ++h) {
                const int row = (half * 6 + co) * 5 + h;
                if (row < 40) o[static_cast<long long>(row) * gN[u]] = (co & 1) ? T[u][co >> 1][h].y : T[u][co >> 1][h].x;
              }
          }
        break;
      }
      const int k = l - 1;  // next layer down: x_{k+1} = PReLU_k(u_k) (+ x_k for k >= 1)
      const float slope = sm.slope[k];
      const float cnt = cg == 0 ? 1.f : 0.f;  // the lanes sharing a column hold the same values: one of them counts
#pragma unroll
      for (int u = 0; u < JX; ++u) {
        float uv[6][5];  // all 30 loads are issued before the first use
#pragma unroll
        for (int co = 0; co < 6; ++co)
#pragma unroll
          for (int h = 0; h < 5; ++h)
            uv[co][h] = on[u] ? __ldg(Usave + k * layer_stride + gO[u] +
                                      static_cast<long long>((half * 6 + co) * 5 + h) * gN[u])
                              : 1.f;
#pragma unroll
        for (int co = 0; co < 6; ++co)
#pragma unroll
          for (int h = 0; h < 5; ++h) {
            float& Gr = (co & 1) ? G[u][co >> 1][h].y : G[u][co >> 1][h].x;
            const float Tv = (co & 1) ? T[u][co >> 1][h].y : T[u][co >> 1][h].x;
            const float g = (l == 4) ? Tv : Gr + Tv;
            Gr = g;
            if (on[u]) {
              const int row = (half * 6 + co) * 5 + h;
              const bool pos = uv[co][h] > 0.f;
              if (!pos) dsl[k] = fmaf(g * cnt, uv[co][h], dsl[k]);
              DU[row * Wp + sc[u]] = pos ? g : slope * g;
            }
          }
      }
      PHASE(7 + (4 - l) * 6);
      cp_async_wait_all();
      PHASE(8 + (4 - l) * 6);
      __syncthreads();
    }
  }
#ifdef SSTG_PHASE_CLOCKS
  if (threadIdx.x == 0) g_phase[blockIdx.x * 40 + 38] = clock64();
#endif

  // ---- PReLU slope gradients: CTA reduction in warp order, then the partial row
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float s = warp_sum(dsl[k]);
    if (lane == 0) sm.red[warp][k] = s;
  }
  __syncthreads();
  if (tid < 4) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += sm.red[w][tid];
    gacc[P_PRELUS - P_TP0_W + tid] = s;
  }
  __syncthreads();
  float* out = partial + static_cast<size_t>(blockIdx.x) * B_GRAD_STRIDE;
  for (int i = tid; i < B_GRAD_STRIDE; i += X_THREADS) out[i] = gacc[i];
}

}  // namespace

#ifdef SSTG_PHASE_CLOCKS
extern "C" int sstg_debug_phase_clocks(long long* host_out, int n) {
  return cudaMemcpyFromSymbol(host_out, g_phase, sizeof(long long) * (n < 160 * 40 ? n : 160 * 40)) == cudaSuccess ? 0 : -3;
}
#endif

int txpcnn_bwd_rows(int n_tiles) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return n_tiles < sms ? n_tiles : sms;
}

int launch_txpcnn_backward(const float* dOut, const float* Usave, const float* Xsave, long long layer_stride,
                           const float* H, const int32_t* seg_off, const int32_t* tile_off, int n_tiles, int max_n,
                           const float* params, float* dH, float* partial, cudaStream_t st) {
  if (n_tiles == 0) return SSTG_OK;
  const int Wp = txpcnn_tile_width(max_n);
  const size_t smem = ((sizeof(BSmem) + 15) & ~size_t(15)) + sizeof(float) * (2 * X_ROWS * Wp + B_GRAD_STRIDE);
  const int grid = txpcnn_bwd_rows(n_tiles);
  auto launch = [&](auto kern) -> int {
    SSTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                  "txpcnn_bwd smem attr");
    SSTG_CUDA_TRY(launch_pdl(kern, grid, X_THREADS, smem, st, dOut, Usave, Xsave, layer_stride, H, seg_off, tile_off,
                             n_tiles, Wp, params, dH, partial),
                  "txpcnn_bwd launch");
    return SSTG_OK;
  };
  if (max_n > X_COLS) return launch(txpcnn_bwd_kernel<2>);
  return launch(txpcnn_bwd_kernel<1>);
}

}  // namespace sstg
