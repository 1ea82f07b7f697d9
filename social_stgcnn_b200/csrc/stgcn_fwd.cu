// Kernel (2): the whole st_gcn block of reference model.py:145-155, fused, eval mode.
//
//   g[c,t,v]  = Wg[c,:] . v[t,v,:] + bg[c]                          1x1 conv      (model.py:66)
//   y[c,t,w]  = sum_v g[c,t,v] * A[t,v,w]                          einsum        (model.py:67)
//   p         = PReLU1(BN1(y))                                                    (model.py:114-115)
//   z[c,t,w]  = b2[c] + sum_ci sum_kt W2[c,ci,kt] p[ci,t+kt-1,w]   (3,1) conv    (model.py:116-122)
//   h         = PReLU2(BN2(z) + BNr(Wr . v + br))                                 (model.py:123,147-153)
//
// The kernel is HBM-read bound: per scene it must read the (8,N,N) adjacency once
// (32 N^2 bytes) against ~90 N^2 flops.  Layout of the work:
//   * persistent CTAs (one per SM) walk scenes s = blockIdx.x, +gridDim.x, ...
//   * warp t of the CTA owns timestep t.  It streams its own A[t] slice through a private
//     shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//     lane 0 re-arms a stage as soon as the warp has consumed it, so the ring keeps
//     prefetching ACROSS scene boundaries and the copy engine never waits on the math.
//     Because a ragged packed batch cannot guarantee 16-byte aligned rows, each copy is
//     the 16-byte-aligned superset of the rows it needs and the consumer indexes into it
//     with the residual offset.
//   * lanes own output columns w = lane + 32 j.  The aggregation is issue-limited, not FMA-limited
//     (profiles/r2d: 19.5 instructions per 64-element row, 62 % of the issue slots, 2,460 instructions per
//     warp and scene of which only 1,090 were the row loop), so every FMA is a packed FFMA2 over a PAIR OF
//     ROWS: accumulator (sum over even rows, sum over odd rows) += (A[v][w], A[v+1][w]) * (g[c][v], g[c][v+1]).
//     The pair of A values is two LDS.32 into adjacent registers (any N, no alignment requirement), the g
//     pairs of FOUR rows of a channel are one broadcast LDS.128 out of a [row quad][channel][4] tile, and
//     a chunk always holds a multiple of four rows: 8.25 + ~2 instructions per 64-element row.
//   * BN (eval) is folded into scale/offset pairs once per CTA; the temporal conv needs the
//     neighbouring timesteps, i.e. other warps' results, through a double-buffered
//     shared tile and ONE __syncthreads per scene; it runs as FFMA2 over the lane's column pair
//     (lane, lane + 32) against (w, w) pairs kept duplicated in shared memory.
#include <cstdlib>

#include "sstg_common.cuh"

namespace sstg {

namespace {

constexpr int S_WARPS = T_OBS;  // one warp per observed timestep
constexpr int S_THREADS = S_WARPS * 32;
constexpr int S_MAX_STAGES = 4;
constexpr int S_PAD = 8;       // floats of slack per stage for the aligned-superset copies
constexpr int G_QUAD = 4 * C;  // floats per row quad of the g tile: [channel][4 rows]
#ifndef SSTG_STGCN_UNROLL
#define SSTG_STGCN_UNROLL 2
#endif
constexpr int S_UNROLL = SSTG_STGCN_UNROLL;  // row quads per trip of the aggregation loop (A/B knob: 1 / 2 / 4)

// packed FP32 (sm_100 FFMA2): d.xy = a.xy * b.xy + d.xy, two IEEE fp32 FMAs in one issue slot
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

struct alignas(16) FoldedParams {
  float2 w2d[C][3][6];  // [ci][kt][co] = (w, w) of the temporal conv with BN2 folded in; co padded to 6 (48-byte rows)
  float2 b2d[6];        // (b, b)
  float2 rwd[2][6];     // [x|y][co] = (w, w) of the residual 1x1 conv with BN_res folded in
  float2 rbd[6];
  float gw[C][2], gb[C];
  float s1[C], o1[C];
  float a1, a2;
};

struct SmemLayout {
  int NP;           // padded max pedestrians
  int chunk_floats; // payload floats per ring stage
  int p_bufs;       // 1 or 2
  size_t off_bar, off_rtab, off_g, off_p, off_ring, total;
};

__host__ __device__ inline SmemLayout make_layout(int max_n, size_t budget, int stages) {
  SmemLayout L;
  L.NP = (max_n + 3) & ~3;
  size_t o = (sizeof(FoldedParams) + 15) & ~size_t(15);
  L.off_bar = o;
  o += sizeof(uint64_t) * S_WARPS * S_MAX_STAGES;
  L.off_rtab = o;
  o += (sizeof(uint16_t) * (L.NP + 1) + 15) & ~size_t(15);
  L.off_g = o;
  o += sizeof(float) * S_WARPS * (L.NP / 4) * G_QUAD;
  L.off_p = o;
  L.p_bufs = (L.NP <= 128) ? 2 : 1;
  o += sizeof(float) * L.p_bufs * C * T_OBS * L.NP;
  L.off_ring = o;
  // the ring is followed by 32 * 8 floats of slack: lanes whose column lies beyond N read (and ignore) whatever
  // follows their row inside the allocation
  const size_t slack = sizeof(float) * 32 * 8;
  size_t per_stage = (budget - o - slack) / (S_WARPS * stages) / sizeof(float);
  per_stage &= ~size_t(3);
  long cf = static_cast<long>(per_stage) - S_PAD;
  if (cf > 4096) cf = 4096;  // 16 KB copies are plenty
  L.chunk_floats = static_cast<int>(cf);
  L.total = o + sizeof(float) * S_WARPS * stages * (L.chunk_floats + S_PAD) + slack;
  return L;
}

// rows of an N x N slice per ring stage: the slice is cut into the fewest chunks that fit a stage, of EQUAL
// size rounded up to a multiple of four rows (N = 64 with room for 19 rows: 4 x 16, not 20 + 20 + 20 + 4)
__host__ __device__ inline int rows_per_chunk(int chunk_floats, int N) {
  const int rmax = chunk_floats / N;            // >= 4: checked at launch
  if (rmax >= N) return N;
  const int r4 = rmax & ~3;
#ifdef SSTG_STGCN_UNBALANCED
  return r4;
#endif
  const int n = (N + r4 - 1) / r4;              // chunks per slice
  return (((N + n - 1) / n) + 3) & ~3;          // <= r4 because r4 is a multiple of four
}

// cursor over one warp's sequence of (scene, row-chunk) copies
struct Cursor {
  int s;            // scene index (>= S: exhausted)
  int N;            // pedestrians of that scene
  int r0;           // next row of the slice
  int R;            // rows per chunk for this scene
  const float* sl;  // base of this warp's A[t] slice
};

__device__ __forceinline__ void cursor_open(Cursor& c, int s, int S, int t, const uint16_t* __restrict__ rtab,
                                            const int32_t* __restrict__ seg_off,
                                            const int64_t* __restrict__ a_off, const float* __restrict__ A) {
  c.s = s;
  c.r0 = 0;
  if (s < S) {
    c.N = seg_off[s + 1] - seg_off[s];
    c.R = rtab[c.N];
    c.sl = A + a_off[s] + static_cast<size_t>(t) * c.N * c.N;
  } else {
    c.N = 1;
    c.R = 1;
    c.sl = A;
  }
}

// CPS = CTAs per SM the kernel is compiled for (register cap 255 / 128): two co-resident CTAs
// (16 warps) hide the shared-memory and mbarrier latencies of each other's math phases.
template <int NJ, int CPS, int S_STAGES>
__global__ void __launch_bounds__(S_THREADS, CPS)
stgcn_fwd_kernel(const float* __restrict__ V, const float* __restrict__ A, const int32_t* __restrict__ seg_off,
                 const int64_t* __restrict__ a_off, int S, int max_n, const float* __restrict__ params,
                 const float* __restrict__ bn, float* __restrict__ H, int smem_budget) {
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemLayout L = make_layout(max_n, static_cast<size_t>(smem_budget), S_STAGES);
  FoldedParams& fp = *reinterpret_cast<FoldedParams*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, t = tid >> 5;
  const int NP = L.NP;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar) + t * S_MAX_STAGES;
  uint16_t* rtab = reinterpret_cast<uint16_t*>(smem + L.off_rtab);                                  // [N] -> rows per chunk
  float* gq = reinterpret_cast<float*>(smem + L.off_g) + static_cast<size_t>(t) * (NP / 4) * G_QUAD;  // [v/4][c][4]
  float* pbuf = reinterpret_cast<float*>(smem + L.off_p);                                           // [buf][c][t][NP]
  const int stage_floats = L.chunk_floats + S_PAD;
  float* ring = reinterpret_cast<float*>(smem + L.off_ring) + static_cast<size_t>(t) * S_STAGES * stage_floats;

  // ---- fold parameters (BatchNorm in eval form: y = (x - rm) / sqrt(rv + eps) * gamma + beta)
  if (tid < C) {
    const int c = tid;
    const float s1 = params[P_BN1_G + c] / sqrtf(bn[5 + c] + BN_EPS);
    fp.s1[c] = s1;
    fp.o1[c] = params[P_BN1_B + c] - bn[0 + c] * s1;
    const float s2 = params[P_BN2_G + c] / sqrtf(bn[15 + c] + BN_EPS);
    for (int ci = 0; ci < C; ++ci)
      for (int k = 0; k < 3; ++k) {
        const float w = params[P_TCN_W + (c * C + ci) * 3 + k] * s2;
        fp.w2d[ci][k][c] = make_float2(w, w);
      }
    const float b2 = (params[P_TCN_B + c] - bn[10 + c]) * s2 + params[P_BN2_B + c];
    fp.b2d[c] = make_float2(b2, b2);
    const float sr = params[P_BNR_G + c] / sqrtf(bn[25 + c] + BN_EPS);
    const float r0 = params[P_RES_W + c * 2 + 0] * sr, r1 = params[P_RES_W + c * 2 + 1] * sr;
    fp.rwd[0][c] = make_float2(r0, r0);
    fp.rwd[1][c] = make_float2(r1, r1);
    const float rb = (params[P_RES_B + c] - bn[20 + c]) * sr + params[P_BNR_B + c];
    fp.rbd[c] = make_float2(rb, rb);
    fp.gw[c][0] = params[P_GCN_W + c * 2 + 0];
    fp.gw[c][1] = params[P_GCN_W + c * 2 + 1];
    fp.gb[c] = params[P_GCN_B + c];
    if (c == 0) {
      fp.a1 = params[P_PRELU1];
      fp.a2 = params[P_PRELU2];
    }
  }
  for (int n = tid + 1; n <= NP; n += S_THREADS) rtab[n] = static_cast<uint16_t>(rows_per_chunk(L.chunk_floats, n));
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < S_STAGES; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // ---- producer prologue: fill this warp's ring
  // A is read exactly once: an L2 evict-first policy on the bulk copies keeps the stream from displacing V, the
  // offsets and H's write-back lines (3.49 -> 3.43 ms per 131,072 scenes of 64)
  const uint64_t l2_pol = l2_evict_first_policy();
  Cursor pc;
  cursor_open(pc, blockIdx.x, S, t, rtab, seg_off, a_off, A);
  auto produce = [&](int stage) {
    if (pc.s >= S) return;
    int rows = pc.N - pc.r0;
    if (rows > pc.R) rows = pc.R;
    const float* src = pc.sl + pc.r0 * pc.N;
    const uintptr_t b = reinterpret_cast<uintptr_t>(src);
    const uintptr_t b0 = b & ~uintptr_t(15);
    const uint32_t bytes = (static_cast<uint32_t>(b - b0) + static_cast<uint32_t>(rows * pc.N) * 4u + 15u) & ~15u;
    if (lane == 0) {
      mbar_arrive_expect_tx(&full[stage], bytes);
      bulk_g2s_hint(ring + stage * stage_floats, reinterpret_cast<const void*>(b0), bytes, &full[stage], l2_pol);
    }
    pc.r0 += rows;
    if (pc.r0 >= pc.N) cursor_open(pc, pc.s + gridDim.x, S, t, rtab, seg_off, a_off, A);
  };
#pragma unroll
  for (int i = 0; i < S_STAGES; ++i) produce(i);

  uint32_t consumed = 0;  // chunks consumed by this warp so far (stage = consumed % STAGES)
  // V[t] of the first scene, as (column 2 i, column 2 i + 1) pairs of the lane's columns lane + 32 j.  The next
  // scene's offsets travel with its V prefetch, so the scene top does not wait for a dependent global load ->
  // use chain (3.44 -> 3.33 ms per 131,072 scenes of 64, profiles/r2a_ab_stgcn_po.txt).
  constexpr int NP2 = (NJ + 1) / 2;
  float2 vx[NP2], vy[NP2];
  int nx_n0 = 0, nx_N = 1;
  long long nx_ao = 0;
  // (a two-level pipeline -- offsets two scenes ahead, V one scene ahead -- measured SLOWER at the 128-register
  //  cap of the two-CTA form: 3.00 vs 2.82 ms, profiles/r2g_ab_stgcn.txt)
  auto load_v = [&](int s) {
    if (s < S) {
      const int n0 = seg_off[s], N = seg_off[s + 1] - n0;
      nx_n0 = n0;
      nx_N = N;
      nx_ao = a_off[s];
      const float2* vp = reinterpret_cast<const float2*>(V) + (static_cast<size_t>(n0) * T_OBS + static_cast<size_t>(t) * N);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        const float2 v = k < N ? vp[k] : make_float2(0.f, 0.f);
        if (j & 1) {
          vx[j >> 1].y = v.x;
          vy[j >> 1].y = v.y;
        } else {
          vx[j >> 1].x = v.x;
          vy[j >> 1].x = v.y;
        }
      }
      if (NJ & 1) vx[NP2 - 1].y = vy[NP2 - 1].y = 0.f;
    }
  };
  load_v(blockIdx.x);

  int it = 0;
  for (int s = blockIdx.x; s < S; s += gridDim.x, ++it) {
    const int n0 = nx_n0;
    const int N = nx_N;
    const float* sl = A + nx_ao + static_cast<size_t>(t) * N * N;
    const int R = rtab[N];

    // ---- phase 1: g for this warp's timestep, rows (= nodes) owned by the lane, into the [quad][c][4] tile;
    // the rows that pad N to a multiple of four get g = 0
    float2 rx[NP2], ry[NP2];
#pragma unroll
    for (int i = 0; i < NP2; ++i) {
      rx[i] = vx[i];
      ry[i] = vy[i];
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = lane + 32 * j;
      const float x = (j & 1) ? vx[j >> 1].y : vx[j >> 1].x;
      const float y = (j & 1) ? vy[j >> 1].y : vy[j >> 1].x;
      if (k < ((N + 3) & ~3)) {
        float* g = gq + (k >> 2) * G_QUAD + (k & 3);
        const bool in = k < N;
#pragma unroll
        for (int c = 0; c < C; ++c) g[c * 4] = in ? fmaf(fp.gw[c][1], y, fmaf(fp.gw[c][0], x, fp.gb[c])) : 0.f;
      }
    }
    __syncwarp();
    load_v(s + gridDim.x);  // register prefetch of the next scene's V[t] and offsets

    // ---- phase 2: y[c, t, w] = sum_v g[c, t, v] A[t, v, w], streamed through the ring
    float2 acc[C][NJ];  // (sum over even rows, sum over odd rows) of column lane + 32 j
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[c][j] = make_float2(0.f, 0.f);

    for (int r0 = 0; r0 < N; r0 += R) {
      const int rows = (N - r0 < R) ? (N - r0) : R;
      const int stage = consumed % S_STAGES;
      const uint32_t parity = (consumed / S_STAGES) & 1;
      const float* src = sl + r0 * N;
      const int off = static_cast<int>((reinterpret_cast<uintptr_t>(src) & 15) >> 2);
      const float* row = ring + stage * stage_floats + off + lane;   // column lane of the chunk's first row
      const float* g = gq + (r0 >> 2) * G_QUAD;                      // r0 is a multiple of four whenever rows >= 4
      mbar_wait(&full[stage], parity);
      int q = 0;
#pragma unroll S_UNROLL
      for (; q + 4 <= rows; q += 4) {
        float4 g4[C];
#pragma unroll
        for (int c = 0; c < C; ++c) g4[c] = *reinterpret_cast<const float4*>(g + c * 4);
        g += G_QUAD;
        float2 a01[NJ], a23[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {   // lanes beyond N read (finite or not) neighbours: their sums are never used
          a01[j] = make_float2(row[32 * j], row[N + 32 * j]);
          a23[j] = make_float2(row[2 * N + 32 * j], row[3 * N + 32 * j]);
        }
        row += 4 * N;
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int j = 0; j < NJ; ++j) ffma2(acc[c][j], a01[j], make_float2(g4[c].x, g4[c].y));
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int j = 0; j < NJ; ++j) ffma2(acc[c][j], a23[j], make_float2(g4[c].z, g4[c].w));
      }
      for (; q < rows; ++q) {  // N not a multiple of four: the last rows of the slice, one at a time
        const int v = r0 + q;
        const float* gs = gq + (v >> 2) * G_QUAD + (v & 3);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float a = row[32 * j];
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c][j].x = fmaf(gs[c * 4], a, acc[c][j].x);
        }
        row += N;
      }
      __syncwarp();
      produce(stage);
      ++consumed;
    }

    // ---- BN1 + PReLU1 -> shared p[c][t][w]
    float* p = pbuf + static_cast<size_t>(L.p_bufs == 2 ? (it & 1) : 0) * C * T_OBS * NP;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        if (k < N) p[(c * T_OBS + t) * NP + k] = prelu(fmaf(acc[c][j].x + acc[c][j].y, fp.s1[c], fp.o1[c]), fp.a1);
      }
    __syncthreads();

    // ---- phase 3: (3,1) temporal conv (+BN2) + residual branch (+BN_res) + PReLU2, two columns per FFMA2
    float* Hs = H + static_cast<size_t>(n0) * (C * T_OBS) + t * N + lane;
    float2 z[C][NP2];
#pragma unroll
    for (int co = 0; co < C; ++co)
#pragma unroll
      for (int i = 0; i < NP2; ++i) {
        z[co][i] = fp.rbd[co];
        ffma2(z[co][i], fp.rwd[0][co], rx[i]);
        ffma2(z[co][i], fp.rwd[1][co], ry[i]);
        z[co][i].x += fp.b2d[co].x;
        z[co][i].y += fp.b2d[co].y;
      }
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      const int tt = t + kt - 1;
      if (tt >= 0 && tt < T_OBS) {
        const float* pr = p + tt * NP + lane;
#pragma unroll
        for (int ci = 0; ci < C; ++ci) {
          float2 pv[NP2];
#pragma unroll
          for (int i = 0; i < NP2; ++i)   // the second column of an odd NJ lies beyond NP: re-read the first
            pv[i] = make_float2(pr[ci * T_OBS * NP + 64 * i], pr[ci * T_OBS * NP + ((2 * i + 1 < NJ) ? 64 * i + 32 : 64 * i)]);
#pragma unroll
          for (int co = 0; co < C; ++co)
#pragma unroll
            for (int i = 0; i < NP2; ++i) ffma2(z[co][i], fp.w2d[ci][kt][co], pv[i]);
        }
      }
    }
#pragma unroll
    for (int co = 0; co < C; ++co)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float h = (j & 1) ? z[co][j >> 1].y : z[co][j >> 1].x;
        if (lane + 32 * j < N) Hs[co * T_OBS * N + 32 * j] = prelu(h, fp.a2);
      }
    if (L.p_bufs == 1) __syncthreads();
  }
}

template <int NJ, int CPS, int STAGES>
int launch_nj(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off, int S, int max_n,
              const float* params, const float* bn, float* H, int sm_count, cudaStream_t st) {
  int dev = 0, sms = 148, max_optin = 0;
  SSTG_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
  SSTG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "sm count");
  SSTG_CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev), "smem optin");
  // 227 KB for one CTA per SM; two CTAs share the SM's 228 KB (1 KB reserved per CTA)
  const int budget = CPS == 1 ? max_optin : (max_optin + 1024) / 2 - 1024;
  const SmemLayout L = make_layout(max_n, static_cast<size_t>(budget), STAGES);
  if (L.chunk_floats < 4 * L.NP) return SSTG_ETOOBIG;   // a ring stage holds at least four rows
  auto kern = stgcn_fwd_kernel<NJ, CPS, STAGES>;
  SSTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total)),
                "stgcn smem attr");
  if (sm_count > 0 && sm_count < sms) sms = sm_count;
  const long long want = static_cast<long long>(sms) * CPS;
  const int grid = static_cast<int>(S < want ? S : want);
  kern<<<grid, S_THREADS, L.total, st>>>(V, A, seg_off, a_off, S, max_n, params, bn, H, budget);
  SSTG_CUDA_TRY(cudaGetLastError(), "stgcn launch");
  return SSTG_OK;
}

// SSTG_STGCN_CPS=1|2 overrides the CTAs-per-SM choice (tuning knob; default 2 for N <= 64).
// SSTG_STGCN_STAGES=2|3|4: ring depth per warp when two CTAs share an SM (tuning knob; default 2:
// measured 5.40 TB/s vs 4.92 (3) and 4.73 (4) at N = 64 -- fewer, larger bulk copies per slice).
int stgcn_stages_default() {
  const char* e = getenv("SSTG_STGCN_STAGES");
  if (e && e[0] >= '2' && e[0] <= '4') return e[0] - '0';
  return 2;
}
int stgcn_cps_default(int max_n) {
  const char* e = getenv("SSTG_STGCN_CPS");
  if (e && (e[0] == '1' || e[0] == '2')) return e[0] - '0';
  return max_n <= 64 ? 2 : 1;
}

}  // namespace

int launch_stgcn_forward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off, int S,
                         int max_n, const float* params, const float* bn, float* H, int sm_count,
                         cudaStream_t st) {
  if (S == 0) return SSTG_OK;
  const bool two = stgcn_cps_default(max_n) == 2 && max_n <= 64;
  const int stages = stgcn_stages_default();
#define SSTG_ARGS V, A, seg_off, a_off, S, max_n, params, bn, H, sm_count, st
  if (max_n <= 32) {
    if (!two) return launch_nj<1, 1, 4>(SSTG_ARGS);
    return stages == 2 ? launch_nj<1, 2, 2>(SSTG_ARGS) : stages == 3 ? launch_nj<1, 2, 3>(SSTG_ARGS)
                                                                     : launch_nj<1, 2, 4>(SSTG_ARGS);
  }
  if (max_n <= 64) {
    if (!two) return launch_nj<2, 1, 4>(SSTG_ARGS);
    return stages == 2 ? launch_nj<2, 2, 2>(SSTG_ARGS) : stages == 3 ? launch_nj<2, 2, 3>(SSTG_ARGS)
                                                                     : launch_nj<2, 2, 4>(SSTG_ARGS);
  }
  if (max_n <= 128) return launch_nj<4, 1, 4>(SSTG_ARGS);
  return launch_nj<8, 1, 4>(SSTG_ARGS);
#undef SSTG_ARGS
}

}  // namespace sstg
