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
//   * lanes own output columns w; each A element is read from shared memory exactly once
//     and feeds 5 FMAs (one per channel) against a broadcast g[., t, v].
//   * BN (eval) is folded into scale/offset pairs once per CTA; the temporal conv needs the
//     neighbouring timesteps, i.e. other warps' results, through a double-buffered
//     shared tile and ONE __syncthreads per scene.
#include <cstdlib>

#include "sstg_common.cuh"

namespace sstg {

namespace {

constexpr int S_WARPS = T_OBS;  // one warp per observed timestep
constexpr int S_THREADS = S_WARPS * 32;
constexpr int S_MAX_STAGES = 4;
constexpr int S_PAD = 8;  // floats of slack per stage for the aligned-superset copies

struct FoldedParams {
  float gw[C][2], gb[C];
  float s1[C], o1[C];
  float w2[C][C][3], b2[C];  // BN2 folded in
  float rw[C][2], rb[C];     // BN_res folded in
  float a1, a2;
};

struct SmemLayout {
  int NP;           // padded max pedestrians
  int chunk_floats; // payload floats per ring stage
  int p_bufs;       // 1 or 2
  size_t off_bar, off_g, off_p, off_ring, total;
};

__host__ __device__ inline SmemLayout make_layout(int max_n, size_t budget, int stages) {
  SmemLayout L;
  L.NP = (max_n + 3) & ~3;
  size_t o = (sizeof(FoldedParams) + 15) & ~size_t(15);
  L.off_bar = o;
  o += sizeof(uint64_t) * S_WARPS * S_MAX_STAGES;
  L.off_g = o;
  o += sizeof(float) * S_WARPS * L.NP * 8;
  L.off_p = o;
  L.p_bufs = (L.NP <= 128) ? 2 : 1;
  o += sizeof(float) * L.p_bufs * C * T_OBS * L.NP;
  L.off_ring = o;
  size_t per_stage = (budget - o) / (S_WARPS * stages) / sizeof(float);
  per_stage &= ~size_t(3);
  long cf = static_cast<long>(per_stage) - S_PAD;
  if (cf > 4096) cf = 4096;  // 16 KB copies are plenty
  L.chunk_floats = static_cast<int>(cf);
  L.total = o + sizeof(float) * S_WARPS * stages * (L.chunk_floats + S_PAD);
  return L;
}

// cursor over one warp's sequence of (scene, row-chunk) copies
struct Cursor {
  int s;            // scene index (>= S: exhausted)
  int N;            // pedestrians of that scene
  int r0;           // next row of the slice
  int R;            // rows per chunk for this scene
  const float* sl;  // base of this warp's A[t] slice
};

__device__ __forceinline__ void cursor_open(Cursor& c, int s, int S, int t, int chunk_floats,
                                            const int32_t* __restrict__ seg_off,
                                            const int64_t* __restrict__ a_off, const float* __restrict__ A) {
  c.s = s;
  c.r0 = 0;
  if (s < S) {
    c.N = seg_off[s + 1] - seg_off[s];
    c.R = chunk_floats / c.N;
    if (c.R > c.N) c.R = c.N;
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
  float* gbuf = reinterpret_cast<float*>(smem + L.off_g) + static_cast<size_t>(t) * NP * 8;  // [v][8]
  float* pbuf = reinterpret_cast<float*>(smem + L.off_p);                                    // [buf][c][t][NP]
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
      for (int k = 0; k < 3; ++k) fp.w2[c][ci][k] = params[P_TCN_W + (c * C + ci) * 3 + k] * s2;
    fp.b2[c] = (params[P_TCN_B + c] - bn[10 + c]) * s2 + params[P_BN2_B + c];
    const float sr = params[P_BNR_G + c] / sqrtf(bn[25 + c] + BN_EPS);
    fp.rw[c][0] = params[P_RES_W + c * 2 + 0] * sr;
    fp.rw[c][1] = params[P_RES_W + c * 2 + 1] * sr;
    fp.rb[c] = (params[P_RES_B + c] - bn[20 + c]) * sr + params[P_BNR_B + c];
    fp.gw[c][0] = params[P_GCN_W + c * 2 + 0];
    fp.gw[c][1] = params[P_GCN_W + c * 2 + 1];
    fp.gb[c] = params[P_GCN_B + c];
    if (c == 0) {
      fp.a1 = params[P_PRELU1];
      fp.a2 = params[P_PRELU2];
    }
  }
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
  cursor_open(pc, blockIdx.x, S, t, L.chunk_floats, seg_off, a_off, A);
  auto produce = [&](int stage) {
    if (pc.s >= S) return;
    int rows = pc.N - pc.r0;
    if (rows > pc.R) rows = pc.R;
    const float* src = pc.sl + static_cast<size_t>(pc.r0) * pc.N;
    const uintptr_t b = reinterpret_cast<uintptr_t>(src);
    const uintptr_t b0 = b & ~uintptr_t(15);
    const uintptr_t e1 = (b + static_cast<uintptr_t>(rows) * pc.N * 4 + 15) & ~uintptr_t(15);
    if (lane == 0) {
      const uint32_t bytes = static_cast<uint32_t>(e1 - b0);
      mbar_arrive_expect_tx(&full[stage], bytes);
      bulk_g2s_hint(ring + stage * stage_floats, reinterpret_cast<const void*>(b0), bytes, &full[stage], l2_pol);
    }
    pc.r0 += rows;
    if (pc.r0 >= pc.N) cursor_open(pc, pc.s + gridDim.x, S, t, L.chunk_floats, seg_off, a_off, A);
  };
#pragma unroll
  for (int i = 0; i < S_STAGES; ++i) produce(i);

  uint32_t consumed = 0;  // chunks consumed by this warp so far (stage = consumed % STAGES)
  // V[t] of the first scene
  float vx[NJ], vy[NJ];
#ifndef SSTG_STGCN_NO_PREFETCH_OFFSETS
  // The next scene's offsets travel with its V prefetch, so the scene top does not wait for a dependent global
  // load -> use chain (6 % of the stall samples sat on its first consumer, profiles/r1t_ncu_full_summary.csv):
  // 3.44 -> 3.33 ms per 131,072 scenes of 64, bit-identical (profiles/r2a_ab_stgcn_po.txt).
  // -DSSTG_STGCN_NO_PREFETCH_OFFSETS restores the in-loop loads for A/B runs.
  int nx_n0 = 0, nx_N = 1;
  long long nx_ao = 0;
#endif
  auto load_v = [&](int s) {
    if (s < S) {
      const int n0 = seg_off[s], N = seg_off[s + 1] - n0;
#ifndef SSTG_STGCN_NO_PREFETCH_OFFSETS
      nx_n0 = n0;
      nx_N = N;
      nx_ao = a_off[s];
#endif
      const float2* vp = reinterpret_cast<const float2*>(V) + (static_cast<size_t>(n0) * T_OBS + static_cast<size_t>(t) * N);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        const float2 v = k < N ? vp[k] : make_float2(0.f, 0.f);
        vx[j] = v.x;
        vy[j] = v.y;
      }
    }
  };
  load_v(blockIdx.x);

  int it = 0;
  for (int s = blockIdx.x; s < S; s += gridDim.x, ++it) {
#ifndef SSTG_STGCN_NO_PREFETCH_OFFSETS
    const int n0 = nx_n0;
    const int N = nx_N;
    const float* sl = A + nx_ao + static_cast<size_t>(t) * N * N;
#else
    const int n0 = seg_off[s];
    const int N = seg_off[s + 1] - n0;
    const float* sl = A + a_off[s] + static_cast<size_t>(t) * N * N;
#endif
    int R = L.chunk_floats / N;
    if (R > N) R = N;

    // ---- phase 1: g for this warp's timestep, columns owned by the lane
    float rx[NJ], ry[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = lane + 32 * j;
      rx[j] = vx[j];
      ry[j] = vy[j];
      if (k < N) {
        float4 g4;
        g4.x = fmaf(fp.gw[0][1], vy[j], fmaf(fp.gw[0][0], vx[j], fp.gb[0]));
        g4.y = fmaf(fp.gw[1][1], vy[j], fmaf(fp.gw[1][0], vx[j], fp.gb[1]));
        g4.z = fmaf(fp.gw[2][1], vy[j], fmaf(fp.gw[2][0], vx[j], fp.gb[2]));
        g4.w = fmaf(fp.gw[3][1], vy[j], fmaf(fp.gw[3][0], vx[j], fp.gb[3]));
        *reinterpret_cast<float4*>(gbuf + k * 8) = g4;
        gbuf[k * 8 + 4] = fmaf(fp.gw[4][1], vy[j], fmaf(fp.gw[4][0], vx[j], fp.gb[4]));
      }
    }
    __syncwarp();
    load_v(s + gridDim.x);  // register prefetch of the next scene's V[t]

    // ---- phase 2: y[c, t, w] = sum_v g[c, t, v] A[t, v, w], streamed through the ring
    float acc[C][NJ];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[c][j] = 0.f;

    for (int r0 = 0; r0 < N; r0 += R) {
      const int rows = (N - r0 < R) ? (N - r0) : R;
      const int stage = consumed % S_STAGES;
      const uint32_t parity = (consumed / S_STAGES) & 1;
      const float* src = sl + static_cast<size_t>(r0) * N;
      const int off = static_cast<int>((reinterpret_cast<uintptr_t>(src) & 15) >> 2);
      const float* tile = ring + stage * stage_floats + off;
      mbar_wait(&full[stage], parity);
#pragma unroll 4
      for (int r = 0; r < rows; ++r) {
        const float* g = gbuf + (r0 + r) * 8;
        const float4 g4 = *reinterpret_cast<const float4*>(g);
        const float g5 = g[4];
        const float* row = tile + r * N;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int k = lane + 32 * j;
          const float a = k < N ? row[k] : 0.f;
          acc[0][j] = fmaf(g4.x, a, acc[0][j]);
          acc[1][j] = fmaf(g4.y, a, acc[1][j]);
          acc[2][j] = fmaf(g4.z, a, acc[2][j]);
          acc[3][j] = fmaf(g4.w, a, acc[3][j]);
          acc[4][j] = fmaf(g5, a, acc[4][j]);
        }
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
        if (k < N) p[(c * T_OBS + t) * NP + k] = prelu(fmaf(acc[c][j], fp.s1[c], fp.o1[c]), fp.a1);
      }
    __syncthreads();

    // ---- phase 3: (3,1) temporal conv (+BN2) + residual branch (+BN_res) + PReLU2
    float* Hs = H + static_cast<size_t>(n0) * (C * T_OBS);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = lane + 32 * j;
      if (k < N) {
        float z[C];
#pragma unroll
        for (int co = 0; co < C; ++co) z[co] = fp.b2[co];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          const int tt = t + kt - 1;
          if (tt >= 0 && tt < T_OBS) {
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
              const float pv = p[(ci * T_OBS + tt) * NP + k];
#pragma unroll
              for (int co = 0; co < C; ++co) z[co] = fmaf(fp.w2[co][ci][kt], pv, z[co]);
            }
          }
        }
#pragma unroll
        for (int co = 0; co < C; ++co) {
          const float res = fmaf(fp.rw[co][1], ry[j], fmaf(fp.rw[co][0], rx[j], fp.rb[co]));
          Hs[(co * T_OBS + t) * N + k] = prelu(z[co] + res, fp.a2);
        }
      }
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
  if (L.chunk_floats < L.NP) return SSTG_ETOOBIG;
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
