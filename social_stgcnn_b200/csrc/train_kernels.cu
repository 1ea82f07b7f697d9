// Training-path kernels, part 1: st_gcn block in TRAIN mode (forward + backward), the
// bivariate NLL with its gradient, the BatchNorm running-statistics recurrence, the
// deterministic gradient reduction and the SGD update.
//
// Reference semantics (train.py:28-79, model.py:145-155, metrics.py:84-113):
//   * the reference runs batch 1, so BatchNorm2d in train mode normalises with the batch
//     statistics of ONE scene: per channel over its 8*N elements (biased variance),
//     eps 1e-5; running stats get momentum 0.1 and the unbiased variance, scene after
//     scene in loader order;
//   * loss of an optimiser group = (sum of the per-scene mean NLLs of all scenes but the
//     last one of the group) / batch_size; one backward; plain SGD.
// A packed batch therefore keeps per-scene statistics: one CTA owns one scene at a time
// (warp t <-> observed timestep t, lanes <-> pedestrian columns) and the three BN layers
// cost CTA-wide reductions, two-pass (mean, then centred squares) like ATen.
//
// Gradients of the 142 st_gcn parameters are accumulated per thread in registers over
// all scenes the CTA walks, reduced once per CTA in a fixed order and written to a
// per-CTA partial row; sstg_grad_finalize sums the rows in ascending CTA order, so the
// result is bit-reproducible for a given grid (no float atomics anywhere).
#include "sstg_common.cuh"
#include "bn_update.cuh"

namespace sstg {

namespace {

// -DSSTG_PHASE_CLOCKS: thread 0 of every CTA stamps clock64() at the phase boundaries of its first scene
// (profiling aid for the latency-bound small-group case; read back with sstg_debug_phase_clocks_st).
#ifdef SSTG_PHASE_CLOCKS
__device__ long long g_phase_st[300 * 24];
#define PHASE_ST(k)                                                                                      \
  do {                                                                                                   \
    if (threadIdx.x == 0 && s == static_cast<int>(blockIdx.x)) g_phase_st[blockIdx.x * 24 + (k)] = clock64(); \
  } while (0)
#else
#define PHASE_ST(k)
#endif

constexpr int R_WARPS = T_OBS;
constexpr int R_THREADS = R_WARPS * 32;
constexpr int ST_PARAMS = P_TP0_W;  // 142 st_gcn parameters come first in the flat block

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// CTA-wide sum of NV values per thread; every thread receives the totals (summed over the
// 8 warps in ascending order).  `red` is shared scratch [R_WARPS][NV].
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == 0) red[warp * NV + i] = s;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < R_WARPS; ++w) s += red[w * NV + i];
    v[i] = s;
  }
  __syncthreads();
}

struct TrainParams {
  float gw[C][2], gb[C];
  float g1[C], b1[C], a1;
  float w2[C][C][3], b2[C];
  float g2[C], be2[C];
  float rw[C][2], rb[C];
  float gr[C], ber[C], a2;
};

__device__ __forceinline__ void load_train_params(TrainParams& tp, const float* __restrict__ params) {
  const int tid = threadIdx.x;
  if (tid < C) {
    const int c = tid;
    tp.gw[c][0] = params[P_GCN_W + 2 * c];
    tp.gw[c][1] = params[P_GCN_W + 2 * c + 1];
    tp.gb[c] = params[P_GCN_B + c];
    tp.g1[c] = params[P_BN1_G + c];
    tp.b1[c] = params[P_BN1_B + c];
    for (int ci = 0; ci < C; ++ci)
      for (int k = 0; k < 3; ++k) tp.w2[c][ci][k] = params[P_TCN_W + (c * C + ci) * 3 + k];
    tp.b2[c] = params[P_TCN_B + c];
    tp.g2[c] = params[P_BN2_G + c];
    tp.be2[c] = params[P_BN2_B + c];
    tp.rw[c][0] = params[P_RES_W + 2 * c];
    tp.rw[c][1] = params[P_RES_W + 2 * c + 1];
    tp.rb[c] = params[P_RES_B + c];
    tp.gr[c] = params[P_BNR_G + c];
    tp.ber[c] = params[P_BNR_B + c];
    if (c == 0) {
      tp.a1 = params[P_PRELU1];
      tp.a2 = params[P_PRELU2];
    }
  }
}

// ---------------------------------------------------------------------------------------
// st_gcn forward, train mode.  Saves Y (pre-BN1), Z (pre-BN2), H and the per-scene batch
// statistics stats[s] = {mean1[5], var1[5], mean2[5], var2[5], mean_r[5], var_r[5]}
// (biased variances).
// ---------------------------------------------------------------------------------------
template <int NJ>
__global__ void __launch_bounds__(R_THREADS)
stgcn_train_fwd_kernel(const float* __restrict__ V, const float* __restrict__ A,
                       const int32_t* __restrict__ seg_off, const int64_t* __restrict__ a_off, int S, int NP,
                       const float* __restrict__ params, float* __restrict__ Y, float* __restrict__ Z,
                       float* __restrict__ H, float* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrainParams& tp = *reinterpret_cast<TrainParams*>(smem_raw);
  float* red = reinterpret_cast<float*>(smem_raw + ((sizeof(TrainParams) + 15) & ~size_t(15)));  // [8][10]
  float* gbuf = red + R_WARPS * 10;                                                             // [8][NP][8]
  float* pbuf = gbuf + R_WARPS * NP * 8;                                                        // [5][8][NP]
  const int tid = threadIdx.x, lane = tid & 31, t = tid >> 5;
#ifdef SSTG_PHASE_CLOCKS
  if (tid == 0) g_phase_st[blockIdx.x * 24 + 12] = clock64();
#endif
  load_train_params(tp, params);
  pdl_wait();
  pdl_release();
  __syncthreads();
  float* gw = gbuf + static_cast<size_t>(t) * NP * 8;

  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    PHASE_ST(13);
    const int n0 = seg_off[s], N = seg_off[s + 1] - n0;
    const float invM = 1.f / static_cast<float>(T_OBS * N);
    const float2* vp = reinterpret_cast<const float2*>(V) + (static_cast<size_t>(n0) * T_OBS + static_cast<size_t>(t) * N);
    float vx[NJ], vy[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = lane + 32 * j;
      const float2 v = k < N ? vp[k] : make_float2(0.f, 0.f);
      vx[j] = v.x;
      vy[j] = v.y;
      if (k < N) {
#pragma unroll
        for (int c = 0; c < C; ++c) gw[k * 8 + c] = fmaf(tp.gw[c][1], v.y, fmaf(tp.gw[c][0], v.x, tp.gb[c]));
      }
    }
    __syncwarp();
    // y[c, t, w] = sum_v g[c, t, v] A[t, v, w]  (rows of A read coalesced, lanes = columns w)
    float y[C][NJ];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int j = 0; j < NJ; ++j) y[c][j] = 0.f;
    const float* At = A + a_off[s] + static_cast<size_t>(t) * N * N;
#pragma unroll 16
    for (int v = 0; v < N; ++v) {
      const float4 g4 = *reinterpret_cast<const float4*>(gw + v * 8);
      const float g5 = gw[v * 8 + 4];
      const float* row = At + static_cast<size_t>(v) * N;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        const float a = k < N ? row[k] : 0.f;
        y[0][j] = fmaf(g4.x, a, y[0][j]);
        y[1][j] = fmaf(g4.y, a, y[1][j]);
        y[2][j] = fmaf(g4.z, a, y[2][j]);
        y[3][j] = fmaf(g4.w, a, y[3][j]);
        y[4][j] = fmaf(g5, a, y[4][j]);
      }
    }
    PHASE_ST(14);
    // residual branch pre-BN: r = Wr v + br
    float r[C][NJ];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int j = 0; j < NJ; ++j) r[c][j] = fmaf(tp.rw[c][1], vy[j], fmaf(tp.rw[c][0], vx[j], tp.rb[c]));
    // ---- batch statistics of y and r (two-pass)
    float m[10];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float sy = 0.f, sr = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (lane + 32 * j < N) {
          sy += y[c][j];
          sr += r[c][j];
        }
      m[c] = sy;
      m[5 + c] = sr;
    }
    block_sum<10>(m, red);
#pragma unroll
    for (int i = 0; i < 10; ++i) m[i] *= invM;
    float q[10];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float sy = 0.f, sr = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (lane + 32 * j < N) {
          const float dy = y[c][j] - m[c], dr = r[c][j] - m[5 + c];
          sy = fmaf(dy, dy, sy);
          sr = fmaf(dr, dr, sr);
        }
      q[c] = sy;
      q[5 + c] = sr;
    }
    block_sum<10>(q, red);
#pragma unroll
    for (int i = 0; i < 10; ++i) q[i] *= invM;  // biased variances
    if (tid < C) {
      float* st = stats + static_cast<size_t>(s) * SSTG_N_BN;
      st[tid] = m[tid];
      st[5 + tid] = q[tid];
      st[20 + tid] = m[5 + tid];
      st[25 + tid] = q[5 + tid];
    }
    PHASE_ST(15);
    // ---- BN1 + PReLU1 -> shared p; save y
    float* Ys = Y + static_cast<size_t>(n0) * 40;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float isd = 1.0f / sqrtf(q[c] + BN_EPS);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        if (k < N) {
          Ys[(c * T_OBS + t) * N + k] = y[c][j];
          const float xh = (y[c][j] - m[c]) * isd;
          pbuf[(c * T_OBS + t) * NP + k] = prelu(fmaf(xh, tp.g1[c], tp.b1[c]), tp.a1);
        }
      }
    }
    __syncthreads();
    PHASE_ST(16);
    // ---- (3,1) temporal conv
    float z[C][NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = lane + 32 * j;
#pragma unroll
      for (int co = 0; co < C; ++co) z[co][j] = tp.b2[co];
      if (k < N) {
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          const int tt = t + kt - 1;
          if (tt >= 0 && tt < T_OBS) {
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
              const float pv = pbuf[(ci * T_OBS + tt) * NP + k];
#pragma unroll
              for (int co = 0; co < C; ++co) z[co][j] = fmaf(tp.w2[co][ci][kt], pv, z[co][j]);
            }
          }
        }
      }
    }
    PHASE_ST(17);
    float mz[5], qz[5];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float sz = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (lane + 32 * j < N) sz += z[c][j];
      mz[c] = sz;
    }
    block_sum<5>(mz, red);  // also orders the pbuf reads above before the next scene's writes
#pragma unroll
    for (int c = 0; c < C; ++c) mz[c] *= invM;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float sz = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (lane + 32 * j < N) {
          const float d = z[c][j] - mz[c];
          sz = fmaf(d, d, sz);
        }
      qz[c] = sz;
    }
    block_sum<5>(qz, red);
#pragma unroll
    for (int c = 0; c < C; ++c) qz[c] *= invM;
    if (tid < C) {
      float* st = stats + static_cast<size_t>(s) * SSTG_N_BN;
      st[10 + tid] = mz[tid];
      st[15 + tid] = qz[tid];
    }
    PHASE_ST(18);
    float* Zs = Z + static_cast<size_t>(n0) * 40;
    float* Hs = H + static_cast<size_t>(n0) * 40;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float is2 = 1.0f / sqrtf(qz[c] + BN_EPS);
      const float isr = 1.0f / sqrtf(q[5 + c] + BN_EPS);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        if (k < N) {
          Zs[(c * T_OBS + t) * N + k] = z[c][j];
          const float bz = fmaf((z[c][j] - mz[c]) * is2, tp.g2[c], tp.be2[c]);
          const float br = fmaf((r[c][j] - m[5 + c]) * isr, tp.gr[c], tp.ber[c]);
          Hs[(c * T_OBS + t) * N + k] = prelu(bz + br, tp.a2);
        }
      }
    }
    PHASE_ST(19);
  }
}

// ---------------------------------------------------------------------------------------
// st_gcn backward.  dH (5,8,N) blocks in; partial[blockIdx.x][0..141] out.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(R_THREADS)
stgcn_train_bwd_kernel(const float* __restrict__ V, const float* __restrict__ A,
                       const int32_t* __restrict__ seg_off, const int64_t* __restrict__ a_off, int S, int NP,
                       const float* __restrict__ params, const float* __restrict__ Y,
                       const float* __restrict__ Z, const float* __restrict__ stats,
                       const float* __restrict__ dH, float* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrainParams& tp = *reinterpret_cast<TrainParams*>(smem_raw);
  float* red = reinterpret_cast<float*>(smem_raw + ((sizeof(TrainParams) + 15) & ~size_t(15)));  // [8][128]
  float* gsum = red + R_WARPS * 128;                     // [144] per-CTA sums of the per-scene-reduced grads
  float* pb = gsum + 144;                                // [5][8][NP]  p = PReLU1(BN1(y))
  float* dz = pb + C * T_OBS * NP;                       // [5][8][NP]  dL/dz
  float* dq = dz + C * T_OBS * NP;                       // [5][8][NP]  dL/dq, then dL/dy in place
  const int tid = threadIdx.x, lane = tid & 31, t = tid >> 5;
#ifdef SSTG_PHASE_CLOCKS
  if (tid == 0) g_phase_st[blockIdx.x * 24 + 0] = clock64();
#endif
  load_train_params(tp, params);
  if (tid < 144) gsum[tid] = 0.f;
  pdl_wait();
  pdl_release();
  __syncthreads();

  // per-thread gradient partials carried over all scenes of this CTA
  float dWr[C][2], dbr[C], dW2[C][C][3], db2[C], da1 = 0.f, da2 = 0.f;
  float dWg[C][2], dbg[C];  // per lane (lane <-> source node), warp-summed once at the end
#pragma unroll
  for (int c = 0; c < C; ++c) {
    dWr[c][0] = dWr[c][1] = dbr[c] = db2[c] = dWg[c][0] = dWg[c][1] = dbg[c] = 0.f;
#pragma unroll
    for (int ci = 0; ci < C; ++ci) dW2[c][ci][0] = dW2[c][ci][1] = dW2[c][ci][2] = 0.f;
  }

  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    PHASE_ST(1);
    const int n0 = seg_off[s], N = seg_off[s + 1] - n0;
    const float M = static_cast<float>(T_OBS * N), invM = 1.f / M;
    const float* st = stats + static_cast<size_t>(s) * SSTG_N_BN;
    float mu1[C], is1[C], mu2[C], is2[C], mur[C], isr[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      mu1[c] = st[c];
      is1[c] = 1.0f / sqrtf(st[5 + c] + BN_EPS);
      mu2[c] = st[10 + c];
      is2[c] = 1.0f / sqrtf(st[15 + c] + BN_EPS);
      mur[c] = st[20 + c];
      isr[c] = 1.0f / sqrtf(st[25 + c] + BN_EPS);
    }
    const float2* vp = reinterpret_cast<const float2*>(V) + (static_cast<size_t>(n0) * T_OBS + static_cast<size_t>(t) * N);
    const float* Ys = Y + static_cast<size_t>(n0) * 40;
    const float* Zs = Z + static_cast<size_t>(n0) * 40;
    const float* Gs = dH + static_cast<size_t>(n0) * 40;

    // ---- phase A: d_s = dH * PReLU2'(s);  sums for BN2 / BN_res backward
    float sA[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) sA[i] = 0.f;
    for (int k = lane; k < N; k += 32) {
      const float2 v = vp[k];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float zc = Zs[(c * T_OBS + t) * N + k];
        const float rc = fmaf(tp.rw[c][1], v.y, fmaf(tp.rw[c][0], v.x, tp.rb[c]));
        const float xh2 = (zc - mu2[c]) * is2[c], xhr = (rc - mur[c]) * isr[c];
        const float sv = fmaf(xh2, tp.g2[c], tp.be2[c]) + fmaf(xhr, tp.gr[c], tp.ber[c]);
        const float g = Gs[(c * T_OBS + t) * N + k];
        const float ds = sv > 0.f ? g : tp.a2 * g;
        if (!(sv > 0.f)) da2 = fmaf(g, sv, da2);
        sA[c] += ds;
        sA[5 + c] = fmaf(ds, xh2, sA[5 + c]);
        sA[10 + c] = fmaf(ds, xhr, sA[10 + c]);
      }
    }
    PHASE_ST(2);
    block_sum<15>(sA, red);
    PHASE_ST(3);
    if (tid < 15) gsum[tid] += sA[tid];  // [0..4] dbeta2 (= dbeta_r), [5..9] dgamma2, [10..14] dgamma_r

    // ---- phase B: d_z, d_r; p = PReLU1(BN1(y)) to shared
    for (int k = lane; k < N; k += 32) {
      const float2 v = vp[k];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float zc = Zs[(c * T_OBS + t) * N + k];
        const float rc = fmaf(tp.rw[c][1], v.y, fmaf(tp.rw[c][0], v.x, tp.rb[c]));
        const float xh2 = (zc - mu2[c]) * is2[c], xhr = (rc - mur[c]) * isr[c];
        const float sv = fmaf(xh2, tp.g2[c], tp.be2[c]) + fmaf(xhr, tp.gr[c], tp.ber[c]);
        const float g = Gs[(c * T_OBS + t) * N + k];
        const float ds = sv > 0.f ? g : tp.a2 * g;
        const float dzc = tp.g2[c] * is2[c] * invM * (M * ds - sA[c] - xh2 * sA[5 + c]);
        const float drc = tp.gr[c] * isr[c] * invM * (M * ds - sA[c] - xhr * sA[10 + c]);
        dz[(c * T_OBS + t) * NP + k] = dzc;
        dWr[c][0] = fmaf(drc, v.x, dWr[c][0]);
        dWr[c][1] = fmaf(drc, v.y, dWr[c][1]);
        dbr[c] += drc;
        const float yc = Ys[(c * T_OBS + t) * N + k];
        const float qv = fmaf((yc - mu1[c]) * is1[c], tp.g1[c], tp.b1[c]);
        pb[(c * T_OBS + t) * NP + k] = prelu(qv, tp.a1);
      }
    }
    __syncthreads();
    PHASE_ST(4);

    // ---- phase C: temporal-conv backward (weights, bias, input), PReLU1 backward, BN1 sums
    float sC[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) sC[i] = 0.f;
    for (int k = lane; k < N; k += 32) {
      float dzo[C];
#pragma unroll
      for (int co = 0; co < C; ++co) {
        dzo[co] = dz[(co * T_OBS + t) * NP + k];
        db2[co] += dzo[co];
      }
      float dp[C];
#pragma unroll
      for (int ci = 0; ci < C; ++ci) dp[ci] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        const int ti = t + kt - 1;   // z[t] read p[t + kt - 1]
        if (ti >= 0 && ti < T_OBS) {
#pragma unroll
          for (int ci = 0; ci < C; ++ci) {
            const float pv = pb[(ci * T_OBS + ti) * NP + k];
#pragma unroll
            for (int co = 0; co < C; ++co) dW2[co][ci][kt] = fmaf(dzo[co], pv, dW2[co][ci][kt]);
          }
        }
        const int tz = t - kt + 1;      // p[t] was read by z[t - kt + 1] with tap kt
        if (tz >= 0 && tz < T_OBS) {
#pragma unroll
          for (int co = 0; co < C; ++co) {
            const float dv = dz[(co * T_OBS + tz) * NP + k];
#pragma unroll
            for (int ci = 0; ci < C; ++ci) dp[ci] = fmaf(tp.w2[co][ci][kt], dv, dp[ci]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float yc = Ys[(c * T_OBS + t) * N + k];
        const float xh1 = (yc - mu1[c]) * is1[c];
        const float qv = fmaf(xh1, tp.g1[c], tp.b1[c]);
        const float dqc = qv > 0.f ? dp[c] : tp.a1 * dp[c];
        if (!(qv > 0.f)) da1 = fmaf(dp[c], qv, da1);
        dq[(c * T_OBS + t) * NP + k] = dqc;
        sC[c] += dqc;
        sC[5 + c] = fmaf(dqc, xh1, sC[5 + c]);
      }
    }
    PHASE_ST(5);
    block_sum<10>(sC, red);
    PHASE_ST(6);
    if (tid < 10) gsum[16 + tid] += sC[tid];  // [16..20] dbeta1, [21..25] dgamma1

    // ---- phase D: BN1 backward -> d_y (in place in dq; only this warp touches row t)
    for (int k = lane; k < N; k += 32) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float yc = Ys[(c * T_OBS + t) * N + k];
        const float xh1 = (yc - mu1[c]) * is1[c];
        const float d = dq[(c * T_OBS + t) * NP + k];
        dq[(c * T_OBS + t) * NP + k] = tp.g1[c] * is1[c] * invM * (M * d - sC[c] - xh1 * sC[5 + c]);
      }
    }
    __syncwarp();
    PHASE_ST(7);
    // ---- phase E: d_g[c, t, v] = sum_w d_y[c, t, w] A[t, v, w];  dWg, dbg.
    // Lanes <-> source nodes v, each lane walks ITS row of A (w ascending, 16 loads in flight) against
    // d_y broadcast from shared memory: no cross-lane reduction and no load->shuffle chain per node --
    // the per-node latency chain was what set the duration of a small group.  A[t] is at most 16 KB per
    // warp at N = 64 and stays in L1 across the strided row walk.
    const float* At = A + a_off[s] + static_cast<size_t>(t) * N * N;
    const float* dyt = dq + t * NP;
    for (int v = lane; v < N; v += 32) {
      const float* row = At + static_cast<size_t>(v) * N;
      float pg[C] = {0.f, 0.f, 0.f, 0.f, 0.f};
      int w = 0;
      for (; w + 16 <= N; w += 16) {
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = __ldg(row + w + j);
#pragma unroll
        for (int j = 0; j < 16; ++j)
#pragma unroll
          for (int c = 0; c < C; ++c) pg[c] = fmaf(dyt[c * T_OBS * NP + w + j], a[j], pg[c]);
      }
      for (; w + 4 <= N; w += 4) {
        float a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = __ldg(row + w + j);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < C; ++c) pg[c] = fmaf(dyt[c * T_OBS * NP + w + j], a[j], pg[c]);
      }
      for (; w < N; ++w) {
        const float a = __ldg(row + w);
#pragma unroll
        for (int c = 0; c < C; ++c) pg[c] = fmaf(dyt[c * T_OBS * NP + w], a, pg[c]);
      }
      const float2 vv = vp[v];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dWg[c][0] = fmaf(pg[c], vv.x, dWg[c][0]);
        dWg[c][1] = fmaf(pg[c], vv.y, dWg[c][1]);
        dbg[c] += pg[c];
      }
    }
    PHASE_ST(8);
    __syncthreads();  // pb / dz / dq are rewritten by the next scene
    PHASE_ST(9);
  }
#ifdef SSTG_PHASE_CLOCKS
  if (tid == 0) g_phase_st[blockIdx.x * 24 + 10] = clock64();
#endif

  // ---- CTA reduction of the carried partials, fixed order, then the partial row.
  // 112 per-thread values (97 carried + dWg/dbg), padded to 128: a warp reduce-scatter (at each butterfly
  // step a lane hands over the half its partner keeps: 64+32+16+8+4 shuffles instead of 112 x 5) leaves lane
  // L with the warp totals of values 4L..4L+3; thread i < 112 then adds the 8 warps in ascending order and
  // stores parameter i of the row -- no serial tail on one thread.
  {
    float v1[128];
    int i = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      v1[i++] = dWr[c][0];
      v1[i++] = dWr[c][1];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) v1[i++] = dbr[c];
#pragma unroll
    for (int co = 0; co < C; ++co)
#pragma unroll
      for (int ci = 0; ci < C; ++ci)
#pragma unroll
        for (int k = 0; k < 3; ++k) v1[i++] = dW2[co][ci][k];
#pragma unroll
    for (int c = 0; c < C; ++c) v1[i++] = db2[c];
    v1[i++] = da1;
    v1[i++] = da2;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      v1[i++] = dWg[c][0];
      v1[i++] = dWg[c][1];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) v1[i++] = dbg[c];
#pragma unroll
    for (int j = 112; j < 128; ++j) v1[j] = 0.f;
#pragma unroll
    for (int half = 64; half >= 4; half >>= 1) {
      const int o = half >> 2;                       // lane offsets 16, 8, 4, 2, 1
      const bool up = (lane & o) != 0;               // this lane keeps the upper half
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float send = up ? v1[j] : v1[j + half];
        const float keep = up ? v1[j + half] : v1[j];
        v1[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    *reinterpret_cast<float4*>(red + t * 128 + 4 * lane) = make_float4(v1[0], v1[1], v1[2], v1[3]);
    __syncthreads();
    float* out = partial + static_cast<size_t>(blockIdx.x) * ST_PARAMS;
    if (tid < 112) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < R_WARPS; ++w) sum += red[w * 128 + tid];
      const int dst = tid < 10 ? P_RES_W + tid
                    : tid < 15 ? P_RES_B + tid - 10
                    : tid < 90 ? P_TCN_W + tid - 15
                    : tid < 95 ? P_TCN_B + tid - 90
                    : tid == 95 ? P_PRELU1
                    : tid == 96 ? P_PRELU2
                    : tid < 107 ? P_GCN_W + tid - 97
                                : P_GCN_B + tid - 107;
      out[dst] = sum;
    } else if (tid >= 128 && tid < 128 + C) {
      const int c = tid - 128;
      out[P_BN2_B + c] = gsum[c];
      out[P_BNR_B + c] = gsum[c];
      out[P_BN2_G + c] = gsum[5 + c];
      out[P_BNR_G + c] = gsum[10 + c];
      out[P_BN1_B + c] = gsum[16 + c];
      out[P_BN1_G + c] = gsum[21 + c];
    }
#ifdef SSTG_PHASE_CLOCKS
    if (tid == 0) g_phase_st[blockIdx.x * 24 + 11] = clock64();
#endif
  }
}

// ---------------------------------------------------------------------------------------
// NLL forward + gradient w.r.t. V_pred (metrics.py:84-113): one CTA per scene, warp t <-> predicted
// timestep t, lanes <-> pedestrians -- an optimiser group is ~128 small scenes, so the kernel is
// latency-bound and the longest scene sets its duration; 12 warps cut that chain 12-fold.
//   loss[s]  = mean over the 12*N_s elements of -log(max(pdf, 1e-20))
//   dVpred   = weight[s] / (12 N_s) * d(-log pdf)/dV_pred      (zero where the clamp is active)
// The scene sum is taken per warp, then over the 12 warps in ascending order (fixed order).
// ---------------------------------------------------------------------------------------
constexpr int NLL_THREADS = T_PRED * 32;
__global__ void __launch_bounds__(NLL_THREADS)
nll_fwd_bwd_kernel(const float* __restrict__ Vpred, const float* __restrict__ Vtr,
                   const int32_t* __restrict__ seg_off, int S, const float* __restrict__ weight,
                   float* __restrict__ loss, float* __restrict__ dV, int vtr_is_rel) {
  __shared__ float part[T_PRED];
  const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_release();
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    const int n0 = seg_off[s], N = seg_off[s + 1] - n0;
    const float* vp = Vpred + static_cast<size_t>(n0) * 60;
    float* dv = dV + static_cast<size_t>(n0) * 60;
    const float2* vt = reinterpret_cast<const float2*>(Vtr) + static_cast<size_t>(n0) * T_PRED;
    const int tot = T_PRED * N;
    const float wgt = (weight != nullptr ? weight[s] : 1.f) / static_cast<float>(tot);
    float sum = 0.f;
    for (int n = lane; n < N; n += 32) {
      const int i = t * N + n;
      float2 tr;
      if (vtr_is_rel) {  // Vtr is pred_traj_rel [P,2,12] itself: the target V_tr (utils.py:44) is its transpose
        const float* pr = Vtr + static_cast<size_t>(n0 + n) * (2 * T_PRED) + t;
        tr = make_float2(__ldg(pr), __ldg(pr + T_PRED));
      } else {
        tr = vt[i];
      }
      const float m0 = vp[i], m1 = vp[12 * N + i], l2 = vp[24 * N + i], l3 = vp[36 * N + i], l4 = vp[48 * N + i];
      const float nx = tr.x - m0;
      const float ny = tr.y - m1;
      const float sx = expf(l2);
      const float sy = expf(l3);
      const float rho = tanhf(l4);
      const float sxsy = sx * sy;
      const float qx = nx / sx, qy = ny / sy;
      const float cross = __fdiv_rn(__fmul_rn(__fmul_rn(rho, nx), ny), sxsy);
      const float z = __fsub_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(2.f, cross));
      const float q = __fsub_rn(1.f, __fmul_rn(rho, rho));
      float pdf = expf(__fdiv_rn(-z, __fmul_rn(2.f, q)));
      pdf = __fdiv_rn(pdf, __fmul_rn(6.283185307179586f, __fmul_rn(sxsy, sqrtf(q))));
      sum += -logf(fmaxf(pdf, 1e-20f));
      float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f;
      if (pdf >= 1e-20f) {  // torch.clamp passes the gradient where input >= min
        const float iq = 1.f / q;
        g0 = -(qx / sx - rho * qy / sx) * iq;
        g1 = -(qy / sy - rho * qx / sy) * iq;
        g2 = 1.f - (qx * qx - cross) * iq;
        g3 = 1.f - (qy * qy - cross) * iq;
        g4 = -(nx * ny) / sxsy + z * rho * iq - rho;
      }
      dv[i] = g0 * wgt;
      dv[12 * N + i] = g1 * wgt;
      dv[24 * N + i] = g2 * wgt;
      dv[36 * N + i] = g3 * wgt;
      dv[48 * N + i] = g4 * wgt;
    }
    sum = warp_sum(sum);
    if (lane == 0) part[t] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot_s = 0.f;
#pragma unroll
      for (int w = 0; w < T_PRED; ++w) tot_s += part[w];
      loss[s] = tot_s / static_cast<float>(tot);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// BatchNorm running statistics: r <- (1 - m) r + m s_i, scene after scene, unbiased var.
// 30 threads; ATen's CPU kernel evaluates the update in double, so do we.
// ---------------------------------------------------------------------------------------
// (the recurrence itself lives in bn_update.cuh: the TXP-CNN training forward runs it as an extra CTA)
__global__ void __launch_bounds__(256)
bn_running_update_kernel(const float* __restrict__ stats, const int32_t* __restrict__ seg_off, int S,
                         double momentum, float* __restrict__ bn, const int32_t* __restrict__ abort_flag) {
  // a failed peer exchange (peer_exchange.cu) leaves the statistics undefined: leave the buffers alone
  if (abort_flag && *abort_flag) return;
  bn_running_update_body(stats, seg_off, S, momentum, bn);
}

// grad[i] = sum over CTA rows (ascending) of the two partial buffers; parameters that are
// never executed (tpcnns.4, prelus.4 -- model.py:191) get zero.
// ascending-row sum of one column; the loads of 16 rows are in flight together, the additions stay in order
__device__ __forceinline__ float column_sum(const float* __restrict__ p, int rows, int stride) {
  float s = 0.f;
  int r = 0;
  for (; r + 16 <= rows; r += 16) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __ldg(p + static_cast<size_t>(r + j) * stride);
#pragma unroll
    for (int j = 0; j < 16; ++j) s += v[j];
  }
  for (; r + 4 <= rows; r += 4) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(p + static_cast<size_t>(r + j) * stride);
#pragma unroll
    for (int j = 0; j < 4; ++j) s += v[j];
  }
  for (; r < rows; ++r) s += __ldg(p + static_cast<size_t>(r) * stride);
  return s;
}

__device__ __forceinline__ float grad_finalize_one(int i, const float* __restrict__ part_st, int rows_st,
                                                   const float* __restrict__ part_tx, int rows_tx, int tx_stride) {
  return i < ST_PARAMS ? column_sum(part_st + i, rows_st, ST_PARAMS)
                       : column_sum(part_tx + (i - ST_PARAMS), rows_tx, tx_stride);
}

__global__ void grad_finalize_kernel(const float* __restrict__ part_st, int rows_st,
                                     const float* __restrict__ part_tx, int rows_tx, int tx_stride,
                                     float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= SSTG_N_PARAMS) return;
  grad[i] = grad_finalize_one(i, part_st, rows_st, part_tx, rows_tx, tx_stride);
}

// SGD (train.py:197: lr only) with optional clip_grad_norm_ (train.py:69-70).
__global__ void __launch_bounds__(1024)
sgd_step_kernel(float* __restrict__ params, const float* __restrict__ grad, int n, float lr, float clip_norm,
                float* __restrict__ grad_norm_out) {
  __shared__ float red[32];
  __shared__ float coef_s;
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ss = fmaf(grad[i], grad[i], ss);
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) tot += red[w];
    const float nrm = sqrtf(tot);
    if (grad_norm_out) *grad_norm_out = nrm;
    float coef = 1.f;
    if (clip_norm > 0.f) {
      coef = clip_norm / (nrm + 1e-6f);
      if (coef > 1.f) coef = 1.f;
    }
    coef_s = coef;
  }
  __syncthreads();
  const float step = lr * coef_s;
  for (int i = threadIdx.x; i < n; i += blockDim.x) params[i] = fmaf(-step, grad[i], params[i]);
}

// group loss = sum_s weight[s] * loss[s] (train.py:58-66), fixed order, fp64 accumulate
__device__ __forceinline__ void group_loss_body(const float* __restrict__ loss, const float* __restrict__ weight,
                                                int S, float* __restrict__ out) {
  double acc = 0.0;
  for (int s = threadIdx.x; s < S; s += 32) acc += static_cast<double>(loss[s]) * (weight ? weight[s] : 1.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) *out = static_cast<float>(acc);
}
__global__ void group_loss_kernel(const float* __restrict__ loss, const float* __restrict__ weight, int S,
                                  float* __restrict__ out) {
  group_loss_body(loss, weight, S, out);
}

// The tail of one optimiser group in ONE launch (the group is latency-bound: four dependent launches of
// ~3-16 us each become one): CTAs 0..n-3 reduce the gradient partial rows and, when no clipping is asked
// for, apply SGD on the spot; CTA n-2 runs the BatchNorm recurrence; CTA n-1 the group loss.
constexpr int TAIL_FIN_CTAS = (SSTG_N_PARAMS + 255) / 256;
__global__ void __launch_bounds__(256)
train_tail_kernel(const float* __restrict__ part_st, int rows_st, const float* __restrict__ part_tx, int rows_tx,
                  int tx_stride, float* __restrict__ grad, float* __restrict__ params, float lr,
                  const float* __restrict__ stats, const int32_t* __restrict__ seg_off, int S, double momentum,
                  float* __restrict__ bn, const float* __restrict__ loss, const float* __restrict__ weight,
                  float* __restrict__ group_loss) {
  pdl_wait();  // writes `params`: never releases its dependents early
  if (blockIdx.x < TAIL_FIN_CTAS) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= SSTG_N_PARAMS) return;
    const float g = grad_finalize_one(i, part_st, rows_st, part_tx, rows_tx, tx_stride);
    grad[i] = g;
    if (lr != 0.f) params[i] = fmaf(-lr, g, params[i]);
  } else if (blockIdx.x == TAIL_FIN_CTAS) {
    if (momentum >= 0.0) bn_running_update_body(stats, seg_off, S, momentum, bn);
  } else if (group_loss != nullptr && threadIdx.x < 32) {
    group_loss_body(loss, weight, S, group_loss);
  }
}

size_t train_fwd_smem(int NP) {
  return ((sizeof(TrainParams) + 15) & ~size_t(15)) + sizeof(float) * (R_WARPS * 10 + R_WARPS * NP * 8 + C * T_OBS * NP);
}
size_t train_bwd_smem(int NP) {
  return ((sizeof(TrainParams) + 15) & ~size_t(15)) + sizeof(float) * (R_WARPS * 128 + 144 + 3 * C * T_OBS * NP);
}

template <int NJ>
int launch_train_fwd_nj(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off, int S, int NP,
                        const float* params, float* Y, float* Z, float* H, float* stats, int grid,
                        cudaStream_t st) {
  const size_t smem = train_fwd_smem(NP);
  auto kern = stgcn_train_fwd_kernel<NJ>;
  if (smem > 48 * 1024)
    SSTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                  "stgcn_train_fwd smem attr");
  SSTG_CUDA_TRY(launch_pdl(kern, grid, R_THREADS, smem, st, V, A, seg_off, a_off, S, NP, params, Y, Z, H, stats),
                "stgcn_train_fwd launch");
  return SSTG_OK;
}

}  // namespace

#ifdef SSTG_PHASE_CLOCKS
extern "C" int sstg_debug_phase_clocks_st(long long* host_out, int n) {
  return cudaMemcpyFromSymbol(host_out, g_phase_st, sizeof(long long) * (n < 300 * 24 ? n : 300 * 24)) == cudaSuccess ? 0 : -3;
}
#endif

int train_grid(int S) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return S < 148 ? S : 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  const int cap = 2 * sms;
  return S < cap ? S : cap;
}

int launch_stgcn_train_forward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off, int S,
                               int max_n, const float* params, float* Y, float* Z, float* H, float* stats,
                               cudaStream_t st) {
  if (S == 0) return SSTG_OK;
  const int NP = (max_n + 3) & ~3;
  const int grid = train_grid(S);
  if (max_n <= 32) return launch_train_fwd_nj<1>(V, A, seg_off, a_off, S, NP, params, Y, Z, H, stats, grid, st);
  if (max_n <= 64) return launch_train_fwd_nj<2>(V, A, seg_off, a_off, S, NP, params, Y, Z, H, stats, grid, st);
  if (max_n <= 128) return launch_train_fwd_nj<4>(V, A, seg_off, a_off, S, NP, params, Y, Z, H, stats, grid, st);
  return launch_train_fwd_nj<8>(V, A, seg_off, a_off, S, NP, params, Y, Z, H, stats, grid, st);
}

int stgcn_train_bwd_rows(int S) { return train_grid(S); }

int launch_stgcn_train_backward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off, int S,
                                int max_n, const float* params, const float* Y, const float* Z,
                                const float* stats, const float* dH, float* partial, cudaStream_t st) {
  if (S == 0) return SSTG_OK;
  const int NP = (max_n + 3) & ~3;
  const size_t smem = train_bwd_smem(NP);
  auto kern = stgcn_train_bwd_kernel;
  if (smem > 48 * 1024)
    SSTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                  "stgcn_train_bwd smem attr");
  SSTG_CUDA_TRY(launch_pdl(kern, train_grid(S), R_THREADS, smem, st, V, A, seg_off, a_off, S, NP, params, Y, Z, stats,
                           dH, partial),
                "stgcn_train_bwd launch");
  return SSTG_OK;
}

int launch_nll_fwd_bwd(const float* Vpred, const float* Vtr, const int32_t* seg_off, int S, const float* weight,
                       float* loss, float* dV, cudaStream_t st, int vtr_is_rel) {
  if (S == 0) return SSTG_OK;
  const int cap = 8 * 148;
  SSTG_CUDA_TRY(launch_pdl(nll_fwd_bwd_kernel, S < cap ? S : cap, NLL_THREADS, 0, st, Vpred, Vtr, seg_off, S, weight,
                           loss, dV, vtr_is_rel),
                "nll_fwd_bwd launch");
  return SSTG_OK;
}

// CUDA loads kernels lazily, and a first-time load may synchronise the context: it must not happen while
// another kernel of this process is waiting for a peer (peer_exchange.cu) -- load ahead of time.
int preload_bn_running_update() {
  cudaFuncAttributes fa;
  SSTG_CUDA_TRY(cudaFuncGetAttributes(&fa, bn_running_update_kernel), "bn_running_update preload");
  return SSTG_OK;
}

int launch_bn_running_update(const float* stats, const int32_t* seg_off, int S, double momentum, float* bn,
                             cudaStream_t st, const int32_t* abort_flag) {
  if (S == 0) return SSTG_OK;
  bn_running_update_kernel<<<1, 256, 0, st>>>(stats, seg_off, S, momentum, bn, abort_flag);
  SSTG_CUDA_TRY(cudaGetLastError(), "bn_running_update launch");
  return SSTG_OK;
}

int launch_grad_finalize(const float* part_st, int rows_st, const float* part_tx, int rows_tx, int tx_stride,
                         float* grad, cudaStream_t st) {
  grad_finalize_kernel<<<(SSTG_N_PARAMS + 255) / 256, 256, 0, st>>>(part_st, rows_st, part_tx, rows_tx, tx_stride,
                                                                   grad);
  SSTG_CUDA_TRY(cudaGetLastError(), "grad_finalize launch");
  return SSTG_OK;
}

int launch_train_tail(const float* part_st, int rows_st, const float* part_tx, int rows_tx, int tx_stride,
                      float* grad, float* params, float lr, const float* stats, const int32_t* seg_off, int S,
                      double momentum, float* bn, const float* loss, const float* weight, float* group_loss,
                      cudaStream_t st) {
  SSTG_CUDA_TRY(launch_pdl(train_tail_kernel, TAIL_FIN_CTAS + 2, 256, 0, st, part_st, rows_st, part_tx, rows_tx,
                           tx_stride, grad, params, lr, stats, seg_off, S, momentum, bn, loss, weight, group_loss),
                "train_tail launch");
  return SSTG_OK;
}

int launch_group_loss(const float* loss, const float* weight, int S, float* out, cudaStream_t st) {
  group_loss_kernel<<<1, 32, 0, st>>>(loss, weight, S, out);
  SSTG_CUDA_TRY(cudaGetLastError(), "group_loss launch");
  return SSTG_OK;
}

int launch_sgd_step(float* params, const float* grad, int n, float lr, float clip_norm, float* grad_norm_out,
                    cudaStream_t st) {
  sgd_step_kernel<<<1, 1024, 0, st>>>(params, grad, n, lr, clip_norm, grad_norm_out);
  SSTG_CUDA_TRY(cudaGetLastError(), "sgd_step launch");
  return SSTG_OK;
}

}  // namespace sstg
