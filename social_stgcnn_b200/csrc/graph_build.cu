// Kernel (1): per-scene graph construction for a segment-packed batch of scenes.
//
// Replaces reference utils.py:23-53 (anorm + seq_to_graph + networkx
// normalized_laplacian_matrix).  For every scene s and timestep t:
//     W_hk = 1/||rel_h - rel_k||  (0 when the distance is 0),  W_hh = 1
//     d_h  = sum_k W_hk,  A = D^-1/2 (D - W) D^-1/2
// One CTA walks scenes (grid-stride); each of its 8 warps owns one timestep slice.
// Lanes own columns k; because W is symmetric the column sum a lane accumulates while
// the warp walks the rows h = 0..N-1 IS the row sum d_k, in exactly the ascending order
// scipy's csr mat-vec uses -- so no cross-lane reduction is needed at all and the fp64
// mode reproduces networkx bit for bit.  Rows are written as coalesced 128-byte segments.
//
// Two arithmetic modes (template F64):
//   F64  : dx,dy,dx^2,dy^2,+ in fp32 (separate roundings, as 0-dim fp32 torch tensors do),
//          then sqrt, 1/x, row sum, ds = 1/sqrt(d) and ds_h*((D-W)_hk*ds_k) in fp64, cast
//          to fp32 -- the reference's rounding points.
//   fp32 : rsqrt on the same fp32 squared distance, 4-way split fp32 row sums; within
//          ~3e-7 relative of the fp64 result (tests bound it at 2e-6; bar is 1e-5).
#include <cstdlib>

#include "sstg_common.cuh"

namespace sstg {

namespace {

constexpr int G_WARPS = 8;
constexpr int G_THREADS = G_WARPS * 32;

template <bool F64>
struct RealT {
  using type = float;
};
template <>
struct RealT<true> {
  using type = double;
};

// squared distance exactly as the reference: (x1-x2)**2 + (y1-y2)**2 on fp32 tensors
__device__ __forceinline__ float sqdist(float xh, float yh, float xk, float yk) {
  const float dx = __fsub_rn(xh, xk);
  const float dy = __fsub_rn(yh, yk);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
__device__ __forceinline__ double weight(float s, double) {
  const double nrm = sqrt(static_cast<double>(s));  // math.sqrt (utils.py:24)
  return nrm == 0.0 ? 0.0 : 1.0 / nrm;              // utils.py:25-27
}
// fp32 mode: one FMA for the squared distance, one MUFU for the inverse distance
__device__ __forceinline__ float weight(float s, float) { return inv_dist(s); }
__device__ __forceinline__ float sqdist_mode(float xh, float yh, float xk, float yk, float) {
  const float dx = xh - xk, dy = yh - yk;
  return fmaf(dx, dx, dy * dy);
}
__device__ __forceinline__ float sqdist_mode(float xh, float yh, float xk, float yk, double) {
  return sqdist(xh, yh, xk, yk);
}

template <bool F64, int NJ>
__global__ void __launch_bounds__(G_THREADS)
graph_build_kernel(const float* __restrict__ rel, const int32_t* __restrict__ seg_off,
                   const int64_t* __restrict__ a_off, int S, int T, int NP, float* __restrict__ V,
                   float* __restrict__ A, int flags) {
  using Real = typename RealT<F64>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);  // [T][NP]
  float* ys = xs + T * NP;                         // [T][NP]
  Real* dsw = reinterpret_cast<Real*>(ys + T * NP) + (threadIdx.x >> 5) * NP;  // [warp][NP]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();     // first kernels of the training chain: the launch overlaps the predecessor's tail
  pdl_release();
  const bool norm_lap = flags & SSTG_GRAPH_NORM_LAP;
  const bool v_only = flags & SSTG_GRAPH_V_ONLY;

  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    const int n0 = seg_off[s];
    const int N = seg_off[s + 1] - n0;
    const float* r = rel + static_cast<size_t>(n0) * 2 * T;
    const int tot = N * 2 * T;
    for (int i = tid; i < tot; i += G_THREADS) {  // (N,2,T) -> xs/ys[t][p]
      const int p = i / (2 * T), rem = i - p * 2 * T;
      const int c = rem / T, t = rem - c * T;
      (c ? ys : xs)[t * NP + p] = r[i];
    }
    __syncthreads();
    float2* Vs = reinterpret_cast<float2*>(V + static_cast<size_t>(n0) * T * 2);  // utils.py:42
    for (int i = tid; i < N * T; i += G_THREADS) {
      const int t = i / N, p = i - t * N;
      Vs[i] = make_float2(xs[t * NP + p], ys[t * NP + p]);
    }
    if (!v_only) {
      float* As = A + a_off[s];
      for (int t = warp; t < T; t += G_WARPS) {
        const float* xt = xs + t * NP;
        const float* yt = ys + t * NP;
        float xk[NJ], yk[NJ];
        Real dsk[NJ];
        Real lkk[NJ];  // (D - W)_kk = d_k - 1
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int k = lane + 32 * j;
          xk[j] = k < N ? xt[k] : 0.f;
          yk[j] = k < N ? yt[k] : 0.f;
        }
        if (norm_lap) {
          // ---- pass 1: d_k = sum_h W_hk (ascending h; W_kk = 1)
          Real dk[NJ];
          if (F64) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) dk[j] = Real(0);
            for (int h = 0; h < N; ++h) {
              const float xh = xt[h], yh = yt[h];
#pragma unroll
              for (int j = 0; j < NJ; ++j) {
                Real w = weight(sqdist_mode(xh, yh, xk[j], yk[j], Real(0)), Real(0));
                if (h == lane + 32 * j) w = Real(1);
                dk[j] += w;
              }
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) lkk[j] = dk[j] - Real(1);
          } else {
            Real acc[NJ][4];
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = Real(0);
            int h = 0;
            for (; h + 4 <= N; h += 4) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float xh = xt[h + u], yh = yt[h + u];
#pragma unroll
                for (int j = 0; j < NJ; ++j) acc[j][u] += weight(sqdist_mode(xh, yh, xk[j], yk[j], Real(0)), Real(0));
              }
            }
            for (; h < N; ++h) {
              const float xh = xt[h], yh = yt[h];
#pragma unroll
              for (int j = 0; j < NJ; ++j) acc[j][0] += weight(sqdist_mode(xh, yh, xk[j], yk[j], Real(0)), Real(0));
            }
            // the diagonal contributed w(0) = 0; d = 1 + off-diagonal sum
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              lkk[j] = (acc[j][0] + acc[j][1]) + (acc[j][2] + acc[j][3]);
              dk[j] = Real(1) + lkk[j];
            }
          }
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int k = lane + 32 * j;
            dsk[j] = F64 ? Real(1.0 / sqrt(static_cast<double>(dk[j]))) : Real(rsqrt_fast(static_cast<float>(dk[j])));
            if (k < N) dsw[k] = dsk[j];
          }
          __syncwarp();
        }
        // ---- pass 2: A_hk = ds_h * ((D - W)_hk * ds_k), one coalesced row per iteration
        float* At = As + static_cast<size_t>(t) * N * N;
        for (int h = 0; h < N; ++h) {
          const float xh = xt[h], yh = yt[h];
          const Real dsh = norm_lap ? dsw[h] : Real(1);
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int k = lane + 32 * j;
            const Real w = weight(sqdist_mode(xh, yh, xk[j], yk[j], Real(0)), Real(0));
            float a;
            if (norm_lap && !F64) {
              // -(w ds_h) ds_k + 0: the fused form yields +0.0 for absent edges without a select
              const float wv = (h == k) ? -static_cast<float>(lkk[j]) : static_cast<float>(w);
              a = fmaf(-(wv * static_cast<float>(dsh)), static_cast<float>(dsk[j]), 0.f);
            } else if (norm_lap) {
              const Real l = (h == k) ? lkk[j] : -w;
              a = static_cast<float>(dsh * (l * dsk[j]));
              if (h != k && w == Real(0)) a = 0.f;  // absent edge: +0.0, not -0.0
            } else {
              a = (h == k) ? 1.f : static_cast<float>(w);
            }
            if (k < N) At[static_cast<size_t>(h) * N + k] = a;
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Fast path (fp32, normalised Laplacian, N <= 64): every weight is computed ONCE.
// A work item is (scene, group of 4 timesteps); warp <-> timestep.  Pass 1 walks the rows,
// lanes own columns: w = rsqrt(|d|^2) goes to a per-warp shared W tile while the lane
// accumulates its column sum (= row sum, W symmetric).  Then ds = rsqrt(1 + sum) and the
// diagonal of the tile is overwritten with -(d - 1), so that pass 2 is the branch-free
// A = fma(-(w ds_h), ds_k, +0)  streamed from shared memory as 16-byte rows: full 128-byte
// store segments, and absent edges come out as +0.0 like the reference's.
// ---------------------------------------------------------------------------------------
constexpr int F_WARPS = 4;
constexpr int F_THREADS = F_WARPS * 32;
#ifndef SSTG_G_U1
#define SSTG_G_U1 1
#endif
constexpr int G_U1 = SSTG_G_U1;  // row groups (of 4) of pass 1 per trip (A/B knob)

// Requires T % 4 == 0 (8 and 12 in every configuration) and N <= 64.
template <int NJ>
__global__ void __launch_bounds__(F_THREADS)
graph_build_fast_kernel(const float* __restrict__ rel, const int32_t* __restrict__ seg_off,
                        const int64_t* __restrict__ a_off, int S, int T, int NP, int wtile,
                        float* __restrict__ V, float* __restrict__ A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);   // [4][NP]
  float* ys = xs + F_WARPS * NP;                    // [4][NP]
  float* dsm = ys + F_WARPS * NP;                   // [4][NP]
  float* Wt = dsm + F_WARPS * NP + static_cast<size_t>(threadIdx.x >> 5) * wtile;  // [warp][N*N]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  pdl_release();
  const int groups = T / F_WARPS;
  const long long items = static_cast<long long>(S) * groups;

  // Two-level software pipeline: the offsets (seg_off, a_off) of the item AFTER the next one and the
  // displacement float4 of the NEXT item are in flight while the current item is computed, so neither the
  // dependent offset -> address -> data chain nor the data load is ever waited for.
  long long it = blockIdx.x;
  int t0 = 0, n0 = 0, N = 0;             // item whose `pre` is in registers (current at the loop top)
  long long ao = 0;
  int q_n0 = 0, q_N = 0;                 // offsets of the item after it
  long long q_ao = 0;
  float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_offsets = [&](long long item) {
    const int sc = static_cast<int>(item / groups);
    q_n0 = seg_off[sc];
    q_N = seg_off[sc + 1] - q_n0;
    q_ao = a_off[sc];
  };
  auto load_rel = [&](long long item) {  // uses the offsets already in q_*
    const int sc = static_cast<int>(item / groups);
    t0 = static_cast<int>(item - static_cast<long long>(sc) * groups) * F_WARPS;
    n0 = q_n0;
    N = q_N;
    ao = q_ao;
    if (tid < 2 * N)  // tid = p*2 + c: one 16-byte run of 4 timesteps
      pre = *reinterpret_cast<const float4*>(rel + (static_cast<size_t>(n0) * 2 + tid) * T + t0);
  };
  if (it < items) {
    load_offsets(it);
    load_rel(it);
    if (it + gridDim.x < items) load_offsets(it + gridDim.x);
  }

  while (it < items) {
    const int ct0 = t0, cn0 = n0, cN = N;
    const long long cao = ao;
    if (tid < 2 * cN) {
      float* dst = ((tid & 1) ? ys : xs) + (tid >> 1);
      dst[0] = pre.x;
      dst[NP] = pre.y;
      dst[2 * NP] = pre.z;
      dst[3 * NP] = pre.w;
    }
    __syncthreads();
    it += gridDim.x;
    if (it < items) {
      load_rel(it);                                                 // offsets arrived an iteration ago
      if (it + gridDim.x < items) load_offsets(it + gridDim.x);     // consumed an iteration from now
    }

    float2* Vs = reinterpret_cast<float2*>(V + static_cast<size_t>(cn0) * T * 2) + static_cast<size_t>(ct0) * cN;
    for (int i = tid; i < cN * F_WARPS; i += F_THREADS) {
      const int tt = i / cN, pp = i - tt * cN;
      Vs[i] = make_float2(xs[tt * NP + pp], ys[tt * NP + pp]);
    }
    {
      const float* xt = xs + warp * NP;
      const float* yt = ys + warp * NP;
      float* dst = dsm + warp * NP;
      float xk[NJ], yk[NJ], acc[NJ][2];
      bool kv[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        kv[j] = k < cN;
        xk[j] = kv[j] ? xt[k] : 0.f;
        yk[j] = kv[j] ? yt[k] : 0.f;
        acc[j][0] = acc[j][1] = 0.f;
      }
      // ---- pass 1: weights once, column sums on the fly (4 rows per iteration)
      float* wp = Wt + lane;
      int h = 0;
      float4 xn = *reinterpret_cast<const float4*>(xt), yn = *reinterpret_cast<const float4*>(yt);
#ifdef SSTG_G_NO_PASS1   // profiling knob (never in the shipped build): time the store phase alone
      h = cN;
#endif
#pragma unroll G_U1
      for (; h + 4 <= cN; h += 4) {
        const float4 xh = xn, yh = yn;  // loaded one group ahead (the read past row N stays inside the padded row)
        xn = *reinterpret_cast<const float4*>(xt + ((h + 4 < NP) ? h + 4 : h));
        yn = *reinterpret_cast<const float4*>(yt + ((h + 4 < NP) ? h + 4 : h));
        const float xa[4] = {xh.x, xh.y, xh.z, xh.w}, ya[4] = {yh.x, yh.y, yh.z, yh.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const float dx = xa[u] - xk[j], dy = ya[u] - yk[j];
            const float w = inv_dist(fmaf(dx, dx, dy * dy));
            acc[j][u & 1] += w;
            if (kv[j]) wp[32 * j] = w;
          }
          wp += cN;
        }
      }
      for (; h < cN; ++h) {
        const float xh = xt[h], yh = yt[h];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float dx = xh - xk[j], dy = yh - yk[j];
          const float w = inv_dist(fmaf(dx, dx, dy * dy));
          acc[j][0] += w;
          if (kv[j]) wp[32 * j] = w;
        }
        wp += cN;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        const float lkk = acc[j][0] + acc[j][1];  // the diagonal contributed w(0) = 0: d = 1 + off-diagonal sum
        if (kv[j]) {
          dst[k] = rsqrt_fast(1.f + lkk);
          Wt[k * cN + k] = -lkk;  // so that -(W_kk) ds_k ds_k = (d_k - 1) / d_k
        }
      }
      __syncwarp();
      // ---- pass 2: A_hk = -(w_hk ds_h) ds_k, streamed out
      float* At = A + cao + static_cast<size_t>(ct0 + warp) * cN * cN;
      if ((cN & 3) == 0) {
        const int lpr = cN >> 2;           // 16-byte chunks per row (<= 16 for N <= 64)
        const int rpp = 32 / lpr;          // rows per warp pass
        const int rr = lane / lpr, c4 = (lane - rr * lpr) * 4;
        const float4 dk = *reinterpret_cast<const float4*>(dst + c4);
        if (rr < rpp) {
          const float* wsrc = Wt + rr * cN + c4;
          float* adst = At + static_cast<size_t>(rr) * cN + c4;
          const int stride = rpp * cN;
          int hh = rr;
          for (; hh + 3 * rpp < cN; hh += 4 * rpp) {
            float4 w4[4];
            float dh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              w4[u] = *reinterpret_cast<const float4*>(wsrc + u * stride);
              dh[u] = dst[hh + u * rpp];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float4 o;
              o.x = fmaf(-(w4[u].x * dh[u]), dk.x, 0.f);
              o.y = fmaf(-(w4[u].y * dh[u]), dk.y, 0.f);
              o.z = fmaf(-(w4[u].z * dh[u]), dk.z, 0.f);
              o.w = fmaf(-(w4[u].w * dh[u]), dk.w, 0.f);
#ifdef SSTG_G_NO_STORE   // profiling knob (never in the shipped build): time the arithmetic alone
              if (o.x == 123.456f)
#endif
              *reinterpret_cast<float4*>(adst + u * stride) = o;  // plain stores: st.global.cs (evict-first) measured 3 % slower
            }
            wsrc += 4 * stride;
            adst += 4 * stride;
          }
          for (; hh < cN; hh += rpp) {
            const float4 w4 = *reinterpret_cast<const float4*>(wsrc);
            const float dh = dst[hh];
            float4 o;
            o.x = fmaf(-(w4.x * dh), dk.x, 0.f);
            o.y = fmaf(-(w4.y * dh), dk.y, 0.f);
            o.z = fmaf(-(w4.z * dh), dk.z, 0.f);
            o.w = fmaf(-(w4.w * dh), dk.w, 0.f);
            *reinterpret_cast<float4*>(adst) = o;
            wsrc += stride;
            adst += stride;
          }
        }
      } else {
        float dk[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) dk[j] = kv[j] ? dst[lane + 32 * j] : 0.f;
        for (int hh = 0; hh < cN; ++hh) {
          const float dh = dst[hh];
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int k = lane + 32 * j;
            if (kv[j]) At[static_cast<size_t>(hh) * cN + k] = fmaf(-(Wt[hh * cN + k] * dh), dk[j], 0.f);
          }
        }
      }
    }
    __syncthreads();
  }
}

// The same work with WARP-level items (scene, timestep): no CTA barrier anywhere.  In the kernel above the four
// warps of a CTA stage their inputs together, so they also run pass 1 (arithmetic only) and pass 2 (the store
// burst) together, and an SM holds just three such groups: built without stores the kernel takes 2.81 ms, without
// pass 1 3.02 ms, with both 3.55 ms -- the phases overlap only partly.  Here every warp fetches its own
// timestep's displacements (4-byte loads at a 2 T float stride: the 8 timesteps of a pedestrian share one
// 32-byte sector, so DRAM traffic is unchanged and the repeats hit L2), prefetches the next item into registers
// and walks its items independently: 12 actors per SM instead of 3.  Arithmetic and results are identical.
template <int NJ>
__global__ void __launch_bounds__(F_THREADS)
graph_build_warp_kernel(const float* __restrict__ rel, const int32_t* __restrict__ seg_off,
                        const int64_t* __restrict__ a_off, int S, int T, int NP, int wtile,
                        float* __restrict__ V, float* __restrict__ A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* xs = reinterpret_cast<float*>(smem_raw) + warp * NP;                      // [warp][NP]
  float* ys = reinterpret_cast<float*>(smem_raw) + (F_WARPS + warp) * NP;          // [warp][NP]
  float* dst = reinterpret_cast<float*>(smem_raw) + (2 * F_WARPS + warp) * NP;     // [warp][NP] ds
  float* Wt = reinterpret_cast<float*>(smem_raw) + 3 * F_WARPS * NP + static_cast<size_t>(warp) * wtile;  // [warp][N*N]
  pdl_wait();
  pdl_release();
  const long long items = static_cast<long long>(S) * T;
  const long long W = static_cast<long long>(gridDim.x) * F_WARPS;
  long long it = static_cast<long long>(blockIdx.x) * F_WARPS + warp;
  // Software pipeline per warp, in registers.  Under the store stream a global load takes 6-7 us -- longer than
  // one item (~6 us) -- so the displacements run TWO items ahead and the offsets they need two more.
  struct Meta {
    int t, n0, N;
    long long ao;
  };
  auto load_offsets = [&](long long item, Meta& o) {
    if (item < items) {
      const int sc = static_cast<int>(item / T);
      o.t = static_cast<int>(item - static_cast<long long>(sc) * T);
      o.n0 = seg_off[sc];
      o.N = seg_off[sc + 1] - o.n0;
      o.ao = a_off[sc];
    }
  };
  auto load_rel = [&](long long item, const Meta& m, float (&x)[NJ], float (&y)[NJ]) {  // offsets of `item` are in `m`
    if (item < items) {
      const float* r = rel + static_cast<size_t>(m.n0) * 2 * T + m.t;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        if (k < m.N) {
          x[j] = r[static_cast<size_t>(k) * 2 * T];
          y[j] = r[static_cast<size_t>(k) * 2 * T + T];
        }
      }
    }
  };
  Meta m0 = {0, 0, 0, 0}, m1 = m0, m2 = m0, o3 = m0, o4 = m0;   // items it, it+W, it+2W (data), it+3W, it+4W (offsets)
  float px0[NJ], py0[NJ], px1[NJ], py1[NJ], px2[NJ], py2[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) px0[j] = py0[j] = px1[j] = py1[j] = px2[j] = py2[j] = 0.f;
  load_offsets(it, m0);
  load_offsets(it + W, m1);
  load_offsets(it + 2 * W, m2);
  load_offsets(it + 3 * W, o3);
  load_offsets(it + 4 * W, o4);
  load_rel(it, m0, px0, py0);
  load_rel(it + W, m1, px1, py1);
  load_rel(it + 2 * W, m2, px2, py2);
  int ct = 0, cn0 = 0, cN = 0;
  long long cao = 0;
  while (it < items) {
    ct = m0.t;
    cn0 = m0.n0;
    cN = m0.N;
    cao = m0.ao;
    // stage this warp's x / y rows and write the node tensor V[t] (utils.py:42)
    float2* Vs = reinterpret_cast<float2*>(V + static_cast<size_t>(cn0) * T * 2) + static_cast<size_t>(ct) * cN;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = lane + 32 * j;
      if (k < cN) {
        xs[k] = px0[j];
        ys[k] = py0[j];
        Vs[k] = make_float2(px0[j], py0[j]);
      }
    }
    __syncwarp();
    // rotate; request the displacements of item it + 3W (offsets two trips old) and the offsets of it + 5W
    m0 = m1;
    m1 = m2;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      px0[j] = px1[j], py0[j] = py1[j];
      px1[j] = px2[j], py1[j] = py2[j];
    }
    m2 = o3;
    load_rel(it + 3 * W, m2, px2, py2);
    o3 = o4;
    load_offsets(it + 5 * W, o4);
    it += W;
    const float* xt = xs;
    const float* yt = ys;
    const int ct0 = ct - warp;   // the shared body below addresses its slice as ct0 + warp
    {
      float xk[NJ], yk[NJ], acc[NJ][2];
      bool kv[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        kv[j] = k < cN;
        xk[j] = kv[j] ? xt[k] : 0.f;
        yk[j] = kv[j] ? yt[k] : 0.f;
        acc[j][0] = acc[j][1] = 0.f;
      }
      // ---- pass 1: weights once, column sums on the fly (4 rows per iteration)
      float* wp = Wt + lane;
      int h = 0;
      float4 xn = *reinterpret_cast<const float4*>(xt), yn = *reinterpret_cast<const float4*>(yt);
#ifdef SSTG_G_NO_PASS1   // profiling knob (never in the shipped build): time the store phase alone
      h = cN;
#endif
#pragma unroll G_U1
      for (; h + 4 <= cN; h += 4) {
        const float4 xh = xn, yh = yn;  // loaded one group ahead (the read past row N stays inside the padded row)
        xn = *reinterpret_cast<const float4*>(xt + ((h + 4 < NP) ? h + 4 : h));
        yn = *reinterpret_cast<const float4*>(yt + ((h + 4 < NP) ? h + 4 : h));
        const float xa[4] = {xh.x, xh.y, xh.z, xh.w}, ya[4] = {yh.x, yh.y, yh.z, yh.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const float dx = xa[u] - xk[j], dy = ya[u] - yk[j];
            const float w = inv_dist(fmaf(dx, dx, dy * dy));
            acc[j][u & 1] += w;
            if (kv[j]) wp[32 * j] = w;
          }
          wp += cN;
        }
      }
      for (; h < cN; ++h) {
        const float xh = xt[h], yh = yt[h];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float dx = xh - xk[j], dy = yh - yk[j];
          const float w = inv_dist(fmaf(dx, dx, dy * dy));
          acc[j][0] += w;
          if (kv[j]) wp[32 * j] = w;
        }
        wp += cN;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int k = lane + 32 * j;
        const float lkk = acc[j][0] + acc[j][1];  // the diagonal contributed w(0) = 0: d = 1 + off-diagonal sum
        if (kv[j]) {
          dst[k] = rsqrt_fast(1.f + lkk);
          Wt[k * cN + k] = -lkk;  // so that -(W_kk) ds_k ds_k = (d_k - 1) / d_k
        }
      }
      __syncwarp();
      // ---- pass 2: A_hk = -(w_hk ds_h) ds_k, streamed out
      float* At = A + cao + static_cast<size_t>(ct0 + warp) * cN * cN;
      if ((cN & 3) == 0) {
        const int lpr = cN >> 2;           // 16-byte chunks per row (<= 16 for N <= 64)
        const int rpp = 32 / lpr;          // rows per warp pass
        const int rr = lane / lpr, c4 = (lane - rr * lpr) * 4;
        const float4 dk = *reinterpret_cast<const float4*>(dst + c4);
        if (rr < rpp) {
          const float* wsrc = Wt + rr * cN + c4;
          float* adst = At + static_cast<size_t>(rr) * cN + c4;
          const int stride = rpp * cN;
          int hh = rr;
          for (; hh + 3 * rpp < cN; hh += 4 * rpp) {
            float4 w4[4];
            float dh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              w4[u] = *reinterpret_cast<const float4*>(wsrc + u * stride);
              dh[u] = dst[hh + u * rpp];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float4 o;
              o.x = fmaf(-(w4[u].x * dh[u]), dk.x, 0.f);
              o.y = fmaf(-(w4[u].y * dh[u]), dk.y, 0.f);
              o.z = fmaf(-(w4[u].z * dh[u]), dk.z, 0.f);
              o.w = fmaf(-(w4[u].w * dh[u]), dk.w, 0.f);
#ifdef SSTG_G_NO_STORE   // profiling knob (never in the shipped build): time the arithmetic alone
              if (o.x == 123.456f)
#endif
              *reinterpret_cast<float4*>(adst + u * stride) = o;  // plain stores: st.global.cs (evict-first) measured 3 % slower
            }
            wsrc += 4 * stride;
            adst += 4 * stride;
          }
          for (; hh < cN; hh += rpp) {
            const float4 w4 = *reinterpret_cast<const float4*>(wsrc);
            const float dh = dst[hh];
            float4 o;
            o.x = fmaf(-(w4.x * dh), dk.x, 0.f);
            o.y = fmaf(-(w4.y * dh), dk.y, 0.f);
            o.z = fmaf(-(w4.z * dh), dk.z, 0.f);
            o.w = fmaf(-(w4.w * dh), dk.w, 0.f);
            *reinterpret_cast<float4*>(adst) = o;
            wsrc += stride;
            adst += stride;
          }
        }
      } else {
        float dk[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) dk[j] = kv[j] ? dst[lane + 32 * j] : 0.f;
        for (int hh = 0; hh < cN; ++hh) {
          const float dh = dst[hh];
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int k = lane + 32 * j;
            if (kv[j]) At[static_cast<size_t>(hh) * cN + k] = fmaf(-(Wt[hh * cN + k] * dh), dk[j], 0.f);
          }
        }
      }
    }
    __syncwarp();   // the next item overwrites this warp's rows and tile
  }
}

template <int NJ>
int launch_fast(const float* rel, const int32_t* seg_off, const int64_t* a_off, int S, int T, int max_n, float* V,
                float* A, cudaStream_t st) {
  const int NP = (max_n + 3) & ~3;
  const int wtile = (max_n * max_n + 3) & ~3;
  const size_t smem = sizeof(float) * (static_cast<size_t>(3 * F_WARPS * NP) + static_cast<size_t>(F_WARPS) * wtile);
  const char* coop = getenv("SSTG_GRAPH_COOP");   // A/B knob: 1 = the CTA-cooperative kernel (4 timesteps per item)
  auto kern = (coop && coop[0] == '1') ? graph_build_fast_kernel<NJ> : graph_build_warp_kernel<NJ>;
  if (smem > 48 * 1024)
    SSTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                  "graph_build_fast smem attr");
  int dev = 0, sms = 148;
  SSTG_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
  SSTG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "sm count");
  int per_sm = 1;
  SSTG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, F_THREADS, smem), "occupancy");
  if (per_sm < 1) per_sm = 1;
  const long long items = static_cast<long long>(S) * (T / F_WARPS);
  const long long want = static_cast<long long>(sms) * per_sm;
  const int grid = static_cast<int>(items < want ? items : want);
  SSTG_CUDA_TRY(launch_pdl(kern, grid, F_THREADS, smem, st, rel, seg_off, a_off, S, T, NP, wtile, V, A),
                "graph_build_fast launch");
  return SSTG_OK;
}

template <bool F64, int NJ>
int launch_nj(const float* rel, const int32_t* seg_off, const int64_t* a_off, int S, int T, int NP, float* V,
              float* A, int flags, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(2 * T * NP) * sizeof(float) +
                      static_cast<size_t>(G_WARPS * NP) * (F64 ? sizeof(double) : sizeof(float));
  auto kern = graph_build_kernel<F64, NJ>;
  if (smem > 48 * 1024)
    SSTG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                  "graph_build smem attr");
  int dev = 0, sms = 148;
  SSTG_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
  SSTG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "sm count");
  int per_sm = 1;
  SSTG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G_THREADS, smem), "occupancy");
  if (per_sm < 1) per_sm = 1;
  const long long want = static_cast<long long>(sms) * per_sm;
  const int grid = static_cast<int>(S < want ? S : want);
  SSTG_CUDA_TRY(launch_pdl(kern, grid, G_THREADS, smem, st, rel, seg_off, a_off, S, T, NP, V, A, flags),
                "graph_build launch");
  return SSTG_OK;
}

template <bool F64>
int launch_f(const float* rel, const int32_t* seg_off, const int64_t* a_off, int S, int T, int max_n, float* V,
             float* A, int flags, cudaStream_t st) {
  const int NP = (max_n + 3) & ~3;
  if (max_n <= 32) return launch_nj<F64, 1>(rel, seg_off, a_off, S, T, NP, V, A, flags, st);
  if (max_n <= 64) return launch_nj<F64, 2>(rel, seg_off, a_off, S, T, NP, V, A, flags, st);
  if (max_n <= 128) return launch_nj<F64, 4>(rel, seg_off, a_off, S, T, NP, V, A, flags, st);
  return launch_nj<F64, 8>(rel, seg_off, a_off, S, T, NP, V, A, flags, st);
}

}  // namespace

int launch_graph_build(const float* rel, const int32_t* seg_off, const int64_t* a_off, int S, int T, int max_n,
                       float* V, float* A, int flags, cudaStream_t st) {
  if (S == 0) return SSTG_OK;
  if (flags & SSTG_GRAPH_FP64) return launch_f<true>(rel, seg_off, a_off, S, T, max_n, V, A, flags, st);
  if ((flags & SSTG_GRAPH_NORM_LAP) && !(flags & SSTG_GRAPH_V_ONLY) && max_n <= 64 && (T & 3) == 0) {
    if (max_n <= 32) return launch_fast<1>(rel, seg_off, a_off, S, T, max_n, V, A, st);
    return launch_fast<2>(rel, seg_off, a_off, S, T, max_n, V, A, st);
  }
  return launch_f<false>(rel, seg_off, a_off, S, T, max_n, V, A, flags, st);
}

}  // namespace sstg
