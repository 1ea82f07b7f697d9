// Kernels (4a) and (4b): bivariate-Gaussian NLL (forward) and K-sample trajectory
// sampling with per-pedestrian ADE/FDE, min over the samples.
//
// (4b) replaces the per-scene body of reference test() (test.py:59-123) together with
// metrics.ade / fde / seq_to_nodes / nodes_rel_to_nodes_abs (metrics.py:21-75):
//   sx = exp(p2), sy = exp(p3), rho = tanh(p4)                                  test.py:59-61
//   cov = [[sx*sx, (rho*sx)*sy], [., sy*sy]],  L = chol(cov)                    test.py:63-70
//   sample_k[t] = mean[t] + L[t] eps_k[t]                                       test.py:98
//   abs_k[t] = fp32( fp64(cumsum_fp32(sample_k)[t]) + last_obs )                metrics.py:69-75
//   ADE_k = (1/12) sum_t ||abs_k[t] - gt_abs[t]||,  FDE_k = ||.|| at t = 11      metrics.py:21-53
//   min over k independently for ADE and FDE, per pedestrian                    test.py:121-123
// The work is embarrassingly parallel over pedestrians, so the kernel ignores scene
// boundaries except for addressing: a CTA takes 64 consecutive pedestrians of the packed
// batch; per (t, pedestrian) the distribution parameters and the ground-truth absolute
// position are computed once into shared memory; then 4 warps-pairs split the K samples
// (thread = (sample class k mod 4, pedestrian)), so a warp reads 32 consecutive
// pedestrians of one (k, t) row of the noise tensor -- fully coalesced float2 loads --
// or draws the noise in-kernel with Philox4x32-10 + Box-Muller.
//
// (4a) replaces metrics.bivariate_loss (metrics.py:84-113), one warp per scene.
#include <type_traits>

#include "sstg_common.cuh"

namespace sstg {

namespace {

#ifndef SSTG_E_KU
#define SSTG_E_KU 5
#endif
constexpr int E_KU = SSTG_E_KU;  // samples a thread advances together: 5 (K = 20 -> 4 sample classes, 128 threads) or 4 (5 classes,
                                 // 160 threads: no unpaired fifth sample, every trajectory operation packed); A/B knob -DSSTG_E_KU=4|5
static_assert(E_KU == 4 || E_KU == 5, "E_KU");
constexpr int E_PEDS = 32;  // pedestrians per CTA: small CTAs (128 threads, 64 registers in the throughput form) so that
                            // 8 fit an SM and one CTA's latency-bound prologue overlaps the sampling loops of the others
constexpr int E_SUB = 20 / E_KU;
constexpr int E_THREADS = E_PEDS * E_SUB;
constexpr int E_CTAS = E_KU == 5 ? 8 : 6;   // CTAs per SM of the throughput form at 64 registers

struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __device__ static __forceinline__ uint4 run(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
      const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
      c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
      k.x += W0;
      k.y += W1;
    }
    return c;
  }
};

// two standard normals from two uint32 (Box-Muller).  u1 in (0,1], u2 in [0,1) are built from
// the top 23 bits with the exponent trick (no int->float conversion), the transcendental part
// is three MUFU ops (lg2, sqrt, sin/cos approx) -- the noise only has to be N(0,1), and
// tests/philox_ref.py replays exactly this recipe.
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// u1 is in [2^-23, 1]: never denormal, so the .ftz form returns the same bits as __log2f without its
// denormal-input fix-up (FSETP + FMUL + FADD per call)
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed FP32 (sm_100 FADD2 / FMUL2 / FFMA2): two IEEE fp32 operations per issue slot, each component rounded
// exactly like the scalar instruction -- the sampling loop advances two samples of a pedestrian per instruction
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)),
        "l"(reinterpret_cast<const unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float2 bc2(const float v) { return make_float2(v, v); }

// The two Box-Muller transforms of one Philox output, (r.x, r.y) -> e0 and (r.z, r.w) -> e1, with their arithmetic
// packed pairwise: u1 = 2 - f(a) in (0, 1], u2 = f(b) - 1 in [0, 1) with f(w) = as_float((w >> 9) | 0x3f800000),
// radius sqrt(-2 ln u1) = sqrt(-1.386... * log2 u1), angle 2 pi u2, e = radius * (cos, sin).  Every operation is
// rounded exactly as in the scalar recipe tests/philox_ref.py replays (2 - x == fma(x, -1, 2) exactly).
__device__ __forceinline__ void box_muller2(const uint4 r, float2& e0, float2& e1) {
  const float2 fa = make_float2(__uint_as_float((r.x >> 9) | 0x3f800000u), __uint_as_float((r.z >> 9) | 0x3f800000u));
  const float2 fb = make_float2(__uint_as_float((r.y >> 9) | 0x3f800000u), __uint_as_float((r.w >> 9) | 0x3f800000u));
  const float2 u1 = fma2(fa, bc2(-1.0f), bc2(2.0f));                 // (0, 1]
  const float2 u2 = add2(fb, bc2(-1.0f));                            // [0, 1)
  const float2 t = mul2(bc2(-1.3862943611198906f), make_float2(lg2_approx(u1.x), lg2_approx(u1.y)));
  const float2 rr = make_float2(sqrt_approx(t.x), sqrt_approx(t.y));  // sqrt(-2 ln u1)
  const float2 ang = mul2(bc2(6.283185307179586f), u2);
  float sn0, cs0, sn1, cs1;
  __sincosf(ang.x, &sn0, &cs0);
  __sincosf(ang.y, &sn1, &cs1);
  e0 = mul2(bc2(rr.x), make_float2(cs0, sn0));
  e1 = mul2(bc2(rr.y), make_float2(cs1, sn1));
}

// The same transform for the first (SECOND = false: words x, y) or second (words z, w) normal pair of TWO Philox
// outputs at once, packed ACROSS the two draws: ex = (e.x of draw a, e.x of draw b), ey likewise -- the layout the
// packed trajectory arithmetic consumes, so no register-pair shuffling sits between the two.  Same operations,
// same roundings per component as box_muller2.
template <bool SECOND>
__device__ __forceinline__ void box_muller_pair(const uint4 ra, const uint4 rb, float2& ex, float2& ey) {
  const uint32_t a0 = SECOND ? ra.z : ra.x, a1 = SECOND ? ra.w : ra.y;
  const uint32_t b0 = SECOND ? rb.z : rb.x, b1 = SECOND ? rb.w : rb.y;
  const float2 fa = make_float2(__uint_as_float((a0 >> 9) | 0x3f800000u), __uint_as_float((b0 >> 9) | 0x3f800000u));
  const float2 fb = make_float2(__uint_as_float((a1 >> 9) | 0x3f800000u), __uint_as_float((b1 >> 9) | 0x3f800000u));
  const float2 u1 = fma2(fa, bc2(-1.0f), bc2(2.0f));
  const float2 u2 = add2(fb, bc2(-1.0f));
  const float2 t = mul2(bc2(-1.3862943611198906f), make_float2(lg2_approx(u1.x), lg2_approx(u1.y)));
  const float2 rr = make_float2(sqrt_approx(t.x), sqrt_approx(t.y));
  const float2 ang = mul2(bc2(6.283185307179586f), u2);
  float sna, csa, snb, csb;
  __sincosf(ang.x, &sna, &csa);
  __sincosf(ang.y, &snb, &csb);
  ex = mul2(rr, make_float2(csa, csb));
  ey = mul2(rr, make_float2(sna, snb));
}

// scene of packed pedestrian p: largest s with seg_off[s] <= p
__device__ __forceinline__ int find_scene(const int32_t* __restrict__ seg_off, int S, int p) {
  int lo = 0, hi = S;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (seg_off[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

// Of the two sample pairs (0,1), (2,3) of a thread, how many advance as PACKED FP32 pairs in the throughput form (the rest,
// and sample 4, as scalar chains; same roundings either way).  ncu shows the heavy half of the FMA pipe as the busiest unit
// (74 %: Philox's IMAD.WIDE plus the packed FP32), but un-packing does not relieve it: measured 2.13 / 2.18 / 2.24 ms per
// 65,536 scenes x 64 for 2 / 1 / 0 packed pairs (profiles/r1an_ab_traj_packed.txt).  A/B knob: -DSSTG_E_TRAJ_PACKED=0|1|2.
#ifndef SSTG_E_TRAJ_PACKED
#define SSTG_E_TRAJ_PACKED 2
#endif
constexpr int E_TRAJ_PACKED = SSTG_E_TRAJ_PACKED;

// EXACT  : the reference's rounding points (separately rounded mean + L eps, fp64 sqrt and
//          fp64 sequential sum of the 12 distances, metrics.py:33-34,50); otherwise fp32 with FMA.
// PHILOX : draw the noise in-kernel instead of reading eps.
// EXTRA  : eps_out / pred_abs requested (raw_data_dict, test.py:105).
template <bool EXACT, bool PHILOX, bool EXTRA>
__global__ void __launch_bounds__(E_THREADS, (EXACT || EXTRA) ? 4 : E_CTAS)
sample_ade_fde_kernel(const float* __restrict__ Vpred, const float* __restrict__ Vtr,
                      const float* __restrict__ last_obs, const int32_t* __restrict__ seg_off, int S, int K,
                      const float* __restrict__ eps, uint64_t seed, float* __restrict__ eps_out,
                      double* __restrict__ min_ade, double* __restrict__ min_fde, float* __restrict__ pred_abs,
                      int vtr_is_rel, int uniform_n) {
  using Acc = typename std::conditional<EXACT, double, float>::type;
  // [t][field][ped]: mux, muy, L00, L10, L11, gtx, gty
  __shared__ float par[T_PRED][7][E_PEDS];
  __shared__ Acc red[2][E_SUB][E_PEDS];
  __shared__ int sc_n0[E_PEDS], sc_N[E_PEDS];
  const int tid = threadIdx.x;
  const int P = seg_off[S];
  const int p_base = blockIdx.x * E_PEDS;

  // ---- scene of every pedestrian of the block.  Warp 0 finds the scene of the block's first
  // pedestrian with a 32-ary search (4 dependent load rounds for 1M scenes instead of 20); the 64
  // pedestrians then need at most 6 more steps inside [s0, s0 + 64), or none at all when the
  // whole block lies in one scene (the dense-crowd case).
  __shared__ int s0_sh;
  if (uniform_n > 0) {  // every scene has uniform_n pedestrians (dense synthetic crowds): no search at all
    if (tid < E_PEDS) {
      const int p = p_base + tid;
      sc_n0[tid] = p < P ? (p / uniform_n) * uniform_n : 0;
      sc_N[tid] = p < P ? uniform_n : 0;
    }
  } else {
    if (tid < 32) {
      int lo = 0, hi = S;  // seg_off[lo] <= p_base < seg_off[hi]
      while (hi - lo > 1) {
        const int step = (hi - lo + 31) >> 5;
        const int c = lo + tid * step;
        const bool le = c < hi && seg_off[c] <= p_base;
        const unsigned m = __ballot_sync(0xffffffffu, le);
        const int last = 31 - __clz(m);  // lane 0 always votes (c == lo)
        lo += last * step;
        hi = (lo + step < hi) ? lo + step : hi;
      }
      if (tid == 0) s0_sh = lo;
    }
    __syncthreads();
    if (tid < E_PEDS) {
      const int p = p_base + tid;
      if (p < P) {
        int lo = s0_sh, hi = (lo + E_PEDS < S) ? lo + E_PEDS : S;
        if (seg_off[lo + 1] > p) {
          hi = lo + 1;
        } else {
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (seg_off[mid] <= p) lo = mid; else hi = mid;
          }
        }
        sc_n0[tid] = seg_off[lo];
        sc_N[tid] = seg_off[lo + 1] - seg_off[lo];
      } else {
        sc_n0[tid] = 0;
        sc_N[tid] = 0;
      }
    }
  }
  __syncthreads();
  // ---- distribution parameters per (t, pedestrian) and the ground-truth absolute trajectory.  All global
  // loads of the prologue -- 15 parameter loads per thread, and for the first warp the pedestrian's 24
  // ground-truth displacements and start position -- are issued before any of the arithmetic, so that a CTA
  // pays ONE memory latency here, not two.
  {
    constexpr int PER = (T_PRED * E_PEDS + E_THREADS - 1) / E_THREADS;  // 3
    float in[PER][5];
    bool ok[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = tid + u * E_THREADS;
      const bool inr = i < T_PRED * E_PEDS;
      const int t = inr ? i / E_PEDS : 0, pl = inr ? i - t * E_PEDS : 0;
      const int N = sc_N[pl];
      ok[u] = inr && N > 0;
      const int n0 = sc_n0[pl], nl = p_base + pl - n0;
      const float* vp = Vpred + static_cast<size_t>(n0) * 60 + t * N + nl;  // (c, t, n) at (c*12 + t)*N + n
#pragma unroll
      for (int c = 0; c < 5; ++c) in[u][c] = ok[u] ? vp[static_cast<size_t>(c) * 12 * N] : 0.f;
    }
    const bool gt_thread = tid < E_PEDS && sc_N[tid < E_PEDS ? tid : 0] > 0;
    float rx[T_PRED], ry[T_PRED];
    float2 gt_init = make_float2(0.f, 0.f);
    if (gt_thread) {
      const int N = sc_N[tid], n0 = sc_n0[tid], nl = p_base + tid - n0;
      gt_init = reinterpret_cast<const float2*>(last_obs)[p_base + tid];
      if (vtr_is_rel) {  // Vtr is pred_traj_rel [P,2,12]: this pedestrian's 24 floats are contiguous
        const float4* pr = reinterpret_cast<const float4*>(Vtr + static_cast<size_t>(p_base + tid) * (2 * T_PRED));
#pragma unroll
        for (int i = 0; i < T_PRED / 4; ++i) {
          const float4 a = pr[i], b = pr[T_PRED / 4 + i];
          rx[4 * i] = a.x, rx[4 * i + 1] = a.y, rx[4 * i + 2] = a.z, rx[4 * i + 3] = a.w;
          ry[4 * i] = b.x, ry[4 * i + 1] = b.y, ry[4 * i + 2] = b.z, ry[4 * i + 3] = b.w;
        }
      } else {
        const float2* vt = reinterpret_cast<const float2*>(Vtr) + static_cast<size_t>(n0) * T_PRED + nl;
#pragma unroll
        for (int t = 0; t < T_PRED; ++t) {
          const float2 v = vt[t * N];
          rx[t] = v.x;
          ry[t] = v.y;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      if (!ok[u]) continue;
      const int i = tid + u * E_THREADS;
      const int t = i / E_PEDS, pl = i - t * E_PEDS;
      const float sx = expf(in[u][2]), sy = expf(in[u][3]), rho = tanhf(in[u][4]);
      const float a = __fmul_rn(sx, sx);
      const float b = __fmul_rn(__fmul_rn(rho, sx), sy);
      const float c = __fmul_rn(sy, sy);
      const float L00 = sqrtf(a);                // potrf
      const float L10 = __fdiv_rn(b, L00);
      const float L11 = sqrtf(fmaf(-L10, L10, c));
      par[t][0][pl] = in[u][0];
      par[t][1][pl] = in[u][1];
      par[t][2][pl] = L00;
      par[t][3][pl] = L10;
      par[t][4][pl] = L11;
    }
    // ground-truth absolute trajectory: sequential fp32 prefix sum, then the start position.
    // metrics.py:69-75 adds the start in fp64 and stores fp32; fp32(fp64(a) + fp64(b)) equals the
    // fp32 sum (double rounding through a 53 >= 2*24+2 bit format is innocuous), so __fadd_rn is bit-exact.
    if (gt_thread) {
      float cx = 0.f, cy = 0.f;
#pragma unroll
      for (int t = 0; t < T_PRED; ++t) {
        cx = __fadd_rn(cx, rx[t]);
        cy = __fadd_rn(cy, ry[t]);
        par[t][5][tid] = __fadd_rn(cx, gt_init.x);
        par[t][6][tid] = __fadd_rn(cy, gt_init.y);
      }
    }
  }
  __syncthreads();

  const int sub = tid / E_PEDS, pl = tid - sub * E_PEDS;
  const int N = sc_N[pl];
  Acc best_a = Acc(3.0e38f), best_f = Acc(3.0e38f);
  if (N > 0) {
    const int p = p_base + pl;
    const int n0 = sc_n0[pl], nl = p - n0;
    const float2 init = reinterpret_cast<const float2*>(last_obs)[p];
    const size_t eb = static_cast<size_t>(n0) * K * (T_PRED * 2);  // scene block (K,12,N,2)
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    for (int kb = sub; kb < K; kb += E_SUB * E_KU) {
      float cx[E_KU], cy[E_KU];
      Acc acc[E_KU], fd[E_KU];
#pragma unroll
      for (int i = 0; i < E_KU; ++i) {
        cx[i] = cy[i] = 0.f;
        acc[i] = fd[i] = Acc(0);
      }
      // throughput path (fp32 distances, nothing extra to store): samples (0,1) and (2,3) of the thread advance
      // as packed pairs, sample 4 alone -- bit-identical to the scalar loop, ~10 % fewer instructions
      constexpr bool PACK = !EXACT && !EXTRA;
      float2 cxp[2], cyp[2], accp[2], fdp[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) cxp[q] = cyp[q] = accp[q] = fdp[q] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int tp = 0; tp < T_PRED / 2; ++tp) {  // not unrolled: keeps the 84 shared-memory parameter loads from being hoisted
        float2 e2[2][E_KU];
        float2 exq[2][2], eyq[2][2];  // PACK && PHILOX: noise of the sample pairs (0,1), (2,3), packed across the pair
        auto draw = [&](int i) {
          const int k = kb + E_SUB * i;
          return Philox::run(make_uint4(static_cast<uint32_t>(k * (T_PRED / 2) + tp), 0u, static_cast<uint32_t>(p), 0u),
                             key);
        };
        if constexpr (PACK && PHILOX) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint4 ra = draw(2 * q), rb = draw(2 * q + 1);
            box_muller_pair<false>(ra, rb, exq[0][q], eyq[0][q]);
            box_muller_pair<true>(ra, rb, exq[1][q], eyq[1][q]);
          }
          if constexpr (E_KU == 5) box_muller2(draw(4), e2[0][4], e2[1][4]);
        } else {
#pragma unroll
          for (int i = 0; i < E_KU; ++i) {
            const int k = kb + E_SUB * i;
            if (PHILOX) {
              box_muller2(draw(i), e2[0][i], e2[1][i]);
            } else if (k < K) {
              const size_t ei = eb + (static_cast<size_t>(k * T_PRED + 2 * tp) * N + nl) * 2;
              e2[0][i] = *reinterpret_cast<const float2*>(eps + ei);
              e2[1][i] = *reinterpret_cast<const float2*>(eps + ei + static_cast<size_t>(N) * 2);
            } else {
              e2[0][i] = e2[1][i] = make_float2(0.f, 0.f);
            }
          }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int t = 2 * tp + hh;
          const float mux = par[t][0][pl], muy = par[t][1][pl], l00 = par[t][2][pl], l10 = par[t][3][pl],
                      l11 = par[t][4][pl], gx = par[t][5][pl], gy = par[t][6][pl];
          if constexpr (PACK) {
            const float2 ngx = bc2(-gx), ngy = bc2(-gy), ix = bc2(init.x), iy = bc2(init.y);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float2 ex = PHILOX ? exq[hh][q] : make_float2(e2[hh][2 * q].x, e2[hh][2 * q + 1].x);
              const float2 ey = PHILOX ? eyq[hh][q] : make_float2(e2[hh][2 * q].y, e2[hh][2 * q + 1].y);
              if (q < E_TRAJ_PACKED) {
                const float2 sx = fma2(bc2(l00), ex, bc2(mux));
                const float2 sy = fma2(bc2(l11), ey, fma2(bc2(l10), ex, bc2(muy)));
                cxp[q] = add2(cxp[q], sx);
                cyp[q] = add2(cyp[q], sy);
                const float2 dx = add2(add2(cxp[q], ix), ngx);   // (cx + init.x) - gx, each rounded like the scalar path
                const float2 dy = add2(add2(cyp[q], iy), ngy);
                const float2 s2 = fma2(dx, dx, mul2(dy, dy));
                const float2 d = make_float2(sqrt_approx(s2.x), sqrt_approx(s2.y));
                accp[q] = add2(accp[q], d);
                fdp[q] = d;
              } else {  // the same pair as two scalar chains (light half of the FMA pipe; see E_TRAJ_PACKED)
                cxp[q].x = __fadd_rn(cxp[q].x, fmaf(l00, ex.x, mux));
                cxp[q].y = __fadd_rn(cxp[q].y, fmaf(l00, ex.y, mux));
                cyp[q].x = __fadd_rn(cyp[q].x, fmaf(l11, ey.x, fmaf(l10, ex.x, muy)));
                cyp[q].y = __fadd_rn(cyp[q].y, fmaf(l11, ey.y, fmaf(l10, ex.y, muy)));
                const float dx0 = __fsub_rn(__fadd_rn(cxp[q].x, init.x), gx), dx1 = __fsub_rn(__fadd_rn(cxp[q].y, init.x), gx);
                const float dy0 = __fsub_rn(__fadd_rn(cyp[q].x, init.y), gy), dy1 = __fsub_rn(__fadd_rn(cyp[q].y, init.y), gy);
                const float2 d = make_float2(sqrt_approx(fmaf(dx0, dx0, __fmul_rn(dy0, dy0))),
                                             sqrt_approx(fmaf(dx1, dx1, __fmul_rn(dy1, dy1))));
                accp[q].x = __fadd_rn(accp[q].x, d.x);
                accp[q].y = __fadd_rn(accp[q].y, d.y);
                fdp[q] = d;
              }
            }
          }
#pragma unroll
          for (int i = PACK ? 4 : 0; i < E_KU; ++i) {
            const float2 e = e2[hh][i];
            float sxv, syv;
            if (EXACT) {  // sample = mean + L @ eps with the separate roundings of the bmm + add of rsample
              sxv = __fadd_rn(mux, __fmul_rn(l00, e.x));
              syv = __fadd_rn(muy, __fadd_rn(__fmul_rn(l10, e.x), __fmul_rn(l11, e.y)));
            } else {
              sxv = fmaf(l00, e.x, mux);
              syv = fmaf(l11, e.y, fmaf(l10, e.x, muy));
            }
            cx[i] = __fadd_rn(cx[i], sxv);
            cy[i] = __fadd_rn(cy[i], syv);
            const float ax = __fadd_rn(cx[i], init.x);
            const float ay = __fadd_rn(cy[i], init.y);
            if (EXTRA) {
              const int k = kb + E_SUB * i;
              if (k < K) {
                const size_t ei = eb + (static_cast<size_t>(k * T_PRED + t) * N + nl) * 2;
                if (eps_out != nullptr) *reinterpret_cast<float2*>(eps_out + ei) = e;
                if (pred_abs != nullptr) *reinterpret_cast<float2*>(pred_abs + ei) = make_float2(ax, ay);
              }
            }
            const float dx = __fsub_rn(ax, gx);
            const float dy = __fsub_rn(ay, gy);
            if (EXACT) {
              const double d = sqrt(static_cast<double>(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
              acc[i] += d;
              fd[i] = d;
            } else {
              const float d = sqrt_approx(fmaf(dx, dx, dy * dy));
              acc[i] += d;
              fd[i] = d;
            }
          }
        }
      }
      if constexpr (PACK) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          acc[2 * q] = accp[q].x, acc[2 * q + 1] = accp[q].y;
          fd[2 * q] = fdp[q].x, fd[2 * q + 1] = fdp[q].y;
        }
      }
#pragma unroll
      for (int i = 0; i < E_KU; ++i)
        if (kb + E_SUB * i < K) {
          const Acc ade = acc[i] / Acc(T_PRED);
          best_a = ade < best_a ? ade : best_a;
          best_f = fd[i] < best_f ? fd[i] : best_f;
        }
    }
  }
  red[0][sub][pl] = best_a;
  red[1][sub][pl] = best_f;
  __syncthreads();
  if (tid < E_PEDS && sc_N[tid] > 0) {
    Acc a = red[0][0][tid], f = red[1][0][tid];
#pragma unroll
    for (int u = 1; u < E_SUB; ++u) {
      a = red[0][u][tid] < a ? red[0][u][tid] : a;
      f = red[1][u][tid] < f ? red[1][u][tid] : f;
    }
    min_ade[p_base + tid] = static_cast<double>(a);
    min_fde[p_base + tid] = static_cast<double>(f);
  }
}

// ---- (4a) NLL forward: one warp per scene, metrics.py:84-113 term by term
__global__ void __launch_bounds__(256)
nll_forward_kernel(const float* __restrict__ Vpred, const float* __restrict__ Vtr,
                   const int32_t* __restrict__ seg_off, int S, float* __restrict__ loss) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= S) return;
  const int n0 = seg_off[warp], N = seg_off[warp + 1] - n0;
  const float* vp = Vpred + static_cast<size_t>(n0) * 60;
  const float2* vt = reinterpret_cast<const float2*>(Vtr) + static_cast<size_t>(n0) * T_PRED;
  float sum = 0.f;
  const int tot = T_PRED * N;
  for (int i = lane; i < tot; i += 32) {  // i = t*N + n
    const float2 tr = vt[i];
    const float normx = tr.x - vp[i];
    const float normy = tr.y - vp[12 * N + i];
    const float sx = expf(vp[24 * N + i]);
    const float sy = expf(vp[36 * N + i]);
    const float corr = tanhf(vp[48 * N + i]);
    const float sxsy = sx * sy;
    const float qx = normx / sx, qy = normy / sy;
    const float z = __fsub_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)),
                              __fmul_rn(2.f, __fdiv_rn(__fmul_rn(__fmul_rn(corr, normx), normy), sxsy)));
    const float neg_rho = __fsub_rn(1.f, __fmul_rn(corr, corr));
    float result = expf(__fdiv_rn(-z, __fmul_rn(2.f, neg_rho)));
    const float denom = __fmul_rn(6.283185307179586f, __fmul_rn(sxsy, sqrtf(neg_rho)));
    result = __fdiv_rn(result, denom);
    sum += -logf(fmaxf(result, 1e-20f));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) loss[warp] = sum / static_cast<float>(tot);
}

}  // namespace

int launch_nll_forward(const float* Vpred, const float* Vtr, const int32_t* seg_off, int S, float* loss,
                       cudaStream_t st) {
  if (S == 0) return SSTG_OK;
  const int warps_per_block = 8;
  const int grid = (S + warps_per_block - 1) / warps_per_block;
  nll_forward_kernel<<<grid, 256, 0, st>>>(Vpred, Vtr, seg_off, S, loss);
  SSTG_CUDA_TRY(cudaGetLastError(), "nll launch");
  return SSTG_OK;
}

int launch_sample_ade_fde(const float* Vpred, const float* Vtr, const float* last_obs, const int32_t* seg_off,
                          int S, int P, int K, const float* eps, uint64_t seed, float* eps_out, double* min_ade,
                          double* min_fde, float* pred_abs, int flags, cudaStream_t st) {
  if (S == 0 || P == 0) return SSTG_OK;
  const int grid = (P + E_PEDS - 1) / E_PEDS;
  const bool exact = flags & SSTG_EVAL_FP64_DIST, philox = eps == nullptr;
  const bool extra = eps_out != nullptr || pred_abs != nullptr;
  const int rel_in = (flags & SSTG_EVAL_VTR_IS_REL) ? 1 : 0;
  const int uni = (flags & SSTG_EVAL_UNIFORM_N) ? P / S : 0;  // caller guarantees N_s == P / S for every scene
#define SSTG_LAUNCH_SAMPLE(EX, PH, XT)                                                                         \
  sample_ade_fde_kernel<EX, PH, XT><<<grid, E_THREADS, 0, st>>>(Vpred, Vtr, last_obs, seg_off, S, K, eps, seed, \
                                                                eps_out, min_ade, min_fde, pred_abs, rel_in, uni)
  if (exact) {
    if (philox) { if (extra) SSTG_LAUNCH_SAMPLE(true, true, true); else SSTG_LAUNCH_SAMPLE(true, true, false); }
    else        { if (extra) SSTG_LAUNCH_SAMPLE(true, false, true); else SSTG_LAUNCH_SAMPLE(true, false, false); }
  } else {
    if (philox) { if (extra) SSTG_LAUNCH_SAMPLE(false, true, true); else SSTG_LAUNCH_SAMPLE(false, true, false); }
    else        { if (extra) SSTG_LAUNCH_SAMPLE(false, false, true); else SSTG_LAUNCH_SAMPLE(false, false, false); }
  }
#undef SSTG_LAUNCH_SAMPLE
  SSTG_CUDA_TRY(cudaGetLastError(), "sample_ade_fde launch");
  return SSTG_OK;
}

}  // namespace sstg

// ---- stand-alone trajectory helpers behind the reference's metrics.py API ---------------
namespace sstg {
namespace {

// nodes_rel_to_nodes_abs (metrics.py:69-75): abs[s,p] = fp32( fp64(sequential fp32 prefix sum) + init[p] )
// with init in fp64 (the reference's init node is a float64 numpy array).  One thread per pedestrian.
__global__ void rel_to_abs_kernel(const float* __restrict__ rel, const double* __restrict__ init,
                                  const int32_t* __restrict__ seg_off, int S, int T, float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = seg_off[S];
  if (p >= P) return;
  const int s = find_scene(seg_off, S, p);
  const int n0 = seg_off[s], N = seg_off[s + 1] - n0, nl = p - n0;
  const float2* r = reinterpret_cast<const float2*>(rel) + static_cast<size_t>(n0) * T + nl;
  float2* o = reinterpret_cast<float2*>(out) + static_cast<size_t>(n0) * T + nl;
  const double ix = init[2 * static_cast<size_t>(p)], iy = init[2 * static_cast<size_t>(p) + 1];
  float cx = 0.f, cy = 0.f;
  for (int t = 0; t < T; ++t) {
    const float2 v = r[static_cast<size_t>(t) * N];
    cx = __fadd_rn(cx, v.x);
    cy = __fadd_rn(cy, v.y);
    o[static_cast<size_t>(t) * N] = make_float2(static_cast<float>(static_cast<double>(cx) + ix),
                                                static_cast<float>(static_cast<double>(cy) + iy));
  }
}

// metrics.ade / metrics.fde (metrics.py:21-53) per pedestrian: sum over t of the fp64 sqrt of the
// fp32 squared distance (sequential), and the distance at the last step.
__global__ void traj_error_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                  const int32_t* __restrict__ seg_off, int S, int T,
                                  double* __restrict__ sum_dist, double* __restrict__ final_dist) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = seg_off[S];
  if (p >= P) return;
  const int s = find_scene(seg_off, S, p);
  const int n0 = seg_off[s], N = seg_off[s + 1] - n0, nl = p - n0;
  const float2* a = reinterpret_cast<const float2*>(pred) + static_cast<size_t>(n0) * T + nl;
  const float2* b = reinterpret_cast<const float2*>(tgt) + static_cast<size_t>(n0) * T + nl;
  double acc = 0.0, d = 0.0;
  for (int t = 0; t < T; ++t) {
    const float2 u = a[static_cast<size_t>(t) * N], v = b[static_cast<size_t>(t) * N];
    const float dx = __fsub_rn(u.x, v.x), dy = __fsub_rn(u.y, v.y);
    d = sqrt(static_cast<double>(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
    acc += d;
  }
  sum_dist[p] = acc;
  final_dist[p] = d;
}

}  // namespace

namespace {
// Compact host format -> the fp32 tensors of TrajectoryDataset, bit for bit.
// The dataset's positions are decimal numbers with 4 places (utils.py:145 np.around(.., 4) =
// rint(x * 1e4) / 1e4 in float64), so a track is exactly: an int32 start (units of 1e-4) plus 19
// int16 steps.  Decoding repeats the reference's arithmetic: pos64 = k / 1e4 (float64 division),
// rel = fp32(pos64[t] - pos64[t-1]) (utils.py:153-155,180-183), last_obs = fp32(pos64[7]).
// One thread per (pedestrian, coordinate); 20 frames = obs 8 + pred 12.
__global__ void decode_tracks_kernel(const int32_t* __restrict__ start, const int16_t* __restrict__ delta, int P,
                                     float* __restrict__ obs_rel, float* __restrict__ pred_rel,
                                     float* __restrict__ last_obs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // p*2 + c
  if (i >= 2 * P) return;
  const int2* dp = reinterpret_cast<const int2*>(delta + static_cast<size_t>(i) * 20);  // 40 bytes, 8-byte aligned
  int16_t d[20];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const int2 v = dp[q];
    d[4 * q] = static_cast<int16_t>(v.x & 0xffff);
    d[4 * q + 1] = static_cast<int16_t>(v.x >> 16);
    d[4 * q + 2] = static_cast<int16_t>(v.y & 0xffff);
    d[4 * q + 3] = static_cast<int16_t>(v.y >> 16);
  }
  int k = start[i];
  double prev = static_cast<double>(k) / 10000.0;
  float r[20];
  r[0] = 0.f;  // utils.py:153: the first relative displacement is zero
  float last = static_cast<float>(prev);
#pragma unroll
  for (int t = 1; t < 20; ++t) {
    k += d[t];
    const double cur = static_cast<double>(k) / 10000.0;
    r[t] = static_cast<float>(cur - prev);
    prev = cur;
    if (t == T_OBS - 1) last = static_cast<float>(cur);
  }
  float4* o = reinterpret_cast<float4*>(obs_rel + static_cast<size_t>(i) * T_OBS);
  o[0] = make_float4(r[0], r[1], r[2], r[3]);
  o[1] = make_float4(r[4], r[5], r[6], r[7]);
  float4* q = reinterpret_cast<float4*>(pred_rel + static_cast<size_t>(i) * T_PRED);
  q[0] = make_float4(r[8], r[9], r[10], r[11]);
  q[1] = make_float4(r[12], r[13], r[14], r[15]);
  q[2] = make_float4(r[16], r[17], r[18], r[19]);
  last_obs[i] = last;
}
}  // namespace

int launch_decode_tracks(const int32_t* start, const int16_t* delta, int P, float* obs_rel, float* pred_rel,
                         float* last_obs, cudaStream_t st) {
  if (P == 0) return SSTG_OK;
  decode_tracks_kernel<<<(2 * P + 255) / 256, 256, 0, st>>>(start, delta, P, obs_rel, pred_rel, last_obs);
  SSTG_CUDA_TRY(cudaGetLastError(), "decode_tracks launch");
  return SSTG_OK;
}

int launch_rel_to_abs(const float* rel, const double* init, const int32_t* seg_off, int S, int P, int T, float* out,
                      cudaStream_t st) {
  if (S == 0 || P == 0) return SSTG_OK;
  rel_to_abs_kernel<<<(P + 127) / 128, 128, 0, st>>>(rel, init, seg_off, S, T, out);
  SSTG_CUDA_TRY(cudaGetLastError(), "rel_to_abs launch");
  return SSTG_OK;
}

int launch_traj_error(const float* pred, const float* tgt, const int32_t* seg_off, int S, int P, int T,
                      double* sum_dist, double* final_dist, cudaStream_t st) {
  if (S == 0 || P == 0) return SSTG_OK;
  traj_error_kernel<<<(P + 127) / 128, 128, 0, st>>>(pred, tgt, seg_off, S, T, sum_dist, final_dist);
  SSTG_CUDA_TRY(cudaGetLastError(), "traj_error launch");
  return SSTG_OK;
}

}  // namespace sstg
