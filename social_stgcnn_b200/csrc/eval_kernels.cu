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
#include "sstg_common.cuh"

namespace sstg {

namespace {

constexpr int E_PEDS = 64;
constexpr int E_SUB = 4;
constexpr int E_THREADS = E_PEDS * E_SUB;

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

// two standard normals from two uint32 (Box-Muller on (0,1) uniforms)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (static_cast<float>(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(6.283185307179586f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
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

template <bool FP64_DIST>
__global__ void __launch_bounds__(E_THREADS)
sample_ade_fde_kernel(const float* __restrict__ Vpred, const float* __restrict__ Vtr,
                      const float* __restrict__ last_obs, const int32_t* __restrict__ seg_off, int S, int K,
                      const float* __restrict__ eps, uint64_t seed, float* __restrict__ eps_out,
                      double* __restrict__ min_ade, double* __restrict__ min_fde, float* __restrict__ pred_abs) {
  // [t][field][ped]: mux, muy, L00, L10, L11, gtx, gty
  __shared__ float par[T_PRED][7][E_PEDS];
  __shared__ double red[2][E_SUB][E_PEDS];
  __shared__ int sc_n0[E_PEDS], sc_N[E_PEDS];
  const int tid = threadIdx.x;
  const int P = seg_off[S];
  const int p_base = blockIdx.x * E_PEDS;

  if (tid < E_PEDS) {
    const int p = p_base + tid;
    if (p < P) {
      const int s = find_scene(seg_off, S, p);
      sc_n0[tid] = seg_off[s];
      sc_N[tid] = seg_off[s + 1] - seg_off[s];
    } else {
      sc_n0[tid] = 0;
      sc_N[tid] = 0;
    }
  }
  __syncthreads();
  // ---- distribution parameters per (t, pedestrian)
  for (int i = tid; i < T_PRED * E_PEDS; i += E_THREADS) {
    const int t = i / E_PEDS, pl = i - t * E_PEDS;
    const int N = sc_N[pl];
    if (N == 0) continue;
    const int n0 = sc_n0[pl], nl = p_base + pl - n0;
    const float* vp = Vpred + static_cast<size_t>(n0) * 60 + t * N + nl;  // (c, t, n) at (c*12 + t)*N + n
    const float mux = vp[0], muy = vp[12 * N], p2 = vp[24 * N], p3 = vp[36 * N], p4 = vp[48 * N];
    const float sx = expf(p2), sy = expf(p3), rho = tanhf(p4);
    const float a = __fmul_rn(sx, sx);
    const float b = __fmul_rn(__fmul_rn(rho, sx), sy);
    const float c = __fmul_rn(sy, sy);
    const float L00 = sqrtf(a);                // potrf
    const float L10 = __fdiv_rn(b, L00);
    const float L11 = sqrtf(fmaf(-L10, L10, c));
    par[t][0][pl] = mux;
    par[t][1][pl] = muy;
    par[t][2][pl] = L00;
    par[t][3][pl] = L10;
    par[t][4][pl] = L11;
  }
  // ---- ground-truth absolute trajectory: sequential fp32 prefix sum + fp64 add of the start
  if (tid < E_PEDS && sc_N[tid] > 0) {
    const int N = sc_N[tid], n0 = sc_n0[tid], nl = p_base + tid - n0;
    const float2* vt = reinterpret_cast<const float2*>(Vtr) + static_cast<size_t>(n0) * T_PRED + nl;
    const float2 init = reinterpret_cast<const float2*>(last_obs)[p_base + tid];
    float cx = 0.f, cy = 0.f;
#pragma unroll
    for (int t = 0; t < T_PRED; ++t) {
      const float2 v = vt[t * N];
      cx = __fadd_rn(cx, v.x);
      cy = __fadd_rn(cy, v.y);
      par[t][5][tid] = static_cast<float>(static_cast<double>(cx) + static_cast<double>(init.x));
      par[t][6][tid] = static_cast<float>(static_cast<double>(cy) + static_cast<double>(init.y));
    }
  }
  __syncthreads();

  const int sub = tid / E_PEDS, pl = tid - sub * E_PEDS;
  const int N = sc_N[pl];
  double best_a = 1e300, best_f = 1e300;
  if (N > 0) {
    const int p = p_base + pl;
    const int n0 = sc_n0[pl], nl = p - n0;
    const float2 init = reinterpret_cast<const float2*>(last_obs)[p];
    const double ix = init.x, iy = init.y;
    const size_t eb = static_cast<size_t>(n0) * K * (T_PRED * 2);  // scene block (K,12,N,2)
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    for (int k = sub; k < K; k += E_SUB) {
      float cx = 0.f, cy = 0.f;
      double acc = 0.0, fde = 0.0;
      float2 e_next = make_float2(0.f, 0.f);
#pragma unroll
      for (int t = 0; t < T_PRED; ++t) {
        const size_t ei = eb + (static_cast<size_t>(k * T_PRED + t) * N + nl) * 2;
        float2 e;
        if (eps != nullptr) {
          e = *reinterpret_cast<const float2*>(eps + ei);
        } else if ((t & 1) == 0) {
          const uint4 r = Philox::run(make_uint4(static_cast<uint32_t>(k * (T_PRED / 2) + t / 2), 0u,
                                                 static_cast<uint32_t>(p), 0u), key);
          e = box_muller(r.x, r.y);
          e_next = box_muller(r.z, r.w);
        } else {
          e = e_next;
        }
        if (eps_out != nullptr) *reinterpret_cast<float2*>(eps_out + ei) = e;
        // sample = mean + L @ eps, separate roundings like the bmm + add of rsample
        const float sxv = __fadd_rn(par[t][0][pl], __fmul_rn(par[t][2][pl], e.x));
        const float syv = __fadd_rn(par[t][1][pl],
                                    __fadd_rn(__fmul_rn(par[t][3][pl], e.x), __fmul_rn(par[t][4][pl], e.y)));
        cx = __fadd_rn(cx, sxv);
        cy = __fadd_rn(cy, syv);
        const float ax = static_cast<float>(static_cast<double>(cx) + ix);
        const float ay = static_cast<float>(static_cast<double>(cy) + iy);
        if (pred_abs != nullptr) *reinterpret_cast<float2*>(pred_abs + ei) = make_float2(ax, ay);
        const float dx = __fsub_rn(ax, par[t][5][pl]);
        const float dy = __fsub_rn(ay, par[t][6][pl]);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const double d = FP64_DIST ? sqrt(static_cast<double>(d2)) : static_cast<double>(sqrtf(d2));
        acc += d;
        if (t == T_PRED - 1) fde = d;
      }
      const double ade = acc / static_cast<double>(T_PRED);
      best_a = ade < best_a ? ade : best_a;
      best_f = fde < best_f ? fde : best_f;
    }
  }
  red[0][sub][pl] = best_a;
  red[1][sub][pl] = best_f;
  __syncthreads();
  if (tid < E_PEDS && sc_N[tid] > 0) {
    double a = red[0][0][tid], f = red[1][0][tid];
#pragma unroll
    for (int u = 1; u < E_SUB; ++u) {
      a = red[0][u][tid] < a ? red[0][u][tid] : a;
      f = red[1][u][tid] < f ? red[1][u][tid] : f;
    }
    min_ade[p_base + tid] = a;
    min_fde[p_base + tid] = f;
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
  if (flags & SSTG_EVAL_FP64_DIST)
    sample_ade_fde_kernel<true><<<grid, E_THREADS, 0, st>>>(Vpred, Vtr, last_obs, seg_off, S, K, eps, seed,
                                                            eps_out, min_ade, min_fde, pred_abs);
  else
    sample_ade_fde_kernel<false><<<grid, E_THREADS, 0, st>>>(Vpred, Vtr, last_obs, seg_off, S, K, eps, seed,
                                                             eps_out, min_ade, min_fde, pred_abs);
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
