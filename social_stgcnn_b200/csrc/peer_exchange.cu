// Gradient exchange of the scene-sharded training group over NVLink peer memory, fused with the
// optimiser step (the reference trains on one device, train.py:28-79, so the group is the
// unit that shards).
//
// Every rank owns a MAILBOX in its own HBM that the other ranks of the node map through CUDA IPC:
//
//   per parity p in {0,1}:   slot[src] (7,564 floats: flat gradient | group loss), src = 0..world-1
//                            stats[max_scenes x 30]  per-scene BatchNorm statistics in GLOBAL scene order
//   flags[world] (uint32)    flags[src] = last epoch whose data `src` has fully delivered here
//
// One kernel per rank and group (`peer_exchange_sgd_kernel`, one CTA):
//   push   this rank's gradient, loss and its own scenes' statistics rows are STORED into the mailbox of
//          every rank (its own included) -- NVLink stores are fire-and-forget, no remote load is ever issued;
//   signal __threadfence_system, then a release store of the epoch into flags[rank] of every mailbox;
//   wait   acquire-poll the LOCAL flags until every source has delivered this epoch (bounded by a
//          timeout: on expiry *status is set and nothing is applied -- the GPU never hangs on a dead peer;
//          the status word is sticky and a rank whose status is set takes no part in later exchanges, so the
//          failure reaches every rank within one more timeout);
//   apply  sum the slots in rank order (fixed order => every replica gets the same bits), optional
//          global-norm clipping, SGD update of the resident weights, statistics copied out for the
//          running-statistics recurrence (bn_running_update_kernel, which every rank runs over the
//          global scene order).
// Parity = epoch & 1 double-buffers the payload: a rank can only start pushing epoch e+2 after it has
// seen every flag of epoch e+1, which every peer sets only after its own epoch-e kernel has finished
// reading -- so a slot is never overwritten while it is being read.
#include <cstring>

#include "sstg_common.cuh"

namespace sstg {
namespace {

constexpr int X_THREADS = 1024;
constexpr int SLOT_FLOATS = 7564;  // gradient + group loss; a multiple of 4
static_assert(SLOT_FLOATS >= SSTG_N_PARAMS + 1 && SLOT_FLOATS % 4 == 0, "slot layout");

struct PeerBoxes {
  float* box[SSTG_MAX_RANKS];
};

__host__ __device__ inline int64_t parity_floats(int world, int max_scenes) {
  return int64_t(world) * SLOT_FLOATS + ((int64_t(max_scenes) * SSTG_N_BN + 3) & ~int64_t(3));
}
__host__ __device__ inline int64_t flags_offset_floats(int world, int max_scenes) {
  return 2 * parity_floats(world, max_scenes);
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float warp_sum_x(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One rank's arguments.  The production call launches ONE CTA with one of these; a process that holds several
// ranks of a group on ONE device (the single-GPU test harness) launches them as the CTAs of ONE cooperative
// grid -- CTAs of a cooperative launch are co-resident by contract, separate launches that spin on each other
// are not (B200_PROFILING.md: kernels that wait on one another must not be separate launches on one GPU).
struct RankArgs {
  const float* grad_local;
  const float* loss_local;
  const float* stats_local;
  float* params;
  float* grad_out;
  float* loss_out;
  float* stats_out;
  int32_t* status;
  int rank, scene_lo, s_local;
};
struct RankSet {
  RankArgs r[SSTG_MAX_RANKS];
};

__global__ void __launch_bounds__(X_THREADS)
peer_exchange_sgd_kernel(PeerBoxes boxes, RankSet set, int world, uint32_t epoch, int max_scenes, int s_all, float lr,
                         float clip_norm, unsigned long long timeout_ns) {
  __shared__ float red[X_THREADS / 32];
  __shared__ float coef_s;
  __shared__ int late_s;
  const RankArgs& a = set.r[blockIdx.x];
  const int rank = a.rank;
  const float* __restrict__ grad_local = a.grad_local;
  const float* __restrict__ stats_local = a.stats_local;
  float* __restrict__ params = a.params;
  float* __restrict__ grad_out = a.grad_out;
  float* __restrict__ stats_out = a.stats_out;
  const int tid = threadIdx.x;
  // A rank whose earlier exchange failed takes no further part (the status word is sticky): it neither pushes
  // nor applies, so its peers run into their own time-out and every replica stops at the next check instead of
  // training on with weights that may differ.
  if (*static_cast<volatile int32_t*>(a.status) != 0) return;
  const int64_t pbase = int64_t(epoch & 1u) * parity_floats(world, max_scenes);
  const int64_t stats_base = pbase + int64_t(world) * SLOT_FLOATS;
  const int64_t flags_base = flags_offset_floats(world, max_scenes);
  if (tid == 0) late_s = 0;

  // ---- push: one pass over the local payload, stored to every mailbox (rotated start spreads the links)
  const float gl = a.loss_local ? *a.loss_local : 0.f;
  for (int i = tid; i < SLOT_FLOATS; i += X_THREADS) {
    const float v = i < SSTG_N_PARAMS ? grad_local[i] : (i == SSTG_N_PARAMS ? gl : 0.f);
    for (int d = 0; d < world; ++d) {
      int dst = rank + 1 + d;
      if (dst >= world) dst -= world;
      boxes.box[dst][pbase + int64_t(rank) * SLOT_FLOATS + i] = v;
    }
  }
  for (int i = tid; i < a.s_local * SSTG_N_BN; i += X_THREADS) {
    const float v = stats_local[i];
    for (int d = 0; d < world; ++d) {
      int dst = rank + 1 + d;
      if (dst >= world) dst -= world;
      boxes.box[dst][stats_base + int64_t(a.scene_lo) * SSTG_N_BN + i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  // ---- signal, then wait for every source on the local flags
  if (tid < world)
    st_release_sys(reinterpret_cast<uint32_t*>(boxes.box[tid] + flags_base) + rank, epoch);
  if (tid < world) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(boxes.box[rank] + flags_base) + tid;
    const uint64_t t0 = global_ns();
    // epochs only grow; the signed difference tolerates the 32-bit wrap
    while (static_cast<int32_t>(ld_acquire_sys(f) - epoch) < 0) {
      if (global_ns() - t0 > timeout_ns) {
        atomicExch(&late_s, 1 + tid);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (late_s) {
    if (tid == 0) *a.status = late_s;  // 1 + the first source that never arrived
    return;
  }

  // ---- apply: sum in rank order, clip, SGD
  const float* mine = boxes.box[rank] + pbase;
  float ss = 0.f;
  for (int i = tid; i < SLOT_FLOATS; i += X_THREADS) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += __ldcg(mine + int64_t(r) * SLOT_FLOATS + i);
    if (i < SSTG_N_PARAMS) {
      grad_out[i] = s;
      ss = fmaf(s, s, ss);
    } else if (i == SSTG_N_PARAMS && a.loss_out) {
      *a.loss_out = s;
    }
  }
  const float* st_in = boxes.box[rank] + stats_base;
  for (int i = tid; i < s_all * SSTG_N_BN; i += X_THREADS) stats_out[i] = __ldcg(st_in + i);
  ss = warp_sum_x(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < X_THREADS / 32; ++w) tot += red[w];
    float coef = 1.f;
    if (clip_norm > 0.f) {
      coef = clip_norm / (sqrtf(tot) + 1e-6f);
      if (coef > 1.f) coef = 1.f;
    }
    coef_s = coef;
  }
  __syncthreads();
  if (lr != 0.f) {
    const float step = lr * coef_s;
    // each thread re-reads exactly the grad_out elements it wrote above
    for (int i = tid; i < SSTG_N_PARAMS; i += X_THREADS) params[i] = fmaf(-step, grad_out[i], params[i]);
  }
}

// Argument checks shared by the two entry points; fills one RankArgs.
int fill_rank(const sstg_peer_rank& in, int world, int max_scenes, int s_all, RankArgs* out) {
  if (in.rank < 0 || in.rank >= world || in.s_local < 0 || in.scene_lo < 0 ||
      int64_t(in.scene_lo) + in.s_local > s_all)
    return SSTG_EINVAL;
  if (!in.grad_local || !in.params || !in.grad_out || !in.status || (in.s_local > 0 && !in.stats_local))
    return SSTG_EINVAL;
  if (s_all > 0 && (!in.stats_out || !in.seg_off_all || !in.bn)) return SSTG_EINVAL;
  (void)max_scenes;
  out->grad_local = in.grad_local;
  out->loss_local = in.loss_local;
  out->stats_local = in.stats_local;
  out->params = in.params;
  out->grad_out = in.grad_out;
  out->loss_out = in.loss_out;
  out->stats_out = in.stats_out;
  out->status = in.status;
  out->rank = in.rank;
  out->scene_lo = in.scene_lo;
  out->s_local = in.s_local;
  return SSTG_OK;
}

int exchange_launch(void* const* mailboxes, int world, const sstg_peer_rank* ranks, int n_local, uint32_t epoch,
                    int max_scenes, int s_all, float lr, float clip_norm, double bn_momentum, uint64_t timeout_ns,
                    cudaStream_t st) {
  if (world < 1 || world > SSTG_MAX_RANKS || epoch == 0 || !ranks || n_local < 1 || n_local > world)
    return SSTG_EINVAL;
  if (max_scenes < 0 || s_all < 0 || s_all > max_scenes || !mailboxes) return SSTG_EINVAL;
  if (!(bn_momentum >= 0.0 && bn_momentum <= 1.0)) return SSTG_EINVAL;
  RankSet set;
  std::memset(&set, 0, sizeof(set));
  unsigned seen = 0;
  for (int i = 0; i < n_local; ++i) {
    const int rc = fill_rank(ranks[i], world, max_scenes, s_all, &set.r[i]);
    if (rc != SSTG_OK) return rc;
    if (seen & (1u << ranks[i].rank)) return SSTG_EINVAL;  // a rank listed twice
    seen |= 1u << ranks[i].rank;
  }
  // All kernels of this call are loaded BEFORE the first one is launched: CUDA's lazy loading may synchronise
  // the context on a first launch, and with the exchange kernel already waiting for peers that would stall this
  // rank.
  {
    cudaFuncAttributes fa;
    SSTG_CUDA_TRY(cudaFuncGetAttributes(&fa, peer_exchange_sgd_kernel), "peer_exchange preload");
    const int rc = preload_bn_running_update();
    if (rc != SSTG_OK) return rc;
  }
  PeerBoxes boxes;
  for (int r = 0; r < SSTG_MAX_RANKS; ++r) {
    boxes.box[r] = r < world ? static_cast<float*>(mailboxes[r]) : nullptr;
    if (r < world && !boxes.box[r]) return SSTG_EINVAL;
  }
  unsigned long long tmo = timeout_ns ? timeout_ns : 30000000000ull;
  if (n_local == 1) {
    peer_exchange_sgd_kernel<<<1, X_THREADS, 0, st>>>(boxes, set, world, epoch, max_scenes, s_all, lr, clip_norm, tmo);
    SSTG_CUDA_TRY(cudaGetLastError(), "peer_exchange_sgd launch");
  } else {
    // several ranks of the group live on THIS device: their CTAs wait on one another, so they must be
    // co-resident -- one cooperative launch (fails cleanly instead of dead-locking if they would not fit)
    void* args[] = {&boxes, &set, &world, &epoch, &max_scenes, &s_all, &lr, &clip_norm, &tmo};
    SSTG_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(peer_exchange_sgd_kernel), dim3(n_local),
                                              dim3(X_THREADS), args, 0, st),
                  "peer_exchange_sgd cooperative launch");
  }
  for (int i = 0; i < n_local; ++i) {
    const int rc = launch_bn_running_update(ranks[i].stats_out, ranks[i].seg_off_all, s_all, bn_momentum, ranks[i].bn,
                                            st, ranks[i].status);
    if (rc != SSTG_OK) return rc;
  }
  return SSTG_OK;
}

}  // namespace
}  // namespace sstg

using namespace sstg;

extern "C" {

int64_t sstg_peer_mailbox_bytes(int32_t world, int32_t max_scenes) {
  if (world < 1 || world > SSTG_MAX_RANKS || max_scenes < 0) return SSTG_EINVAL;
  return (flags_offset_floats(world, max_scenes) + 2 * SSTG_MAX_RANKS) * int64_t(sizeof(float));
}

int sstg_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle) {
  if (bytes <= 0 || !ptr || !handle) return SSTG_EINVAL;
  void* p = nullptr;
  SSTG_CUDA_TRY(cudaMalloc(&p, static_cast<size_t>(bytes)), "peer mailbox cudaMalloc");
  cudaError_t e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    note_cuda_error(e, "peer mailbox export");
    cudaFree(p);
    return SSTG_ECUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == SSTG_PEER_HANDLE_BYTES, "handle size");
  std::memcpy(handle, &h, sizeof(h));
  *ptr = p;
  return SSTG_OK;
}

int sstg_peer_open(const unsigned char* handle, void** ptr) {
  if (!handle || !ptr) return SSTG_EINVAL;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  SSTG_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "peer mailbox cudaIpcOpenMemHandle");
  *ptr = p;
  return SSTG_OK;
}

int sstg_peer_close(void* ptr) {
  if (!ptr) return SSTG_OK;
  SSTG_CUDA_TRY(cudaIpcCloseMemHandle(ptr), "peer mailbox cudaIpcCloseMemHandle");
  return SSTG_OK;
}

int sstg_peer_free(void* ptr) {
  if (!ptr) return SSTG_OK;
  SSTG_CUDA_TRY(cudaFree(ptr), "peer mailbox cudaFree");
  return SSTG_OK;
}

int sstg_peer_exchange_apply(void* const* mailboxes, int32_t world, int32_t rank, uint32_t epoch,
                             int32_t max_scenes, const float* grad_local, const float* loss_local,
                             const float* stats_local, int32_t scene_lo, int32_t s_local,
                             const int32_t* seg_off_all, int32_t s_all, float* params, float* bn, float lr,
                             float clip_norm, double bn_momentum, float* grad_out, float* loss_out,
                             float* stats_out, int32_t* status, uint64_t timeout_ns, void* stream) {
  sstg_peer_rank one;
  one.rank = rank;
  one.scene_lo = scene_lo;
  one.s_local = s_local;
  one.reserved = 0;
  one.grad_local = grad_local;
  one.loss_local = loss_local;
  one.stats_local = stats_local;
  one.seg_off_all = seg_off_all;
  one.params = params;
  one.bn = bn;
  one.grad_out = grad_out;
  one.loss_out = loss_out;
  one.stats_out = stats_out;
  one.status = status;
  return exchange_launch(mailboxes, world, &one, 1, epoch, max_scenes, s_all, lr, clip_norm, bn_momentum, timeout_ns,
                         static_cast<cudaStream_t>(stream));
}

int sstg_peer_exchange_apply_local(void* const* mailboxes, int32_t world, const sstg_peer_rank* ranks,
                                   int32_t n_local, uint32_t epoch, int32_t max_scenes, int32_t s_all, float lr,
                                   float clip_norm, double bn_momentum, uint64_t timeout_ns, void* stream) {
  return exchange_launch(mailboxes, world, ranks, n_local, epoch, max_scenes, s_all, lr, clip_norm, bn_momentum,
                         timeout_ns, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
