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
//          timeout: on expiry *status is set and nothing is applied -- the GPU never hangs on a dead peer);
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

__global__ void __launch_bounds__(X_THREADS)
peer_exchange_sgd_kernel(PeerBoxes boxes, int world, int rank, uint32_t epoch, int max_scenes,
                         const float* __restrict__ grad_local, const float* __restrict__ loss_local,
                         const float* __restrict__ stats_local, int scene_lo, int s_local, int s_all,
                         float* __restrict__ params, float lr, float clip_norm, float* __restrict__ grad_out,
                         float* __restrict__ loss_out, float* __restrict__ stats_out,
                         int32_t* __restrict__ status, unsigned long long timeout_ns) {
  __shared__ float red[X_THREADS / 32];
  __shared__ float coef_s;
  __shared__ int late_s;
  const int tid = threadIdx.x;
  const int64_t pbase = int64_t(epoch & 1u) * parity_floats(world, max_scenes);
  const int64_t stats_base = pbase + int64_t(world) * SLOT_FLOATS;
  const int64_t flags_base = flags_offset_floats(world, max_scenes);
  if (tid == 0) late_s = 0;

  // ---- push: one pass over the local payload, stored to every mailbox (rotated start spreads the links)
  const float gl = loss_local ? *loss_local : 0.f;
  for (int i = tid; i < SLOT_FLOATS; i += X_THREADS) {
    const float v = i < SSTG_N_PARAMS ? grad_local[i] : (i == SSTG_N_PARAMS ? gl : 0.f);
    for (int d = 0; d < world; ++d) {
      int dst = rank + 1 + d;
      if (dst >= world) dst -= world;
      boxes.box[dst][pbase + int64_t(rank) * SLOT_FLOATS + i] = v;
    }
  }
  for (int i = tid; i < s_local * SSTG_N_BN; i += X_THREADS) {
    const float v = stats_local[i];
    for (int d = 0; d < world; ++d) {
      int dst = rank + 1 + d;
      if (dst >= world) dst -= world;
      boxes.box[dst][stats_base + int64_t(scene_lo) * SSTG_N_BN + i] = v;
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
    if (tid == 0) *status = late_s;  // 1 + the first source that never arrived
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
    } else if (i == SSTG_N_PARAMS && loss_out) {
      *loss_out = s;
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
  if (world < 1 || world > SSTG_MAX_RANKS || rank < 0 || rank >= world || epoch == 0) return SSTG_EINVAL;
  if (max_scenes < 0 || s_all < 0 || s_all > max_scenes || s_local < 0 || scene_lo < 0 ||
      int64_t(scene_lo) + s_local > s_all)
    return SSTG_EINVAL;
  if (!mailboxes || !grad_local || !params || !grad_out || !status || (s_local > 0 && !stats_local)) return SSTG_EINVAL;
  if (s_all > 0 && (!stats_out || !seg_off_all || !bn)) return SSTG_EINVAL;
  if (!(bn_momentum >= 0.0 && bn_momentum <= 1.0)) return SSTG_EINVAL;
  // Both kernels of this call are loaded BEFORE the first one is launched: CUDA's lazy loading may
  // synchronise the context on a first launch, and with the exchange kernel already waiting for peers that
  // would stall this rank (and deadlock ranks that share one process, as the single-GPU tests do).
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  peer_exchange_sgd_kernel<<<1, X_THREADS, 0, st>>>(boxes, world, rank, epoch, max_scenes, grad_local, loss_local,
                                                    stats_local, scene_lo, s_local, s_all, params, lr, clip_norm,
                                                    grad_out, loss_out, stats_out, status,
                                                    timeout_ns ? timeout_ns : 5000000000ull);
  SSTG_CUDA_TRY(cudaGetLastError(), "peer_exchange_sgd launch");
  return launch_bn_running_update(stats_out, seg_off_all, s_all, bn_momentum, bn, st, status);
}

}  // extern "C"
