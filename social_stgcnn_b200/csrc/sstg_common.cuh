// Shared device/host helpers for the sm_100a kernels of libsstg.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sstg.h"

namespace sstg {

// ---- parameter block offsets: named_parameters() order of reference model.py:157-178
constexpr int P_GCN_W = 0;     // st_gcns.0.gcn.conv.weight (5,2,1,1)
constexpr int P_GCN_B = 10;    // .bias (5)
constexpr int P_BN1_G = 15;    // tcn.0.weight
constexpr int P_BN1_B = 20;    // tcn.0.bias
constexpr int P_PRELU1 = 25;   // tcn.1.weight
constexpr int P_TCN_W = 26;    // tcn.2.weight (5,5,3,1)
constexpr int P_TCN_B = 101;   // tcn.2.bias
constexpr int P_BN2_G = 106;   // tcn.3.weight
constexpr int P_BN2_B = 111;   // tcn.3.bias
constexpr int P_RES_W = 116;   // residual.0.weight (5,2,1,1)
constexpr int P_RES_B = 126;   // residual.0.bias
constexpr int P_BNR_G = 131;   // residual.1.weight
constexpr int P_BNR_B = 136;   // residual.1.bias
constexpr int P_PRELU2 = 141;  // prelu.weight
constexpr int P_TP0_W = 142;   // tpcnns.0.weight (12,8,3,3)
constexpr int P_TP0_B = 1006;
constexpr int P_TP1_W = 1018;  // tpcnns.k.weight (12,12,3,3), stride 1308 per layer
constexpr int P_TP_STRIDE = 1308;
constexpr int P_TPOUT_W = 6250;  // tpcnn_ouput.weight
constexpr int P_TPOUT_B = 7546;
constexpr int P_PRELUS = 7558;   // prelus.0..4
// bn block: rm1[5] rv1[5] rm2[5] rv2[5] rmr[5] rvr[5]
constexpr float BN_EPS = 1e-5f;

constexpr int T_OBS = SSTG_T_OBS;
constexpr int T_PRED = SSTG_T_PRED;
constexpr int C = SSTG_C;

// thread-local last CUDA error text (sstg_api.cu)
void note_cuda_error(cudaError_t e, const char* where);

#define SSTG_CUDA_TRY(expr, where)                 \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) {                       \
      ::sstg::note_cuda_error(_e, where);          \
      return SSTG_ECUDA;                           \
    }                                              \
  } while (0)

#ifdef __CUDACC__
// ---- mbarrier + bulk async copy (TMA, 1-D) wrappers ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same with an L2 evict-first policy: for data that is read exactly once
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// shared -> global bulk copy (bulk-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- programmatic dependent launch (the training step is a chain of short kernels: each one's prologue --
// weight staging out of `params`, table set-up -- runs underneath its predecessor's main body).
// Rules kept by every kernel that uses them: nothing produced by an earlier kernel is read before pdl_wait();
// dependents are released only AFTER this kernel's own wait (so a prologue never overlaps the kernel two
// places back); kernels that WRITE `params` or the plan tables never release early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float prelu(float x, float a) { return x >= 0.f ? x : a * x; }
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// inverse distance from the squared distance; a squared distance below FLT_MIN (only the
// exact-zero case occurs with real coordinates) is the reference's "distance 0 -> weight 0"
__device__ __forceinline__ float inv_dist(float sq) { return sq >= 1.17549435e-38f ? rsqrt_fast(sq) : 0.f; }

#endif  // __CUDACC__

int sstg_xtile_pairs();  // sstg_api.cu: column pairs per TXP-CNN tile, 64 or (SSTG_XTILE_PAIRS=32) 32 -- read by the planner AND the launcher
bool pdl_enabled();  // sstg_api.cu: false when the environment has SSTG_NO_PDL=1 (read at every launch)

#ifdef __CUDACC__
// <<<grid, block, smem, st>>> with programmatic stream serialization allowed (the kernel must call pdl_wait()).
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;   // SSTG_NO_PDL=1: plain stream order (A/B runs, tests)
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// launchers (one per .cu file)
int launch_graph_build(const float* rel, const int32_t* seg_off, const int64_t* a_off, int S, int T, int max_n,
                       float* V, float* A, int flags, cudaStream_t st);
int launch_stgcn_forward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off, int S,
                         int max_n, const float* params, const float* bn, float* H, int sm_count,
                         cudaStream_t st);
int launch_txpcnn_forward(const float* H, const int32_t* seg_off, const int32_t* tile_off, int n_tiles, int max_n,
                          const float* params, float* out, float* Usave, float* Xsave, long long layer_stride,
                          int tile_cols, cudaStream_t st, const float* bn_stats = nullptr, int bn_scenes = 0,
                          double bn_momentum = 0.0, float* bn = nullptr);
int launch_rel_to_abs(const float* rel, const double* init, const int32_t* seg_off, int S, int P, int T, float* out,
                      cudaStream_t st);
int launch_traj_error(const float* pred, const float* tgt, const int32_t* seg_off, int S, int P, int T,
                      double* sum_dist, double* final_dist, cudaStream_t st);
int launch_stgcn_forward_fused(const float* rel, const int32_t* seg_off, int S, int max_n, const float* params,
                               const float* bn, float* Pws, float* H, cudaStream_t st);
int launch_decode_tracks(const int32_t* start, const int16_t* delta, int P, float* obs_rel, float* pred_rel,
                         float* last_obs, cudaStream_t st);
int txpcnn_tile_width(int max_n);
int txpcnn_wino_mode();
int launch_txpcnn_forward_wino(const float* H, const int32_t* seg_off, const int32_t* tile_off, int n_tiles,
                               const float* params, float* out, int mode, int tile_cols_hint, cudaStream_t st);
int txpcnn_bwd_rows(int n_tiles);
int launch_txpcnn_backward(const float* dOut, const float* Usave, const float* Xsave, long long layer_stride,
                           const float* H, const int32_t* seg_off, const int32_t* tile_off, int n_tiles, int max_n,
                           const float* params, float* dH, float* partial, cudaStream_t st);
int stgcn_train_bwd_rows(int S);
int launch_stgcn_train_forward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off, int S,
                               int max_n, const float* params, float* Y, float* Z, float* H, float* stats,
                               cudaStream_t st);
int launch_stgcn_train_backward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off, int S,
                                int max_n, const float* params, const float* Y, const float* Z,
                                const float* stats, const float* dH, float* partial, cudaStream_t st);
int launch_nll_fwd_bwd(const float* Vpred, const float* Vtr, const int32_t* seg_off, int S, const float* weight,
                       float* loss, float* dV, cudaStream_t st, int vtr_is_rel = 0);
int launch_bn_running_update(const float* stats, const int32_t* seg_off, int S, double momentum, float* bn,
                             cudaStream_t st, const int32_t* abort_flag = nullptr);
int preload_bn_running_update();
int launch_grad_finalize(const float* part_st, int rows_st, const float* part_tx, int rows_tx, int tx_stride,
                         float* grad, cudaStream_t st);
int launch_group_loss(const float* loss, const float* weight, int S, float* out, cudaStream_t st);
int launch_train_tail(const float* part_st, int rows_st, const float* part_tx, int rows_tx, int tx_stride,
                      float* grad, float* params, float lr, const float* stats, const int32_t* seg_off, int S,
                      double momentum, float* bn, const float* loss, const float* weight, float* group_loss,
                      cudaStream_t st);
int launch_sgd_step(float* params, const float* grad, int n, float lr, float clip_norm, float* grad_norm_out,
                    cudaStream_t st);
int launch_nll_forward(const float* Vpred, const float* Vtr, const int32_t* seg_off, int S, float* loss,
                       cudaStream_t st);
int launch_sample_ade_fde(const float* Vpred, const float* Vtr, const float* last_obs, const int32_t* seg_off,
                          int S, int P, int K, const float* eps, uint64_t seed, float* eps_out, double* min_ade,
                          double* min_fde, float* pred_abs, int flags, cudaStream_t st);

}  // namespace sstg
