/*
 * sstg.h -- C ABI of the B200-native Social-STGCNN hot path (libsstg.so).
 *
 * The reference (GRatTWCU/Social-STGCNN) is pure Python and has no FFI boundary; its
 * boundary is the Python API.  Each entry point below therefore cites the reference
 * FUNCTION it replaces (file:line under the reference tree), and
 * social_stgcnn_b200/{utils,model,metrics}.py put the reference's own signatures back on
 * top of these calls (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - all tensors are fp32, densely packed, 16-byte aligned at the base;
 *   - a batch is a list of S scenes with ragged pedestrian counts N_s, segment-packed:
 *       seg_off[S+1]  int32  prefix sum of N_s              (pedestrian index space, P = seg_off[S])
 *       a_off[S+1]    int64  float offset of scene s's (T,N_s,N_s) adjacency block; every
 *                            entry is a multiple of 4 floats (16 B) -- see sstg_plan_host
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - nothing here allocates or synchronises: the caller owns every buffer, the calls
 *     are stream-ordered and thread-safe (no global state);
 *   - return value: 0 on success, a negative SSTG_E* code otherwise (sstg_error_string).
 *
 * Model dimensions are those of every shipped checkpoint (train.py:127-136):
 * input_feat 2, output_feat 5, obs_seq_len 8, pred_seq_len 12, n_stgcnn 1, n_txpcnn 5, kernel 3.
 */
#ifndef SSTG_H_
#define SSTG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSTG_T_OBS 8
#define SSTG_T_PRED 12
#define SSTG_C_IN 2
#define SSTG_C 5
#define SSTG_N_PARAMS 7563 /* named_parameters() order of model.py:157-178 */
#define SSTG_N_BN 30       /* running_mean[5], running_var[5] of tcn.0, tcn.3, residual.1 */
#define SSTG_MAX_PEDS 256  /* largest scene the kernels accept (shared-memory bound) */
#define SSTG_XTILE_COLS 128 /* TXP-CNN tile: whole scenes up to this many pedestrian columns */
#define SSTG_XTILE_MAX_SCENES 32
#define SSTG_XTILE_MAX_PAIRS 64  /* a tile holds at most 64 column pairs: sum over its scenes of ceil(N_s / 2) */
#define SSTG_ST_PARAMS 142          /* st_gcn parameters: the first 142 floats of the block */
#define SSTG_TXP_GRAD_STRIDE 7424   /* row length of the TXP-CNN partial-gradient buffer (7421 padded) */
#define SSTG_SAVE_FLOATS_PER_PED 600 /* activations kept for backward: Y 40, Z 40, H 40, U 4x60, X 4x60 */

/* flags of sstg_graph_build */
#define SSTG_GRAPH_NORM_LAP 1 /* norm_lap_matr=True (utils.py:48-50) */
#define SSTG_GRAPH_FP64 2     /* fp64 islands exactly as the reference (bit-exact vs networkx); default fp32 */
#define SSTG_GRAPH_V_ONLY 4   /* only the node tensor V (the A_tr the reference builds is never consumed) */

/* flags of sstg_sample_ade_fde */
#define SSTG_EVAL_FP64_DIST 1 /* per-step distance sqrt in fp64 exactly as math.sqrt in metrics.py:33,50 */
#define SSTG_EVAL_UNIFORM_N 4 /* every scene has the same pedestrian count (= total_peds / S): skips the scene search */
#define SSTG_EVAL_VTR_IS_REL 2 /* `Vtr` is pred_traj_rel [P,2,12] itself (utils.py:183), not the (12,N,2) V_tr blocks */

#define SSTG_OK 0
#define SSTG_EINVAL (-1)    /* bad argument (null pointer, negative size, misaligned) */
#define SSTG_ETOOBIG (-2)   /* a scene exceeds SSTG_MAX_PEDS */
#define SSTG_ECUDA (-3)     /* a CUDA runtime call failed; see sstg_last_cuda_error */
#define SSTG_EARCH (-4)     /* device is not sm_100 */

int sstg_version(void);
const char* sstg_error_string(int code);
/* text of the last CUDA error seen by THIS thread inside the library ("" if none) */
const char* sstg_last_cuda_error(void);

/* ---- host-side planning (pure CPU, no CUDA) ---------------------------------------
 * From the per-scene pedestrian counts compute the packed-batch offsets and the
 * TXP-CNN tile table (whole scenes per tile; at most SSTG_XTILE_COLS columns and
 * SSTG_XTILE_MAX_SCENES scenes and SSTG_XTILE_MAX_PAIRS column pairs, narrower -- down to 8 columns -- for batches too small to
 * fill the GPU with full tiles: cap = clamp(ceil(P / 296), 8, 128)).  seg_off_host[S+1], a_off_host[S+1], tile_off_host[S+1] are
 * caller-provided; tile t covers scenes [tile_off[t], tile_off[t+1]).
 * Outputs *n_tiles, *max_n, *a_floats (total floats of the packed A buffer, padded) and
 * *tile_cols (may be NULL): shared-memory columns of the widest tile, for sstg_txpcnn_forward.
 * Replaces the `seq_start_end` bookkeeping of TrajectoryDataset (utils.py:186-190). */
int sstg_plan_host(const int32_t* n_peds_host, int32_t S, int32_t T,
                   int32_t* seg_off_host, int64_t* a_off_host, int32_t* tile_off_host,
                   int32_t* n_tiles, int32_t* max_n, int64_t* a_floats, int32_t* tile_cols);

/* ---- (1) graph construction --------------------------------------------------------
 * Replaces utils.seq_to_graph + anorm (utils.py:23-53) incl. the networkx
 * normalized_laplacian_matrix call (utils.py:49-50), for every scene of a packed batch.
 *   rel   [P,2,T]   per-pedestrian relative displacements, the layout of
 *                   TrajectoryDataset.obs_traj_rel / pred_traj_rel (utils.py:180-183)
 *   V     scene s at float offset seg_off[s]*T*2, layout (T,N_s,2)   (utils.py:42)
 *   A     scene s at float offset a_off[s],       layout (T,N_s,N_s) (utils.py:43-50)
 * A may be NULL iff SSTG_GRAPH_V_ONLY. */
int sstg_graph_build(const float* rel, const int32_t* seg_off, const int64_t* a_off,
                     int32_t S, int32_t T, int32_t max_n,
                     float* V, float* A, int32_t flags, void* stream);

/* ---- (2) st_gcn block, eval mode ----------------------------------------------------
 * Replaces st_gcn.forward (model.py:145-155) = ConvTemporalGraphical (model.py:64-68:
 * 1x1 conv + einsum('nctv,tvw->nctw')) + BN + PReLU + (3,1) conv + BN + residual + PReLU,
 * BatchNorm in eval form from the running statistics `bn`.
 *   V  as written by sstg_graph_build with T = 8;  A likewise
 *   H  scene s at float offset seg_off[s]*40, layout (5,8,N_s)   [model.py:185 output]
 * `sm_count` <= 0 means "use the device's SM count" (persistent grid). */
int sstg_stgcn_forward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off,
                       int32_t S, int32_t max_n, const float* params, const float* bn,
                       float* H, int32_t sm_count, void* stream);

/* ---- (1)+(2) fused: graph construction + st_gcn without materialising A ------------------
 * For a packed evaluation that needs V_pred / ADE / FDE but not the adjacency itself
 * (test.py's loop): seq_to_graph (utils.py:23-53, fp32 mode, normalised Laplacian) and
 * st_gcn.forward (model.py:145-155, eval BN) computed straight from the displacements, A kept
 * in shared memory -- 32 N^2 bytes per scene neither written nor read.  N <= 64 only
 * (SSTG_ETOOBIG otherwise: use sstg_graph_build + sstg_stgcn_forward).
 *   rel  [P,2,8] obs_traj_rel;  Pws [P*40] scratch;  H as sstg_stgcn_forward */
int sstg_stgcn_forward_fused(const float* rel, const int32_t* seg_off, int32_t S, int32_t max_n,
                             const float* params, const float* bn, float* Pws, float* H, void* stream);

/* ---- (3) TXP-CNN --------------------------------------------------------------------
 * Replaces model.py:187-195: flat view (5,8,N)->(8,5,N), tpcnns[0..3] (3x3, time as
 * channels) with PReLU and residuals, tpcnn_ouput, flat view (12,5,N)->(5,12,N).
 *   H    as written by sstg_stgcn_forward
 *   out  scene s at float offset seg_off[s]*60, layout (5,12,N_s) == V_pred[0] of
 *        social_stgcnn.forward (model.py:198)
 *   tile_off [n_tiles+1] from sstg_plan_host (device copy); tile_cols from sstg_plan_host
 *   (0 = unknown: the kernel then reserves shared memory for the widest possible tile).
 * Inference runs the Winograd F(2,3) kernel (txpcnn_wino.cu: 64-pair tiles, or 128-pair tiles when the batch
 * holds scenes of more than 128 pedestrians), the training forward the direct FFMA2 kernel (txpcnn_fwd.cu).
 * NOTE (the library's only global state): when the batch fills the GPU (n_tiles >= 2 x SMs) the Winograd
 * kernel reads its transformed filters from a per-device __constant__ bank that this call refreshes in
 * stream order (the head of `out` serves as staging before the kernel overwrites it).  Calls with DIFFERENT
 * `params` must therefore not be in flight on two streams of one device at the same time; calls with the
 * same weights may.  Environment SSTG_TXP_FILTERS=smem selects the variant without the bank;
 * SSTG_TXP_PREFETCH=1 an experimental variant that fetches the next tile's input during the current tile
 * (same results; not the default). */
int sstg_txpcnn_forward(const float* H, const int32_t* seg_off, const int32_t* tile_off,
                        int32_t n_tiles, int32_t max_n, int32_t tile_cols, const float* params,
                        float* out, void* stream);

/* ---- (4a) bivariate Gaussian NLL, forward -------------------------------------------
 * Replaces metrics.bivariate_loss (metrics.py:84-113) per scene.
 *   Vpred  (5,12,N_s) blocks as written by sstg_txpcnn_forward
 *   Vtr    (12,N_s,2) blocks (sstg_graph_build on pred_traj_rel with T = 12)
 *   loss   [S] per-scene mean over the 12*N_s elements */
int sstg_nll_forward(const float* Vpred, const float* Vtr, const int32_t* seg_off, int32_t S,
                     float* loss, void* stream);

/* ---- (4b) K-sample trajectory sampling + ADE/FDE, min over samples ------------------
 * Replaces the per-scene body of test() (test.py:59-123) with metrics.ade / fde /
 * nodes_rel_to_nodes_abs (metrics.py:21-75): Cholesky factor of the 2x2 covariance,
 * sample_k = mean + L eps_k, rel->abs prefix sums from the last observed position,
 * per-pedestrian ADE/FDE, independent minima over the K samples.
 *   last_obs [P,2]     obs_traj[:, :, -1]                      (test.py:84-85); P = total_peds = seg_off[S]
 *   eps      scene s at float offset seg_off[s]*K*24, layout (K,12,N_s,2); standard-normal
 *            noise to inject.  NULL = draw in-kernel with Philox4x32-10 keyed by `seed`
 *            (counter = (k*6 + t/2, 0, global pedestrian index)).
 *   eps_out  optional, same layout: the noise actually used (for replay); may be NULL
 *   min_ade, min_fde [P] float64                                (test.py:121-123)
 *   pred_abs optional (K,12,N_s,2) blocks at seg_off[s]*K*24: raw_data_dict['pred'] (test.py:105) */
int sstg_sample_ade_fde(const float* Vpred, const float* Vtr, const float* last_obs,
                        const int32_t* seg_off, int32_t S, int32_t total_peds, int32_t K,
                        const float* eps, uint64_t seed, float* eps_out,
                        double* min_ade, double* min_fde, float* pred_abs,
                        int32_t flags, void* stream);

/* ---- stand-alone trajectory helpers (the reference's metrics.py API) ------------------
 * sstg_rel_to_abs: metrics.nodes_rel_to_nodes_abs (metrics.py:69-75) for packed (T,N_s,2)
 * blocks (scene s at float offset seg_off[s]*T*2): out[t,p] = init[p] + sum_{u<=t} rel[u,p]
 * with the reference's rounding points.  init [P,2] float64 (the reference passes a float64 array). */
int sstg_rel_to_abs(const float* rel, const double* init, const int32_t* seg_off, int32_t S,
                    int32_t total_peds, int32_t T, float* out, void* stream);

/* sstg_traj_error: the inner sums of metrics.ade / metrics.fde (metrics.py:21-53) per
 * pedestrian for packed (T,N_s,2) blocks: sum_dist[p] = sum_t ||pred - tgt|| (fp64),
 * final_dist[p] = ||pred - tgt|| at t = T-1. */
int sstg_traj_error(const float* pred, const float* tgt, const int32_t* seg_off, int32_t S,
                    int32_t total_peds, int32_t T, double* sum_dist, double* final_dist, void* stream);

/* ---- compact host format (input side of the path) ---------------------------------------
 * The dataset's positions are 4-decimal numbers (utils.py:145), so a 20-frame track is exactly
 * an int32 start position plus 19 int16 steps, both in units of 1e-4: 88 bytes per pedestrian
 * instead of the 168 bytes of obs_traj_rel + pred_traj_rel + obs_traj[:, :, -1] in fp32.
 * sstg_decode_tracks reproduces those three tensors BIT FOR BIT on the device (float64 k/1e4 and
 * differences, cast to fp32: utils.py:145,153-155,180-183), halving the PCIe traffic of a
 * host-fed evaluation.
 *   start [P,2] int32;  delta [P,2,20] int16 (delta[.,.,0] ignored, 16-byte rows for alignment)
 *   obs_rel [P,2,8], pred_rel [P,2,12], last_obs [P,2]  fp32 outputs */
int sstg_decode_tracks(const int32_t* start, const int16_t* delta, int32_t total_peds,
                       float* obs_rel, float* pred_rel, float* last_obs, void* stream);

/* ===================== training path (train.py:28-79) =================================
 * The reference trains with batch 1 and accumulates the loss of `batch_size` scenes before
 * one backward (train.py:58-74).  A packed batch reproduces that exactly: BatchNorm uses
 * the statistics of each scene separately (train mode, model.py:114,121,141), the loss
 * weights carry the 1/batch_size and the "last scene of the group is not scored" rule
 * (train.py:58-66), and the gradient of the whole group comes out of one backward. */

/* Rows of the two partial-gradient buffers sstg_backward needs for a batch of S scenes in
 * n_tiles TXP-CNN tiles on the current device: partial_st[rows_st][SSTG_ST_PARAMS],
 * partial_tx[rows_tx][SSTG_TXP_GRAD_STRIDE]. */
int sstg_train_partial_rows(int32_t S, int32_t n_tiles, int32_t* rows_st, int32_t* rows_tx);

/* social_stgcnn.forward in TRAIN mode (model.py:182-198 with nn.BatchNorm2d batch statistics).
 *   save         [total_peds * SSTG_SAVE_FLOATS_PER_PED] activations for sstg_backward
 *   scene_stats  [S][30]  per scene: mean,var(biased) of BN tcn.0, tcn.3, residual.1
 *   Vpred        (5,12,N_s) blocks like sstg_txpcnn_forward */
int sstg_forward_train(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off,
                       const int32_t* tile_off, int32_t S, int32_t n_tiles, int32_t max_n, int32_t total_peds,
                       const float* params, float* save, float* scene_stats, float* Vpred, void* stream);

/* BatchNorm running statistics after forwarding the S scenes in order (momentum 0.1, unbiased
 * variance -- torch.nn.BatchNorm2d as used at model.py:114,121,141).  bn[30] is updated in place. */
int sstg_bn_running_update(const float* scene_stats, const int32_t* seg_off, int32_t S, double momentum,
                           float* bn, void* stream);

/* metrics.bivariate_loss forward + its gradient (metrics.py:84-113 under autograd):
 *   loss[s]   per-scene mean NLL;   dVpred = scene_weight[s] * d loss[s] / d Vpred  ((5,12,N_s) blocks)
 *   scene_weight may be NULL (= 1).  For train.py's accumulation: 1/batch_size, 0 for the last
 *   scene of each optimiser group (train.py:58-66). */
int sstg_nll_backward(const float* Vpred, const float* Vtr, const int32_t* seg_off, int32_t S,
                      const float* scene_weight, float* loss, float* dVpred, void* stream);

/* Backward of the whole model (loss.backward(), train.py:67): from dVpred to the flat gradient
 * grad[SSTG_N_PARAMS] in named_parameters() order (zeros for tpcnns.4 / prelus.4, which
 * model.py:191 never executes).  dH_ws [total_peds*40] is scratch.  Deterministic: no atomics. */
int sstg_backward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off,
                  const int32_t* tile_off, int32_t S, int32_t n_tiles, int32_t max_n, int32_t total_peds,
                  const float* params, const float* save, const float* scene_stats, const float* dVpred,
                  float* dH_ws, float* partial_st, float* partial_tx, float* grad, void* stream);

/* optimizer.step() of train.py:197 (SGD, lr only) with the optional clip_grad_norm_ of
 * train.py:69-70 (clip_norm <= 0: off).  grad_norm_out (1 float, may be NULL) gets ||grad||_2. */
int sstg_sgd_step(float* params, const float* grad, int32_t n, float lr, float clip_norm,
                  float* grad_norm_out, void* stream);

/* ---- one optimiser group of train() in ONE call (train.py:38-74 for `batch_size` scenes) --------
 * Same results as: graph build (obs: V + A) -> sstg_forward_train -> sstg_nll_backward -> sstg_backward ->
 * sstg_bn_running_update -> [sstg_sgd_step if lr != 0], all on `stream` -- issued as 7 kernels chained by
 * programmatic dependent launch (the NLL kernel reads pred_rel directly, the BatchNorm recurrence rides in the
 * TXP-CNN forward launch, gradient reduction + group loss + un-clipped SGD share the last launch).
 *   obs_rel [P,2,8], pred_rel [P,2,12]     TrajectoryDataset.obs_traj_rel / pred_traj_rel rows of the group
 *   params [7563], bn [30]                 updated in place
 *   scene_weight [S]                       1/batch_size, 0 for the unscored last scene (may be NULL = 1)
 *   ws                                     caller-owned scratch of >= sstg_train_step_workspace() floats, 16-byte aligned
 *   loss [S] per-scene NLL, grad [7563], group_loss [1] = sum_s weight[s]*loss[s] (may be NULL)
 *   graph_flags                            SSTG_GRAPH_FP64 or 0 (NORM_LAP is implied)
 *   bn_momentum                            0.1 in the reference; NEGATIVE = leave bn[] alone (a scene-sharded
 *                                          group applies the recurrence after the exchange, see below); the
 *                                          statistics then stay in ws at sstg_train_step_stats_offset() floats */
int64_t sstg_train_step_workspace(int32_t S, int32_t n_tiles, int32_t total_peds, int64_t a_floats);
int64_t sstg_train_step_stats_offset(int32_t S, int32_t n_tiles, int32_t total_peds, int64_t a_floats);
int sstg_train_step(const float* obs_rel, const float* pred_rel, const int32_t* seg_off, const int64_t* a_off,
                    const int32_t* tile_off, int32_t S, int32_t n_tiles, int32_t max_n, int32_t total_peds,
                    int64_t a_floats, float* params, float* bn, const float* scene_weight, float* ws,
                    int64_t ws_floats, float lr, float clip_norm, double bn_momentum, int32_t graph_flags,
                    float* loss, float* grad, float* group_loss, void* stream);

/* ---- one optimiser group sharded over the GPUs of a node: gradient exchange over NVLink peer memory,
 * fused with the optimiser step (no reference counterpart: train.py runs on one device; this is the
 * exchange step of SURVEY 8e).  Every rank allocates one MAILBOX (sstg_peer_alloc), the 64-byte CUDA IPC
 * handles are exchanged by the host (any transport), every rank maps the others' (sstg_peer_open).
 * Per group each rank runs sstg_train_step(lr = 0, bn_momentum < 0) on its contiguous shard with the
 * group's scene weights, then sstg_peer_exchange_apply, which in ONE kernel stores its gradient / loss /
 * own statistics rows into every mailbox, signals, waits for all sources on its local flags, sums the
 * gradients in rank order (bit-identical replicas), clips and applies SGD; the BatchNorm recurrence over the
 * GLOBAL scene order follows on the same stream.
 *   mailboxes [world]   HOST array of device pointers, [rank] = this rank's own mailbox
 *   epoch               1, 2, 3, ... identical on all ranks, one per call (selects the payload parity)
 *   max_scenes          the capacity the mailboxes were sized for (same on all ranks)
 *   grad_local [7563], loss_local [1] (NULL = 0), stats_local [s_local,30]: this rank's contribution;
 *   scene_lo, s_local   its scene range inside the group of s_all scenes; seg_off_all [s_all+1] (device)
 *   grad_out [7563], loss_out [1] (may be NULL), stats_out [s_all,30]: the reduced group quantities
 *   status [1] int32    device word, 0 on success; 1+r if source r had not delivered within timeout_ns
 *                       (0 = 30 s): then nothing is applied -- the kernel never spins unbounded.  The word is
 *                       sticky: while it is non-zero this rank's calls neither push nor apply, so its peers
 *                       time out as well and no replica trains on (the caller clears it to recover).
 * sstg_peer_exchange_apply_local: the same exchange for a process that holds n_local ranks of the group on
 * ONE device (single-GPU test harness): their CTAs wait on one another, so they run as ONE cooperative
 * launch instead of n_local launches that are not guaranteed to be co-resident. */
#define SSTG_MAX_RANKS 8
#define SSTG_PEER_HANDLE_BYTES 64
int64_t sstg_peer_mailbox_bytes(int32_t world, int32_t max_scenes);
int sstg_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle);
int sstg_peer_open(const unsigned char* handle, void** ptr);
int sstg_peer_close(void* ptr);
int sstg_peer_free(void* ptr);
typedef struct sstg_peer_rank {
  int32_t rank, scene_lo, s_local, reserved;
  const float* grad_local;
  const float* loss_local;
  const float* stats_local;
  const int32_t* seg_off_all;
  float* params;
  float* bn;
  float* grad_out;
  float* loss_out;
  float* stats_out;
  int32_t* status;
} sstg_peer_rank;
int sstg_peer_exchange_apply_local(void* const* mailboxes, int32_t world, const sstg_peer_rank* ranks,
                                   int32_t n_local, uint32_t epoch, int32_t max_scenes, int32_t s_all, float lr,
                                   float clip_norm, double bn_momentum, uint64_t timeout_ns, void* stream);
int sstg_peer_exchange_apply(void* const* mailboxes, int32_t world, int32_t rank, uint32_t epoch,
                             int32_t max_scenes, const float* grad_local, const float* loss_local,
                             const float* stats_local, int32_t scene_lo, int32_t s_local,
                             const int32_t* seg_off_all, int32_t s_all, float* params, float* bn, float lr,
                             float clip_norm, double bn_momentum, float* grad_out, float* loss_out,
                             float* stats_out, int32_t* status, uint64_t timeout_ns, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSTG_H_ */
