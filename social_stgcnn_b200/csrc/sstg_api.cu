// extern "C" surface of libsstg.so (declared in include/sstg.h): argument validation,
// host-side planning, and dispatch to the kernel launchers.  No torch types, no
// allocation, no synchronisation.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sstg_common.cuh"

namespace sstg {

static thread_local char g_last_cuda_error[256] = "";

void note_cuda_error(cudaError_t e, const char* where) {
  std::snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", where, cudaGetErrorName(e),
                cudaGetErrorString(e));
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sstg_xtile_pairs() {
  const char* e = std::getenv("SSTG_XTILE_PAIRS");
  return (e && e[0] == '3' && e[1] == '2') ? 32 : SSTG_XTILE_MAX_PAIRS;
}

bool pdl_enabled() {
  const char* e = std::getenv("SSTG_NO_PDL");
  return !(e && e[0] == '1');
}

}  // namespace sstg

using namespace sstg;

extern "C" {

int sstg_version(void) { return 110; }  // 110: + sstg_peer_exchange_apply_local, sticky exchange status

const char* sstg_error_string(int code) {
  switch (code) {
    case SSTG_OK: return "ok";
    case SSTG_EINVAL: return "invalid argument";
    case SSTG_ETOOBIG: return "scene larger than SSTG_MAX_PEDS";
    case SSTG_ECUDA: return "CUDA runtime error";
    case SSTG_EARCH: return "device is not sm_100";
    default: return "unknown error";
  }
}

const char* sstg_last_cuda_error(void) { return g_last_cuda_error; }

int sstg_plan_host(const int32_t* n_peds_host, int32_t S, int32_t T, int32_t* seg_off_host, int64_t* a_off_host,
                   int32_t* tile_off_host, int32_t* n_tiles, int32_t* max_n, int64_t* a_floats,
                   int32_t* tile_cols) {
  if (S < 0 || T <= 0 || (S > 0 && !n_peds_host) || !seg_off_host || !a_off_host || !tile_off_host || !n_tiles ||
      !max_n || !a_floats)
    return SSTG_EINVAL;
  // TXP-CNN tile width: whole scenes up to SSTG_XTILE_COLS columns -- but a SMALL batch is cut into
  // narrower tiles so that it still spreads over the SMs (a tile is one CTA): aim at two tiles per
  // SM of a 148-SM part, never below 8 columns.  Large batches get the full 128-column tiles.
  int64_t total = 0;
  for (int32_t s = 0; s < S; ++s) total += n_peds_host[s] > 0 ? n_peds_host[s] : 0;
  int64_t cap64 = (total + 295) / 296;
  int32_t cap = static_cast<int32_t>(cap64 < 8 ? 8 : (cap64 > SSTG_XTILE_COLS ? SSTG_XTILE_COLS : cap64));
  const int32_t max_pairs = ::sstg::sstg_xtile_pairs();
  if (cap > 2 * max_pairs) cap = 2 * max_pairs;
  int64_t ped = 0, af = 0;
  int32_t mx = 0, nt = 0, cols = 0, scenes = 0, pairs = 0;
  int32_t wcol = 1, wmax = 0;  // shared-memory columns of the current tile in the odd-start layout of txpcnn_wino.cu
  seg_off_host[0] = 0;
  a_off_host[0] = 0;
  tile_off_host[0] = 0;
  for (int32_t s = 0; s < S; ++s) {
    const int32_t n = n_peds_host[s];
    if (n < 1) return SSTG_EINVAL;
    if (n > SSTG_MAX_PEDS) return SSTG_ETOOBIG;
    if (n > mx) mx = n;
    // close the current TXP-CNN tile if this scene does not fit any more
    if (scenes > 0 && (cols + n > cap || scenes == SSTG_XTILE_MAX_SCENES ||
                       pairs + ((n + 1) >> 1) > max_pairs)) {
      tile_off_host[++nt] = s;
      cols = 0;
      scenes = 0;
      pairs = 0;
      wcol = 1;
    }
    cols += n;
    pairs += (n + 1) >> 1;
    wcol += n + 1;
    wcol += (~wcol) & 1;  // the next scene starts at an odd column
    if (wcol + 1 > wmax) wmax = wcol + 1;
    ++scenes;
    ped += n;
    if (ped > INT32_MAX / 2) return SSTG_ETOOBIG;  // pedestrian indices are int32 (element offsets are 64-bit)
    seg_off_host[s + 1] = static_cast<int32_t>(ped);
    af += static_cast<int64_t>(T) * n * n;
    af = (af + 3) & ~int64_t(3);  // every scene block starts 16-byte aligned
    a_off_host[s + 1] = af;
  }
  if (S > 0) tile_off_host[++nt] = S;
  *n_tiles = nt;
  *max_n = mx;
  *a_floats = af;
  if (tile_cols) *tile_cols = wmax;
  return SSTG_OK;
}

int sstg_graph_build(const float* rel, const int32_t* seg_off, const int64_t* a_off, int32_t S, int32_t T,
                     int32_t max_n, float* V, float* A, int32_t flags, void* stream) {
  const bool v_only = flags & SSTG_GRAPH_V_ONLY;
  if (S < 0 || T <= 0 || T > 64) return SSTG_EINVAL;
  if (S == 0) return SSTG_OK;  // an empty batch is a no-op whatever the other arguments
  if (max_n < 1) return SSTG_EINVAL;
  if (max_n > SSTG_MAX_PEDS) return SSTG_ETOOBIG;
  if (!rel || !seg_off || !V || (!v_only && (!A || !a_off))) return SSTG_EINVAL;
  if (!aligned16(V) || (!v_only && !aligned16(A))) return SSTG_EINVAL;
  return launch_graph_build(rel, seg_off, a_off, S, T, max_n, V, A, flags, static_cast<cudaStream_t>(stream));
}

int sstg_stgcn_forward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off, int32_t S,
                       int32_t max_n, const float* params, const float* bn, float* H, int32_t sm_count,
                       void* stream) {
  if (S < 0) return SSTG_EINVAL;
  if (S == 0) return SSTG_OK;
  if (max_n < 1) return SSTG_EINVAL;
  if (max_n > SSTG_MAX_PEDS) return SSTG_ETOOBIG;
  if (!V || !A || !seg_off || !a_off || !params || !bn || !H) return SSTG_EINVAL;
  if (!aligned16(V) || !aligned16(A) || !aligned16(H)) return SSTG_EINVAL;
  return launch_stgcn_forward(V, A, seg_off, a_off, S, max_n, params, bn, H, sm_count,
                              static_cast<cudaStream_t>(stream));
}

int sstg_stgcn_forward_fused(const float* rel, const int32_t* seg_off, int32_t S, int32_t max_n,
                             const float* params, const float* bn, float* Pws, float* H, void* stream) {
  if (S < 0) return SSTG_EINVAL;
  if (S == 0) return SSTG_OK;
  if (max_n < 1) return SSTG_EINVAL;
  if (max_n > 64) return SSTG_ETOOBIG;
  if (!rel || !seg_off || !params || !bn || !Pws || !H || !aligned16(rel)) return SSTG_EINVAL;
  return launch_stgcn_forward_fused(rel, seg_off, S, max_n, params, bn, Pws, H, static_cast<cudaStream_t>(stream));
}

int sstg_txpcnn_forward(const float* H, const int32_t* seg_off, const int32_t* tile_off, int32_t n_tiles,
                        int32_t max_n, int32_t tile_cols, const float* params, float* out, void* stream) {
  if (n_tiles < 0) return SSTG_EINVAL;
  if (n_tiles == 0) return SSTG_OK;
  if (max_n < 1) return SSTG_EINVAL;
  if (max_n > SSTG_MAX_PEDS) return SSTG_ETOOBIG;
  if (!H || !seg_off || !tile_off || !params || !out) return SSTG_EINVAL;
  return launch_txpcnn_forward(H, seg_off, tile_off, n_tiles, max_n, params, out, nullptr, nullptr, 0, tile_cols,
                               static_cast<cudaStream_t>(stream));
}

int sstg_nll_forward(const float* Vpred, const float* Vtr, const int32_t* seg_off, int32_t S, float* loss,
                     void* stream) {
  if (S < 0) return SSTG_EINVAL;
  if (S == 0) return SSTG_OK;
  if (!Vpred || !Vtr || !seg_off || !loss) return SSTG_EINVAL;
  return launch_nll_forward(Vpred, Vtr, seg_off, S, loss, static_cast<cudaStream_t>(stream));
}

int sstg_sample_ade_fde(const float* Vpred, const float* Vtr, const float* last_obs, const int32_t* seg_off,
                        int32_t S, int32_t total_peds, int32_t K, const float* eps, uint64_t seed, float* eps_out,
                        double* min_ade, double* min_fde, float* pred_abs, int32_t flags, void* stream) {
  if (S < 0 || total_peds < 0 || K < 1) return SSTG_EINVAL;
  if (S == 0 || total_peds == 0) return SSTG_OK;
  if (!Vpred || !Vtr || !last_obs || !seg_off || !min_ade || !min_fde) return SSTG_EINVAL;
  return launch_sample_ade_fde(Vpred, Vtr, last_obs, seg_off, S, total_peds, K, eps, seed, eps_out, min_ade,
                               min_fde, pred_abs, flags, static_cast<cudaStream_t>(stream));
}

int sstg_rel_to_abs(const float* rel, const double* init, const int32_t* seg_off, int32_t S, int32_t total_peds,
                    int32_t T, float* out, void* stream) {
  if (S < 0 || total_peds < 0 || T < 1) return SSTG_EINVAL;
  if (S == 0 || total_peds == 0) return SSTG_OK;
  if (!rel || !init || !seg_off || !out) return SSTG_EINVAL;
  return launch_rel_to_abs(rel, init, seg_off, S, total_peds, T, out, static_cast<cudaStream_t>(stream));
}

int sstg_traj_error(const float* pred, const float* tgt, const int32_t* seg_off, int32_t S, int32_t total_peds,
                    int32_t T, double* sum_dist, double* final_dist, void* stream) {
  if (S < 0 || total_peds < 0 || T < 1) return SSTG_EINVAL;
  if (S == 0 || total_peds == 0) return SSTG_OK;
  if (!pred || !tgt || !seg_off || !sum_dist || !final_dist) return SSTG_EINVAL;
  return launch_traj_error(pred, tgt, seg_off, S, total_peds, T, sum_dist, final_dist,
                           static_cast<cudaStream_t>(stream));
}

int sstg_decode_tracks(const int32_t* start, const int16_t* delta, int32_t total_peds, float* obs_rel,
                       float* pred_rel, float* last_obs, void* stream) {
  if (total_peds < 0) return SSTG_EINVAL;
  if (total_peds == 0) return SSTG_OK;
  if (!start || !delta || !obs_rel || !pred_rel || !last_obs) return SSTG_EINVAL;
  if (!aligned16(delta) || !aligned16(obs_rel) || !aligned16(pred_rel)) return SSTG_EINVAL;
  return launch_decode_tracks(start, delta, total_peds, obs_rel, pred_rel, last_obs,
                              static_cast<cudaStream_t>(stream));
}

// ---- training path ------------------------------------------------------------------
int sstg_train_partial_rows(int32_t S, int32_t n_tiles, int32_t* rows_st, int32_t* rows_tx) {
  if (S < 0 || n_tiles < 0 || !rows_st || !rows_tx) return SSTG_EINVAL;
  *rows_st = stgcn_train_bwd_rows(S);
  *rows_tx = txpcnn_bwd_rows(n_tiles);
  return SSTG_OK;
}

int sstg_forward_train(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off,
                       const int32_t* tile_off, int32_t S, int32_t n_tiles, int32_t max_n, int32_t total_peds,
                       const float* params, float* save, float* scene_stats, float* Vpred, void* stream) {
  if (S < 0 || n_tiles < 0 || total_peds < 0) return SSTG_EINVAL;
  if (S == 0) return SSTG_OK;
  if (max_n < 1) return SSTG_EINVAL;
  if (max_n > SSTG_MAX_PEDS) return SSTG_ETOOBIG;
  if (!V || !A || !seg_off || !a_off || !tile_off || !params || !save || !scene_stats || !Vpred) return SSTG_EINVAL;
  if (!aligned16(V)) return SSTG_EINVAL;
  const long long P = total_peds;
  float* Y = save;
  float* Z = save + P * 40;
  float* H = save + P * 80;
  float* U = save + P * 120;
  float* X = save + P * 360;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = launch_stgcn_train_forward(V, A, seg_off, a_off, S, max_n, params, Y, Z, H, scene_stats, st);
  if (rc != SSTG_OK) return rc;
  return launch_txpcnn_forward(H, seg_off, tile_off, n_tiles, max_n, params, Vpred, U, X, P * 60, 0, st);
}

int sstg_bn_running_update(const float* scene_stats, const int32_t* seg_off, int32_t S, double momentum,
                           float* bn, void* stream) {
  if (S < 0 || !(momentum >= 0.0 && momentum <= 1.0)) return SSTG_EINVAL;
  if (S == 0) return SSTG_OK;
  if (!scene_stats || !seg_off || !bn) return SSTG_EINVAL;
  return launch_bn_running_update(scene_stats, seg_off, S, momentum, bn, static_cast<cudaStream_t>(stream));
}

int sstg_nll_backward(const float* Vpred, const float* Vtr, const int32_t* seg_off, int32_t S,
                      const float* scene_weight, float* loss, float* dVpred, void* stream) {
  if (S < 0) return SSTG_EINVAL;
  if (S == 0) return SSTG_OK;
  if (!Vpred || !Vtr || !seg_off || !loss || !dVpred) return SSTG_EINVAL;
  return launch_nll_fwd_bwd(Vpred, Vtr, seg_off, S, scene_weight, loss, dVpred, static_cast<cudaStream_t>(stream));
}

int sstg_backward(const float* V, const float* A, const int32_t* seg_off, const int64_t* a_off,
                  const int32_t* tile_off, int32_t S, int32_t n_tiles, int32_t max_n, int32_t total_peds,
                  const float* params, const float* save, const float* scene_stats, const float* dVpred,
                  float* dH_ws, float* partial_st, float* partial_tx, float* grad, void* stream) {
  if (S < 0 || n_tiles < 0 || total_peds < 0 || (S > 0 && max_n < 1)) return SSTG_EINVAL;
  if (max_n > SSTG_MAX_PEDS) return SSTG_ETOOBIG;
  if (!grad) return SSTG_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (S == 0) {
    SSTG_CUDA_TRY(cudaMemsetAsync(grad, 0, sizeof(float) * SSTG_N_PARAMS, st), "grad memset");
    return SSTG_OK;
  }
  if (!V || !A || !seg_off || !a_off || !tile_off || !params || !save || !scene_stats || !dVpred || !dH_ws ||
      !partial_st || !partial_tx)
    return SSTG_EINVAL;
  const long long P = total_peds;
  const float* Y = save;
  const float* Z = save + P * 40;
  const float* H = save + P * 80;
  const float* U = save + P * 120;
  const float* X = save + P * 360;
  int rc = launch_txpcnn_backward(dVpred, U, X, P * 60, H, seg_off, tile_off, n_tiles, max_n, params, dH_ws,
                                  partial_tx, st);
  if (rc != SSTG_OK) return rc;
  rc = launch_stgcn_train_backward(V, A, seg_off, a_off, S, max_n, params, Y, Z, scene_stats, dH_ws, partial_st, st);
  if (rc != SSTG_OK) return rc;
  return launch_grad_finalize(partial_st, stgcn_train_bwd_rows(S), partial_tx, txpcnn_bwd_rows(n_tiles),
                              SSTG_TXP_GRAD_STRIDE, grad, st);
}

int sstg_sgd_step(float* params, const float* grad, int32_t n, float lr, float clip_norm, float* grad_norm_out,
                  void* stream) {
  if (n < 0 || !params || !grad) return SSTG_EINVAL;
  if (n == 0) return SSTG_OK;
  return launch_sgd_step(params, grad, n, lr, clip_norm, grad_norm_out, static_cast<cudaStream_t>(stream));
}

namespace {
struct TrainWs {
  int64_t V, A, Vtr, save, stats, vp, dvp, dH, pst, ptx, total;
};
inline int64_t up4(int64_t x) { return (x + 3) & ~int64_t(3); }
TrainWs train_ws_layout(int32_t S, int32_t n_tiles, int32_t P, int64_t a_floats) {
  TrainWs w;
  int64_t o = 0;
  w.V = o;      o += up4(int64_t(P) * 16);
  w.A = o;      o += up4(a_floats);
  w.Vtr = o;    o += up4(int64_t(P) * 24);
  w.save = o;   o += up4(int64_t(P) * SSTG_SAVE_FLOATS_PER_PED);
  w.stats = o;  o += up4(int64_t(S) * SSTG_N_BN);
  w.vp = o;     o += up4(int64_t(P) * 60);
  w.dvp = o;    o += up4(int64_t(P) * 60);
  w.dH = o;     o += up4(int64_t(P) * 40);
  w.pst = o;    o += up4(int64_t(stgcn_train_bwd_rows(S)) * SSTG_ST_PARAMS);
  w.ptx = o;    o += up4(int64_t(txpcnn_bwd_rows(n_tiles)) * SSTG_TXP_GRAD_STRIDE);
  w.total = o;
  return w;
}
}  // namespace

int64_t sstg_train_step_workspace(int32_t S, int32_t n_tiles, int32_t total_peds, int64_t a_floats) {
  if (S < 0 || n_tiles < 0 || total_peds < 0 || a_floats < 0) return SSTG_EINVAL;
  return train_ws_layout(S, n_tiles, total_peds, a_floats).total;
}

int64_t sstg_train_step_stats_offset(int32_t S, int32_t n_tiles, int32_t total_peds, int64_t a_floats) {
  if (S < 0 || n_tiles < 0 || total_peds < 0 || a_floats < 0) return SSTG_EINVAL;
  return train_ws_layout(S, n_tiles, total_peds, a_floats).stats;
}

int sstg_train_step(const float* obs_rel, const float* pred_rel, const int32_t* seg_off, const int64_t* a_off,
                    const int32_t* tile_off, int32_t S, int32_t n_tiles, int32_t max_n, int32_t total_peds,
                    int64_t a_floats, float* params, float* bn, const float* scene_weight, float* ws,
                    int64_t ws_floats, float lr, float clip_norm, double bn_momentum, int32_t graph_flags,
                    float* loss, float* grad, float* group_loss, void* stream) {
  if (S < 0 || n_tiles < 0 || total_peds < 0 || a_floats < 0 || (S > 0 && max_n < 1)) return SSTG_EINVAL;
  if (max_n > SSTG_MAX_PEDS) return SSTG_ETOOBIG;
  if (!params || !bn || !grad || !loss || !(bn_momentum <= 1.0)) return SSTG_EINVAL;
  if (S == 0) {
    if (group_loss)
      SSTG_CUDA_TRY(cudaMemsetAsync(group_loss, 0, sizeof(float), static_cast<cudaStream_t>(stream)), "loss memset");
    return sstg_backward(nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 1, 0, nullptr, nullptr, nullptr, nullptr,
                         nullptr, nullptr, nullptr, grad, stream);
  }
  if (!obs_rel || !pred_rel || !seg_off || !a_off || !tile_off || !ws || !aligned16(ws)) return SSTG_EINVAL;
  const TrainWs w = train_ws_layout(S, n_tiles, total_peds, a_floats);
  if (ws_floats < w.total) return SSTG_EINVAL;
  const int gf = SSTG_GRAPH_NORM_LAP | (graph_flags & SSTG_GRAPH_FP64);
  int rc = sstg_graph_build(obs_rel, seg_off, a_off, S, SSTG_T_OBS, max_n, ws + w.V, ws + w.A, gf, stream);
  if (rc) return rc;
  // (no V_tr pass: the NLL kernel reads the targets straight out of pred_traj_rel)
  {
    // forward: the BatchNorm running-statistics recurrence rides along as an extra CTA of the TXP-CNN launch
    const long long P = total_peds;
    float* save = ws + w.save;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = launch_stgcn_train_forward(ws + w.V, ws + w.A, seg_off, a_off, S, max_n, params, save, save + P * 40,
                                    save + P * 80, ws + w.stats, st);
    if (rc) return rc;
    rc = launch_txpcnn_forward(save + P * 80, seg_off, tile_off, n_tiles, max_n, params, ws + w.vp, save + P * 120,
                               save + P * 360, P * 60, 0, st, ws + w.stats, S, bn_momentum,
                               bn_momentum >= 0.0 ? bn : nullptr);
    if (rc) return rc;
    rc = launch_nll_fwd_bwd(ws + w.vp, pred_rel, seg_off, S, scene_weight, loss, ws + w.dvp, st, /*vtr_is_rel=*/1);
    if (rc) return rc;
  }
  // backward without its own gradient reduction: the reduction, the BN recurrence, the group loss and
  // (without clipping) the SGD update share ONE launch
  {
    const long long P = total_peds;
    const float* save = ws + w.save;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = launch_txpcnn_backward(ws + w.dvp, save + P * 120, save + P * 360, P * 60, save + P * 80, seg_off, tile_off,
                                n_tiles, max_n, params, ws + w.dH, ws + w.ptx, st);
    if (rc) return rc;
    rc = launch_stgcn_train_backward(ws + w.V, ws + w.A, seg_off, a_off, S, max_n, params, save, save + P * 40,
                                     ws + w.stats, ws + w.dH, ws + w.pst, st);
    if (rc) return rc;
    const bool clip = clip_norm > 0.f;
    rc = launch_train_tail(ws + w.pst, stgcn_train_bwd_rows(S), ws + w.ptx, txpcnn_bwd_rows(n_tiles),
                           SSTG_TXP_GRAD_STRIDE, grad, params, clip ? 0.f : lr, ws + w.stats, seg_off, S,
                           /*BN recurrence already done=*/-1.0, bn, loss, scene_weight, group_loss, st);
    if (rc) return rc;
    if (clip && lr != 0.f) rc = sstg_sgd_step(params, grad, SSTG_N_PARAMS, lr, clip_norm, nullptr, stream);
  }
  return rc;
}

}  // extern "C"
