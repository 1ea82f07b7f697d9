"""CPU oracle for the Social-STGCNN hot path -- TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (numpy for the fp64 "islands", torch-functional
fp32 for the convolutional arithmetic) of the reference algorithm on the path
    seq_to_graph -> social_stgcnn.forward -> bivariate_loss / test() sampling + ade/fde.
It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product (`social_stgcnn_b200`) never imports it and has no CPU fallback.

Parity pin: every function below is checked against outputs of the real
reference (imported from /root/reference in the build container) by
`tests/golden/make_golden.py`; the resulting vectors are committed under
`tests/golden/*.npz` and re-checked by `tests/test_oracle_golden.py`
(`-m "not gpu"`).  The reference itself ships no tests or golden vectors
(SURVEY.md section 4), so those generated fixtures are the pin.

Every function cites the reference file:line it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

T_OBS = 8
T_PRED = 12
C_IN = 2
C_OUT = 5
N_PARAMS = 7563

# named_parameters() order of reference model.py:157-178.
PARAM_LAYOUT = [
    ("st_gcns.0.gcn.conv.weight", (5, 2, 1, 1)),
    ("st_gcns.0.gcn.conv.bias", (5,)),
    ("st_gcns.0.tcn.0.weight", (5,)),
    ("st_gcns.0.tcn.0.bias", (5,)),
    ("st_gcns.0.tcn.1.weight", (1,)),
    ("st_gcns.0.tcn.2.weight", (5, 5, 3, 1)),
    ("st_gcns.0.tcn.2.bias", (5,)),
    ("st_gcns.0.tcn.3.weight", (5,)),
    ("st_gcns.0.tcn.3.bias", (5,)),
    ("st_gcns.0.residual.0.weight", (5, 2, 1, 1)),
    ("st_gcns.0.residual.0.bias", (5,)),
    ("st_gcns.0.residual.1.weight", (5,)),
    ("st_gcns.0.residual.1.bias", (5,)),
    ("st_gcns.0.prelu.weight", (1,)),
    ("tpcnns.0.weight", (12, 8, 3, 3)),
    ("tpcnns.0.bias", (12,)),
    ("tpcnns.1.weight", (12, 12, 3, 3)),
    ("tpcnns.1.bias", (12,)),
    ("tpcnns.2.weight", (12, 12, 3, 3)),
    ("tpcnns.2.bias", (12,)),
    ("tpcnns.3.weight", (12, 12, 3, 3)),
    ("tpcnns.3.bias", (12,)),
    ("tpcnns.4.weight", (12, 12, 3, 3)),
    ("tpcnns.4.bias", (12,)),
    ("tpcnn_ouput.weight", (12, 12, 3, 3)),
    ("tpcnn_ouput.bias", (12,)),
    ("prelus.0.weight", (1,)),
    ("prelus.1.weight", (1,)),
    ("prelus.2.weight", (1,)),
    ("prelus.3.weight", (1,)),
    ("prelus.4.weight", (1,)),
]
BN_LAYOUT = [
    "st_gcns.0.tcn.0.running_mean", "st_gcns.0.tcn.0.running_var",
    "st_gcns.0.tcn.3.running_mean", "st_gcns.0.tcn.3.running_var",
    "st_gcns.0.residual.1.running_mean", "st_gcns.0.residual.1.running_var",
]
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------
# (1) graph construction                                   utils.py:23-53
# --------------------------------------------------------------------------
def seq_to_graph(seq_rel, norm_lap_matr: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """`seq_to_graph` (utils.py:29-53) with `anorm` (utils.py:23-27) and
    networkx `normalized_laplacian_matrix` (call site utils.py:49-50) restated
    with the reference's rounding points:

      * dx, dy, dx^2, dy^2 and their sum are separate fp32 roundings (0-dim
        torch fp32 tensors, utils.py:24);
      * sqrt, reciprocal, the row sum d_h = sum_k W_hk (column-ascending, the
        order scipy's csr mat-vec uses), ds = 1/sqrt(d) and
        ds_h * ((D-W)_hk * ds_k) are fp64 (networkx laplacianmatrix.py);
      * the final cast to fp32 is utils.py:52-53.

    seq_rel: (N, 2, T) float32.  Returns V (T, N, 2) f32 and A (T, N, N) f32.
    (`seq_` -- absolute positions -- only supplies shapes in the reference.)
    """
    rel = np.ascontiguousarray(np.asarray(seq_rel, dtype=np.float32))
    if rel.ndim != 3 or rel.shape[1] != 2:
        raise ValueError("seq_rel must be (N, 2, T)")
    N, _, T = rel.shape
    V = np.ascontiguousarray(rel.transpose(2, 0, 1))            # (T, N, 2)  utils.py:42
    x = V[:, :, 0]
    y = V[:, :, 1]
    dx = x[:, :, None] - x[:, None, :]                           # fp32
    dy = y[:, :, None] - y[:, None, :]
    s = dx * dx + dy * dy                                        # fp32 mul, mul, add
    assert s.dtype == np.float32
    nrm = np.sqrt(s.astype(np.float64))                          # math.sqrt utils.py:24
    with np.errstate(divide="ignore"):
        W = np.where(nrm == 0.0, 0.0, 1.0 / nrm)                 # utils.py:25-27
    idx = np.arange(N)
    W[:, idx, idx] = 1.0                                         # utils.py:43
    if not norm_lap_matr:
        return V, W.astype(np.float32)
    d = np.add.accumulate(W, axis=2)[:, :, -1]                   # sequential fp64 row sum
    ds = 1.0 / np.sqrt(d)
    L = -W
    L[:, idx, idx] = d - 1.0                                     # (D - W)_hh
    A = ds[:, :, None] * (L * ds[:, None, :])
    A[W == 0.0] = 0.0                                            # absent edges are +0.0
    return V, A.astype(np.float32)


# --------------------------------------------------------------------------
# parameter helpers
# --------------------------------------------------------------------------
def flatten_params(state: Dict[str, torch.Tensor]) -> np.ndarray:
    """state_dict -> flat fp32 block in named_parameters() order (7,563)."""
    parts = [np.asarray(state[k].detach().cpu().numpy(), np.float32).reshape(-1) for k, _ in PARAM_LAYOUT]
    flat = np.concatenate(parts)
    assert flat.size == N_PARAMS
    return flat


def flatten_bn(state: Dict[str, torch.Tensor]) -> np.ndarray:
    """state_dict -> 30 floats: (rm, rv) of BN1, BN2, BN_res."""
    return np.concatenate([np.asarray(state[k].detach().cpu().numpy(), np.float32).reshape(-1) for k in BN_LAYOUT])


def unflatten_params(flat, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    flat = np.asarray(flat)
    for k, shp in PARAM_LAYOUT:
        n = int(np.prod(shp))
        out[k] = torch.tensor(flat[off:off + n].reshape(shp), dtype=dtype)
        off += n
    return out


def unflatten_bn(flat, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    flat = np.asarray(flat)
    return {k: torch.tensor(flat[5 * i:5 * i + 5], dtype=dtype) for i, k in enumerate(BN_LAYOUT)}


# --------------------------------------------------------------------------
# (2)+(3) model forward                                     model.py:18-198
# --------------------------------------------------------------------------
def _bn(x, P, S, prefix, training, stats_out):
    """BatchNorm2d (model.py:114,121,141): eps 1e-5, momentum 0.1, affine."""
    rm = S[prefix + ".running_mean"]
    rv = S[prefix + ".running_var"]
    if training:
        # batch statistics over (N, H, W) of THIS scene; running stats updated
        # with the unbiased variance (torch.nn.BatchNorm2d semantics).
        rm_new, rv_new = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm_new, rv_new, P[prefix + ".weight"], P[prefix + ".bias"],
                         True, BN_MOMENTUM, BN_EPS)
        if stats_out is not None:
            stats_out[prefix + ".running_mean"] = rm_new
            stats_out[prefix + ".running_var"] = rv_new
        return y
    return F.batch_norm(x, rm, rv, P[prefix + ".weight"], P[prefix + ".bias"], False, BN_MOMENTUM, BN_EPS)


def stgcn_forward(P, S, v, A, training=False, stats_out=None):
    """st_gcn block (model.py:145-155) incl. ConvTemporalGraphical (model.py:64-68).
    v: (1, 2, T_obs, N); A: (T_obs, N, N).  Returns h (1, 5, T_obs, N)."""
    assert A.shape[0] == T_OBS                                              # model.py:65
    p = "st_gcns.0."
    res = F.conv2d(v, P[p + "residual.0.weight"], P[p + "residual.0.bias"])  # model.py:134-139
    res = _bn(res, P, S, p + "residual.1", training, stats_out)             # model.py:141
    g = F.conv2d(v, P[p + "gcn.conv.weight"], P[p + "gcn.conv.bias"])       # model.py:66
    y = torch.einsum("nctv,tvw->nctw", g, A).contiguous()                   # model.py:67
    y = _bn(y, P, S, p + "tcn.0", training, stats_out)                      # model.py:114
    y = F.prelu(y, P[p + "tcn.1.weight"])                                   # model.py:115
    y = F.conv2d(y, P[p + "tcn.2.weight"], P[p + "tcn.2.bias"], padding=(1, 0))  # model.py:116-122
    y = _bn(y, P, S, p + "tcn.3", training, stats_out)                      # model.py:123 (Dropout p=0)
    return F.prelu(y + res, P[p + "prelu.weight"])                          # model.py:150-153


def txpcnn_forward(P, h):
    """TXP-CNN (model.py:187-195).  h: (1, 5, 8, N) contiguous -> (1, 5, 12, N).
    Both `view`s are flat reinterpretations, not transposes (model.py:187,195)."""
    n = h.shape[3]
    x = h.contiguous().view(1, T_OBS, C_OUT, n)                             # model.py:187
    x = F.prelu(F.conv2d(x, P["tpcnns.0.weight"], P["tpcnns.0.bias"], padding=1), P["prelus.0.weight"])  # :189
    for k in (1, 2, 3):                                                     # range(1, n_txpcnn-1)  :191
        x = F.prelu(F.conv2d(x, P[f"tpcnns.{k}.weight"], P[f"tpcnns.{k}.bias"], padding=1),
                    P[f"prelus.{k}.weight"]) + x                            # :192
    x = F.conv2d(x, P["tpcnn_ouput.weight"], P["tpcnn_ouput.bias"], padding=1)  # :194
    return x.contiguous().view(1, C_OUT, T_PRED, n)                         # :195


def model_forward(P, S, V_obs, A_obs, training=False, stats_out=None):
    """social_stgcnn.forward (model.py:182-198) with the caller glue of
    train.py:48-56 / test.py:38-53.  V_obs: (T_obs, N, 2); A_obs: (T_obs, N, N).
    Returns V_pred in the (1, 5, T_pred, N) reference layout."""
    dt = P["tpcnns.0.weight"].dtype
    v = torch.as_tensor(V_obs, dtype=dt)[None].permute(0, 3, 1, 2)          # train.py:48
    a = torch.as_tensor(A_obs, dtype=dt)
    h = stgcn_forward(P, S, v, a, training, stats_out)
    return txpcnn_forward(P, h)


def vpred_tn5(out):
    """(1, 5, T_pred, N) -> (T_pred, N, 5) as train.py:52-56 does (no squeeze of N)."""
    return out.permute(0, 2, 3, 1)[0]


# --------------------------------------------------------------------------
# (4a) bivariate Gaussian NLL                               metrics.py:84-113
# --------------------------------------------------------------------------
def bivariate_loss(V_pred, V_trgt):
    """V_pred (T, N, 5), V_trgt (T, N, 2) -> 0-dim tensor (metrics.py:84-113)."""
    normx = V_trgt[:, :, 0] - V_pred[:, :, 0]
    normy = V_trgt[:, :, 1] - V_pred[:, :, 1]
    sx = torch.exp(V_pred[:, :, 2])
    sy = torch.exp(V_pred[:, :, 3])
    corr = torch.tanh(V_pred[:, :, 4])
    sxsy = sx * sy
    z = (normx / sx) ** 2 + (normy / sy) ** 2 - 2 * ((corr * normx * normy) / sxsy)
    neg_rho = 1 - corr ** 2
    result = torch.exp(-z / (2 * neg_rho))
    denom = 2 * np.pi * (sxsy * torch.sqrt(neg_rho))
    result = result / denom
    result = -torch.log(torch.clamp(result, min=1e-20))
    return torch.mean(result)


def train_step(P, S, scenes, batch_size):
    """One optimiser group of train() (train.py:28-79): every scene is forwarded
    in train mode (BN batch statistics per scene, running stats updated in
    order); loss = (sum of the first len-1 per-scene means) / batch_size -- the
    last scene of the group is forwarded but not scored (train.py:58-66).

    scenes: list of (V_obs (8,N,2), A_obs (8,N,N), V_tr (12,N,2)).
    Returns (loss float, flat grad (7563,) fp32 with zeros for never-used
    parameters, updated BN state dict, per-scene losses)."""
    P = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    S = {k: v.clone() for k, v in S.items()}
    total, per_scene = None, []
    for i, (V, A, Vt) in enumerate(scenes):
        upd = {}
        out = model_forward(P, S, V, A, training=True, stats_out=upd)
        S.update(upd)
        if i == len(scenes) - 1:
            break
        l = bivariate_loss(vpred_tn5(out), torch.as_tensor(Vt, dtype=out.dtype))
        per_scene.append(float(l.detach()))
        total = l if total is None else total + l
    loss = total / batch_size
    loss.backward()
    grad = np.concatenate([
        (P[k].grad.numpy().reshape(-1) if P[k].grad is not None else np.zeros(int(np.prod(shp))))
        for k, shp in PARAM_LAYOUT]).astype(np.float32 if loss.dtype == torch.float32 else np.float64)
    return float(loss), grad, S, per_scene


# --------------------------------------------------------------------------
# (4b) sampling + ADE/FDE                     test.py:59-126, metrics.py:21-75
# --------------------------------------------------------------------------
def cholesky_2x2(V_pred):
    """test.py:59-70: sx, sy, corr, cov and the lower Cholesky factor that
    MultivariateNormal builds, in closed form with fp32 roundings
    (potrf on [[a, b], [b, c]]: L00 = sqrt(a), L10 = b / L00,
    L11 = sqrt(fma(-L10, L10, c))).
    Pin status: against torch 2.11 CPU (MKL 2024.2 spotrf) L00 is bit-equal and
    L11 is bit-equal GIVEN MKL's L10, but MKL's L10 is not a correctly rounded
    quotient (it is within +-2 ulp of b / L00 on 37 % of entries and matches no
    closed form we tried), so sampled positions agree with the reference to
    ~1e-6 relative, not bit-for-bit ("Cholesky" hard part).
    V_pred: (T, N, 5) float32 numpy.  Returns mean (T,N,2), L00, L10, L11 (T,N)."""
    V = np.asarray(V_pred, np.float32)
    t = torch.from_numpy(np.ascontiguousarray(V))
    sx = torch.exp(t[:, :, 2]).numpy()
    sy = torch.exp(t[:, :, 3]).numpy()
    corr = torch.tanh(t[:, :, 4]).numpy()
    a = sx * sx
    b = corr * sx * sy                   # (corr*sx)*sy  test.py:65
    c = sy * sy
    L00 = np.sqrt(a)
    L10 = b / L00
    L11 = np.sqrt((c.astype(np.float64) - L10.astype(np.float64) * L10.astype(np.float64)).astype(np.float32))
    return V[:, :, 0:2], L00, L10, L11


def rel_to_abs(nodes_rel, init):
    """nodes_rel_to_nodes_abs (metrics.py:69-75): sequential fp32 prefix sum,
    fp64 add of the start node, store to fp32."""
    nodes_rel = np.asarray(nodes_rel, np.float32)
    cs = np.add.accumulate(nodes_rel, axis=0, dtype=np.float32)
    return (cs.astype(np.float64) + np.asarray(init, np.float64)[None]).astype(np.float32)


def sample_ade_fde(V_pred, V_tr, last_obs, eps):
    """test.py:59-123 for one scene with INJECTED standard-normal noise.

    V_pred (12, N, 5) f32; V_tr (12, N, 2) f32 (relative targets);
    last_obs (N, 2) f32 = obs_traj[:, :, -1]; eps (K, 12, N, 2) f32.
    Sample k = mean + L @ eps[k]  (MultivariateNormal.rsample).
    Returns (min_ade (N,), min_fde (N,)) float64, and the K absolute
    predicted trajectories (K, 12, N, 2) f32 (raw_data_dict['pred'], test.py:105)."""
    mean, L00, L10, L11 = cholesky_2x2(V_pred)
    eps = np.asarray(eps, np.float32)
    K = eps.shape[0]
    init = np.asarray(last_obs, np.float32)
    tgt_abs = rel_to_abs(V_tr, init)                                       # test.py:84-85
    N = mean.shape[1]
    ade = np.empty((K, N), np.float64)
    fde = np.empty((K, N), np.float64)
    preds = np.empty((K,) + mean.shape, np.float32)
    for k in range(K):
        e0, e1 = eps[k, :, :, 0], eps[k, :, :, 1]
        sx_ = mean[:, :, 0] + L00 * e0
        sy_ = mean[:, :, 1] + (L10 * e0 + L11 * e1)
        pred_abs = rel_to_abs(np.stack([sx_, sy_], -1), init)              # test.py:103-104
        preds[k] = pred_abs
        dx = pred_abs[:, :, 0] - tgt_abs[:, :, 0]                           # fp32, metrics.py:33
        dy = pred_abs[:, :, 1] - tgt_abs[:, :, 1]
        d = np.sqrt((dx * dx + dy * dy).astype(np.float64))                # math.sqrt -> fp64
        acc = np.zeros(N, np.float64)
        for t in range(d.shape[0]):                                        # sequential += over t
            acc = acc + d[t]
        ade[k] = acc / d.shape[0]                                          # metrics.py:34
        fde[k] = d[-1]                                                      # metrics.py:50-51
    return ade.min(0), fde.min(0), preds                                   # test.py:121-123


# --------------------------------------------------------------------------
# whole-scene inference (what bench.py's CPU legs time)
# --------------------------------------------------------------------------
def infer_scene(P, S, obs_rel, pred_rel, last_obs, eps):
    """graph build + eval forward + K-sample ADE/FDE for ONE scene, batch 1,
    exactly the per-scene work of test.py (dataset graph build included)."""
    V_obs, A_obs = seq_to_graph(obs_rel)
    V_tr = np.ascontiguousarray(np.asarray(pred_rel, np.float32).transpose(2, 0, 1))
    with torch.no_grad():
        out = model_forward(P, S, V_obs, A_obs)
    V_pred = vpred_tn5(out).numpy()
    return sample_ade_fde(V_pred, V_tr, last_obs, eps)[:2]


def synthetic_scenes(num_scenes, n_peds, seed=1234):
    """Synthetic crowd generator of SURVEY 8(d) (mirrors the random-walk data of
    complete_nuscenes_setup.py:259-286 and the rel/rounding rules of
    utils.py:145,153-155).  Returns obs_rel (S*N,2,8), pred_rel (S*N,2,12),
    obs_abs (S*N,2,8) float32 -- packed like TrajectoryDataset.obs_traj_rel."""
    g = torch.Generator().manual_seed(seed)
    P_ = num_scenes * n_peds
    start = (torch.rand(P_, 2, generator=g) * 20 - 10)
    vel = (torch.rand(P_, 2, generator=g) - 0.5)
    steps = vel[:, :, None] + 0.1 * torch.randn(P_, 2, 20, generator=g)
    steps[:, :, 0] = 0
    pos = start[:, :, None] + torch.cumsum(steps, dim=2)
    pos = np.around(pos.double().numpy(), decimals=4)                      # utils.py:145
    rel = np.zeros_like(pos)
    rel[:, :, 1:] = pos[:, :, 1:] - pos[:, :, :-1]                          # utils.py:153-155
    pos32, rel32 = pos.astype(np.float32), rel.astype(np.float32)
    return (np.ascontiguousarray(rel32[:, :, :T_OBS]), np.ascontiguousarray(rel32[:, :, T_OBS:]),
            np.ascontiguousarray(pos32[:, :, :T_OBS]))
