"""Packed-batch host layer over the C ABI: plans, device buffers, and the
inference pipeline  graph build -> st_gcn -> TXP-CNN -> K-sample ADE/FDE.

torch is used for device memory, streams and (in `parallel.py`) torch.distributed
only; every arithmetic step is a libsstg.so kernel.  The layouts are the
reference's own: pedestrians packed as `TrajectoryDataset.obs_traj_rel` packs them
(utils.py:180-190), per-scene V (T,N,2), A (T,N,N), V_pred (5,12,N).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib as L


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(dev: torch.device):
    return torch.cuda.current_stream(dev).cuda_stream


def _need(t: torch.Tensor, dtype, what: str, dev=None):
    if not isinstance(t, torch.Tensor) or t.dtype != dtype:
        raise TypeError(f"{what} must be a {dtype} tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device (no CPU fallback)")
    if dev is not None and t.device != dev:
        raise RuntimeError(f"{what} is on {t.device}, expected {dev}")
    if not t.is_contiguous():
        raise RuntimeError(f"{what} must be contiguous")
    return t


class ScenePlan:
    """Host-side description of a ragged batch of scenes (sstg_plan_host)."""

    def __init__(self, n_peds, T: int = L.T_OBS):
        n = np.ascontiguousarray(np.asarray(n_peds, dtype=np.int32).reshape(-1))
        S = int(n.size)
        self.S, self.T = S, int(T)
        self.n_peds = n
        self.seg_off = np.zeros(S + 1, np.int32)
        self.a_off = np.zeros(S + 1, np.int64)
        self.tile_off = np.zeros(S + 1, np.int32)
        nt, mx, af, tc = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int64), np.zeros(1, np.int32)
        lib = L.load()
        L.check(lib.sstg_plan_host(n.ctypes.data, S, self.T, self.seg_off.ctypes.data, self.a_off.ctypes.data,
                                   self.tile_off.ctypes.data, nt.ctypes.data, mx.ctypes.data, af.ctypes.data,
                                   tc.ctypes.data), "sstg_plan_host")
        self.n_tiles, self.max_n, self.a_floats, self.tile_cols = int(nt[0]), int(mx[0]), int(af[0]), int(tc[0])
        self.tile_off = self.tile_off[: self.n_tiles + 1].copy()
        self.P = int(self.seg_off[-1])
        self.uniform = bool(S > 0 and int(n.min()) == int(n.max()))
        self._dev: Dict[torch.device, "DevicePlan"] = {}
        self._cache: Dict[tuple, list] = {}

    def chunks(self, chunk_scenes: int, taper: bool = False):
        """Contiguous sub-plans of at most `chunk_scenes` scenes: [(scene lo, hi, ped lo, hi, ScenePlan)].
        Identical pedestrian-count patterns share one sub-plan (and its device offsets).
        `taper`: a quarter-size first and last chunk -- in a copy/compute pipeline the first chunk's copy and
        the last chunk's kernels are the parts nothing overlaps, so they are kept short."""
        chunk_scenes = max(1, int(chunk_scenes))
        key = ("chunks", chunk_scenes, bool(taper))
        got = self._cache.get(key)
        if got is None:
            q = chunk_scenes // 4
            if taper and q >= 1 and self.S > 2 * chunk_scenes:
                mid = self.S - 2 * q
                cuts = [0, q] + [q + k * chunk_scenes for k in range(1, mid // chunk_scenes + 1)]
                if mid % chunk_scenes:
                    cuts.append(self.S - q)
                cuts.append(self.S)
            else:
                cuts = list(range(0, self.S, chunk_scenes)) + [self.S]
            got, uniq = [], {}
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                if hi <= lo:
                    continue
                n = self.n_peds[lo:hi]
                sub = uniq.get(n.tobytes())
                if sub is None:
                    sub = uniq[n.tobytes()] = self if (lo == 0 and hi == self.S) else ScenePlan(n, self.T)
                got.append((lo, hi, int(self.seg_off[lo]), int(self.seg_off[hi]), sub))
            self._cache[key] = got
        return got

    def on(self, device) -> "DevicePlan":
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ScenePlan.on(): a CUDA device is required (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device not in self._dev:
            # one host->device copy for the three offset tables: [a_off int64 | seg_off int32 | tile_off int32]
            nb_a, nb_s, nb_t = self.a_off.nbytes, self.seg_off.nbytes, self.tile_off.nbytes
            host = np.empty(nb_a + nb_s + nb_t, np.uint8)
            host[:nb_a] = self.a_off.view(np.uint8)
            host[nb_a:nb_a + nb_s] = self.seg_off.view(np.uint8)
            host[nb_a + nb_s:] = self.tile_off.view(np.uint8)
            buf = torch.from_numpy(host).to(device)
            self._dev[device] = DevicePlan(
                self, device,
                buf[nb_a:nb_a + nb_s].view(torch.int32),
                buf[:nb_a].view(torch.int64),
                buf[nb_a + nb_s:].view(torch.int32))
        return self._dev[device]


@dataclass
class DevicePlan:
    host: ScenePlan
    device: torch.device
    seg_off: torch.Tensor
    a_off: torch.Tensor
    tile_off: torch.Tensor


def plan_groups(n_peds_list, device, T: int = L.T_OBS):
    """ScenePlans of many batches (e.g. all optimiser groups of an epoch) whose offset tables reach the
    device in ONE host->device copy instead of one per plan.  Returns the list of ScenePlans with
    `.on(device)` already resolved."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("plan_groups(): a CUDA device is required (no CPU fallback)")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    plans = [ScenePlan(n, T) for n in n_peds_list]
    up16 = lambda x: (x + 15) & ~15
    offs, total = [], 0
    for p in plans:
        o_a = total
        o_s = o_a + up16(p.a_off.nbytes)
        o_t = o_s + up16(p.seg_off.nbytes)
        total = o_t + up16(p.tile_off.nbytes)
        offs.append((o_a, o_s, o_t))
    host = np.zeros(max(total, 16), np.uint8)
    for p, (o_a, o_s, o_t) in zip(plans, offs):
        host[o_a:o_a + p.a_off.nbytes] = p.a_off.view(np.uint8)
        host[o_s:o_s + p.seg_off.nbytes] = p.seg_off.view(np.uint8)
        host[o_t:o_t + p.tile_off.nbytes] = p.tile_off.view(np.uint8)
    buf = torch.from_numpy(host).to(device)
    for p, (o_a, o_s, o_t) in zip(plans, offs):
        p._dev[device] = DevicePlan(p, device,
                                    buf[o_s:o_s + p.seg_off.nbytes].view(torch.int32),
                                    buf[o_a:o_a + p.a_off.nbytes].view(torch.int64),
                                    buf[o_t:o_t + p.tile_off.nbytes].view(torch.int32))
    return plans


# --------------------------------------------------------------------------
# functional ops (device tensors in, device tensors out; `out=` to reuse buffers)
# --------------------------------------------------------------------------
def graph_build(dp: DevicePlan, rel: torch.Tensor, T: int, *, norm_lap: bool = True, fp64: bool = False,
                v_only: bool = False, V: Optional[torch.Tensor] = None, A: Optional[torch.Tensor] = None):
    """utils.seq_to_graph (utils.py:29-53) for every scene of the packed batch.
    rel: [P, 2, T] fp32.  Returns (V packed (T,N,2) blocks, A packed (T,N,N) blocks or None)."""
    h, dev = dp.host, dp.device
    _need(rel, torch.float32, "rel", dev)
    if tuple(rel.shape) != (h.P, 2, T):
        raise ValueError(f"rel must be [{h.P}, 2, {T}], got {tuple(rel.shape)}")
    if not v_only and T != h.T:
        raise ValueError(f"plan was made for T={h.T}; build a ScenePlan(n_peds, T={T}) for this adjacency")
    if V is None:
        V = torch.empty(h.P * T * 2, dtype=torch.float32, device=dev)
    if A is None and not v_only:
        A = torch.empty(h.a_floats, dtype=torch.float32, device=dev)
    flags = (L.GRAPH_NORM_LAP if norm_lap else 0) | (L.GRAPH_FP64 if fp64 else 0) | (L.GRAPH_V_ONLY if v_only else 0)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_graph_build(_ptr(rel), _ptr(dp.seg_off), _ptr(dp.a_off), h.S, T, h.max_n,
                                          _ptr(V), _ptr(A), flags, _stream(dev)), "sstg_graph_build")
    return V, (None if v_only else A)


def stgcn_forward(dp: DevicePlan, V, A, params, bn, *, H=None, sm_count: int = 0):
    """st_gcn.forward (model.py:145-155), eval-mode BN, for the packed batch -> H (5,8,N) blocks."""
    h, dev = dp.host, dp.device
    for t, w in ((V, "V"), (A, "A"), (params, "params"), (bn, "bn")):
        _need(t, torch.float32, w, dev)
    if params.numel() != L.N_PARAMS or bn.numel() != L.N_BN:
        raise ValueError("params must hold 7563 floats and bn 30 floats")
    if V.numel() < h.P * L.T_OBS * 2 or A.numel() < h.a_floats:
        raise ValueError("V / A buffers are smaller than the plan requires")
    if h.T != L.T_OBS:
        raise AssertionError(f"A.size(0) must equal the st_gcn kernel size {L.T_OBS}")  # model.py:65
    if H is None:
        H = torch.empty(h.P * 40, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_stgcn_forward(_ptr(V), _ptr(A), _ptr(dp.seg_off), _ptr(dp.a_off), h.S, h.max_n,
                                            _ptr(params), _ptr(bn), _ptr(H), sm_count, _stream(dev)),
                "sstg_stgcn_forward")
    return H


def stgcn_forward_fused(dp: DevicePlan, obs_rel, params, bn, *, H=None, Pws=None):
    """seq_to_graph + st_gcn.forward (eval) straight from obs_traj_rel [P,2,8] with the adjacency kept on
    chip (sstg_stgcn_forward_fused; N <= 64) -> H (5,8,N) blocks."""
    h, dev = dp.host, dp.device
    for t, w in ((obs_rel, "obs_rel"), (params, "params"), (bn, "bn")):
        _need(t, torch.float32, w, dev)
    if tuple(obs_rel.shape) != (h.P, 2, L.T_OBS):
        raise ValueError(f"obs_rel must be [{h.P}, 2, {L.T_OBS}]")
    if h.max_n > 64:
        raise ValueError("the fused graph + st_gcn path handles scenes of up to 64 pedestrians")
    if H is None:
        H = torch.empty(h.P * 40, dtype=torch.float32, device=dev)
    if Pws is None:
        Pws = torch.empty(h.P * 40, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_stgcn_forward_fused(_ptr(obs_rel), _ptr(dp.seg_off), h.S, h.max_n, _ptr(params),
                                                  _ptr(bn), _ptr(Pws), _ptr(H), _stream(dev)),
                "sstg_stgcn_forward_fused")
    return H


def txpcnn_forward(dp: DevicePlan, H, params, *, out=None):
    """model.py:187-195 for the packed batch -> V_pred (5,12,N) blocks."""
    h, dev = dp.host, dp.device
    _need(H, torch.float32, "H", dev)
    _need(params, torch.float32, "params", dev)
    if out is None:
        out = torch.empty(h.P * 60, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_txpcnn_forward(_ptr(H), _ptr(dp.seg_off), _ptr(dp.tile_off), h.n_tiles, h.max_n,
                                             h.tile_cols, _ptr(params), _ptr(out), _stream(dev)),
                "sstg_txpcnn_forward")
    return out


def nll_forward(dp: DevicePlan, Vpred, Vtr, *, loss=None):
    """metrics.bivariate_loss (metrics.py:84-113) per scene -> loss[S]."""
    h, dev = dp.host, dp.device
    _need(Vpred, torch.float32, "Vpred", dev)
    _need(Vtr, torch.float32, "Vtr", dev)
    if loss is None:
        loss = torch.empty(h.S, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_nll_forward(_ptr(Vpred), _ptr(Vtr), _ptr(dp.seg_off), h.S, _ptr(loss), _stream(dev)),
                "sstg_nll_forward")
    return loss


def sample_ade_fde(dp: DevicePlan, Vpred, Vtr, last_obs, K: int = 20, *, eps=None, seed: int = 0,
                   want_eps: bool = False, want_pred: bool = False, fp64_dist: bool = True,
                   min_ade=None, min_fde=None, vtr_is_rel: bool = False):
    """test.py:59-123 for the packed batch.  Returns dict(min_ade[P] f64, min_fde[P] f64,
    eps (noise used, if want_eps), pred_abs (K,12,N,2 blocks, if want_pred)).
    vtr_is_rel: `Vtr` is pred_traj_rel [P,2,12] itself (saves building the V_tr blocks)."""
    h, dev = dp.host, dp.device
    for t, w in ((Vpred, "Vpred"), (Vtr, "Vtr"), (last_obs, "last_obs")):
        _need(t, torch.float32, w, dev)
    if last_obs.numel() != h.P * 2:
        raise ValueError("last_obs must be [P, 2]")
    n_eps = h.P * K * L.T_PRED * 2
    if eps is not None:
        _need(eps, torch.float32, "eps", dev)
        if eps.numel() != n_eps:
            raise ValueError(f"eps must hold {n_eps} floats ((K,12,N,2) per scene)")
    if min_ade is None:
        min_ade = torch.empty(h.P, dtype=torch.float64, device=dev)
    if min_fde is None:
        min_fde = torch.empty(h.P, dtype=torch.float64, device=dev)
    eps_out = torch.empty(n_eps, dtype=torch.float32, device=dev) if want_eps else None
    pred = torch.empty(n_eps, dtype=torch.float32, device=dev) if want_pred else None
    with torch.cuda.device(dev):
        L.check(L.load().sstg_sample_ade_fde(_ptr(Vpred), _ptr(Vtr), _ptr(last_obs), _ptr(dp.seg_off), h.S, h.P, K,
                                             _ptr(eps), seed & (2 ** 64 - 1), _ptr(eps_out), _ptr(min_ade),
                                             _ptr(min_fde), _ptr(pred),
                                             (L.EVAL_FP64_DIST if fp64_dist else 0) |
                                             (L.EVAL_VTR_IS_REL if vtr_is_rel else 0) |
                                             (L.EVAL_UNIFORM_N if h.uniform else 0),
                                             _stream(dev)), "sstg_sample_ade_fde")
    return {"min_ade": min_ade, "min_fde": min_fde, "eps": eps_out, "pred_abs": pred}


def pack_tracks(traj):
    """Host side of the compact input format (include/sstg.h: sstg_decode_tracks): positions
    [P, 2, 20] (any float dtype; rounded to 4 decimals like utils.py:145) -> (start int32 [P,2],
    delta int16 [P,2,20]) in units of 1e-4.  Raises OverflowError if a per-frame step exceeds
    +-3.2767 m or a start position +-214 km (then the fp32 interface must be used)."""
    k = np.rint(np.asarray(traj, np.float64) * 1e4).astype(np.int64)
    if k.ndim != 3 or k.shape[1] != 2 or k.shape[2] != L.T_OBS + L.T_PRED:
        raise ValueError("traj must be [P, 2, 20]")
    d = np.zeros(k.shape, np.int64)
    d[:, :, 1:] = k[:, :, 1:] - k[:, :, :-1]
    if np.abs(d).max(initial=0) > 32767 or np.abs(k[:, :, 0]).max(initial=0) > 2 ** 31 - 1:
        raise OverflowError("track does not fit the int32 start / int16 step format")
    return np.ascontiguousarray(k[:, :, 0].astype(np.int32)), np.ascontiguousarray(d.astype(np.int16))


def decode_tracks(start, delta, *, obs_rel=None, pred_rel=None, last_obs=None):
    """sstg_decode_tracks: compact (start, delta) on the device -> obs_traj_rel [P,2,8],
    pred_traj_rel [P,2,12], obs_traj[:, :, -1] [P,2], bit-identical to TrajectoryDataset's tensors."""
    dev = start.device
    _need(start, torch.int32, "start", dev)
    _need(delta, torch.int16, "delta", dev)
    P = start.shape[0]
    if tuple(start.shape) != (P, 2) or tuple(delta.shape) != (P, 2, L.T_OBS + L.T_PRED):
        raise ValueError("start must be [P,2] int32 and delta [P,2,20] int16")
    f = torch.float32
    obs_rel = torch.empty(P, 2, L.T_OBS, dtype=f, device=dev) if obs_rel is None else obs_rel
    pred_rel = torch.empty(P, 2, L.T_PRED, dtype=f, device=dev) if pred_rel is None else pred_rel
    last_obs = torch.empty(P, 2, dtype=f, device=dev) if last_obs is None else last_obs
    with torch.cuda.device(dev):
        L.check(L.load().sstg_decode_tracks(_ptr(start), _ptr(delta), P, _ptr(obs_rel), _ptr(pred_rel),
                                            _ptr(last_obs), _stream(dev)), "sstg_decode_tracks")
    return obs_rel, pred_rel, last_obs


# --------------------------------------------------------------------------
# the inference pipeline
# --------------------------------------------------------------------------
class InferenceEngine:
    """Weights resident on one GPU + reusable workspaces; runs the whole hot path
    on a packed batch: seq_to_graph -> social_stgcnn.forward (eval) -> test()'s
    K-sample ADE/FDE.  `graph_fp64` / `fp64_dist` select the reference's fp64
    rounding islands (bit-level parity) or the fp32 fast path (within 1e-5)."""

    def __init__(self, params, bn, device="cuda"):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("InferenceEngine needs a CUDA device (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        L.load()
        self.device = device
        self.params = torch.as_tensor(np.asarray(params, np.float32)).reshape(-1).to(device).contiguous()
        self.bn = torch.as_tensor(np.asarray(bn, np.float32)).reshape(-1).to(device).contiguous()
        if self.params.numel() != L.N_PARAMS or self.bn.numel() != L.N_BN:
            raise ValueError("params must hold 7563 floats and bn 30 floats")
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self._cap: Dict[str, Optional[torch.Tensor]] = {}
        self._pin: Dict[tuple, Dict[str, torch.Tensor]] = {}

    @classmethod
    def from_state_dict(cls, sd, device="cuda"):
        params, bn, _ = flatten_state_dict(sd)
        return cls(params, bn, device)

    def load_weights(self, params, bn):
        """Replace the resident weights (e.g. with a TrainEngine's current ones) without reallocating."""
        self.params.copy_(torch.as_tensor(params).reshape(-1))
        self.bn.copy_(torch.as_tensor(bn).reshape(-1))

    def workspace(self, plan: ScenePlan) -> Dict[str, torch.Tensor]:
        """Views, sized for `plan`, of one grow-only set of device buffers (so a smaller plan --
        e.g. the chunks of infer_host -- reuses the memory of a larger one)."""
        need = {"V_obs": (plan.P * L.T_OBS * 2, torch.float32), "A_obs": (plan.a_floats, torch.float32),
                "V_tr": (plan.P * L.T_PRED * 2, torch.float32), "H": (plan.P * 40, torch.float32),
                "Pws": (plan.P * 40, torch.float32),
                "V_pred": (plan.P * 60, torch.float32), "min_ade": (plan.P, torch.float64),
                "min_fde": (plan.P, torch.float64)}
        key = (plan.S, plan.P, plan.a_floats)
        ws = self._ws.get(key)
        if ws is None:
            for k, (n, dt) in need.items():
                cur = self._cap.get(k)
                if cur is None or cur.numel() < n:
                    self._cap[k] = None          # drop the old buffer before allocating the bigger one
                    self._cap[k] = torch.empty(max(n, 1), dtype=dt, device=self.device)
                    self._ws = {}                # views of the old buffer are stale
            ws = {k: self._cap[k][:n] for k, (n, dt) in need.items()}
            if len(self._ws) > 64:
                self._ws = {}
            self._ws[key] = ws
        return ws

    def infer(self, plan: ScenePlan, obs_rel, pred_rel, last_obs, K: int = 20, *, eps=None, seed: int = 0,
              graph_fp64: bool = False, fp64_dist: bool = False, want_eps: bool = False, want_pred: bool = False,
              min_ade=None, min_fde=None, want_vtr: bool = True, fused: bool = False):
        """Device-resident inputs: obs_rel [P,2,8], pred_rel [P,2,12], last_obs [P,2].
        Returns the workspace dict (views valid until the next call) plus eps / pred_abs."""
        dp = plan.on(self.device)
        ws = self.workspace(plan)
        fused = bool(fused) and plan.max_n <= 64 and not graph_fp64
        if fused:   # graph + st_gcn with the adjacency kept on chip: A_obs / V_obs are NOT produced
            stgcn_forward_fused(dp, obs_rel, self.params, self.bn, H=ws["H"], Pws=ws["Pws"])
        else:
            graph_build(dp, obs_rel, L.T_OBS, fp64=graph_fp64, V=ws["V_obs"], A=ws["A_obs"])
            stgcn_forward(dp, ws["V_obs"], ws["A_obs"], self.params, self.bn, H=ws["H"])
        if want_vtr:  # the (12,N,2) target blocks (V_tr of the dataset tuple; needed by the NLL, not by ADE/FDE)
            graph_build(dp, pred_rel, L.T_PRED, v_only=True, V=ws["V_tr"])
        txpcnn_forward(dp, ws["H"], self.params, out=ws["V_pred"])
        r = sample_ade_fde(dp, ws["V_pred"], ws["V_tr"] if want_vtr else pred_rel, last_obs, K, eps=eps, seed=seed,
                           want_eps=want_eps, want_pred=want_pred, fp64_dist=fp64_dist, vtr_is_rel=not want_vtr,
                           min_ade=ws["min_ade"] if min_ade is None else min_ade,
                           min_fde=ws["min_fde"] if min_fde is None else min_fde)
        out = dict(ws)
        if fused:
            out["A_obs"] = out["V_obs"] = None
        out["min_ade"], out["min_fde"] = r["min_ade"], r["min_fde"]
        out["eps"], out["pred_abs"] = r["eps"], r["pred_abs"]
        return out

    def capture(self, plan: ScenePlan, K: int = 20, *, seed: int = 0, graph_fp64: bool = False,
                fp64_dist: bool = False, fused: bool = False, want_vtr: bool = True) -> "GraphedInference":
        """The whole pipeline of `infer` for a FIXED plan captured once into a CUDA graph: small batches (a test split
        of a few hundred scenes) are launch-bound -- five to six kernels of a few microseconds each -- and a graph
        replay issues them with one driver call.  Returns a GraphedInference whose static input tensors are filled by
        the caller and whose `replay()` returns the same dict as `infer`."""
        return GraphedInference(self, plan, K, seed=seed, graph_fp64=graph_fp64, fp64_dist=fp64_dist, fused=fused,
                                want_vtr=want_vtr)

    def infer_host(self, plan: ScenePlan, obs_rel, pred_rel, last_obs, K: int = 20, *, chunk_scenes: int = 8192,
                   seed: int = 0, compact=None, fused: bool = True, taper: bool = True, **kw):
        """End-to-end call with HOST buffers (numpy or CPU tensors, ideally pinned): the batch is
        cut into chunks of `chunk_scenes` scenes and software-pipelined over three streams --
        H2D of chunk i+1 (copy stream) overlaps the kernels of chunk i (current stream) and the
        D2H of chunk i-1's per-pedestrian minima (second copy stream).  PCIe carries 168 B per
        pedestrian in and 16 B out, so the steady state is max(H2D time, kernel time); the first and the
        last chunk are a quarter of the size (`taper`), which shortens the pipeline's fill and drain.
        Only the per-pedestrian minima leave the device, so by default (`fused`) graph build and
        st_gcn run as the fused pair that keeps the adjacency on chip (scenes of <= 64 pedestrians).
        `compact=(start, delta)` (pack_tracks) replaces the three fp32 host tensors by the 88 B per
        pedestrian fixed-point format, decoded bit-exactly on the device (sstg_decode_tracks).
        Returns (min_ade, min_fde) as numpy float64 arrays of length P (pinned, reused)."""
        dev = self.device
        if compact is None:
            obs_rel, pred_rel, last_obs = (torch.as_tensor(obs_rel), torch.as_tensor(pred_rel),
                                           torch.as_tensor(last_obs))
        else:
            c_start, c_delta = torch.as_tensor(compact[0]), torch.as_tensor(compact[1])
        chunks = plan.chunks(chunk_scenes, taper)
        key = (plan.P, max(c[4].P for c in chunks))
        st = self._pin.get(key)
        if st is None:
            mp = key[1]
            f = torch.float32
            st = {
                "in": [{"obs": torch.empty(mp, 2, L.T_OBS, dtype=f, device=dev),
                        "pred": torch.empty(mp, 2, L.T_PRED, dtype=f, device=dev),
                        "last": torch.empty(mp, 2, dtype=f, device=dev),
                        "start": torch.empty(mp, 2, dtype=torch.int32, device=dev),
                        "delta": torch.empty(mp, 2, L.T_OBS + L.T_PRED, dtype=torch.int16, device=dev)}
                       for _ in range(2)],
                "d_out": [torch.empty(plan.P, dtype=torch.float64, device=dev) for _ in range(2)],
                "h_out": [torch.empty(plan.P, dtype=torch.float64).pin_memory() for _ in range(2)],
                "s_h2d": torch.cuda.Stream(dev), "s_d2h": torch.cuda.Stream(dev),
                "free": [torch.cuda.Event() for _ in range(2)],
            }
            self._pin = {key: st}
        main = torch.cuda.current_stream(dev)
        s_h2d, s_d2h = st["s_h2d"], st["s_d2h"]
        s_h2d.wait_stream(main)          # earlier work on the caller's stream (e.g. a previous call's readers)
        s_d2h.wait_stream(main)
        for i, (lo, hi, p_lo, p_hi, sub) in enumerate(chunks):
            b = st["in"][i & 1]
            n = p_hi - p_lo
            with torch.cuda.stream(s_h2d):
                if i >= 2:
                    s_h2d.wait_event(st["free"][i & 1])      # the kernels of chunk i-2 are done with this buffer
                if compact is None:
                    b["obs"][:n].copy_(obs_rel[p_lo:p_hi], non_blocking=True)
                    b["pred"][:n].copy_(pred_rel[p_lo:p_hi], non_blocking=True)
                    b["last"][:n].copy_(last_obs[p_lo:p_hi], non_blocking=True)
                else:
                    b["start"][:n].copy_(c_start[p_lo:p_hi], non_blocking=True)
                    b["delta"][:n].copy_(c_delta[p_lo:p_hi], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(s_h2d)
            main.wait_event(ready)
            if compact is not None:
                decode_tracks(b["start"][:n], b["delta"][:n], obs_rel=b["obs"][:n], pred_rel=b["pred"][:n],
                              last_obs=b["last"][:n])
            self.infer(sub, b["obs"][:n], b["pred"][:n], b["last"][:n], K, seed=seed + 0x9E3779B1 * i, want_vtr=False,
                       fused=fused,
                       min_ade=st["d_out"][0][p_lo:p_hi], min_fde=st["d_out"][1][p_lo:p_hi], **kw)
            st["free"][i & 1].record(main)
            s_d2h.wait_event(st["free"][i & 1])
            with torch.cuda.stream(s_d2h):
                # contiguous 1-D slices: a strided D2H copy would go through a pageable staging buffer and block
                st["h_out"][0][p_lo:p_hi].copy_(st["d_out"][0][p_lo:p_hi], non_blocking=True)
                st["h_out"][1][p_lo:p_hi].copy_(st["d_out"][1][p_lo:p_hi], non_blocking=True)
        main.wait_stream(s_d2h)
        main.synchronize()
        return st["h_out"][0].numpy(), st["h_out"][1].numpy()

    @staticmethod
    def h2d_bytes(plan: ScenePlan, compact: bool = False) -> int:
        if compact:
            return plan.P * (2 * 4 + 2 * (L.T_OBS + L.T_PRED) * 2)
        return plan.P * (2 * L.T_OBS + 2 * L.T_PRED + 2) * 4

    @staticmethod
    def d2h_bytes(plan: ScenePlan) -> int:
        return plan.P * 2 * 8


class GraphedInference:
    """`InferenceEngine.infer` for one fixed ScenePlan as a CUDA graph (torch.cuda.CUDAGraph: the capture records the
    kernel launches libsstg.so issues on the capturing stream).  Inputs live in static device tensors
    (`obs_rel`, `pred_rel`, `last_obs`): copy new data into them, call `replay()`, read the results -- views of static
    buffers, valid until the next replay.  The engine's weights are read at replay time, so `load_weights` is seen.
    The Philox seed is baked into the graph (kernel arguments are captured by value)."""

    def __init__(self, engine: "InferenceEngine", plan: ScenePlan, K: int, **kw):
        dev = engine.device
        self.engine, self.plan, self.K = engine, plan, int(K)
        f = torch.float32
        self.obs_rel = torch.zeros(plan.P, 2, L.T_OBS, dtype=f, device=dev)
        self.pred_rel = torch.zeros(plan.P, 2, L.T_PRED, dtype=f, device=dev)
        self.last_obs = torch.zeros(plan.P, 2, dtype=f, device=dev)
        # a private engine view: its own workspace, the parent's weight tensors (shared storage)
        self._eng = InferenceEngine.__new__(InferenceEngine)
        self._eng.device, self._eng.params, self._eng.bn = dev, engine.params, engine.bn
        self._eng._ws, self._eng._cap, self._eng._pin = {}, {}, {}
        plan.on(dev)                                   # offset tables uploaded before the capture
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                  # warm-up outside the capture: module loading, workspace allocation
            self._eng.infer(plan, self.obs_rel, self.pred_rel, self.last_obs, self.K, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._eng.infer(plan, self.obs_rel, self.pred_rel, self.last_obs, self.K, **kw)

    def replay(self, obs_rel=None, pred_rel=None, last_obs=None):
        """Optionally copy new inputs (device or host tensors) into the static buffers, then launch the graph."""
        for dst, src in ((self.obs_rel, obs_rel), (self.pred_rel, pred_rel), (self.last_obs, last_obs)):
            if src is not None:
                dst.copy_(torch.as_tensor(src), non_blocking=True)
        self.graph.replay()
        return self.out


# --------------------------------------------------------------------------
# training path: functional ops
# --------------------------------------------------------------------------
def partial_rows(plan: ScenePlan):
    """Rows of the two partial-gradient buffers sstg_backward needs on the current device."""
    a, b = np.zeros(1, np.int32), np.zeros(1, np.int32)
    L.check(L.load().sstg_train_partial_rows(plan.S, plan.n_tiles, a.ctypes.data, b.ctypes.data),
            "sstg_train_partial_rows")
    return int(a[0]), int(b[0])


def forward_train(dp: DevicePlan, V, A, params, *, save=None, scene_stats=None, out=None):
    """social_stgcnn.forward in train mode (per-scene BatchNorm batch statistics) for the packed
    batch.  Returns (V_pred (5,12,N) blocks, save, scene_stats [S,30])."""
    h, dev = dp.host, dp.device
    for t, w in ((V, "V"), (A, "A"), (params, "params")):
        _need(t, torch.float32, w, dev)
    if h.T != L.T_OBS:
        raise AssertionError(f"A.size(0) must equal the st_gcn kernel size {L.T_OBS}")  # model.py:65
    if save is None:
        save = torch.empty(h.P * L.SAVE_FLOATS_PER_PED, dtype=torch.float32, device=dev)
    if scene_stats is None:
        scene_stats = torch.empty(h.S, L.N_BN, dtype=torch.float32, device=dev)
    if out is None:
        out = torch.empty(h.P * 60, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_forward_train(_ptr(V), _ptr(A), _ptr(dp.seg_off), _ptr(dp.a_off), _ptr(dp.tile_off),
                                            h.S, h.n_tiles, h.max_n, h.P, _ptr(params), _ptr(save),
                                            _ptr(scene_stats), _ptr(out), _stream(dev)), "sstg_forward_train")
    return out, save, scene_stats


def bn_running_update(dp: DevicePlan, scene_stats, bn, momentum: float = 0.1):
    """Apply the S per-scene BatchNorm updates to the running statistics bn[30], in scene order."""
    h, dev = dp.host, dp.device
    _need(scene_stats, torch.float32, "scene_stats", dev)
    _need(bn, torch.float32, "bn", dev)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_bn_running_update(_ptr(scene_stats), _ptr(dp.seg_off), h.S, float(momentum), _ptr(bn),
                                                _stream(dev)), "sstg_bn_running_update")
    return bn


def nll_backward(dp: DevicePlan, Vpred, Vtr, scene_weight=None, *, loss=None, dVpred=None):
    """bivariate_loss per scene and scene_weight[s] * d loss[s] / d V_pred."""
    h, dev = dp.host, dp.device
    _need(Vpred, torch.float32, "Vpred", dev)
    _need(Vtr, torch.float32, "Vtr", dev)
    if scene_weight is not None:
        _need(scene_weight, torch.float32, "scene_weight", dev)
        if scene_weight.numel() != h.S:
            raise ValueError("scene_weight must hold one float per scene")
    if loss is None:
        loss = torch.empty(h.S, dtype=torch.float32, device=dev)
    if dVpred is None:
        dVpred = torch.empty(h.P * 60, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_nll_backward(_ptr(Vpred), _ptr(Vtr), _ptr(dp.seg_off), h.S, _ptr(scene_weight),
                                           _ptr(loss), _ptr(dVpred), _stream(dev)), "sstg_nll_backward")
    return loss, dVpred


def backward(dp: DevicePlan, V, A, params, save, scene_stats, dVpred, *, grad=None, ws=None):
    """loss.backward() for the packed batch: flat gradient [7563] in named_parameters() order."""
    h, dev = dp.host, dp.device
    for t, w in ((V, "V"), (A, "A"), (params, "params"), (save, "save"), (scene_stats, "scene_stats"),
                 (dVpred, "dVpred")):
        _need(t, torch.float32, w, dev)
    if grad is None:
        grad = torch.empty(L.N_PARAMS, dtype=torch.float32, device=dev)
    if ws is None:
        ws = {}
    with torch.cuda.device(dev):
        rs, rt = partial_rows(h)
        f = torch.float32
        if ws.get("dH") is None or ws["dH"].numel() < h.P * 40:
            ws["dH"] = torch.empty(max(1, h.P * 40), dtype=f, device=dev)
        if ws.get("pst") is None or ws["pst"].numel() < rs * L.ST_PARAMS:
            ws["pst"] = torch.empty(max(1, rs * L.ST_PARAMS), dtype=f, device=dev)
        if ws.get("ptx") is None or ws["ptx"].numel() < rt * L.TXP_GRAD_STRIDE:
            ws["ptx"] = torch.empty(max(1, rt * L.TXP_GRAD_STRIDE), dtype=f, device=dev)
        L.check(L.load().sstg_backward(_ptr(V), _ptr(A), _ptr(dp.seg_off), _ptr(dp.a_off), _ptr(dp.tile_off), h.S,
                                       h.n_tiles, h.max_n, h.P, _ptr(params), _ptr(save), _ptr(scene_stats),
                                       _ptr(dVpred), _ptr(ws["dH"]), _ptr(ws["pst"]), _ptr(ws["ptx"]), _ptr(grad),
                                       _stream(dev)), "sstg_backward")
    return grad


def sgd_step(params, grad, lr: float, clip_norm: float = 0.0, grad_norm_out=None):
    """optimizer.step() of train.py:197 (+ clip_grad_norm_ of train.py:69-70 when clip_norm > 0)."""
    dev = params.device
    _need(params, torch.float32, "params", dev)
    _need(grad, torch.float32, "grad", dev)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_sgd_step(_ptr(params), _ptr(grad), params.numel(), float(lr), float(clip_norm or 0.0),
                                       _ptr(grad_norm_out), _stream(dev)), "sstg_sgd_step")
    return params


# named_parameters() order / shapes of the reference model (model.py:157-178) and its
# BatchNorm buffers: the layout of the flat parameter block and of bn[30].
PARAM_LAYOUT = (
    [("st_gcns.0.gcn.conv.weight", (5, 2, 1, 1)), ("st_gcns.0.gcn.conv.bias", (5,)),
     ("st_gcns.0.tcn.0.weight", (5,)), ("st_gcns.0.tcn.0.bias", (5,)), ("st_gcns.0.tcn.1.weight", (1,)),
     ("st_gcns.0.tcn.2.weight", (5, 5, 3, 1)), ("st_gcns.0.tcn.2.bias", (5,)),
     ("st_gcns.0.tcn.3.weight", (5,)), ("st_gcns.0.tcn.3.bias", (5,)),
     ("st_gcns.0.residual.0.weight", (5, 2, 1, 1)), ("st_gcns.0.residual.0.bias", (5,)),
     ("st_gcns.0.residual.1.weight", (5,)), ("st_gcns.0.residual.1.bias", (5,)), ("st_gcns.0.prelu.weight", (1,)),
     ("tpcnns.0.weight", (12, 8, 3, 3)), ("tpcnns.0.bias", (12,))]
    + [(f"tpcnns.{k}.{w}", shp) for k in range(1, 5) for w, shp in (("weight", (12, 12, 3, 3)), ("bias", (12,)))]
    + [("tpcnn_ouput.weight", (12, 12, 3, 3)), ("tpcnn_ouput.bias", (12,))]
    + [(f"prelus.{k}.weight", (1,)) for k in range(5)])
BN_MODULES = ("st_gcns.0.tcn.0", "st_gcns.0.tcn.3", "st_gcns.0.residual.1")


def flatten_state_dict(sd):
    """The reference's 40-key state_dict (val_best.pth) -> (params[7563], bn[30], num_batches_tracked)."""
    params = np.concatenate([np.asarray(sd[k].detach().cpu().numpy(), np.float32).reshape(-1)
                             for k, _ in PARAM_LAYOUT])
    bn = np.concatenate([np.asarray(sd[f"{m}.{b}"].detach().cpu().numpy(), np.float32).reshape(-1)
                         for m in BN_MODULES for b in ("running_mean", "running_var")])
    nbt = int(sd.get(BN_MODULES[0] + ".num_batches_tracked", 0))
    if params.size != L.N_PARAMS or bn.size != L.N_BN:
        raise ValueError("state_dict does not have the reference model's shapes")
    return params, bn, nbt


def unflatten_state_dict(params, bn, num_batches_tracked: int = 0):
    """(params[7563], bn[30]) -> the reference's state_dict (CPU tensors, torch.save-able as val_best.pth)."""
    from collections import OrderedDict
    params = np.asarray(params, np.float32).reshape(-1)
    bn = np.asarray(bn, np.float32).reshape(-1)
    flat, off = {}, 0
    for k, shp in PARAM_LAYOUT:
        n = int(np.prod(shp))
        flat[k] = torch.from_numpy(params[off:off + n].reshape(shp).copy())
        off += n
    bnt = {}
    for i, m in enumerate(BN_MODULES):
        bnt[m + ".running_mean"] = torch.from_numpy(bn[10 * i:10 * i + 5].copy())
        bnt[m + ".running_var"] = torch.from_numpy(bn[10 * i + 5:10 * i + 10].copy())
        bnt[m + ".num_batches_tracked"] = torch.tensor(int(num_batches_tracked))
    sd = OrderedDict()                      # nn.Module.state_dict() order: a module's params, then its buffers
    for k, _ in PARAM_LAYOUT:
        sd[k] = flat[k]
        mod = k.rsplit(".", 1)[0]
        if mod in BN_MODULES and k.endswith(".bias"):
            for b in ("running_mean", "running_var", "num_batches_tracked"):
                sd[f"{mod}.{b}"] = bnt[f"{mod}.{b}"]
    return sd


def group_weights(S: int, batch_size: int) -> np.ndarray:
    """Loss weights of one optimiser group of train() (train.py:58-66): every scene but the
    last is scored with weight 1/batch_size; the last one is forwarded (it updates the BN
    running statistics) but contributes no loss."""
    w = np.full(S, 1.0 / batch_size, np.float32)
    if S > 0:
        w[-1] = 0.0
    return w


class TrainEngine:
    """Weights + BN buffers resident on one GPU; `step` runs one optimiser group of the
    reference's train() (train.py:28-79) on a packed ragged batch:
    graph build -> train-mode forward -> NLL + gradient -> backward -> BN running stats -> SGD."""

    def __init__(self, params, bn, device="cuda", num_batches_tracked: int = 0):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("TrainEngine needs a CUDA device (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        L.load()
        self.device = device
        self.params = torch.as_tensor(np.asarray(params, np.float32)).reshape(-1).to(device).contiguous().clone()
        self.bn = torch.as_tensor(np.asarray(bn, np.float32)).reshape(-1).to(device).contiguous().clone()
        if self.params.numel() != L.N_PARAMS or self.bn.numel() != L.N_BN:
            raise ValueError("params must hold 7563 floats and bn 30 floats")
        self.num_batches_tracked = int(num_batches_tracked)
        self.grad = torch.zeros(L.N_PARAMS, dtype=torch.float32, device=device)
        self._ws: Dict[str, torch.Tensor] = {}

    @classmethod
    def from_state_dict(cls, sd, device="cuda"):
        params, bn, nbt = flatten_state_dict(sd)
        return cls(params, bn, device, nbt)

    def state_dict(self):
        """The reference's state_dict (40 keys) of the current weights: `torch.save(eng.state_dict(),
        'val_best.pth')` is loadable by the reference's test.py:183-186 unchanged."""
        return unflatten_state_dict(self.params.cpu().numpy(), self.bn.cpu().numpy(), self.num_batches_tracked)

    def forward_backward(self, plan: ScenePlan, obs_rel, pred_rel, scene_weight, *, graph_fp64: bool = True):
        """Everything of a group except the parameter update.  Returns dict(loss[S], grad[7563],
        scene_stats[S,30], V_pred)."""
        dp = plan.on(self.device)
        V, A = graph_build(dp, obs_rel, L.T_OBS, fp64=graph_fp64)
        V_tr, _ = graph_build(dp, pred_rel, L.T_PRED, v_only=True)
        vp, save, stats = forward_train(dp, V, A, self.params)
        w = torch.as_tensor(np.asarray(scene_weight, np.float32)).to(self.device) \
            if not isinstance(scene_weight, torch.Tensor) else scene_weight
        loss, dvp = nll_backward(dp, vp, V_tr, w)
        backward(dp, V, A, self.params, save, stats, dvp, grad=self.grad, ws=self._ws)
        return {"loss": loss, "grad": self.grad, "scene_stats": stats, "V_pred": vp, "weight": w}

    def _step_buffers(self, plan: ScenePlan, batch_size: int):
        """Grow-only scratch of sstg_train_step, per-scene loss, and the cached loss weights."""
        lib = L.load()
        need = int(lib.sstg_train_step_workspace(plan.S, plan.n_tiles, plan.P, plan.a_floats))
        if need < 0:
            L.check(need, "sstg_train_step_workspace")
        ws = self._ws.get("step")
        if ws is None or ws.numel() < need:
            self._ws["step"] = None
            ws = self._ws["step"] = torch.empty(max(need, 4), dtype=torch.float32, device=self.device)
        loss = self._ws.get("loss")
        if loss is None or loss.numel() < plan.S + 1:
            loss = self._ws["loss"] = torch.empty(max(plan.S + 1, 129), dtype=torch.float32, device=self.device)
        wkey = ("w", plan.S, batch_size)
        w = self._ws.get(wkey)
        if w is None:
            w = self._ws[wkey] = torch.from_numpy(group_weights(plan.S, batch_size)).to(self.device)
        return ws, need, loss, w

    def step(self, plan: ScenePlan, obs_rel, pred_rel, batch_size: int, lr: float, clip_grad=None,
             *, graph_fp64: bool = False):
        """One optimiser group in ONE library call (sstg_train_step): returns the group loss
        (sum_i w_i * l_i) as a 0-dim device tensor (valid until the next step).
        graph_fp64=True builds A with the reference's fp64 islands (bit-identical to networkx);
        the default fp32 graph is within 2e-6 of it."""
        dev = self.device
        dp = plan.on(dev)
        _need(obs_rel, torch.float32, "obs_rel", dev)
        _need(pred_rel, torch.float32, "pred_rel", dev)
        if tuple(obs_rel.shape) != (plan.P, 2, L.T_OBS) or tuple(pred_rel.shape) != (plan.P, 2, L.T_PRED):
            raise ValueError(f"obs_rel / pred_rel must be [{plan.P}, 2, 8] / [{plan.P}, 2, 12]")
        ws, need, loss, w = self._step_buffers(plan, batch_size)
        with torch.cuda.device(dev):
            L.check(L.load().sstg_train_step(
                _ptr(obs_rel), _ptr(pred_rel), _ptr(dp.seg_off), _ptr(dp.a_off), _ptr(dp.tile_off), plan.S,
                plan.n_tiles, plan.max_n, plan.P, plan.a_floats, _ptr(self.params), _ptr(self.bn), _ptr(w), _ptr(ws),
                need, float(lr), float(clip_grad or 0.0), 0.1, L.GRAPH_FP64 if graph_fp64 else 0,
                _ptr(loss), _ptr(self.grad), loss.data_ptr() + 4 * plan.S, _stream(dev)), "sstg_train_step")
        self.num_batches_tracked += plan.S
        self.last_loss = loss[:plan.S]
        return loss[plan.S]

    def shard_forward_backward(self, plan: Optional[ScenePlan], obs_rel, pred_rel, scene_weight, *,
                               graph_fp64: bool = False):
        """This rank's part of a scene-sharded optimiser group in one library call (sstg_train_step with
        lr = 0 and the BN recurrence deferred): returns device pointers-as-tensors (grad_local[7563],
        weighted loss[1], scene_stats[S,30]) for PeerExchange.exchange.  `plan` None = empty shard."""
        dev = self.device
        gl = self._ws.get("grad_local")
        if gl is None:
            gl = self._ws["grad_local"] = torch.zeros(L.N_PARAMS + 1, dtype=torch.float32, device=dev)
        if plan is None or plan.S == 0:
            gl.zero_()
            return gl, gl[L.N_PARAMS:], None
        dp = plan.on(dev)
        _need(obs_rel, torch.float32, "obs_rel", dev)
        _need(pred_rel, torch.float32, "pred_rel", dev)
        _need(scene_weight, torch.float32, "scene_weight", dev)
        if tuple(obs_rel.shape) != (plan.P, 2, L.T_OBS) or tuple(pred_rel.shape) != (plan.P, 2, L.T_PRED):
            raise ValueError(f"obs_rel / pred_rel must be [{plan.P}, 2, 8] / [{plan.P}, 2, 12]")
        if scene_weight.numel() != plan.S:
            raise ValueError("scene_weight must hold one weight per scene of the shard")
        ws, need, loss, _ = self._step_buffers(plan, 1)
        lib = L.load()
        off = int(lib.sstg_train_step_stats_offset(plan.S, plan.n_tiles, plan.P, plan.a_floats))
        with torch.cuda.device(dev):
            L.check(lib.sstg_train_step(
                _ptr(obs_rel), _ptr(pred_rel), _ptr(dp.seg_off), _ptr(dp.a_off), _ptr(dp.tile_off), plan.S,
                plan.n_tiles, plan.max_n, plan.P, plan.a_floats, _ptr(self.params), _ptr(self.bn), _ptr(scene_weight),
                _ptr(ws), need, 0.0, 0.0, -1.0, L.GRAPH_FP64 if graph_fp64 else 0,
                _ptr(loss), _ptr(gl), gl.data_ptr() + 4 * L.N_PARAMS, _stream(dev)), "sstg_train_step")
        self.last_loss = loss[:plan.S]
        return gl, gl[L.N_PARAMS:], ws[off:off + plan.S * L.N_BN].view(plan.S, L.N_BN)

    def step_unfused(self, plan: ScenePlan, obs_rel, pred_rel, batch_size: int, lr: float, clip_grad=None,
             *, graph_fp64: bool = True):
        """One optimiser group: returns the group loss (sum_i w_i * l_i) as a 0-dim device tensor."""
        r = self.forward_backward(plan, obs_rel, pred_rel, group_weights(plan.S, batch_size), graph_fp64=graph_fp64)
        bn_running_update(plan.on(self.device), r["scene_stats"], self.bn)
        self.num_batches_tracked += plan.S
        sgd_step(self.params, self.grad, lr, clip_grad or 0.0)
        return (r["loss"] * r["weight"]).sum()
