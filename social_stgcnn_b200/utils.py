"""Drop-in for the reference's utils.py: `anorm`, `seq_to_graph`, `poly_fit`, `read_file`,
`TrajectoryDataset` -- same names, arguments, return layouts and error behaviour
(reference utils.py:23-228), with every graph built by the CUDA kernel of libsstg.so.

What changed underneath:
  * seq_to_graph runs sstg_graph_build in its fp64-island mode, which reproduces the
    networkx normalized Laplacian bit for bit (tests/test_gpu_parity.py);
  * TrajectoryDataset builds the graphs of ALL scenes (both halves) in two kernel launches
    on a packed ragged batch instead of O(scenes * T * N^2) Python calls, and keeps the
    packed device tensors (`.packed`) for the batched pipeline next to the per-scene CPU
    tensors the reference exposes (`v_obs`, `A_obs`, `v_pred`, `A_pred`);
  * the text parsing/windowing is vectorised numpy (same scenes, same order).
There is no CPU fallback for the graph math: without a CUDA device these raise.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
from torch.utils.data import Dataset

from .engine import ScenePlan, graph_build


def _cuda_device(t=None):
    if isinstance(t, torch.Tensor) and t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("social_stgcnn_b200 needs a CUDA device: graph construction runs in libsstg.so "
                           "(there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def anorm(p1, p2):
    """utils.py:23-27 -- scalar helper kept for API parity (the kernels inline it)."""
    dx = np.float32(p1[0]) - np.float32(p2[0])
    dy = np.float32(p1[1]) - np.float32(p2[1])
    norm = math.sqrt(np.float32(np.float32(dx * dx) + np.float32(dy * dy)))
    return 0 if norm == 0 else 1 / norm


def seq_to_graph(seq_, seq_rel, norm_lap_matr=True):
    """utils.py:29-53.  seq_, seq_rel: (N, 2, T) tensors (extra singleton dims squeezed).
    Returns V (T, N, 2) and A (T, N, N) float32 tensors on the device of `seq_rel`
    (CPU tensors in, CPU tensors out -- the arithmetic still runs on the GPU)."""
    seq_ = torch.as_tensor(seq_).squeeze()
    rel_in = torch.as_tensor(seq_rel)
    rel = rel_in.squeeze()
    if rel.dim() == 2:                       # a single pedestrian squeezed away
        rel = rel[None]
    if rel.dim() != 3 or rel.shape[1] != 2:
        raise IndexError(f"seq_rel must be (N, 2, T) after squeeze, got {tuple(rel_in.shape)}")
    n, _, T = rel.shape
    dev = _cuda_device(rel_in)
    rel_d = rel.to(device=dev, dtype=torch.float32).contiguous()
    plan = ScenePlan([n], T=T)
    V, A = graph_build(plan.on(dev), rel_d, T, norm_lap=bool(norm_lap_matr), fp64=True)
    V = V.view(T, n, 2)
    A = A[: T * n * n].view(T, n, n)
    if not rel_in.is_cuda:
        V, A = V.cpu(), A.cpu()
    return V, A


def poly_fit(traj, traj_len, threshold):
    """utils.py:56-71: 1.0 if the 2nd-order fit residual of the last traj_len points is >= threshold."""
    t = np.linspace(0, traj_len - 1, traj_len)
    res_x = np.polyfit(t, traj[0, -traj_len:], 2, full=True)[1]
    res_y = np.polyfit(t, traj[1, -traj_len:], 2, full=True)[1]
    return 1.0 if res_x + res_y >= threshold else 0.0


def _poly_fit_batch(seq, traj_len, threshold):
    """poly_fit (utils.py:56-71) for every pedestrian of a window in two least-squares solves:
    np.polyfit accepts a 2-D right-hand side, so the residuals come from the same LAPACK path."""
    n = seq.shape[0]
    t = np.linspace(0, traj_len - 1, traj_len)
    res_x = np.polyfit(t, seq[:, 0, -traj_len:].T, 2, full=True)[1]
    res_y = np.polyfit(t, seq[:, 1, -traj_len:].T, 2, full=True)[1]
    if res_x.size != n or res_y.size != n:      # rank-deficient fit returns no residuals; mirror the scalar path
        return np.array([poly_fit(s, traj_len, threshold) for s in seq])
    return (res_x + res_y >= threshold).astype(np.float64)


def read_file(_path, delim="\t"):
    """utils.py:72-83: rows of `<frame> <ped> <x> <y>`; ValueError on a wrong delimiter."""
    delim = {"tab": "\t", "space": " "}.get(delim, delim)
    rows = []
    with open(_path, "r") as f:
        for line in f:
            rows.append([float(tok) for tok in line.strip().split(delim)])
    return np.asarray(rows)


def _scenes_of_file(data, seq_len, skip, min_ped, pred_len, threshold):
    """Sliding windows of one file (utils.py:121-170), vectorised: a pedestrian belongs to a
    window iff it is present at the window's first and last frame (pad_front == 0 and
    pad_end == seq_len, utils.py:146-149).  Yields (abs (n,2,L), rel (n,2,L), non_linear (n,),
    peds_seen) per kept window, pedestrians in ascending id order like np.unique.
    Memory is O(rows) like the reference's: the rows are sorted by frame once, a window is the
    contiguous row range of its seq_len frames (no dense pedestrian x frame table, which long
    recordings with many track ids would blow up)."""
    frames, fi = np.unique(data[:, 0], return_inverse=True)
    F = frames.size
    by_frame = np.argsort(fi, kind="stable")                      # frame order, file order inside a frame
    fi_s = fi[by_frame]
    ped_s = data[by_frame, 1]
    xy_s = np.around(data[by_frame, 2:4], decimals=4)             # utils.py:145
    frame_start = np.searchsorted(fi_s, np.arange(F + 1))         # rows of frame f: [frame_start[f], frame_start[f+1])
    num_sequences = int(math.ceil((F - seq_len + 1) / skip))
    for idx in range(0, num_sequences * skip + 1, skip):          # "+ 1" as in utils.py:129
        lo, hi = frame_start[min(idx, F)], frame_start[min(idx + seq_len, F)]
        ids, inv, cnt = np.unique(ped_s[lo:hi], return_inverse=True, return_counts=True)
        seen = int(ids.size)
        if idx + seq_len > F:                                     # trailing short window: nobody can span it
            yield None, None, None, seen
            continue
        rel_f = fi_s[lo:hi] - idx
        first = np.bincount(inv[rel_f == 0], minlength=seen) > 0
        last = np.bincount(inv[rel_f == seq_len - 1], minlength=seen) > 0
        ok = first & last
        if np.any(ok & (cnt != seq_len)):
            raise ValueError("could not broadcast input array: a pedestrian spans the window but misses "
                             "frames inside it")               # what utils.py:158 would raise
        n = int(ok.sum())
        if n <= min_ped:                                          # utils.py:165
            yield None, None, None, seen
            continue
        rows = np.nonzero(ok[inv])[0]                             # window rows of the kept pedestrians ...
        rows = rows[np.argsort(inv[rows], kind="stable")]         # ... by pedestrian id, frame order inside
        seq = xy_s[lo + rows].reshape(n, seq_len, 2).transpose(0, 2, 1)   # (n, 2, L)
        rel = np.zeros_like(seq)
        rel[:, :, 1:] = seq[:, :, 1:] - seq[:, :, :-1]              # utils.py:153-155
        nl = _poly_fit_batch(seq, pred_len, threshold)
        yield seq, rel, nl, seen


def load_scenes(data_dir, obs_len, pred_len, skip, threshold, min_ped, delim):
    """Host-side half of TrajectoryDataset.__init__ (utils.py:114-170): parse every file of
    `data_dir` in os.listdir order and window it.  Pure numpy, no GPU."""
    seq_len = obs_len + pred_len
    seqs, rels, masks, non_linear, counts, max_peds = [], [], [], [], [], 0
    for path in [os.path.join(data_dir, p) for p in os.listdir(data_dir)]:      # utils.py:114 order
        data = read_file(path, delim)
        for seq, rel, nl, seen in _scenes_of_file(data, seq_len, skip, min_ped, pred_len, threshold):
            max_peds = max(max_peds, seen)
            if seq is None:
                continue
            seqs.append(seq)
            rels.append(rel)
            non_linear.append(nl)
            masks.append(np.ones((seq.shape[0], seq_len)))
            counts.append(seq.shape[0])
    return seqs, rels, masks, non_linear, counts, max_peds


class _LazyPacked(dict):
    """The dataset's packed device tensors; "A_tr" is computed the first time it is asked for."""

    def __init__(self, make_a_tr, items):
        super().__init__(items)
        self._make_a_tr = make_a_tr

    def __missing__(self, key):
        if key != "A_tr":
            raise KeyError(key)
        self["A_tr"] = self._make_a_tr()
        return dict.__getitem__(self, "A_tr")


class _LazyAdjacencyList:
    """`dataset.A_pred`: a list of per-scene (T_pred, N, N) views, materialised on first use."""

    def __init__(self, dataset):
        self._ds, self._views = dataset, None

    def _get(self):
        if self._views is None:
            self._views = self._ds._pred_adjacency_views()
        return self._views

    def __getitem__(self, i):
        return self._get()[i]

    def __len__(self):
        return self._ds.num_seq

    def __iter__(self):
        return iter(self._get())


class TrajectoryDataset(Dataset):
    """Dataloder for the Trajectory datasets -- drop-in for utils.py:86-228."""

    def __init__(self, data_dir, obs_len=8, pred_len=8, skip=1, threshold=0.002, min_ped=1, delim="\t",
                 norm_lap_matr=True, device=None):
        super().__init__()
        self.max_peds_in_frame = 0
        self.data_dir = data_dir
        self.obs_len, self.pred_len, self.skip = obs_len, pred_len, skip
        self.seq_len = obs_len + pred_len
        self.delim, self.norm_lap_matr = delim, norm_lap_matr

        seqs, rels, masks, non_linear, counts, self.max_peds_in_frame = load_scenes(
            data_dir, obs_len, pred_len, skip, threshold, min_ped, delim)

        self.num_seq = len(seqs)
        from ._lib import MAX_PEDS
        if counts and max(counts) > MAX_PEDS:
            raise ValueError(f"{data_dir}: a window holds {max(counts)} pedestrians, but the kernels of libsstg.so "
                             f"accept at most SSTG_MAX_PEDS = {MAX_PEDS} per scene (include/sstg.h; shared-memory "
                             "bound of the st_gcn / graph kernels)")
        seq_all = np.concatenate(seqs, axis=0)
        rel_all = np.concatenate(rels, axis=0)
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).type(torch.float)
        self.obs_traj = f32(seq_all[:, :, :obs_len])
        self.pred_traj = f32(seq_all[:, :, obs_len:])
        self.obs_traj_rel = f32(rel_all[:, :, :obs_len])
        self.pred_traj_rel = f32(rel_all[:, :, obs_len:])
        self.loss_mask = f32(np.concatenate(masks, axis=0))
        self.non_linear_ped = f32(np.concatenate(non_linear))
        cum = [0] + np.cumsum(counts).tolist()
        self.seq_start_end = list(zip(cum, cum[1:]))
        # the same tracks in the 88 B/pedestrian fixed-point host format (None if a step overflows int16)
        try:
            from .engine import pack_tracks
            self.compact = pack_tracks(seq_all) if seq_all.shape[2] == 20 else None
        except OverflowError:
            self.compact = None

        # ---- graphs: two launches over the packed batch instead of 2 * num_seq Python graph builds
        dev = torch.device(device) if device is not None else _cuda_device()
        self.plan_obs = ScenePlan(counts, T=obs_len)
        self.plan_pred = ScenePlan(counts, T=pred_len)
        obs_rel_d = self.obs_traj_rel.to(dev).contiguous()
        pred_rel_d = self.pred_traj_rel.to(dev).contiguous()
        V_o, A_o = graph_build(self.plan_obs.on(dev), obs_rel_d, obs_len, norm_lap=norm_lap_matr, fp64=True)
        # The pred-half adjacency A_tr (utils.py:209-211) is 60 % of the reference's graph work and nothing on
        # the path consumes it: only V_tr is built here; A_tr is built on first access of
        # `A_pred[...]`, item [9] of __getitem__, or `packed["A_tr"]`.
        V_p, _ = graph_build(self.plan_pred.on(dev), pred_rel_d, pred_len, v_only=True)
        self.packed = _LazyPacked(self._build_pred_adjacency,
                                  {"device": dev, "obs_rel": obs_rel_d, "pred_rel": pred_rel_d, "V_obs": V_o,
                                   "A_obs": A_o, "V_tr": V_p,
                                   "last_obs": self.obs_traj[:, :, -1].to(dev).contiguous()})
        V_oc, A_oc, V_pc = V_o.cpu(), A_o.cpu(), V_p.cpu()
        self.v_obs, self.A_obs, self.v_pred = [], [], []
        self._counts, self._cum = list(counts), cum
        for i, n in enumerate(counts):
            s = cum[i]
            ao = int(self.plan_obs.a_off[i])
            self.v_obs.append(V_oc[s * obs_len * 2:(s + n) * obs_len * 2].view(obs_len, n, 2))
            self.A_obs.append(A_oc[ao:ao + obs_len * n * n].view(obs_len, n, n))
            self.v_pred.append(V_pc[s * pred_len * 2:(s + n) * pred_len * 2].view(pred_len, n, 2))
        self.A_pred = _LazyAdjacencyList(self)

    def _build_pred_adjacency(self):
        """A_tr of every scene, packed on the device (one launch over the whole dataset)."""
        dev = self.packed["device"]
        _, A_p = graph_build(self.plan_pred.on(dev), self.packed["pred_rel"], self.pred_len,
                             norm_lap=self.norm_lap_matr, fp64=True)
        return A_p

    def _pred_adjacency_views(self):
        A_pc = self.packed["A_tr"].cpu()
        out = []
        for i, n in enumerate(self._counts):
            ap = int(self.plan_pred.a_off[i])
            out.append(A_pc[ap:ap + self.pred_len * n * n].view(self.pred_len, n, n))
        return out

    def __len__(self):
        return self.num_seq

    def __getitem__(self, index):
        start, end = self.seq_start_end[index]
        return [
            self.obs_traj[start:end, :], self.pred_traj[start:end, :],
            self.obs_traj_rel[start:end, :], self.pred_traj_rel[start:end, :],
            self.non_linear_ped[start:end], self.loss_mask[start:end, :],
            self.v_obs[index], self.A_obs[index], self.v_pred[index], self.A_pred[index],
        ]
