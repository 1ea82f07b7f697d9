"""Drop-in for the reference's metrics.py: `ade`, `fde`, `seq_to_nodes`,
`nodes_rel_to_nodes_abs`, `closer_to_zero`, `bivariate_loss` -- same names, arguments and
return types (reference metrics.py:21-113), computed by libsstg.so kernels.

  * bivariate_loss is a torch.autograd.Function over sstg_nll_backward (loss and its
    gradient in one launch), so `loss.backward()` keeps working (train.py:59-67);
  * ade / fde / nodes_rel_to_nodes_abs take and return host numpy arrays / Python floats
    like the reference, but the arithmetic (fp32 differences, fp64 sqrt and sums; the
    sequential fp32 prefix sum) runs on the GPU with the reference's rounding points;
  * seq_to_nodes is a pure re-layout (no arithmetic) and stays in numpy.
There is no CPU fallback for the arithmetic.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import engine as E


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("social_stgcnn_b200.metrics needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _pack_scenes(arrs, counts):
    """list of (T, >=count, 2) arrays -> packed (T, count, 2) blocks, the plan, T."""
    T = int(arrs[0].shape[0])
    blocks = []
    for a, c in zip(arrs, counts):
        a = np.asarray(a)
        if a.ndim == 2:                      # a squeezed single pedestrian (T, 2)
            a = a[:, None, :]
        if a.shape[0] != T:
            raise ValueError("all scenes must have the same number of time steps")
        blocks.append(np.ascontiguousarray(a[:, :c, :], dtype=np.float32).reshape(-1))
    plan = E.ScenePlan([int(c) for c in counts])
    return np.concatenate(blocks), plan, T


def _traj_error(predAll, targetAll, count_):
    dev = _dev()
    pred, plan, T = _pack_scenes(predAll, count_)
    tgt, _, T2 = _pack_scenes(targetAll, count_)
    if T != T2:
        raise ValueError("pred and target have different lengths")
    dp = plan.on(dev)
    sd = torch.empty(plan.P, dtype=torch.float64, device=dev)
    fd = torch.empty(plan.P, dtype=torch.float64, device=dev)
    p_d, t_d = torch.from_numpy(pred).to(dev), torch.from_numpy(tgt).to(dev)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_traj_error(p_d.data_ptr(), t_d.data_ptr(), dp.seg_off.data_ptr(), plan.S, plan.P, T,
                                         sd.data_ptr(), fd.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream), "sstg_traj_error")
    return sd.cpu().numpy(), fd.cpu().numpy(), plan, T


def ade(predAll, targetAll, count_):
    """metrics.py:21-36: mean over scenes of the scene's mean displacement error."""
    sd, _, plan, T = _traj_error(predAll, targetAll, count_)
    tot = 0.0
    for s in range(plan.S):
        a, b = int(plan.seg_off[s]), int(plan.seg_off[s + 1])
        sum_ = 0.0
        for v in sd[a:b]:
            sum_ += float(v)
        tot += sum_ / ((b - a) * T)
    return tot / plan.S


def fde(predAll, targetAll, count_):
    """metrics.py:39-53: mean over scenes of the scene's mean final displacement error."""
    _, fd, plan, _ = _traj_error(predAll, targetAll, count_)
    tot = 0.0
    for s in range(plan.S):
        a, b = int(plan.seg_off[s]), int(plan.seg_off[s + 1])
        sum_ = 0.0
        for v in fd[a:b]:
            sum_ += float(v)
        tot += sum_ / (b - a)
    return tot / plan.S


def seq_to_nodes(seq_):
    """metrics.py:56-67: (1, N, 2, T) -> (T, N, 2) float64 (squeezed)."""
    seq_ = np.asarray(seq_)
    max_nodes = seq_.shape[1]
    s = seq_.squeeze()
    if s.ndim == 2:                          # N == 1 squeezed away: the reference iterates len(step_) == 2 rows
        s = s[None]
    V = np.zeros((s.shape[2], max_nodes, 2))
    V[:, :s.shape[0], :] = np.transpose(s, (2, 0, 1))
    return V.squeeze()


def nodes_rel_to_nodes_abs(nodes, init_node):
    """metrics.py:69-75: running sum of the relative steps plus the start node."""
    nodes = np.asarray(nodes)
    init_node = np.asarray(init_node)
    if nodes.ndim != 3:
        raise IndexError("nodes must be (T, N, 2)")
    T, N, _ = nodes.shape
    dev = _dev()
    plan = E.ScenePlan([N])
    dp = plan.on(dev)
    rel = torch.from_numpy(np.ascontiguousarray(nodes, dtype=np.float32)).to(dev)
    ini = torch.from_numpy(np.ascontiguousarray(init_node[:N], dtype=np.float64)).to(dev)
    out = torch.empty_like(rel)
    with torch.cuda.device(dev):
        L.check(L.load().sstg_rel_to_abs(rel.data_ptr(), ini.data_ptr(), dp.seg_off.data_ptr(), 1, N, T,
                                         out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                "sstg_rel_to_abs")
    return out.cpu().numpy().astype(nodes.dtype, copy=False).squeeze()


def closer_to_zero(current, new_v):
    """metrics.py:77-82."""
    dec = min([(abs(current), current), (abs(new_v), new_v)])[1]
    return dec != current


class _BivariateLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, V_pred, V_trgt):
        T, N = V_pred.shape[0], V_pred.shape[1]
        if (T * N) % L.T_PRED:
            raise ValueError("bivariate_loss kernel needs T*N to be a multiple of 12")
        vp = V_pred.permute(2, 0, 1).contiguous().float()          # (5, T, N): the kernel's block layout
        vt = V_trgt.contiguous().float()
        dp = E.ScenePlan([T * N // L.T_PRED]).on(vp.device)
        loss, dv = E.nll_backward(dp, vp, vt)
        ctx.save_for_backward(dv)
        ctx.shape = (T, N)
        return loss[0].clone()

    @staticmethod
    def backward(ctx, g):
        (dv,) = ctx.saved_tensors
        T, N = ctx.shape
        return (dv.view(5, T, N) * g).permute(1, 2, 0), None


def bivariate_loss(V_pred, V_trgt):
    """metrics.py:84-113.  V_pred (T, N, 5) = (mu_x, mu_y, log s_x, log s_y, atanh rho),
    V_trgt (T, N, 2) -> 0-dim tensor with autograd."""
    if not (isinstance(V_pred, torch.Tensor) and V_pred.is_cuda):
        raise RuntimeError("social_stgcnn_b200.bivariate_loss needs CUDA tensors (there is no CPU fallback)")
    if V_pred.dim() != 3 or V_pred.shape[2] != 5 or V_trgt.shape[:2] != V_pred.shape[:2]:
        raise ValueError(f"expected V_pred (T,N,5) and V_trgt (T,N,2), got {tuple(V_pred.shape)} {tuple(V_trgt.shape)}")
    return _BivariateLoss.apply(V_pred, V_trgt.to(V_pred.device))
