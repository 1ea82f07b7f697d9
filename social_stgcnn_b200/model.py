"""Drop-in for the reference's model.py: `social_stgcnn(v, a) -> (V_pred, a)` as an nn.Module
with the reference's constructor, submodule tree and 40-key `state_dict()` (reference
model.py:157-198), whose forward AND backward run in libsstg.so.

  * the submodules (`st_gcns.0.gcn.conv`, `st_gcns.0.tcn.{0..4}`, `st_gcns.0.residual.{0,1}`,
    `st_gcns.0.prelu`, `tpcnns.{0..4}`, `tpcnn_ouput` (sic), `prelus.{0..4}`) are parameter
    containers created in the reference's definition order, so `load_state_dict` of the shipped
    checkpoints, default initialisation under a seed, and `optimizer` parameter order all match;
    none of their own forward methods is ever called;
  * eval mode  -> sstg_stgcn_forward + sstg_txpcnn_forward (BatchNorm from the running stats);
  * train mode -> sstg_forward_train (BatchNorm batch statistics of this scene, running stats
    updated like nn.BatchNorm2d) under a torch.autograd.Function whose backward is
    sstg_backward; gradients reach `.grad` of the module's parameters through autograd's own
    slicing of the flat parameter block.  `tpcnns.4` / `prelus.4` are never executed
    (model.py:191) and keep `grad is None`, as in the reference.
There is no CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import engine as E

_UNUSED = ("tpcnns.4.weight", "tpcnns.4.bias", "prelus.4.weight")
_BN_ORDER = ("st_gcns.0.tcn.0", "st_gcns.0.tcn.3", "st_gcns.0.residual.1")
_PLANS = {}


def _plan(n: int) -> E.ScenePlan:
    p = _PLANS.get(n)
    if p is None:
        p = _PLANS[n] = E.ScenePlan([n])
    return p


class _GraphConv(nn.Module):
    """Parameter container for ConvTemporalGraphical (model.py:18-68): `.conv` is the 1x1 Conv2d."""

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=(1, 1))


class st_gcn(nn.Module):
    """Parameter container for the st_gcn block (model.py:71-155); computed by the fused kernels."""

    def __init__(self, in_channels, out_channels, kernel_size, use_mdn=False, stride=1, dropout=0, residual=True):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0] % 2 == 1
        self.use_mdn = use_mdn
        self.gcn = _GraphConv(in_channels, out_channels, kernel_size[1])
        self.tcn = nn.Sequential(
            nn.BatchNorm2d(out_channels),
            nn.PReLU(),
            nn.Conv2d(out_channels, out_channels, (kernel_size[0], 1), (stride, 1), ((kernel_size[0] - 1) // 2, 0)),
            nn.BatchNorm2d(out_channels),
            nn.Dropout(dropout, inplace=True),
        )
        self.residual = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(stride, 1)),
            nn.BatchNorm2d(out_channels),
        )
        self.prelu = nn.PReLU()


class _TrainForward(torch.autograd.Function):
    """social_stgcnn.forward in train mode for ONE scene; backward = sstg_backward."""

    @staticmethod
    def forward(ctx, flat, V, A, module):
        n = V.shape[1]
        dp = _plan(n).on(V.device)
        vp, save, stats = E.forward_train(dp, V, A, flat)
        module._update_running_stats(dp, stats)
        ctx.dp = dp
        ctx.save_for_backward(flat, V, A, save, stats)
        return vp.view(1, L.C_OUT, L.T_PRED, n)

    @staticmethod
    def backward(ctx, g):
        flat, V, A, save, stats = ctx.saved_tensors
        g = g.contiguous().view(-1)
        grad = E.backward(ctx.dp, V, A, flat, save, stats, g)
        return grad, None, None, None


class social_stgcnn(nn.Module):
    def __init__(self, n_stgcnn=1, n_txpcnn=1, input_feat=2, output_feat=5, seq_len=8, pred_seq_len=12,
                 kernel_size=3):
        super().__init__()
        self.n_stgcnn = n_stgcnn
        self.n_txpcnn = n_txpcnn
        self.st_gcns = nn.ModuleList()
        self.st_gcns.append(st_gcn(input_feat, output_feat, (kernel_size, seq_len)))
        for _ in range(1, n_stgcnn):
            self.st_gcns.append(st_gcn(output_feat, output_feat, (kernel_size, seq_len)))
        self.tpcnns = nn.ModuleList()
        self.tpcnns.append(nn.Conv2d(seq_len, pred_seq_len, 3, padding=1))
        for _ in range(1, n_txpcnn):
            self.tpcnns.append(nn.Conv2d(pred_seq_len, pred_seq_len, 3, padding=1))
        self.tpcnn_ouput = nn.Conv2d(pred_seq_len, pred_seq_len, 3, padding=1)
        self.prelus = nn.ModuleList()
        for _ in range(n_txpcnn):
            self.prelus.append(nn.PReLU())
        self._cfg = (n_stgcnn, n_txpcnn, input_feat, output_feat, seq_len, pred_seq_len, kernel_size)

    # ---- flat views of the parameters / buffers in the C ABI's order --------------------
    def _flat_params(self):
        named = dict(self.named_parameters())
        parts = []
        for k, p in named.items():      # named_parameters() order == the ABI's parameter block order
            parts.append((p.detach() if k in _UNUSED else p).reshape(-1))
        return torch.cat(parts)

    def _bn_modules(self):
        s = self.st_gcns[0]
        return (s.tcn[0], s.tcn[3], s.residual[1])

    def _flat_bn(self):
        return torch.cat([t for m in self._bn_modules() for t in (m.running_mean, m.running_var)])

    @torch.no_grad()
    def _update_running_stats(self, dp, stats):
        bn = self._flat_bn()
        E.bn_running_update(dp, stats, bn, momentum=self._bn_modules()[0].momentum)
        for i, m in enumerate(self._bn_modules()):
            m.running_mean.copy_(bn[10 * i:10 * i + 5])
            m.running_var.copy_(bn[10 * i + 5:10 * i + 10])
            m.num_batches_tracked += 1

    def forward(self, v, a):
        if self._cfg != (1, 5, 2, 5, 8, 12, 3):
            raise NotImplementedError(
                "the CUDA path implements the shipped configuration (n_stgcnn=1, n_txpcnn=5, input_feat=2, "
                f"output_feat=5, seq_len=8, pred_seq_len=12, kernel_size=3); got {self._cfg}")
        if not (isinstance(v, torch.Tensor) and v.is_cuda and a.is_cuda):
            raise RuntimeError("social_stgcnn_b200.social_stgcnn needs CUDA tensors (there is no CPU fallback)")
        assert a.size(0) == self.st_gcns[0].gcn.kernel_size           # model.py:65
        if v.dim() != 4 or v.shape[0] != 1 or v.shape[1] != 2 or v.shape[2] != L.T_OBS:
            raise ValueError(f"v must be (1, 2, {L.T_OBS}, N), got {tuple(v.shape)}")
        n = v.shape[3]
        if a.dim() != 3 or a.shape[1] != n or a.shape[2] != n:
            raise ValueError(f"a must be ({L.T_OBS}, {n}, {n}), got {tuple(a.shape)}")
        # (1,2,8,N) -> V (8,N,2): free when v is the permuted view of V_obs the callers pass (train.py:48)
        V = v[0].permute(1, 2, 0).contiguous().float()
        A = a.contiguous().float()
        if A.data_ptr() % 16:
            A = A.clone()
        flat = self._flat_params()
        if self.training:
            out = _TrainForward.apply(flat, V, A, self)
        else:
            with torch.no_grad():
                dp = _plan(n).on(V.device)
                H = E.stgcn_forward(dp, V, A, flat.detach(), self._flat_bn())
                out = E.txpcnn_forward(dp, H, flat.detach()).view(1, L.C_OUT, L.T_PRED, n)
        return out, a
