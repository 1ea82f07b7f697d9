"""Multi-GPU layer: one process per GPU (torch.distributed, NCCL over NVLink), scenes sharded
contiguously across ranks.

Scenes are independent everywhere on the path (the reference runs batch 1, so even BatchNorm
sees one scene at a time), therefore:
  * inference / evaluation needs NO data-path collective -- each rank runs the whole pipeline
    on its shard; the dataset-level ADE/FDE (test.py:125-126: mean over ALL pedestrians) is one
    3-element all-reduce of (sum min-ADE, sum min-FDE, pedestrians);
  * training needs ONE all-reduce (sum) per optimiser group over a flat buffer that carries the
    7,563-float gradient (the loss weights 1/batch_size and "last scene unscored" are applied
    locally, so the sum of the rank gradients IS the group gradient), the group loss, and the
    per-scene BatchNorm statistics (30 floats per scene, each rank filling only its own rows, so
    the sum is the concatenation): every rank then applies the running-statistics recurrence in
    the global scene order and all replicas stay bit-identical.
The collectives are torch.distributed calls (plumbing); the arithmetic before and after them
is libsstg.so.  The helpers below take any process group, so the host logic is tested with
gloo on CPU tensors (tests/test_parallel_gloo.py) and runs unchanged over NCCL.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L


def scene_cost(n_peds) -> np.ndarray:
    """Relative cost of a scene: the adjacency traffic 32*N^2 B dominates graph build and st_gcn,
    the TXP-CNN / sampling work is linear in N."""
    n = np.asarray(n_peds, np.float64)
    return 32.0 * n * n + 1200.0 * n


def shard_bounds(n_peds, world: int) -> List[Tuple[int, int]]:
    """Contiguous partition of the scenes into `world` shards with balanced cost.
    Returns [(lo, hi)] per rank; shards may be empty when there are fewer scenes than ranks."""
    if world < 1:
        raise ValueError("world must be >= 1")
    n = np.asarray(n_peds).reshape(-1)
    S = int(n.size)
    if S == 0:
        return [(0, 0)] * world
    c = np.cumsum(scene_cost(n))
    total = c[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(c, target, side="left"))
        # put the boundary scene on the side that leaves the smaller imbalance
        if i < S and (c[i] - target) < (target - (c[i - 1] if i > 0 else 0.0)):
            i += 1
        cuts.append(min(max(i, cuts[-1]), S))
    cuts.append(S)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_slice(n_peds, world: int, rank: int):
    """(scene lo, scene hi, pedestrian lo, pedestrian hi) of `rank`."""
    lo, hi = shard_bounds(n_peds, world)[rank]
    seg = np.concatenate([[0], np.cumsum(np.asarray(n_peds).reshape(-1))])
    return lo, hi, int(seg[lo]), int(seg[hi])


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def reduce_scores(sum_ade, sum_fde, n_peds_local, group=None, device=None):
    """Dataset-level (ADE, FDE) = sums over all ranks / total pedestrians (test.py:125-126)."""
    t = torch.tensor([float(sum_ade), float(sum_fde), float(n_peds_local)], dtype=torch.float64, device=device)
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    tot = max(float(t[2]), 1.0)
    return float(t[0]) / tot, float(t[1]) / tot, int(t[2])


def allreduce_gradient(grad: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the flat gradient over the ranks, in place (30 KB: latency-bound over NVLink)."""
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return grad


def gather_scene_stats(stats_local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gather the per-scene BatchNorm statistics [S_r, 30] of every rank into the global
    scene order [sum S_r, 30] (shards are contiguous, so rank order == scene order)."""
    world, _ = _world(group)
    if world == 1:
        return stats_local
    counts = [int(c) for c in counts]
    mx = max(max(counts), 1)
    pad = torch.zeros(mx, L.N_BN, dtype=stats_local.dtype, device=stats_local.device)
    if stats_local.numel():
        pad[: stats_local.shape[0]] = stats_local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def bind_process_to_gpu_numa(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (read from sysfs), BEFORE the
    pinned host buffers are allocated, so that first-touch places them next to the GPU's PCIe root
    and host->device copies of all ranks do not cross the socket interconnect.  Returns the node, or
    None when the topology is not exposed (then nothing changes)."""
    import os
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


class ShardedEvaluator:
    """test.py's evaluation over a scene-sharded dataset: this rank's InferenceEngine on this
    rank's shard, then the 3-element score reduction."""

    def __init__(self, engine, group=None):
        self.engine, self.group = engine, group

    def evaluate(self, n_peds, obs_rel, pred_rel, last_obs, K: int = 20, seed: int = 0, **kw):
        """Inputs describe the WHOLE dataset (host or device tensors packed like
        TrajectoryDataset.obs_traj_rel); each rank slices its own shard."""
        from .engine import ScenePlan
        world, rank = _world(self.group)
        lo, hi, p_lo, p_hi = shard_slice(n_peds, world, rank)
        dev = self.engine.device
        if hi == lo:
            return reduce_scores(0.0, 0.0, 0, self.group, dev)
        plan = ScenePlan(np.asarray(n_peds)[lo:hi])
        cut = lambda t: torch.as_tensor(t)[p_lo:p_hi].to(dev, non_blocking=True).contiguous()
        r = self.engine.infer(plan, cut(obs_rel), cut(pred_rel), cut(last_obs), K, seed=seed + rank, **kw)
        return reduce_scores(r["min_ade"].sum(), r["min_fde"].sum(), plan.P, self.group, dev)


class DistributedTrainer:
    """One optimiser group of train() (train.py:28-79) spread over the ranks of `group`.

    Every rank holds identical weights and BN buffers (a TrainEngine).  `step` takes the
    WHOLE group (all ranks pass the same n_peds / tensors, or each its own slice with
    `presliced=True`), runs graph build + train-mode forward + backward on this rank's
    contiguous shard with the group's loss weights, all-reduces the flat gradient, gathers
    the BN statistics, and applies the identical SGD step everywhere."""

    def __init__(self, engine, group=None):
        self.engine, self.group = engine, group
        self._flat = {}

    def step(self, n_peds, obs_rel, pred_rel, batch_size: int, lr: float, clip_grad: Optional[float] = None,
             presliced: bool = False):
        from . import engine as E
        world, rank = _world(self.group)
        n_peds = np.asarray(n_peds, np.int32).reshape(-1)
        bounds = shard_bounds(n_peds, world)
        lo, hi = bounds[rank]
        seg = np.concatenate([[0], np.cumsum(n_peds)])
        eng, dev = self.engine, self.engine.device
        w_all = E.group_weights(int(n_peds.size), batch_size)
        loss_local = torch.zeros(1, dtype=torch.float64, device=dev)
        if hi > lo:
            plan = E.ScenePlan(n_peds[lo:hi])
            if presliced:
                o, p = obs_rel, pred_rel
            else:
                o = torch.as_tensor(obs_rel)[seg[lo]:seg[hi]].to(dev).contiguous()
                p = torch.as_tensor(pred_rel)[seg[lo]:seg[hi]].to(dev).contiguous()
            r = eng.forward_backward(plan, o, p, w_all[lo:hi])
            stats = r["scene_stats"]
            loss_local += (r["loss"].double() * r["weight"].double()).sum()
        else:
            eng.grad.zero_()
            stats = torch.zeros(0, L.N_BN, dtype=torch.float32, device=dev)
        # ONE collective per optimiser group: [gradient 7563 | group loss 1 | per-scene BN statistics S x 30]
        # in a single flat buffer, all-reduced (sum).  Every rank fills only its own scenes' rows of the
        # statistics block, so the sum IS the concatenation in global scene order.
        S_all = int(n_peds.size)
        if world > 1:
            flat = self._flat.get(S_all)
            if flat is None:
                flat = self._flat[S_all] = torch.empty(L.N_PARAMS + 1 + S_all * L.N_BN, dtype=torch.float32, device=dev)
            flat[:L.N_PARAMS].copy_(eng.grad)
            flat[L.N_PARAMS] = loss_local[0].float()
            st = flat[L.N_PARAMS + 1:].view(S_all, L.N_BN)
            st.zero_()
            if hi > lo:
                st[lo:hi].copy_(stats)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            eng.grad.copy_(flat[:L.N_PARAMS])
            loss_local = flat[L.N_PARAMS:L.N_PARAMS + 1].double()
            stats_all = st
        else:
            stats_all = stats
        full = E.ScenePlan(n_peds)
        E.bn_running_update(full.on(dev), stats_all.contiguous(), eng.bn)
        eng.num_batches_tracked += int(n_peds.size)
        E.sgd_step(eng.params, eng.grad, lr, clip_grad or 0.0)
        return loss_local[0]
