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


def _coll_device(group=None, device=None):
    """Where a small host-logic collective has to live: the GPU for NCCL, the CPU for gloo."""
    backend = dist.get_backend(group) if (dist.is_available() and dist.is_initialized()) else "gloo"
    if "nccl" in str(backend) and device is not None and torch.device(device).type == "cuda":
        return torch.device(device)
    return torch.device("cpu")


def broadcast_order(order, group=None, device=None, src: int = 0) -> np.ndarray:
    """Rank `src`'s epoch order (the DataLoader shuffle of train.py:173-177) on every rank.  The sharded
    optimiser group is only meaningful when all ranks cut the SAME scene sequence into groups."""
    order = np.ascontiguousarray(np.asarray(order, np.int64).reshape(-1))
    world, _ = _world(group)
    if world == 1:
        return order
    t = torch.from_numpy(order.copy()).to(_coll_device(group, device))
    dist.broadcast(t, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    return t.cpu().numpy()


def assert_same_across_ranks(arrays, group=None, device=None, what: str = "arguments") -> None:
    """Raise ValueError on EVERY rank when the ranks do not hold byte-identical `arrays` (one 4-element
    all-reduce of two 63-bit digests and their negations: max(x) == -max(-x) iff all ranks agree).  It is
    also a barrier."""
    world, _ = _world(group)
    if world == 1:
        return
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str((a.dtype.str, a.shape)).encode())
        h.update(a.tobytes())
    d = h.digest()
    x0 = int.from_bytes(d[:8], "little") >> 1
    x1 = int.from_bytes(d[8:], "little") >> 1
    t = torch.tensor([x0, -x0, x1, -x1], dtype=torch.int64, device=_coll_device(group, device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    t = t.cpu()
    if int(t[0]) != -int(t[1]) or int(t[2]) != -int(t[3]):
        raise ValueError(f"{what} differ between ranks: every rank of a sharded group must pass the same values "
                         "(draw the shuffle once and broadcast it -- parallel.broadcast_order)")


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


class PeerExchange:
    """The mailboxes of sstg_peer_exchange_apply (include/sstg.h) for one rank: its own mailbox in HBM
    plus the CUDA-IPC mappings of every other rank's.  The 64-byte handles travel through
    torch.distributed (plumbing); the exchange itself is libsstg.so's kernel storing over NVLink.

    All ranks must sit on one node (NVLink/NVSwitch peer access) and call `exchange` in lockstep."""

    def __init__(self, device, max_scenes: int = 256, group=None, timeout_s: float = 30.0, *, _shared=None):
        import atexit
        import ctypes as C
        import weakref
        # the IPC mappings and the cudaMalloc'd mailbox are not torch memory: release them when the object goes
        # away or at interpreter exit (close() is idempotent)
        self._own, self._opened = None, []
        ref = weakref.ref(self)
        atexit.register(lambda: ref() is not None and ref()._close_quietly())
        lib = L.load()
        self.device = torch.device(device)
        self.group, self.max_scenes, self.epoch = group, int(max_scenes), 0
        self.timeout_ns = int(timeout_s * 1e9)
        self._emulated = _shared is not None
        if _shared is not None:                      # several ranks emulated inside one process (tests)
            self.world, self.rank, ptrs = _shared
        else:
            self.world, self.rank = _world(group)
            nbytes = int(lib.sstg_peer_mailbox_bytes(self.world, self.max_scenes))
            if nbytes < 0:
                L.check(nbytes, "sstg_peer_mailbox_bytes")
            # Every rank takes part in the same two exchanges whatever happens locally (a rank that failed sends
            # its error text instead of a handle), so a failure surfaces on ALL ranks instead of leaving the
            # others stuck in a collective.
            import socket
            own, handle, err = C.c_void_p(), C.create_string_buffer(64), None
            try:
                with torch.cuda.device(self.device):
                    L.check(lib.sstg_peer_alloc(nbytes, C.byref(own), handle), "sstg_peer_alloc")
                self._own = own.value
            except Exception as e:                                   # noqa: BLE001 -- reported to every rank below
                err = f"rank {self.rank}: {e}"
            mine = (socket.gethostname(), handle.raw if err is None else None, err)
            everyone = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(everyone, mine, group=group)
            else:
                everyone = [mine]
            errs = [h[2] for h in everyone if h[2]]
            if not errs and any(h[0] != mine[0] for h in everyone):
                errs = ["PeerExchange needs all ranks on one node (peer memory over NVLink)"]
            ptrs = []
            if not errs:
                try:
                    for r, (_, raw, _e) in enumerate(everyone):
                        if r == self.rank:
                            ptrs.append(self._own)
                            continue
                        q = C.c_void_p()
                        with torch.cuda.device(self.device):
                            L.check(lib.sstg_peer_open(raw, C.byref(q)), f"sstg_peer_open(rank {r})")
                        self._opened.append(q.value)
                        ptrs.append(q.value)
                    torch.cuda.synchronize(self.device)
                except Exception as e:                               # noqa: BLE001
                    errs = [f"rank {self.rank}: {e}"]
            if self.world > 1:                   # second exchange doubles as the barrier: nobody stores into a
                opened = [None] * self.world     # mailbox that is not mapped and zeroed on every rank yet
                dist.all_gather_object(opened, errs, group=group)
                errs = [x for lst in opened for x in lst]
            if errs:
                self.close()
                raise RuntimeError("peer-memory exchange unavailable: " + "; ".join(sorted(set(errs))))
        self._boxes = (C.c_void_p * self.world)(*ptrs)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.stats = torch.zeros(max(self.max_scenes, 1), L.N_BN, dtype=torch.float32, device=self.device)

    @classmethod
    def local_group(cls, world: int, max_scenes: int, device, timeout_s: float = 30.0):
        """`world` exchanges that live in THIS process on one GPU (each rank's kernel goes on its own
        stream): the protocol, runnable on a single-GPU box."""
        import ctypes as C
        lib = L.load()
        device = torch.device(device)
        nbytes = int(lib.sstg_peer_mailbox_bytes(world, max_scenes))
        if nbytes < 0:
            L.check(nbytes, "sstg_peer_mailbox_bytes")
        ptrs = []
        for _ in range(world):
            q, h = C.c_void_p(), C.create_string_buffer(64)
            with torch.cuda.device(device):
                L.check(lib.sstg_peer_alloc(nbytes, C.byref(q), h), "sstg_peer_alloc")
            ptrs.append(q.value)
        out = [cls(device, max_scenes, None, timeout_s, _shared=(world, r, ptrs)) for r in range(world)]
        out[0]._owned_all = ptrs
        local = _LocalRanks(world, device)
        for pe in out:
            pe._local = local
        return out

    def exchange(self, engine, grad_local, loss_local, stats_local, scene_lo: int, s_local: int, seg_off_all,
                 s_all: int, lr: float, clip_norm: float = 0.0, bn_momentum: float = 0.1, stream=None):
        """Push this rank's contribution, wait for all, apply SGD + the BN recurrence to `engine`
        (params / bn / grad in place).  Asynchronous: `check()` reports a peer that never arrived."""
        if s_all > self.max_scenes:
            raise ValueError(f"group of {s_all} scenes exceeds the mailbox capacity {self.max_scenes}")
        self.epoch += 1
        ptr = lambda t: t.data_ptr() if isinstance(t, torch.Tensor) else (t or None)
        if self._emulated:
            # ranks that share this process and device: queue the arguments; the group goes out as ONE
            # cooperative launch when its last rank has called (or at flush() / check())
            one = L.PeerRank(self.rank, int(scene_lo), int(s_local), 0, ptr(grad_local), ptr(loss_local),
                             ptr(stats_local), ptr(seg_off_all), engine.params.data_ptr(), engine.bn.data_ptr(),
                             engine.grad.data_ptr(), self.loss.data_ptr(), self.stats.data_ptr(),
                             self.status.data_ptr())
            self._local.submit(self, one, (self.epoch, int(s_all), float(lr), float(clip_norm or 0.0),
                                           float(bn_momentum)),
                               keep=(grad_local, loss_local, stats_local, seg_off_all, engine))
            return self.loss[0]
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            L.check(L.load().sstg_peer_exchange_apply(
                self._boxes, self.world, self.rank, self.epoch, self.max_scenes, ptr(grad_local), ptr(loss_local),
                ptr(stats_local), int(scene_lo), int(s_local), ptr(seg_off_all), int(s_all), engine.params.data_ptr(),
                engine.bn.data_ptr(), float(lr), float(clip_norm or 0.0), float(bn_momentum), engine.grad.data_ptr(),
                self.loss.data_ptr(), self.stats.data_ptr(), self.status.data_ptr(), self.timeout_ns, stream),
                "sstg_peer_exchange_apply")
        return self.loss[0]

    def check(self, shared: bool = True):
        """Synchronise and raise if any exchange since the last check failed.  The status word is sticky: after a
        time-out this rank's exchange kernels refuse to run (no push, no update), so its peers time out too and
        NO replica trains on.  With `shared` (real multi-process groups) the status is max-reduced over the
        ranks first, so that every rank raises at the same check instead of the late one alone."""
        if self._emulated:
            self._local.flush()
            if self._local.done is not None:
                self._local.done.synchronize()        # the group ran on its own stream
        st = self.status
        if shared and not self._emulated and self.world > 1 and dist.is_available() and dist.is_initialized():
            st = st.clone()
            dist.all_reduce(st, op=dist.ReduceOp.MAX, group=self.group)
        st = int(st.item())
        if st:
            own = int(self.status.item())
            who = f"rank {own - 1} did not deliver to rank {self.rank}" if own else \
                f"an exchange failed on another rank (first late source there: rank {st - 1})"
            raise RuntimeError(f"peer exchange: {who} within {self.timeout_ns / 1e9:.1f} s; the replicas may have "
                               "diverged -- restart from the last checkpoint")

    def wait_previous(self):
        """Order the caller's current stream after the last exchange this rank took part in.  A no-op for real
        ranks (their exchange runs on the caller's stream); ranks emulated in one process are launched together
        on the group's stream."""
        if self._emulated:
            self._local.wait(torch.cuda.current_stream(self.device))

    def _close_quietly(self):
        try:
            self.close()
        except Exception:                                           # noqa: BLE001 -- interpreter / context teardown
            pass

    def __del__(self):
        self._close_quietly()

    def close(self):
        lib = L.load()
        if not self._opened and not self._own and not getattr(self, "_owned_all", None):
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for q in self._opened:
                lib.sstg_peer_close(q)
            self._opened = []
            for q in getattr(self, "_owned_all", []) or ([self._own] if self._own else []):
                lib.sstg_peer_free(q)
            self._own, self._owned_all = None, []


class _LocalRanks:
    """Several ranks of one exchange group inside ONE process on ONE device (PeerExchange.local_group): their
    exchange kernels wait on one another, so they are issued as ONE cooperative launch
    (sstg_peer_exchange_apply_local) -- separate launches are not guaranteed to be co-resident."""

    def __init__(self, world: int, device):
        self.world, self.device = int(world), torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.pending, self.common, self.done = {}, None, None

    def submit(self, pe, one, common, keep):
        if pe.rank in self.pending or (self.common is not None and common != self.common):
            self.flush()                              # a rank came round again before the others: issue what there is
        self.common = common
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.pending[pe.rank] = (pe, one, ev, keep)
        if len(self.pending) == self.world:
            self.flush()

    def flush(self):
        if not self.pending:
            return
        import ctypes as C
        items = [self.pending[r] for r in sorted(self.pending)]
        pe0 = items[0][0]
        arr = (L.PeerRank * len(items))(*[it[1] for it in items])
        epoch, s_all, lr, clip, mom = self.common
        for it in items:
            self.stream.wait_event(it[2])             # every queued rank's shard results are ready
        with torch.cuda.device(self.device):
            L.check(L.load().sstg_peer_exchange_apply_local(pe0._boxes, self.world, arr, len(items), epoch,
                                                            pe0.max_scenes, s_all, lr, clip, mom, pe0.timeout_ns,
                                                            self.stream.cuda_stream),
                    "sstg_peer_exchange_apply_local")
        self.done = torch.cuda.Event()
        self.done.record(self.stream)
        self._keep = [it[3] for it in items]          # until the next flush: the launch reads these buffers
        self.pending, self.common = {}, None

    def wait(self, stream):
        if self.done is not None:
            stream.wait_event(self.done)


class _ShardedGroup:
    """Host metadata of one optimiser group as this rank sees it (cached per pedestrian-count vector)."""

    def __init__(self, n_peds, batch_size, world, rank, dev, upload: bool = True):
        from . import engine as E
        self.n_peds = n_peds
        self.S_all = int(n_peds.size)
        self.lo, self.hi = shard_bounds(n_peds, world)[rank]
        seg = np.concatenate([[0], np.cumsum(n_peds)]).astype(np.int64)
        self.p_lo, self.p_hi = int(seg[self.lo]), int(seg[self.hi])
        self.w_all = E.group_weights(self.S_all, batch_size)
        self._seg32 = seg.astype(np.int32)
        self._w_local = np.ascontiguousarray(self.w_all[self.lo:self.hi], np.float32)
        self.full_plan = None
        self.plan = self.w_dev = self.seg_all_dev = None
        if upload:                      # one group at a time: three small host->device copies
            self.plan = E.ScenePlan(n_peds[self.lo:self.hi]) if self.hi > self.lo else None
            self.w_dev = torch.from_numpy(self._w_local).to(dev) if self.plan is not None else None
            self.seg_all_dev = torch.from_numpy(self._seg32).to(dev)

    @staticmethod
    def upload_many(groups, dev):
        """Device side of many groups (an epoch) in TWO host->device copies: the local plans' offset tables
        (engine.plan_groups) and one arena holding every group's global seg_off and local loss weights."""
        from . import engine as E
        todo = [g for g in groups if g.seg_all_dev is None]
        if not todo:
            return
        local = [g for g in todo if g.hi > g.lo]
        for g, pl in zip(local, E.plan_groups([g.n_peds[g.lo:g.hi] for g in local], dev)):
            g.plan = pl
        offs, total = [], 0
        for g in todo:
            o_s = total
            o_w = o_s + ((g._seg32.size + 3) & ~3)
            total = o_w + ((g._w_local.size + 3) & ~3)
            offs.append((o_s, o_w))
        host = np.zeros(max(total, 4), np.int32)
        for g, (o_s, o_w) in zip(todo, offs):
            host[o_s:o_s + g._seg32.size] = g._seg32
            host[o_w:o_w + g._w_local.size] = g._w_local.view(np.int32)
        buf = torch.from_numpy(host).to(dev)
        for g, (o_s, o_w) in zip(todo, offs):
            g.seg_all_dev = buf[o_s:o_s + g._seg32.size]
            g.w_dev = buf[o_w:o_w + g._w_local.size].view(torch.float32) if g.hi > g.lo else None


class DistributedTrainer:
    """One optimiser group of train() (train.py:28-79) spread over the ranks of `group`.

    Every rank holds identical weights and BN buffers (a TrainEngine).  `step` takes the
    WHOLE group (all ranks pass the same n_peds / tensors, or each its own slice with
    `presliced=True`), runs graph build + train-mode forward + backward on this rank's
    contiguous shard with the group's loss weights, exchanges gradient / loss / BN statistics and
    applies the identical SGD step and BN recurrence everywhere.

    exchange = "peer": libsstg.so's one-kernel exchange over NVLink peer memory (PeerExchange), fused
                       with the SGD step -- two library calls per group, no torch op in between;
               "nccl": one torch.distributed all-reduce of [gradient | loss | statistics rows];
               "auto": "peer" when every rank is on CUDA on one node and world <= SSTG_MAX_RANKS (8), else
                       "nccl"; the environment variable SSTG_EXCHANGE=nccl|peer decides for "auto" only -- an
                       explicit "peer" / "nccl" argument is never overridden."""

    def __init__(self, engine, group=None, exchange: str = "auto", max_scenes: int = 256, *,
                 peer_timeout_s: float = 30.0, _emulate=None):
        import os
        self.engine, self.group = engine, group
        self._flat, self._groups = {}, {}
        self._wr = None
        if _emulate is not None:       # a PeerExchange.local_group member: several ranks inside one process (tests)
            self.peer, self.exchange = _emulate, "peer"
            self._wr = (_emulate.world, _emulate.rank)
            self.world, self.rank = self._wr
            return
        world, rank = _world(group)
        self.world, self.rank = world, rank
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be auto, peer or nccl")
        mode = exchange
        if exchange == "auto" and os.environ.get("SSTG_EXCHANGE"):
            mode = os.environ["SSTG_EXCHANGE"]
            if mode not in ("auto", "peer", "nccl"):
                raise ValueError("SSTG_EXCHANGE must be auto, peer or nccl")
        if mode == "peer" and world > L.MAX_RANKS:
            raise ValueError(f"exchange='peer' handles at most SSTG_MAX_RANKS = {L.MAX_RANKS} ranks on one node "
                             f"(world size is {world}); use exchange='nccl'")
        self.peer = None
        if 1 < world <= L.MAX_RANKS and (mode == "peer" or (mode == "auto" and engine.device.type == "cuda")):
            try:
                self.peer = PeerExchange(engine.device, max_scenes, group, timeout_s=peer_timeout_s)
                ok = 1
            except Exception as e:
                if mode == "peer":
                    raise
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({e}); using the NCCL all-reduce")
                ok = 0
            t = torch.tensor([ok], dtype=torch.int32, device=engine.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            if int(t[0]) == 0 and self.peer is not None:
                self.peer.close()
                self.peer = None
        self.exchange = "peer" if self.peer is not None else ("nccl" if world > 1 else "none")

    def check(self):
        """Raise on every rank if a peer-memory exchange since the last check failed on any rank (no-op for
        the NCCL exchange, whose collectives fail by themselves)."""
        if self.peer is not None:
            self.peer.check(shared=self._wr is None)

    def prepare(self, n_peds_list, batch_size: int):
        """Host preparation of many groups at once (driver.train_epoch calls it with the epoch's groups): their
        offset tables, global segment offsets and loss weights reach the device in two copies instead of
        three per group."""
        world, rank = self._wr or _world(self.group)
        new = []
        for n in n_peds_list:
            n = np.ascontiguousarray(np.asarray(n, np.int32).reshape(-1))
            key = (n.tobytes(), batch_size)
            if key not in self._groups:
                if len(self._groups) >= 4096:
                    self._groups.clear()
                g = self._groups[key] = _ShardedGroup(n, batch_size, world, rank, self.engine.device, upload=False)
                new.append(g)
        _ShardedGroup.upload_many(new, self.engine.device)

    def _group(self, n_peds, batch_size):
        key = (n_peds.tobytes(), batch_size)
        g = self._groups.get(key)
        if g is None:
            if len(self._groups) >= 4096:
                self._groups.clear()
            world, rank = self._wr or _world(self.group)
            g = self._groups[key] = _ShardedGroup(n_peds, batch_size, world, rank, self.engine.device)
        return g

    def step(self, n_peds, obs_rel, pred_rel, batch_size: int, lr: float, clip_grad: Optional[float] = None,
             presliced: bool = False, graph_fp64: bool = False):
        from . import engine as E
        world, rank = self._wr or _world(self.group)
        n_peds = np.ascontiguousarray(np.asarray(n_peds, np.int32).reshape(-1))
        g = self._group(n_peds, batch_size)
        lo, hi = g.lo, g.hi
        eng, dev = self.engine, self.engine.device
        o = p = None
        if hi > lo:
            if presliced:
                o, p = obs_rel, pred_rel
            else:
                o = torch.as_tensor(obs_rel)[g.p_lo:g.p_hi].to(dev).contiguous()
                p = torch.as_tensor(pred_rel)[g.p_lo:g.p_hi].to(dev).contiguous()
        if self.peer is not None:
            # library call 1: this shard's graph build + forward + backward (no update);
            # library call 2: exchange over peer memory + SGD + BN recurrence in global scene order
            self.peer.wait_previous()
            grad_l, loss_l, stats_l = eng.shard_forward_backward(g.plan, o, p, g.w_dev, graph_fp64=graph_fp64)
            out = self.peer.exchange(eng, grad_l, loss_l, stats_l, lo, hi - lo, g.seg_all_dev, g.S_all, lr,
                                     clip_grad or 0.0)
            eng.num_batches_tracked += g.S_all
            return out
        loss_local = torch.zeros(1, dtype=torch.float64, device=dev)
        if hi > lo:
            r = eng.forward_backward(g.plan, o, p, g.w_dev, graph_fp64=graph_fp64)
            stats = r["scene_stats"]
            loss_local += (r["loss"].double() * r["weight"].double()).sum()
        else:
            eng.grad.zero_()
            stats = torch.zeros(0, L.N_BN, dtype=torch.float32, device=dev)
        # ONE collective per optimiser group: [gradient 7563 | group loss 1 | per-scene BN statistics S x 30]
        # in a single flat buffer, all-reduced (sum).  Every rank fills only its own scenes' rows of the
        # statistics block, so the sum IS the concatenation in global scene order.
        S_all = g.S_all
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
        if g.full_plan is None:
            g.full_plan = E.ScenePlan(n_peds)
        E.bn_running_update(g.full_plan.on(dev), stats_all.contiguous(), eng.bn)
        eng.num_batches_tracked += S_all
        E.sgd_step(eng.params, eng.grad, lr, clip_grad or 0.0)
        return loss_local[0]
