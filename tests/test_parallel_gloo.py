"""CPU tests of the multi-GPU host logic with world_size 2 over gloo (no GPU, no kernels):
scene sharding, the score reduction, the gradient all-reduce, the BatchNorm-statistics
gather in global scene order, and the DistributedTrainer orchestration with the device work
replaced by deterministic stand-ins that live HERE (test infrastructure, not product code)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_and_balance():
    from social_stgcnn_b200.parallel import scene_cost, shard_bounds, shard_slice
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for n in ([64] * 1000, rng.integers(2, 58, 947).tolist(), [2, 3], [5], []):
            b = shard_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == len(n)
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))        # contiguous, disjoint, complete
            assert all(lo <= hi for lo, hi in b)
            if len(n) >= 100:
                c = scene_cost(n)
                loads = [c[lo:hi].sum() for lo, hi in b]
                assert max(loads) <= 1.05 * (sum(loads) / world) + c.max()
    lo, hi, p_lo, p_hi = shard_slice([3, 4, 5, 6], 2, 1)
    assert (lo, hi) == (2, 4) or (lo, hi) == (3, 4)
    assert p_hi == 18 and p_lo == sum([3, 4, 5, 6][:lo])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bn_recurrence(bn, stats, n_peds, momentum=0.1):
    """float64 stand-in for sstg_bn_running_update (the CUDA kernel is tested on the GPU)."""
    bn = bn.clone().double()
    for s in range(stats.shape[0]):
        M = 8.0 * n_peds[s]
        for i in range(30):
            v = float(stats[s, i])
            if (i // 5) % 2 == 1:
                v = v * M / (M - 1.0)
            bn[i] = momentum * v + (1 - momentum) * bn[i]
    return bn.float()


class _FakePlan:
    def __init__(self, n_peds, T=8):
        self.n_peds = np.asarray(n_peds, np.int32)
        self.S = int(self.n_peds.size)

    def on(self, dev):
        return self


class _FakeEngine:
    """Stands in for TrainEngine: 'gradient' and 'statistics' are simple functions of the scene
    ids, so the expected global result is known in closed form."""

    def __init__(self):
        self.device = torch.device("cpu")
        self.params = torch.arange(7563, dtype=torch.float32) * 1e-3
        self.bn = torch.ones(30)
        self.grad = torch.zeros(7563)
        self.num_batches_tracked = 0
        self.seen = []

    def forward_backward(self, plan, obs_rel, pred_rel, weight, **kw):
        ids = obs_rel[:, 0, 0].long()                      # pedestrian rows carry their global scene id
        first = torch.unique_consecutive(ids)
        self.seen = first.tolist()
        w = torch.as_tensor(weight)
        assert len(first) == plan.S == len(w)
        self.grad.zero_()
        for sid, wi in zip(first.tolist(), w.tolist()):
            self.grad += wi * (sid + 1)
        stats = torch.stack([torch.full((30,), float(s)) + torch.arange(30) * 0.01 for s in first.tolist()])
        loss = first.double().float() * 0.5
        return {"loss": loss, "grad": self.grad, "scene_stats": stats, "weight": w, "V_pred": None}


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from social_stgcnn_b200 import engine as E
        from social_stgcnn_b200 import parallel as Pl
        # ---- score reduction
        a, f, n = Pl.reduce_scores(1.5 * (rank + 1), 3.0 * (rank + 1), 10 * (rank + 1))
        assert n == 30 and abs(a - 4.5 / 30) < 1e-15 and abs(f - 9.0 / 30) < 1e-15
        # ---- gradient all-reduce
        g = torch.full((7563,), float(rank + 1))
        Pl.allreduce_gradient(g)
        assert torch.all(g == 3.0)
        # ---- ragged stats gather keeps the global scene order
        counts = [3, 1]
        mine = torch.arange(counts[rank] * 30, dtype=torch.float32).view(counts[rank], 30) + 1000 * rank
        allst = Pl.gather_scene_stats(mine, counts)
        assert allst.shape == (4, 30) and allst[0, 0] == 0 and allst[2, 29] == 89 and allst[3, 0] == 1000
        # ---- DistributedTrainer orchestration with stand-ins for the device work
        n_peds = [4, 2, 9, 3, 5, 7, 2, 6]
        seg = np.concatenate([[0], np.cumsum(n_peds)])
        obs = torch.zeros(int(seg[-1]), 2, 8)
        for s in range(len(n_peds)):
            obs[seg[s]:seg[s + 1], 0, 0] = s
        eng = _FakeEngine()
        updates = {}

        def fake_bn(dp, stats, bn, momentum=0.1):
            updates["stats"] = stats.clone()
            bn.copy_(_bn_recurrence(bn, stats, dp.n_peds, momentum))

        def fake_sgd(params, grad, lr, clip=0.0, out=None):
            params.sub_(lr * grad)

        E_bn, E_sgd, E_plan = E.bn_running_update, E.sgd_step, E.ScenePlan
        E.bn_running_update, E.sgd_step, E.ScenePlan = fake_bn, fake_sgd, _FakePlan
        try:
            tr = Pl.DistributedTrainer(eng)
            loss = tr.step(n_peds, obs, obs.clone(), batch_size=8, lr=0.5)
        finally:
            E.bn_running_update, E.sgd_step, E.ScenePlan = E_bn, E_sgd, E_plan
        bounds = Pl.shard_bounds(n_peds, world)
        assert eng.seen == list(range(*bounds[rank]))                       # this rank forwarded its shard only
        w = E.group_weights(8, 8)
        exp_grad = sum(float(w[s]) * (s + 1) for s in range(8))             # all scenes, last one unscored
        assert torch.allclose(eng.grad, torch.full((7563,), exp_grad))
        assert torch.allclose(eng.params, torch.arange(7563, dtype=torch.float32) * 1e-3 - 0.5 * exp_grad)
        exp_loss = sum(float(w[s]) * 0.5 * s for s in range(8))
        assert abs(float(loss) - exp_loss) < 1e-6
        assert updates["stats"].shape == (8, 30)
        assert torch.equal(updates["stats"][:, 0], torch.arange(8, dtype=torch.float32))   # global order
        assert eng.num_batches_tracked == 8
        assert tr.world == world and tr.rank == rank and tr.exchange == "nccl"
        tr.check()                                                         # no peer exchange: nothing to report
        # ---- the epoch shuffle: rank 0's permutation reaches every rank; a rank with its own order is caught
        mine = np.random.default_rng(100 + rank).permutation(37)           # what an unseeded fit() would draw
        order = Pl.broadcast_order(mine)
        assert np.array_equal(order, np.random.default_rng(100).permutation(37))
        Pl.assert_same_across_ranks((order, np.arange(5)))                 # agree: passes (and is a barrier)
        try:
            Pl.assert_same_across_ranks((mine,), what="epoch order")       # disagree: EVERY rank raises
            raise AssertionError("mismatch not detected")
        except ValueError as e:
            assert "epoch order differ between ranks" in str(e)
        # both replicas end bit-identical
        blob = torch.cat([eng.params, eng.bn])
        both = [torch.empty_like(blob) for _ in range(world)]
        dist.all_gather(both, blob)
        assert torch.equal(both[0], both[1])
        ret[rank] = "ok"
    except Exception as e:                                                  # surfaced by the parent
        import traceback
        ret[rank] = "".join(traceback.format_exception(e))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world2_gloo_host_logic():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)
