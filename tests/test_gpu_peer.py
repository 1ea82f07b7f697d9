"""GPU tests of the peer-memory gradient exchange (sstg_peer_exchange_apply, include/sstg.h): the
scene-sharded optimiser group must reproduce the single-GPU group (TrainEngine.step).

On a one-GPU box the ranks are emulated inside one process (PeerExchange.local_group: one mailbox per rank,
the same kernel and protocol; because the ranks' CTAs wait on one another they are issued as ONE cooperative
launch, sstg_peer_exchange_apply_local -- never as separate launches that may not be co-resident); with
>= 2 GPUs a second test runs real processes over CUDA IPC + NVLink and compares with the NCCL exchange.

Tolerances: the sharded gradient is the same per-scene terms summed in a different order (per-rank
partial sums, then rank order), so weights after a step agree to fp32 rounding of the gradient
(1e-5 of the largest entry, scaled by lr); the replicas among themselves must be BIT-identical."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def _groups(n_groups, scenes, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_groups):
        n = rng.integers(2, 40, scenes).astype(np.int32)
        P = int(n.sum())
        obs = (rng.standard_normal((P, 2, 8)) * 0.3).astype(np.float32)
        pred = (rng.standard_normal((P, 2, 12)) * 0.3).astype(np.float32)
        out.append((n, obs, pred))
    return out


@pytest.mark.parametrize("world", [2, 3, 8])
def test_emulated_ranks_match_single_gpu(golden, dev, world):
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import parallel as Pl
    params, bn = golden.params("eth")
    scenes, lr = 12, 1e-3
    groups = _groups(4, scenes, 5 + world)            # 4 groups: both payload parities are reused
    single = sstg.TrainEngine(params, bn, dev)
    peers = Pl.PeerExchange.local_group(world, 16, dev)
    engines = [sstg.TrainEngine(params, bn, dev) for _ in range(world)]
    trainers = [Pl.DistributedTrainer(e, _emulate=p) for e, p in zip(engines, peers)]
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    # allocate every engine's scratch up front: nothing below may need a cudaMalloc while a rank's
    # kernel is already waiting for the others
    for n, obs, pred in groups:
        o, p = torch.from_numpy(obs).to(dev), torch.from_numpy(pred).to(dev)
        for tr in trainers:
            g = tr._group(np.ascontiguousarray(n), scenes)
            if g.plan is not None:
                tr.engine.shard_forward_backward(g.plan, o[g.p_lo:g.p_hi], p[g.p_lo:g.p_hi], g.w_dev)
    torch.cuda.synchronize()
    try:
        for n, obs, pred in groups:
            o, p = torch.from_numpy(obs).to(dev), torch.from_numpy(pred).to(dev)
            ref_loss = single.step(sstg.ScenePlan(n), o, p, scenes, lr, clip_grad=10.0)
            torch.cuda.synchronize()
            losses = []
            for r in range(world):                       # every rank's work is enqueued before anyone is waited on
                with torch.cuda.stream(streams[r]):
                    losses.append(trainers[r].step(n, o, p, scenes, lr, clip_grad=10.0))
            torch.cuda.synchronize()
            for pe in peers:
                pe.check()
            for r in range(world):
                assert abs(float(losses[r]) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss)) + 1e-6
            g_ref = single.grad.cpu().numpy()
            scale = float(np.abs(g_ref).max())
            for r in range(world):
                assert np.abs(engines[r].grad.cpu().numpy() - g_ref).max() <= 1e-5 * scale
                assert torch.equal(engines[r].params, engines[0].params)          # replicas: same bits
                assert torch.equal(engines[r].bn, engines[0].bn)
                assert torch.equal(engines[r].grad, engines[0].grad)
            assert np.allclose(engines[0].bn.cpu().numpy(), single.bn.cpu().numpy(), rtol=1e-5, atol=1e-7)
            assert np.abs(engines[0].params.cpu().numpy() - single.params.cpu().numpy()).max() <= 2e-5 * scale * lr + 1e-9
        assert engines[0].num_batches_tracked == single.num_batches_tracked == 4 * scenes
    finally:
        peers[0].close()


def test_missing_peer_times_out_and_leaves_weights_alone(golden, dev):
    """A rank that never shows up: the kernel gives up after the timeout, reports the missing source in
    the status word and applies nothing (no unbounded spin on the GPU)."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import parallel as Pl
    params, bn = golden.params("eth")
    peers = Pl.PeerExchange.local_group(2, 16, dev, timeout_s=0.05)
    try:
        eng = sstg.TrainEngine(params, bn, dev)
        tr = Pl.DistributedTrainer(eng, _emulate=peers[0])
        n, obs, pred = _groups(1, 6, 3)[0]
        tr.step(n, torch.from_numpy(obs).to(dev), torch.from_numpy(pred).to(dev), 6, 1e-3)
        with pytest.raises(RuntimeError, match="rank 1 did not deliver"):
            peers[0].check()
        assert np.array_equal(eng.params.cpu().numpy(), params.astype(np.float32).reshape(-1))
        assert np.array_equal(eng.bn.cpu().numpy(), bn.astype(np.float32).reshape(-1))
        # the status word is sticky: the failed rank takes no part in later exchanges (returns at once, pushes
        # nothing, applies nothing) until the caller clears it
        flags_before = peers[0].epoch
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        tr.step(n, torch.from_numpy(obs).to(dev), torch.from_numpy(pred).to(dev), 6, 1e-3)
        with pytest.raises(RuntimeError, match="rank 1 did not deliver"):
            tr.check()
        t1.record()
        torch.cuda.synchronize()
        assert t0.elapsed_time(t1) < 45.0                                  # no second 50 ms wait
        assert peers[0].epoch == flags_before + 1
        assert np.array_equal(eng.params.cpu().numpy(), params.astype(np.float32).reshape(-1))
    finally:
        peers[0].close()


def test_mailbox_capacity_is_checked(golden, dev):
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import parallel as Pl
    params, bn = golden.params("eth")
    peers = Pl.PeerExchange.local_group(1, 4, dev)
    try:
        tr = Pl.DistributedTrainer(sstg.TrainEngine(params, bn, dev), _emulate=peers[0])
        n, obs, pred = _groups(1, 6, 4)[0]
        with pytest.raises(ValueError, match="mailbox capacity"):
            tr.step(n, torch.from_numpy(obs).to(dev), torch.from_numpy(pred).to(dev), 6, 1e-3)
    finally:
        peers[0].close()


# ---- real processes, real peer memory (needs >= 2 GPUs; the round-end one-GPU box skips it) ----------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, params, bn, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import social_stgcnn_b200 as sstg
        from social_stgcnn_b200 import parallel as Pl
        scenes, lr = 16, 1e-3
        groups = _groups(5, scenes, 11)
        e_peer, e_nccl, e_one = (sstg.TrainEngine(params, bn, dev) for _ in range(3))
        t_peer = Pl.DistributedTrainer(e_peer, exchange="peer", max_scenes=scenes)
        t_nccl = Pl.DistributedTrainer(e_nccl, exchange="nccl")
        assert t_peer.exchange == "peer" and t_nccl.exchange == "nccl"
        for n, obs, pred in groups:
            o, p = torch.from_numpy(obs).to(dev), torch.from_numpy(pred).to(dev)
            l1 = t_peer.step(n, o, p, scenes, lr)
            l2 = t_nccl.step(n, o, p, scenes, lr)
            l0 = e_one.step(sstg.ScenePlan(n), o, p, scenes, lr)
            t_peer.peer.check()
            assert abs(float(l1) - float(l0)) <= 1e-5 * abs(float(l0)) + 1e-6, (float(l1), float(l0))
            assert abs(float(l2) - float(l0)) <= 1e-5 * abs(float(l0)) + 1e-6
        scale = float(e_one.grad.abs().max())
        assert float((e_peer.grad - e_one.grad).abs().max()) <= 1e-5 * scale
        assert float((e_peer.params - e_one.params).abs().max()) <= 1e-4 * scale * lr + 1e-9
        assert torch.allclose(e_peer.bn, e_one.bn, rtol=1e-5, atol=1e-7)
        assert torch.allclose(e_peer.bn, e_nccl.bn, rtol=1e-6, atol=1e-8)
        blob = torch.cat([e_peer.params, e_peer.bn])
        both = [torch.empty_like(blob) for _ in range(world)]
        dist.all_gather(both, blob)
        assert all(torch.equal(both[0], b) for b in both)                       # replicas: same bits
        # ---- a whole shuffled epoch through the driver (batched host preparation, DistributedTrainer.prepare)
        from social_stgcnn_b200 import driver
        n_all = np.concatenate([g[0] for g in groups])
        obs_all = torch.from_numpy(np.concatenate([g[1] for g in groups])).to(dev)
        pred_all = torch.from_numpy(np.concatenate([g[2] for g in groups])).to(dev)

        class Split:
            packed = {"device": dev, "obs_rel": obs_all, "pred_rel": pred_all}
            plan_obs = sstg.ScenePlan(n_all)

            def __len__(self):
                return int(n_all.size)

        order = np.random.default_rng(4).permutation(int(n_all.size))
        la = driver.train_epoch(Split(), e_peer, scenes, lr, order=order, trainer=t_peer)
        lb = driver.train_epoch(Split(), e_one, scenes, lr, order=order)
        assert abs(la - lb) <= 1e-5 * abs(lb) + 1e-7, (la, lb)
        assert float((e_peer.params - e_one.params).abs().max()) <= 2e-4 * scale * lr + 1e-9
        assert torch.allclose(e_peer.bn, e_one.bn, rtol=1e-5, atol=1e-7)
        t_peer.peer.close()
        ret[rank] = "ok"
    except Exception as e:
        import traceback
        ret[rank] = "".join(traceback.format_exception(e))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_real_processes_over_nvlink(golden):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    params, bn = golden.params("eth")
    world = min(torch.cuda.device_count(), 8)          # every GPU of the box: 2, 4 or 8 ranks
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_rank_main, args=(world, _free_port(), params, bn, ret), nprocs=world, join=True)
        assert dict(ret) == {r: "ok" for r in range(world)}, dict(ret)
