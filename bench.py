#!/usr/bin/env python
"""bench.py -- headline benchmark of the Social-STGCNN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): scenes/sec of  graph build + ST-GCN/TXP-CNN forward + 20-sample
ADE/FDE  on synthetic dense crowds (64 pedestrians per scene, obs 8 -> pred 12), sharded by
scene across the GPUs with no data-path collective (weak scaling: a fixed shard per GPU).
One "step" = one pass of the whole hot path over this rank's resident batch of scenes.

The JSON line carries: `value` (device-resident inputs), `e2e` (host buffers, H2D + D2H in
the timed region, through InferenceEngine.infer_host), `roofline` for the st_gcn kernel (the
kernel BASELINE.json states the HBM target on) plus `kernels` with every kernel's share,
`cpu_baseline` (the oracle port timed on the host cores, rank 0 / N=1 only), `clocks`.

`--impl reference` times the reference's CPU algorithm (oracle port -- the Python reference
itself cannot travel to the GPU box) with all host cores on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scenes_per_sec_graph_fwd_ade_fde"
UNIT = "scenes/s"
N_PEDS = 64
K_SAMPLES = 20


def synth_batch(num_scenes, n_peds, seed, compact=False):
    """Synthetic crowds of SURVEY 8(d): random starts in (-10,10)^2, per-pedestrian velocity
    U(-.5,.5)^2 + N(0, .1^2) per frame, positions rounded to 4 decimals exactly as the dataset does
    (utils.py:145: rint(x*1e4)/1e4 in float64), rel = first difference in float64 cast to fp32 with
    rel[:, :, 0] = 0 (utils.py:153-155,180-183).  Returns CPU tensors (obs_rel [P,2,8],
    pred_rel [P,2,12], last_obs [P,2]) and, with compact=True, also the same tracks in the
    fixed-point host format (start int32 [P,2], delta int16 [P,2,20])."""
    import torch
    g = torch.Generator().manual_seed(seed)
    P = num_scenes * n_peds
    start = torch.rand(P, 2, generator=g) * 20 - 10
    vel = torch.rand(P, 2, generator=g) - 0.5
    steps = vel[:, :, None] + 0.1 * torch.randn(P, 2, 20, generator=g)
    steps[:, :, 0] = 0
    k = torch.round((start[:, :, None] + torch.cumsum(steps, 2)).double() * 1e4)       # integer positions, units 1e-4
    pos = k / 1e4
    rel = torch.zeros_like(pos)
    rel[:, :, 1:] = pos[:, :, 1:] - pos[:, :, :-1]
    out = (rel[:, :, :8].float().contiguous(), rel[:, :, 8:].float().contiguous(), pos[:, :, 7].float().contiguous())
    if compact:
        d = torch.zeros_like(k)
        d[:, :, 1:] = k[:, :, 1:] - k[:, :, :-1]
        out = out + (k[:, :, 0].to(torch.int32).contiguous(), d.to(torch.int16).contiguous())
    return out


def workload_name(scenes_per_gpu):
    return (f"synthetic dense crowds: {scenes_per_gpu} scenes/GPU x {N_PEDS} peds (config 3: 1M scenes at 8 GPUs with "
            f"the default shard), obs 8 -> pred 12, K={K_SAMPLES} Philox samples, eth checkpoint weights")


def bench_config(scenes_per_gpu, world):
    """`config` of the JSON line -- built by ONE function for both arms, so that the reference arm reports
    exactly the configuration of ours (what differs between the arms, e.g. the bounded CPU sample of the
    reference arm, is reported next to it, not inside it)."""
    a_gb = 32.0 * N_PEDS * N_PEDS * scenes_per_gpu / 1e9
    return {"workload": workload_name(scenes_per_gpu), "scenes_per_gpu": scenes_per_gpu, "peds_per_scene": N_PEDS,
            "ksteps": K_SAMPLES, "graph_mode": "fp32",
            "l2_policy": "inputs larger than L2 (A is %.1f GB per step)" % a_gb,
            "parallelism": f"scene-sharded x{world}, no data-path collective"}


def load_weights():
    w = np.load(os.path.join(ROOT, "tests", "golden", "weights.npz"))
    return w["eth.params"], w["eth.bn"]


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""
    BITS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
            0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
            0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        return {"sm_mhz": (float(np.median(self.sm)) if self.sm else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def profiled_traffic():
    """DRAM bytes of the st_gcn kernel from the committed ncu --set full capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum for one launch of `scenes` scenes), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))["stgcn_fwd_kernel"]
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------
# CPU legs: the oracle port (the checker) timed on host cores
# ---------------------------------------------------------------------------------
def _cpu_worker(args):
    import torch
    torch.set_num_threads(1)
    from oracle import stgcnn_oracle as O
    params, bn, obs_rel, pred_rel, last, n_peds, seed = args
    P_, S_ = O.unflatten_params(params), O.unflatten_bn(bn)
    g = torch.Generator().manual_seed(seed)
    n_sc = obs_rel.shape[0] // n_peds
    tot = 0.0
    for i in range(n_sc):
        sl = slice(i * n_peds, (i + 1) * n_peds)
        eps = torch.randn(K_SAMPLES, 12, n_peds, 2, generator=g).numpy()       # the reference draws its own noise
        a, f = O.infer_scene(P_, S_, obs_rel[sl], pred_rel[sl], last[sl], eps)
        tot += float(a.sum() + f.sum())
    return n_sc, tot


def cpu_port_throughput(n_scenes, procs, seed=0):
    """scenes/s of the oracle port (graph build + eval forward + 20-sample ADE/FDE, batch 1
    per scene as the reference does) using `procs` host processes."""
    import multiprocessing as mp
    params, bn = load_weights()
    obs_rel, pred_rel, last = [t.numpy() for t in synth_batch(n_scenes, N_PEDS, 4321 + seed)]
    per = max(1, n_scenes // procs)
    jobs = []
    for w in range(procs):
        lo, hi = w * per * N_PEDS, (w + 1) * per * N_PEDS
        if lo >= obs_rel.shape[0]:
            break
        jobs.append((params, bn, obs_rel[lo:hi], pred_rel[lo:hi], last[lo:hi], N_PEDS, seed + w))
    t0 = time.perf_counter()
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            res = pool.map(_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    done = sum(r[0] for r in res)
    return done / dt, done, dt


def cpu_port_phases(n_scenes=32, seed=0):
    """Single-process time of the oracle port's three phases on `n_scenes` scenes of the workload (SURVEY 8d asks for
    graph build and forward / evaluation separately): seconds per scene of seq_to_graph, of the eval forward and of the
    20-sample ADE/FDE."""
    import torch
    torch.set_num_threads(1)
    from oracle import stgcnn_oracle as O
    params, bn = load_weights()
    P_, S_ = O.unflatten_params(params), O.unflatten_bn(bn)
    obs_rel, pred_rel, last = [t.numpy() for t in synth_batch(n_scenes, N_PEDS, 8765 + seed)]
    g = torch.Generator().manual_seed(seed)
    tg = tf = ts = 0.0
    for i in range(n_scenes):
        sl = slice(i * N_PEDS, (i + 1) * N_PEDS)
        eps = torch.randn(K_SAMPLES, 12, N_PEDS, 2, generator=g).numpy()
        t0 = time.perf_counter()
        V_obs, A_obs = O.seq_to_graph(obs_rel[sl])
        t1 = time.perf_counter()
        with torch.no_grad():
            out = O.model_forward(P_, S_, V_obs, A_obs)
        V_pred = O.vpred_tn5(out).numpy()
        t2 = time.perf_counter()
        O.sample_ade_fde(V_pred, np.ascontiguousarray(pred_rel[sl].transpose(2, 0, 1)), last[sl], eps)
        t3 = time.perf_counter()
        tg, tf, ts = tg + (t1 - t0), tf + (t2 - t1), ts + (t3 - t2)
    return {"graph_build_ms_per_scene": 1e3 * tg / n_scenes, "forward_ms_per_scene": 1e3 * tf / n_scenes,
            "sample_ade_fde_ms_per_scene": 1e3 * ts / n_scenes, "scenes": n_scenes, "threads": 1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 256 * cores          # ~2-3 s of CPU work per step on all cores
    for _ in range(args.warmup):
        cpu_port_throughput(cores, cores)
    t0 = time.perf_counter()
    done = 0
    for s in range(args.steps):
        _, d, _ = cpu_port_throughput(per_step, cores, seed=s)
        done += d
    dt = time.perf_counter() - t0
    v = done / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.scenes_per_gpu, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} scenes of this workload per step (bounded CPU sample): {done} scenes "
                                   f"in {dt:.1f}s, oracle port (vectorised numpy graph build + torch CPU fp32 forward "
                                   f"+ numpy sampling), one process per core"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------
# graph-construction leg (BASELINE.json config 5): seq_to_graph on packed tiny scenes
# ---------------------------------------------------------------------------------
def run_graph_leg(dev, steps=5, num_scenes=262144):
    """scenes/s of TrajectoryDataset's graph construction (utils.py:201-211: seq_to_graph of the
    observed half, T = 8, AND of the predicted half, T = 12, normalised Laplacian) on packed ragged
    scenes shaped like the nuScenes-mini / ETH-UCY windows (2..5 pedestrians), in the fp64-island
    mode that is bit-identical to networkx, and in the fp32 mode."""
    import torch
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E
    rng = np.random.default_rng(5)
    n = rng.integers(2, 6, num_scenes).astype(np.int32)
    obs, pred, _ = synth_batch(int(n.sum()), 1, 99)
    obs, pred = obs.to(dev), pred.to(dev)
    p8, p12 = sstg.ScenePlan(n, T=8), sstg.ScenePlan(n, T=12)
    d8, d12 = p8.on(dev), p12.on(dev)
    V8 = torch.empty(p8.P * 16, device=dev); A8 = torch.empty(p8.a_floats, device=dev)
    V12 = torch.empty(p12.P * 24, device=dev); A12 = torch.empty(p12.a_floats, device=dev)
    out = {"scenes": num_scenes, "peds_per_scene": "2..5", "what": "seq_to_graph of both halves (T=8 and T=12) per scene"}
    for mode, fp64 in (("fp64_exact", True), ("fp32", False)):
        def one():
            E.graph_build(d8, obs, 8, fp64=fp64, V=V8, A=A8)
            E.graph_build(d12, pred, 12, fp64=fp64, V=V12, A=A12)
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        out[mode + "_scenes_per_s"] = num_scenes * steps / (e0.elapsed_time(e1) * 1e-3)
    return out


# ---------------------------------------------------------------------------------
# ragged inference leg (BASELINE.json config 1 shapes): real-data-sized scenes, packed
# ---------------------------------------------------------------------------------
def run_ragged_leg(dev, steps=5, num_scenes=65536):
    """scenes/s of graph build + forward + 20-sample ADE/FDE on a packed batch of ETH/UCY-shaped scenes
    (2..57 pedestrians, median ~6): the shape of the reference's own test.py workload, where it runs one
    scene per iteration."""
    import torch
    import social_stgcnn_b200 as sstg
    params, bn = load_weights()
    eng = sstg.InferenceEngine(params, bn, dev)
    n = synth_ragged_counts(num_scenes, 17)
    obs, pred, last = synth_batch(int(n.sum()), 1, 4242)
    obs, pred, last = obs.to(dev), pred.to(dev), last.to(dev)
    plan = sstg.ScenePlan(n)
    out = {"scenes": num_scenes, "pedestrians": int(n.sum()), "peds_per_scene": "2..57, log-normal, median 6"}
    for name, fused in (("scenes_per_s", False), ("fused_scenes_per_s", True)):
        for _ in range(3):
            eng.infer(plan, obs, pred, last, K_SAMPLES, seed=1, want_vtr=False, fused=fused)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            eng.infer(plan, obs, pred, last, K_SAMPLES, seed=1, want_vtr=False, fused=fused)
        e1.record()
        torch.cuda.synchronize()
        out[name] = num_scenes * steps / (e0.elapsed_time(e1) * 1e-3)
    return out


# ---------------------------------------------------------------------------------
# ETH test split through the drop-in API (BASELINE.json config 1): the reference's own CPU-runnable case
# ---------------------------------------------------------------------------------
def run_eth_leg(dev):
    """test.py's job on the shipped eth/test file (70 scenes, 181 pedestrians) with the eth checkpoint:
    TrajectoryDataset (file parsing + windowing + graph build) and driver.evaluate (forward + KSTEPS=20
    sampling + ADE/FDE with the reference's fp64 rounding islands), wall clock incl. every copy."""
    import time
    import torch
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import driver
    from social_stgcnn_b200.utils import TrajectoryDataset
    path = os.path.join(ROOT, "tests", "golden", "datasets", "eth", "test") + "/"
    if not os.path.isdir(path):
        return None
    params, bn = load_weights()
    eng = sstg.InferenceEngine(params, bn, dev)
    d = TrajectoryDataset(path, obs_len=8, pred_len=12, skip=1, norm_lap_matr=True)      # warm (page cache, kernels)
    driver.evaluate(d, eng, KSTEPS=K_SAMPLES, seed=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = TrajectoryDataset(path, obs_len=8, pred_len=12, skip=1, norm_lap_matr=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    reps = 20
    for i in range(reps):
        ade_, fde_, _ = driver.evaluate(d, eng, KSTEPS=K_SAMPLES, seed=i)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    driver.evaluate(d, eng, KSTEPS=K_SAMPLES, seed=0, graph=True)          # captures the split's pipeline once
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    for i in range(reps):
        driver.evaluate(d, eng, KSTEPS=K_SAMPLES, seed=0, graph=True)      # replays (each ends with the D2H of the sums)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    return {"scenes": len(d), "pedestrians": int(d.obs_traj.shape[0]), "dataset_build_ms": (t1 - t0) * 1e3,
            "evaluate_ms": (t2 - t1) / reps * 1e3, "scenes_per_s_evaluate": len(d) * reps / (t2 - t1),
            "evaluate_ms_cuda_graph": (t4 - t3) / reps * 1e3,
            "ade": ade_, "fde": fde_,
            "what": "utils.TrajectoryDataset + driver.evaluate (= test.py:18-127) on datasets/eth/test, eth checkpoint; "
                    "the unmodified reference needs ~10 minutes for its five test splits on 8 CPU cores (SURVEY 8c)"}


# ---------------------------------------------------------------------------------
# large-scene leg (BASELINE.json config 4): 256 pedestrians per scene, N x N aggregation stress
# ---------------------------------------------------------------------------------
def run_large_leg(dev, steps=5, num_scenes=2048, n_peds=256):
    """scenes/s of graph build + forward + 20-sample ADE/FDE on scenes of 256 pedestrians (A is 2 MB per
    scene: the st_gcn kernel streams it once), and the st_gcn kernel's share of the HBM roofline there."""
    import torch
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E
    params, bn = load_weights()
    eng = sstg.InferenceEngine(params, bn, dev)
    obs, pred, last = synth_batch(num_scenes, n_peds, 777)
    obs, pred, last = obs.to(dev), pred.to(dev), last.to(dev)
    plan = sstg.ScenePlan(np.full(num_scenes, n_peds, np.int32))
    dp = plan.on(dev)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms = timed(lambda: eng.infer(plan, obs, pred, last, K_SAMPLES, seed=1, want_vtr=False))
    V, A = E.graph_build(dp, obs, 8)
    prm, bnb = eng.params, eng.bn
    H = torch.empty(plan.P * 40, dtype=torch.float32, device=dev)
    ms_s = timed(lambda: E.stgcn_forward(dp, V, A, prm, bnb, H=H))
    bytes_s = (64 * n_peds + 32 * n_peds * n_peds + 160 * n_peds) * num_scenes
    peak, _ = measured_peak_gbs()
    gbs = bytes_s / (ms_s * 1e-3) / 1e9
    return {"scenes": num_scenes, "peds_per_scene": n_peds, "scenes_per_s": num_scenes / (ms * 1e-3), "ms_per_step": ms,
            "stgcn_GBps": gbs, "stgcn_frac_of_hbm": gbs / peak if peak else None,
            "what": "graph build + st_gcn + TXP-CNN + 20-sample ADE/FDE, device-resident, A materialised "
                    f"({plan.a_floats * 4 / 1e9:.1f} GB per step: larger than L2)"}


# ---------------------------------------------------------------------------------
# training leg (BASELINE.json config 2): optimiser groups of 128 ragged scenes
# ---------------------------------------------------------------------------------
def synth_ragged_counts(num_scenes, seed):
    """Pedestrian counts shaped like the ETH/UCY train splits (median 4-6,
    mean ~11, p95 ~36, max 57): log-normal, clipped to [2, 57]."""
    rng = np.random.default_rng(seed)
    n = np.rint(np.exp(rng.normal(1.8, 0.9, num_scenes))).astype(np.int32)
    return np.clip(n, 2, 57)


def run_train_leg(dev, world, rank, steps, warmup, batch=128):
    """scenes/s of train() (graph build + train-mode forward + NLL + backward + BN stats + SGD)
    on packed groups of `batch` scenes; with world > 1 every group is sharded over the ranks and
    gradient / loss / BN statistics are exchanged over NVLink peer memory by libsstg's fused
    exchange+SGD kernel (the NCCL all-reduce variant is timed beside it).  Device timing, max over ranks."""
    import torch
    import torch.distributed as dist
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import parallel as Pl
    params, bn = load_weights()
    groups = []
    for g in range(8):                                    # 8 distinct groups, cycled
        n = synth_ragged_counts(batch, 100 + g)
        obs, pred, _ = synth_batch(int(n.sum()), 1, 7000 + g)
        groups.append((n, sstg.ScenePlan(n), obs.to(dev), pred.to(dev)))

    def timed(exchange):
        eng = sstg.TrainEngine(params, bn, dev)
        trainer = Pl.DistributedTrainer(eng, exchange=exchange, max_scenes=batch, peer_timeout_s=2.0) \
            if world > 1 else None

        def one(i):
            n, plan, obs, pred = groups[i % len(groups)]
            if trainer is not None:
                trainer.step(n, obs, pred, batch, 1e-4)
            else:
                eng.step(plan, obs, pred, batch, 1e-4)

        for i in range(warmup):
            one(i)
        if trainer is not None and trainer.peer is not None:
            # a peer that cannot be reached (IPC / peer access refused on this box) shows up here, after at most
            # warmup x 2 s; every rank learns about it and the leg is re-timed through NCCL
            bad = torch.tensor([int(trainer.peer.status.item() != 0)], dtype=torch.int32, device=dev)
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            if int(bad[0]):
                trainer.peer.close()
                return None, "peer-failed"
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        if trainer is not None and trainer.peer is not None:
            trainer.peer.check()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        used = trainer.exchange if trainer is not None else "none"
        if trainer is not None and trainer.peer is not None:
            trainer.peer.close()
        return float(t[0]), used

    ms, exchange = timed("auto")
    nccl = None
    if exchange == "peer-failed":
        ms, exchange = timed("nccl")
    if world > 1 and exchange == "peer":
        ms_nccl, _ = timed("nccl")
        nccl = {"value": batch * steps / (ms_nccl * 1e-3), "unit": "scenes/s", "ms_per_group": ms_nccl / steps,
                "what": "same sharded group, exchange through one torch.distributed (NCCL) all-reduce instead"}
    peds = float(np.mean([g[0].sum() for g in groups]))
    # the same loop through the public driver (driver.train_epoch: shuffled epoch order, per-epoch gather and
    # plan upload, one library call per group), wall clock, on an ETH-train-sized synthetic split
    epoch = None
    if world == 1:
        import time
        from social_stgcnn_b200 import driver
        S_ep = 2816
        n_ep = synth_ragged_counts(S_ep, 5)
        o_ep, p_ep, _ = synth_batch(int(n_ep.sum()), 1, 9)

        class _Split:
            packed = {"device": dev, "obs_rel": o_ep.to(dev), "pred_rel": p_ep.to(dev)}
            plan_obs = sstg.ScenePlan(n_ep)

            def __len__(self):
                return S_ep

        eng_ep = sstg.TrainEngine(params, bn, dev)
        rng = np.random.default_rng(0)
        driver.train_epoch(_Split(), eng_ep, batch, 1e-4, order=rng.permutation(S_ep))
        torch.cuda.synchronize()
        per_epoch = []
        for _ in range(7):
            t0 = time.perf_counter()
            driver.train_epoch(_Split(), eng_ep, batch, 1e-4, order=rng.permutation(S_ep))   # ends with a D2H of the loss
            per_epoch.append(time.perf_counter() - t0)
        dt = float(np.median(per_epoch))
        epoch = {"value": S_ep / dt, "unit": "scenes/s", "ms_per_epoch": dt * 1e3, "scenes": S_ep,
                 "ms_per_epoch_all": [round(x * 1e3, 3) for x in per_epoch],
                 "what": "driver.train_epoch (the train() replacement) on 2,816 ragged scenes, shuffled order, "
                         "wall clock incl. the per-epoch host preparation; median of 7 epochs"}
    indep = None
    if world > 1:
        # the reference's own multi-GPU recipe (train.sh:4-18): independent trainings side by side, one per GPU,
        # no collective at all -- every rank trains its own replica on its own groups
        eng2 = sstg.TrainEngine(params, bn, dev)
        for i in range(warmup):
            eng2.step(groups[i % len(groups)][1], groups[i % len(groups)][2], groups[i % len(groups)][3], batch, 1e-4)
        dist.barrier()
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for i in range(steps):
            g = groups[i % len(groups)]
            eng2.step(g[1], g[2], g[3], batch, 1e-4)
        i1.record()
        torch.cuda.synchronize()
        t2 = torch.tensor([i0.elapsed_time(i1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        indep = {"value": world * batch * steps / (float(t2[0]) * 1e-3), "unit": "scenes/s",
                 "what": "independent trainings, one per GPU, as the reference's train.sh launches them (no collective)"}
    return {"value": batch * steps / (ms * 1e-3), "unit": "scenes/s", "ms_per_group": ms / steps, "batch": batch,
            "exchange": exchange, "nccl_exchange": nccl, "independent_replicas": indep, "epoch_through_driver": epoch,
            "groups_timed": steps, "mean_peds_per_group": peds, "n_gpus": world,
            "what": "train(): graph build + train-mode fwd (per-scene BN stats) + NLL + full bwd + BN running stats "
                    "+ SGD, one packed ragged group of 128 ETH/UCY-shaped scenes per step (8 distinct groups cycled; a "
                    "group's working set, ~3 MB, is L2-resident by nature -- the step is latency-bound, not bandwidth-bound)"
                    + ("; group sharded over ranks, gradient / loss / BN statistics exchanged by libsstg's peer-memory kernel "
                       "over NVLink fused with SGD" if exchange == "peer" else
                       "; group sharded over ranks, one NCCL all-reduce of [gradient | loss | BN statistics]"
                       if world > 1 else "")}


# ---------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # rank 0's stdout is ONE JSON line: the image exports NCCL_DEBUG=VERSION, whose only effect is a banner on
    # stdout (NCCL ignores NCCL_DEBUG_FILE at that level); higher levels are sent to stderr instead
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    from social_stgcnn_b200.parallel import bind_process_to_gpu_numa
    numa_node = bind_process_to_gpu_numa(local)   # before any pinned allocation: first touch lands next to the GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, N = args.scenes_per_gpu, N_PEDS
    params, bn = load_weights()
    eng = sstg.InferenceEngine(params, bn, dev)
    plan = sstg.ScenePlan(np.full(S, N, np.int32))
    obs_h, pred_h, last_h, start_h, delta_h = synth_batch(S, N, 1234 + rank, compact=True)   # this rank's shard
    obs_h, pred_h, last_h = obs_h.pin_memory(), pred_h.pin_memory(), last_h.pin_memory()
    start_h, delta_h = start_h.pin_memory(), delta_h.pin_memory()
    obs, pred, last = obs_h.to(dev), pred_h.to(dev), last_h.to(dev)
    dp = plan.on(dev)
    ws = eng.workspace(plan)
    names = ["graph_build", "stgcn_forward", "txpcnn_forward", "sample_ade_fde"]

    def step(ev=None):
        def mark(i):
            if ev is not None:
                ev[i].record()
        mark(0)
        E.graph_build(dp, obs, 8, V=ws["V_obs"], A=ws["A_obs"])
        mark(1)
        E.stgcn_forward(dp, ws["V_obs"], ws["A_obs"], eng.params, eng.bn, H=ws["H"])
        mark(2)
        E.txpcnn_forward(dp, ws["H"], eng.params, out=ws["V_pred"])
        mark(3)
        # the ground-truth track is read straight from pred_traj_rel (SSTG_EVAL_VTR_IS_REL): no V_tr blocks needed
        E.sample_ade_fde(dp, ws["V_pred"], pred, last, K_SAMPLES, seed=args.seed, fp64_dist=False,
                         min_ade=ws["min_ade"], min_fde=ws["min_fde"], vtr_is_rel=True)
        mark(4)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for k in range(args.steps):
        step(evs[k])
    t_end.record()
    barrier()
    sampler.stop_flag = True
    ms_total = t_start.elapsed_time(t_end)
    kern_ms = {n: float(np.mean([evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(args.steps)]))
               for i, n in enumerate(names)}
    sum_a, sum_f = float(ws["min_ade"].sum()), float(ws["min_fde"].sum())

    # ---- the same step with graph build + st_gcn fused (adjacency kept on chip, never written to HBM)
    def step_fused():
        E.stgcn_forward_fused(dp, obs, eng.params, eng.bn, H=ws["H"], Pws=ws["Pws"])
        E.txpcnn_forward(dp, ws["H"], eng.params, out=ws["V_pred"])
        E.sample_ade_fde(dp, ws["V_pred"], pred, last, K_SAMPLES, seed=args.seed, fp64_dist=False,
                         min_ade=ws["min_ade"], min_fde=ws["min_fde"], vtr_is_rel=True)
    for _ in range(args.warmup):
        step_fused()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        step_fused()
    f1.record()
    barrier()
    ms_fused = f0.elapsed_time(f1)

    # ---- e2e: host buffers -> H2D -> pipeline -> D2H, through the public API
    for _ in range(max(1, min(2, args.warmup))):
        eng.infer_host(plan, obs_h, pred_h, last_h, K_SAMPLES, seed=args.seed, chunk_scenes=args.chunk_scenes)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        a_h, f_h = eng.infer_host(plan, obs_h, pred_h, last_h, K_SAMPLES, seed=args.seed,
                                  chunk_scenes=args.chunk_scenes)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    # ---- e2e with the compact fixed-point host format (88 B instead of 168 B per pedestrian over PCIe)
    a_ref, f_ref = a_h.copy(), f_h.copy()
    for _ in range(2):
        a_c, f_c = eng.infer_host(plan, None, None, None, K_SAMPLES, seed=args.seed, chunk_scenes=args.chunk_scenes,
                                  compact=(start_h, delta_h))
    compact_identical = bool(np.array_equal(a_c, a_ref) and np.array_equal(f_c, f_ref))
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        eng.infer_host(plan, None, None, None, K_SAMPLES, seed=args.seed, chunk_scenes=args.chunk_scenes,
                       compact=(start_h, delta_h))
    k1.record()
    barrier()
    ms_e2e_c = k0.elapsed_time(k1)

    # ---- H2D copy bandwidth of this box (context for e2e: PCIe carries 168 B per pedestrian)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    obs.copy_(obs_h, non_blocking=True)
    pred.copy_(pred_h, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = (obs_h.numel() + pred_h.numel()) * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    # ---- live device copy bandwidth (the same recipe as MEASURED_PEAKS.json: b.copy_(a), read + write bytes)
    ca = torch.empty(256 * 1024 * 1024, dtype=torch.float32, device=dev)
    cb = torch.empty_like(ca)
    best = 1e9
    for _ in range(6):
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        cb.copy_(ca)
        q1.record()
        torch.cuda.synchronize()
        best = min(best, q0.elapsed_time(q1))
    copy_gbs = 2 * ca.numel() * 4 / (best * 1e-3) / 1e9
    del ca, cb
    train = graph_leg = ragged_leg = large_leg = eth_leg = None
    if not args.no_train:
        train = run_train_leg(dev, world, rank, steps=max(20, 10 * args.steps), warmup=5)
        if rank == 0:
            graph_leg = run_graph_leg(dev)
            ragged_leg = run_ragged_leg(dev)
            large_leg = run_large_leg(dev)
            eth_leg = run_eth_leg(dev)

    t = torch.tensor([ms_total, ms_e2e, ms_e2e_c, ms_fused], dtype=torch.float64, device=dev)
    red = torch.tensor([sum_a, sum_f, float(plan.P)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # max over ranks
        dist.all_reduce(red, op=dist.ReduceOp.SUM)        # dataset-level ADE/FDE: the only exchange, outside timing
    ms_total, ms_e2e, ms_e2e_c, ms_fused = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        total_scenes = S * world * args.steps
        value = total_scenes / (ms_total * 1e-3)
        sg = kern_ms["stgcn_forward"]
        bytes_s = (64 * N + 32 * N * N + 160 * N) * S            # SURVEY 8(d) row "S", per launch
        achieved = bytes_s / (sg * 1e-3) / 1e9
        ksum = sum(kern_ms.values())
        alg = {"graph_build": (128 * N + 32 * N * N) * S, "stgcn_forward": bytes_s,
               "txpcnn_forward": 400 * N * S, "sample_ade_fde": 352 * N * S}
        kernels = {n: {"ms": round(kern_ms[n], 4), "share": round(kern_ms[n] / ksum, 4),
                       "GBps": round(alg[n] / (kern_ms[n] * 1e-3) / 1e9, 1)} for n in names}
        kernels["txpcnn_forward"]["TFLOPs_fp32"] = round(60480 * N * S / (kern_ms["txpcnn_forward"] * 1e-3) / 1e12, 2)
        kernels["txpcnn_forward"]["bound"] = "fp32 FMA (not HBM, not tensor): 60,480 flop/ped-column vs 400 B"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(S, world),
            "host": {"numa_node_rank0": numa_node, "cpus": os.cpu_count()},
            "roofline": {"kernel": "stgcn_fwd_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (tr["dram_bytes"] * S / tr["scenes"] if (tr := profiled_traffic()) else None),
                         "traffic_source": (f"ncu --set full, {tr['file']}: {tr['dram_bytes']} B for a launch of "
                                            f"{tr['scenes']} scenes, scaled to this launch" if tr else None),
                         "algorithmic_bytes_per_launch": bytes_s, "peak_source": peak_src,
                         "copy_GBps_live": round(copy_gbs, 1),
                         "share_of_step": kern_ms["stgcn_forward"] / ksum},
            "kernels": kernels,
            "e2e": {"value": S * world * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": eng.h2d_bytes(plan), "d2h_bytes_per_step": eng.d2h_bytes(plan),
                    "ms_per_step": ms_e2e / args.steps, "h2d_GBps_measured": round(h2d_gbs, 1),
                    "how": f"InferenceEngine.infer_host: pinned host buffers, {len(plan.chunks(args.chunk_scenes, True))} chunks of <= "
                           f"{args.chunk_scenes} scenes (quarter-size first and last) software-pipelined over 3 streams (H2D | kernels | D2H); "
                           f"graph build + st_gcn as the fused pair (only ADE/FDE leave the device)"},
            "value_fused": {"value": total_scenes / (ms_fused * 1e-3), "unit": UNIT, "ms_per_step": ms_fused / args.steps,
                            "what": "same work with sstg_stgcn_forward_fused: graph build + st_gcn without "
                                    "materialising A (3 + 1 kernels); used by the host-fed e2e paths, not by `value`"},
            "e2e_compact": {"value": S * world * args.steps / (ms_e2e_c * 1e-3), "unit": UNIT,
                            "h2d_bytes_per_step": eng.h2d_bytes(plan, compact=True),
                            "d2h_bytes_per_step": eng.d2h_bytes(plan), "ms_per_step": ms_e2e_c / args.steps,
                            "identical_to_fp32_interface": compact_identical,
                            "how": "same pipeline fed the fixed-point host format (int32 start + int16 steps in units "
                                   "of 1e-4, sstg_decode_tracks rebuilds the fp32 tensors bit for bit on the device)"},
            # per step: graph, st_gcn, TXP-CNN (Winograd), sampling + the 8-CTA filter-transform kernel that refreshes
            # the Winograd constant bank before every TXP-CNN launch (plus one 32 KB device-to-device copy, not a kernel)
            "gpu_launches": (4 + int(plan.n_tiles >= 2 * torch.cuda.get_device_properties(dev).multi_processor_count
                                     and os.environ.get("SSTG_TXP_FILTERS", "")[:1] != "s"
                                     and os.environ.get("SSTG_TXP_WINO", "2")[:1] != "0")) * args.steps,
            "train": train,
            "graph_construction": graph_leg,
            "ragged_inference": ragged_leg,
            "large_scenes_256": large_leg,
            "eth_test_split": eth_leg,
            "clocks": sampler.result(),
            "result": {"ade": red[0].item() / red[2].item(), "fde": red[1].item() / red[2].item(),
                       "peds": int(red[2].item())},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v1, d1, t1 = cpu_port_throughput(max(4, cores), 1)      # calibrate
            n = int(max(cores, min(4096, 12.0 * v1 * cores)))
            v, done, dt = cpu_port_throughput(n, cores)
            ref_t = ""
            p = os.path.join(ROOT, "tests", "golden", "REFERENCE_TIMING.txt")
            if os.path.exists(p):
                ref_t = " | " + open(p).read().strip()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{done} scenes of {N} peds in {dt:.1f}s, oracle port, one process per "
                                              f"core (single process: {v1:.1f} scenes/s){ref_t}",
                                    "phases_single_thread": cpu_port_phases()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes-per-gpu", type=int, default=131072)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-throughput leg")
    ap.add_argument("--chunk-scenes", type=int, default=8192, help="scenes per pipeline chunk of the e2e path")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch ourselves under torchrun, one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
