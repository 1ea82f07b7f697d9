"""Timing probe of the training step (one optimiser group of 128 ragged scenes) and of a large
synthetic group; prints per-step device time.  Run under ncu for the per-kernel launch list."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import social_stgcnn_b200 as sstg

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--dense", type=int, default=0, help="if > 0: N pedestrians per scene instead of the ragged mix")
a = ap.parse_args()
dev = torch.device("cuda", 0)
params, bn = bench.load_weights()
eng = sstg.TrainEngine(params, bn, dev)
n = np.full(a.batch, a.dense, np.int32) if a.dense else bench.synth_ragged_counts(a.batch, 100)
obs, pred, _ = bench.synth_batch(int(n.sum()), 1, 7000)
obs, pred = obs.to(dev), pred.to(dev)
plan = sstg.ScenePlan(n)
for _ in range(5):
    eng.step(plan, obs, pred, a.batch, 1e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(a.steps):
    eng.step(plan, obs, pred, a.batch, 1e-4)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"batch {a.batch} scenes, {int(n.sum())} peds: {ms*1e3:.1f} us/step device, {(t1-t0)/a.steps*1e6:.1f} us/step host issue, "
      f"{a.batch/ms*1e3:.0f} scenes/s")
