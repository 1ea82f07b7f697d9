"""Host + device cost of driver.train_epoch per optimiser group on an ETH-train-sized synthetic dataset
(2,816 ragged scenes, shuffled order as the reference's DataLoader does)."""
import os
import sys
import time
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import social_stgcnn_b200 as sstg
from social_stgcnn_b200 import driver

dev = torch.device("cuda", 0)
params, bn = bench.load_weights()
S = 2816
n = bench.synth_ragged_counts(S, 5)
obs, pred, _ = bench.synth_batch(int(n.sum()), 1, 9)
ds = SimpleNamespace(packed={"device": dev, "obs_rel": obs.to(dev), "pred_rel": pred.to(dev)},
                     plan_obs=sstg.ScenePlan(n), __len__=None)


class DS:
    packed = ds.packed
    plan_obs = ds.plan_obs

    def __len__(self):
        return S


d = DS()
eng = sstg.TrainEngine(params, bn, dev)
rng = np.random.default_rng(0)
driver.train_epoch(d, eng, 128, 1e-4, order=rng.permutation(S))
torch.cuda.synchronize()
t0 = time.perf_counter()
E = 5
for _ in range(E):
    driver.train_epoch(d, eng, 128, 1e-4, order=rng.permutation(S))
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / E
print(f"train_epoch: {dt*1e3:.2f} ms per epoch of {S} scenes = {dt/(S/128)*1e6:.0f} us per group, {S/dt:.0f} scenes/s (wall clock)")
