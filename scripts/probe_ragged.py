"""Per-kernel CUDA-event times of the inference pipeline on a packed batch of ETH/UCY-shaped scenes
(2..57 pedestrians, median ~6: bench.py's ragged leg), plain and fused front."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import social_stgcnn_b200 as sstg
from social_stgcnn_b200 import engine as E
import bench

dev = torch.device("cuda", 0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
params, bn = bench.load_weights()
eng = sstg.InferenceEngine(params, bn, dev)
n = bench.synth_ragged_counts(S, 17)
obs, pred, last = [t.to(dev) for t in bench.synth_batch(int(n.sum()), 1, 4242)]
plan = sstg.ScenePlan(n)
dp = plan.on(dev)
ws = eng.workspace(plan)
steps = [
    ("graph", lambda: E.graph_build(dp, obs, 8, V=ws["V_obs"], A=ws["A_obs"])),
    ("stgcn", lambda: E.stgcn_forward(dp, ws["V_obs"], ws["A_obs"], eng.params, eng.bn, H=ws["H"])),
    ("fused_front", lambda: E.stgcn_forward_fused(dp, obs, eng.params, eng.bn, H=ws["H"], Pws=ws["Pws"])),
    ("txp", lambda: E.txpcnn_forward(dp, ws["H"], eng.params, out=ws["V_pred"])),
    ("sample", lambda: E.sample_ade_fde(dp, ws["V_pred"], pred, last, 20, fp64_dist=False, vtr_is_rel=True,
                                        min_ade=ws["min_ade"], min_fde=ws["min_fde"])),
]
print(f"{S} scenes, {plan.P} pedestrians, A {plan.a_floats * 4 / 1e6:.1f} MB, tiles {plan.n_tiles}")
for name, fn in steps:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:12s} {np.median(ts) * 1e3:9.1f} us   {np.median(ts) * 1e6 / S:7.2f} ns/scene   {np.median(ts) * 1e6 / plan.P:6.2f} ns/ped")
