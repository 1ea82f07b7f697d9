"""Can the HBM-bound st_gcn kernel and the ALU-bound sampling kernel share the SMs?  Times S (st_gcn on half A) and
E (sampling on half B) back to back on one stream and concurrently on two streams, with st_gcn's persistent grid
reduced to `--stgcn-sms` x 2 CTAs so that the sampling kernel's CTAs find room beside it."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import social_stgcnn_b200 as sstg
from social_stgcnn_b200 import engine as E
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=65536)
ap.add_argument("--stgcn-sms", type=int, nargs="+", default=[0, 74, 111])
a = ap.parse_args()
dev = torch.device("cuda", 0)
S, N = a.scenes, 64
params, bn = bench.load_weights()
eng = sstg.InferenceEngine(params, bn, dev)
plan = sstg.ScenePlan(np.full(S, N, np.int32))
dp = plan.on(dev)
obs, pred, last = [t.to(dev) for t in bench.synth_batch(S, N, 1)]
V, A = E.graph_build(dp, obs, 8)
H = E.stgcn_forward(dp, V, A, eng.params, eng.bn)
vp = E.txpcnn_forward(dp, H, eng.params)
H2 = torch.empty_like(H)
ade = torch.empty(plan.P, dtype=torch.float64, device=dev)
fde = torch.empty_like(ade)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run_S(sms):
    E.stgcn_forward(dp, V, A, eng.params, eng.bn, H=H2, sm_count=sms)


def run_E():
    E.sample_ade_fde(dp, vp, pred, last, 20, fp64_dist=False, vtr_is_rel=True, min_ade=ade, min_fde=fde)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


print(f"{S} scenes x {N}: E alone {timed(run_E):.3f} ms")
for sms in a.stgcn_sms:
    tS = timed(lambda: run_S(sms))

    def both():
        main = torch.cuda.current_stream(dev)
        s1.wait_stream(main)
        s2.wait_stream(main)
        with torch.cuda.stream(s1):
            run_S(sms)
        with torch.cuda.stream(s2):
            run_E()
        main.wait_stream(s1)
        main.wait_stream(s2)

    print(f"st_gcn grid {'full' if sms <= 0 else 2 * sms:>4}: S alone {tS:.3f} ms, S || E {timed(both):.3f} ms")
