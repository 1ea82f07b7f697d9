"""Quick per-kernel timing probe on synthetic N-pedestrian scenes (CUDA events)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import social_stgcnn_b200 as sstg
from social_stgcnn_b200 import engine as E

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=32768)
ap.add_argument("--peds", type=int, default=64)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--fp64", action="store_true")
ap.add_argument("--only", default="", help="comma-separated substrings: time only the matching kernels (the others run once)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
S, N = a.scenes, a.peds
g = torch.Generator(device="cpu").manual_seed(0)
P = S * N
vel = (torch.rand(P, 2, 1, generator=g) - 0.5)
obs_rel = (vel + 0.1 * torch.randn(P, 2, 8, generator=g)); obs_rel[:, :, 0] = 0
pred_rel = (vel + 0.1 * torch.randn(P, 2, 12, generator=g))
last = torch.rand(P, 2, generator=g) * 20 - 10
obs_rel, pred_rel, last = obs_rel.to(dev), pred_rel.to(dev), last.to(dev)
w = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights.npz"))
eng = sstg.InferenceEngine(w["eth.params"], w["eth.bn"], dev)
plan = sstg.ScenePlan(np.full(S, N, np.int32))
dp = plan.on(dev)
ws = eng.workspace(plan)

def timeit(fn, name, bytes_=None, flops=None):
    if a.only and not any(k in name for k in a.only.split(",")):
        fn()
        return 0.0
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    s = f"{name:28s} {ms:9.3f} ms  {S/ms/1e3:9.3f} M scenes/s"
    if bytes_: s += f"  {bytes_*S/ms/1e6:9.1f} GB/s"
    if flops: s += f"  {flops*S/ms/1e9:9.2f} TFLOP/s"
    print(s, flush=True)
    return ms

tot = 0
tot += timeit(lambda: E.graph_build(dp, obs_rel, 8, fp64=a.fp64, V=ws["V_obs"], A=ws["A_obs"]), "graph_build(obs)", 128*N+32*N*N)
tot += timeit(lambda: E.graph_build(dp, pred_rel, 12, v_only=True, V=ws["V_tr"]), "graph_build(pred,V only)", 192*N)
tot += timeit(lambda: E.stgcn_forward(dp, ws["V_obs"], ws["A_obs"], eng.params, eng.bn, H=ws["H"]), "stgcn_forward", 64*N+32*N*N+160*N)
tot += timeit(lambda: E.txpcnn_forward(dp, ws["H"], eng.params, out=ws["V_pred"]), "txpcnn_forward", 400*N, 60480*N)
tot += timeit(lambda: E.sample_ade_fde(dp, ws["V_pred"], ws["V_tr"], last, 20, fp64_dist=a.fp64, min_ade=ws["min_ade"], min_fde=ws["min_fde"]), "sample_ade_fde(philox)", 352*N)
print(f"sum of kernels {tot:.3f} ms -> {S/tot/1e3:.3f} M scenes/s")
timeit(lambda: eng.infer(plan, obs_rel, pred_rel, last, 20, graph_fp64=a.fp64, fp64_dist=a.fp64), "pipeline infer()")
