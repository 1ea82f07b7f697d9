"""A/B several builds of libsstg.so in ONE process on the dense-crowd workload (CUDA events).

    python scripts/ab_libs.py [--scenes 131072] [--peds 64] [--rounds 3] lib_a.so lib_b.so@SSTG_TXP_WINO=4 ...

The four inference kernels are timed for every library, interleaved over `rounds` passes (so that clock or
thermal drift hits all builds alike); every output is compared BITWISE with the first library's.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import social_stgcnn_b200 as sstg
from social_stgcnn_b200 import _lib as L
from social_stgcnn_b200 import engine as E

ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+")
ap.add_argument("--scenes", type=int, default=131072)
ap.add_argument("--peds", type=int, default=64)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--only", default="", help="comma-separated substrings of kernel names to time")
a = ap.parse_args()
dev = torch.device("cuda", 0)
S, N = a.scenes, a.peds
P = S * N
g = torch.Generator(device="cpu").manual_seed(0)
vel = torch.rand(P, 2, 1, generator=g) - 0.5
obs_rel = vel + 0.1 * torch.randn(P, 2, 8, generator=g)
obs_rel[:, :, 0] = 0
pred_rel = vel + 0.1 * torch.randn(P, 2, 12, generator=g)
last = torch.rand(P, 2, generator=g) * 20 - 10
obs_rel, pred_rel, last = obs_rel.to(dev), pred_rel.to(dev), last.to(dev)
w = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights.npz"))


_env_set = []


def use(spec):
    """spec = path[@KEY=VAL[,KEY=VAL...]]: the library to load and the tuning knobs (read by the library with getenv
    at every launch) to set while it is in use."""
    path, _, env = spec.partition("@")
    for k in _env_set:
        os.environ.pop(k, None)
    _env_set.clear()
    for kv in filter(None, env.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
        _env_set.append(k)
    L._lib = None
    L.LIB_PATH = os.path.abspath(path)
    L.load()


use(a.libs[0])
eng = sstg.InferenceEngine(w["eth.params"], w["eth.bn"], dev)
plan = sstg.ScenePlan(np.full(S, N, np.int32))
dp = plan.on(dev)
ws = eng.workspace(plan)
steps = [
    ("graph", lambda: E.graph_build(dp, obs_rel, 8, V=ws["V_obs"], A=ws["A_obs"]), lambda: ws["A_obs"][: 1 << 24]),
    ("stgcn", lambda: E.stgcn_forward(dp, ws["V_obs"], ws["A_obs"], eng.params, eng.bn, H=ws["H"]), lambda: ws["H"]),
    ("txp", lambda: E.txpcnn_forward(dp, ws["H"], eng.params, out=ws["V_pred"]), lambda: ws["V_pred"]),
    ("sample", lambda: E.sample_ade_fde(dp, ws["V_pred"], pred_rel, last, 20, fp64_dist=False, vtr_is_rel=True,
                                        min_ade=ws["min_ade"], min_fde=ws["min_fde"]),
     lambda: torch.cat([ws["min_ade"], ws["min_fde"]])),
]
Hf = torch.empty_like(ws["H"])
steps.append(("fused_front", lambda: E.stgcn_forward_fused(dp, obs_rel, eng.params, eng.bn, H=Hf, Pws=ws["Pws"]), lambda: Hf))
# the same kernel with INJECTED noise (the <EXACT=0, PHILOX=0> instantiation) on the first scenes of the batch
S2 = min(S, 4096)
P2 = S2 * N
dp2 = sstg.ScenePlan(np.full(S2, N, np.int32)).on(dev)
eps2 = torch.randn(P2 * 20 * 24, generator=torch.Generator(device="cpu").manual_seed(1)).to(dev)
ade2 = torch.empty(P2, dtype=torch.float64, device=dev)
fde2 = torch.empty(P2, dtype=torch.float64, device=dev)
steps.append(("sample_eps", lambda: E.sample_ade_fde(dp2, ws["V_pred"][: P2 * 60], pred_rel[:P2], last[:P2], 20, eps=eps2,
                                                    fp64_dist=False, vtr_is_rel=True, min_ade=ade2, min_fde=fde2),
              lambda: torch.cat([ade2, fde2])))
ref = {}
times = {lib: {name: [] for name, _, _ in steps} for lib in a.libs}
for rnd in range(a.rounds):
    for lib in a.libs:
        use(lib)
        for name, fn, out in steps:
            timed = not a.only or any(k in name for k in a.only.split(","))
            fn()
            if rnd == 0:
                o = out().clone()
                if name not in ref:
                    ref[name] = o
                elif not torch.equal(ref[name].view(torch.int32), o.view(torch.int32)):
                    print(f"!! {lib}: {name} differs from {a.libs[0]} "
                          f"(max abs {float((ref[name] - o).abs().max()):.3e})", flush=True)
            if not timed:
                continue
            torch.cuda.synchronize()
            for _ in range(a.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                times[lib][name].append(e0.elapsed_time(e1))
for lib in a.libs:
    row = "  ".join(f"{name} {np.median(v):7.3f}" for name, v in times[lib].items() if v)
    print(f"{os.path.basename(lib):40s} {row}", flush=True)
