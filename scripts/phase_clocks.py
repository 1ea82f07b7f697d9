"""Per-phase clock64() breakdown of the latency-bound training kernels on one 128-scene group.

    python scripts/phase_clocks.py --build      # here (no GPU needed): nvcc -DSSTG_PHASE_CLOCKS -> _scratch/libsstg_phase.so
    gpurun -- 'SSTG_LIB=/root/repo/_scratch/libsstg_phase.so python scripts/phase_clocks.py'

With -DSSTG_PHASE_CLOCKS thread 0 of every CTA stamps clock64() at the phase boundaries of its first tile /
scene (txpcnn_bwd.cu, txpcnn_fwd.cu, train_kernels.cu); the extra entry points sstg_debug_phase_clocks* copy
the stamps back.  Printed: the CTA that owns the largest, the median and the smallest tile / scene."""
import argparse
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ALT = os.path.join(ROOT, "_scratch", "libsstg_phase.so")


def build():
    from social_stgcnn_b200 import build as B
    obj = os.path.join(ROOT, "_scratch", "obj")
    os.makedirs(obj, exist_ok=True)
    objs = []
    for src in B.SOURCES:
        o = os.path.join(obj, src.replace(".cu", ".o"))
        subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-DSSTG_PHASE_CLOCKS", "-c", os.path.join(B.CSRC, src), "-o", o],
                       check=True, capture_output=True)
        objs.append(o)
    subprocess.run([B._nvcc(), "-shared", "-o", ALT] + objs + ["-cudart", "shared", "-Xlinker",
                                                               "-rpath,/usr/local/cuda/lib64"], check=True)
    print(ALT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--batch", type=int, default=128)
    a = ap.parse_args()
    if a.build:
        return build()
    import numpy as np
    import torch
    import bench
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import _lib as L
    lib = L.load()
    if not hasattr(lib, "sstg_debug_phase_clocks"):
        raise SystemExit("this libsstg.so was built without -DSSTG_PHASE_CLOCKS: run with --build first and set SSTG_LIB")
    dev = torch.device("cuda", 0)
    params, bn = bench.load_weights()
    eng = sstg.TrainEngine(params, bn, dev)
    n = bench.synth_ragged_counts(a.batch, 100)
    obs, pred, _ = bench.synth_batch(int(n.sum()), 1, 7000)
    obs, pred = obs.to(dev), pred.to(dev)
    plan = sstg.ScenePlan(n)
    for _ in range(5):
        eng.step(plan, obs, pred, a.batch, 1e-4)
    torch.cuda.synchronize()

    def read(fn, rows, width):
        buf = (ctypes.c_longlong * (rows * width))()
        f = getattr(lib, fn)
        f.argtypes = [ctypes.c_void_p, ctypes.c_int]
        assert f(buf, rows * width) == 0
        return np.array(buf).reshape(rows, width)

    nt = plan.n_tiles
    cols = np.array([n[plan.tile_off[t]:plan.tile_off[t + 1]].sum() for t in range(nt)])
    order = np.argsort(-cols)
    pick = [int(order[0]), int(order[nt // 2]), int(order[-1])]
    xb = read("sstg_debug_phase_clocks", 160, 40)
    print(f"TXP-CNN backward ({nt} tiles, max {cols.max()} columns); cycles")
    for r in pick:
        c = xb[r]
        print(f" cta {r} cols {cols[r]} total {c[38] - c[39]}: setup {c[1] - c[0]} first loads {c[2] - c[1]}")
        for li in range(5):
            b = 3 + li * 6
            line = f"   L{4 - li}: dW {c[b + 1] - c[b]}  conv^T (+wait for dW, copy issue) {c[b + 2] - c[b + 1]}"
            if li < 4:
                line += f"  du-update {c[b + 4] - c[b + 3]}  copy wait + barrier {c[b + 6] - c[b + 4]}"
            print(line)
    xf = read("sstg_debug_phase_clocks_xf", 160, 16)
    print("TXP-CNN forward (training variant)")
    for r in pick:
        c = xf[r]
        line, prev = f" cta {r} cols {cols[r]}: weights + wait for predecessor {c[1] - c[0]} setup+load {c[2] - c[1]}", c[2]
        for l in range(5):
            line += f" | L{l} conv {c[3 + 2 * l] - prev} post {(c[4 + 2 * l] if l < 4 else c[12]) - c[3 + 2 * l]}"
            prev = c[4 + 2 * l]
        print(line)
    st = read("sstg_debug_phase_clocks_st", 300, 24)
    o2 = np.argsort(-n)
    print("st_gcn train forward / backward")
    for r in (int(o2[0]), int(o2[len(n) // 2]), int(o2[-1])):
        c = st[r]
        print(f" scene {r} N {n[r]}: FWD aggregate {c[14] - c[13]} stats1 {c[15] - c[14]} bn1 {c[16] - c[15]} tconv {c[17] - c[16]} "
              f"stats2 {c[18] - c[17]} out {c[19] - c[18]}")
        print(f"            BWD A {c[2] - c[1]} sum15 {c[3] - c[2]} B {c[4] - c[3]} C {c[5] - c[4]} sum10 {c[6] - c[5]} "
              f"D {c[7] - c[6]} E {c[8] - c[7]} final {c[11] - c[10]}")


if __name__ == "__main__":
    main()
