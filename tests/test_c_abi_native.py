"""The C ABI driven from plain C (examples/c_abi_eval.c): no Python and no torch between the caller and
libsstg.so.  CPU: the example compiles against include/sstg.h and links with the library (gcc, no nvcc).
GPU: the binary runs the whole path on golden scenes and its outputs are checked against the reference's
golden vectors (graphs bit-exact, V_pred / NLL within 1e-5) and against the ctypes-driven engine (ADE/FDE
bit-identical for the same Philox seed)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp):
    from social_stgcnn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from social_stgcnn_b200.build import build
        build()
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = os.path.join(tmp, "c_abi_eval")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = [cc, "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "c_abi_eval.c"), "-o", exe, "-L", libdir, "-lsstg",
           "-L", os.path.join(CUDA, "lib64"), "-lcudart", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{CUDA}/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_example_compiles_and_links_against_the_header(tmp_path):
    exe = _build(str(tmp_path))
    assert os.path.exists(exe)
    # usage error path (no GPU touched): exit code 1
    assert subprocess.run([exe], capture_output=True).returncode == 1


@pytest.mark.gpu
def test_c_caller_reproduces_the_reference(golden, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import social_stgcnn_b200 as sstg
    exe = _build(str(tmp_path))
    names = [n for n in golden.names if n.startswith("eth")][:24] + [n for n in golden.names if n.startswith("univ")][:3]
    names, n_peds, obs_rel, pred_rel, last_obs = golden.pack(names)
    params, bn = golden.params("eth")
    S, K, seed = len(names), 20, 12345
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.array([S, K], np.int32).tobytes())
        f.write(np.array([seed], np.uint64).tobytes())
        for a, dt in ((n_peds, np.int32), (params, np.float32), (bn, np.float32), (obs_rel, np.float32),
                      (pred_rel, np.float32), (last_obs, np.float32)):
            f.write(np.ascontiguousarray(a, dt).tobytes())
    r = subprocess.run([exe, inp, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    a_floats = int(np.frombuffer(raw[:8], np.int64)[0])
    P = int(n_peds.sum())
    o = 8
    A = np.frombuffer(raw[o:o + 4 * a_floats], np.float32); o += 4 * a_floats
    vp = np.frombuffer(raw[o:o + 4 * P * 60], np.float32); o += 4 * P * 60
    loss = np.frombuffer(raw[o:o + 4 * S], np.float32); o += 4 * S
    ade = np.frombuffer(raw[o:o + 8 * P], np.float64); o += 8 * P
    fde = np.frombuffer(raw[o:o + 8 * P], np.float64); o += 8 * P
    assert o == len(raw)
    plan = sstg.ScenePlan(n_peds)
    assert plan.a_floats == a_floats
    for i, name in enumerate(names):
        n = int(n_peds[i])
        a = A[plan.a_off[i]:plan.a_off[i] + 8 * n * n].reshape(8, n, n)
        ref = golden.graph[f"{name}.A_obs"]
        assert np.array_equal(a.view(np.uint32), ref.view(np.uint32)), name            # networkx, bit for bit
        if name.startswith("eth"):
            got = vp[plan.seg_off[i] * 60:(plan.seg_off[i] + n) * 60].reshape(5, 12, n)
            want = golden.fwd[f"eth.{name}.eval"].reshape(got.shape)       # the reference model's eval forward
            assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), name
    # the ctypes-driven engine on the same inputs: identical library, identical results
    dev = torch.device("cuda", 0)
    eng = sstg.InferenceEngine(params, bn, dev)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    rr = eng.infer(plan, t(obs_rel), t(pred_rel), t(last_obs), K, seed=seed, graph_fp64=True, fp64_dist=True)
    lref = sstg.nll_forward(plan.on(dev), rr["V_pred"], rr["V_tr"]).cpu().numpy()
    torch.cuda.synchronize()
    assert np.array_equal(vp.view(np.uint32), rr["V_pred"].cpu().numpy().view(np.uint32))
    assert np.array_equal(loss.view(np.uint32), lref.view(np.uint32))
    assert np.array_equal(ade, rr["min_ade"].cpu().numpy()) and np.array_equal(fde, rr["min_fde"].cpu().numpy())
    ref_nll = golden.loss["eth.nll_eval"][:24]
    assert np.abs(loss[:24] - ref_nll).max() <= 1e-5 * np.maximum(np.abs(ref_nll), 1.0).max()
