"""CPU test of bench.py's contract for the reference arm (the only arm that runs without a GPU):
one JSON line with the keys the driver reads, our arm's metric / unit / config, and the
cpu_baseline / e2e objects of the tier contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scenes_per_sec_graph_fwd_ade_fde" and d["unit"] == "scenes/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and d["config"]["peds_per_scene"] == 64
    # the reference arm reports exactly the configuration our arm reports (both call bench_config)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(131072, 1)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": bench_config(') == 2


def test_our_arm_refuses_to_run_without_cuda():
    """No CPU fallback: without a CUDA device the product arm fails loudly instead of timing the oracle."""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
