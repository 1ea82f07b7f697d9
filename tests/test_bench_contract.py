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


import pytest  # noqa: E402


@pytest.mark.gpu
def test_our_arm_prints_one_contract_line_on_the_gpu():
    """bench.py (our arm) on a small shard: ONE JSON line on stdout with every key of the driver's contract and of the
    tier contract (roofline, e2e with the bytes it copies, gpu_launches, clocks), a config equal to bench_config()."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    S = 4096
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3",
                        "--scenes-per-gpu", str(S), "--no-train", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["metric"] == bench.METRIC and d["unit"] == "scenes/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["config"] == bench.bench_config(S, 1)
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - S) < 1e-6 * S
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["kernel"] == "stgcn_fwd_kernel"
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and rf["achieved"] > 0
    assert rf["algorithmic_bytes_per_launch"] == (64 * 64 + 32 * 64 * 64 + 160 * 64) * S
    e = d["e2e"]
    assert e["unit"] == "scenes/s" and e["value"] > 0
    assert e["h2d_bytes_per_step"] == S * 64 * 168 and e["d2h_bytes_per_step"] == S * 64 * 16
    assert d["gpu_launches"] in (4 * 2, 5 * 2)           # + the Winograd filter refresh when the grid is full
    assert set(d["kernels"]) == {"graph_build", "stgcn_forward", "txpcnn_forward", "sample_ade_fde"}
    assert abs(sum(k["share"] for k in d["kernels"].values()) - 1.0) < 1e-3
    c = d["clocks"]
    assert "sm_mhz" in c and "reasons" in c and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    assert d["e2e_compact"]["identical_to_fp32_interface"] is True
    assert 0.1 < d["result"]["ade"] < 5.0 and d["result"]["peds"] == S * 64
