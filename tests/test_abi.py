"""CPU tests of the boundary: libsstg.so loads, exports every symbol include/sstg.h declares,
the host-side planner behaves, and argument errors come back as codes (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from social_stgcnn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from social_stgcnn_b200.build import build
        build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "sstg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sstg_[a-z0-9_]+)\s*\(", hdr))
    assert {"sstg_graph_build", "sstg_stgcn_forward", "sstg_txpcnn_forward", "sstg_nll_forward",
            "sstg_sample_ade_fde", "sstg_plan_host", "sstg_version"} <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/sstg.h but not exported"
    assert lib.sstg_version() >= 100


def test_header_constants_match_python_binding():
    from social_stgcnn_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sstg.h")).read()
    c = {k: int(v) for k, v in re.findall(r"#define\s+(SSTG_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", hdr)}
    assert (c["SSTG_T_OBS"], c["SSTG_T_PRED"], c["SSTG_C"]) == (_lib.T_OBS, _lib.T_PRED, _lib.C_OUT)
    assert c["SSTG_N_PARAMS"] == _lib.N_PARAMS == 7563 and c["SSTG_N_BN"] == _lib.N_BN
    assert c["SSTG_MAX_PEDS"] == _lib.MAX_PEDS and c["SSTG_XTILE_COLS"] == _lib.XTILE_COLS
    assert (c["SSTG_GRAPH_NORM_LAP"], c["SSTG_GRAPH_FP64"], c["SSTG_GRAPH_V_ONLY"]) == (1, 2, 4)


def test_plan_host_offsets_and_tiles(lib):
    import social_stgcnn_b200 as sstg
    n = [2, 3, 64, 64, 5, 130, 2]
    p = sstg.ScenePlan(n)
    assert p.S == 7 and p.P == sum(n) and p.max_n == 130
    assert list(p.seg_off) == list(np.concatenate([[0], np.cumsum(n)]))
    # every adjacency block starts 16-byte aligned and is large enough
    for i, k in enumerate(n):
        assert p.a_off[i] % 4 == 0 and p.a_off[i + 1] - p.a_off[i] >= 8 * k * k
    assert p.a_floats == p.a_off[-1]
    # tiles: whole scenes; column cap = clamp(ceil(P / 296), 8, 128) unless a single scene is larger
    cap = min(128, max(8, -(-p.P // 296)))
    assert cap == 8 and list(p.tile_off) == [0, 2, 3, 4, 5, 6, 7] and p.n_tiles == 6
    for t in range(p.n_tiles):
        cols = p.seg_off[p.tile_off[t + 1]] - p.seg_off[p.tile_off[t]]
        assert cols <= cap or p.tile_off[t + 1] - p.tile_off[t] == 1
    big = sstg.ScenePlan([64] * 4096)                       # a large batch gets the full 128-column tiles
    assert big.n_tiles == 2048 and np.all(np.diff(big.tile_off) == 2)


def test_plan_host_many_tiny_scenes_cap(lib):
    import social_stgcnn_b200 as sstg
    p = sstg.ScenePlan([2] * 100)
    assert np.all(np.diff(p.tile_off) <= 32) and p.tile_off[-1] == 100
    p12 = sstg.ScenePlan([3, 5], T=12)
    assert p12.a_off[1] == 108 and p12.a_off[2] == 108 + 300


def test_plan_host_rejects_bad_input(lib):
    import social_stgcnn_b200 as sstg
    with pytest.raises(RuntimeError, match="invalid argument"):
        sstg.ScenePlan([2, 0, 3])
    with pytest.raises(RuntimeError, match="SSTG_MAX_PEDS"):
        sstg.ScenePlan([2, 257])
    assert sstg.ScenePlan([]).P == 0


def test_null_and_range_errors_return_codes(lib):
    """Argument validation happens before any CUDA call, so it is testable without a GPU."""
    z = ctypes.c_void_p(0)
    assert lib.sstg_graph_build(z, z, z, 1, 8, 4, z, z, 1, z) == -1            # null pointers
    assert lib.sstg_graph_build(z, z, z, 1, 8, 1000, z, z, 1, z) == -2          # scene too big
    assert lib.sstg_graph_build(z, z, z, 0, 8, 4, z, z, 1, z) == 0              # empty batch is a no-op
    assert lib.sstg_stgcn_forward(z, z, z, z, 1, 4, z, z, z, 0, z) == -1
    assert lib.sstg_txpcnn_forward(z, z, z, 0, 4, 0, z, z, z) == 0
    assert lib.sstg_sample_ade_fde(z, z, z, z, 1, 4, 0, z, 0, z, z, z, z, 0, z) == -1   # K < 1
    assert lib.sstg_error_string(-2).decode().startswith("scene larger")


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under social_stgcnn_b200/ may reference it."""
    pkg = os.path.join(ROOT, "social_stgcnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_cpu_tensors_are_refused(lib):
    import torch
    import social_stgcnn_b200 as sstg
    with pytest.raises(RuntimeError, match="CUDA"):
        sstg.InferenceEngine(np.zeros(7563, np.float32), np.zeros(30, np.float32), "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        sstg.ScenePlan([2, 3]).on("cpu")


def test_peer_mailbox_and_workspace_host_arithmetic(lib):
    """Host-only entry points of the sharded-training exchange (no GPU needed)."""
    # mailbox = 2 parities x (world slots of 7,564 floats + stats rows) + flag words
    b = lib.sstg_peer_mailbox_bytes(2, 128)
    assert b == (2 * (2 * 7564 + 128 * 30) + 16) * 4
    assert lib.sstg_peer_mailbox_bytes(8, 0) == (2 * 8 * 7564 + 16) * 4
    assert lib.sstg_peer_mailbox_bytes(0, 4) < 0 and lib.sstg_peer_mailbox_bytes(9, 4) < 0
    assert lib.sstg_peer_mailbox_bytes(2, -1) < 0
    # the statistics block of sstg_train_step's workspace lies inside it, 16-byte aligned
    total = lib.sstg_train_step_workspace(12, 5, 100, 4000)
    off = lib.sstg_train_step_stats_offset(12, 5, 100, 4000)
    assert 0 < off < total and off % 4 == 0 and off + 12 * 30 <= total
    assert lib.sstg_train_step_stats_offset(-1, 0, 0, 0) < 0
    # argument validation happens before anything touches the device
    assert lib.sstg_peer_alloc(0, None, None) == -1
    assert lib.sstg_peer_open(None, None) == -1
    assert lib.sstg_peer_close(None) == 0 and lib.sstg_peer_free(None) == 0
    assert lib.sstg_peer_exchange_apply(None, 0, 0, 1, 0, None, None, None, 0, 0, None, 0, None, None, 0.0, 0.0, 0.1,
                                        None, None, None, None, 0, None) == -1
