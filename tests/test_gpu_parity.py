"""GPU parity tests (run with -m gpu on a B200): every kernel, called through the C ABI
(ctypes -> libsstg.so), against the committed golden vectors of the real reference and
against the oracle on the same seeded inputs.

Tolerances: integer/index outputs bit-exact; fp32 outputs within 1e-5 relative
(BASELINE.json north_star), scaled by the tensor's max magnitude; the fp64-island
modes are checked bit-exactly where the reference itself is reproducible."""
import numpy as np
import pytest
import torch

from tests.common import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def batch(golden, dev):
    import social_stgcnn_b200 as sstg
    names, n_peds, obs_rel, pred_rel, last_obs = golden.pack()
    plan = sstg.ScenePlan(n_peds)
    t = lambda a: torch.from_numpy(a).to(dev)
    return dict(names=names, plan=plan, dp=plan.on(dev), obs_rel=t(obs_rel), pred_rel=t(pred_rel),
                last_obs=t(last_obs), n_peds=n_peds)


def _scene_A(A, plan, i, T=8):
    n = int(plan.n_peds[i])
    o = int(plan.a_off[i])
    return A[o:o + T * n * n].reshape(T, n, n)


def test_graph_fp64_bit_exact(golden, batch, dev):
    """seq_to_graph in fp64-island mode == networkx, bit for bit, on every golden scene
    (ragged batch N = 2..64 incl. the zero-distance case)."""
    import social_stgcnn_b200 as sstg
    plan = batch["plan"]
    V, A = sstg.graph_build(batch["dp"], batch["obs_rel"], 8, fp64=True)
    torch.cuda.synchronize()
    V, A = V.cpu().numpy(), A.cpu().numpy()
    n_el = 0
    for i, name in enumerate(batch["names"]):
        n = int(plan.n_peds[i])
        v = V[plan.seg_off[i] * 16:(plan.seg_off[i] + n) * 16].reshape(8, n, 2)
        assert np.array_equal(v, golden.graph[f"{name}.V_obs"]), name
        a_ref = golden.graph[f"{name}.A_obs"]
        a = _scene_A(A, plan, i)
        assert np.array_equal(a.view(np.uint32), a_ref.view(np.uint32)), (name, np.abs(a - a_ref).max())
        n_el += a.size
    assert n_el == 75040


def test_graph_pred_half_T12(golden, batch, dev):
    """The pred-half graph (utils.py:209) with T = 12, bit-exact in fp64 mode."""
    import social_stgcnn_b200 as sstg
    plan12 = sstg.ScenePlan(batch["n_peds"], T=12)
    V, A = sstg.graph_build(plan12.on(dev), batch["pred_rel"], 12, fp64=True)
    torch.cuda.synchronize()
    V, A = V.cpu().numpy(), A.cpu().numpy()
    for i, name in enumerate(batch["names"]):
        n = int(plan12.n_peds[i])
        a = _scene_A(A, plan12, i, 12)
        a_ref = golden.graph[f"{name}.A_tr"]
        assert np.array_equal(a.view(np.uint32), a_ref.view(np.uint32)), name
        v = V[plan12.seg_off[i] * 24:(plan12.seg_off[i] + n) * 24].reshape(12, n, 2)
        assert np.array_equal(v, golden.graph[f"{name}.V_tr"]), name


def test_graph_fp32_within_tolerance(golden, batch, dev):
    """Default fp32 mode: every element within 2e-6 of the reference (bar: 1e-5),
    relative to the element itself (floor 1e-3 for the near-zero diagonal of sparse rows)."""
    import social_stgcnn_b200 as sstg
    plan = batch["plan"]
    _, A = sstg.graph_build(batch["dp"], batch["obs_rel"], 8, fp64=False)
    torch.cuda.synchronize()
    A = A.cpu().numpy()
    worst = 0.0
    for i, name in enumerate(batch["names"]):
        a_ref = golden.graph[f"{name}.A_obs"].astype(np.float64)
        a = _scene_A(A, plan, i).astype(np.float64)
        assert np.array_equal(a == 0, a_ref == 0), name
        err = np.abs(a - a_ref) / np.maximum(np.abs(a_ref), 1e-3)
        worst = max(worst, err.max())
    assert worst < 2e-6, worst


def test_graph_no_laplacian(golden, batch, dev):
    """norm_lap_matr=False returns the raw inverse-distance adjacency (utils.py:43-47)."""
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    plan = batch["plan"]
    _, A = sstg.graph_build(batch["dp"], batch["obs_rel"], 8, norm_lap=False, fp64=True)
    torch.cuda.synchronize()
    A = A.cpu().numpy()
    for i, name in list(enumerate(batch["names"]))[::9]:
        _, w = O.seq_to_graph(golden.graph[f"{name}.obs_rel"], norm_lap_matr=False)
        assert np.array_equal(_scene_A(A, plan, i), w), name


@pytest.mark.parametrize("ckpt", ["eth", "hotel", "univ", "zara1", "zara2"])
def test_forward_eval_vs_reference(golden, batch, dev, ckpt):
    """st_gcn + TXP-CNN on the reference's own A (fp64-mode graph == reference bitwise):
    V_pred within 1e-5 of the reference's eval forward for every golden scene."""
    import social_stgcnn_b200 as sstg
    plan, dp = batch["plan"], batch["dp"]
    params, bn = golden.params(ckpt)
    p, b = torch.from_numpy(params).to(dev), torch.from_numpy(bn).to(dev)
    V, A = sstg.graph_build(dp, batch["obs_rel"], 8, fp64=True)
    H = sstg.stgcn_forward(dp, V, A, p, b)
    out = sstg.txpcnn_forward(dp, H, p)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    worst = 0.0
    for i, name in enumerate(batch["names"]):
        n = int(plan.n_peds[i])
        o = out[plan.seg_off[i] * 60:(plan.seg_off[i] + n) * 60].reshape(1, 5, 12, n)
        ref = golden.fwd[f"{ckpt}.{name}.eval"]
        worst = max(worst, rel_err(o, ref))
    assert worst < TOL, worst


def test_stgcn_block_vs_oracle(golden, batch, dev):
    """The st_gcn kernel alone (H) against the oracle's st_gcn block, per scene."""
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    plan, dp = batch["plan"], batch["dp"]
    params, bn = golden.params("eth")
    P, S = O.unflatten_params(params), O.unflatten_bn(bn)
    V, A = sstg.graph_build(dp, batch["obs_rel"], 8, fp64=True)
    H = sstg.stgcn_forward(dp, V, A, torch.from_numpy(params).to(dev), torch.from_numpy(bn).to(dev))
    torch.cuda.synchronize()
    H = H.cpu().numpy()
    worst = 0.0
    for i, name in enumerate(batch["names"]):
        n = int(plan.n_peds[i])
        v = torch.from_numpy(golden.graph[f"{name}.V_obs"])[None].permute(0, 3, 1, 2)
        with torch.no_grad():
            ref = O.stgcn_forward(P, S, v, torch.from_numpy(golden.graph[f"{name}.A_obs"])).numpy()
        h = H[plan.seg_off[i] * 40:(plan.seg_off[i] + n) * 40].reshape(1, 5, 8, n)
        worst = max(worst, rel_err(h, ref))
    assert worst < TOL, worst


@pytest.mark.parametrize("counts", [
    [1, 2, 3, 4, 5, 6, 7, 8, 9, 30, 31, 32],            # one column per lane; row tails of 1..3 after the row quads
    [33, 62, 63, 64, 64, 65 - 2, 34, 47, 64, 2, 61],    # two columns per lane, two CTAs per SM
    [65, 66, 96, 127, 128, 3, 100],                     # four columns per lane, several ring chunks per slice
    [129, 130, 200, 255, 256, 256, 131],                # eight columns per lane, 4-row chunks at N = 256
])
def test_stgcn_kernel_shapes_vs_oracle(golden, dev, counts):
    """The st_gcn kernel on synthetic ragged batches that cover every template instantiation, the row-quad /
    row-tail split of the FFMA2 aggregation (N mod 4 = 0..3), equal-sized ring chunks and lanes whose columns
    lie beyond N -- against the oracle's st_gcn block (model.py:145-155 restated) on the SAME V and A.  A is
    an arbitrary dense matrix here: the block does not depend on it being a Laplacian.  Tolerance 1e-5 of
    the largest |H| of the scene (fp32: the kernel sums even and odd rows separately)."""
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    params, bn = golden.params("hotel")
    P, S = O.unflatten_params(params), O.unflatten_bn(bn)
    n = np.asarray(counts, np.int32)
    plan = sstg.ScenePlan(n)
    dp = plan.on(dev)
    rng = np.random.default_rng(int(n.sum()))
    V = (rng.standard_normal(plan.P * 16) * 0.5).astype(np.float32)
    A = np.zeros(plan.a_floats, np.float32)
    blocks = []
    for i, k in enumerate(n):
        a = (rng.standard_normal((8, k, k)) / np.sqrt(k)).astype(np.float32)
        A[plan.a_off[i]:plan.a_off[i] + 8 * k * k] = a.ravel()
        blocks.append(a)
    H = sstg.stgcn_forward(dp, torch.from_numpy(V).to(dev), torch.from_numpy(A).to(dev),
                           torch.from_numpy(params).to(dev), torch.from_numpy(bn).to(dev))
    torch.cuda.synchronize()
    H = H.cpu().numpy()
    for i, k in enumerate(n):
        o = int(plan.seg_off[i])
        v = torch.from_numpy(V[o * 16:(o + k) * 16].reshape(8, k, 2))[None].permute(0, 3, 1, 2)
        with torch.no_grad():
            ref = O.stgcn_forward(P, S, v, torch.from_numpy(blocks[i])).numpy()
        h = H[o * 40:(o + k) * 40].reshape(1, 5, 8, k)
        assert np.isfinite(h).all()
        assert rel_err(h, ref) < TOL, (i, int(k), rel_err(h, ref))


def test_nll_forward(golden, batch, dev):
    """bivariate_loss per scene against the reference's eval NLL (loss_golden)."""
    import social_stgcnn_b200 as sstg
    dp = batch["dp"]
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    r = eng.infer(batch["plan"], batch["obs_rel"], batch["pred_rel"], batch["last_obs"], graph_fp64=True)
    loss = sstg.nll_forward(dp, r["V_pred"], r["V_tr"])
    torch.cuda.synchronize()
    ref = golden.loss["eth.nll_eval"]
    got = loss.cpu().numpy().astype(np.float64)
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < TOL, err.max()
    assert abs(got[:70].mean() - 0.84885728) < 1e-5


def test_sample_ade_fde_injected_noise(golden, batch, dev):
    """Whole pipeline with the reference's injected noise against test()'s own numbers.

    ADE/FDE are differences of fp32 ABSOLUTE positions (metrics.py:69-75 stores them in
    fp32), so their noise floor is one fp32 ulp of the position magnitude, not of the ADE
    value: a 1e-6 relative change in V_pred (conv summation order) moves a position by
    one ulp.  Bars: sampled absolute trajectories within 1e-5 relative; per-pedestrian
    min ADE/FDE within 1e-5 x the scene's position scale (and within 1e-4 of their own
    value); dataset-level eth ADE/FDE within 1e-5 relative."""
    import social_stgcnn_b200 as sstg
    plan = batch["plan"]
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    eps = np.concatenate([golden.sample[f"{n}.eps"].reshape(-1) for n in batch["names"]]).astype(np.float32)
    r = eng.infer(plan, batch["obs_rel"], batch["pred_rel"], batch["last_obs"], K=20,
                  eps=torch.from_numpy(eps).to(dev), graph_fp64=True, fp64_dist=True, want_pred=True)
    torch.cuda.synchronize()
    ade, fde = r["min_ade"].cpu().numpy(), r["min_fde"].cpu().numpy()
    pred = r["pred_abs"].cpu().numpy()
    worst_pos = worst_own = 0.0
    tot_a = tot_f = 0.0
    tot_n = 0
    for i, name in enumerate(batch["names"]):
        n, o = int(plan.n_peds[i]), int(plan.seg_off[i])
        ra, rf = golden.sample[f"{name}.min_ade"], golden.sample[f"{name}.min_fde"]
        got = pred[o * 480:(o + n) * 480].reshape(20, 12, n, 2)
        scale = max(1.0, float(np.abs(got).max()))
        da, df = np.abs(ade[o:o + n] - ra), np.abs(fde[o:o + n] - rf)
        worst_pos = max(worst_pos, da.max() / scale, df.max() / scale)
        worst_own = max(worst_own, (da / ra).max(), (df / rf).max())
        if f"{name}.pred_abs" in golden.sample:
            assert rel_err(got, golden.sample[f"{name}.pred_abs"]) < TOL, name
        if name.startswith("eth"):
            tot_a += ade[o:o + n].sum(); tot_f += fde[o:o + n].sum(); tot_n += n
    assert worst_pos < TOL, worst_pos
    assert worst_own < 1e-3, worst_own   # tiny near-hit FDEs: same absolute error, larger relative
    assert abs(tot_a / tot_n - golden.sample["eth.ade"][0]) < TOL * golden.sample["eth.ade"][0]
    assert abs(tot_f / tot_n - golden.sample["eth.fde"][0]) < TOL * golden.sample["eth.fde"][0]


def test_sample_kernel_vs_oracle_same_inputs(golden, batch, dev):
    """The sampling/metric kernel in isolation: fed the oracle-visible V_pred it produced,
    fp64-distance mode must match the oracle to 1e-12 (same rounding points)."""
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    plan = batch["plan"]
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    r = eng.infer(plan, batch["obs_rel"], batch["pred_rel"], batch["last_obs"], K=20, seed=7,
                  graph_fp64=True, fp64_dist=True, want_eps=True)
    torch.cuda.synchronize()
    vp, vtr = r["V_pred"].cpu().numpy(), r["V_tr"].cpu().numpy()
    eps = r["eps"].cpu().numpy()
    ade, fde = r["min_ade"].cpu().numpy(), r["min_fde"].cpu().numpy()
    last = batch["last_obs"].cpu().numpy()
    for i, name in list(enumerate(batch["names"]))[::5] + [(len(batch["names"]) - 2, batch["names"][-2])]:
        n, o = int(plan.n_peds[i]), int(plan.seg_off[i])
        v = vp[o * 60:(o + n) * 60].reshape(5, 12, n).transpose(1, 2, 0)
        t = vtr[o * 24:(o + n) * 24].reshape(12, n, 2)
        e = eps[o * 480:(o + n) * 480].reshape(20, 12, n, 2)
        oa, of, _ = O.sample_ade_fde(v, t, last[o:o + n], e)
        # exp/tanh differ by an ulp between CUDA libm and torch CPU: 1e-6, not bitwise
        assert np.abs(ade[o:o + n] - oa).max() <= 2e-6 * max(1.0, oa.max()), name
        assert np.abs(fde[o:o + n] - of).max() <= 2e-6 * max(1.0, of.max()), name


def test_philox_noise_statistics_and_replay(golden, batch, dev):
    """In-kernel Philox4x32-10 + Box-Muller: the noise the kernel reports (eps_out) is
    standard normal, differs per seed, and matches a numpy Philox replica."""
    import social_stgcnn_b200 as sstg
    from tests.philox_ref import philox_normals
    plan = batch["plan"]
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    r1 = eng.infer(plan, batch["obs_rel"], batch["pred_rel"], batch["last_obs"], K=20, seed=1234, want_eps=True)
    e1 = r1["eps"].cpu().numpy().copy()
    r2 = eng.infer(plan, batch["obs_rel"], batch["pred_rel"], batch["last_obs"], K=20, seed=99, want_eps=True)
    e2 = r2["eps"].cpu().numpy()
    assert abs(e1.mean()) < 0.01 and abs(e1.std() - 1.0) < 0.01
    assert np.abs(e1 - e2).max() > 1.0
    # replica: scene 0 (N = n0) and the last scene
    for i in (0, len(batch["names"]) - 1):
        n, o = int(plan.n_peds[i]), int(plan.seg_off[i])
        ref = philox_normals(1234, np.arange(o, o + n), 20)          # (K,12,n,2)
        got = e1[o * 480:(o + n) * 480].reshape(20, 12, n, 2)
        assert np.abs(got - ref).max() < 2e-4


def test_infer_host_pipelined_chunks_equal_direct_calls(golden, batch, dev):
    """InferenceEngine.infer_host (host buffers, 3-stream software pipeline over chunks of scenes)
    returns exactly what direct device-resident calls on the same chunks return."""
    import social_stgcnn_b200 as sstg
    plan = batch["plan"]
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    obs_h = batch["obs_rel"].cpu().pin_memory()
    pred_h = batch["pred_rel"].cpu().pin_memory()
    last_h = batch["last_obs"].cpu().pin_memory()
    for chunk, fused in ((7, True), (32, False), (10 ** 6, True)):
        a_h, f_h = eng.infer_host(plan, obs_h, pred_h, last_h, 20, chunk_scenes=chunk, seed=5, fused=fused)
        a_h, f_h = a_h.copy(), f_h.copy()
        for i, (lo, hi, p_lo, p_hi, sub) in enumerate(plan.chunks(chunk, taper=True)):
            r = eng.infer(sub, batch["obs_rel"][p_lo:p_hi].contiguous(), batch["pred_rel"][p_lo:p_hi].contiguous(),
                          batch["last_obs"][p_lo:p_hi].contiguous(), 20, seed=5 + 0x9E3779B1 * i, fused=fused)
            torch.cuda.synchronize()
            assert np.array_equal(a_h[p_lo:p_hi], r["min_ade"].cpu().numpy()), (chunk, i)
            assert np.array_equal(f_h[p_lo:p_hi], r["min_fde"].cpu().numpy()), (chunk, i)
    assert len(plan.chunks(7)) == (plan.S + 6) // 7


@pytest.mark.parametrize("ckpt", ["eth", "hotel", "univ", "zara1", "zara2"])
def test_fused_graph_stgcn_front(golden, batch, dev, ckpt):
    """sstg_stgcn_forward_fused (graph + st_gcn with the adjacency kept in shared memory) against
    the unfused pair on the same inputs and, through the TXP-CNN, against the reference's V_pred."""
    import social_stgcnn_b200 as sstg
    plan, dp = batch["plan"], batch["dp"]
    params, bn = golden.params(ckpt)
    p, b = torch.from_numpy(params).to(dev), torch.from_numpy(bn).to(dev)
    V, A = sstg.graph_build(dp, batch["obs_rel"], 8, fp64=True)
    H_ref = sstg.stgcn_forward(dp, V, A, p, b)
    H = sstg.stgcn_forward_fused(dp, batch["obs_rel"], p, b)
    out = sstg.txpcnn_forward(dp, H, p)
    torch.cuda.synchronize()
    H, H_ref, out = H.cpu().numpy(), H_ref.cpu().numpy(), out.cpu().numpy()
    worst_h = worst = 0.0
    for i, name in enumerate(batch["names"]):
        n, o = int(plan.n_peds[i]), int(plan.seg_off[i])
        worst_h = max(worst_h, rel_err(H[o * 40:(o + n) * 40], H_ref[o * 40:(o + n) * 40]))
        got = out[o * 60:(o + n) * 60].reshape(1, 5, 12, n)
        worst = max(worst, rel_err(got, golden.fwd[f"{ckpt}.{name}.eval"]))
    assert worst_h < TOL, worst_h
    assert worst < TOL, worst
    big = sstg.ScenePlan([65, 3])
    with pytest.raises(ValueError):
        sstg.stgcn_forward_fused(big.on(dev), torch.zeros(68, 2, 8, device=dev), p, b)


def test_edge_cases_single_pedestrian_odd_T_and_empty(golden, dev):
    """Edges the packed layout must survive: scenes of ONE pedestrian (A = [[0]]: d = 1), scenes of two,
    mixed with a 64-pedestrian scene; a graph with T = 5 (general kernel, T % 4 != 0); an empty batch."""
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    n_list = [1, 2, 1, 64, 1, 3]
    obs_rel, pred_rel, obs_abs = O.synthetic_scenes(1, sum(n_list), seed=31)
    last = np.ascontiguousarray(obs_abs[:, :, -1])
    plan = sstg.ScenePlan(n_list)
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    P_, S_ = O.unflatten_params(params), O.unflatten_bn(bn)
    for fused in (False, True):
        r = eng.infer(plan, t(obs_rel), t(pred_rel), t(last), K=20, seed=2, fused=fused, fp64_dist=True, want_eps=True,
                      graph_fp64=not fused)
        torch.cuda.synchronize()
        vp, eps = r["V_pred"].cpu().numpy(), r["eps"].cpu().numpy()
        ade = r["min_ade"].cpu().numpy()
        for i, n in enumerate(n_list):
            o = int(plan.seg_off[i])
            V_ref, A_ref = O.seq_to_graph(obs_rel[o:o + n])
            if not fused:
                a = r["A_obs"].cpu().numpy()[plan.a_off[i]:plan.a_off[i] + 8 * n * n].reshape(8, n, n)
                assert np.array_equal(a.view(np.uint32), A_ref.view(np.uint32)), n
                if n == 1:
                    assert np.all(a == 0)
            with torch.no_grad():
                out = O.model_forward(P_, S_, V_ref, A_ref).numpy()
            got = vp[o * 60:(o + n) * 60].reshape(1, 5, 12, n)
            assert rel_err(got, out) < TOL, (n, fused)
            vt = np.ascontiguousarray(pred_rel[o:o + n].transpose(2, 0, 1))
            oa, _, _ = O.sample_ade_fde(got[0].transpose(1, 2, 0), vt, last[o:o + n],
                                        eps[o * 480:(o + n) * 480].reshape(20, 12, n, 2))
            assert np.abs(ade[o:o + n] - oa).max() < 2e-6 * max(1.0, oa.max())
    # T = 5 graph (neither 8 nor 12): the general kernel, bit-exact in fp64 mode, close in fp32 mode
    rel5 = np.ascontiguousarray(obs_rel[:, :, :5])
    p5 = sstg.ScenePlan(n_list, T=5)
    for fp64 in (True, False):
        V, A = sstg.graph_build(p5.on(dev), t(rel5), 5, fp64=fp64)
        torch.cuda.synchronize()
        A = A.cpu().numpy()
        for i, n in enumerate(n_list):
            o = int(p5.seg_off[i])
            _, A_ref = O.seq_to_graph(rel5[o:o + n])
            a = A[p5.a_off[i]:p5.a_off[i] + 5 * n * n].reshape(5, n, n)
            if fp64:
                assert np.array_equal(a.view(np.uint32), A_ref.view(np.uint32)), n
            else:
                assert np.abs(a - A_ref).max() < 2e-6
    # empty batch: every entry point is a no-op
    e = sstg.ScenePlan([])
    z = torch.zeros(0, 2, 8, device=dev)
    V, A = sstg.graph_build(e.on(dev), z, 8)
    H = sstg.stgcn_forward(e.on(dev), V, torch.zeros(4, device=dev), t(params), t(bn))
    out = sstg.txpcnn_forward(e.on(dev), H, t(params))
    assert V.numel() == 0 and out.numel() == 0
