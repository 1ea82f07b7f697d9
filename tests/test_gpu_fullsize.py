"""GPU parity at BASELINE.json's full sizes (10^5 scenes, 10^6..10^7 pedestrians), where the
oracle cannot run everything: size-independent properties plus oracle spot checks.

  * replication: a ragged batch made of R copies of the 77 golden scenes must give R bit-identical
    copies of every tensor (V, A, H, V_pred, NLL, injected-noise ADE/FDE) -- catches any offset /
    overflow / tiling error that only shows far into a large packed batch -- and copy 0 is
    checked against the reference's golden vectors;
  * dense crowds (32,768 scenes x 64): Laplacian identities (symmetry, D^1/2 in the null space),
    finiteness, min-over-K monotonicity (the k-th Philox sample does not depend on K), chunked
    host pipeline == one-shot pipeline, and oracle spot checks on random scenes."""
import os

import numpy as np
import pytest
import torch

from tests.common import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def _rows_equal_first(x, R):
    x = x.view(R, -1)
    return bool((x == x[0:1]).all())


def test_replicated_ragged_batch_bitwise(golden, dev):
    import social_stgcnn_b200 as sstg
    R = 1400                                                    # 107,800 scenes, 520,800 pedestrians
    names, n_peds, obs_rel, pred_rel, last_obs = golden.pack()
    S0, P0 = len(names), int(n_peds.sum())
    plan = sstg.ScenePlan(np.tile(n_peds, R))
    assert plan.S == R * S0 and plan.P == R * P0
    t = lambda a: torch.from_numpy(np.tile(a, (R,) + (1,) * (a.ndim - 1))).to(dev)
    obs, pred, last = t(obs_rel), t(pred_rel), t(last_obs)
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    eps0 = np.concatenate([golden.sample[f"{n}.eps"].reshape(-1) for n in names]).astype(np.float32)
    eps = torch.from_numpy(eps0).to(dev).repeat(R)
    for exact in (True, False):
        r = eng.infer(plan, obs, pred, last, K=20, eps=eps, graph_fp64=exact, fp64_dist=exact)
        loss = sstg.nll_forward(plan.on(dev), r["V_pred"], r["V_tr"])
        torch.cuda.synchronize()
        assert plan.a_floats % R == 0
        for k in ("V_obs", "A_obs", "V_tr", "H", "V_pred", "min_ade", "min_fde"):
            assert torch.isfinite(r[k]).all(), k
            assert _rows_equal_first(r[k], R), (k, exact)
        assert _rows_equal_first(loss, R)
        # copy 0 (and therefore every copy) against the reference's golden vectors
        vp = r["V_pred"][:P0 * 60].cpu().numpy()
        A = r["A_obs"][:plan.a_floats // R].cpu().numpy()
        ade = r["min_ade"][:P0].cpu().numpy()
        for i in (0, 33, 70, 75, 76):
            name, n, o = names[i], int(n_peds[i]), int(plan.seg_off[i])
            got = vp[o * 60:(o + n) * 60].reshape(1, 5, 12, n)
            assert rel_err(got, golden.fwd[f"eth.{name}.eval"]) < 1e-5, name
            a = A[plan.a_off[i]:plan.a_off[i] + 8 * n * n].reshape(8, n, n)
            ref = golden.graph[f"{name}.A_obs"]
            if exact:
                assert np.array_equal(a.view(np.uint32), ref.view(np.uint32)), name
            else:
                assert np.abs(a - ref).max() < 2e-6
            scale = max(1.0, np.abs(golden.graph[f"{name}.obs"]).max())
            assert np.abs(ade[o:o + n] - golden.sample[f"{name}.min_ade"]).max() < 2e-5 * scale, name
    ref_l = golden.loss["eth.nll_eval"]
    got_l = loss[:S0].cpu().numpy()
    assert (np.abs(got_l - ref_l) / np.maximum(np.abs(ref_l), 1.0)).max() < 1e-5


def test_dense_crowds_properties(golden, dev):
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    S, N = 32768, 64
    obs_rel, pred_rel, obs_abs = O.synthetic_scenes(S, N, seed=2024)
    last = np.ascontiguousarray(obs_abs[:, :, -1])
    plan = sstg.ScenePlan(np.full(S, N, np.int32))
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    obs, pred, lst = (torch.from_numpy(a).to(dev) for a in (obs_rel, pred_rel, last))
    r = eng.infer(plan, obs, pred, lst, K=20, seed=9)
    torch.cuda.synchronize()
    A = r["A_obs"].view(S, 8, N, N)
    for k in ("A_obs", "H", "V_pred", "min_ade", "min_fde"):
        assert torch.isfinite(r[k]).all(), k
    # symmetric normalised Laplacian: A == A^T (to an ulp), 0 <= A_hh < 1, and D^1/2 1 is in its null
    # space, with the degrees d_h taken from the raw adjacency W (norm_lap_matr=False) summed in fp64
    assert float((A - A.transpose(2, 3)).abs().max()) < 1e-6
    diag = torch.diagonal(A, dim1=2, dim2=3)
    assert float(diag.min()) >= 0.0 and float(diag.max()) < 1.0
    _, W = sstg.graph_build(plan.on(dev), obs, 8, norm_lap=False, fp64=True)
    sqrt_d = W.view(S, 8, N, N).double().sum(3).sqrt()
    resid = torch.einsum("sthk,stk->sth", A.double(), sqrt_d)
    bound = torch.einsum("sthk,stk->sth", A.double().abs(), sqrt_d)
    assert bool((resid.abs() <= 1e-5 * bound + 1e-9).all()), float((resid.abs() / (bound + 1e-9)).max())
    del W, resid, bound
    assert float(A[:, 0].abs().max()) == 0.0                       # frame 0: all displacements are zero (utils.py:153-155)
    ade20, fde20 = r["min_ade"].clone(), r["min_fde"].clone()
    # min over K is monotone in K: sample k is the same Philox draw whatever K is
    r10 = eng.infer(plan, obs, pred, lst, K=10, seed=9)
    torch.cuda.synchronize()
    assert bool((ade20 <= r10["min_ade"]).all()) and bool((fde20 <= r10["min_fde"]).all())
    assert float((r10["min_ade"] - ade20).max()) > 0.0
    # the chunked host pipeline reproduces one-shot runs of the same chunks (seed + 0x9E3779B1 * chunk)
    a_h, f_h = eng.infer_host(plan, torch.from_numpy(obs_rel).pin_memory(), torch.from_numpy(pred_rel).pin_memory(),
                              torch.from_numpy(last).pin_memory(), 20, chunk_scenes=S, seed=9, fused=False)
    assert np.array_equal(a_h, ade20.cpu().numpy()) and np.array_equal(f_h, fde20.cpu().numpy())
    # ... and the fused front (adjacency kept on chip) moves no per-pedestrian minimum by more than 1e-5 of
    # the position scale: ADE/FDE are continuous in V_pred, and V_pred changes by rounding only
    rfu = eng.infer(plan, obs, pred, lst, K=20, seed=9, fused=True)
    torch.cuda.synchronize()
    assert rfu["A_obs"] is None
    scale = float(np.abs(obs_abs).max())
    assert float((rfu["min_ade"] - ade20).abs().max()) < 1e-5 * scale
    assert float((rfu["min_fde"] - fde20).abs().max()) < 1e-5 * scale
    # oracle spot checks in exact mode, replaying the kernel's own noise
    rx = eng.infer(plan, obs, pred, lst, K=20, seed=9, graph_fp64=True, fp64_dist=True, want_eps=True)
    torch.cuda.synchronize()
    P_, S_ = O.unflatten_params(params), O.unflatten_bn(bn)
    Ax, vp = rx["A_obs"].view(S, 8, N, N), rx["V_pred"].view(S, 5, 12, N)
    eps = rx["eps"].view(S, 20, 12, N, 2)
    for s in (0, 12345, S - 1):
        sl = slice(s * N, (s + 1) * N)
        V_ref, A_ref = O.seq_to_graph(obs_rel[sl])
        assert np.array_equal(Ax[s].cpu().numpy().view(np.uint32), A_ref.view(np.uint32)), s
        with torch.no_grad():
            out = O.model_forward(P_, S_, V_ref, A_ref)
        got = vp[s].cpu().numpy()
        assert rel_err(got[None], out.numpy()) < 1e-5, s
        vt = np.ascontiguousarray(pred_rel[sl].transpose(2, 0, 1))
        oa, of, _ = O.sample_ade_fde(got.transpose(1, 2, 0), vt, last[sl], eps[s].cpu().numpy())
        assert np.abs(rx["min_ade"][sl].cpu().numpy() - oa).max() < 2e-6 * max(1.0, oa.max()), s
        assert np.abs(rx["min_fde"][sl].cpu().numpy() - of).max() < 2e-6 * max(1.0, of.max()), s


def test_full_shard_of_config3_is_consistent_with_small_batches(golden, dev):
    """BASELINE config 3 at the bench's full per-GPU size (131,072 scenes x 64 pedestrians: 8.4 M pedestrians, a 17 GB
    adjacency buffer of 2^32 floats): the adjacency, H and V_pred of scenes at the start, in the middle
    (around the 2^31-float mark of A, where a 32-bit signed element index would wrap) and at the end of the batch are BIT-IDENTICAL to what the same scenes give as a
    batch of their own -- scenes are independent, so any offset overflow or cross-scene leak would show -- and
    the per-pedestrian minima are finite, ordered (ADE <= FDE is not required, but both > 0) and monotone in K."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 40e9:
        pytest.skip("needs ~40 GB of free device memory")
    S, N = 131072, 64
    g = torch.Generator().manual_seed(77)
    P = S * N
    vel = torch.rand(P, 2, 1, generator=g) - 0.5
    obs = vel + 0.1 * torch.randn(P, 2, 8, generator=g)
    obs[:, :, 0] = 0
    pred = vel + 0.1 * torch.randn(P, 2, 12, generator=g)
    last = torch.rand(P, 2, generator=g) * 20 - 10
    obs, pred, last = obs.to(dev), pred.to(dev), last.to(dev)
    params, bn = golden.params("zara1")
    eng = sstg.InferenceEngine(params, bn, dev)
    plan = sstg.ScenePlan(np.full(S, N, np.int32))
    assert plan.a_floats == S * 8 * N * N == 2 ** 32
    r = eng.infer(plan, obs, pred, last, K=20, seed=3, want_vtr=False)
    torch.cuda.synchronize()
    for k in ("H", "V_pred", "min_ade", "min_fde"):
        assert torch.isfinite(r[k]).all(), k
    assert float(r["min_ade"].min()) > 0.0 and float(r["min_fde"].min()) > 0.0
    ade20 = r["min_ade"].clone()
    mid = (2 ** 31) // (8 * N * N) - 128          # the scenes whose adjacency blocks straddle float offset 2^31
    eng2 = sstg.InferenceEngine(params, bn, dev)
    for lo in (0, mid, S - 256):
        sub = sstg.ScenePlan(np.full(256, N, np.int32))
        sl = slice(lo * N, (lo + 256) * N)
        dps = sub.on(dev)
        V, A = E.graph_build(dps, obs[sl].contiguous(), 8)
        H = E.stgcn_forward(dps, V, A, eng2.params, eng2.bn)
        vp = E.txpcnn_forward(dps, H, eng2.params)
        torch.cuda.synchronize()
        a_big = r["A_obs"][lo * 8 * N * N:(lo + 256) * 8 * N * N]
        assert torch.equal(a_big.view(torch.int32), A.view(torch.int32)), lo
        assert torch.equal(r["H"][lo * N * 40:(lo + 256) * N * 40].view(torch.int32), H.view(torch.int32)), lo
        assert torch.equal(r["V_pred"][lo * N * 60:(lo + 256) * N * 60].view(torch.int32), vp.view(torch.int32)), lo
    r10 = eng.infer(plan, obs, pred, last, K=10, seed=3, want_vtr=False)
    torch.cuda.synchronize()
    assert bool((ade20 <= r10["min_ade"]).all())


def test_large_scenes_256_pedestrians(golden, dev):
    """BASELINE config 4: 256 pedestrians per scene (the N x N aggregation stress; one A[t] slice is
    256 KB > shared memory, so st_gcn streams it in row chunks; TXP-CNN tiles are single scenes).
    Mixed with 200 / 129 / 128 / 65 to cross every kernel-variant boundary.  Graph bit-exact in
    fp64 mode, forward / NLL / ADE / FDE against the oracle, and one training group against the
    fp64 oracle."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E
    from oracle import stgcnn_oracle as O
    n_list = [256, 200, 129, 128, 65, 256, 3]
    obs_rel, pred_rel, obs_abs = O.synthetic_scenes(1, sum(n_list), seed=256)
    last = np.ascontiguousarray(obs_abs[:, :, -1])
    plan = sstg.ScenePlan(n_list)
    params, bn = golden.params("univ") if "univ.params" in golden.weights.files else golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    obs, pred, lst = (torch.from_numpy(a).to(dev) for a in (obs_rel, pred_rel, last))
    rx = eng.infer(plan, obs, pred, lst, K=20, seed=4, graph_fp64=True, fp64_dist=True, want_eps=True)
    loss = sstg.nll_forward(plan.on(dev), rx["V_pred"], rx["V_tr"])
    rf = eng.infer(plan, obs, pred, lst, K=20, seed=4)          # fast modes, separate call (workspace reuse)
    torch.cuda.synchronize()
    A_fast = rf["A_obs"].cpu().numpy()
    rx2 = eng.infer(plan, obs, pred, lst, K=20, seed=4, graph_fp64=True, fp64_dist=True, want_eps=True)
    torch.cuda.synchronize()
    A, vp, eps = rx2["A_obs"].cpu().numpy(), rx2["V_pred"].cpu().numpy(), rx2["eps"].cpu().numpy()
    ade, fde = rx2["min_ade"].cpu().numpy(), rx2["min_fde"].cpu().numpy()
    P_, S_ = O.unflatten_params(params), O.unflatten_bn(bn)
    for i, n in enumerate(n_list):
        o, ao = int(plan.seg_off[i]), int(plan.a_off[i])
        V_ref, A_ref = O.seq_to_graph(obs_rel[o:o + n])
        a = A[ao:ao + 8 * n * n].reshape(8, n, n)
        assert np.array_equal(a.view(np.uint32), A_ref.view(np.uint32)), n
        af = A_fast[ao:ao + 8 * n * n].reshape(8, n, n).astype(np.float64)
        assert np.array_equal(af == 0, A_ref == 0)
        assert (np.abs(af - A_ref) / np.maximum(np.abs(A_ref), 1e-3)).max() < 3e-6, n
        with torch.no_grad():
            out = O.model_forward(P_, S_, V_ref, A_ref)
        got = vp[o * 60:(o + n) * 60].reshape(1, 5, 12, n)
        assert rel_err(got, out.numpy()) < 1e-5, n
        vt = np.ascontiguousarray(pred_rel[o:o + n].transpose(2, 0, 1))
        l_ref = float(O.bivariate_loss(O.vpred_tn5(out), torch.from_numpy(vt)))
        assert abs(float(loss[i]) - l_ref) < 1e-5 * max(1.0, abs(l_ref)), n
        e = eps[o * 480:(o + n) * 480].reshape(20, 12, n, 2)
        oa, of, _ = O.sample_ade_fde(got[0].transpose(1, 2, 0), vt, last[o:o + n], e)
        assert np.abs(ade[o:o + n] - oa).max() < 2e-6 * max(1.0, oa.max()), n
        assert np.abs(fde[o:o + n] - of).max() < 2e-6 * max(1.0, of.max()), n
    # one optimiser group at these sizes against the fp64 oracle
    tr = sstg.TrainEngine(params, bn, dev)
    r = tr.forward_backward(plan, obs, pred, E.group_weights(len(n_list), 8))
    torch.cuda.synchronize()
    scenes = []
    for i, n in enumerate(n_list):
        o = int(plan.seg_off[i])
        V, A_ = O.seq_to_graph(obs_rel[o:o + n])
        scenes.append((V, A_, np.ascontiguousarray(pred_rel[o:o + n].transpose(2, 0, 1))))
    lo, go, _, _ = O.train_step(O.unflatten_params(params, torch.float64), O.unflatten_bn(bn, torch.float64), scenes, 8)
    got_loss = float((r["loss"].double() * r["weight"].double()).sum())
    assert abs(got_loss - lo) < 1e-5 * max(1.0, abs(lo))
    g = r["grad"].cpu().numpy().astype(np.float64)
    assert np.abs(g - go).max() < 5e-5 * np.abs(go).max()


def test_inference_kernels_stay_inside_their_buffers(golden, dev):
    """compute-sanitizer is not available on this pool: every output buffer of the inference kernels
    (V, A, H, the fused front's scratch, V_pred, the per-pedestrian minima) is carved out of a sentinel-filled
    arena with guard bands; after graph build, st_gcn (plain and fused), TXP-CNN (Winograd and direct) and
    sampling, every sentinel must be intact -- ragged, tile-boundary and 256-pedestrian shapes."""
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    params, bn = golden.params("eth")
    prm, bnb = torch.from_numpy(params).to(dev), torch.from_numpy(bn).to(dev)
    SENT = 0x7fc0dead
    G = 4096
    for n_list in ([2, 3, 2, 57, 5, 2, 31, 33], [64] * 40, [127, 128, 129, 2], [256, 3, 255], [2] * 300):
        plan = sstg.ScenePlan(n_list)
        dp = plan.on(dev)
        P = plan.P
        obs_rel, pred_rel, obs_abs = O.synthetic_scenes(1, P, seed=11)
        o, p, l = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (obs_rel, pred_rel, obs_abs[:, :, -1]))
        sizes = {"V": P * 16, "A": plan.a_floats, "H": P * 40, "H2": P * 40, "Pws": P * 40, "out": P * 60,
                 "out2": P * 60, "ade": 2 * P, "fde": 2 * P}                       # fp64 minima: 2 floats each
        sizes = {k: (v + 3) & ~3 for k, v in sizes.items()}
        arena = torch.empty(sum(sizes.values()) + G * (len(sizes) + 1), dtype=torch.int32, device=dev)
        arena.fill_(SENT)
        fl, buf, cur = arena.view(torch.float32), {}, G
        mask = torch.ones_like(arena, dtype=torch.bool)
        for k, sz in sizes.items():
            buf[k] = fl[cur:cur + sz]
            mask[cur:cur + sz] = False
            cur += sz + G
        V, A = sstg.graph_build(dp, o, 8, V=buf["V"][:P * 16], A=buf["A"][:plan.a_floats])
        H = sstg.stgcn_forward(dp, V, A, prm, bnb, H=buf["H"][:P * 40])
        if plan.max_n <= 64:
            sstg.stgcn_forward_fused(dp, o, prm, bnb, H=buf["H2"][:P * 40], Pws=buf["Pws"][:P * 40])
        out = sstg.txpcnn_forward(dp, H, prm, out=buf["out"][:P * 60])
        os.environ["SSTG_TXP_WINO"] = "0"                                          # the direct FFMA2 kernel as well
        try:
            sstg.txpcnn_forward(dp, H, prm, out=buf["out2"][:P * 60])
        finally:
            os.environ.pop("SSTG_TXP_WINO", None)
        sstg.sample_ade_fde(dp, out, p, l, 20, seed=1, fp64_dist=False, vtr_is_rel=True,
                            min_ade=buf["ade"][:2 * P].view(torch.float64), min_fde=buf["fde"][:2 * P].view(torch.float64))
        torch.cuda.synchronize()
        assert bool((arena[mask] == SENT).all()), n_list
        assert bool(torch.isfinite(buf["out"][:P * 60]).all()) and bool(torch.isfinite(buf["ade"][:2 * P].view(torch.float64)).all())
        assert float((buf["out"][:P * 60] - buf["out2"][:P * 60]).abs().max()) < 1e-4


def test_winograd_constant_bank_filters_equal_shared_memory_filters(golden, dev):
    """Full grids run the Winograd TXP-CNN with its transformed filters in a __constant__ bank (uniform-register
    operands), refreshed in stream order before every launch; SSTG_TXP_FILTERS=smem keeps them in shared
    memory.  Same arithmetic: the outputs must be bit-identical, also when the weights change from one launch
    to the next on the same stream."""
    import social_stgcnn_b200 as sstg
    rng = np.random.default_rng(12)
    n = np.concatenate([np.full(600, 64), rng.integers(2, 58, 4000)]).astype(np.int32)
    plan = sstg.ScenePlan(n)
    assert plan.n_tiles >= 2 * 148                      # the constant-bank variant is selected
    dp = plan.on(dev)
    H = torch.from_numpy(rng.standard_normal(plan.P * 40).astype(np.float32)).to(dev)
    outs = {}
    for ck in ("eth", "zara1"):
        prm = torch.from_numpy(golden.params(ck)[0]).to(dev)
        a = sstg.txpcnn_forward(dp, H, prm).clone()
        os.environ["SSTG_TXP_FILTERS"] = "smem"
        try:
            b = sstg.txpcnn_forward(dp, H, prm).clone()
        finally:
            os.environ.pop("SSTG_TXP_FILTERS", None)
        torch.cuda.synchronize()
        assert torch.equal(a, b), ck
        outs[ck] = (prm, a)
    assert not torch.equal(outs["eth"][1], outs["zara1"][1])
    # alternate the weights launch after launch without synchronising in between
    seq = [sstg.txpcnn_forward(dp, H, outs[ck][0]).clone() for ck in ("eth", "zara1", "eth", "zara1")]
    torch.cuda.synchronize()
    for got, ck in zip(seq, ("eth", "zara1", "eth", "zara1")):
        assert torch.equal(got, outs[ck][1]), ck


def test_winograd_for_scenes_of_up_to_256_pedestrians_equals_the_direct_kernel(golden, dev):
    """Batches with scenes of 129..256 pedestrians run the Winograd TXP-CNN in its 128-pair form (256-thread CTAs; a scene
    of more than 128 columns is a tile of its own, smaller scenes still share tiles): the constant-bank and the
    shared-memory filter variants must be bit-identical, and both within 1e-5 of the direct FFMA2 kernel
    (SSTG_TXP_WINO=0) that served these sizes before -- on a batch large enough for the full grid and on a small one."""
    import social_stgcnn_b200 as sstg
    rng = np.random.default_rng(21)
    for counts in (np.concatenate([np.full(150, 256), rng.integers(129, 257, 120), rng.integers(2, 129, 300)]),
                   np.array([256, 3, 200, 129, 64, 255, 130, 7])):
        n = rng.permutation(counts).astype(np.int32)
        plan = sstg.ScenePlan(n)
        dp = plan.on(dev)
        H = torch.from_numpy(rng.standard_normal(plan.P * 40).astype(np.float32)).to(dev)
        prm = torch.from_numpy(golden.params("univ")[0]).to(dev)
        a = sstg.txpcnn_forward(dp, H, prm).clone()
        os.environ["SSTG_TXP_FILTERS"] = "smem"
        try:
            b = sstg.txpcnn_forward(dp, H, prm).clone()
        finally:
            os.environ.pop("SSTG_TXP_FILTERS", None)
        os.environ["SSTG_TXP_WINO"] = "0"
        try:
            c = sstg.txpcnn_forward(dp, H, prm).clone()
        finally:
            os.environ.pop("SSTG_TXP_WINO", None)
        torch.cuda.synchronize()
        assert torch.isfinite(a).all()
        assert torch.equal(a, b)
        scale = float(c.abs().max())
        assert float((a - c).abs().max()) <= 1e-5 * scale, float((a - c).abs().max()) / scale


@pytest.mark.skipif(os.environ.get("SSTG_TEST_EXPERIMENTAL") != "1",
                    reason="SSTG_TXP_PREFETCH is an unmeasured A/B variant: run with SSTG_TEST_EXPERIMENTAL=1 (under a timeout)")
def test_winograd_next_tile_prefetch_variant_is_bit_identical(golden, dev):
    """SSTG_TXP_PREFETCH=1 fetches the next tile's 40 input rows into a staging tile while the current tile's
    layers 1..4 run.  Same arithmetic in the same order: bit-identical outputs, for dense tiles (2 x 64), ragged
    tiles (many small scenes per tile, odd and even pedestrian counts) and a grid with a ragged last wave."""
    import social_stgcnn_b200 as sstg
    rng = np.random.default_rng(13)
    for n in (np.full(4096, 64), np.concatenate([np.full(600, 64), rng.integers(2, 58, 4000)]),
              rng.integers(2, 9, 40000), np.concatenate([rng.integers(100, 129, 300), rng.integers(2, 58, 3000)])):
        plan = sstg.ScenePlan(n.astype(np.int32))
        assert plan.n_tiles >= 2 * 148                      # the constant-bank variant is selected
        dp = plan.on(dev)
        H = torch.from_numpy(rng.standard_normal(plan.P * 40).astype(np.float32)).to(dev)
        prm = torch.from_numpy(golden.params("eth")[0]).to(dev)
        a = sstg.txpcnn_forward(dp, H, prm).clone()
        os.environ["SSTG_TXP_PREFETCH"] = "1"
        try:
            b = sstg.txpcnn_forward(dp, H, prm).clone()
        finally:
            os.environ.pop("SSTG_TXP_PREFETCH", None)
        torch.cuda.synchronize()
        assert torch.equal(a, b), (int(n.min()), int(n.max()), len(n))
