"""GPU parity at BASELINE.json's full sizes (10^5 scenes, 10^6..10^7 pedestrians), where the
oracle cannot run everything: size-independent properties plus oracle spot checks.

  * replication: a ragged batch made of R copies of the 77 golden scenes must give R bit-identical
    copies of every tensor (V, A, H, V_pred, NLL, injected-noise ADE/FDE) -- catches any offset /
    overflow / tiling error that only shows far into a large packed batch -- and copy 0 is
    checked against the reference's golden vectors;
  * dense crowds (32,768 scenes x 64): Laplacian identities (symmetry, D^1/2 in the null space),
    finiteness, min-over-K monotonicity (the k-th Philox sample does not depend on K), chunked
    host pipeline == one-shot pipeline, and oracle spot checks on random scenes."""
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
                              torch.from_numpy(last).pin_memory(), 20, chunk_scenes=S, seed=9)
    assert np.array_equal(a_h, ade20.cpu().numpy()) and np.array_equal(f_h, fde20.cpu().numpy())
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
