"""CPU tests: the oracle (oracle/stgcnn_oracle.py) against the committed golden vectors
that tests/golden/make_golden.py produced by running the REAL reference.  This is the
pin that lets the GPU parity tests trust the oracle."""
import numpy as np
import torch

from oracle import stgcnn_oracle as O
from tests.common import rel_err


def test_param_layout_matches_header():
    assert sum(int(np.prod(s)) for _, s in O.PARAM_LAYOUT) == O.N_PARAMS == 7563
    assert len(O.BN_LAYOUT) * 5 == 30


def test_graph_bit_exact_against_networkx(golden):
    """seq_to_graph restatement == reference (networkx) bit for bit: obs (T=8) and pred (T=12)
    halves of 77 scenes, N = 2..64, incl. the zero-distance (anorm == 0) scene."""
    n_el = 0
    for name in golden.names:
        for rel_k, v_k, a_k in (("obs_rel", "V_obs", "A_obs"), ("pred_rel", "V_tr", "A_tr")):
            V, A = O.seq_to_graph(golden.graph[f"{name}.{rel_k}"])
            assert np.array_equal(V, golden.graph[f"{name}.{v_k}"]), name
            ref = golden.graph[f"{name}.{a_k}"]
            assert np.array_equal(A.view(np.uint32), ref.view(np.uint32)), name
            n_el += ref.size
    assert n_el == 187600


def test_graph_known_values(golden):
    """eth/test scene 0, frame 1 -> +-0.82324237; frame 0 is all zeros."""
    A = golden.graph["eth000.A_obs"]
    assert np.all(A[0] == 0)
    assert np.allclose(A[1], [[0.82324237, -0.82324237], [-0.82324237, 0.82324237]], atol=1e-8)
    _, A2 = O.seq_to_graph(golden.graph["eth000.obs_rel"])
    assert np.array_equal(A2, A)


def test_graph_zero_distance_and_no_laplacian(golden):
    rel = golden.graph["syn9dup.obs_rel"]
    _, W = O.seq_to_graph(rel, norm_lap_matr=False)
    assert np.all(W[1:, 1, 3] == 0) and np.all(W[1:, 3, 1] == 0)      # duplicated pedestrian -> weight 0
    assert np.all(W[:, np.arange(9), np.arange(9)] == 1)
    _, A = O.seq_to_graph(rel)
    assert np.all(np.signbit(A[A == 0]) == False)                     # absent edges are +0.0


def test_forward_eval_and_train(golden):
    for ckpt in ("eth", "hotel", "univ", "zara1", "zara2"):    # the five shipped checkpoints
        params, bn = golden.params(ckpt)
        P, S = O.unflatten_params(params), O.unflatten_bn(bn)
        for name in golden.names:
            with torch.no_grad():
                out = O.model_forward(P, S, golden.graph[f"{name}.V_obs"], golden.graph[f"{name}.A_obs"]).numpy()
            assert rel_err(out, golden.fwd[f"{ckpt}.{name}.eval"]) < 2e-6, (ckpt, name)
        for key in [k for k in golden.fwd.files if k.startswith(ckpt) and k.endswith(".train")]:
            name = key.split(".")[1]
            upd = {}
            with torch.no_grad():
                out = O.model_forward(P, S, golden.graph[f"{name}.V_obs"], golden.graph[f"{name}.A_obs"],
                                      training=True, stats_out=upd).numpy()
            assert rel_err(out, golden.fwd[key]) < 2e-5, key
            S2 = dict(S)
            S2.update(upd)
            assert np.allclose(O.flatten_bn(S2), golden.fwd[key + "_bn"], rtol=1e-5, atol=1e-7)


def test_known_prediction_values(golden):
    """eth checkpoint, eval, scene 0."""
    vp = golden.fwd["eth.eth000.eval"][0].transpose(1, 2, 0)          # (12, N, 5)
    assert np.allclose(vp[0, 0], [-0.45499301, 0.04672549, -1.52478576, -2.34347057, -0.98509097], atol=2e-6)
    assert np.allclose(vp[11, 1], [-0.56813139, 0.07984990, -1.78517723, -2.03921986, 0.07730427], atol=2e-6)


def test_nll_and_mean_over_eth(golden):
    params, bn = golden.params("eth")
    P, S = O.unflatten_params(params), O.unflatten_bn(bn)
    got = []
    for name in golden.names:
        with torch.no_grad():
            out = O.model_forward(P, S, golden.graph[f"{name}.V_obs"], golden.graph[f"{name}.A_obs"])
            got.append(float(O.bivariate_loss(O.vpred_tn5(out), torch.from_numpy(golden.graph[f"{name}.V_tr"]))))
    ref = golden.loss["eth.nll_eval"]
    assert np.abs(np.array(got) - ref).max() < 2e-6 * max(1.0, np.abs(ref).max())
    assert abs(ref[0] - 2.2498154640) < 1e-6 and abs(ref[:70].mean() - 0.84885728) < 1e-6


def test_train_step_semantics(golden):
    """train.py:28-79: loss = (sum of the first len-1 scene losses)/bs, last scene forwarded but
    unscored, BN running stats advanced by every scene, tpcnns.4 / prelus.4 never get a gradient."""
    params, bn = golden.params("eth")
    P, S = O.unflatten_params(params), O.unflatten_bn(bn)
    for tag in ("g0", "g1", "g2"):
        names = [str(n) for n in golden.loss[f"train.{tag}.names"]]
        scenes = [(golden.graph[f"{n}.V_obs"], golden.graph[f"{n}.A_obs"], golden.graph[f"{n}.V_tr"]) for n in names]
        loss, grad, S2, _ = O.train_step(P, S, scenes, 4)
        assert abs(loss - golden.loss[f"train.{tag}.loss"][0]) < 1e-5 * max(1.0, abs(loss))
        g_ref = golden.loss[f"train.{tag}.grad"]
        assert np.abs(grad - g_ref).max() < 2e-5 * np.abs(g_ref).max()
        none = golden.loss[f"train.{tag}.grad_is_none"]
        assert none.sum() == 1309 and np.all(grad[none] == 0)
        assert np.allclose(O.flatten_bn(S2), golden.loss[f"train.{tag}.bn"], rtol=1e-5, atol=1e-7)
        assert int(golden.loss[f"train.{tag}.nbt"][0]) == int(golden.weights["eth.nbt"][0]) + len(names)


def test_sampling_ade_fde(golden):
    """test.py:59-123 with injected noise.  Against the reference: ~1e-6 (MKL's potrf L10 is not
    a correctly rounded quotient, and numpy-scalar **2 goes through powf) -- documented in the
    oracle; bar here 3e-6, far inside the 1e-5 budget."""
    params, bn = golden.params("eth")
    P, S = O.unflatten_params(params), O.unflatten_bn(bn)
    tot_a = tot_f = 0.0
    tot_n = 0
    for name in golden.names:
        with torch.no_grad():
            out = O.model_forward(P, S, golden.graph[f"{name}.V_obs"], golden.graph[f"{name}.A_obs"])
        vp = O.vpred_tn5(out).numpy()
        a, f, pred = O.sample_ade_fde(vp, golden.graph[f"{name}.V_tr"], golden.graph[f"{name}.obs"][:, :, -1],
                                      golden.sample[f"{name}.eps"])
        ra, rf = golden.sample[f"{name}.min_ade"], golden.sample[f"{name}.min_fde"]
        assert np.abs(a - ra).max() < 3e-6 * ra.max() and np.abs(f - rf).max() < 3e-6 * rf.max(), name
        if f"{name}.pred_abs" in golden.sample:
            assert rel_err(pred, golden.sample[f"{name}.pred_abs"]) < 1e-6
        if name.startswith("eth"):
            tot_a += a.sum(); tot_f += f.sum(); tot_n += a.size
    assert tot_n == 181
    assert abs(tot_a / tot_n - golden.sample["eth.ade"][0]) < 1e-6
    assert abs(tot_f / tot_n - golden.sample["eth.fde"][0]) < 1e-6


def test_rel_to_abs_rounding_points():
    rng = np.random.default_rng(3)
    rel = rng.standard_normal((12, 5, 2)).astype(np.float32)
    init = (rng.standard_normal((5, 2)) * 10).astype(np.float32)
    out = O.rel_to_abs(rel, init)
    ref = np.zeros_like(rel)
    for s in range(12):                      # metrics.py:69-75 semantics, spelled out
        for p in range(5):
            acc = np.float32(0)
            for u in range(s + 1):
                acc = np.float32(acc + rel[u, p])
            ref[s, p] = (acc.astype(np.float64) + init[p].astype(np.float64)).astype(np.float32)
    assert np.array_equal(out, ref)
