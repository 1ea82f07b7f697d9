"""GPU tests of the drop-in Python surface (social_stgcnn_b200.{model,metrics,utils}): the
reference's own call sequences (train.py:38-74, test.py:38-123, utils.py:86-228) run against
the CUDA-backed replacements and are compared with the golden vectors of the real reference."""
import math
import os

import numpy as np
import pytest
import torch

from tests.common import GOLD, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def _state_dict(golden, ckpt):
    from oracle import stgcnn_oracle as O
    params, bn = golden.params(ckpt)
    sd = dict(O.unflatten_params(params))
    sd.update(O.unflatten_bn(bn))
    nbt = int(golden.weights[f"{ckpt}.nbt"][0])
    for p in ("st_gcns.0.tcn.0", "st_gcns.0.tcn.3", "st_gcns.0.residual.1"):
        sd[p + ".num_batches_tracked"] = torch.tensor(nbt)
    return sd


def _model(golden, dev, ckpt="eth"):
    from social_stgcnn_b200.model import social_stgcnn
    m = social_stgcnn(n_stgcnn=1, n_txpcnn=5, output_feat=5, seq_len=8, kernel_size=3, pred_seq_len=12)
    m.load_state_dict(_state_dict(golden, ckpt), strict=True)       # exactly the checkpoint's 40 keys
    return m.to(dev)


def _scene(golden, name, dev):
    g = golden.graph
    t = lambda k: torch.from_numpy(g[f"{name}.{k}"]).to(dev)
    return t("V_obs")[None], t("A_obs")[None], t("V_tr")[None]        # the DataLoader's leading 1


def test_eval_forward_is_the_reference(golden, dev):
    m = _model(golden, dev).eval()
    for name in ("eth000", "univ000", "syn64", "syn9dup"):
        V_obs, A_obs, _ = _scene(golden, name, dev)
        a_in = A_obs.squeeze(0)
        with torch.no_grad():
            vp, a_out = m(V_obs.permute(0, 3, 1, 2), a_in)            # test.py:38-40
        assert a_out is a_in                                          # model.py:198 returns `a` itself
        assert rel_err(vp.cpu().numpy(), golden.fwd[f"eth.{name}.eval"]) < 1e-5, name
    with pytest.raises(AssertionError):                               # model.py:65
        m(V_obs.permute(0, 3, 1, 2), a_in[:7])


def test_train_loop_of_the_reference(golden, dev):
    """train.py:38-74 verbatim on the drop-ins (bs = 4, group g0): accumulated loss, .grad of every
    parameter, BN buffers and num_batches_tracked equal the reference's."""
    from social_stgcnn_b200.metrics import bivariate_loss
    Lz = golden.loss
    names = [str(n) for n in Lz["train.g0.names"]]
    m = _model(golden, dev).train()
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    batch_size, loader_len, turn_point = 4, len(names), len(names) - 1
    batch_count, is_fst_loss, loss_batch = 0, True, 0.0
    for cnt, name in enumerate(names):
        batch_count += 1
        V_obs, A_obs, V_tr = _scene(golden, name, dev)
        opt.zero_grad()
        V_pred, _ = m(V_obs.permute(0, 3, 1, 2), A_obs.squeeze(0))
        V_pred = V_pred.permute(0, 2, 3, 1).squeeze(0)
        if batch_count % batch_size != 0 and cnt != turn_point:
            l = bivariate_loss(V_pred, V_tr.squeeze(0))
            loss = l if is_fst_loss else loss + l
            is_fst_loss = False
        else:
            loss = loss / batch_size
            is_fst_loss = True
            loss.backward()
            opt.step()
            loss_batch += loss.item()
    assert loader_len == 4
    ref_loss = float(Lz["train.g0.loss"][0])
    assert abs(loss_batch - ref_loss) < 1e-5 * max(1.0, abs(ref_loss))
    ref_grad, none_mask = Lz["train.g0.grad"], Lz["train.g0.grad_is_none"]
    off = 0
    scale = np.abs(ref_grad).max()
    for k, p in m.named_parameters():
        n = p.numel()
        if none_mask[off]:
            assert p.grad is None, k                                   # never executed (model.py:191)
        else:
            err = np.abs(p.grad.cpu().numpy().reshape(-1) - ref_grad[off:off + n]).max() / scale
            assert err < 5e-5, (k, err)
        off += n
    from oracle import stgcnn_oracle as O
    assert np.allclose(O.flatten_bn(m.state_dict()), Lz["train.g0.bn"], rtol=1e-5, atol=1e-7)
    assert int(m.state_dict()["st_gcns.0.tcn.0.num_batches_tracked"]) == int(Lz["train.g0.nbt"][0])


def test_bivariate_loss_value_and_autograd(golden, dev):
    from oracle import stgcnn_oracle as O
    from social_stgcnn_b200.metrics import bivariate_loss
    rng = np.random.default_rng(11)
    vp = rng.normal(0, 0.6, (12, 7, 5)).astype(np.float32)
    vt = rng.normal(0, 0.6, (12, 7, 2)).astype(np.float32)
    a = torch.tensor(vp, device=dev, requires_grad=True)
    l = bivariate_loss(a, torch.tensor(vt, device=dev))
    (3.0 * l).backward()
    b = torch.tensor(vp, dtype=torch.float64, requires_grad=True)
    lr = O.bivariate_loss(b, torch.tensor(vt, dtype=torch.float64))
    (3.0 * lr).backward()
    assert l.dim() == 0 and abs(float(l) - float(lr)) < 1e-5 * abs(float(lr))
    assert np.abs(a.grad.cpu().numpy() - b.grad.numpy()).max() < 1e-5 * np.abs(b.grad.numpy()).max()


def test_ade_fde_and_rel_to_abs(golden, dev):
    from oracle import stgcnn_oracle as O
    from social_stgcnn_b200 import metrics as M
    rng = np.random.default_rng(2)
    preds = [rng.normal(0, 2, (12, n, 2)).astype(np.float32) for n in (3, 1, 8)]
    tgts = [rng.normal(0, 2, (12, n, 2)).astype(np.float32) for n in (3, 1, 8)]
    counts = [3, 1, 6]

    def ref(final_only):
        tot = 0.0
        for p, t, c in zip(preds, tgts, counts):
            s = 0.0
            for i in range(c):
                for k in (range(11, 12) if final_only else range(12)):
                    dx, dy = p[k, i, 0] - t[k, i, 0], p[k, i, 1] - t[k, i, 1]
                    s += math.sqrt(np.float32(dx * dx) + np.float32(dy * dy))
            tot += s / (c if final_only else c * 12)
        return tot / len(preds)

    assert abs(M.ade(preds, tgts, counts) - ref(False)) < 1e-12 * ref(False) + 1e-15
    assert abs(M.fde(preds, tgts, counts) - ref(True)) < 1e-12 * ref(True) + 1e-15
    rel = rng.normal(0, 0.3, (12, 5, 2)).astype(np.float32)
    init = rng.normal(0, 8, (5, 2))
    assert np.array_equal(M.nodes_rel_to_nodes_abs(rel, init), O.rel_to_abs(rel, init))
    seq = rng.normal(0, 1, (1, 4, 2, 8)).astype(np.float32)
    V = M.seq_to_nodes(seq)
    assert V.shape == (8, 4, 2) and V.dtype == np.float64 and np.array_equal(V[3, 2], seq[0, 2, :, 3])


def test_seq_to_graph_and_dataset(golden, dev):
    """utils.seq_to_graph on CPU tensors (the reference's calling convention) and the whole
    TrajectoryDataset on the shipped eth/test file: 70 scenes, 181 pedestrians, every A_obs /
    A_tr bit-identical to the reference's networkx graphs."""
    from social_stgcnn_b200.utils import TrajectoryDataset, seq_to_graph
    g = golden.graph
    V, A = seq_to_graph(torch.from_numpy(g["univ000.obs"]), torch.from_numpy(g["univ000.obs_rel"]), True)
    assert not V.is_cuda and np.array_equal(A.numpy().view(np.uint32), g["univ000.A_obs"].view(np.uint32))
    assert np.array_equal(V.numpy(), g["univ000.V_obs"])
    d = TrajectoryDataset(os.path.join(GOLD, "datasets", "eth", "test") + "/", obs_len=8, pred_len=12, skip=1,
                          norm_lap_matr=True)
    assert len(d) == 70 and d.obs_traj.shape[0] == 181
    # the pred-half adjacency is built on first use only; V_tr is there from the start
    assert "A_tr" not in d.packed and "V_tr" in d.packed and len(d.A_pred) == 70
    for i in range(70):
        item = d[i]
        assert len(item) == 10
        assert "A_tr" in d.packed
        name = f"eth{i:03d}"
        assert np.array_equal(item[2].numpy(), g[f"{name}.obs_rel"]), name
        assert np.array_equal(item[6].numpy(), g[f"{name}.V_obs"]), name
        assert np.array_equal(item[7].numpy().view(np.uint32), g[f"{name}.A_obs"].view(np.uint32)), name
        assert np.array_equal(item[8].numpy(), g[f"{name}.V_tr"]), name
        assert np.array_equal(item[9].numpy().view(np.uint32), g[f"{name}.A_tr"].view(np.uint32)), name
    with pytest.raises(ValueError):
        TrajectoryDataset(os.path.join(GOLD, "datasets", "eth", "test") + "/", obs_len=8, pred_len=12, delim=" ")


def test_driver_evaluate_is_test_py(golden, dev):
    """driver.evaluate == test() of the reference on eth/test with the same injected noise:
    dataset ADE/FDE within 1e-5 relative; raw_data_dict trajectories within 1e-5."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import driver
    from social_stgcnn_b200.utils import TrajectoryDataset
    d = TrajectoryDataset(os.path.join(GOLD, "datasets", "eth", "test") + "/", obs_len=8, pred_len=12, skip=1,
                          norm_lap_matr=True, device=dev)
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    eps = np.concatenate([golden.sample[f"eth{i:03d}.eps"].reshape(-1) for i in range(70)]).astype(np.float32)
    ade_, fde_, raw = driver.evaluate(d, eng, KSTEPS=20, eps=torch.from_numpy(eps).to(dev), want_raw=True)
    ra, rf = float(golden.sample["eth.ade"][0]), float(golden.sample["eth.fde"][0])
    assert abs(ade_ - ra) < 1e-5 * ra and abs(fde_ - rf) < 1e-5 * rf, (ade_, ra, fde_, rf)
    assert sorted(raw) == list(range(1, 71))
    for i in (0, 17, 69):
        name = f"eth{i:03d}"
        got = np.stack(raw[i + 1]["pred"])
        assert got.shape == golden.sample[f"{name}.pred_abs"].shape
        assert rel_err(got, golden.sample[f"{name}.pred_abs"]) < 1e-5, name
        assert np.array_equal(raw[i + 1]["obs"][-1], golden.graph[f"{name}.obs"][:, :, -1]) or \
            rel_err(raw[i + 1]["obs"][-1], golden.graph[f"{name}.obs"][:, :, -1]) < 1e-6
    # Philox mode: statistically the same score (eth/test unseeded reference run: 0.73 / 1.22)
    a2, f2, _ = driver.evaluate(d, eng, KSTEPS=20, seed=3)
    assert 0.6 < a2 < 0.9 and 1.0 < f2 < 1.5


def test_driver_train_epoch_vs_oracle(golden, dev):
    """driver.train_epoch (two optimiser groups of 4 on the first 8 eth/test scenes, lr 0.01)
    against the oracle replaying train.py's loop with SGD: returned loss, final weights, BN."""
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    from social_stgcnn_b200 import driver
    from social_stgcnn_b200.utils import TrajectoryDataset
    d = TrajectoryDataset(os.path.join(GOLD, "datasets", "eth", "test") + "/", obs_len=8, pred_len=12, skip=1,
                          norm_lap_matr=True, device=dev)
    params, bn = golden.params("eth")
    eng = sstg.TrainEngine(params, bn, dev)
    order = np.arange(8)

    class _Sub:                                       # the first 8 scenes as an "epoch"
        packed, plan_obs = d.packed, d.plan_obs
        def __len__(self):
            return 8
    got = driver.train_epoch(_Sub(), eng, batch_size=4, lr=0.01, order=order)
    torch.cuda.synchronize()
    P, S = O.unflatten_params(params), O.unflatten_bn(bn)
    flat = params.astype(np.float32).copy()
    tot = 0.0
    for g0 in (0, 4):
        scenes = []
        for i in range(g0, g0 + 4):
            g = golden.graph
            scenes.append((g[f"eth{i:03d}.V_obs"], g[f"eth{i:03d}.A_obs"], g[f"eth{i:03d}.V_tr"]))
        lo, go, S, _ = O.train_step(O.unflatten_params(flat), S, scenes, 4)
        flat = (flat - np.float32(0.01) * go).astype(np.float32)
        tot += lo
    assert abs(got - tot / 8) < 1e-5 * max(1.0, abs(tot / 8)), (got, tot / 8)
    w = eng.params.cpu().numpy()
    assert np.abs(w - flat).max() < 1e-5 * np.abs(flat).max()
    assert np.allclose(eng.bn.cpu().numpy(), O.flatten_bn(S), rtol=1e-5, atol=1e-7)
    assert eng.num_batches_tracked == 8


def test_driver_fit_writes_reference_artifacts(golden, dev, tmp_path):
    """driver.fit (main() of train.py): two epochs on the eth/test split; the artefacts are the
    reference's (args.pkl, val_best.pth with the 40-key state_dict, metrics.pkl,
    constant_metrics.pkl) and the checkpoint loads into the drop-in module and reproduces the
    engine's validation loss."""
    import pickle
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import driver
    from social_stgcnn_b200.model import social_stgcnn
    from social_stgcnn_b200.utils import TrajectoryDataset
    d = TrajectoryDataset(os.path.join(GOLD, "datasets", "eth", "test") + "/", obs_len=8, pred_len=12, skip=1,
                          norm_lap_matr=True, device=dev)
    params, bn = golden.params("eth")
    eng = sstg.TrainEngine(params, bn, dev, num_batches_tracked=int(golden.weights["eth.nbt"][0]))
    ck = str(tmp_path / "ck") + "/"
    metrics, const = driver.fit(d, d, eng, num_epochs=2, batch_size=16, lr=0.01, use_lrschd=True, lr_sh_rate=1,
                                checkpoint_dir=ck, args={"tag": "t"}, seed=0, log=None)
    assert len(metrics["train_loss"]) == 2 and all(np.isfinite(metrics["val_loss"]))
    assert sorted(os.listdir(ck)) == ["args.pkl", "constant_metrics.pkl", "metrics.pkl", "val_best.pth"]
    assert pickle.load(open(ck + "metrics.pkl", "rb")) == metrics
    sd = torch.load(ck + "val_best.pth")
    m = social_stgcnn(n_stgcnn=1, n_txpcnn=5, output_feat=5, seq_len=8, kernel_size=3, pred_seq_len=12)
    m.load_state_dict(sd, strict=True)
    best = const["min_val_epoch"]
    ie = sstg.InferenceEngine.from_state_dict(sd, dev)
    assert abs(driver.validate_epoch(d, ie, 16) - metrics["val_loss"][best]) < 1e-6
    assert int(sd["st_gcns.0.tcn.0.num_batches_tracked"]) == int(golden.weights["eth.nbt"][0]) + 70 * (best + 1)


def test_compact_host_format_is_bit_exact(golden, dev):
    """sstg_decode_tracks rebuilds obs_traj_rel / pred_traj_rel / obs_traj[:, :, -1] of the dataset
    bit for bit from the 88-byte fixed-point tracks, and the host pipeline fed with them returns
    exactly what the fp32 interface returns."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200.utils import TrajectoryDataset
    d = TrajectoryDataset(os.path.join(GOLD, "datasets", "eth", "test") + "/", obs_len=8, pred_len=12, skip=1,
                          norm_lap_matr=True, device=dev)
    assert d.compact is not None
    start, delta = (torch.from_numpy(a) for a in d.compact)
    o, p, l = sstg.decode_tracks(start.to(dev), delta.to(dev))
    torch.cuda.synchronize()
    assert torch.equal(o.cpu(), d.obs_traj_rel) and torch.equal(p.cpu(), d.pred_traj_rel)
    assert torch.equal(l.cpu(), d.obs_traj[:, :, -1])
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    a1, f1 = eng.infer_host(d.plan_obs, d.obs_traj_rel.pin_memory(), d.pred_traj_rel.pin_memory(),
                            d.obs_traj[:, :, -1].contiguous().pin_memory(), 20, chunk_scenes=16, seed=1)
    a1, f1 = a1.copy(), f1.copy()
    a2, f2 = eng.infer_host(d.plan_obs, None, None, None, 20, chunk_scenes=16, seed=1,
                            compact=(start.pin_memory(), delta.pin_memory()))
    assert np.array_equal(a1, a2) and np.array_equal(f1, f2)


def test_validate_epoch_and_sharded_evaluator(golden, dev):
    """driver.validate_epoch == vald() of train.py:81-122 replayed on the reference's own per-scene eval
    NLLs (golden), and parallel.ShardedEvaluator at world size 1 == driver.evaluate."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import driver, parallel
    from social_stgcnn_b200.utils import TrajectoryDataset
    d = TrajectoryDataset(os.path.join(GOLD, "datasets", "eth", "test") + "/", obs_len=8, pred_len=12, skip=1,
                          norm_lap_matr=True, device=dev)
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    nll = golden.loss["eth.nll_eval"][:70]
    for bs in (16, 128, 7):
        ref = sum(float(nll[g0:g0 + bs][:-1].sum()) / bs for g0 in range(0, 70, bs)) / 70
        got = driver.validate_epoch(d, eng, bs)
        assert abs(got - ref) < 1e-5 * max(1.0, abs(ref)), (bs, got, ref)
    a1, f1, _ = driver.evaluate(d, eng, KSTEPS=20, seed=5, exact=False)
    ev = parallel.ShardedEvaluator(eng)
    a2, f2, n = ev.evaluate(d.plan_obs.n_peds, d.obs_traj_rel, d.pred_traj_rel, d.obs_traj[:, :, -1].contiguous(),
                            K=20, seed=5)
    assert n == 181 and abs(a1 - a2) < 1e-12 and abs(f1 - f2) < 1e-12


@pytest.mark.parametrize("split", ["test", "val"])
def test_nuscenes_dummy_dataset_graphs_are_the_reference(dev, split):
    """BASELINE config 5 (nuScenes_mini graph construction): TrajectoryDataset(delim=' ') on the reference's own
    nuScenes stand-in data == the unmodified reference (real networkx normalized Laplacian), scene by scene:
    V_obs, V_tr bit-exact and A_obs (T = 8), A_tr (T = 12, built lazily) BIT-IDENTICAL
    (tests/golden/nuscenes_golden.npz, made by tests/golden/make_nuscenes_golden.py)."""
    import os
    from tests.common import GOLD
    from social_stgcnn_b200.utils import TrajectoryDataset
    g = np.load(os.path.join(GOLD, "nuscenes_golden.npz"))
    d = TrajectoryDataset(os.path.join(GOLD, "datasets", "nuscenes_mini_dummy", split) + "/", obs_len=8, pred_len=12,
                          skip=1, norm_lap_matr=True, delim=" ", device=dev)
    n = g[f"{split}.n_peds"]
    assert len(d) == len(n) and [e - s for s, e in d.seq_start_end] == n.tolist()
    for name, lst in (("V_obs", d.v_obs), ("A_obs", d.A_obs), ("V_tr", d.v_pred), ("A_tr", d.A_pred)):
        got = np.concatenate([np.asarray(t.numpy(), np.float32).ravel() for t in lst])
        ref = g[f"{split}.{name}"]
        assert got.shape == ref.shape, name
        assert np.array_equal(got.view(np.int32), ref.view(np.int32)), (name, float(np.abs(got - ref).max()))
    item = d[len(d) - 1]
    assert len(item) == 10 and tuple(item[7].shape) == (8, int(n[-1]), int(n[-1])) and tuple(item[9].shape) == (12, int(n[-1]), int(n[-1]))


def test_graph_captured_evaluation_equals_the_plain_pipeline(golden, dev):
    """InferenceEngine.capture / driver.evaluate(graph=True): the split's whole pipeline replayed from ONE CUDA graph
    gives bit-identical per-pedestrian minima and the same dataset ADE/FDE as the kernel-by-kernel calls, sees new
    weights (load_weights) and new inputs (replay(...)) without a re-capture."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import driver
    from social_stgcnn_b200.utils import TrajectoryDataset
    d = TrajectoryDataset(os.path.join(GOLD, "datasets", "eth", "test") + "/", obs_len=8, pred_len=12, skip=1,
                          norm_lap_matr=True, device=dev)
    params, bn = golden.params("eth")
    eng = sstg.InferenceEngine(params, bn, dev)
    a0, f0, _ = driver.evaluate(d, eng, KSTEPS=20, seed=4)
    ref_ade = eng.infer(d.plan_obs, d.packed["obs_rel"], d.packed["pred_rel"], d.packed["last_obs"], 20, seed=4,
                        graph_fp64=True, fp64_dist=True)["min_ade"].clone()
    for _ in range(3):                                   # first call captures, the others replay
        a1, f1, _ = driver.evaluate(d, eng, KSTEPS=20, seed=4, graph=True)
        assert (a1, f1) == (a0, f0)
    g = next(iter(d._graphs.values()))
    torch.cuda.synchronize()
    assert torch.equal(g.out["min_ade"], ref_ade)
    # new weights are seen by the captured graph (it reads the engine's parameter tensors at replay time)
    p2, b2 = golden.params("zara1")
    eng.load_weights(p2, b2)
    a2, f2, _ = driver.evaluate(d, eng, KSTEPS=20, seed=4, graph=True)
    a3, f3, _ = driver.evaluate(d, eng, KSTEPS=20, seed=4)
    assert (a2, f2) == (a3, f3) and (a2, f2) != (a0, f0)
    # new inputs through replay(): the same scenes with the displacement order of the pedestrians' coordinates swapped
    sw = d.packed["obs_rel"].flip(1).contiguous()
    r = g.replay(obs_rel=sw)
    torch.cuda.synchronize()
    want = eng.infer(d.plan_obs, sw, d.packed["pred_rel"], d.packed["last_obs"], 20, seed=4, graph_fp64=True,
                     fp64_dist=True)["min_ade"]
    assert torch.equal(r["min_ade"], want)
