"""GPU parity tests of the TRAINING path (run with -m gpu on a B200), through the C ABI:
train-mode forward (per-scene BatchNorm statistics), NLL gradient, full backward, BN running
statistics and SGD, against (a) the golden vectors the REAL reference's train() produced
(tests/golden/loss_golden.npz, forward_golden.npz) and (b) the oracle on seeded inputs.

Tolerances: loss within 1e-5 relative (north_star); gradients within 5e-5 of the largest
gradient entry of the tensor group (fp32 autograd vs fp32 kernels, different summation
order); BN buffers within 1e-5 relative."""
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


def _pack(golden, names):
    g = golden.graph
    n_peds = np.array([g[f"{n}.obs_rel"].shape[0] for n in names], np.int32)
    obs_rel = np.concatenate([g[f"{n}.obs_rel"] for n in names]).astype(np.float32)
    pred_rel = np.concatenate([g[f"{n}.pred_rel"] for n in names]).astype(np.float32)
    return n_peds, obs_rel, pred_rel


def _param_groups():
    from oracle import stgcnn_oracle as O
    off, out = 0, []
    for k, shp in O.PARAM_LAYOUT:
        n = int(np.prod(shp))
        out.append((k, off, off + n))
        off += n
    return out


def _grad_report(got, ref):
    """Worst error per parameter tensor, relative to the largest |ref| entry of the whole gradient."""
    scale = max(float(np.abs(ref).max()), 1e-30)
    rows = []
    for k, a, b in _param_groups():
        rows.append((float(np.abs(got[a:b] - ref[a:b]).max()) / scale, k))
    return sorted(rows, reverse=True)


@pytest.mark.parametrize("ckpt", ["eth", "zara1"])
def test_forward_train_vs_reference(golden, dev, ckpt):
    """Train-mode forward (BN batch statistics of each scene) and the BN buffer update, one
    scene at a time from the checkpoint state, against the reference's own train-mode forward."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E
    params, bn = golden.params(ckpt)
    keys = [k for k in golden.fwd.files if k.startswith(ckpt + ".") and k.endswith(".train")]
    assert len(keys) == 4
    names = [k.split(".")[1] for k in keys]
    n_peds, obs_rel, _ = _pack(golden, names)
    plan = sstg.ScenePlan(n_peds)
    dp = plan.on(dev)
    p = torch.from_numpy(params).to(dev)
    V, A = sstg.graph_build(dp, torch.from_numpy(obs_rel).to(dev), 8, fp64=True)
    vp, save, stats = E.forward_train(dp, V, A, p)
    torch.cuda.synchronize()
    vp = vp.cpu().numpy()
    for i, (key, name) in enumerate(zip(keys, names)):
        n, o = int(n_peds[i]), int(plan.seg_off[i])
        got = vp[o * 60:(o + n) * 60].reshape(1, 5, 12, n)
        assert rel_err(got, golden.fwd[key]) < 2e-5, (key, rel_err(got, golden.fwd[key]))
        # BN buffers after forwarding THIS scene alone from the checkpoint state
        one = sstg.ScenePlan([n])
        b = torch.from_numpy(bn.copy()).to(dev)
        E.bn_running_update(one.on(dev), stats[i:i + 1].contiguous(), b)
        ref_bn = golden.fwd[key + "_bn"]
        assert np.allclose(b.cpu().numpy(), ref_bn, rtol=1e-5, atol=1e-7), (key, np.abs(b.cpu().numpy() - ref_bn).max())


@pytest.mark.parametrize("tag", ["g0", "g1", "g2"])
def test_train_group_vs_reference(golden, dev, tag):
    """One optimiser group of the reference's train() (bs = 4, the last scene unscored; g2 is
    the short trailing group): loss, the 7,563-float gradient, BN buffers."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E
    Lz = golden.loss
    names = [str(n) for n in Lz[f"train.{tag}.names"]]
    n_peds, obs_rel, pred_rel = _pack(golden, names)
    params, bn = golden.params("eth")
    eng = sstg.TrainEngine(params, bn, dev)
    plan = sstg.ScenePlan(n_peds)
    r = eng.forward_backward(plan, torch.from_numpy(obs_rel).to(dev), torch.from_numpy(pred_rel).to(dev),
                             E.group_weights(len(names), 4))
    E.bn_running_update(plan.on(dev), r["scene_stats"], eng.bn)
    torch.cuda.synchronize()
    loss = float((r["loss"].double() * r["weight"].double()).sum())
    ref_loss = float(Lz[f"train.{tag}.loss"][0])
    assert abs(loss - ref_loss) < 1e-5 * max(1.0, abs(ref_loss)), (loss, ref_loss)
    grad, ref = r["grad"].cpu().numpy(), Lz[f"train.{tag}.grad"]
    rep = _grad_report(grad, ref)
    assert rep[0][0] < 5e-5, rep[:6]
    none_mask = Lz[f"train.{tag}.grad_is_none"]
    assert np.all(grad[none_mask] == 0.0)               # tpcnns.4 / prelus.4 are never executed
    assert np.allclose(eng.bn.cpu().numpy(), Lz[f"train.{tag}.bn"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("n_list", [
    [2, 3, 33, 64, 7, 130, 5, 40],
    # the switch points of the small-tile forms of the TXP-CNN kernels (<= 32: 4 lanes per column, <= 64: 2,
    # above: 1; > 128: two columns per thread) and several tiny scenes sharing one tile
    [32, 63, 65, 2, 2, 3, 2, 31, 128, 129],
    [2, 2, 2, 3, 2, 4, 2, 2, 3, 2, 2, 5],
], ids=["mixed", "mode-boundaries", "tiny-shared-tiles"])
def test_train_step_vs_fp64_oracle_ragged(golden, dev, n_list):
    """A ragged synthetic group (N = 2 .. 130, spanning several TXP-CNN tiles and the NJ = 1..8
    kernel variants) against the oracle evaluated in fp64 on the same inputs."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E
    from oracle import stgcnn_oracle as O
    obs_rel, pred_rel, _ = O.synthetic_scenes(1, sum(n_list), seed=77)
    params, bn = golden.params("zara1")
    eng = sstg.TrainEngine(params, bn, dev)
    plan = sstg.ScenePlan(n_list)
    bs = len(n_list)
    r = eng.forward_backward(plan, torch.from_numpy(obs_rel).to(dev), torch.from_numpy(pred_rel).to(dev),
                             E.group_weights(len(n_list), bs))
    E.bn_running_update(plan.on(dev), r["scene_stats"], eng.bn)
    torch.cuda.synchronize()
    P = O.unflatten_params(params, torch.float64)
    S = O.unflatten_bn(bn, torch.float64)
    scenes = []
    for i, n in enumerate(n_list):
        o = int(plan.seg_off[i])
        V, A = O.seq_to_graph(obs_rel[o:o + n])
        scenes.append((V, A, np.ascontiguousarray(pred_rel[o:o + n].transpose(2, 0, 1))))
    lo, go, So, per = O.train_step(P, S, scenes, bs)
    loss = float((r["loss"].double() * r["weight"].double()).sum())
    assert abs(loss - lo) < 1e-5 * max(1.0, abs(lo)), (loss, lo)
    got_l = r["loss"].cpu().numpy()[:-1]
    assert np.abs(got_l - np.array(per)).max() < 1e-5 * max(1.0, np.abs(per).max())
    rep = _grad_report(r["grad"].cpu().numpy().astype(np.float64), go.astype(np.float64))
    assert rep[0][0] < 5e-5, rep[:6]
    assert np.allclose(eng.bn.cpu().numpy(), O.flatten_bn(So), rtol=1e-5, atol=1e-7)


def test_nll_gradient_vs_autograd_incl_clamp(golden, dev):
    """sstg_nll_backward against autograd through the reference formula, including an element
    whose pdf underflows below the 1e-20 clamp (loss 46.05..., zero gradient; metrics.py:110)."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E
    from oracle import stgcnn_oracle as O
    rng = np.random.default_rng(5)
    n_list = [3, 9, 64]
    plan = sstg.ScenePlan(n_list)
    P_ = plan.P
    vp = rng.normal(0, 0.7, (P_ * 60,)).astype(np.float32)
    vt_blocks, ref_l, ref_g = [], [], []
    w = np.array([0.25, 0.5, 1.0], np.float32)
    for i, n in enumerate(n_list):
        o = int(plan.seg_off[i])
        blk = vp[o * 60:(o + n) * 60].reshape(5, 12, n)
        tgt = rng.normal(0, 0.5, (12, n, 2)).astype(np.float32)
        if i == 1:
            blk[2, 0, 0] = -6.0          # sigma_x = e^-6 and a 3-unit miss: exp(-z/2q) underflows
            tgt[0, 0, 0] = blk[0, 0, 0] + 3.0
        vt_blocks.append(tgt.reshape(-1))
        t = torch.tensor(blk.transpose(1, 2, 0).copy(), dtype=torch.float64, requires_grad=True)
        l = O.bivariate_loss(t, torch.tensor(tgt, dtype=torch.float64))
        (l * float(w[i])).backward()
        ref_l.append(float(l))
        ref_g.append(t.grad.numpy().transpose(2, 0, 1).reshape(-1))
    vt = np.concatenate(vt_blocks)
    loss, dv = E.nll_backward(plan.on(dev), torch.from_numpy(vp).to(dev), torch.from_numpy(vt).to(dev),
                              torch.from_numpy(w).to(dev))
    torch.cuda.synchronize()
    loss, dv = loss.cpu().numpy(), dv.cpu().numpy()
    ref_g = np.concatenate(ref_g)
    # the fp64 reference does not underflow where fp32 does: compare the clamped element separately
    o1 = int(plan.seg_off[1])
    clamp_idx = [o1 * 60 + c * 12 * 9 for c in range(5)]
    assert np.all(dv[clamp_idx] == 0.0)
    mask = np.ones_like(dv, bool)
    mask[clamp_idx] = False
    assert np.abs(dv[mask] - ref_g[mask]).max() < 1e-5 * np.abs(ref_g[mask]).max()
    assert abs(loss[0] - ref_l[0]) < 1e-5 * abs(ref_l[0]) and abs(loss[2] - ref_l[2]) < 1e-5 * abs(ref_l[2])
    # scene 1: the clamped element contributes -log(1e-20)
    t32 = torch.tensor(vp[o1 * 60:(o1 + 9) * 60].reshape(5, 12, 9).transpose(1, 2, 0).copy())
    l32 = float(O.bivariate_loss(t32, torch.tensor(vt_blocks[1].reshape(12, 9, 2))))
    assert abs(loss[1] - l32) < 1e-5 * abs(l32)


def test_backward_is_deterministic_and_sgd(golden, dev):
    """Two runs give bit-identical gradients (fixed-order reductions, no atomics); the SGD
    kernel applies p -= lr * g and clip_grad_norm_ semantics (train.py:69-74)."""
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import engine as E
    from oracle import stgcnn_oracle as O
    n_list = [4, 57, 2, 21, 64, 3, 3, 12]
    obs_rel, pred_rel, _ = O.synthetic_scenes(1, sum(n_list), seed=3)
    params, bn = golden.params("eth")
    eng = sstg.TrainEngine(params, bn, dev)
    plan = sstg.ScenePlan(n_list)
    o, p = torch.from_numpy(obs_rel).to(dev), torch.from_numpy(pred_rel).to(dev)
    g1 = eng.forward_backward(plan, o, p, E.group_weights(8, 8))["grad"].clone()
    g2 = eng.forward_backward(plan, o, p, E.group_weights(8, 8))["grad"].clone()
    torch.cuda.synchronize()
    assert torch.equal(g1, g2)
    g = g1.cpu().numpy().astype(np.float64)
    nrm = float(np.sqrt((g * g).sum()))
    # plain SGD
    w0 = eng.params.clone()
    out = torch.zeros(1, device=dev)
    E.sgd_step(eng.params, g1, lr=0.01, grad_norm_out=out)
    assert abs(float(out) - nrm) < 1e-5 * nrm
    assert np.allclose(eng.params.cpu().numpy(), w0.cpu().numpy() - 0.01 * g1.cpu().numpy(), rtol=0, atol=1e-7)
    # clipped: coefficient = max_norm / (norm + 1e-6), applied only if < 1
    eng.params.copy_(w0)
    clip = nrm / 4
    E.sgd_step(eng.params, g1, lr=0.01, clip_norm=clip)
    coef = clip / (nrm + 1e-6)
    assert np.allclose(eng.params.cpu().numpy(), w0.cpu().numpy() - 0.01 * coef * g1.cpu().numpy(), rtol=0, atol=1e-7)
    # a whole step moves the weights and counts the forwarded scenes
    eng.params.copy_(w0)
    loss = eng.step(plan, o, p, batch_size=8, lr=0.01)
    assert np.isfinite(float(loss)) and eng.num_batches_tracked == 8
    assert not torch.equal(eng.params, w0)


def test_fused_step_call_equals_separate_calls(golden, dev):
    """sstg_train_step (one library call per optimiser group) == the same group issued call by
    call through the functional ops: identical weights, BN buffers, loss, bit for bit."""
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    n_list = [5, 3, 64, 2, 17, 31, 4, 9, 2, 2, 57, 6]
    obs_rel, pred_rel, _ = O.synthetic_scenes(1, sum(n_list), seed=21)
    params, bn = golden.params("hotel") if "hotel.params" in golden.weights.files else golden.params("eth")
    o, p = torch.from_numpy(obs_rel).to(dev), torch.from_numpy(pred_rel).to(dev)
    plan = sstg.ScenePlan(n_list)
    a, b = sstg.TrainEngine(params, bn, dev), sstg.TrainEngine(params, bn, dev)
    for it in range(3):
        exact = bool(it & 1)                       # both graph modes
        la = a.step(plan, o, p, batch_size=12, lr=0.01, clip_grad=10.0, graph_fp64=exact)
        lb = b.step_unfused(plan, o, p, batch_size=12, lr=0.01, clip_grad=10.0, graph_fp64=exact)
        torch.cuda.synchronize()
        assert torch.equal(a.params, b.params) and torch.equal(a.bn, b.bn), it
        assert abs(float(la) - float(lb)) < 1e-6 * max(1.0, abs(float(lb)))
    for it in range(2):                              # no clipping: SGD is applied inside the tail kernel
        la = a.step(plan, o, p, batch_size=12, lr=0.01)
        lb = b.step_unfused(plan, o, p, batch_size=12, lr=0.01, graph_fp64=False)
        torch.cuda.synchronize()
        assert torch.equal(a.params, b.params) and torch.equal(a.bn, b.bn), it
        assert torch.equal(a.grad, b.grad)
    assert a.num_batches_tracked == b.num_batches_tracked == 60


def test_train_step_stays_inside_its_buffers(golden, dev):
    """compute-sanitizer is not available on this pool, so the one-call training step is run with every
    buffer it writes (workspace of exactly sstg_train_step_workspace floats, loss, grad) embedded in a
    sentinel-filled arena: all sentinels must survive, for tile shapes that hit every small-tile form."""
    import ctypes
    import social_stgcnn_b200 as sstg
    from social_stgcnn_b200 import _lib as L
    from social_stgcnn_b200 import engine as E
    from oracle import stgcnn_oracle as O
    lib = L.load()
    params, bn = golden.params("eth")
    SENT = 0x7fc0dead                                      # a NaN bit pattern no kernel produces
    for n_list in ([2, 2, 3, 2, 5, 2], [32, 33, 64, 65, 2, 31], [128, 129, 7], [57] * 5):
        plan = sstg.ScenePlan(n_list)
        dp = plan.on(dev)
        obs_rel, pred_rel, _ = O.synthetic_scenes(1, sum(n_list), seed=3)
        o, p = torch.from_numpy(obs_rel).to(dev), torch.from_numpy(pred_rel).to(dev)
        need = int(lib.sstg_train_step_workspace(plan.S, plan.n_tiles, plan.P, plan.a_floats))
        G = 4096                                           # guard floats on every side
        sizes = [need, plan.S + 1, L.N_PARAMS]             # ws | loss (+ group loss) | grad
        arena = torch.full((sum(sizes) + G * (len(sizes) + 1),), 0, dtype=torch.int32, device=dev)
        arena.fill_(SENT)
        offs, cur = [], G
        for sz in sizes:
            offs.append(cur)
            cur += sz + G
        fl = arena.view(torch.float32)
        ws, loss, grad = (fl[o_:o_ + sz] for o_, sz in zip(offs, sizes))
        assert ws.data_ptr() % 16 == 0
        prm = torch.from_numpy(params.astype(np.float32)).to(dev)
        bnb = torch.from_numpy(bn.astype(np.float32)).to(dev)
        w = torch.from_numpy(E.group_weights(plan.S, plan.S)).to(dev)
        rc = lib.sstg_train_step(o.data_ptr(), p.data_ptr(), dp.seg_off.data_ptr(), dp.a_off.data_ptr(),
                                 dp.tile_off.data_ptr(), plan.S, plan.n_tiles, plan.max_n, plan.P, plan.a_floats,
                                 prm.data_ptr(), bnb.data_ptr(), w.data_ptr(), ws.data_ptr(), need, 0.01, 0.0, 0.1, 0,
                                 loss.data_ptr(), grad.data_ptr(), loss.data_ptr() + 4 * plan.S,
                                 torch.cuda.current_stream(dev).cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        mask = torch.ones_like(arena, dtype=torch.bool)
        for o_, sz in zip(offs, sizes):
            mask[o_:o_ + sz] = False
        assert bool((arena[mask] == SENT).all()), n_list     # nothing outside the three buffers was touched
        assert bool(torch.isfinite(grad).all()) and bool(torch.isfinite(loss).all())
        assert prm.numel() == L.N_PARAMS and bool(torch.isfinite(prm).all())


def test_dependent_launch_chain_equals_plain_stream_order(golden, dev):
    """The training chain is launched with programmatic dependent launch (prologues overlap the predecessor).
    With SSTG_NO_PDL=1 the same kernels run in plain stream order: 60 optimiser steps over changing ragged
    groups must leave bit-identical weights, BN buffers and losses either way."""
    import os
    import social_stgcnn_b200 as sstg
    from oracle import stgcnn_oracle as O
    params, bn = golden.params("eth")
    rng = np.random.default_rng(8)
    groups = []
    for g in range(6):
        n = rng.integers(2, 58, 24).astype(np.int32)
        obs_rel, pred_rel, _ = O.synthetic_scenes(1, int(n.sum()), seed=40 + g)
        groups.append((sstg.ScenePlan(n), torch.from_numpy(obs_rel).to(dev), torch.from_numpy(pred_rel).to(dev)))

    def run(no_pdl):
        if no_pdl:
            os.environ["SSTG_NO_PDL"] = "1"
        try:
            eng = sstg.TrainEngine(params, bn, dev)
            losses = []
            for it in range(60):
                plan, o, p = groups[it % len(groups)]
                losses.append(eng.step(plan, o, p, 24, 0.01, clip_grad=(10.0 if it % 3 == 0 else None)).clone())
            torch.cuda.synchronize()
            return eng.params.clone(), eng.bn.clone(), torch.stack(losses)
        finally:
            os.environ.pop("SSTG_NO_PDL", None)

    a, b = run(False), run(True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert bool(torch.isfinite(a[2]).all())
