#!/usr/bin/env python
"""Generate the committed golden fixtures by running the REAL reference.

Run in the build container only (needs /root/reference, read-only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference modules (utils.py, model.py, metrics.py, train.py and the `test`
function of test.py) are imported / executed unmodified from /root/reference;
nothing of theirs is copied into this repo.  Outputs (small .npz files plus one
dataset text file used as a data fixture) land in tests/golden/.  While
generating, every value is also compared with oracle/stgcnn_oracle.py, and the
script aborts if the oracle does not reproduce the reference (bit-exact for
graphs and the conv arithmetic; ~1e-7 for ADE/FDE given the same trajectories,
~1e-6 end to end through MKL's Cholesky -- see the oracle's docstrings).
"""
import ast
import os
import shutil
import sys
import time

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import utils as ref_utils            # noqa: E402  (reference)
import model as ref_model            # noqa: E402
import metrics as ref_metrics        # noqa: E402
import train as ref_train            # noqa: E402
import torch.distributions.multivariate_normal as torchdist  # noqa: E402
from torch.utils.data import DataLoader, Dataset             # noqa: E402

from oracle import stgcnn_oracle as O  # noqa: E402

torch.set_num_threads(8)
CKPTS = ["eth", "hotel", "univ", "zara1", "zara2"]


def load_ckpt(name):
    sd = torch.load(f"{REF}/checkpoint/social-stgcnn-{name}/val_best.pth", map_location="cpu")
    m = ref_model.social_stgcnn(n_stgcnn=1, n_txpcnn=5, output_feat=5, seq_len=8, kernel_size=3, pred_seq_len=12)
    m.load_state_dict(sd)
    return m, sd


def dataset_without_graphs(path):
    """Reference TrajectoryDataset with the (minutes-long) graph build stubbed;
    graphs are then built with the real seq_to_graph only for chosen scenes."""
    real = ref_utils.seq_to_graph
    ref_utils.seq_to_graph = lambda a, b, c=True: (torch.zeros(1), torch.zeros(1))
    try:
        d = ref_utils.TrajectoryDataset(path, obs_len=8, pred_len=12, skip=1, norm_lap_matr=True)
    finally:
        ref_utils.seq_to_graph = real
    return d


def extract_test_fn():
    """test.py executes at import (no __main__ guard): pull out `test` via ast
    and exec the unmodified function source with the globals it expects."""
    src = open(f"{REF}/test.py").read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "test"][0]
    code = compile(ast.Module(body=[fn], type_ignores=[]), f"{REF}/test.py", "exec")
    ns = {"torch": torch, "np": np, "torchdist": torchdist, "copy": __import__("copy"),
          "device": torch.device("cpu"), "math": __import__("math")}
    for k in ("seq_to_nodes", "nodes_rel_to_nodes_abs", "ade", "fde"):
        ns[k] = getattr(ref_metrics, k)
    exec(code, ns)
    return ns


class ListDataset(Dataset):
    def __init__(self, items):
        self.items = items

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def scene_tuple(obs, pred, obs_rel, pred_rel):
    """The 10-tuple of TrajectoryDataset.__getitem__ (utils.py:217-228) built with
    the real seq_to_graph."""
    v_o, a_o = ref_utils.seq_to_graph(obs, obs_rel, True)
    v_p, a_p = ref_utils.seq_to_graph(pred, pred_rel, True)
    n = obs.shape[0]
    return [obs, pred, obs_rel, pred_rel, torch.zeros(n), torch.ones(n, 20), v_o, a_o, v_p, a_p]


def main():
    t00 = time.time()
    out = {}

    # ---------------- weights of the five shipped checkpoints ----------------
    W = {}
    for c in CKPTS:
        _, sd = load_ckpt(c)
        W[f"{c}.params"] = O.flatten_params(sd)
        W[f"{c}.bn"] = O.flatten_bn(sd)
        W[f"{c}.nbt"] = np.array([int(sd["st_gcns.0.tcn.0.num_batches_tracked"])], np.int64)
    np.savez_compressed(f"{HERE}/weights.npz", **W)

    # ---------------- scenes: all of eth/test + picked univ/test + synthetic ---
    os.makedirs(f"{HERE}/datasets/eth/test", exist_ok=True)
    shutil.copyfile(f"{REF}/datasets/eth/test/biwi_eth.txt", f"{HERE}/datasets/eth/test/biwi_eth.txt")
    d_eth = dataset_without_graphs(f"{REF}/datasets/eth/test/")
    d_univ = dataset_without_graphs(f"{REF}/datasets/univ/test/")
    cases = []   # (name, obs, pred, obs_rel, pred_rel)
    for i, (s, e) in enumerate(d_eth.seq_start_end):
        cases.append((f"eth{i:03d}", d_eth.obs_traj[s:e], d_eth.pred_traj[s:e],
                      d_eth.obs_traj_rel[s:e], d_eth.pred_traj_rel[s:e]))
    sizes = [e - s for s, e in d_univ.seq_start_end]
    picked = []
    for want in (57, 31, 21, 6, 3):
        idx = [i for i, n in enumerate(sizes) if n == want][:1]
        picked += idx
    for i in picked:
        s, e = d_univ.seq_start_end[i]
        cases.append((f"univ{i:03d}", d_univ.obs_traj[s:e], d_univ.pred_traj[s:e],
                      d_univ.obs_traj_rel[s:e], d_univ.pred_traj_rel[s:e]))
    obs_rel, pred_rel, obs_abs = O.synthetic_scenes(1, 64, seed=1234)
    pos_pred = obs_abs[:, :, -1:] + np.cumsum(pred_rel, axis=2)
    cases.append(("syn64", torch.from_numpy(obs_abs), torch.from_numpy(pos_pred.astype(np.float32)),
                  torch.from_numpy(obs_rel), torch.from_numpy(pred_rel)))
    obs_rel, pred_rel, obs_abs = O.synthetic_scenes(1, 9, seed=99)
    # duplicate one pedestrian's motion -> zero distance -> anorm == 0 branch (utils.py:25-26)
    obs_rel[3] = obs_rel[1]
    pred_rel[3] = pred_rel[1]
    pos_pred = obs_abs[:, :, -1:] + np.cumsum(pred_rel, axis=2)
    cases.append(("syn9dup", torch.from_numpy(obs_abs), torch.from_numpy(pos_pred.astype(np.float32)),
                  torch.from_numpy(obs_rel), torch.from_numpy(pred_rel)))

    names = [c[0] for c in cases]
    print("cases:", len(cases), "N:", sorted({c[1].shape[0] for c in cases}))

    # ---------------- (1) graph golden ----------------
    G = {"names": np.array(names)}
    tuples = {}
    t0 = time.time()
    n_el = 0
    for name, obs, pred, orel, prel in cases:
        tup = scene_tuple(obs, pred, orel, prel)
        tuples[name] = tup
        v_o, a_o, v_p, a_p = [t.numpy() for t in tup[6:10]]
        G[f"{name}.obs"] = obs.numpy()
        G[f"{name}.obs_rel"] = orel.numpy()
        G[f"{name}.pred_rel"] = prel.numpy()
        G[f"{name}.V_obs"] = v_o
        G[f"{name}.A_obs"] = a_o
        G[f"{name}.V_tr"] = v_p
        G[f"{name}.A_tr"] = a_p
        # oracle must be bit-identical
        for rel, v_ref, a_ref in ((orel, v_o, a_o), (prel, v_p, a_p)):
            v_or, a_or = O.seq_to_graph(rel.numpy())
            assert np.array_equal(v_or, v_ref), name
            assert np.array_equal(a_or.view(np.uint32), a_ref.view(np.uint32)), (name, np.abs(a_or - a_ref).max())
            n_el += a_ref.size
    print(f"graph: reference took {time.time()-t0:.1f}s; oracle bit-identical on {n_el} A elements")
    np.savez_compressed(f"{HERE}/graph_golden.npz", **G)

    # ---------------- (2)+(3) forward golden (eval + train mode) ----------------
    Fw = {"names": np.array(names)}
    worst = 0.0
    for c in CKPTS:                                        # all five shipped checkpoints
        m, sd = load_ckpt(c)
        P = {k: sd[k].float() for k, _ in O.PARAM_LAYOUT}
        S = {k: sd[k].float() for k in O.BN_LAYOUT}
        for name in names:
            tup = tuples[name]
            V_obs, A_obs = tup[6][None], tup[7][None]
            m.eval()
            with torch.no_grad():
                vp, a_ret = m(V_obs.permute(0, 3, 1, 2), A_obs.squeeze(0))
            assert a_ret.data_ptr() == A_obs.data_ptr()
            Fw[f"{c}.{name}.eval"] = vp.numpy()
            with torch.no_grad():
                o = O.model_forward(P, S, tup[6].numpy(), tup[7].numpy()).numpy()
            err = np.abs(o - vp.numpy()).max() / max(1.0, np.abs(vp.numpy()).max())
            worst = max(worst, err)
            assert err < 2e-6, (c, name, err)
        if c not in ("eth", "zara1"):
            continue
        # train mode: forward + BN buffer update, one scene at a time from the checkpoint state
        for name in ("eth000", "eth010", f"univ{picked[0]:03d}", "syn64"):
            m.load_state_dict(sd)
            m.train()
            tup = tuples[name]
            with torch.no_grad():
                vp, _ = m(tup[6][None].permute(0, 3, 1, 2), tup[7])
            Fw[f"{c}.{name}.train"] = vp.numpy()
            sd2 = m.state_dict()
            Fw[f"{c}.{name}.train_bn"] = O.flatten_bn(sd2)
            upd = {}
            with torch.no_grad():
                o = O.model_forward(P, S, tup[6].numpy(), tup[7].numpy(), training=True, stats_out=upd).numpy()
            assert np.abs(o - vp.numpy()).max() < 2e-5 * max(1.0, np.abs(o).max()), (c, name)
            S2 = dict(S); S2.update(upd)
            assert np.allclose(O.flatten_bn(S2), O.flatten_bn(sd2), rtol=1e-6, atol=1e-7)
    print(f"forward: oracle vs reference worst rel err {worst:.2e}")
    np.savez_compressed(f"{HERE}/forward_golden.npz", **Fw)

    # ---------------- (4a) loss golden ----------------
    m, sd = load_ckpt("eth")
    P = {k: sd[k].float() for k, _ in O.PARAM_LAYOUT}
    S = {k: sd[k].float() for k in O.BN_LAYOUT}
    m.eval()
    Lz = {}
    nll = []
    for name in names:
        tup = tuples[name]
        with torch.no_grad():
            vp, _ = m(tup[6][None].permute(0, 3, 1, 2), tup[7])
            vp = vp.permute(0, 2, 3, 1).squeeze(0)
            l = ref_metrics.bivariate_loss(vp, tup[8])
            lo = O.bivariate_loss(vp, tup[8])
        assert float(l) == float(lo)
        nll.append(float(l))
    Lz["names"] = np.array(names)
    Lz["eth.nll_eval"] = np.array(nll, np.float64)
    mean70 = float(np.mean(np.array(nll[:70], np.float32).astype(np.float64)))
    print("mean eval NLL over the 70 eth/test scenes:", mean70, "(SURVEY: 0.84885728)")
    assert abs(mean70 - 0.84885728) < 1e-6

    # train(): the reference's accumulate-and-backward, bs=4, lr=0 so weights stay put.
    class Args:
        batch_size = 4
        clip_grad = None
    for tag, group in (("g0", ["eth000", "eth001", "eth002", "eth003"]),
                       ("g1", ["eth020", f"univ{picked[1]:03d}", "syn9dup", "eth040"]),
                       ("g2", ["syn64", f"univ{picked[0]:03d}", "eth005"])):      # short last group (turn_point)
        m.load_state_dict(sd)
        opt = torch.optim.SGD(m.parameters(), lr=0.0)
        loader = DataLoader(ListDataset([tuples[n] for n in group]), batch_size=1, shuffle=False, num_workers=0)
        ret = ref_train.train(0, m, loader, opt, Args, torch.device("cpu"))
        grad = np.concatenate([(p.grad.numpy().reshape(-1) if p.grad is not None else np.zeros(p.numel(), np.float32))
                               for _, p in m.named_parameters()]).astype(np.float32)
        none_mask = np.concatenate([np.full(p.numel(), p.grad is None) for _, p in m.named_parameters()])
        Lz[f"train.{tag}.names"] = np.array(group)
        Lz[f"train.{tag}.loss"] = np.array([ret * len(group)], np.float64)   # train() returns loss/batch_count
        Lz[f"train.{tag}.grad"] = grad
        Lz[f"train.{tag}.grad_is_none"] = none_mask
        Lz[f"train.{tag}.bn"] = O.flatten_bn(m.state_dict())
        Lz[f"train.{tag}.nbt"] = np.array([int(m.state_dict()["st_gcns.0.tcn.0.num_batches_tracked"])])
        scenes = [(tuples[n][6].numpy(), tuples[n][7].numpy(), tuples[n][8].numpy()) for n in group]
        lo, go, So, _ = O.train_step(P, S, scenes, 4)
        assert abs(lo - ret * len(group)) < 1e-6 * max(1, abs(lo)), (lo, ret * len(group))
        gscale = np.abs(grad).max()
        assert np.abs(go - grad).max() < 2e-5 * gscale, (tag, np.abs(go - grad).max(), gscale)
        assert np.allclose(O.flatten_bn(So), O.flatten_bn(m.state_dict()), rtol=1e-5, atol=1e-7)
        print(f"train {tag}: loss {lo:.7f} |grad|max {gscale:.4f} oracle-vs-ref grad err {np.abs(go-grad).max():.2e}")
    np.savez_compressed(f"{HERE}/loss_golden.npz", **Lz)

    # ---------------- (4b) sampling + ADE/FDE golden with injected noise ----------------
    ns = extract_test_fn()
    real_std_normal = torchdist._standard_normal
    Sg = {"names": np.array(names)}
    m.load_state_dict(sd)
    feed = {}

    def fake_std_normal(shape, dtype, device):
        e = feed["eps"][feed["k"]]
        feed["k"] += 1
        assert tuple(shape) == tuple(e.shape), (shape, e.shape)
        return torch.from_numpy(e.copy())

    tot_a, tot_f, tot_n = 0.0, 0.0, 0
    worst_s = 0.0
    n_exact = 0
    for name in names:
        tup = tuples[name]
        n = tup[0].shape[0]
        g = torch.Generator().manual_seed(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        eps = torch.randn(20, 12, n, 2, generator=g).numpy()
        feed["eps"], feed["k"] = eps, 0
        ns["loader_test"] = DataLoader(ListDataset([tup]), batch_size=1, shuffle=False, num_workers=0)
        ns["model"] = m
        torchdist._standard_normal = fake_std_normal
        try:
            ade_, fde_, raw = ns["test"](KSTEPS=20)
        finally:
            torchdist._standard_normal = real_std_normal
        assert feed["k"] == 20
        preds = np.stack([np.asarray(p, np.float32).reshape(12, n, 2) for p in raw[1]["pred"]])
        trgt = np.asarray(raw[1]["trgt"], np.float32).reshape(12, n, 2)
        # per-pedestrian minima, via the reference's own ade/fde on its own trajectories
        pa = np.array([[ref_metrics.ade([preds[k][:, i:i + 1]], [trgt[:, i:i + 1]], [1]) for i in range(n)] for k in range(20)])
        pf = np.array([[ref_metrics.fde([preds[k][:, i:i + 1]], [trgt[:, i:i + 1]], [1]) for i in range(n)] for k in range(20)])
        assert abs(pa.min(0).mean() - ade_) < 1e-12 and abs(pf.min(0).mean() - fde_) < 1e-12
        Sg[f"{name}.eps"] = eps
        Sg[f"{name}.min_ade"] = pa.min(0)
        Sg[f"{name}.min_fde"] = pf.min(0)
        if n <= 9:
            Sg[f"{name}.pred_abs"] = preds
        # oracle, fed the REFERENCE's V_pred (isolates the sampling/metric restatement)
        with torch.no_grad():
            vp, _ = m(tup[6][None].permute(0, 3, 1, 2), tup[7])
        vp = vp.permute(0, 2, 3, 1).squeeze(0).numpy()
        oa, of, op = O.sample_ade_fde(vp, tup[8].numpy(), tup[0][:, :, -1].numpy(), eps)
        # MKL's potrf L10 is +-2 ulp off the closed form (see oracle.cholesky_2x2): not bit-exact.
        perr = np.abs(op - preds).max() / max(1.0, np.abs(preds).max())
        aerr = np.abs(oa - pa.min(0)).max() / pa.min(0).max()
        ferr = np.abs(of - pf.min(0)).max() / pf.min(0).max()
        worst_s = max(worst_s, perr, aerr, ferr)
        assert perr <= 1e-6 and aerr <= 2e-6 and ferr <= 2e-6, (name, perr, aerr, ferr)
        # ... and, fed the reference's own sampled trajectories, equal up to the reference's
        # numpy-scalar `**2` (glibc powf: 1 fp32 ulp off x*x on ~0.07 % of values, metrics.py:33,50).
        tg = O.rel_to_abs(tup[8].numpy(), tup[0][:, :, -1].numpy())
        assert np.array_equal(tg, trgt), name
        dxy = preds - trgt[None]
        dd = np.sqrt((dxy[..., 0] * dxy[..., 0] + dxy[..., 1] * dxy[..., 1]).astype(np.float64))
        acc = np.zeros((20, n))
        for t in range(12):
            acc = acc + dd[:, t]
        e1 = np.abs(dd[:, -1].min(0) - pf.min(0)).max() / pf.min(0).max()
        e2 = np.abs((acc / 12).min(0) - pa.min(0)).max() / pa.min(0).max()
        n_exact += int(np.array_equal(dd[:, -1].min(0), pf.min(0)))
        assert e1 <= 1.3e-7 and e2 <= 1.3e-7, (name, e1, e2)
        if name.startswith("eth"):
            tot_a += pa.min(0).sum(); tot_f += pf.min(0).sum(); tot_n += n
    Sg["eth.ade"] = np.array([tot_a / tot_n])
    Sg["eth.fde"] = np.array([tot_f / tot_n])
    print(f"sampling: oracle vs reference worst rel err {worst_s:.2e}; metric restatement bit-exact FDE on {n_exact}/{len(names)} scenes")
    print(f"sampling: eth/test injected-noise ADE {tot_a/tot_n:.6f} FDE {tot_f/tot_n:.6f} over {tot_n} peds")
    np.savez_compressed(f"{HERE}/sample_golden.npz", **Sg)

    # timing of the unmodified reference per scene (reported next to the CPU baseline)
    tup = tuples["syn64"]
    t0 = time.time()
    ref_utils.seq_to_graph(tup[0], tup[2], True)
    tg = time.time() - t0
    print(f"reference seq_to_graph N=64, T=8: {tg:.2f}s/scene on this box")
    with open(f"{HERE}/REFERENCE_TIMING.txt", "w") as f:
        f.write(f"unmodified reference, this container ({os.cpu_count()} cores): seq_to_graph N=64 T=8 {tg:.3f} s/scene\n")
    print(f"done in {time.time()-t00:.0f}s")


if __name__ == "__main__":
    main()
