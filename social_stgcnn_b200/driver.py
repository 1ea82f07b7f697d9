"""Packed-batch replacements for the reference's two scene loops: `test()` (test.py:18-127) and
`train()` / `vald()` (train.py:28-122).  Same inputs (a TrajectoryDataset, the weights), same
return values, but every scene of the split goes through the kernels in one packed batch
(evaluation) or one packed batch per optimiser group (training) instead of a batch-1
DataLoader loop.  Host code here is orchestration only (index gathers, bookkeeping).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib as L
from . import engine as E


def _packed(dataset, dev):
    pk = getattr(dataset, "packed", None)
    if pk is None or torch.device(pk["device"]) != dev:
        raise RuntimeError("dataset must be a social_stgcnn_b200.utils.TrajectoryDataset built on the engine's device")
    return pk


def _rel_to_abs(dp, rel_blocks, init, T):
    out = torch.empty_like(rel_blocks)
    dev = dp.device
    with torch.cuda.device(dev):
        L.check(L.load().sstg_rel_to_abs(rel_blocks.data_ptr(), init.data_ptr(), dp.seg_off.data_ptr(), dp.host.S,
                                         dp.host.P, T, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                "sstg_rel_to_abs")
    return out


def evaluate(dataset, engine: E.InferenceEngine, KSTEPS: int = 20, seed: int = 0, eps: Optional[torch.Tensor] = None,
             want_raw: bool = False, exact: bool = True, graph: bool = False):
    """test() of the reference (test.py:18-127) over a whole split.

    Returns (ade_, fde_, raw_data_dict): the means over ALL pedestrians of the per-pedestrian
    minima over KSTEPS samples (test.py:121-126), and -- if `want_raw` -- the reference's
    raw_data_dict {step: {'obs', 'trgt', 'pred': [KSTEPS arrays]}} of absolute trajectories
    (test.py:87-105), step counted from 1.  `exact` keeps the reference's fp64 rounding islands
    (graph, distances); `eps` injects the standard-normal noise (layout (K,12,N,2) per scene).
    `graph=True` (in-kernel noise, no raw output): the split's pipeline is captured once into a CUDA graph
    (InferenceEngine.capture, cached on the dataset per engine / KSTEPS / seed / mode) and replayed -- a test split
    of a few hundred scenes is launch-bound, and the replay is one driver call instead of five kernel launches."""
    dev = engine.device
    pk = _packed(dataset, dev)
    plan = dataset.plan_obs
    if graph and eps is None and not want_raw:
        cache = dataset.__dict__.setdefault("_graphs", {})
        key = (id(engine), int(KSTEPS), int(seed), bool(exact))
        g = cache.get(key)
        if g is None:
            g = cache[key] = engine.capture(plan, KSTEPS, seed=seed, graph_fp64=exact, fp64_dist=exact)
            g.replay(pk["obs_rel"], pk["pred_rel"], pk["last_obs"])     # the split's tensors into the static buffers, once
        else:
            g.replay()
        r = g.out
    else:
        r = engine.infer(plan, pk["obs_rel"], pk["pred_rel"], pk["last_obs"], KSTEPS, eps=eps, seed=seed,
                         graph_fp64=exact, fp64_dist=exact, want_pred=want_raw)
    P = plan.P
    ade_ = float(r["min_ade"].sum()) / P
    fde_ = float(r["min_fde"].sum()) / P
    raw: Dict[int, dict] = {}
    if want_raw:
        dp = plan.on(dev)
        first = dataset.obs_traj[:, :, 0].to(dev).double().contiguous()
        last = dataset.obs_traj[:, :, -1].to(dev).double().contiguous()
        obs_abs = _rel_to_abs(dp, r["V_obs"], first, L.T_OBS).cpu().numpy()
        tgt_abs = _rel_to_abs(dp, r["V_tr"], last, L.T_PRED).cpu().numpy()
        pred = r["pred_abs"].cpu().numpy()
        for s in range(plan.S):
            o, n = int(plan.seg_off[s]), int(plan.n_peds[s])
            pr = pred[o * KSTEPS * 24:(o + n) * KSTEPS * 24].reshape(KSTEPS, L.T_PRED, n, 2)
            raw[s + 1] = {"obs": obs_abs[o * 16:(o + n) * 16].reshape(L.T_OBS, n, 2).copy(),
                          "trgt": tgt_abs[o * 24:(o + n) * 24].reshape(L.T_PRED, n, 2).copy(),
                          "pred": [pr[k].copy() for k in range(KSTEPS)]}
    return ade_, fde_, raw


def epoch_rows(seg_off, n_peds, order):
    """Pedestrian rows of the packed dataset in epoch order, and each scene's [start, end) inside that
    gathered order: after ONE index_select with `rows`, every optimiser group is a contiguous slice."""
    order = np.asarray(order, np.int64)
    counts = np.asarray(n_peds, np.int64)[order]
    ped_end = np.cumsum(counts)
    ped_start = ped_end - counts
    total = int(ped_end[-1]) if order.size else 0
    rows = np.repeat(np.asarray(seg_off, np.int64)[order] - ped_start, counts) + np.arange(total, dtype=np.int64)
    return rows, ped_start, ped_end


def train_epoch(dataset, engine: E.TrainEngine, batch_size: int = 128, lr: float = 0.01,
                clip_grad: Optional[float] = None, order=None, trainer=None, check_every: int = 32):
    """train() of the reference (train.py:28-79) over one epoch.

    `order` is the scene order of the epoch (the DataLoader's shuffle; default: dataset order).
    Scenes are taken in groups of `batch_size`; each group is one packed forward/backward and one
    SGD step, with the reference's accumulation rule (the last scene of a group is forwarded but
    not scored, divisor batch_size; train.py:58-66).  Returns loss_batch / batch_count exactly as
    the reference does (sum of the group losses / number of scenes, train.py:76-79).
    `trainer` (parallel.DistributedTrainer) spreads every group over the ranks.  Every rank must then
    pass the SAME `order` (fit() broadcasts rank 0's shuffle): this is verified with one small collective
    per epoch, which also lines the ranks up before the first peer-memory exchange, and the exchange status
    is checked on all ranks together every `check_every` groups and at the end of the epoch."""
    dev = engine.device
    pk = _packed(dataset, dev)
    S = len(dataset)
    order = np.arange(S) if order is None else np.asarray(order)
    n_all = np.asarray(dataset.plan_obs.n_peds)
    seg = np.asarray(dataset.plan_obs.seg_off, np.int64)
    if S == 0:
        return 0.0
    sharded = trainer is not None and getattr(trainer, "world", 1) > 1
    if sharded:
        # Ranks that disagreed on the epoch order would shard DIFFERENT scenes of every group (some counted twice,
        # some never, BatchNorm rows overlapping) without any error.  The collective doubles as the barrier that
        # bounds host-side skew (dataset parsing, a checkpoint write, lazy module loading) before the first
        # exchange kernel starts its bounded wait.
        from . import parallel as Pl
        Pl.assert_same_across_ranks((np.asarray(order, np.int64), n_all[np.asarray(order, np.int64)],
                                     np.asarray([batch_size], np.int64)), trainer.group, dev,
                                    what="train_epoch: scene order / pedestrian counts / batch_size")
    # ---- the epoch is prepared in one pass (the per-group device time is ~0.13 ms, so anything done per
    # group on the host decides the epoch time): ONE gather of the pedestrian rows in epoch order -- every
    # group is then a contiguous slice -- and ONE upload of all the groups' offset tables
    rows, ped_start, ped_end = epoch_rows(seg, n_all, order)
    rows_d = torch.from_numpy(rows).to(dev)
    obs_e, pred_e = pk["obs_rel"].index_select(0, rows_d), pk["pred_rel"].index_select(0, rows_d)
    bounds = list(range(0, S, batch_size)) + [S]
    groups = [order[a:b] for a, b in zip(bounds[:-1], bounds[1:])]
    plans = E.plan_groups([n_all[g] for g in groups], dev) if trainer is None else None
    if trainer is not None and hasattr(trainer, "prepare"):
        trainer.prepare([n_all[g] for g in groups], batch_size)
    losses = torch.zeros(len(groups), dtype=torch.float32, device=dev)
    for gi, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
        # (a trailing group of ONE scene has no scored scene: zero loss, zero gradient -- the
        #  reference crashes there on a freed autograd graph, SURVEY a-8; no shipped split hits it)
        p0, p1 = int(ped_start[a]), int(ped_end[b - 1])
        obs, pred = obs_e[p0:p1], pred_e[p0:p1]
        if trainer is not None:
            losses[gi].copy_(trainer.step(n_all[groups[gi]], obs, pred, batch_size, lr, clip_grad, presliced=False))
            if sharded and check_every > 0 and (gi + 1) % check_every == 0 and gi + 1 < len(groups):
                trainer.check()         # a rank that never delivered: every rank raises, together
        else:
            losses[gi].copy_(engine.step(plans[gi], obs, pred, batch_size, lr, clip_grad))
    if trainer is not None:
        trainer.check()
    return float(losses.double().sum()) / S


def vald_accumulate(scene_loss, batch_size: int) -> float:
    """The bookkeeping of vald() (train.py:81-122) on the per-scene eval losses, in loader order: groups of
    `batch_size` scenes, the last scene of a group unscored, divisor batch_size, result = sum of the group
    losses / number of scenes.  A trailing group of exactly ONE scene (len % batch_size == 1) scores nothing
    itself: there the reference divides the PREVIOUS group's already divided loss by batch_size once more
    and adds that (train.py:115-119: `loss` still holds the last group's value); with no previous group it
    raises UnboundLocalError -- mirrored here."""
    loss = np.asarray(scene_loss, np.float64)
    S = int(loss.size)
    if S == 0:
        return 0.0
    tot, prev = 0.0, None
    for g0 in range(0, S, batch_size):
        g = loss[g0:g0 + batch_size]
        if g.size == 1:
            if prev is None:
                raise UnboundLocalError("cannot access local variable 'loss' where it is not associated with a "
                                        "value (vald() of train.py with a single scene)")
            cur = prev / batch_size
        else:
            cur = float(g[:-1].sum()) / batch_size
        tot += cur
        prev = cur
    return tot / S


def validate_epoch(dataset, engine: E.InferenceEngine, batch_size: int = 128):
    """vald() of the reference (train.py:81-122): eval-mode forward, same accumulation rule,
    no parameter update.  Returns loss_batch / batch_count.  The dataset's packed V_obs / A_obs / V_tr
    blocks (built once, bit-identical to the reference's graphs) are fed straight to the forward
    kernels: no graph build and no sampling per epoch."""
    dev = engine.device
    pk = _packed(dataset, dev)
    plan = dataset.plan_obs
    if plan.S == 0:
        return 0.0
    dp = plan.on(dev)
    ws = engine.workspace(plan)
    E.stgcn_forward(dp, pk["V_obs"], pk["A_obs"], engine.params, engine.bn, H=ws["H"])
    E.txpcnn_forward(dp, ws["H"], engine.params, out=ws["V_pred"])
    loss = E.nll_forward(dp, ws["V_pred"], pk["V_tr"]).double().cpu().numpy()
    return vald_accumulate(loss, batch_size)


def fit(train_set, val_set, engine: E.TrainEngine, *, num_epochs: int = 250, batch_size: int = 128, lr: float = 0.01,
        use_lrschd: bool = False, lr_sh_rate: int = 150, clip_grad: Optional[float] = None,
        checkpoint_dir: Optional[str] = None, args=None, seed: Optional[int] = None, trainer=None, log=print):
    """main() of the reference (train.py:124-246) on the packed path: per epoch a shuffled
    train_epoch and a validate_epoch, StepLR(step_size=lr_sh_rate, gamma=0.2) when `use_lrschd`,
    and -- in `checkpoint_dir` -- exactly the artefacts the reference writes: `args.pkl`,
    `val_best.pth` (the 40-key state_dict, on every new best validation loss), `metrics.pkl`,
    `constant_metrics.pkl`.  Returns (metrics, constant_metrics).
    With a `trainer` (parallel.DistributedTrainer) every epoch's shuffle is drawn on rank 0 and broadcast,
    and rank 0 alone writes the artefacts (each file is written to a temporary name and renamed)."""
    import os
    import pickle
    from . import parallel as Pl
    world, rank = (trainer.world, trainer.rank) if trainer is not None else (1, 0)
    group = trainer.group if trainer is not None else None
    rng = np.random.default_rng(seed)
    val_eng = E.InferenceEngine(engine.params.cpu().numpy(), engine.bn.cpu().numpy(), engine.device)
    writer = bool(checkpoint_dir) and rank == 0       # replicas are identical: rank 0 alone writes the artefacts

    def dump(name, obj, save=pickle.dump):
        # whole-file replace: a reader (or a crash) never sees a half-written checkpoint
        path = os.path.join(checkpoint_dir, name)
        tmp = f"{path}.tmp.{os.getpid()}"
        with open(tmp, "wb") as fp:
            save(obj, fp)
        os.replace(tmp, path)

    if writer:
        os.makedirs(checkpoint_dir, exist_ok=True)
        if args is not None:
            dump("args.pkl", args)
    metrics = {"train_loss": [], "val_loss": []}
    constant_metrics = {"min_val_epoch": -1, "min_val_loss": 9999999999999999}
    cur_lr = lr
    for epoch in range(num_epochs):
        order = rng.permutation(len(train_set))                 # DataLoader(shuffle=True), train.py:173-177
        if world > 1:                                           # every rank trains on RANK 0's shuffle
            order = Pl.broadcast_order(order, group, engine.device)
        metrics["train_loss"].append(train_epoch(train_set, engine, batch_size, cur_lr, clip_grad, order, trainer))
        val_eng.load_weights(engine.params, engine.bn)
        metrics["val_loss"].append(validate_epoch(val_set, val_eng, batch_size))
        if use_lrschd and (epoch + 1) % lr_sh_rate == 0:        # StepLR.step() once per epoch, train.py:199-200,224
            cur_lr *= 0.2
        if metrics["val_loss"][-1] < constant_metrics["min_val_loss"]:
            constant_metrics["min_val_loss"] = metrics["val_loss"][-1]
            constant_metrics["min_val_epoch"] = epoch
            if writer:
                dump("val_best.pth", engine.state_dict(), torch.save)
        if log and rank == 0:
            log(f"epoch {epoch}: train_loss {metrics['train_loss'][-1]:.6f} val_loss {metrics['val_loss'][-1]:.6f} "
                f"lr {cur_lr:g} {constant_metrics}")
        if writer:
            dump("metrics.pkl", metrics)
            dump("constant_metrics.pkl", constant_metrics)
    return metrics, constant_metrics
