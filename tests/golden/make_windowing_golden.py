#!/usr/bin/env python
"""Pin utils.load_scenes (the vectorised windowing) to the REAL reference on every shipped split.

Run in the build container only (needs /root/reference, read-only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_windowing_golden.py

For every datasets/<name>/<split> directory of the reference, the unmodified reference
TrajectoryDataset (utils.py:86-228, graph build stubbed -- it is not part of the windowing) is run and a
digest of what the windowing produces is written to tests/golden/windowing_golden.json: scene count,
pedestrian count, max_peds_in_frame and the SHA-256 of obs_traj | pred_traj | obs_traj_rel | pred_traj_rel |
non_linear_ped | seq_start_end.  tests/test_host_logic.py recomputes the digests with utils.load_scenes (when
/root/reference is present; the files do not travel).  The script aborts if our windowing differs.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
import utils as ref_utils  # noqa: E402  (the reference)


def digest(obs, pred, obs_rel, pred_rel, nl, sse):
    h = hashlib.sha256()
    for a in (obs, pred, obs_rel, pred_rel, nl):
        h.update(np.ascontiguousarray(a, np.float32).tobytes())
    h.update(np.ascontiguousarray(np.asarray(sse, np.int64)).tobytes())
    return h.hexdigest()


def ours(path):
    from social_stgcnn_b200.utils import load_scenes
    seqs, rels, masks, nl, counts, max_peds = load_scenes(path, 8, 12, 1, 0.002, 1, "\t")
    seq_all, rel_all = np.concatenate(seqs), np.concatenate(rels)
    cum = [0] + np.cumsum(counts).tolist()
    return {"scenes": len(seqs), "peds": int(seq_all.shape[0]), "max_peds_in_frame": int(max_peds),
            "sha256": digest(seq_all[:, :, :8], seq_all[:, :, 8:], rel_all[:, :, :8], rel_all[:, :, 8:],
                             np.concatenate(nl), list(zip(cum, cum[1:])))}


def main():
    ref_utils.seq_to_graph = lambda a, b, c=True: (torch.zeros(1), torch.zeros(1))
    out = {}
    for name in ("eth", "hotel", "univ", "zara1", "zara2"):
        for split in ("test", "val", "train"):
            path = f"{REF}/datasets/{name}/{split}/"
            d = ref_utils.TrajectoryDataset(path, obs_len=8, pred_len=12, skip=1, norm_lap_matr=True)
            ref = {"scenes": len(d), "peds": int(d.obs_traj.shape[0]), "max_peds_in_frame": int(d.max_peds_in_frame),
                   "sha256": digest(d.obs_traj.numpy(), d.pred_traj.numpy(), d.obs_traj_rel.numpy(),
                                    d.pred_traj_rel.numpy(), d.non_linear_ped.numpy(), d.seq_start_end)}
            mine = ours(path)
            print(name, split, ref, "OK" if mine == ref else f"MISMATCH {mine}", flush=True)
            if mine != ref:
                raise SystemExit("utils.load_scenes does not reproduce the reference's windowing")
            out[f"{name}/{split}"] = ref
    with open(os.path.join(HERE, "windowing_golden.json"), "w") as fp:
        json.dump(out, fp, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
