#!/usr/bin/env python
"""Golden vectors for BASELINE.json config 5 (nuScenes_mini graph construction), made by the REAL reference.

Run in the build container only (needs /root/reference, read-only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_nuscenes_golden.py

The shipped datasets/nuScenes_mini/ holds no trajectories and the nuScenes devkit is absent, so the
reference's own stand-in is used: `complete_nuscenes_setup.create_dummy_data()` (complete_nuscenes_setup.py:240-290;
5 / 2 / 1 files, 2-5 pedestrians, 50-99 frames, seeded with np.random.seed(i + ord(split[0]))), executed
unmodified in a scratch directory.  Its output files are committed as a DATA fixture under
tests/golden/datasets/nuscenes_mini_dummy/ (space-delimited ETH format, as the reference writes it).  Then the
unmodified reference `TrajectoryDataset(..., delim=' ')` (utils.py:86-228, with the real networkx graph build)
is run on the test and val splits and its per-scene V_obs / A_obs / V_tr / A_tr are stored in
tests/golden/nuscenes_golden.npz; the windowing digests of all three splits are appended to
tests/golden/windowing_golden.json.  The script aborts if utils.load_scenes disagrees with the reference.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import complete_nuscenes_setup as ref_setup   # noqa: E402  (the reference)
import utils as ref_utils                     # noqa: E402  (the reference)
from make_windowing_golden import digest      # noqa: E402

OUT = os.path.join(HERE, "datasets", "nuscenes_mini_dummy")


def ours(path):
    from social_stgcnn_b200.utils import load_scenes
    seqs, rels, masks, nl, counts, max_peds = load_scenes(path, 8, 12, 1, 0.002, 1, " ")
    seq_all, rel_all = np.concatenate(seqs), np.concatenate(rels)
    cum = [0] + np.cumsum(counts).tolist()
    return {"scenes": len(seqs), "peds": int(seq_all.shape[0]), "max_peds_in_frame": int(max_peds),
            "sha256": digest(seq_all[:, :, :8], seq_all[:, :, 8:], rel_all[:, :, :8], rel_all[:, :, 8:],
                             np.concatenate(nl), list(zip(cum, cum[1:])))}


def main():
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            ref_setup.setup_directories()
            ref_setup.create_dummy_data()
        finally:
            os.chdir(cwd)
        if os.path.isdir(OUT):
            shutil.rmtree(OUT)
        shutil.copytree(os.path.join(tmp, "datasets", "nuscenes_mini"), OUT)
    gold = {}
    digests = json.load(open(os.path.join(HERE, "windowing_golden.json")))
    real = ref_utils.seq_to_graph
    for split in ("test", "val", "train"):
        path = os.path.join(OUT, split) + "/"
        if split == "train":    # windowing only: the networkx graph build of 300+ scenes adds nothing new
            ref_utils.seq_to_graph = lambda a, b, c=True: (torch.zeros(1), torch.zeros(1))
        d = ref_utils.TrajectoryDataset(path, obs_len=8, pred_len=12, skip=1, norm_lap_matr=True, delim=" ")
        ref_utils.seq_to_graph = real
        ref = {"scenes": len(d), "peds": int(d.obs_traj.shape[0]), "max_peds_in_frame": int(d.max_peds_in_frame),
               "sha256": digest(d.obs_traj.numpy(), d.pred_traj.numpy(), d.obs_traj_rel.numpy(),
                                d.pred_traj_rel.numpy(), d.non_linear_ped.numpy(), d.seq_start_end)}
        mine = ours(path)
        print(split, ref, "OK" if mine == ref else f"MISMATCH {mine}", flush=True)
        if mine != ref:
            raise SystemExit("utils.load_scenes does not reproduce the reference's windowing")
        digests[f"nuscenes_mini_dummy/{split}"] = ref
        if split != "train":
            gold[f"{split}.n_peds"] = np.array([e - s for s, e in d.seq_start_end], np.int32)
            for name, lst in (("V_obs", d.v_obs), ("A_obs", d.A_obs), ("V_tr", d.v_pred), ("A_tr", d.A_pred)):
                gold[f"{split}.{name}"] = np.concatenate([np.asarray(t.numpy(), np.float32).ravel() for t in lst])
    np.savez_compressed(os.path.join(HERE, "nuscenes_golden.npz"), **gold)
    with open(os.path.join(HERE, "windowing_golden.json"), "w") as fp:
        json.dump(digests, fp, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
