"""Shared helpers for the tests: golden fixtures (generated from the real reference by
tests/golden/make_golden.py) and packing of golden scenes into a ragged batch."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


class Golden:
    def __init__(self):
        self.graph = np.load(os.path.join(GOLD, "graph_golden.npz"))
        self.fwd = np.load(os.path.join(GOLD, "forward_golden.npz"))
        self.loss = np.load(os.path.join(GOLD, "loss_golden.npz"))
        self.sample = np.load(os.path.join(GOLD, "sample_golden.npz"))
        self.weights = np.load(os.path.join(GOLD, "weights.npz"))
        self.names = [str(n) for n in self.graph["names"]]

    def params(self, ckpt="eth"):
        return self.weights[f"{ckpt}.params"], self.weights[f"{ckpt}.bn"]

    def pack(self, names=None):
        """Concatenate golden scenes the way TrajectoryDataset packs pedestrians."""
        names = list(names or self.names)
        g = self.graph
        n_peds = [g[f"{n}.obs_rel"].shape[0] for n in names]
        obs_rel = np.concatenate([g[f"{n}.obs_rel"] for n in names]).astype(np.float32)
        pred_rel = np.concatenate([g[f"{n}.pred_rel"] for n in names]).astype(np.float32)
        last_obs = np.concatenate([g[f"{n}.obs"][:, :, -1] for n in names]).astype(np.float32)
        return names, np.array(n_peds, np.int32), obs_rel, pred_rel, last_obs


def rel_err(a, b, scale=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    s = max(np.abs(b).max(), 1e-30) if scale is None else scale
    return float(np.abs(a - b).max() / s)
