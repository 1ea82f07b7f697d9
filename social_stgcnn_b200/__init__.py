"""social_stgcnn_b200 -- B200-native (sm_100a) hot path of Social-STGCNN.

Drop-in surface (same names and signatures as the reference's utils.py / model.py /
metrics.py):  seq_to_graph, TrajectoryDataset, social_stgcnn, bivariate_loss, ade, fde.
Packed-batch surface: ScenePlan, InferenceEngine and the functional ops in `engine`.
Everything computes in libsstg.so (hand-written CUDA, C ABI in include/sstg.h); there
is no CPU fallback.
"""
from ._lib import LIB_PATH, load as load_library  # noqa: F401
from .engine import (DevicePlan, InferenceEngine, ScenePlan, graph_build, nll_forward,  # noqa: F401
                     sample_ade_fde, stgcn_forward, txpcnn_forward)

__all__ = ["ScenePlan", "DevicePlan", "InferenceEngine", "graph_build", "stgcn_forward", "txpcnn_forward",
           "nll_forward", "sample_ade_fde", "load_library", "LIB_PATH"]
