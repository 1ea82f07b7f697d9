"""social_stgcnn_b200 -- B200-native (sm_100a) hot path of Social-STGCNN.

Drop-in surface (same names and signatures as the reference's utils.py / model.py /
metrics.py):  seq_to_graph, TrajectoryDataset, social_stgcnn, bivariate_loss, ade, fde.
Packed-batch surface: ScenePlan, InferenceEngine and the functional ops in `engine`.
Everything computes in libsstg.so (hand-written CUDA, C ABI in include/sstg.h); there
is no CPU fallback.
"""
from ._lib import LIB_PATH, load as load_library  # noqa: F401
from .engine import (DevicePlan, InferenceEngine, ScenePlan, TrainEngine, backward,  # noqa: F401
                     bn_running_update, decode_tracks, pack_tracks, forward_train, graph_build, group_weights, nll_backward, nll_forward,
                     sample_ade_fde, sgd_step, stgcn_forward, stgcn_forward_fused, txpcnn_forward)

__all__ = ["ScenePlan", "DevicePlan", "InferenceEngine", "TrainEngine", "graph_build", "stgcn_forward",
           "stgcn_forward_fused", "txpcnn_forward", "nll_forward", "sample_ade_fde", "forward_train", "nll_backward", "backward",
           "bn_running_update", "sgd_step", "pack_tracks", "decode_tracks", "group_weights", "load_library", "LIB_PATH"]
