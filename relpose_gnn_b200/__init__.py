"""relpose_gnn_b200 -- B200-native (sm_100a) message-passing hot path of RelPose-GNN.

Public surface (mirrors the reference's Python class API for this path):
  simpleConvEdge_upt   drop-in for niantic.modules.my_gnn_layer.simpleConvEdge_upt (my_gnn_layer.py:277-311)
  RelPoseGNN           the GNN part of PoseNetX_R2.forward (posenet.py:1053-1091)
  PoseNetCriterion     criterion.py:33-60 fused with compute_RP (posenet.py:1021-1031)
  graph.GraphBatch     implicit per-graph edge template (FC enumeration + batch-shared edge dropout)
Everything executes in librpg_b200.so (C ABI: include/rpg.h).  Importing this package does not require a GPU;
calling any op without the built library or without CUDA tensors raises.
"""
from . import graph  # noqa: F401
from . import parallel  # noqa: F401
from .graph import (GraphBatch, apply_edge_mask, attach, batched_edge_index, check_pending, edge_dropout_keep,  # noqa: F401
                    fc_template, knn_graph, mask_edge_index, set_validation)
from .layers import simpleConv, simpleConvEdge, simpleConvEdge_upt  # noqa: F401
from .evaluation import compose_query_pose, pose_errors, qexp, save_poses  # noqa: F401
from .feed import DeviceFeeder, ScalarReadback  # noqa: F401
from .graph_io import collate, load_graph  # noqa: F401
from .model import PoseNetCriterion, RelPoseGNN  # noqa: F401
from .optim import FusedAdam  # noqa: F401

__all__ = ["FusedAdam", "load_graph", "collate", "mask_edge_index", "set_validation", "check_pending", "DeviceFeeder", "ScalarReadback", "apply_edge_mask", "compose_query_pose", "pose_errors", "save_poses", "qexp", "knn_graph", "simpleConv", "simpleConvEdge", "simpleConvEdge_upt", "RelPoseGNN", "PoseNetCriterion", "GraphBatch", "attach", "batched_edge_index",
           "edge_dropout_keep", "fc_template", "graph"]
