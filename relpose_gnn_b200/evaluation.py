"""Evaluation composition of the reference (test.py:222-243): from the predicted relative poses of the edges into the
query image (node 0 of every graph) and the known absolute pose of a reference image to the absolute pose of the query,
with the log-quaternion mapped to a unit quaternion (pose_utils.qexp, pose_utils.py:340-348).  The reference does this
in numpy, one graph per batch; here it is one small kernel over all graphs of the batch."""
import ctypes as C

import numpy as np
import torch

from . import _lib, graph as graph_mod
from .ops import _stream, check


def qexp(v):
    """[n, 3] log-quaternions (CUDA tensor) -> [n, 4] unit quaternions [cos|v|, sinc(|v|/pi) v]."""
    if not v.is_cuda:
        raise ValueError("relpose_gnn_b200.qexp needs a CUDA tensor")
    v = v.float().contiguous().reshape(-1, 3)
    q = torch.empty(v.size(0), 4, dtype=torch.float32, device=v.device)
    check(_lib.load().rpg_qexp(v.data_ptr(), v.size(0), q.data_ptr(), _stream(v)), "rpg_qexp")
    return q


def compose_query_pose(pred_edges, poses_abs, edge_index, ref_node=0, pose_m=None, pose_s=None):
    """pred_edges [Et, 6] (output_R of the model), poses_abs [Nt, 6] (data.y), edge_index of the batch.
    Returns (pred [G, 7], target [G, 7]) = (t, unit quaternion) of every graph's query image (node 0), translations
    un-normalised with (pose_s, pose_m) as in test.py:241-243.  `ref_node` picks the ref_node-th edge into node 0."""
    if not pred_edges.is_cuda:
        raise ValueError("relpose_gnn_b200.compose_query_pose needs CUDA tensors")
    g = graph_mod.from_edge_index(edge_index, poses_abs.size(0))
    into0 = np.flatnonzero(g.host_template()[1] == 0)
    if ref_node >= into0.size:
        raise ValueError(f"ref_node {ref_node}: only {into0.size} edges end in node 0")
    ref_k = int(into0[ref_node])
    pred_edges = pred_edges.float().contiguous()
    poses_abs = poses_abs.float().contiguous()
    out_p = torch.empty(g.G, 7, dtype=torch.float32, device=pred_edges.device)
    out_t = torch.empty(g.G, 7, dtype=torch.float32, device=pred_edges.device)
    m = (C.c_float * 3)(*[float(x) for x in pose_m]) if pose_m is not None else None
    s = (C.c_float * 3)(*[float(x) for x in pose_s]) if pose_s is not None else None
    check(_lib.load().rpg_eval_compose(pred_edges.data_ptr(), poses_abs.data_ptr(), g.byref(), ref_k, m, s,
                                       out_p.data_ptr(), out_t.data_ptr(), _stream(pred_edges)), "rpg_eval_compose")
    return out_p, out_t


def pose_errors(pred7, targ7):
    """Per-pose translation error (same unit as t) and rotation error in degrees of [n, 7] = (t, unit quaternion) poses:
    the reference's `t_criterion` / `quaternion_angular_error` (test.py:202-203; pose_utils.py:420-431).  CUDA tensors."""
    if not pred7.is_cuda or not targ7.is_cuda:
        raise ValueError("relpose_gnn_b200.pose_errors needs CUDA tensors")
    pred7, targ7 = pred7.float().contiguous(), targ7.float().contiguous()
    if pred7.shape != targ7.shape or pred7.dim() != 2 or pred7.size(1) != 7:
        raise ValueError("pose_errors expects two [n, 7] tensors")
    t_err = torch.empty(pred7.size(0), dtype=torch.float32, device=pred7.device)
    q_err = torch.empty_like(t_err)
    check(_lib.load().rpg_pose_errors(pred7.data_ptr(), targ7.data_ptr(), pred7.size(0), t_err.data_ptr(), q_err.data_ptr(),
                                      _stream(pred7)), "rpg_pose_errors")
    return t_err, q_err


def save_poses(pred_poses, rel_paths, p_output, target_poses):
    """The reference's result file (test.py:38-42): an .npz with rel_path, abs_t, abs_q, targ_t, targ_q.
    pred_poses / target_poses: [n, 7] tensors or arrays."""
    pred = pred_poses.detach().cpu().numpy() if torch.is_tensor(pred_poses) else np.asarray(pred_poses)
    targ = target_poses.detach().cpu().numpy() if torch.is_tensor(target_poses) else np.asarray(target_poses)
    if len(rel_paths) != len(pred):
        raise ValueError(f"len(rel_paths): {len(rel_paths)} != {len(pred)} len(pred_poses)")
    np.savez(p_output, rel_path=[str(p) for p in rel_paths], abs_t=pred[:, :3], abs_q=pred[:, 3:], targ_t=targ[:, :3],
             targ_q=targ[:, 3:])
