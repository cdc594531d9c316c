"""fp32 mode (split-bf16 arithmetic) against the fp64 oracle: relative errors of the poses, next to the bf16 mode.
usage: python tests/diag/fp32_accuracy.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import relpose_gnn_b200 as rpg  # noqa: E402
from oracle import restatement as R  # noqa: E402  (test infrastructure: this tool is a checker)

dev = torch.device("cuda:0")
D, N, G = 512, 9, 24
case = R.synth_stack_case(D, N, G, 5150, droprate=0.0)
pn_o, pe_o, _, _ = R.stack_forward(case["params"], case["x"], case["edge_index"], 2, 0.0)
for prec in ("bf16", "fp32"):
    model = rpg.RelPoseGNN(D, D, D, droprate=0.0).to(dev)
    model.load_state_dict({k: v.float() for k, v in case["params"].items()}, strict=False)
    model.precision = prec
    model.gnn1.precision = prec
    with torch.no_grad():
        pn, pe, _ = model(case["x"].float().to(dev), case["edge_index"].to(dev))
    rel = lambda a, b: ((a.double().cpu() - b).norm() / b.norm()).item()  # noqa: E731
    print(f"{prec}: pose_nodes rel {rel(pn, pn_o):.2e}, pose_edges rel {rel(pe, pe_o):.2e}")
