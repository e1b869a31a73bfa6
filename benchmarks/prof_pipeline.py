"""Launch the solve a few times on a BASELINE workload (for `ncu -k regex:... -s N -c 1` captures), then S1
(rdpn_correspond) on the same maps.
usage: prof_pipeline.py [config=ycbv] [pipeline=split] [chunk=0] [weighted=1]"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "benchmarks"))
from pipeline import CONFIGS, sets_for  # noqa: E402
from rdpn6d_b200 import pose_solver  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ycbv"
pipe = sys.argv[2] if len(sys.argv) > 2 else "split"
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
weighted = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
s = sets_for(name, 1)[0]
solver = pose_solver.PoseSolver(inlier_thr=0.005, weighted=weighted, pipeline=pipe, chunk_rois=chunk)
plan = pose_solver.make_plan(solver, s["depth"], s["Kp"], s["coor"][:, 0].contiguous(), s["coor"][:, 1].contiguous(),
                             s["coor"][:, 2].contiguous(), s["mask"], s["extent"], s["hyp_idx"], s["region_idx"], s["anchors"])
for i in range(3):
    plan.launch()
torch.cuda.synchronize()
pose_solver.correspond(s["depth"], s["Kp"], s["coor"][:, 0].contiguous(), s["coor"][:, 1].contiguous(), s["coor"][:, 2].contiguous(),
                       s["mask"], s["extent"], s["region_idx"], s["anchors"])
pose_solver.correspond(s["depth"], s["Kp"], s["coor"][:, 0].contiguous(), s["coor"][:, 1].contiguous(), s["coor"][:, 2].contiguous(),
                       s["mask"], s["extent"], s["region_idx"], s["anchors"])
torch.cuda.synchronize()
