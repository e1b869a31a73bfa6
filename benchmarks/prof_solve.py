"""Launch the fused solver a few times on the bench workload (for an `ncu -k regex:pose_solve -s 2 -c 1` capture).
usage: prof_solve.py [B=1024] [R=64] [H=256] [S=3]   (S > 3: samples of S pairs drawn by the kernel)"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from rdpn6d_b200 import pose_solver, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
S = int(sys.argv[4]) if len(sys.argv) > 4 else 3
models = synth.make_models(8, R, seed=1)
base = synth.make_batch(128, models=models, H=H, seed=20260101, occlusion_max=0.6)
b = synth.tile_batch(base, B)
s = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in b.items()}
solver = pose_solver.PoseSolver(inlier_thr=0.005, num_hyp=H, sample_size=S, seed=1)
plan = pose_solver.make_plan(solver, s["depth"], s["Kp"], s["coor"][:, 0].contiguous(), s["coor"][:, 1].contiguous(),
                             s["coor"][:, 2].contiguous(), s["mask"], s["extent"], s["hyp_idx"] if S == 3 else None, s["region_idx"],
                             s["anchors"])
for i in range(4):
    plan.launch()
torch.cuda.synchronize()
