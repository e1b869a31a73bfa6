import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from rdpn6d_b200 import pose_solver, synth
models = synth.make_models(8, 32, seed=1)
base = synth.make_batch(128, models=models, H=256, seed=20260101, occlusion_max=0.6)
b = synth.tile_batch(base, 8192)
s = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in b.items()}
solver = pose_solver.PoseSolver(inlier_thr=0.005)
plan = pose_solver.make_plan(solver, s["depth"], s["Kp"], s["coor"][:, 0].contiguous(), s["coor"][:, 1].contiguous(), s["coor"][:, 2].contiguous(), s["mask"], s["extent"], s["hyp_idx"], s["region_idx"], s["anchors"])
for i in range(4): plan.launch()
torch.cuda.synchronize()
