#!/usr/bin/env python
"""Per-phase cycle budget of the fused solver (tuning build):
    RDPN_NVCC_EXTRA=-DRDPN_PHASE_CLOCKS python -m rdpn6d_b200.build --force && python benchmarks/phase_clocks.py [B] [R]
Thread 0 of every CTA stamps clock64() at the phase boundaries; prints mean cycles per phase per ROI."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from rdpn6d_b200 import _lib, pose_solver, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
R = int(sys.argv[2]) if len(sys.argv) > 2 else 64
NAMES = ["1 prefetch+minmax", "2 gate", "3a histogram", "3b cursors (warp 0)", "3c slots", "4 hypotheses (+barrier)",
         "5 staging", "6 scoring (+barrier)", "7a best", "7b moments", "7b solve", "outputs"]
models = synth.make_models(8, R, seed=1)
b = synth.tile_batch(synth.make_batch(128, models=models, H=256, seed=20260101, occlusion_max=0.6), B)
s = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in b.items()}
plan = pose_solver.make_plan(pose_solver.PoseSolver(inlier_thr=0.005), s["depth"], s["Kp"], s["coor"][:, 0].contiguous(),
                             s["coor"][:, 1].contiguous(), s["coor"][:, 2].contiguous(), s["mask"], s["extent"], s["hyp_idx"],
                             s["region_idx"], s["anchors"])
L = _lib.lib()
buf = torch.zeros(B, 16, dtype=torch.int64, device="cuda")
L.rdpn_debug_set_phase_clocks.argtypes = [ctypes.c_void_p]
assert L.rdpn_debug_set_phase_clocks(buf.data_ptr()) == 0
for _ in range(3):
    plan.launch()
torch.cuda.synchronize()
c = buf.cpu().double()
d = (c[:, 1:13] - c[:, 0:12])
tot = (c[:, 12] - c[:, 0])
# steady state: ignore the first and last wave
lo, hi = int(0.15 * B), int(0.85 * B)
out = {n: round(float(d[lo:hi, i].mean())) for i, n in enumerate(NAMES)}
out["total_cycles_per_cta"] = round(float(tot[lo:hi].mean()))
print(json.dumps(out, indent=1))
