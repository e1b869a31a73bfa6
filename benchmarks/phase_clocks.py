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
# phase overlap per SM: how many resident CTAs are inside the scoring phase at the same time?
import numpy as np
cn = c.numpy()
smid = cn[:, 15].astype(int)
s0, s1 = cn[:, 7], cn[:, 8]      # scoring interval
b0, b1 = cn[:, 0], cn[:, 12]     # CTA lifetime
hist = np.zeros(8)
res_hist = np.zeros(8)
for sm in np.unique(smid):
    idx = np.nonzero(smid == sm)[0]
    t0, t1 = np.quantile(b0[idx], 0.2), np.quantile(b1[idx], 0.8)
    ts = np.linspace(t0, t1, 400)
    k = ((s0[idx][None, :] <= ts[:, None]) & (ts[:, None] < s1[idx][None, :])).sum(1)
    r = ((b0[idx][None, :] <= ts[:, None]) & (ts[:, None] < b1[idx][None, :])).sum(1)
    hist += np.bincount(np.minimum(k, 7), minlength=8)
    res_hist += np.bincount(np.minimum(r, 7), minlength=8)
print("CTAs resident per SM (time share):", (res_hist / res_hist.sum()).round(3).tolist())
print("CTAs in the scoring phase per SM (time share):", (hist / hist.sum()).round(3).tolist())
lt = (b1 - b0)[lo:hi]
print("lifetime mean %.0f std %.0f min %.0f max %.0f" % (lt.mean(), lt.std(), lt.min(), lt.max()))
