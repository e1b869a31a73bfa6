#!/usr/bin/env python
"""Deployment split of the reference (head outputs on the GPU, loader tensors pinned on the host) through
HostPoseSolver, swept over the pipeline chunk size.  One JSON object per line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rdpn6d_b200 import pose_solver  # noqa: E402

B = bench.ROIS_PER_GPU
batch = bench.make_workload()
pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in batch.items() if v is not None}
dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in batch.items() if v is not None}
cx, cy, cz = [dev["coor"][:, c].contiguous() for c in range(3)]
for chunk in (128, 256, 512, 1024):
    hs = pose_solver.HostPoseSolver(inlier_thr=bench.INLIER_THR, chunk_rois=chunk, count_bytes=True)
    args = (pin["depth"], pin["Kp"], cx, cy, cz, dev["mask"], pin["extent"], pin["hyp_idx"], dev["region_idx"], pin["anchors"])
    call = hs.plan(*args)
    call()
    nbytes = hs.last_h2d_bytes
    hs.set_option(4, 0)
    for _ in range(3):
        call()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    dt = (time.perf_counter() - t0) / n
    print(json.dumps({"bench": "host_mixed", "chunk": chunk, "ms": 1e3 * dt, "rois_per_s": B / dt, "h2d_bytes": nbytes}))
    hs.close()
