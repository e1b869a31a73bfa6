#!/usr/bin/env python
"""Host-buffer plugin call (rdpn_pose_solve_host) on the bench workload (YCB-V maps, 4096 ROIs per call): full copy vs
gated pull, swept over the pull granularity and the pipeline chunk size.  One JSON object per line."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rdpn6d_b200 import _lib  # noqa: E402


def main():
    W = bench.WORKLOADS["ycbv"]
    B, H, R = 4096, W["H"], W["R"]
    batch = bench.tile(bench.make_base("ycbv"), B)
    L = _lib.lib()
    pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in batch.items()
           if v is not None and k in ("depth", "Kp", "mask", "extent", "region_idx", "anchors", "hyp_idx")}
    for c, name in enumerate(("coor_x", "coor_y", "coor_z")):
        pin[name] = torch.from_numpy(np.ascontiguousarray(batch["coor"][:, c])).pin_memory()
    h_pose = torch.empty(B, 12, dtype=torch.float32).pin_memory()
    h_ninl = torch.empty(B, dtype=torch.int32).pin_memory()
    h_stat = torch.empty(B, dtype=torch.int32).pin_memory()
    inp = _lib.RoiInputs(depth=pin["depth"].data_ptr(), Kp=pin["Kp"].data_ptr(), depth_div=None,
                         coor_x=pin["coor_x"].data_ptr(), coor_y=pin["coor_y"].data_ptr(), coor_z=pin["coor_z"].data_ptr(),
                         mask=pin["mask"].data_ptr(), extent=pin["extent"].data_ptr(), region_idx=pin["region_idx"].data_ptr(),
                         anchors=pin["anchors"].data_ptr(), num_regions=R, mask_mode=1, mask_thr=0.5, B=B)
    prm = _lib.SolveParams(inlier_thr=bench.INLIER_THR, num_hyp=H, min_pts=4, min_inliers=4, weighted=1, refit_iters=1,
                           with_scale=0, adaptive=0, confidence=0.995, min_iter=10)
    outs = _lib.SolveOutputs(pose=h_pose.data_ptr(), n_inliers=h_ninl.data_ptr(), status=h_stat.data_ptr())
    ctx = ctypes.c_void_p()
    _lib.check(L.rdpn_ctx_create(0, ctypes.byref(ctx)), "ctx_create")

    def call():
        _lib.check(L.rdpn_pose_solve_host(ctx, ctypes.byref(inp), pin["hyp_idx"].data_ptr(), None, ctypes.byref(prm),
                                          ctypes.byref(outs)), "pose_solve_host")

    ref = None
    combos = [(_lib.TRANSFER_COPY, 2, 256)] + [(_lib.TRANSFER_PULL, g, c) for c in (128, 256, 512, 1024) for g in (1, 2, 4, 8)]
    for mode, gran, chunk in combos:
        L.rdpn_ctx_set_option(ctx, _lib.OPT_TRANSFER, mode)
        L.rdpn_ctx_set_option(ctx, _lib.OPT_PULL_GRANULARITY, gran)
        L.rdpn_ctx_set_option(ctx, _lib.OPT_CHUNK_ROIS, chunk)
        L.rdpn_ctx_set_option(ctx, _lib.OPT_COUNT_BYTES, 1)
        call()
        nbytes = int(L.rdpn_ctx_last_h2d_bytes(ctx))
        L.rdpn_ctx_set_option(ctx, _lib.OPT_COUNT_BYTES, 0)
        for _ in range(3):
            call()
        n = 10
        t0 = time.perf_counter()
        for _ in range(n):
            call()
        dt = (time.perf_counter() - t0) / n
        if ref is None:
            ref = h_pose.clone()
        print(json.dumps({"bench": "host_path", "transfer": "copy" if mode == _lib.TRANSFER_COPY else "pull", "gran_quads": gran,
                          "chunk": chunk, "ms": 1e3 * dt, "rois_per_s": B / dt, "h2d_bytes": nbytes,
                          "GBps": nbytes / dt / 1e9, "identical": bool(torch.equal(ref, h_pose))}))
    L.rdpn_ctx_destroy(ctx)


if __name__ == "__main__":
    main()
