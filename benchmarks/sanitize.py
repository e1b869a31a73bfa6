#!/usr/bin/env python
"""Small driver for compute-sanitizer runs (memcheck / racecheck / synccheck) over every kernel:
    compute-sanitizer --tool racecheck python benchmarks/sanitize.py
Keeps problem sizes tiny (the tools slow kernels down by 10-100x)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rdpn6d_b200 import fps_utils, geometry, pose_from_pred, pose_solver, synth  # noqa: E402


def main():
    torch.cuda.set_device(0)
    tc = lambda x: torch.from_numpy(x).cuda()
    # FPS: thread-block cluster (single and several CTAs), cooperative grid, batched objects, centre row
    for n, k in ((3000, 16), (20000, 12), (40000, 12)):
        fps_utils.fps_indices(tc(synth.fps_cloud(n, seed=1)), k)
    fps_utils.fps_indices_batch([tc(synth.fps_cloud(n, seed=n)) for n in (50, 700, 9000)], 10)
    fps_utils.get_fps_and_center(tc(synth.fps_cloud(5000, seed=2)), 8)
    # S1 + fused solver, anchor and dense mode, multi-chunk case included
    big = [synth.ObjectModel("box", [0.12, 0.12, 0.12], 32, np.random.default_rng(0))]
    for kw in (dict(), dict(dense=True), dict(models=big, dzi_pad_scale=1.0, mask_dropout=0.0)):
        b = synth.make_batch(5, H=64, seed=11, **kw)
        g = {k: (None if v is None else tc(v)) for k, v in b.items()}
        pose_solver.correspond(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"], g["extent"],
                               g["region_idx"], g["anchors"])
        for opts in (dict(), dict(weighted=True, refit_iters=2, adaptive=True)):
            pose_solver.pose_solve(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"], g["extent"],
                                   g["hyp_idx"], g["region_idx"], g["anchors"], want_inlier_mask=True, want_hyp=True, **opts)
        if g["region_idx"] is not None:
            # the three-kernel pipeline (front / score / refit with PDL hand-over flags), in chunks of 2 ROIs, both
            # selection rules, multi-chunk scoring included (the third batch has > 1024 gated points)
            for opts in (dict(), dict(weighted=True, refit_iters=2, adaptive=True), dict(select_rule="min_mean_err")):
                pose_solver.pose_solve(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"],
                                       g["extent"], g["hyp_idx"], g["region_idx"], g["anchors"], want_inlier_mask=True,
                                       want_hyp=True, pipeline="split", chunk_rois=2, **opts)
        # S = 10 pairs per hypothesis (the out-of-line Kabsch of the sample), explicit and kernel-drawn samples
        sel = pose_solver.correspond(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"],
                                     g["extent"], g["region_idx"], g["anchors"])["sel"]
        hyp10 = pose_solver.sample_hypotheses(sel, 32, sample_size=10)
        for h_in in (hyp10, None):
            pose_solver.pose_solve(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"], g["extent"],
                                   h_in, g["region_idx"], g["anchors"], want_hyp=True, num_hyp=32, sample_size=10)
            if g["region_idx"] is not None:
                pose_solver.pose_solve(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"], g["extent"],
                                       h_in, g["region_idx"], g["anchors"], want_hyp=True, num_hyp=32, sample_size=10, pipeline="split")
    # host-buffer plugin call: gated pull (pinned) and full copy (pageable)
    b = synth.make_batch(6, H=32, seed=12)
    for pinned in (True, False):
        t = {k: (None if v is None else (torch.from_numpy(np.ascontiguousarray(v)).pin_memory() if pinned else
                                         torch.from_numpy(np.ascontiguousarray(v)))) for k, v in b.items()}
        cx, cy, cz = [t["coor"][:, c].contiguous() for c in range(3)]
        if pinned:
            cx, cy, cz = cx.pin_memory(), cy.pin_memory(), cz.pin_memory()
        hs = pose_solver.HostPoseSolver(inlier_thr=0.005, chunk_rois=4, count_bytes=True)
        hargs = (t["depth"], t["Kp"], cx, cy, cz, t["mask"], t["extent"], t["hyp_idx"], t["region_idx"], t["anchors"])
        hs(*hargs)
        plans = [hs.plan(*hargs, private_outputs=True) for _ in range(2)]  # two submitted calls in flight
        tickets = [p.submit() for p in plans]
        for p, tk in zip(plans, tickets):
            p.wait(tk)
        hs.close()
    # correspondence features / region targets
    geometry.coor_feat(torch.rand(2, 1, 64, 64, device="cuda"), torch.rand(2, 1, 64, 64, device="cuda"), torch.rand(2, 1, 64, 64, device="cuda"),
                       torch.randn(2, 5, 64, 64, device="cuda"), torch.randn(2, 33, 64, 64, device="cuda"), torch.randn(2, 32, 3, device="cuda"),
                       torch.randn(2, 1, 64, 64, device="cuda"))
    geometry.xyz_to_region(torch.randn(2, 64, 64, 3, device="cuda"), torch.randn(2, 32, 3, device="cuda"))
    # geometry
    geometry.kabsch(torch.randn(3, 100, 3, device="cuda"), torch.randn(3, 100, 3, device="cuda"))
    geometry.region_argmax(torch.randn(2, 33, 64, 64, device="cuda"))
    geometry.backproject_th(torch.rand(2, 30, 40, device="cuda"), torch.eye(3, device="cuda"))
    geometry.roi_intrinsics(torch.eye(3, device="cuda")[None].repeat(4, 1, 1), torch.rand(4, 2, device="cuda") * 100, torch.rand(4, device="cuda") * 100 + 20)
    geometry.roi_crop_depth(torch.rand(1, 120, 160, device="cuda"), torch.rand(4, 2, device="cuda") * 100, torch.rand(4, device="cuda") * 100 + 20)
    pose_from_pred.pose_from_pred_centroid_z(torch.randn(4, 6, device="cuda"), torch.rand(4, 2, device="cuda"), torch.rand(4, 1, device="cuda") + 0.5,
                                             torch.eye(3, device="cuda")[None].repeat(4, 1, 1) * 500, torch.rand(4, 2, device="cuda") * 100,
                                             torch.rand(4, device="cuda") + 0.2, torch.rand(4, 2, device="cuda") * 50 + 10)
    pose_from_pred.pose_from_pred(torch.randn(4, 4, device="cuda"), torch.rand(4, 3, device="cuda") + 0.3)
    pose_from_pred.pose_from_pred_centroid_z_abs(torch.randn(4, 6, device="cuda"), torch.rand(4, 2, device="cuda") * 300, torch.rand(4, 1, device="cuda") + 0.5,
                                                 torch.eye(3, device="cuda")[None].repeat(4, 1, 1) * 500)
    geometry.backproject_v2(torch.rand(30, 40, device="cuda"), np.array([[500.0, 0, 20], [0, 500, 15], [0, 0, 1]]))
    geometry.adi(np.eye(3), np.zeros(3), np.eye(3), np.ones(3) * 0.01, torch.randn(600, 3, device="cuda"))
    geometry.add(np.eye(3), np.zeros(3), np.eye(3), np.ones(3) * 0.01, torch.randn(600, 3, device="cuda"))
    torch.cuda.synchronize()
    print("sanitize driver done")


if __name__ == "__main__":
    main()
