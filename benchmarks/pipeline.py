#!/usr/bin/env python
"""Fused kernel vs three-kernel pipeline on the BASELINE workloads (run on the B200 box); one JSON object per line.

    python benchmarks/pipeline.py [--configs ycbv,lmo] [--chunks 0,2048,8192]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdpn6d_b200 import pose_solver, synth  # noqa: E402

CONFIGS = {
    # name: (B, H, R, n_models, n_symmetric, K, unique, occlusion)
    "ycbv": (8192, 256, 32, 21, 5, synth.K_YCBV, 84, 0.5),
    "lmo": (1024, 256, 64, 8, 0, synth.K_LM, 128, 0.6),
    "ycbv2k": (2048, 256, 32, 21, 5, synth.K_YCBV, 84, 0.5),
    "ycbv4k": (4096, 256, 32, 21, 5, synth.K_YCBV, 84, 0.5),
}


def ev_time(fn, iters, warm=3, streams=None):
    """CUDA-event time per call; with `streams`, call i goes to streams[i % len] (steady state of a serving loop that
    keeps two calls in flight) and the timed region is bracketed on the current stream by fork / join events."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    if streams:
        for s in streams:
            s.wait_stream(cur)
    for i in range(iters):
        fn(i)
    if streams:
        for s in streams:
            cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def sets_for(name, nsets):
    B, H, R, nm, nsym, K, uniq, occ = CONFIGS[name]
    models = synth.make_models(nm, R, seed=7, n_symmetric=nsym)
    base = synth.make_batch(uniq, models=models, H=H, seed=777, K=K, occlusion_max=occ)
    b = synth.tile_batch(base, B)
    out = []
    for i in range(nsets):
        out.append({k: (None if v is None else torch.from_numpy(np.roll(v, 37 * i, axis=0).copy()).cuda()) for k, v in b.items()})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="ycbv,lmo")
    ap.add_argument("--chunks", default="0")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--weighted", type=int, default=1)
    ap.add_argument("--streams", type=int, default=1, help="calls in flight (round-robin over this many streams)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    for name in a.configs.split(","):
        B, H, R = CONFIGS[name][:3]
        nsets = 2 if B >= 8192 else 4
        sets = sets_for(name, nsets)
        ref = None
        for pipe, chunk in [("fused", 0)] + [("split", int(c)) for c in a.chunks.split(",")]:
            solver = pose_solver.PoseSolver(inlier_thr=0.005, weighted=bool(a.weighted), pipeline=pipe, chunk_rois=chunk)
            plans = [pose_solver.make_plan(solver, s["depth"], s["Kp"], s["coor"][:, 0].contiguous(), s["coor"][:, 1].contiguous(),
                                           s["coor"][:, 2].contiguous(), s["mask"], s["extent"], s["hyp_idx"], s["region_idx"],
                                           s["anchors"]) for s in sets]
            strs = [torch.cuda.Stream() for _ in range(a.streams)] if a.streams > 1 else None
            ms = ev_time(lambda i: plans[i % nsets].launch(strs[i % a.streams] if strs else None), a.iters, streams=strs)
            r = plans[0].launch()
            torch.cuda.synchronize()
            rec = {"bench": "pipeline", "config": name, "B": B, "H": H, "R": R, "pipeline": pipe, "chunk_rois": chunk, "streams": a.streams, "ms": ms,
                   "rois_per_s": B / (ms * 1e-3), "solved": float((r.status == 0).float().mean())}
            if ref is None:
                ref = (r.best_h.clone(), r.n_inliers.clone(), r.pose.clone())
            else:
                rec["same_winner"] = bool(torch.equal(ref[0], r.best_h) and torch.equal(ref[1], r.n_inliers))
                rec["max_pose_diff"] = float((ref[2] - r.pose).abs().max())
            print(json.dumps(rec), flush=True)
            del plans


if __name__ == "__main__":
    main()
