#!/usr/bin/env python
"""FPS at the sizes the reference's tools run (tools/lm/1_compute_fps.py:26-35: model meshes of 10^4-10^5 vertices, <= 256
picks, many objects): marginal microseconds per pick of the cluster path vs the cooperative grid, and the batched launch.
One JSON object per line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdpn6d_b200 import fps_utils, synth  # noqa: E402


def ev(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for n in (5_000, 8_192, 20_000, 32_768, 50_000, 65_536):
    t = torch.from_numpy(synth.fps_cloud(n, seed=1)).cuda()
    rec = {"bench": "fps_small", "n": n}
    for name, env, xchg in (("cluster", None, None), ("cluster_barrier", None, "barrier"), ("cooperative", "1", None)):
        os.environ.pop("RDPN_FPS_NO_CLUSTER", None)
        os.environ.pop("RDPN_FPS_EXCHANGE", None)
        if env:
            os.environ["RDPN_FPS_NO_CLUSTER"] = env
        if xchg:
            os.environ["RDPN_FPS_EXCHANGE"] = xchg  # the cluster-barrier exchange the push exchange replaced
        a, b = ev(lambda: fps_utils.fps_indices(t, 64)), ev(lambda: fps_utils.fps_indices(t, 256))
        rec[name + "_us_per_pick"] = 1e3 * (b - a) / 192  # marginal: launch and allocation cancel
        rec[name + "_ms_256"] = b
    os.environ.pop("RDPN_FPS_NO_CLUSTER", None)
    os.environ.pop("RDPN_FPS_EXCHANGE", None)
    print(json.dumps(rec), flush=True)
# 30 objects of 30 000 points, 256 picks: one launch vs object by object
clouds = [torch.from_numpy(synth.fps_cloud(30_000, seed=s)).cuda() for s in range(30)]
one = ev(lambda: fps_utils.fps_indices_batch(clouds, 256), 5)
loop = ev(lambda: [fps_utils.fps_indices(c, 256) for c in clouds], 5)
print(json.dumps({"bench": "fps_batch", "objects": 30, "n": 30_000, "k": 256, "batched_ms": one, "object_by_object_ms": loop,
                  "speedup": loop / one}), flush=True)
