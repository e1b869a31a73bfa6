#!/usr/bin/env python
"""Cycle marks inside the FPS cluster kernel (tuning builds only): where the fixed cost of a pick goes.

Build the library with the marks compiled in and run this on a B200:
    RDPN_NVCC_EXTRA=-DRDPN_FPS_TIMING python -m rdpn6d_b200.build --force && python benchmarks/fps_timing.py
Thread 0 of rank 0 sums clock64() differences between the marks of the push exchange (distance update | warp REDUX |
block barrier | wait on the CTA's mbarrier | selection among the <= 8 CTA candidates) over the picks and prints the
averages at the end of the kernel (device printf).  Rebuild without the flag afterwards: the marks cost ~100 cycles a pick."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rdpn6d_b200 import fps_utils, synth  # noqa: E402

for n in (5000, 8192, 20000, 32768):
    t = torch.from_numpy(synth.fps_cloud(n, seed=1)).cuda()
    fps_utils.fps_indices(t, 256)
    torch.cuda.synchronize()
