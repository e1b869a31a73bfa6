import os, sys, json, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'benchmarks')
from rdpn6d_b200 import fps_utils, synth
def ev(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for n in (5000, 8192, 20000, 32768):
    t = torch.from_numpy(synth.fps_cloud(n, seed=1)).cuda()
    for ct in (128, 256, 512):
        os.environ["RDPN_FPS_CLUSTER_THREADS"] = str(ct)
        for ppt in (1, 2, 4, 8, 16):
            if 8 * ct * ppt < n:
                continue
            os.environ["RDPN_FPS_CLUSTER_PPT"] = str(ppt)
            a, b = ev(lambda: fps_utils.fps_indices(t, 64)), ev(lambda: fps_utils.fps_indices(t, 256))
            print(n, "threads", ct, "ppt", ppt, "C=%d" % -(-n // (ct * ppt)), "us/pick %.3f" % (1e3 * (b - a) / 192), flush=True)
