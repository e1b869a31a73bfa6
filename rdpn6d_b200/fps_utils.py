"""GPU farthest point sampling behind the reference's Python surface.

Mirrors /root/reference/core/csrc/fps/fps_utils.py:6-21 (`farthest_point_sampling(pts, sn,
init_center=False) -> pts[idxs]`) and /root/reference/core/utils/data_utils.py:217-226
(`get_fps_and_center`).  numpy in -> numpy out (as the reference); a CUDA tensor in -> CUDA tensor out.
All compute is the sm_100a kernel in csrc/fps.cu; there is no CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("rdpn6d_b200.fps_utils needs a CUDA device (no CPU fallback)")


def fps_indices(pts, sn, init_center=True, start=None, stream=None):
    """Indices [sn] int32 (CUDA tensor) of the farthest-point samples of pts [N,3] (CUDA, float32).

    init_center=True  -> farthest_point_sampling_init_center (farthest_point_sampling.cpp:186-204)
    init_center=False -> farthest_point_sampling (:166-184) from `start` (random when None, :93-94)
    """
    _require_cuda()
    L = _lib.lib()
    assert pts.dim() == 2 and pts.shape[1] == 3  # fps_utils.py:8
    pts = pts.contiguous().to(torch.float32)
    pn = pts.shape[0]
    idx = torch.zeros(sn, dtype=torch.int32, device=pts.device)
    if pn == 0 or sn == 0:
        return idx
    ws_bytes = int(L.rdpn_fps_workspace_bytes(sn))
    if pn > 1_000_000:  # beyond the register-resident limit the running minima stream from the workspace (include/rdpn6d_b200.h)
        ws_bytes += 4 * pn
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pts.device)
    st = (stream or torch.cuda.current_stream(pts.device)).cuda_stream
    with torch.cuda.device(pts.device):
        if init_center:
            rc = L.rdpn_fps_init_center(pts.data_ptr(), idx.data_ptr(), pn, sn, ws.data_ptr(), ws_bytes, st)
        else:
            if start is None:
                start = int(np.random.randint(pn))
            rc = L.rdpn_fps_from_index(pts.data_ptr(), idx.data_ptr(), pn, sn, int(start), ws.data_ptr(), ws_bytes, st)
    _lib.check(rc, "fps")
    return idx


def farthest_point_sampling(pts, sn, init_center=False):
    """Drop-in for fps_utils.farthest_point_sampling: returns pts[idxs] as float32 [sn,3]."""
    _require_cuda()
    if isinstance(pts, np.ndarray):
        pn, _ = pts.shape
        assert pts.shape[1] == 3
        pts32 = np.ascontiguousarray(pts, np.float32)  # fps_utils.py:10
        idxs = np.ascontiguousarray(np.zeros([sn], np.int32))  # fps_utils.py:11
        L = _lib.lib()
        fn = L.farthest_point_sampling_init_center if init_center else L.farthest_point_sampling
        fn(pts32.ctypes.data_as(ctypes.c_void_p), idxs.ctypes.data_as(ctypes.c_void_p), pn, sn)  # fps_utils.py:16-19
        return pts32[idxs]
    idx = fps_indices(pts, sn, init_center=init_center)
    return pts.contiguous().to(torch.float32)[idx.long()]


def get_fps_and_center(pts, num_fps=8, init_center=True):
    """data_utils.py:217-226: FPS samples with the per-axis mean appended as the last row."""
    if isinstance(pts, np.ndarray):
        avg = [np.average(pts[:, 0]), np.average(pts[:, 1]), np.average(pts[:, 2])]
        fps_pts = farthest_point_sampling(pts, num_fps, init_center=init_center)
        return np.concatenate([fps_pts, np.array([avg])], axis=0)
    _require_cuda()
    L = _lib.lib()
    p32 = pts.contiguous().to(torch.float32)
    idx = fps_indices(p32, num_fps, init_center=init_center)
    out = torch.empty(num_fps, 3, dtype=torch.float32, device=p32.device)
    center = torch.empty(3, dtype=torch.float64, device=p32.device)
    with torch.cuda.device(p32.device):
        rc = L.rdpn_fps_gather(p32.data_ptr(), idx.data_ptr(), p32.shape[0], num_fps, out.data_ptr(),
                               center.data_ptr(), torch.cuda.current_stream(p32.device).cuda_stream)
    _lib.check(rc, "fps_gather")
    return torch.cat([out.to(torch.float64), center[None]], dim=0)


def fps_indices_batch(clouds, sn, starts=None, stream=None):
    """FPS of many clouds in ONE launch (rdpn_fps_batch: one thread-block cluster per object).  clouds: list of [N_i,3]
    CUDA float32 tensors (N_i <= 65 536); starts: None = init_center for every object, else one first index per
    object.  Returns [len(clouds), sn] int32 CUDA tensor, indices relative to each cloud."""
    _require_cuda()
    L = _lib.lib()
    dev = clouds[0].device
    sizes = [int(c.shape[0]) for c in clouds]
    assert all(c.dim() == 2 and c.shape[1] == 3 for c in clouds) and min(sizes) > 0
    pts = torch.cat([c.contiguous().to(torch.float32) for c in clouds], dim=0)
    offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=dev)
    idx = torch.zeros(len(clouds), sn, dtype=torch.int32, device=dev)
    st_t = None if starts is None else torch.as_tensor(starts, dtype=torch.int32, device=dev).contiguous()
    st = (stream or torch.cuda.current_stream(dev)).cuda_stream
    with torch.cuda.device(dev):
        rc = L.rdpn_fps_batch(pts.data_ptr(), offs.data_ptr(), len(clouds), max(sizes), sn,
                              st_t.data_ptr() if st_t is not None else None, idx.data_ptr(), st)
    _lib.check(rc, "fps_batch")
    return idx


def fps_and_center_for_models(clouds, nums_fps=(2, 4, 8, 12, 16, 20, 32, 64, 128, 256)):
    """The loop of tools/*/..._compute_fps.py (e.g. tools/lm/1_compute_fps.py:26-35): for every object cloud and
    every sample count, `get_fps_and_center(pts, n, init_center=True)`.  clouds: dict obj_id -> [N,3] numpy array.
    Returns {str(obj_id): {"fps{n}_and_center": [n+1,3] array}} ready for mmcv.dump(fps_points.pkl).  Every cloud is
    uploaded once and ALL objects run in one launch (one thread-block cluster per object) when the largest cloud has at
    most 65 536 points, object by object otherwise; FPS prefixes differ per count only in length, so the largest
    count is computed and shorter ones are prefixes of it."""
    _require_cuda()
    out = {}
    kmax = max(nums_fps)
    ids = list(clouds.keys())
    p32 = [np.ascontiguousarray(np.asarray(clouds[i]), np.float32) for i in ids]
    dev_clouds = [torch.from_numpy(p).cuda() for p in p32]
    if max(p.shape[0] for p in p32) <= 65536:
        idx_all = fps_indices_batch(dev_clouds, kmax).cpu().numpy()
    else:
        idx_all = np.stack([fps_indices(c, kmax, init_center=True).cpu().numpy() for c in dev_clouds])
    for j, obj_id in enumerate(ids):
        pts = np.asarray(clouds[obj_id])
        avg = np.array([[np.average(pts[:, 0]), np.average(pts[:, 1]), np.average(pts[:, 2])]])
        out[str(obj_id)] = {f"fps{n}_and_center": np.concatenate([p32[j][idx_all[j][:n]], avg], axis=0) for n in nums_fps}
    return out
