"""oracle/fps.py -- TEST INFRASTRUCTURE ONLY.

ctypes front ends for (a) the reference's own FPS build (oracle/_ref/libfps_ref.so, kind
"reference") and (b) the plain-C restatement (oracle/fps_oracle.c, kind "port").  Same calling
convention as the reference wrapper /root/reference/core/csrc/fps/fps_utils.py:6-21.
"""
import ctypes

import numpy as np

from . import libfps_ref, liboracle

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def _prep(pts, sn):
    pts = np.ascontiguousarray(pts, np.float32)  # fps_utils.py:10
    assert pts.ndim == 2 and pts.shape[1] == 3  # fps_utils.py:8
    idxs = np.zeros([sn], np.int32)  # fps_utils.py:11
    return pts, idxs


def fps_indices_port(pts, sn, start=None):
    """Indices from the C restatement.  start=None -> bbox-centre initialisation."""
    pts, idxs = _prep(pts, sn)
    lib = liboracle()
    if start is None:
        lib.oracle_fps_init_center(pts.ctypes.data_as(_f32p), idxs.ctypes.data_as(_i32p),
                                   ctypes.c_int(pts.shape[0]), ctypes.c_int(sn))
    else:
        lib.oracle_fps_from_index(pts.ctypes.data_as(_f32p), idxs.ctypes.data_as(_i32p),
                                  ctypes.c_int(pts.shape[0]), ctypes.c_int(sn), ctypes.c_int(start))
    return idxs


def fps_indices_reference(pts, sn):
    """Indices from the reference's own compiled C++ (init_center=True entry)."""
    lib = libfps_ref()
    if lib is None:
        raise RuntimeError("oracle/_ref/libfps_ref.so not available (reference not mounted and no prebuilt)")
    pts, idxs = _prep(pts, sn)
    lib.farthest_point_sampling_init_center(pts.ctypes.data_as(_f32p), idxs.ctypes.data_as(_i32p),
                                            ctypes.c_int(pts.shape[0]), ctypes.c_int(sn))
    return idxs


def farthest_point_sampling(pts, sn, init_center=True, kind="port"):
    """pts[idxs] as float32, exactly like fps_utils.py:21."""
    if not init_center:
        raise ValueError("the random-start entry is not deterministic (cpp:93-94); use start=")
    idx = fps_indices_reference(pts, sn) if kind == "reference" else fps_indices_port(pts, sn)
    return np.ascontiguousarray(pts, np.float32)[idx]


def get_fps_and_center(pts, num_fps=8, init_center=True, kind="port"):
    """/root/reference/core/utils/data_utils.py:217-226: samples + per-axis mean appended."""
    avg = [np.average(pts[:, 0]), np.average(pts[:, 1]), np.average(pts[:, 2])]
    fps_pts = farthest_point_sampling(pts, num_fps, init_center=init_center, kind=kind)
    return np.concatenate([fps_pts, np.array([avg])], axis=0)
