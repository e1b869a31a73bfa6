"""ctypes binding of librdpn6d_b200.so (the C ABI declared in include/rdpn6d_b200.h).

The library is the only compute path of this package: if it is missing and cannot be built, or a
call returns an error, a RuntimeError is raised.  There is no CPU or PyTorch fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librdpn6d_b200.so")

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p


class RoiInputs(ctypes.Structure):
    """struct rdpn_roi_inputs"""
    _fields_ = [
        ("depth", c_vp), ("Kp", c_vp), ("depth_div", c_vp), ("coor_x", c_vp), ("coor_y", c_vp),
        ("coor_z", c_vp), ("mask", c_vp), ("extent", c_vp), ("region_idx", c_vp), ("anchors", c_vp),
        ("num_regions", ctypes.c_int32), ("mask_mode", ctypes.c_int32), ("mask_thr", ctypes.c_float),
        ("B", ctypes.c_int32),
    ]


class SolveParams(ctypes.Structure):
    """struct rdpn_solve_params"""
    _fields_ = [
        ("inlier_thr", ctypes.c_float), ("num_hyp", ctypes.c_int32), ("min_pts", ctypes.c_int32),
        ("min_inliers", ctypes.c_int32), ("weighted", ctypes.c_int32), ("refit_iters", ctypes.c_int32),
        ("with_scale", ctypes.c_int32), ("adaptive", ctypes.c_int32), ("confidence", ctypes.c_float),
        ("min_iter", ctypes.c_int32), ("seed", ctypes.c_uint32), ("roi_base", ctypes.c_int32),
        ("sample_size", ctypes.c_int32), ("pipeline", ctypes.c_int32), ("chunk_rois", ctypes.c_int32),
        ("select_rule", ctypes.c_int32),
    ]


class SolveOutputs(ctypes.Structure):
    """struct rdpn_solve_outputs"""
    _fields_ = [
        ("pose", c_vp), ("n_inliers", c_vp), ("status", c_vp), ("best_h", c_vp), ("n_sel", c_vp),
        ("inlier_mask", c_vp), ("hyp_counts", c_vp), ("hyp_poses", c_vp), ("scale", c_vp), ("rows16", c_vp),
    ]


# every symbol include/rdpn6d_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "rdpn_version": (ctypes.c_int, []),
    "rdpn_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "farthest_point_sampling": (None, [c_vp, c_vp, ctypes.c_int, ctypes.c_int]),
    "farthest_point_sampling_init_center": (None, [c_vp, c_vp, ctypes.c_int, ctypes.c_int]),
    "rdpn_fps_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "rdpn_fps_init_center": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_size_t, c_vp]),
    "rdpn_fps_from_index": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_size_t, c_vp]),
    "rdpn_fps_gather": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp]),
    "rdpn_roi_intrinsics": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp]),
    "rdpn_backproject": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]),
    "rdpn_correspond": (ctypes.c_int, [ctypes.POINTER(RoiInputs), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "rdpn_pose_solve": (ctypes.c_int, [ctypes.POINTER(RoiInputs), c_vp, c_vp, ctypes.POINTER(SolveParams),
                                       ctypes.POINTER(SolveOutputs), c_vp]),
    "rdpn_pose_solve_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "rdpn_pose_solve_ws": (ctypes.c_int, [ctypes.POINTER(RoiInputs), c_vp, c_vp, ctypes.POINTER(SolveParams),
                                          ctypes.POINTER(SolveOutputs), c_vp, ctypes.c_size_t, c_vp]),
    "rdpn_backproject_kinv": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "rdpn_adi": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]),
    "rdpn_add": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp]),
    "rdpn_assemble_pose": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp]),
    "rdpn_fps_batch": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp]),
    "rdpn_pose_solve_stage_ms": (ctypes.c_int, [ctypes.POINTER(RoiInputs), c_vp, c_vp, ctypes.POINTER(SolveParams),
                                                ctypes.POINTER(SolveOutputs), c_vp, ctypes.c_size_t, c_vp,
                                                ctypes.POINTER(ctypes.c_float)]),
    "rdpn_kabsch": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp]),
    "rdpn_centroid_z_to_pose": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int,
                                               ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp]),
    "rdpn_region_argmax": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp]),
    "rdpn_coor_feat": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, c_vp, ctypes.c_int, c_vp]),
    "rdpn_xyz_to_region": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp]),
    "rdpn_roi_crop_depth": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_int,
                                           ctypes.c_int, c_vp, ctypes.c_int, c_vp]),
    "rdpn_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_vp)]),
    "rdpn_ctx_destroy": (None, [c_vp]),
    "rdpn_ctx_set_option": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int]),
    "rdpn_ctx_last_h2d_bytes": (ctypes.c_ulonglong, [c_vp]),
    "rdpn_ctx_last_transfer": (ctypes.c_int, [c_vp]),
    "rdpn_pose_solve_host": (ctypes.c_int, [c_vp, ctypes.POINTER(RoiInputs), c_vp, c_vp, ctypes.POINTER(SolveParams),
                                            ctypes.POINTER(SolveOutputs)]),
    "rdpn_pose_solve_host_submit": (ctypes.c_int, [c_vp, ctypes.POINTER(RoiInputs), c_vp, c_vp, ctypes.POINTER(SolveParams),
                                                   ctypes.POINTER(SolveOutputs), ctypes.POINTER(ctypes.c_int)]),
    "rdpn_ctx_wait": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "rdpn_launch_count": (ctypes.c_ulonglong, []),
    "rdpn_fp32_peak_probe": (ctypes.c_int, [ctypes.c_int, c_f64p]),
}

MAX_SAMPLE = 16  # RDPN_MAX_SAMPLE
PIPELINE_AUTO, PIPELINE_FUSED, PIPELINE_SPLIT = 0, 1, 2  # RDPN_PIPELINE_*
SELECT_MOST_INLIERS, SELECT_MIN_MEAN_ERR = 0, 1  # RDPN_SELECT_*

# rdpn_ctx_set_option keys / transfer strategies (include/rdpn6d_b200.h)
TRANSFER_AUTO, TRANSFER_COPY, TRANSFER_PULL = 0, 1, 2
OPT_TRANSFER, OPT_PULL_GRANULARITY, OPT_CHUNK_ROIS, OPT_COUNT_BYTES = 1, 2, 3, 4

_lib = None


def lib():
    """Load (building first if the .so is absent or stale) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    try:
        _build.build()
    except Exception as e:
        # No silent stale binary: when the sources are newer than the shipped .so and the rebuild fails, that is an
        # error (a GPU box has nvcc; a box without it must receive a current .so).  RDPN_ALLOW_STALE_LIB=1 opts in to
        # loading the stale library anyway (debugging only).
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "rdpn6d_b200: librdpn6d_b200.so is missing and could not be built (%s). "
                "This package has no CPU fallback." % e)
        if os.environ.get("RDPN_ALLOW_STALE_LIB") != "1":
            raise RuntimeError(
                "rdpn6d_b200: csrc/ or include/ is newer than librdpn6d_b200.so and the rebuild failed (%s). "
                "Rebuild with `python -m rdpn6d_b200.build`, or set RDPN_ALLOW_STALE_LIB=1 to load the stale library." % e)
        import warnings

        warnings.warn("rdpn6d_b200: loading a STALE librdpn6d_b200.so (RDPN_ALLOW_STALE_LIB=1): %s" % e)
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            raise RuntimeError("rdpn6d_b200: %s does not export %s" % (LIB_PATH, name))
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().rdpn_error_string(rc).decode()
        raise RuntimeError("rdpn6d_b200 %s failed: %s (code %d)" % (what, msg, rc))


def launch_count():
    return int(lib().rdpn_launch_count())
