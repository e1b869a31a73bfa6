"""GPU-only `pose_from_pred_centroid_z` (surface B3 of SURVEY.md 8b).

Same signature and argument meaning as
/root/reference/core/gdrn_modeling/models/pose_from_pred_centroid_z.py:11-49.  The reference's test
branch (:52-141) copies every ROI to the host and runs allocentric_to_egocentric in numpy (one
device->host sync per ROI, core/utils/utils.py:39-94); here the whole batch is one kernel and `rot`
stays on the device (documented deviation: the reference returns a CPU tensor, :141).
The differentiable train branch (:144-227) is out of scope for this path: is_train=True raises.
"""
import torch

from . import _lib


def pose_from_pred_centroid_z(pred_rots, pred_centroids, pred_z_vals, roi_cams, roi_centers, resize_ratios, roi_whs,
                              eps=1e-4, is_allo=True, z_type="REL", is_train=False):
    if is_train:
        raise NotImplementedError("rdpn6d_b200 implements the test-time (non-differentiable) branch only")
    if z_type not in ("REL", "ABS"):
        raise ValueError(f"Unknown z_type: {z_type}")  # pose_from_pred_centroid_z.py:86
    if not pred_rots.is_cuda:
        raise RuntimeError("rdpn6d_b200.pose_from_pred_centroid_z needs CUDA tensors (no CPU fallback)")
    if roi_cams.dim() == 2:
        roi_cams.unsqueeze_(0)  # :65-66 (in-place, as the reference)
    assert roi_cams.dim() == 3, roi_cams.dim()
    dev = pred_rots.device
    B = pred_centroids.shape[0]
    f = lambda x: x.detach().to(torch.float32).contiguous()
    if pred_rots.dim() == 3 and pred_rots.shape[-1] == 3:
        rot_in, is6d = f(pred_rots), 0
    elif pred_rots.dim() == 2 and pred_rots.shape[-1] == 6:
        rot_in, is6d = f(pred_rots), 1  # rot_reps.py:34-49 fused in
    else:
        raise RuntimeError(f"Wrong pred_rot_ dim: {tuple(pred_rots.shape)}")
    K = f(roi_cams.expand(B, 3, 3))
    cen, z, ctr, wh = f(pred_centroids), f(pred_z_vals.reshape(B)), f(roi_centers), f(roi_whs)
    rr = f(resize_ratios.reshape(B)) if resize_ratios is not None else None
    rot = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
    trans = torch.empty(B, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().rdpn_centroid_z_to_pose(
            rot_in.data_ptr(), is6d, cen.data_ptr(), z.data_ptr(), K.data_ptr(), ctr.data_ptr(),
            rr.data_ptr() if rr is not None else None, wh.data_ptr(), int(bool(is_allo)), int(z_type == "REL"),
            rot.data_ptr(), trans.data_ptr(), B, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "centroid_z_to_pose")
    return rot, trans


def _rot_arg(pred_rots):
    f = lambda x: x.detach().to(torch.float32).contiguous()
    if pred_rots.dim() == 3 and pred_rots.shape[-1] == 3:
        return f(pred_rots), 0
    if pred_rots.dim() == 2 and pred_rots.shape[-1] == 6:
        return f(pred_rots), 1
    if pred_rots.dim() == 2 and pred_rots.shape[-1] == 4:
        return f(pred_rots), 2  # quaternion (w, x, y, z), "this allows unnormalized quat" (pose_from_pred.py:37)
    raise RuntimeError(f"Wrong pred_rot_ dim: {tuple(pred_rots.shape)}")


def _assemble(pred_rots, t_or_c, z, K, trans_mode, is_allo):
    if not pred_rots.is_cuda:
        raise RuntimeError("rdpn6d_b200.pose_from_pred needs CUDA tensors (no CPU fallback)")
    dev, B = pred_rots.device, pred_rots.shape[0]
    rot_in, kind = _rot_arg(pred_rots)
    rot = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
    trans = torch.empty(B, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().rdpn_assemble_pose(rot_in.data_ptr(), kind, t_or_c.data_ptr(), z.data_ptr() if z is not None else None,
                                           K.data_ptr() if K is not None else None, trans_mode, int(bool(is_allo)),
                                           rot.data_ptr(), trans.data_ptr(), B, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "assemble_pose")
    return rot, trans


def pose_from_pred(pred_rots, pred_transes, eps=1e-4, is_allo=True, is_train=False):
    """core/gdrn_modeling/models/pose_from_pred.py:14-58 (test branch): rotation (3x3 matrices or quaternions) and
    translation given, allocentric -> egocentric for the whole batch in one kernel.  Returns (rot [N,3,3], translation)."""
    if is_train:
        raise NotImplementedError("rdpn6d_b200 implements the test-time (non-differentiable) branch only")
    t = pred_transes.detach().to(torch.float32).contiguous()
    rot, _ = _assemble(pred_rots, t, None, None, 2, is_allo)
    return rot, pred_transes  # :23 `translation = pred_transes` is returned as it came in


def pose_from_pred_centroid_z_abs(pred_rots, pred_centroids, pred_z_vals, roi_cams, eps=1e-4, is_allo=True, is_train=False):
    """core/gdrn_modeling/models/pose_from_pred_centroid_z_abs.py:11-92 (test branch): absolute 2-D object centre and
    absolute z -> translation (:45-49), allocentric -> egocentric.  Returns (rot [N,3,3], translation [N,3])."""
    if is_train:
        raise NotImplementedError("rdpn6d_b200 implements the test-time (non-differentiable) branch only")
    if roi_cams.dim() == 2:
        roi_cams.unsqueeze_(0)  # :24-25 (in-place, as the reference)
    assert roi_cams.dim() == 3, roi_cams.dim()
    B = pred_centroids.shape[0]
    f = lambda x: x.detach().to(torch.float32).contiguous()
    return _assemble(pred_rots, f(pred_centroids), f(pred_z_vals.reshape(B)), f(roi_cams.expand(B, 3, 3)), 1, is_allo)
