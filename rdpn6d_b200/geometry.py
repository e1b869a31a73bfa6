"""Batched GPU versions of the pysixd geometry helpers the path uses (surface B4 of SURVEY.md 8b).

reference function (under /root/reference)                     here
-------------------------------------------------------------  ------------------------------
lib/pysixd/misc.py:334-349   backproject_th(depth, K)           backproject_th (also batched)
lib/pysixd/transform.py:983  superimposition_matrix(v0, v1)     superimposition_matrix, kabsch
core/utils/data_utils.py:111-152 + data_loader.py:553-568      roi_intrinsics
core/gdrn_modeling/models/GDRN.py:206-209                      region_argmax
All compute happens in csrc/*.cu through the C ABI; CUDA tensors only.
"""
import torch

from . import _lib


def _cuda_f32(x, name):
    if not x.is_cuda:
        raise RuntimeError("rdpn6d_b200.geometry: %s must be a CUDA tensor (no CPU fallback)" % name)
    return x.detach().to(torch.float32).contiguous()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def backproject_th(depth, K):
    """misc.py:334-349: depth [H,W] (or [B,H,W]), K [3,3] (or [B,3,3]) -> [H,W,3] (or [B,H,W,3])."""
    single = depth.dim() == 2
    assert depth.dim() in (2, 3), depth.dim()  # misc.py:343
    d = _cuda_f32(depth, "depth")
    d3 = d[None] if single else d
    B, H, W = d3.shape
    Kc = _cuda_f32(K, "K")
    k_stride = 0 if Kc.dim() == 2 else 9
    if Kc.dim() == 3:
        assert Kc.shape[0] == B
    out = torch.empty(B, H, W, 3, dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        rc = _lib.lib().rdpn_backproject(d3.data_ptr(), Kc.data_ptr(), k_stride, out.data_ptr(), B, H, W, _stream(d.device))
    _lib.check(rc, "backproject")
    return out[0] if single else out


def kabsch(src, dst, w=None, scale=False):
    """Batched weighted Kabsch / Umeyama.  src, dst: [B,N,3]; w: [B,N] or None.

    Returns (M [B,3,4] with dst ~ M[:, :, :3] src + M[:, :, 3], scale [B]).  Same solution as
    transform.affine_matrix_from_points(src.T, dst.T, shear=False, scale=scale, usesvd=True).
    """
    s = _cuda_f32(src, "src")
    d = _cuda_f32(dst, "dst")
    if s.dim() != 3 or s.shape[2] != 3 or s.shape != d.shape or s.shape[1] < 3:
        raise ValueError("input arrays are of wrong shape or type")  # transform.py:917-918
    B, N, _ = s.shape
    ww = _cuda_f32(w, "w") if w is not None else None
    if ww is not None:
        assert ww.shape == (B, N)
    M = torch.empty(B, 3, 4, dtype=torch.float32, device=s.device)
    sc = torch.empty(B, dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        rc = _lib.lib().rdpn_kabsch(s.data_ptr(), d.data_ptr(), ww.data_ptr() if ww is not None else None, N,
                                    int(bool(scale)), M.data_ptr(), sc.data_ptr(), B, _stream(s.device))
    _lib.check(rc, "kabsch")
    return M, sc


def superimposition_matrix(v0, v1, scale=False):
    """transform.py:983-1029 for one pair of [3,n] (or [4,n]) CUDA tensors -> 4x4 float32."""
    a = v0[:3].t()[None]
    c = v1[:3].t()[None]
    M, _ = kabsch(a, c, None, scale)
    out = torch.eye(4, dtype=torch.float32, device=M.device)
    out[:3, :4] = M[0]
    return out


def roi_intrinsics(K, center, scale, crop_res=256):
    """K [B,3,3], bbox center [B,2], scale [B] -> Kp [B,4] = (fx', fy', cx', cy') of the crop."""
    Kc = _cuda_f32(K, "K")
    B = Kc.shape[0]
    c = _cuda_f32(center, "center")
    s = _cuda_f32(scale.reshape(B), "scale")
    out = torch.empty(B, 4, dtype=torch.float32, device=Kc.device)
    with torch.cuda.device(Kc.device):
        rc = _lib.lib().rdpn_roi_intrinsics(Kc.data_ptr(), c.data_ptr(), s.data_ptr(), int(crop_res), out.data_ptr(), B,
                                            _stream(Kc.device))
    _lib.check(rc, "roi_intrinsics")
    return out


def roi_crop_depth(depth_imgs, center, scale, img_idx=None, crop_res=256, out_res=64):
    """Full-frame depth [N,H,W] (or [H,W]) -> ROI depth maps [B,64,64]: the pixels (4i,4j) of
    cv2.warpAffine(depth, A, (256,256), INTER_LINEAR) (data_loader.py:532-535, 625) without the 256x256 crop."""
    d = _cuda_f32(depth_imgs, "depth_imgs")
    if d.dim() == 2:
        d = d[None]
    N, H, W = d.shape
    c = _cuda_f32(center, "center")
    B = c.shape[0]
    s = _cuda_f32(scale.reshape(B), "scale")
    idx = None
    if img_idx is not None:
        idx = img_idx.detach().to(torch.int32).contiguous()
        assert idx.shape == (B,) and idx.is_cuda
    out = torch.empty(B, out_res, out_res, dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        rc = _lib.lib().rdpn_roi_crop_depth(d.data_ptr(), H, W, idx.data_ptr() if idx is not None else None, c.data_ptr(),
                                            s.data_ptr(), int(crop_res), int(out_res), out.data_ptr(), B, _stream(d.device))
    _lib.check(rc, "roi_crop_depth")
    return out


def coor_feat(coor_x, coor_y, coor_z, roi_coord_2d, region, fps, mask=None, mask_mode="l1", region_attention=True,
              mask_attention="mul"):
    """The tensor ConvPnPNet's conv stack consumes, in one fused pass (GDRN.py:199-222, conv_pnp_net.py:128-136):
    cat(coor_x, coor_y, coor_z, roi_coord_2d, fps[argmax softmax(region[:,1:])], softmax(region[:,1:]))
    multiplied by (or concatenated with) get_mask_prob(mask).  Returns [B, C, 64, 64]."""
    B = coor_x.shape[0]
    f = lambda x, shp, n: _cuda_f32(x, n).reshape(shp)
    cx, cy, cz = f(coor_x, (B, 4096), "coor_x"), f(coor_y, (B, 4096), "coor_y"), f(coor_z, (B, 4096), "coor_z")
    c2d = f(roi_coord_2d, (B, 5, 4096), "roi_coord_2d")
    reg = _cuda_f32(region, "region")
    R = reg.shape[1] - 1
    reg = reg.reshape(B, R + 1, 4096)
    anc = _cuda_f32(fps, "fps")
    if anc.dim() == 2:
        anc = anc[None].expand(B, -1, -1).contiguous()
    assert anc.shape == (B, R, 3), tuple(anc.shape)
    ma = {"none": 0, "mul": 1, "concat": 2}[mask_attention]
    mm = {"raw": 0, "l1": 1, "bce": 2}[mask_mode.lower()] if isinstance(mask_mode, str) else int(mask_mode)
    mk = f(mask, (B, 4096), "mask") if ma else None
    C = 11 + (R if region_attention else 0) + (1 if ma == 2 else 0)
    out = torch.empty(B, C, 64, 64, dtype=torch.float32, device=cx.device)
    with torch.cuda.device(cx.device):
        rc = _lib.lib().rdpn_coor_feat(cx.data_ptr(), cy.data_ptr(), cz.data_ptr(), c2d.data_ptr(), reg.data_ptr(), anc.data_ptr(),
                                       mk.data_ptr() if mk is not None else None, R, mm, int(bool(region_attention)), ma,
                                       out.data_ptr(), B, _stream(cx.device))
    _lib.check(rc, "coor_feat")
    return out


def xyz_to_region(xyz_crop, fps_points):
    """data_utils.py:229-244 batched: xyz_crop [B,h,w,3] (or [h,w,3]), fps_points [B,R,3] (or [R,3]) ->
    (region ids uint8 [B,h,w] with 0 = background, delta float32 [B,h,w,3])."""
    x = _cuda_f32(xyz_crop, "xyz_crop")
    single = x.dim() == 3
    if single:
        x = x[None]
    B, h, w, _ = x.shape
    fp = _cuda_f32(fps_points, "fps_points")
    if fp.dim() == 2:
        fp = fp[None].expand(B, -1, -1).contiguous()
    R = fp.shape[1]
    region = torch.empty(B, h, w, dtype=torch.uint8, device=x.device)
    delta = torch.empty(B, h, w, 3, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().rdpn_xyz_to_region(x.data_ptr(), fp.data_ptr(), R, h * w, region.data_ptr(), delta.data_ptr(), B,
                                           _stream(x.device))
    _lib.check(rc, "xyz_to_region")
    return (region[0], delta[0]) if single else (region, delta)


def region_argmax(region):
    """GDRN.py:206-209: region [B,R+1,64,64] logits -> uint8 [B,64,64] index in [0,R) (bg channel 0 skipped)."""
    r = _cuda_f32(region, "region")
    B, R1, h, w = r.shape
    assert (h, w) == (64, 64)
    out = torch.empty(B, 64, 64, dtype=torch.uint8, device=r.device)
    with torch.cuda.device(r.device):
        rc = _lib.lib().rdpn_region_argmax(r.data_ptr(), R1 - 1, out.data_ptr(), B, _stream(r.device))
    _lib.check(rc, "region_argmax")
    return out


def _kinv_mats(K, R=None, T=None, dev=None):
    import numpy as np

    Kn = K.detach().cpu().numpy() if torch.is_tensor(K) else np.asarray(K)
    m = np.concatenate([np.linalg.inv(Kn.astype(np.float64)).reshape(9),  # misc.py:294 / :360 (a 3 x 3 host inverse)
                        (np.eye(3) if R is None else np.asarray(R.detach().cpu() if torch.is_tensor(R) else R, np.float64)).reshape(9),
                        (np.zeros(3) if T is None else np.asarray(T.detach().cpu() if torch.is_tensor(T) else T, np.float64)).reshape(3)])
    return torch.from_numpy(m).to(dev)


def backproject_v2(depth, K):
    """lib/pysixd/misc.py:352-371: organised cloud d * Kinv (u, v, 1)^T, [H,W,3] float64 (CUDA in -> CUDA out)."""
    d = _cuda_f32(depth, "depth")
    H, W = d.shape
    out = torch.empty(H, W, 3, dtype=torch.float64, device=d.device)
    mats = _kinv_mats(K, dev=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().rdpn_backproject_kinv(d.data_ptr(), mats.data_ptr(), H, W, out.data_ptr(), _stream(d.device)), "backproject_kinv")
    return out


def calc_emb_bp_fast(depth, R, T, K):
    """lib/pysixd/misc.py:288-316 (= calc_xyz_bp_fast): object coordinates of a rendered depth map,
    (depth != 0) * R^T (d * Kinv (u, v, 1)^T - T), [H,W,3] float64."""
    d = _cuda_f32(depth, "depth")
    H, W = d.shape
    out = torch.empty(H, W, 3, dtype=torch.float64, device=d.device)
    mats = _kinv_mats(K, R, T, dev=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().rdpn_backproject_kinv(d.data_ptr(), mats.data_ptr(), H, W, out.data_ptr(), _stream(d.device)), "backproject_kinv")
    return out


calc_xyz_bp_fast = calc_emb_bp_fast  # misc.py:319


def _model_distance(entry, what, R_est, t_est, R_gt, t_gt, pts):
    import numpy as np

    p = _cuda_f32(pts, "pts")
    n = p.shape[0]
    as_np = lambda x: (x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)).astype(np.float64)
    poses = torch.from_numpy(np.concatenate([as_np(R_est).reshape(9), as_np(t_est).reshape(3), as_np(R_gt).reshape(9),
                                             as_np(t_gt).reshape(3)])).to(p.device)
    scratch = torch.zeros(1 + (n + 255) // 256, dtype=torch.float64, device=p.device)
    out = torch.empty(1, dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(getattr(_lib.lib(), entry)(p.data_ptr(), n, poses.data_ptr(), scratch.data_ptr(), out.data_ptr(), _stream(p.device)), what)
    return float(out.item())


def adi(R_est, t_est, R_gt, t_gt, pts):
    """lib/pysixd/pose_error.py:315-337: Average Distance of model points for objects with Indistinguishable views (the
    symmetric objects of the YCB-V config): mean distance from every ground-truth-posed model point to the nearest
    estimate-posed one.  pts: [n,3] CUDA tensor; poses: anything array-like.  Returns a Python float (as the reference)."""
    return _model_distance("rdpn_adi", "adi", R_est, t_est, R_gt, t_gt, pts)


def add(R_est, t_est, R_gt, t_gt, pts):
    """lib/pysixd/pose_error.py:297-312: Average Distance of model points (objects without indistinguishable views): mean
    distance between the same model point under the estimated and the ground-truth pose.  Arguments as adi."""
    return _model_distance("rdpn_add", "add", R_est, t_est, R_gt, t_gt, pts)


def re(R_est, R_gt):
    """lib/pysixd/pose_error.py:400-415: rotational error in degrees, acos of the clipped (trace - 1) / 2.  3 x 3 host-side
    helper of the evaluation (numpy in -> float out, as the reference)."""
    import numpy as np

    R_est, R_gt = np.asarray(R_est), np.asarray(R_gt)
    assert R_est.shape == R_gt.shape == (3, 3)
    tr = min(float(np.einsum("ij,ij->", R_est, R_gt)), 3.0)  # trace(R_est R_gt^T), capped like the reference
    return float(np.rad2deg(np.arccos(min(1.0, max(-1.0, 0.5 * (tr - 1.0))))))


def te(t_est, t_gt):
    """lib/pysixd/pose_error.py:425-436: translational error, the L2 norm of t_gt - t_est."""
    import numpy as np

    t_est, t_gt = np.asarray(t_est).flatten(), np.asarray(t_gt).flatten()
    assert t_est.size == t_gt.size == 3
    return np.linalg.norm(t_gt - t_est)


def get_closest_rot(rot_est, rot_gt, sym_info):
    """core/utils/pose_utils.py:430-454: among rot_gt and rot_gt @ S for the object's symmetry transforms S ([K,3,3] or
    [3,3] or None), the one with the smallest rotation error to rot_est (strictly smaller replaces: the plain rot_gt wins
    ties).  A 3 x 3 host-side helper of the evaluation, numpy in -> numpy out, as the reference."""
    import numpy as np

    if sym_info is None:
        return rot_gt
    if torch.is_tensor(sym_info):
        sym_info = sym_info.cpu().numpy()
    sym_info = np.asarray(sym_info)
    if sym_info.ndim == 2:
        sym_info = sym_info.reshape((1, 3, 3))

    def re(a, b):  # pose_error.py:400-415
        tr = np.trace(np.asarray(a, np.float64).dot(np.asarray(b, np.float64).T))
        tr = tr if tr <= 3 else 3
        return np.rad2deg(np.arccos(0.5 * (tr - 1.0)))

    r_err, closest = re(rot_est, rot_gt), rot_gt
    for i in range(sym_info.shape[0]):
        cand = rot_gt.dot(sym_info[i])
        cur = re(rot_est, cand)
        if cur < r_err:
            r_err, closest = cur, cand
    return closest


def _axis_rotations(axis, angles):
    """Rodrigues rotations about `axis` (through the origin) for a vector of angles: [n,3,3] float64."""
    import numpy as np

    u = np.asarray(axis, np.float64)[:3]
    u = u / np.sqrt(u @ u)
    K = np.array([[0.0, -u[2], u[1]], [u[2], 0.0, -u[0]], [-u[1], u[0], 0.0]])
    c, s_ = np.cos(angles)[:, None, None], np.sin(angles)[:, None, None]
    return c * np.eye(3) + (1.0 - c) * np.outer(u, u) + s_ * K


def get_symmetry_transformations(model_info, max_sym_disc_step):
    """lib/pysixd/misc.py:206-254: the symmetry set of an object model (models_info.json) as a list of {"R": [3,3],
    "t": [3,1]} -- every discrete symmetry (identity first) composed with every step of every continuous one, the
    continuous ones discretised into ceil(pi / max_sym_disc_step) steps of a full turn about (axis, offset).  Dataset-
    metadata host logic; np.stack of the "R" entries is the `sym_info` get_closest_rot takes."""
    import numpy as np

    discrete = [np.eye(4)] + [np.reshape(m, (4, 4)).astype(np.float64) for m in model_info.get("symmetries_discrete", [])]
    steps = int(np.ceil(np.pi / max_sym_disc_step))
    continuous = []  # (R, t) of the rotation about the offset point: x -> R (x - o) + o
    for sym in model_info.get("symmetries_continuous", []):
        o = np.asarray(sym["offset"], np.float64).reshape(3, 1)
        for R in _axis_rotations(sym["axis"], 2.0 * np.pi / steps * np.arange(1, steps)):
            continuous.append((R, o - R @ o))
    out = []
    for D in discrete:
        Rd, td = D[:3, :3], D[:3, 3:4]
        if continuous:
            out.extend({"R": Rc @ Rd, "t": Rc @ td + tc} for Rc, tc in continuous)
        else:
            out.append({"R": Rd, "t": td})
    return out
