"""oracle/pose_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by rdpn6d_b200/).

CPU restatement (numpy, float32/float64 made explicit) of the reference's test-time
dense-correspondence -> pose path.  The reference has no single entry point for this path
(SURVEY.md section 0): the oracle is the COMPOSITION of the reference functions cited below, with the
RANSAC randomness lifted out into an explicit hypothesis index tensor.

stage                       reference lines followed (under /root/reference)
--------------------------  -------------------------------------------------------------------------
roi_scalars                 core/gdrn_modeling/data_loader.py:472-488
roi_intrinsics (a1)         core/utils/data_utils.py:111-152 (rot=0 closed form), data_loader.py:553-568
backproject_roi (a2)        core/gdrn_modeling/data_loader.py:530-576, 624-627
backproject (generic)       lib/pysixd/misc.py:319-331, 334-349
region_argmax/anchors (a3)  core/gdrn_modeling/models/GDRN.py:206-218
residual de-normalise (a3)  core/gdrn_modeling/gdrn_evaluator.py:102-105, models/conv_pnp_net.py:125-127
out_mask (a4)               core/gdrn_modeling/engine_utils.py:118-136, models/model_utils.py:24-42
gate (a4)                   core/gdrn_modeling/gdrn_evaluator.py:110-117
ransac semantics (a5)       lib/pysixd/misc.py:58-142 (strict '<' inliers :111, strictly-greater best and
                            >=4 rule :121, refit on inliers :123-126, adaptive stop :134-138)
kabsch / umeyama (a7)       lib/pysixd/transform.py:913-929, 940-951, 971-980
rigid apply / residual (a8) lib/pysixd/misc.py:895-905  (float32 loop in oracle/pose_oracle.c)
pose assembly (a9)          core/gdrn_modeling/models/pose_from_pred_centroid_z.py:52-141,
                            core/utils/utils.py:39-94, core/utils/rot_reps.py:34-49
re / te (a10)               lib/pysixd/pose_error.py:400-415, 425-436

dtype decisions (SURVEY.md section 7 "numpy-version semantics"): the reference is pinned to numpy
1.23.4 where float32-array (op) float64-scalar evaluates in float32 with the scalar rounded to
float32 first.  Every elementwise stage below therefore rounds its scalars to float32 and computes
in float32, one IEEE operation at a time; Kabsch follows transform.py in float64.
"""
import ctypes
import math

import numpy as np

from . import liboracle

F32 = np.float32
F64 = np.float64

STATUS_OK = 0
STATUS_FEW_POINTS = 1  # gdrn_evaluator.py:393-395  (<4 correspondences -> -100 fill)
STATUS_T_SANITY = 2  # gdrn_evaluator.py:293-296  (te(t_est, t_net) > 1 m -> keep net t)
STATUS_NO_CONSENSUS = 3  # misc.py:121: no hypothesis reached the >=4 inlier rule


# --------------------------------------------------------------------------------------------
# a1: ROI scalars and crop-adjusted intrinsics
# --------------------------------------------------------------------------------------------
def roi_scalars(bbox_xyxy, im_H, im_W, dzi_pad_scale=1.5, out_res=64):
    """data_loader.py:472-488 -> (center[2], scale, resize_ratio, wh[2]) in float64."""
    x1, y1, x2, y2 = [float(v) for v in bbox_xyxy]
    center = np.array([0.5 * (x1 + x2), 0.5 * (y1 + y2)])
    bw = max(x2 - x1, 1)
    bh = max(y2 - y1, 1)
    scale = max(bh, bw) * dzi_pad_scale
    scale = min(scale, max(im_H, im_W)) * 1.0
    return center, scale, out_res / scale, np.array([bw, bh], dtype=F64)


def roi_affine(center, scale, crop_res=256):
    """2x3 crop affine of data_utils.get_affine_transform (:111-152) for rot=0, restated in closed form.

    The reference stores its three source points in a float32 array (:136-142) before handing them to
    cv2.getAffineTransform, so its scale factors are crop/2 divided by float32-rounded differences, not
    exactly crop/scale.  This restatement repeats those roundings (float64 arithmetic on float32-rounded
    points) and solves the 3-point system in closed form:
        p0 = f32(center); p1 = f32(center + (0, -0.5*f32(scale))); e = f32(p0.y - p1.y)
        p2 = (f32(p1.x - e), p1.y);  p0 -> (c/2, c/2), p1 -> (c/2, 0), p2 -> (0, 0)
    Checked against golden matrices produced by the reference function (tests/golden/affine_golden.npz);
    cv2's LU solve differs from the closed form by ~1e-16 relative.
    """
    x0 = float(F32(center[0]))
    y0 = float(F32(center[1]))
    h = 0.5 * float(F32(scale))
    y1 = float(F32(float(center[1]) - h))
    e1 = y0 - y1
    e = float(F32(e1))
    x2 = float(F32(x0 - e))
    e2 = x0 - x2
    half = 0.5 * float(crop_res)
    a00 = half / e2
    a11 = half / e1
    return np.array([[a00, 0.0, half - a00 * x0], [0.0, a11, half - a11 * y0]], dtype=F64)


def roi_intrinsics(K, center, scale, crop_res=256):
    """K' = [[A],[0,0,1]] . K (data_loader.py:553-564) -> (fx', fy', cx', cy') = K'[0,0], K'[1,1],
    K'[0,2], K'[1,2] (:565-568), float64.  A = roi_affine(center, scale).  Operation order is part of the
    contract (the CUDA kernel repeats it in float64 from float32 center / scale / K)."""
    K = np.asarray(K, dtype=F64)
    A = roi_affine(center, scale, crop_res)
    fxp = A[0, 0] * K[0, 0]
    fyp = A[1, 1] * K[1, 1]
    cxp = A[0, 0] * K[0, 2] + A[0, 2]
    cyp = A[1, 1] * K[1, 2] + A[1, 2]
    return np.array([fxp, fyp, cxp, cyp], dtype=F64)


# --------------------------------------------------------------------------------------------
# a2: back-projection
# --------------------------------------------------------------------------------------------
def backproject_roi(depth, Kp, depth_div=None, stride=4):
    """data_loader.py:563-576 evaluated at the pixels kept by [:, ::4, ::4] (:625).

    depth: [h,w] float32 ROI depth sampled at crop pixels (stride*i, stride*j); Kp = (fx',fy',cx',cy');
    depth_div = resize_ratio (data_loader.py:563) or None for metric depth.
    pt0 = (xmap - cx') * pt2 / fx'  -- sub, mul, div, each rounded to float32.
    Returns [3,h,w] float32.
    """
    depth = np.asarray(depth, dtype=F32)
    h, w = depth.shape
    fx, fy, cx, cy = [F32(v) for v in Kp]
    d = depth if depth_div is None else (depth / F32(depth_div)).astype(F32)
    u = (np.arange(w, dtype=F32) * F32(stride))[None, :]
    v = (np.arange(h, dtype=F32) * F32(stride))[:, None]
    x = ((u - cx) * d) / fx
    y = ((v - cy) * d) / fy
    return np.stack([x.astype(F32), y.astype(F32), d.astype(F32)], axis=0)


def backproject(depth, K):
    """misc.py:319-331 / 334-349: [H,W] depth, 3x3 K -> [H,W,3]; float32, (X*depth)/fx order."""
    depth = np.asarray(depth, dtype=F32)
    H, W = depth.shape
    K = np.asarray(K)
    X = (np.arange(W, dtype=F32) - F32(K[0, 2]))[None, :]
    Y = (np.arange(H, dtype=F32) - F32(K[1, 2]))[:, None]
    return np.stack(((X * depth) / F32(K[0, 0]), (Y * depth) / F32(K[1, 1]), depth), axis=2).astype(F32)


# --------------------------------------------------------------------------------------------
# a3: region arg-max, anchor gather, residual de-normalisation
# --------------------------------------------------------------------------------------------
def region_argmax(region_logits):
    """GDRN.py:206-209: softmax over channels 1..R then argmax -> index in [0,R).

    softmax is monotone per pixel, so argmax over the raw logits of channels 1..R gives the same
    index (first maximum wins, as torch.argmax).  region_logits: [R+1,h,w] -> uint8 [h,w].
    """
    return np.argmax(np.asarray(region_logits)[1:], axis=0).astype(np.uint8)


def denormalise_residual(coor, extent):
    """gdrn_evaluator.py:102-105: (c - 0.5) * extent_c, float32. coor [3,h,w], extent [3]."""
    coor = np.asarray(coor, dtype=F32)
    ext = np.asarray(extent, dtype=F32).reshape(3, 1, 1)
    return ((coor - F32(0.5)) * ext).astype(F32)


# --------------------------------------------------------------------------------------------
# a4: mask probability and gate
# --------------------------------------------------------------------------------------------
MASK_RAW, MASK_L1, MASK_BCE = 0, 1, 2


def out_mask(mask, mode=MASK_L1):
    """engine_utils.py:118-136. L1: (m-min)/(max-min) without eps (flat mask -> NaN -> nothing passes)."""
    m = np.asarray(mask, dtype=F32)
    if mode == MASK_RAW:
        return m
    if mode == MASK_L1:
        mn, mx = m.min(), m.max()
        with np.errstate(invalid="ignore", divide="ignore"):
            return ((m - mn) / (mx - mn)).astype(F32)
    if mode == MASK_BCE:
        return (F32(1) / (F32(1) + np.exp(-m, dtype=F32))).astype(F32)
    raise NotImplementedError(mode)


def gate_thresholds(extent):
    """'0.0001 * extent[c]' (gdrn_evaluator.py:112-114): python float times float32 scalar is float64
    under numpy 1.23 scalar promotion, then rounded to float32 by the array comparison."""
    return np.array([F32(0.0001 * float(F32(e))) for e in extent], dtype=F32)


def gate(mask_prob, delta, extent, z, mask_thr=0.5):
    """gdrn_evaluator.py:110-117 plus the composite's own depth validity rule (z > 0: a 3D-3D pair
    needs a measured depth; the reference's 2D-3D PnP had no such need)."""
    thr = gate_thresholds(extent)
    with np.errstate(invalid="ignore"):
        sel = mask_prob > F32(mask_thr)
    for c in range(3):
        sel &= np.abs(delta[c]) > thr[c]
    sel &= z > F32(0)
    return sel


def correspondences(depth, Kp, coor, mask, extent, region_idx=None, anchors=None, depth_div=None,
                    mask_mode=MASK_L1, mask_thr=0.5, stride=4):
    """Stage S1 for one ROI.  Returns dict(cam[3,P], obj[3,P], w[P], sel[P]) float32 / bool, P=h*w.

    anchor mode (region_idx, anchors given): obj = anchors[region], cam = q - delta  (q - delta =
    R a + t because delta = R (x_obj - a), data_loader.py:883-887).
    dense mode (no anchors): obj = delta = (coor-0.5)*extent is the object coordinate, cam = q.
    """
    q = backproject_roi(depth, Kp, depth_div, stride)
    delta = denormalise_residual(coor, extent)
    mprob = out_mask(mask, mask_mode)
    sel = gate(mprob, delta, extent, q[2], mask_thr)
    if anchors is not None:
        a = np.asarray(anchors, dtype=F32)[np.asarray(region_idx).astype(np.int64)]  # [h,w,3]
        obj = np.ascontiguousarray(a.transpose(2, 0, 1))
        cam = (q - delta).astype(F32)
    else:
        obj = delta
        cam = q
    P = sel.size
    return dict(cam=cam.reshape(3, P), obj=obj.reshape(3, P), w=mprob.reshape(P), sel=sel.reshape(P))


# --------------------------------------------------------------------------------------------
# a7: Kabsch / Umeyama (float64, SVD) -- transform.py:913-980 restated, with optional weights
# --------------------------------------------------------------------------------------------
def kabsch(v0, v1, w=None, scale=False):
    """4x4 float64 M with M @ [v0;1] ~ [v1;1].  v0, v1: [3,n].  w=None is exactly
    affine_matrix_from_points(v0, v1, shear=False, scale=scale, usesvd=True)."""
    v0 = np.array(v0, dtype=F64, copy=True)
    v1 = np.array(v1, dtype=F64, copy=True)
    if v0.shape[0] != 3 or v0.shape[1] < 3 or v0.shape != v1.shape:
        raise ValueError("input arrays are of wrong shape or type")  # transform.py:917-918
    if w is None:
        w = np.ones(v0.shape[1], dtype=F64)
    w = np.asarray(w, dtype=F64)
    sw = w.sum()
    m0 = (v0 * w).sum(axis=1) / sw  # transform.py:921 (mean; weighted generalisation)
    m1 = (v1 * w).sum(axis=1) / sw  # transform.py:925
    v0 -= m0[:, None]
    v1 -= m1[:, None]
    u, s, vh = np.linalg.svd((v1 * w) @ v0.T)  # transform.py:942
    R = u @ vh  # transform.py:944
    if np.linalg.det(R) < 0.0:  # transform.py:945-948
        R -= np.outer(u[:, 2], vh[2, :] * 2.0)
    c = 1.0
    if scale:  # transform.py:971-975
        c = math.sqrt(((v1 * v1) * w).sum() / ((v0 * v0) * w).sum())
    M = np.identity(4)
    M[:3, :3] = c * R
    M[:3, 3] = m1 - c * (R @ m0)  # transform.py:978: inv(M1) . M . M0
    return M


def superimposition_matrix(v0, v1, scale=False):
    """transform.py:983-1029."""
    return kabsch(np.asarray(v0, dtype=F64)[:3], np.asarray(v1, dtype=F64)[:3], scale=scale)


# --------------------------------------------------------------------------------------------
# hypothesis generation: minimal 3-point Kabsch per index triplet
# --------------------------------------------------------------------------------------------
DEGENERATE_SIN2 = 1e-6  # triangles with sin^2(angle at vertex 0) <= this are rejected


def _triangle_ok(p0, p1, p2):
    """float64, one rounding per operation, fixed order (shared contract with the CUDA kernel)."""
    e1 = p1 - p0
    e2 = p2 - p0
    nx = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
    ny = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
    nz = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    a2 = (nx * nx + ny * ny) + nz * nz
    l1 = (e1[:, 0] * e1[:, 0] + e1[:, 1] * e1[:, 1]) + e1[:, 2] * e1[:, 2]
    l2 = (e2[:, 0] * e2[:, 0] + e2[:, 1] * e2[:, 1]) + e2[:, 2] * e2[:, 2]
    return a2 > DEGENERATE_SIN2 * (l1 * l2)


def hypothesis_poses(obj, cam, sel, hyp_idx):
    """obj, cam: [3,P] float32; sel: [P] bool; hyp_idx: [H,S] absolute pixel indices, S >= 3 pairs per hypothesis
    (3 = the minimal 3D-3D sample; the reference's loop draws random_sample_num = 10, misc.py:72,91).

    A hypothesis is valid iff its S pixels passed the gate, are pairwise distinct (misc.py:91 samples without
    replacement; for S = 3 a repeated pixel is a degenerate triangle anyway) and each side has a non-degenerate
    triangle (p0, p_{v-1}, p_v) (misc.py:95-101 carries the same intent as a commented-out determinant check).  Pose = Kabsch of the S pairs in float64 (numpy SVD, as transform.py), rounded to float32.
    Returns Rt[H,12] float32 (R row-major | t interleaved as 3x4) and valid[H] uint8.
    """
    hyp_idx = np.asarray(hyp_idx, dtype=np.int64)
    H, S = hyp_idx.shape
    P = obj.shape[1]
    inb = ((hyp_idx >= 0) & (hyp_idx < P)).all(axis=1)
    idx = np.clip(hyp_idx, 0, P - 1)
    a = obj.T.astype(F64)[idx]  # [H,S,3] (hyp, vertex, xyz)
    c = cam.T.astype(F64)[idx]
    valid = inb & sel[idx].all(axis=1)
    if S > 3:
        srt = np.sort(hyp_idx, axis=1)
        valid &= (srt[:, 1:] != srt[:, :-1]).all(axis=1)
    # non-degeneracy: some triangle (p0, p_{v-1}, p_v), 2 <= v < S, passes the test -- on the object side and on the
    # camera side (S = 3: the one triangle there is).  In anchor mode several sampled pixels may share an anchor, so
    # a single fixed triangle would reject samples the reference's loop solves without trouble.
    ok_a = np.zeros(H, bool)
    ok_c = np.zeros(H, bool)
    for v in range(2, S):
        ok_a |= _triangle_ok(a[:, 0], a[:, v - 1], a[:, v])
        ok_c |= _triangle_ok(c[:, 0], c[:, v - 1], c[:, v])
    valid &= ok_a & ok_c
    Rt = np.zeros((H, 3, 4), dtype=F32)
    if valid.any():
        av = a[valid]
        cv = c[valid]
        ma = av.mean(axis=1, keepdims=True)
        mc = cv.mean(axis=1, keepdims=True)
        a0 = av - ma
        c0 = cv - mc
        cov = np.einsum("hvi,hvj->hij", c0, a0)  # v1 . v0^T per hypothesis
        u, s, vh = np.linalg.svd(cov)
        R = u @ vh
        neg = np.linalg.det(R) < 0
        if neg.any():
            R[neg] -= 2.0 * u[neg][:, :, 2:3] * vh[neg][:, 2:3, :]
        t = mc[:, 0, :] - np.einsum("hij,hj->hi", R, ma[:, 0, :])
        Rt[valid, :, :3] = R.astype(F32)
        Rt[valid, :, 3] = t.astype(F32)
    return Rt.reshape(H, 12), valid.astype(np.uint8)


# --------------------------------------------------------------------------------------------
# a5/a8: scoring (C, float32 + explicit fma) and the RANSAC selection rules
# --------------------------------------------------------------------------------------------
_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)


def _fmix32(x):
    x = np.asarray(x, dtype=np.uint32)
    x = x ^ (x >> np.uint32(16))
    x = x * np.uint32(0x85EBCA6B)
    x = x ^ (x >> np.uint32(13))
    x = x * np.uint32(0xC2B2AE35)
    return x ^ (x >> np.uint32(16))


SAMPLE_REDRAWS = 7  # RDPN_SAMPLE_REDRAWS


def sample_triplets(sel, H, seed, roi_index, sample_size=3):
    """The solver's internal hypothesis sampling (include/rdpn6d_b200.h, rdpn_pose_solve with hyp_idx == NULL):
    the stand-in for np.random.choice at misc.py:91, counter-based so that it is reproducible.  sel: [P] bool gate of
    one ROI; returns [H,S] int32 absolute pixel indices (all -1 when nothing is gated).  S = 3: draws are independent, a
    sample that repeats a pixel is an invalid hypothesis (hypothesis_poses).  S > 3: without replacement as misc.py:91
    samples -- a vertex that repeats an earlier pixel of its sample is re-drawn with the attempt number in bits 20..23 of
    the counter, up to SAMPLE_REDRAWS times (include/rdpn6d_b200.h)."""
    S = int(sample_size)
    g = np.nonzero(np.asarray(sel).reshape(-1))[0]
    n = len(g)
    if n == 0:
        return np.full((H, S), -1, np.int32)
    with np.errstate(over="ignore"):
        kroi = _fmix32(_fmix32(np.uint32(seed) ^ np.uint32(0x9E3779B9)) ^ np.uint32(roi_index & 0xFFFFFFFF))
        hv = np.arange(S * H, dtype=np.uint32)
        key = _fmix32(kroi ^ hv)
    k = (key.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)
    out = g[k.astype(np.int64)].astype(np.int32).reshape(H, S)
    if S > 3:
        for h in range(H):
            for v in range(1, S):
                att = 0
                while out[h, v] in out[h, :v] and att < SAMPLE_REDRAWS:
                    att += 1
                    with np.errstate(over="ignore"):
                        kk = _fmix32(kroi ^ np.uint32(S * h + v) ^ np.uint32(att << 20))
                    out[h, v] = g[int((np.uint64(kk) * np.uint64(n)) >> np.uint64(32))]
    return out


def sq_cut(thr):
    """Smallest float32 x with sqrtf(x) >= thr: (sqrt(d2) < thr) <=> (d2 < sq_cut(thr))."""
    f = liboracle().oracle_sq_cut
    f.restype = ctypes.c_float
    return F32(f(ctypes.c_float(float(thr))))


def score_hypotheses(obj_n3, cam_n3, Rt, valid, thr):
    obj_n3 = np.ascontiguousarray(obj_n3, F32)
    cam_n3 = np.ascontiguousarray(cam_n3, F32)
    Rt = np.ascontiguousarray(Rt, F32)
    valid = np.ascontiguousarray(valid, np.uint8)
    H = Rt.shape[0]
    counts = np.zeros(H, np.int32)
    liboracle().oracle_score_hypotheses(
        obj_n3.ctypes.data_as(_f32p), cam_n3.ctypes.data_as(_f32p), ctypes.c_int(obj_n3.shape[0]),
        Rt.ctypes.data_as(_f32p), valid.ctypes.data_as(_u8p), ctypes.c_int(H),
        ctypes.c_float(float(thr)), counts.ctypes.data_as(_i32p))
    return counts


def inlier_mask(obj_n3, cam_n3, Rt12, thr, want_errs=False):
    obj_n3 = np.ascontiguousarray(obj_n3, F32)
    cam_n3 = np.ascontiguousarray(cam_n3, F32)
    Rt12 = np.ascontiguousarray(Rt12, F32).reshape(12)
    n = obj_n3.shape[0]
    m = np.zeros(n, np.uint8)
    e = np.zeros(n, F32) if want_errs else None
    liboracle().oracle_inlier_mask(
        obj_n3.ctypes.data_as(_f32p), cam_n3.ctypes.data_as(_f32p), ctypes.c_int(n),
        Rt12.ctypes.data_as(_f32p), ctypes.c_float(float(thr)), m.ctypes.data_as(_u8p),
        e.ctypes.data_as(_f32p) if want_errs else None)
    return (m, e) if want_errs else m


def select_best(counts, valid, n_sel, min_inliers=4, adaptive=False, confidence=0.995, min_iter=10):
    """misc.py:121 (strictly greater count and >= 4 -> earliest hypothesis wins ties) and, when
    adaptive, misc.py:134-138 (k = log10(1-conf)/log10(1-w^10); stop once i_ransac > max(k, min_iter)).
    Invalid hypotheses do not consume an iteration (intent of misc.py:95-101).
    Returns (best_h or -1, number of hypotheses examined)."""
    best, best_cnt, i_ransac = -1, 0, 0
    H = len(counts)
    for h in range(H):
        if not valid[h]:
            continue
        i_ransac += 1
        c = int(counts[h])
        if c > best_cnt and c >= min_inliers:
            best, best_cnt = h, c
        if adaptive:
            w = c / float(n_sel)
            with np.errstate(divide="ignore"):
                k = np.log10(1 - confidence) / np.log10(1 - pow(w, 10))
            if i_ransac > max(k, min_iter):
                return best, h + 1
    return best, H


def reference_loop_select(obj_n3, cam_n3, ww, Rt, valid, counts, n, thr, min_inliers=4, weighted=False, adaptive=False,
                          confidence=0.995, min_iter=10, scale=False):
    """Selection rule of the reference loop, lib/pysixd/misc.py:108-138, on precomputed hypothesis poses and counts:
    walk the valid hypotheses in order; (:113-116) a sample fit whose mean residual over ALL points is below the best
    so far becomes the returned pose; (:118-132) a sample fit that raises the best inlier count (>= min_inliers) is refit
    on its inliers and the refit is kept if its mean residual is below the best so far; (:134-138) adaptive stop.
    Mean residuals in float64 (errs.mean(), :113), inlier sets by the FP32 contract the counts come from.
    Returns (pose32[12] or None, source hypothesis or -1, best inlier count, inlier mask of the returned pose)."""
    a64, c64 = obj_n3.astype(F64), cam_n3.astype(F64)

    def mean_err(p12):
        P = np.asarray(p12, F64).reshape(3, 4)
        return float(np.linalg.norm(a64 @ P[:, :3].T + P[:, 3] - c64, axis=1).mean())

    best_err, best_pose, src, best_inl, i_ransac = float("inf"), None, -1, 0, 0
    for h in range(len(counts)):
        if not valid[h]:
            continue
        i_ransac += 1
        e = mean_err(Rt[h])
        if e < best_err:
            best_err, best_pose, src = e, Rt[h].copy(), h
        c = int(counts[h])
        if c > best_inl and c >= min_inliers:
            best_inl = c
            m = inlier_mask(obj_n3, cam_n3, Rt[h], thr).astype(bool)
            if int(m.sum()) >= 3:
                M = kabsch(obj_n3[m].T, cam_n3[m].T, w=(ww[m].astype(F64) if weighted else None), scale=scale)
                pr = M[:3, :4].astype(F32).reshape(12)
                er = mean_err(pr)
                if er < best_err:
                    best_err, best_pose, src = er, pr, h
        if adaptive:
            wr = c / float(n)
            with np.errstate(divide="ignore"):
                k = np.log10(1 - confidence) / np.log10(1 - pow(wr, 10))
            if i_ransac > max(k, min_iter):
                break
    if best_pose is None:
        return None, -1, 0, None
    return best_pose, src, best_inl, inlier_mask(obj_n3, cam_n3, best_pose, thr)


def solve_roi(cam, obj, w, sel, hyp_idx, thr, min_pts=4, min_inliers=4, weighted=False,
              refit_iters=1, adaptive=False, confidence=0.995, min_iter=10, scale=False,
              t_net=None, select_rule="most_inliers"):
    """Stages S3-S5 for one ROI on the S1 output.  cam, obj: [3,P] float32; w: [P]; sel: [P] bool.

    Returns dict(pose[3,4] f32, n_inl, status, best_h, n_sel, counts[H], valid[H], Rt_hyp[H,12],
                 inlier_mask[P] uint8 (pixels used by the LAST refit), scale).
    """
    P = cam.shape[1]
    H = hyp_idx.shape[0]
    out = dict(pose=np.full((3, 4), -100, F32), n_inl=0, status=STATUS_OK, best_h=-1,
               n_sel=int(sel.sum()), counts=np.zeros(H, np.int32), valid=np.zeros(H, np.uint8),
               Rt_hyp=np.zeros((H, 12), F32), inlier_mask=np.zeros(P, np.uint8), scale=1.0)
    n = out["n_sel"]
    if n < min_pts:  # gdrn_evaluator.py:380-395
        out["status"] = STATUS_FEW_POINTS
        return out
    pix = np.nonzero(sel)[0]
    obj_n3 = np.ascontiguousarray(obj[:, pix].T)
    cam_n3 = np.ascontiguousarray(cam[:, pix].T)
    Rt, valid = hypothesis_poses(obj, cam, sel, hyp_idx)
    counts = score_hypotheses(obj_n3, cam_n3, Rt, valid, thr)
    out.update(counts=counts, valid=valid, Rt_hyp=Rt)
    if select_rule == "min_mean_err":  # the reference loop's return value (misc.py:113-132)
        pose32, src, best_inl, m = reference_loop_select(obj_n3, cam_n3, w[pix], Rt, valid, counts, n, thr, min_inliers, weighted,
                                                         adaptive, confidence, min_iter, scale)
        if pose32 is None:
            out["status"] = STATUS_NO_CONSENSUS
            return out
        full = np.zeros(P, np.uint8)
        full[pix] = m
        out.update(best_h=src, n_inl=best_inl, inlier_mask=full, pose=np.asarray(pose32, F32).reshape(3, 4).copy())
        if t_net is not None and te(out["pose"][:, 3], np.asarray(t_net)) > 1.0:
            out["status"] = STATUS_T_SANITY
            out["pose"][:, 3] = np.asarray(t_net, F32)
        return out
    best, _ = select_best(counts, valid, n, min_inliers, adaptive, confidence, min_iter)
    if best < 0:
        out["status"] = STATUS_NO_CONSENSUS
        return out
    out["best_h"] = best
    out["n_inl"] = int(counts[best])
    pose32 = Rt[best]
    for _ in range(max(1, refit_iters)):
        m = inlier_mask(obj_n3, cam_n3, pose32, thr)
        if int(m.sum()) < 3:
            break
        k = m.astype(bool)
        ww = w[pix][k].astype(F64) if weighted else None
        M = kabsch(obj_n3[k].T, cam_n3[k].T, w=ww, scale=scale)  # misc.py:123-126 refit on inliers
        if scale:
            out["scale"] = float(np.cbrt(np.linalg.det(M[:3, :3])))
        pose32 = M[:3, :4].astype(F32).reshape(12)
        full = np.zeros(P, np.uint8)
        full[pix] = m
        out["inlier_mask"] = full
    out["pose"] = pose32.reshape(3, 4).copy()
    if t_net is not None and te(out["pose"][:, 3], np.asarray(t_net)) > 1.0:  # gdrn_evaluator.py:293-296
        out["status"] = STATUS_T_SANITY
        out["pose"][:, 3] = np.asarray(t_net, F32)
    return out


def pose_solve_batch(batch, hyp_idx, thr, **kw):
    """Whole path over a batch dict as produced by rdpn6d_b200.synth.make_batch (numpy arrays):
    depth[B,h,w], Kp[B,4], coor[B,3,h,w], mask[B,h,w], extent[B,3], region_idx[B,h,w] | None,
    anchors[B,R,3] | None, depth_div[B] | None.  Returns list of solve_roi dicts (+ S1 arrays)."""
    B = batch["depth"].shape[0]
    s1_kw = dict(mask_mode=kw.pop("mask_mode", MASK_L1), mask_thr=kw.pop("mask_thr", 0.5))
    t_net = kw.pop("t_net", None)
    outs = []
    for b in range(B):
        c = correspondences(
            batch["depth"][b], batch["Kp"][b], batch["coor"][b], batch["mask"][b], batch["extent"][b],
            None if batch.get("region_idx") is None else batch["region_idx"][b],
            None if batch.get("anchors") is None else batch["anchors"][b],
            None if batch.get("depth_div") is None else batch["depth_div"][b], **s1_kw)
        r = solve_roi(c["cam"], c["obj"], c["w"], c["sel"], hyp_idx[b], thr,
                      t_net=None if t_net is None else t_net[b], **kw)
        r["s1"] = c
        outs.append(r)
    return outs


# --------------------------------------------------------------------------------------------
# a10: tolerance metrics
# --------------------------------------------------------------------------------------------
def re(R_est, R_gt):
    """pose_error.py:400-415, degrees."""
    R_est = np.asarray(R_est, F64)
    R_gt = np.asarray(R_gt, F64)
    assert R_est.shape == R_gt.shape == (3, 3)
    trace = np.trace(R_est @ R_gt.T)
    trace = trace if trace <= 3 else 3
    return float(np.rad2deg(np.arccos(min(1.0, max(-1.0, 0.5 * (trace - 1.0))))))


def re_rad_small(R_est, R_gt):
    """Rotation angle of R_est R_gt^T in radians, from the skew part (acos loses everything below
    ~1e-8 rad near zero; the tolerance tests need 1e-5 rad resolution in float64)."""
    D = np.asarray(R_est, F64) @ np.asarray(R_gt, F64).T
    s = 0.5 * math.sqrt((D[2, 1] - D[1, 2]) ** 2 + (D[0, 2] - D[2, 0]) ** 2 + (D[1, 0] - D[0, 1]) ** 2)
    c = 0.5 * (np.trace(D) - 1.0)
    return float(math.atan2(s, c))


def te(t_est, t_gt):
    """pose_error.py:425-436."""
    t_est = np.asarray(t_est, F64).flatten()
    t_gt = np.asarray(t_gt, F64).flatten()
    assert t_est.size == t_gt.size == 3
    return float(np.linalg.norm(t_gt - t_est))


def transform_pts_Rt(pts, R, t):
    """misc.py:895-905."""
    assert pts.shape[1] == 3
    return (R.dot(pts.T) + np.asarray(t).reshape((3, 1))).T


def transform_pts_batch(pts, R, t=None):
    """misc.py:930-949 (torch, batched): pts [B,P,3], R [B,3,3], t [B,3,1] or None -> [B,P,3] in the dtype of the inputs."""
    pts, R = np.asarray(pts), np.asarray(R)
    bs, n_pts = R.shape[0], pts.shape[1]
    assert pts.shape == (bs, n_pts, 3)
    out = np.matmul(R.reshape(bs, 1, 3, 3), pts.reshape(bs, n_pts, 3, 1))
    if t is not None:
        assert t.shape[0] == bs
        out = out + np.asarray(t).reshape(bs, 1, 3, 1)
    return out[..., 0]


def add(R_est, t_est, R_gt, t_gt, pts):
    """pose_error.py:297-312."""
    return float(np.linalg.norm(transform_pts_Rt(pts, R_est, t_est) - transform_pts_Rt(pts, R_gt, t_gt), axis=1).mean())


# --------------------------------------------------------------------------------------------
# a9: pose assembly from (rot6d | R_allo, centroid, z)
# --------------------------------------------------------------------------------------------
def adi(R_est, t_est, R_gt, t_gt, pts):
    """pose_error.py:315-337: mean nearest-neighbour distance (brute force in float64; the reference uses a cKDTree,
    which returns the same nearest distances)."""
    pe = transform_pts_Rt(np.asarray(pts, F64), np.asarray(R_est, F64), np.asarray(t_est, F64))
    pg = transform_pts_Rt(np.asarray(pts, F64), np.asarray(R_gt, F64), np.asarray(t_gt, F64))
    d = np.sqrt(((pg[:, None, :] - pe[None, :, :]) ** 2).sum(-1))
    return float(d.min(axis=1).mean())


def get_closest_rot(rot_est, rot_gt, sym_info):
    """core/utils/pose_utils.py:430-454."""
    if sym_info is None:
        return rot_gt
    sym_info = np.asarray(sym_info)
    if sym_info.ndim == 2:
        sym_info = sym_info.reshape((1, 3, 3))
    r_err, closest = re(rot_est, rot_gt), rot_gt
    for i in range(sym_info.shape[0]):
        cand = rot_gt.dot(sym_info[i])
        cur = re(rot_est, cand)
        if cur < r_err:
            r_err, closest = cur, cand
    return closest


def rotation_matrix(angle, axis):
    """lib/pysixd/transform.py:296-336 (rotation about an axis through the origin), 3x3 part."""
    d = np.asarray(axis, F64)[:3]
    d = d / math.sqrt(float(np.dot(d, d)))
    sina, cosa = math.sin(angle), math.cos(angle)
    R = np.diag([cosa, cosa, cosa])
    R += np.outer(d, d) * (1.0 - cosa)
    d = d * sina
    R += np.array([[0.0, -d[2], d[1]], [d[2], 0.0, -d[0]], [-d[1], d[0], 0.0]])
    return R


def get_symmetry_transformations(model_info, max_sym_disc_step):
    """lib/pysixd/misc.py:206-254: the discrete symmetries of models_info.json (identity first) combined with the
    discretised continuous ones (ceil(pi / max_sym_disc_step) steps about each axis); list of {"R": [3,3], "t": [3,1]}."""
    trans_disc = [{"R": np.eye(3), "t": np.array([[0, 0, 0]]).T}]
    for sym in model_info.get("symmetries_discrete", []):
        s44 = np.reshape(sym, (4, 4))
        trans_disc.append({"R": s44[:3, :3], "t": s44[:3, 3].reshape((3, 1))})
    trans_cont = []
    for sym in model_info.get("symmetries_continuous", []):
        axis = np.array(sym["axis"])
        offset = np.array(sym["offset"]).reshape((3, 1))
        steps = int(np.ceil(np.pi / max_sym_disc_step))
        step = 2.0 * np.pi / steps
        for i in range(1, steps):
            R = rotation_matrix(i * step, axis)
            trans_cont.append({"R": R, "t": -R.dot(offset) + offset})
    trans = []
    for td in trans_disc:
        if len(trans_cont):
            for tc in trans_cont:
                trans.append({"R": tc["R"].dot(td["R"]), "t": tc["R"].dot(td["t"]) + tc["t"]})
        else:
            trans.append(td)
    return trans


def backproject_v2(depth, K):
    """misc.py:352-371."""
    Kinv = np.linalg.inv(K)
    h, w = depth.shape
    gx, gy = np.meshgrid(np.arange(w), np.arange(h))
    g2 = np.stack([gx, gy, np.ones((h, w))], axis=2)
    return depth.reshape(h, w, 1) * (g2 @ Kinv.T)


def calc_emb_bp_fast(depth, R, T, K):
    """misc.py:288-316."""
    pc = backproject_v2(depth, K) - np.asarray(T, F64).reshape(1, 1, 3)
    return (pc @ np.asarray(R, F64)) * (depth != 0).astype(depth.dtype).reshape(*depth.shape, 1)


def ortho6d_to_mat(p6):
    """rot_reps.py:34-49 (float32, batch [B,6] -> [B,3,3], columns x,y,z)."""
    p6 = np.asarray(p6, F32)

    def _norm(v):  # rot_reps.py normalize_vector: v / max(|v|, 1e-8)
        n = np.sqrt((v * v).sum(axis=1, keepdims=True, dtype=F32)).astype(F32)
        return (v / np.maximum(n, F32(1e-8))).astype(F32)

    x = _norm(p6[:, 0:3])
    z = _norm(np.cross(x, p6[:, 3:6]).astype(F32))
    y = np.cross(z, x).astype(F32)
    return np.stack([x, y, z], axis=2)


def quat2mat(q):
    """transforms3d.quaternions.quat2mat (w, x, y, z; normalises internally; identity below float eps) -- the third-party
    call behind RT_transform.quat_trans_to_pose_m (lib/pysixd/RT_transform.py:177-183)."""
    w, x, y, z = [float(v) for v in q]
    Nq = w * w + x * x + y * y + z * z
    if Nq < np.finfo(np.float64).eps:
        return np.eye(3)
    s = 2.0 / Nq
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ = w * X, w * Y, w * Z
    xX, xY, xZ = x * X, x * Y, x * Z
    yY, yZ, zZ = y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY], [xY + wZ, 1.0 - (xX + zZ), yZ - wX], [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


def axangle2mat(axis, angle):
    """transforms3d.axangles.axangle2mat (Rodrigues), the only transforms3d call in utils.py:39-94."""
    x, y, z = np.asarray(axis, F64) / np.linalg.norm(axis)
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    xs, ys, zs = x * s, y * s, z * s
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    return np.array([[x * xC + c, xyC - zs, zxC + ys],
                     [xyC + zs, y * yC + c, yzC - xs],
                     [zxC - ys, yzC + xs, z * zC + c]])


def allocentric_to_egocentric_mat(R_allo, trans, pose_dtype=F32):
    """utils.py:39-94 for src_type=dst_type='mat', cam_ray=(0,0,1): R_ego = Rodrigues(cam x obj, acos(obj_z)) R_allo.
    pose_dtype: dtype of the [R|t] array the caller hands to the reference function -- float32 from the matrix heads
    (np.hstack of float32 tensors), float64 from the quaternion heads (RT_transform.quat_trans_to_pose_m fills a float64
    array, RT_transform.py:177-183).

    dtype flow of the reference as called from pose_from_pred_centroid_z.py:127-136 (pinned by
    tests/golden/path_golden.npz): the pose is float32, so the object ray is normalised in float32; angle, axis and the
    Rodrigues matrix are float64; the product is rounded to float32 when stored into the float32 ego pose."""
    cam_ray = np.array([0, 0, 1.0])
    trans = np.asarray(trans, pose_dtype)
    obj_ray = trans.copy() / np.linalg.norm(trans)  # in the pose's dtype
    angle = math.acos(cam_ray.dot(obj_ray))
    if angle > 0:
        return np.dot(axangle2mat(np.cross(cam_ray, obj_ray), angle), np.asarray(R_allo, pose_dtype))
    return np.asarray(R_allo, F64).copy()


def pose_from_pred_centroid_z_test(pred_rots, pred_centroids, pred_z_vals, roi_cams, roi_centers,
                                   resize_ratios, roi_whs, is_allo=True, z_type="REL"):
    """pose_from_pred_centroid_z.py:52-141 (rot-matrix branch). Translation in float32 with the
    reference's operation order z*(cx-px)/fx; allo->ego per ROI in float64 then cast to float32."""
    pc = np.asarray(pred_centroids, F32)
    whs = np.asarray(roi_whs, F32)
    ctr = np.asarray(roi_centers, F32)
    K = np.asarray(roi_cams, F32)
    if K.ndim == 2:
        K = K[None]
    cx = pc[:, 0] * whs[:, 0] + ctr[:, 0]
    cy = pc[:, 1] * whs[:, 1] + ctr[:, 1]
    z = np.asarray(pred_z_vals, F32).reshape(-1)
    if z_type == "REL":
        z = z * np.asarray(resize_ratios, F32).reshape(-1)
    elif z_type != "ABS":
        raise ValueError(f"Unknown z_type: {z_type}")
    tx = (z * (cx - K[:, 0, 2])) / K[:, 0, 0]
    ty = (z * (cy - K[:, 1, 2])) / K[:, 1, 1]
    trans = np.stack([tx, ty, z], axis=1).astype(F32)
    rots = np.asarray(pred_rots, F32)
    ego = np.zeros_like(rots)
    for i in range(rots.shape[0]):
        ego[i] = allocentric_to_egocentric_mat(rots[i], trans[i]).astype(F32) if is_allo else rots[i]
    return ego, trans


# --------------------------------------------------------------------------------------------
# a6: the CPU solver the reference actually executes (timed baseline only; parity unpinned)
# --------------------------------------------------------------------------------------------
def pnp_v2_as_run(points_3d, points_2d, K, ransac_reprojErr=3.0, ransac_iter=100):
    """misc.py:145-194 with method=EPnP, ransac=True, as called at gdrn_evaluator.py:382-392.
    Third-party arithmetic (OpenCV): used as a timed baseline and a sanity cross-check only."""
    import cv2

    dist = np.zeros((8, 1), dtype="float64")
    p3 = np.ascontiguousarray(np.expand_dims(points_3d, 0).astype(np.float64))
    p2 = np.ascontiguousarray(np.expand_dims(points_2d, 0).astype(np.float64))
    _, rvec, t, _ = cv2.solvePnPRansac(p3, p2, np.asarray(K, np.float64), dist, flags=cv2.SOLVEPNP_EPNP,
                                       reprojectionError=ransac_reprojErr, iterationsCount=ransac_iter)
    R, _ = cv2.Rodrigues(rvec)
    return np.concatenate([R, t.reshape((3, 1))], axis=-1)


# --------------------------------------------------------------------------------------------
# f4: training-side region targets (adjacent row; kept for the "next" scope)
# --------------------------------------------------------------------------------------------
def xyz_to_region(xyz_crop, fps_points):
    """core/utils/data_utils.py:229-244: nearest-anchor region ids (1..R, 0 = background) and
    delta = xyz - anchor.  xyz_crop [h,w,3], fps_points [R,3] (float64 distances as scipy cdist)."""
    xyz_crop = np.asarray(xyz_crop)
    fps_points = np.asarray(fps_points)
    bh, bw = xyz_crop.shape[:2]
    mask_crop = ((xyz_crop[:, :, 0] != 0) | (xyz_crop[:, :, 1] != 0) | (xyz_crop[:, :, 2] != 0)).astype("uint8")
    diff = xyz_crop.reshape(bh * bw, 1, 3).astype(F64) - fps_points[None].astype(F64)
    dists = np.sqrt((diff * diff).sum(-1))
    region_ids = np.argmin(dists, axis=1).reshape(bh, bw) + 1
    delta = xyz_crop - fps_points[region_ids - 1]
    return mask_crop * region_ids, delta


# --------------------------------------------------------------------------------------------
# f1: ROI depth crop as the loader does it (data_loader.py:532-535 -> data_utils.py:81-96), via OpenCV
# --------------------------------------------------------------------------------------------
def roi_crop_depth_cv2(depth_img, center, scale, crop_res=256, stride=4):
    """cv2.warpAffine(depth, A, (crop,crop), INTER_LINEAR)[::stride, ::stride] with A = roi_affine().
    Third-party arithmetic (OpenCV): a cross-check for rdpn_roi_crop_depth, not a pinned oracle."""
    import cv2

    A = roi_affine(center, scale, crop_res)
    crop = cv2.warpAffine(np.asarray(depth_img, F32), A, (int(crop_res), int(crop_res)), flags=cv2.INTER_LINEAR)
    return crop[::stride, ::stride]


def roi_crop_depth(depth_img, center, scale, crop_res=256, out_res=64):
    """numpy restatement of OpenCV's warpAffine + remap for CV_32F / INTER_LINEAR / BORDER_CONSTANT(0)
    (modules/imgproc/src/imgwarp.cpp; third-party: opencv-python 4.5.5.62 pinned by the reference), evaluated
    only at the pixels the loader keeps (data_loader.py:625).  Bit-identical to cv2 4.13 in this container
    (tests/test_oracle_pose.py); slow Python loops, small cases only."""
    depth = np.asarray(depth_img, F32)
    M = roi_affine(center, scale, crop_res).flatten().copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    H, W = depth.shape
    st = crop_res // out_res
    res = np.zeros((out_res, out_res), F32)
    one = F32(1)

    def tap(yy, xx):
        return depth[yy, xx] if 0 <= xx < W and 0 <= yy < H else F32(0)

    for j in range(out_res):
        y = st * j
        X0 = int(np.rint((M[1] * y + M[2]) * 1024)) + 16
        Y0 = int(np.rint((M[4] * y + M[5]) * 1024)) + 16
        for i in range(out_res):
            x = st * i
            X = (X0 + int(np.rint(M[0] * x * 1024))) >> 5
            Y = (Y0 + int(np.rint(M[3] * x * 1024))) >> 5
            sx, sy = X >> 5, Y >> 5
            fx, fy = F32((X & 31) / 32.0), F32((Y & 31) / 32.0)
            r = F32(tap(sy, sx) * F32((one - fy) * (one - fx)))
            r = F32(r + F32(tap(sy, sx + 1) * F32((one - fy) * fx)))
            r = F32(r + F32(tap(sy + 1, sx) * F32(fy * (one - fx))))
            r = F32(r + F32(tap(sy + 1, sx + 1) * F32(fy * fx)))
            res[j, i] = r
    return res
