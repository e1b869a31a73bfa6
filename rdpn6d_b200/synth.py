"""Deterministic synthetic ROI batches for the dense-correspondence -> pose path (numpy only).

Shapes and statistics follow SURVEY.md section 8(d): 64x64 ROI maps of depth + normalised residual
xyz + mask + region index, per-ROI anchors (farthest-point samples of the object model), extents
and crop-adjusted intrinsics, plus RANSAC hypothesis index triplets drawn from the gated pixels.
Objects are analytic (ellipsoids, boxes) ray-cast through the crop camera so the maps are dense and
the ground-truth pose is known.  Camera constants are the reference's dataset intrinsics
(/root/reference/ref/lm_full.py:106, ref/ycbv.py:89).

This module is input generation only: no CUDA, no oracle import.
"""
import numpy as np

F32 = np.float32

K_LM = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]], dtype=np.float64)
K_YCBV = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]], dtype=np.float64)

ROI = 64  # cfg.MODEL.CDPN.BACKBONE.OUTPUT_RES (configs/_base_/gdrn_base.py:26)
CROP = 256  # INPUT_RES (:25); the 64x64 maps are pixels (4i,4j) of the 256x256 crop (data_loader.py:625)


def random_rotations(rng, n):
    """Uniform SO(3) from normalised quaternions, float64 [n,3,3]."""
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)


def _fps_numpy(pts, k):
    """Plain numpy farthest-point sampling used only to place synthetic anchors."""
    ctr = 0.5 * (pts.max(0) + pts.min(0))
    d = ((pts - ctr) ** 2).sum(1)
    out = []
    cur = int(np.argmax(d))
    d = np.full(len(pts), np.inf)
    for _ in range(k):
        out.append(cur)
        d = np.minimum(d, ((pts - pts[cur]) ** 2).sum(1))
        cur = int(np.argmax(d))
    return pts[out]


class ObjectModel:
    """Analytic object: 'ellipsoid' or 'box' with half-sizes `half` (metres) and R anchors."""

    def __init__(self, kind, half, num_regions, rng):
        self.kind = kind
        self.half = np.asarray(half, dtype=np.float64)
        self.extent = (2.0 * self.half).astype(F32)
        self.surface = self.sample_surface(rng, 5000)
        self.anchors = _fps_numpy(self.surface, num_regions).astype(F32)

    def sample_surface(self, rng, n):
        if self.kind == "ellipsoid":
            v = rng.standard_normal((n, 3))
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            return v * self.half
        u = rng.uniform(-1, 1, (n, 3))
        face = rng.integers(0, 3, n)
        sign = rng.choice([-1.0, 1.0], n)
        u[np.arange(n), face] = sign
        return u * self.half

    def raycast(self, o, d):
        """o, d: [...,3] ray origins/directions in the object frame -> (hit[...], t[...])."""
        if self.kind == "ellipsoid":
            oo = o / self.half
            dd = d / self.half
            a = (dd * dd).sum(-1)
            b = 2 * (oo * dd).sum(-1)
            c = (oo * oo).sum(-1) - 1.0
            disc = b * b - 4 * a * c
            hit = disc > 0
            t = (-b - np.sqrt(np.where(hit, disc, 0.0))) / (2 * a)
            return hit & (t > 0), t
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (-self.half - o) / d
            t2 = (self.half - o) / d
        tn = np.minimum(t1, t2).max(-1)
        tf = np.maximum(t1, t2).min(-1)
        hit = (tn < tf) & (tn > 0)
        return hit, tn


def make_models(n_models, num_regions=32, seed=0, n_symmetric=0):
    """n_models analytic objects with extents ~U(0.08, 0.25) m (LM diameters 0.10-0.28 m,
    ref/lm_full.py:77-98); the last n_symmetric ones have symmetric geometry (spheroid / cube-like
    boxes), mirroring SYM_OBJS of the YCB-V config."""
    rng = np.random.default_rng(seed)
    models = []
    for i in range(n_models):
        half = rng.uniform(0.04, 0.125, 3)
        kind = "ellipsoid" if i % 3 != 2 else "box"
        if i >= n_models - n_symmetric:
            half[1] = half[0]  # revolution / square section
        models.append(ObjectModel(kind, half, num_regions, rng))
    return models


def make_batch(B, models=None, H=256, num_regions=32, K=K_LM, seed=20260101, im_hw=(480, 640),
               dzi_pad_scale=1.5, noise_sigma=0.001, outlier_frac=0.15, mask_dropout=0.10,
               occlusion_max=0.0, dense=False, chunk=256):
    """Build one batch.  Returns a dict of contiguous numpy arrays:

    depth[B,64,64] f32 (metres, 0 = no depth)      Kp[B,4] f32 (fx',fy',cx',cy' of the 256 crop)
    coor[B,3,64,64] f32 (normalised residual)      mask[B,64,64] f32 (raw head output)
    region_idx[B,64,64] u8, anchors[B,R,3] f32     extent[B,3] f32
    hyp_idx[B,H,3] i32 (absolute pixel indices)    gt_pose[B,3,4] f64, model_id[B] i32
    K[B,3,3] f32, bbox_center[B,2] f32, scale[B] f32, resize_ratio[B] f32, roi_wh[B,2] f32
    dense=True: no anchors/region (coor is the normalised object coordinate itself).
    """
    if models is None:
        models = make_models(8, num_regions, seed=seed % 1000)
    rng = np.random.default_rng(seed)
    R_ = len(models[0].anchors)
    out = dict(
        depth=np.zeros((B, ROI, ROI), F32), Kp=np.zeros((B, 4), F32), coor=np.zeros((B, 3, ROI, ROI), F32),
        mask=np.zeros((B, ROI, ROI), F32), region_idx=np.zeros((B, ROI, ROI), np.uint8),
        anchors=np.zeros((B, R_, 3), F32), extent=np.zeros((B, 3), F32), hyp_idx=np.zeros((B, H, 3), np.int32),
        gt_pose=np.zeros((B, 3, 4), np.float64), model_id=np.zeros(B, np.int32), K=np.zeros((B, 3, 3), F32),
        bbox_center=np.zeros((B, 2), F32), scale=np.zeros(B, F32), resize_ratio=np.zeros(B, F32),
        roi_wh=np.zeros((B, 2), F32))
    Rs = random_rotations(rng, B)
    ts = np.stack([rng.uniform(-0.2, 0.2, B), rng.uniform(-0.15, 0.15, B), rng.uniform(0.6, 1.2, B)], 1)
    mids = np.arange(B) % len(models)
    K = np.asarray(K, np.float64)
    jj, ii = np.meshgrid(np.arange(ROI), np.arange(ROI), indexing="ij")  # row j, col i
    ucrop = (4.0 * ii).reshape(-1)
    vcrop = (4.0 * jj).reshape(-1)
    for b in range(B):
        m = models[mids[b]]
        R, t = Rs[b], ts[b]
        # bbox from the projected corners of the extent box (+ detector jitter)
        corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * m.half
        pc = corners @ R.T + t
        uv = (pc[:, :2] / pc[:, 2:3]) * np.array([K[0, 0], K[1, 1]]) + np.array([K[0, 2], K[1, 2]])
        x1, y1 = uv.min(0) + rng.uniform(-3, 3, 2)
        x2, y2 = uv.max(0) + rng.uniform(-3, 3, 2)
        # core/gdrn_modeling/data_loader.py:472-488
        center = np.array([0.5 * (x1 + x2), 0.5 * (y1 + y2)])
        bw, bh = max(x2 - x1, 1), max(y2 - y1, 1)
        scale = min(max(bh, bw) * dzi_pad_scale, max(im_hw)) * 1.0
        s = CROP / scale
        fxp, fyp = s * K[0, 0], s * K[1, 1]
        cxp = s * K[0, 2] + (CROP / 2 - s * center[0])
        cyp = s * K[1, 2] + (CROP / 2 - s * center[1])
        # rays through crop pixels, in the object frame
        dirs = np.stack([(ucrop - cxp) / fxp, (vcrop - cyp) / fyp, np.ones_like(ucrop)], 1)
        o = -(R.T @ t)
        d = dirs @ R  # R^T dir
        hit, tt = m.raycast(o[None, :], d)
        tt = np.where(hit, tt, 0.0)
        xo = o[None, :] + tt[:, None] * d  # object coordinates of the visible surface
        z = tt  # camera depth (dirs has z = 1)
        if occlusion_max > 0 and hit.any():  # rectangular occluder over U(0, occlusion_max) of the object's box
            frac = rng.uniform(0, occlusion_max)
            hh = hit.reshape(ROI, ROI)
            rows, cols = np.nonzero(hh.any(1))[0], np.nonzero(hh.any(0))[0]
            r0, r1, c0, c1 = rows[0], rows[-1] + 1, cols[0], cols[-1] + 1
            oh = int(round((r1 - r0) * np.sqrt(frac)))
            ow = int(round((c1 - c0) * np.sqrt(frac)))
            if oh > 0 and ow > 0:
                oy = r0 if rng.random() < 0.5 else r1 - oh  # slides in from a random corner of the box
                ox = c0 if rng.random() < 0.5 else c1 - ow
                occ = np.zeros((ROI, ROI), bool)
                occ[oy:oy + oh, ox:ox + ow] = True
                hit = hit & ~occ.reshape(-1)
        fg = hit
        if dense:
            delta_cam = xo  # coor is the object coordinate itself
            rid = np.zeros(ROI * ROI, np.int64)
        else:
            d2 = ((xo[:, None, :] - m.anchors[None, :, :].astype(np.float64)) ** 2).sum(-1)
            rid = np.argmin(d2, axis=1)  # core/utils/data_utils.py:229-244 (xyz_to_region)
            delta_cam = (xo - m.anchors[rid].astype(np.float64)) @ R.T  # data_loader.py:883-887
        delta_cam = delta_cam + rng.normal(0, noise_sigma, delta_cam.shape)
        outl = rng.random(ROI * ROI) < outlier_frac
        delta_cam[outl] = rng.uniform(-0.5, 0.5, (int(outl.sum()), 3)) * m.extent
        coor = delta_cam / m.extent + 0.5  # data_loader.py:899-903
        coor[~fg] = 0.0
        rid = np.where(fg, rid, rng.integers(0, R_, ROI * ROI))
        keep = fg & (rng.random(ROI * ROI) >= mask_dropout)
        mask = np.where(keep, rng.uniform(0.55, 0.95, ROI * ROI), rng.uniform(0.05, 0.45, ROI * ROI))
        out["depth"][b] = np.where(fg, z, 0.0).reshape(ROI, ROI)
        out["Kp"][b] = [fxp, fyp, cxp, cyp]
        out["coor"][b] = coor.T.reshape(3, ROI, ROI)
        out["mask"][b] = mask.reshape(ROI, ROI)
        out["region_idx"][b] = rid.reshape(ROI, ROI)
        out["anchors"][b] = m.anchors
        out["extent"][b] = m.extent
        out["gt_pose"][b, :, :3] = R
        out["gt_pose"][b, :, 3] = t
        out["model_id"][b] = mids[b]
        out["K"][b] = K
        out["bbox_center"][b] = center
        out["scale"][b] = scale
        out["resize_ratio"][b] = ROI / scale
        out["roi_wh"][b] = [bw, bh]
        # hypothesis triplets from the (approximately) gated pixel list
        mn, mx = mask.min(), mask.max()
        dl = (out["coor"][b].reshape(3, -1) - F32(0.5)) * m.extent[:, None]
        cand = np.nonzero(fg & ((mask - mn) / (mx - mn) > 0.5) & (np.abs(dl) > 1e-4 * m.extent[:, None]).all(0))[0]
        if len(cand) >= 3:
            out["hyp_idx"][b] = cand[rng.integers(0, len(cand), (H, 3))]
    if dense:
        out["region_idx"] = None
        out["anchors"] = None
    return out


def tile_batch(batch, B):
    """Repeat a batch along dim 0 up to B ROIs (large multi-GPU configs reuse a generated base set)."""
    out = {}
    for k, v in batch.items():
        if v is None:
            out[k] = None
            continue
        reps = -(-B // v.shape[0])
        out[k] = np.ascontiguousarray(np.concatenate([v] * reps, axis=0)[:B])
    return out


def fps_cloud(n, seed=0):
    """anisotropic gaussian cloud, float32 [n,3]."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, 3)) * np.array([0.1, 0.07, 0.05])).astype(F32)
