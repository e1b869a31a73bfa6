"""oracle/gen_golden.py -- generates tests/golden/*.npz by EXECUTING THE REFERENCE in this container.

Run from the repository root:  python -m oracle.gen_golden
Needs /root/reference (read-only mount).  The reference modules that import cleanly here are loaded by
file path (lib/pysixd/transform.py, core/utils/data_utils.py); the reference FPS C++ is compiled by
oracle/Makefile into oracle/_ref/.  Nothing is copied from the reference: only inputs we generate and
the outputs the reference computes for them are stored.

Files written (small, committed):
  fps_golden.npz      clouds (or their seeds) + indices from the reference C++ build
  kabsch_golden.npz   point sets + 4x4 matrices from transform.affine_matrix_from_points /
                      superimposition_matrix, incl. the doctest literal (transform.py:893-898),
                      a reflection case (:945-948) and Umeyama scale (:971-975)
  affine_golden.npz   (center, scale) + 2x3 matrices from data_utils.get_affine_transform (:111-152)
  region_golden.npz   xyz crops + anchors + (region ids, delta) from data_utils.xyz_to_region (:229-244)
  pose_golden.npz     a 4-ROI synthetic batch + the composite's outputs where EVERY Kabsch call
                      (hypotheses and refit) went through the reference's affine_matrix_from_points
  path_golden.npz     outputs of reference functions whose MODULES do not import here (mmcv / detectron2 /
                      transforms3d are absent) but whose bodies are plain numpy / torch: the function source is cut
                      out of the reference file with `ast` and executed as it stands (ref_functions below):
                        gate + de-normalisation   gdrn_evaluator.py:89-126  get_img_model_points_with_coords2d
                        mask post-processing      engine_utils.py:118-136   get_out_mask (L1, BCE)
                        back-projection           misc.py:319-349           backproject, backproject_th
                        rigid apply               misc.py:895-905           transform_pts_Rt
                        re / te                   pose_error.py:400-436
                        rot6d                     rot_reps.py:8-49          ortho6d_to_mat_batch (+ helpers)
                        allo -> ego, pose assembly  utils.py:39-94, pose_from_pred_centroid_z.py:52-141, with the one
                                                  missing third-party call (transforms3d.axangles.axangle2mat =
                                                  Rodrigues' formula) supplied by the oracle
  fps_center_golden.npz  the reference's Python FPS surface (fps_utils.py:6-21 + data_utils.get_fps_and_center :217-226) run
                      from source on top of the reference's own C++ build
  roi_scalars_golden.json  bbox -> (centre, scale, resize ratio, wh): data_loader.py:477-482, :488 run from their source lines
  sampler_golden.json per-rank index ranges of InferenceSampler (my_distributed_sampler.py:170-199) run from source
  rows_golden.json    BOP result rows from GDRN_Evaluator.pose_prediction_to_json (gdrn_evaluator.py:483-513) run from source
  ransac_roi_golden.npz  misc.pnp_ransac_custom (misc.py:58-142) run from source on the correspondences of 4 synthetic
                      ROIs (10 pairs per sample, reference Kabsch, float64 scoring): sampled pixel sets + inlier counts
"""
import importlib.util
import os
import sys

import numpy as np

REF = os.environ.get("RDPN_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_functions(rel, names, env=None):
    """Compile the named top-level functions or methods of a reference file WITHOUT importing the module: their
    source segments are taken verbatim from the file (ast) and executed in a namespace holding numpy / torch / math
    and `env`.  Methods become plain functions (pass self=None).  Nothing is written anywhere."""
    import ast
    import math

    import torch

    src = open(os.path.join(REF, rel)).read()
    tree = ast.parse(src)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            found[node.name] = node
    ns = {"np": np, "torch": torch, "math": math, "random": __import__("random")}
    ns.update(env or {})
    for name in names:
        seg = ast.get_source_segment(src, found[name])
        import textwrap

        exec(compile(textwrap.dedent(seg), os.path.join(REF, rel), "exec"), ns)
    return {n: ns[n] for n in names}


class _Cv2Shim:
    """cv2 stand-in that makes misc.pnp_ransac_custom's 2D-3D solver calls 3D-3D: solvePnP -> the reference's own
    Kabsch (transform.affine_matrix_from_points), projectPoints -> rigid apply, Rodrigues -> identity."""
    SOLVEPNP_ITERATIVE = 0

    def __init__(self, tf):
        self.tf, self.sample_solves, self.proj = tf, 0, []

    def solvePnP(self, mp_, ip_, K_, dist, flags=0):
        if len(mp_) == 10:
            self.sample_solves += 1
        M = self.tf.affine_matrix_from_points(np.asarray(mp_, float).T, np.asarray(ip_, float).T, shear=False, scale=False,
                                         usesvd=True)
        return True, M[:3, :3].copy(), M[:3, 3].copy()

    def projectPoints(self, mp_, R_, T_, K_, dist):
        pts_ = (R_ @ np.asarray(mp_, float).T).T + T_
        self.proj.append((len(self.proj), self.sample_solves, pts_))
        return pts_[:, None, :], None

    def Rodrigues(self, R_):
        return R_, None


def gen_path():
    """Reference functions executed from source (see the module docstring)."""
    import types

    import torch

    from oracle import pose_oracle as po

    rng = np.random.default_rng(23)
    out = {}
    # --- gate + de-normalisation (gdrn_evaluator.py:89-126)
    f = ref_functions("core/gdrn_modeling/gdrn_evaluator.py", ["get_img_model_points_with_coords2d"])
    gate_ref = f["get_img_model_points_with_coords2d"]
    G = 6
    maskp = rng.uniform(0, 1, (G, 64, 64)).astype(np.float32)
    coor = rng.uniform(0, 1, (G, 64, 64, 3)).astype(np.float32)
    coor[:, :8] = 0.5  # rows whose de-normalised residual is exactly 0: must be gated out
    coor[:, 8:12, :, 1] = 0.5 + 5e-5  # |delta_y| = 5e-5 * extent < 1e-4 * extent: out as well
    extent = rng.uniform(0.05, 0.3, (G, 3)).astype(np.float32)
    c2d = rng.uniform(0, 1, (G, 64, 64, 2)).astype(np.float32)
    npts, mpts, ipts = [], [], []
    for i in range(G):
        ip, mp_ = gate_ref(None, maskp[i].copy(), coor[i].copy(), c2d[i].copy(), 480, 640, extent[i], -1, 0.5)
        npts.append(len(mp_))
        mpts.append(mp_)
        ipts.append(ip)
    out.update(gate_mask=maskp, gate_coor=coor, gate_extent=extent, gate_c2d=c2d, gate_n=np.array(npts),
               gate_model_points=np.concatenate(mpts), gate_image_points=np.concatenate(ipts))
    # --- mask post-processing (engine_utils.py:118-136)
    gm = ref_functions("core/gdrn_modeling/engine_utils.py", ["get_out_mask"])["get_out_mask"]
    raw = rng.normal(0.3, 0.4, (5, 1, 64, 64)).astype(np.float32)
    for mode in ("L1", "BCE"):
        cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(CDPN=types.SimpleNamespace(
            ROT_HEAD=types.SimpleNamespace(MASK_LOSS_TYPE=mode))))
        out["mask_" + mode] = gm(cfg, torch.from_numpy(raw)).numpy()
    out["mask_raw"] = raw
    # --- back-projection, rigid apply (misc.py)
    mf = ref_functions("lib/pysixd/misc.py", ["backproject", "backproject_th", "transform_pts_Rt"])
    depth = rng.uniform(0.3, 1.5, (48, 64)).astype(np.float32)
    depth[rng.random((48, 64)) < 0.2] = 0
    K = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])
    out.update(bp_depth=depth, bp_K=K, bp_np=mf["backproject"](depth, K),
               bp_th=mf["backproject_th"](torch.from_numpy(depth), torch.from_numpy(K.astype(np.float32))).numpy())
    pts = rng.uniform(-0.2, 0.2, (50, 3))
    tf = _load("ref_transform", "lib/pysixd/transform.py")
    R = tf.random_rotation_matrix(rng.random(3))[:3, :3]
    t = rng.uniform(-1, 1, 3)
    out.update(rt_pts=pts, rt_R=R, rt_t=t, rt_out=mf["transform_pts_Rt"](pts, R, t))
    # --- re / te (pose_error.py:400-436)
    ef = ref_functions("lib/pysixd/pose_error.py", ["re", "te"])
    Rs = np.stack([tf.random_rotation_matrix(rng.random(3))[:3, :3] for _ in range(8)])
    Rs[1] = Rs[0]  # identical rotations: trace 3 (clamp branch)
    ts = rng.uniform(-1, 1, (8, 3))
    out.update(err_R=Rs, err_t=ts, err_re=np.array([ef["re"](Rs[i], Rs[(i + 1) % 8]) for i in range(8)]),
               err_te=np.array([ef["te"](ts[i], ts[(i + 1) % 8]) for i in range(8)]))
    # --- rot6d (rot_reps.py)
    rf = ref_functions("core/utils/rot_reps.py", ["normalize_vector", "cross_product", "ortho6d_to_mat_batch"],
                       env={"F": torch.nn.functional})  # rot_reps.py:6 import torch.nn.functional as F
    p6 = rng.normal(0, 1, (16, 6)).astype(np.float32)
    out.update(rot6d_in=p6, rot6d_out=rf["ortho6d_to_mat_batch"](torch.from_numpy(p6)).numpy())
    # --- allo -> ego and the test-time pose assembly; axangle2mat (transforms3d, absent) = Rodrigues from the oracle
    uf = ref_functions("core/utils/utils.py", ["allocentric_to_egocentric"], env={"axangle2mat": po.axangle2mat})
    pf = ref_functions("core/gdrn_modeling/models/pose_from_pred_centroid_z.py", ["pose_from_predictions_test"],
                       env={"allocentric_to_egocentric": uf["allocentric_to_egocentric"]})
    n = 12
    rots = np.stack([tf.random_rotation_matrix(rng.random(3))[:3, :3] for _ in range(n)]).astype(np.float32)
    cent = rng.uniform(-0.3, 0.3, (n, 2)).astype(np.float32)
    zv = rng.uniform(0.5, 1.5, (n, 1)).astype(np.float32)
    cams = np.tile(K.astype(np.float32)[None], (n, 1, 1))
    ctr = rng.uniform(100, 500, (n, 2)).astype(np.float32)
    rr = rng.uniform(0.2, 1.5, n).astype(np.float32)
    whs = rng.uniform(40, 200, (n, 2)).astype(np.float32)
    for zt in ("REL", "ABS"):
        ego, tr = pf["pose_from_predictions_test"](torch.from_numpy(rots), torch.from_numpy(cent), torch.from_numpy(zv),
                                                   torch.from_numpy(cams.copy()), torch.from_numpy(ctr), torch.from_numpy(rr),
                                                   torch.from_numpy(whs), is_allo=True, z_type=zt)
        out["assm_rot_" + zt], out["assm_trans_" + zt] = ego.numpy(), tr.numpy()
    out.update(assm_rots=rots, assm_cent=cent, assm_z=zv, assm_cams=cams, assm_ctr=ctr, assm_rr=rr, assm_whs=whs)
    # --- the two sibling heads: pose_from_pred.py:21-58 (rotation + translation given) and
    # pose_from_pred_centroid_z_abs.py:21-92 (absolute 2-D centre + absolute z), test branches, rotation matrices and
    # (unnormalised) quaternions.  RT_transform.quat_trans_to_pose_m (RT_transform.py:177-183) needs transforms3d's quat2mat
    # (absent): the oracle supplies it.
    rt = types.SimpleNamespace(quat_trans_to_pose_m=lambda q, t: np.hstack([po.quat2mat(q), np.asarray(t, np.float64).reshape(3, 1)]))
    env = {"allocentric_to_egocentric": uf["allocentric_to_egocentric"], "RT_transform": rt}
    p1 = ref_functions("core/gdrn_modeling/models/pose_from_pred.py", ["pose_from_predictions_test"], env=env)["pose_from_predictions_test"]
    p2 = ref_functions("core/gdrn_modeling/models/pose_from_pred_centroid_z_abs.py", ["pose_from_predictions_test"],
                       env=env)["pose_from_predictions_test"]
    rng2 = np.random.default_rng(231)  # own stream: the vectors generated further down keep their values
    quats = rng2.normal(0, 1, (n, 4)).astype(np.float32)  # unnormalised on purpose ("this allows unnormalized quat", :37)
    trans = np.stack([rng2.uniform(-0.3, 0.3, n), rng2.uniform(-0.2, 0.2, n), rng2.uniform(0.5, 1.5, n)], 1).astype(np.float32)
    cabs = rng2.uniform(100, 500, (n, 2)).astype(np.float32)
    for tag, rin in (("mat", rots), ("quat", quats)):
        ego, tr = p1(torch.from_numpy(rin), torch.from_numpy(trans), is_allo=True)
        out["pfp_rot_" + tag], out["pfp_trans_" + tag] = ego.numpy(), tr.numpy()
        ego, tr = p2(torch.from_numpy(rin), torch.from_numpy(cabs), torch.from_numpy(zv), torch.from_numpy(cams.copy()), is_allo=True)
        out["pfpabs_rot_" + tag], out["pfpabs_trans_" + tag] = ego.numpy(), tr.numpy()
    out.update(pfp_quats=quats, pfp_trans=trans, pfpabs_cent=cabs)
    # --- the loader's depth back-projection, data_loader.py:530-576 + the [:, ::4, ::4] of :625: the statements are
    # inline code of a detectron2-dependent method, so the source LINES are cut out and executed as they stand with
    # the local variables they expect.  NOTE numpy here (2.x, NEP 50) evaluates float32-array (op) np.float64-scalar
    # in float64, the reference's pinned numpy 1.23 in float32: the stored float32 result equals the float32-step
    # arithmetic of the oracle / kernels to within 1-2 float32 ulps, not bit for bit.
    import textwrap

    import cv2

    du = _load("ref_data_utils", "core/utils/data_utils.py")
    lines = open(os.path.join(REF, "core/gdrn_modeling/data_loader.py")).read().splitlines()
    block = textwrap.dedent("\n".join(lines[529:576]))  # 1-based lines 530..576
    assert block.lstrip().startswith("resize_ratio = out_res / scale") and "depth_xyz = np.concatenate" in block
    n = 6
    dimg = rng.uniform(0.4, 1.6, (480, 640)).astype(np.float32)
    dimg[rng.random((480, 640)) < 0.1] = 0
    Kc = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]], dtype=np.float32)
    centers = rng.uniform(150, 450, (n, 2)).astype(np.float32)
    scales = rng.uniform(60, 300, n).astype(np.float32)
    coord_2d = du.get_2d_coord_np(640, 480, low=0, high=1).transpose(1, 2, 0)
    xyz64, newK = [], []
    for i in range(n):
        env = dict(np=np, cv2=cv2, crop_resize_by_warp_affine=du.crop_resize_by_warp_affine, my_warp_affine=du.my_warp_affine,
                   out_res=64, input_res=256, scale=float(scales[i]), bbox_center=centers[i], depth_img=dimg, K=Kc,
                   coord_2d=coord_2d)
        exec(compile(block, "data_loader.py:530-576", "exec"), env)
        xyz64.append(env["depth_xyz"][:, ::4, ::4].astype("float32"))  # :624-627
        newK.append(env["newCameraK"])
    out.update(loader_depth_img=dimg, loader_K=Kc, loader_centers=centers, loader_scales=scales,
               loader_depth_xyz=np.stack(xyz64), loader_newK=np.stack(newK))
    # --- RANSAC loop rules: misc.pnp_ransac_custom (misc.py:58-142) executed from source with a cv2 SHIM that turns
    # its 2D-3D solver calls into 3D-3D ones (solvePnP -> the reference's own Kabsch, projectPoints -> rigid apply,
    # Rodrigues -> identity): the loop, its sampling, its strict '<' inlier rule and its adaptive stop run as written.
    # Recorded: inlier count of every iteration and the number of iterations executed.  (The loop's SELECTION rule --
    # lowest mean error over all points -- is deliberately not the composite's)
    Cv2Shim = lambda: _Cv2Shim(tf)  # noqa: E731

    cases = []
    for ci, (npt, out_frac, noise) in enumerate([(200, 0.1, 5e-4), (200, 0.45, 1e-3), (150, 0.7, 1e-3), (300, 0.3, 2e-3)]):
        shim = Cv2Shim()
        ransac = ref_functions("lib/pysixd/misc.py", ["pnp_ransac_custom"], env={"cv2": shim})["pnp_ransac_custom"]
        Rg = tf.random_rotation_matrix(rng.random(3))[:3, :3]
        tg = rng.uniform(-0.3, 0.3, 3) + np.array([0, 0, 0.9])
        mpts = rng.uniform(-0.1, 0.1, (npt, 3))
        cpts = (Rg @ mpts.T).T + tg + rng.normal(0, noise, (npt, 3))
        bad = rng.random(npt) < out_frac
        cpts[bad] += rng.uniform(-0.1, 0.1, (int(bad.sum()), 3))
        thr_r = 0.005
        np.random.seed(1000 + ci)
        pose_r = ransac(cpts, mpts, None, ransac_iter=100, ransac_min_iter=10, ransac_reprojErr=thr_r)
        # inlier count of iteration i = first projection recorded after the i-th sample solve
        counts_r, seen = [], 0
        for _, ns, pts_ in shim.proj:
            if ns > seen:
                counts_r.append(int((np.linalg.norm(pts_ - cpts, axis=1) < thr_r).sum()))
                seen = ns
        cases.append((mpts, cpts, np.array(counts_r), shim.sample_solves, pose_r))
        out["ransac%d_model" % ci], out["ransac%d_cam" % ci] = mpts, cpts
        out["ransac%d_counts" % ci], out["ransac%d_iters" % ci] = np.array(counts_r), np.array(shim.sample_solves)
        out["ransac%d_pose" % ci], out["ransac%d_seed" % ci] = pose_r, np.array(1000 + ci)
    out["ransac_thr"], out["ransac_cases"] = np.array(0.005), np.array(len(cases))
    np.savez_compressed(os.path.join(GOLD, "path_golden.npz"), **out)
    print("path_golden.npz", len(out), "gate n:", npts, "ransac iters:", [c[3] for c in cases])


def gen_fps():
    from oracle.fps import fps_indices_reference
    from rdpn6d_b200.synth import fps_cloud

    out = {}
    rng = np.random.default_rng(7)
    small = {
        "gauss_2000": (fps_cloud(2000, seed=1), 64),
        "lattice_1000": (np.stack(np.meshgrid(*[np.arange(10)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32), 200),
        "dups_40": (np.repeat(rng.standard_normal((10, 3)).astype(np.float32), 4, 0), 24),
        "n_eq_k_50": (rng.standard_normal((50, 3)).astype(np.float32), 50),
        "n_lt_k_10": (rng.standard_normal((10, 3)).astype(np.float32), 20),
        "cube_8": (np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], np.float32), 8),
        "single_1": (np.array([[0.5, -1.0, 2.0]], np.float32), 3),
    }
    for name, (pts, k) in small.items():
        out[name + "_pts"] = pts
        out[name + "_idx"] = fps_indices_reference(pts, k)
    # large clouds: store the generator seed only (rdpn6d_b200.synth.fps_cloud) + indices
    for n, k, seed in [(200_000, 64, 3), (1_000_000, 8, 0), (1_000_000, 64, 0), (1_000_000, 512, 0)]:
        out[f"seeded_{n}_{k}_{seed}_idx"] = fps_indices_reference(fps_cloud(n, seed=seed), k)
    np.savez_compressed(os.path.join(GOLD, "fps_golden.npz"), **out)
    print("fps_golden.npz", len(out))


def gen_kabsch(tf):
    rng = np.random.default_rng(11)
    out = {}
    # doctest literal transform.py:893-898 (2-D affine; pins the module itself)
    v0 = [[0, 1031, 1031, 0], [0, 0, 1600, 1600]]
    v1 = [[675, 826, 826, 677], [55, 52, 281, 277]]
    out["doctest_v0"], out["doctest_v1"] = np.array(v0, float), np.array(v1, float)
    out["doctest_M"] = tf.affine_matrix_from_points(v0, v1)
    cases = []
    for i, n in enumerate([3, 4, 10, 100, 2000]):
        R = tf.random_rotation_matrix(rng.random(3))[:3, :3]
        t = rng.uniform(-1, 1, 3)
        a = rng.uniform(-0.2, 0.2, (3, n))
        c = R @ a + t[:, None] + rng.normal(0, 1e-3, (3, n))
        cases.append(("rigid%d" % i, a, c, False))
    a = rng.uniform(-0.2, 0.2, (3, 50))
    c = np.diag([1, 1, -1.0]) @ a + rng.normal(0, 1e-3, (3, 50))  # mirrored set -> det<0 branch
    cases.append(("reflect", a, c, False))
    a = rng.uniform(-0.2, 0.2, (3, 3))
    c = np.diag([1, -1.0, 1]) @ a + 0.3  # flipped triangle
    cases.append(("reflect_tri", a, c, False))
    a = rng.uniform(-0.2, 0.2, (3, 200))
    R = tf.random_rotation_matrix(rng.random(3))[:3, :3]
    c = 1.7 * (R @ a) + np.array([[0.1], [0.2], [0.9]]) + rng.normal(0, 1e-3, (3, 200))
    cases.append(("umeyama", a, c, True))
    a = rng.uniform(-0.2, 0.2, (3, 60))
    a[2] = 0.0  # planar
    c = R @ a + 0.5
    cases.append(("planar", a, c, False))
    for name, a, c, sc in cases:
        out[name + "_v0"], out[name + "_v1"] = a, c
        out[name + "_M"] = tf.affine_matrix_from_points(a, c, shear=False, scale=sc, usesvd=True)
        out[name + "_Msup"] = tf.superimposition_matrix(np.asarray(a, np.float64), np.asarray(c, np.float64), scale=sc)
        out[name + "_scale"] = np.array(sc)
    out["case_names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(GOLD, "kabsch_golden.npz"), **out)
    print("kabsch_golden.npz", len(cases))


def gen_affine(du):
    rng = np.random.default_rng(13)
    centers = rng.uniform(50, 600, (32, 2))
    scales = rng.uniform(20, 640, 32)
    mats = np.stack([du.get_affine_transform(centers[i], float(scales[i]), 0, 256) for i in range(32)])
    mats64 = np.stack([du.get_affine_transform(centers[i], float(scales[i]), 0, 64) for i in range(32)])
    # float32-representable inputs (what the evaluator-side tensors hold: bbox_center.astype("float32"))
    c32 = centers.astype(np.float32).astype(np.float64)
    s32 = scales.astype(np.float32).astype(np.float64)
    mats32 = np.stack([du.get_affine_transform(c32[i], float(s32[i]), 0, 256) for i in range(32)])
    np.savez_compressed(os.path.join(GOLD, "affine_golden.npz"), centers=centers, scales=scales, A256=mats, A64=mats64,
                        centers32=c32, scales32=s32, A256_32=mats32)
    print("affine_golden.npz")


def gen_region(du):
    rng = np.random.default_rng(17)
    xyz = rng.uniform(-0.1, 0.1, (4, 64, 64, 3))
    xyz[:, :10] = 0.0  # background rows
    fps = rng.uniform(-0.1, 0.1, (4, 32, 3))
    ids, deltas = [], []
    for i in range(4):
        r, d = du.xyz_to_region(xyz[i], fps[i])
        ids.append(r)
        deltas.append(d)
    np.savez_compressed(os.path.join(GOLD, "region_golden.npz"), xyz=xyz, fps=fps, region=np.stack(ids), delta=np.stack(deltas))
    print("region_golden.npz")


def gen_pose(tf):
    """Composite outputs with the reference's Kabsch on every solve."""
    from oracle import pose_oracle as po
    from rdpn6d_b200 import synth

    def ref_kabsch(v0, v1, w=None, scale=False):
        assert w is None
        return tf.affine_matrix_from_points(v0, v1, shear=False, scale=scale, usesvd=True)

    def ref_hypothesis_poses(obj, cam, sel, hyp_idx):
        Rt, valid = _orig_hyp(obj, cam, sel, hyp_idx)  # validity rule is ours
        for h in np.nonzero(valid)[0]:
            a = obj[:, hyp_idx[h]].astype(np.float64)
            c = cam[:, hyp_idx[h]].astype(np.float64)
            M = tf.affine_matrix_from_points(a, c, shear=False, scale=False, usesvd=True)
            Rt[h] = M[:3, :4].astype(np.float32).reshape(12)
        return Rt, valid

    _orig_hyp, _orig_k = po.hypothesis_poses, po.kabsch
    po.hypothesis_poses, po.kabsch = ref_hypothesis_poses, ref_kabsch
    try:
        models = synth.make_models(4, 32, seed=5, n_symmetric=1)
        b = synth.make_batch(4, models=models, H=64, seed=20260101)
        thr = 0.005
        res = po.pose_solve_batch(b, b["hyp_idx"], thr)
    finally:
        po.hypothesis_poses, po.kabsch = _orig_hyp, _orig_k
    out = {k: v for k, v in b.items() if v is not None}
    out["thr"] = np.float32(thr)
    out["out_pose"] = np.stack([r["pose"] for r in res])
    out["out_ninl"] = np.array([r["n_inl"] for r in res], np.int32)
    out["out_status"] = np.array([r["status"] for r in res], np.int32)
    out["out_best_h"] = np.array([r["best_h"] for r in res], np.int32)
    out["out_nsel"] = np.array([r["n_sel"] for r in res], np.int32)
    out["out_counts"] = np.stack([r["counts"] for r in res])
    out["out_valid"] = np.stack([r["valid"] for r in res])
    out["out_Rt_hyp"] = np.stack([r["Rt_hyp"] for r in res])
    out["out_inlier_mask"] = np.stack([r["inlier_mask"] for r in res])
    out["out_sel"] = np.stack([r["s1"]["sel"] for r in res])
    out["out_cam"] = np.stack([r["s1"]["cam"] for r in res])
    np.savez_compressed(os.path.join(GOLD, "pose_golden.npz"), **out)
    print("pose_golden.npz", out["out_status"], out["out_ninl"])


def gen_ransac_roi(tf):
    """The reference's RANSAC loop on the correspondences of whole synthetic ROIs: misc.pnp_ransac_custom (misc.py:58-142)
    executed from source behind the 3D-3D cv2 shim on the gated (object, camera) pairs of each ROI -- 10 pairs per
    sample (misc.py:72,91), the reference's Kabsch on every solve, every point scored in float64, adaptive stop.
    Stored: the ROI planes, the sampled pixel sets of every iteration as hyp_idx [B,H,10] (-1 beyond the iteration at
    which the loop stopped), the loop's inlier count per iteration and the pose the function returns (ret_pose).  The solver is run on the same planes with the
    same index sets (sample_size = 10) and must reproduce those counts (tests/test_oracle_pose.py,
    tests/test_pose_solve_gpu.py)."""
    from oracle import pose_oracle as po
    from rdpn6d_b200 import synth

    B, HMAX, S, thr = 4, 128, 10, 0.005
    models = synth.make_models(4, 32, seed=9, n_symmetric=1)
    # two ROIs with 1 mm noise and 15 % outlier pixels (the adaptive rule stops the loop right after min_iter: either w
    # is high, or a contaminated sample scores w^10 < 1 ulp and k = -inf) and two with 4 mm noise and no outliers
    # (every sample scores a middling w, k stays huge: the loop runs all 100 iterations)
    b0 = synth.make_batch(2, models=models, H=8, seed=20260301, occlusion_max=0.3)
    b1 = synth.make_batch(2, models=models, H=8, seed=20260302, occlusion_max=0.3, outlier_frac=0.0, noise_sigma=0.004)
    b = {k: np.concatenate([b0[k], b1[k]]) for k in ("depth", "Kp", "coor", "mask", "extent", "region_idx", "anchors")}
    hyp = np.full((B, HMAX, S), -1, np.int32)
    counts = np.full((B, HMAX), -1, np.int32)
    iters = np.zeros(B, np.int32)
    ret_pose = np.zeros((B, 3, 4), np.float64)  # what the function RETURNS: the lowest-mean-error pose (misc.py:113-132, 139-142)
    for r in range(B):
        c = po.correspondences(b["depth"][r], b["Kp"][r], b["coor"][r], b["mask"][r], b["extent"][r], b["region_idx"][r],
                               b["anchors"][r])
        pix = np.nonzero(c["sel"])[0]
        mpts = c["obj"][:, pix].T.astype(np.float64)
        cpts = c["cam"][:, pix].T.astype(np.float64)
        shim = _Cv2Shim(tf)
        ransac = ref_functions("lib/pysixd/misc.py", ["pnp_ransac_custom"], env={"cv2": shim})["pnp_ransac_custom"]
        np.random.seed(2000 + r)
        ret_pose[r] = ransac(cpts, mpts, None, ransac_iter=100, ransac_min_iter=10, ransac_reprojErr=thr)
        cr, seen = [], 0
        for _, ns, pts_ in shim.proj:  # inlier count of iteration i = first projection after the i-th sample solve
            if ns > seen:
                cr.append(int((np.linalg.norm(pts_ - cpts, axis=1) < thr).sum()))
                seen = ns
        n_it = shim.sample_solves
        assert len(cr) == n_it <= HMAX
        # replay the loop's sampling calls (misc.py:91 is its only use of the generator) and check the replay
        np.random.seed(2000 + r)
        for i in range(n_it):
            idx = np.random.choice(len(pix), S, replace=False)
            M = tf.affine_matrix_from_points(mpts[idx].T, cpts[idx].T, shear=False, scale=False, usesvd=True)
            e = np.linalg.norm((M[:3, :3] @ mpts.T).T + M[:3, 3] - cpts, axis=1)
            assert int((e < thr).sum()) == cr[i], (r, i)
            hyp[r, i] = pix[idx]
        counts[r, :n_it] = cr
        iters[r] = n_it
    out = {k: b[k] for k in ("depth", "Kp", "coor", "mask", "extent", "region_idx", "anchors")}
    out.update(hyp_idx=hyp, counts=counts, iters=iters, thr=np.float32(thr), ret_pose=ret_pose)
    np.savez_compressed(os.path.join(GOLD, "ransac_roi_golden.npz"), **out)
    print("ransac_roi_golden.npz iters", iters, "max counts", counts.max(axis=1))


def gen_metrics():
    """Small helpers on the edges of the path, executed from the reference source: misc.backproject_v2 (misc.py:352-371),
    misc.calc_emb_bp_fast (:288-316), pose_error.adi (pose_error.py:315-337, scipy cKDTree), pose_utils.get_closest_rot
    (pose_utils.py:430-454)."""
    from scipy import spatial

    tf = _load("ref_transform", "lib/pysixd/transform.py")
    rng = np.random.default_rng(77)
    mf = ref_functions("lib/pysixd/misc.py", ["backproject_v2", "calc_emb_bp_fast", "transform_pts_Rt"])
    K = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]])
    depth = rng.uniform(0.4, 1.6, (48, 64)).astype(np.float32)
    depth[rng.random((48, 64)) < 0.2] = 0
    R = tf.random_rotation_matrix(rng.random(3))[:3, :3]
    T = rng.uniform(-0.2, 0.8, 3)
    out = dict(bp_depth=depth, bp_K=K, bp_v2=mf["backproject_v2"](depth, K), bp_R=R, bp_T=T, bp_emb=mf["calc_emb_bp_fast"](depth, R, T, K))
    pf = ref_functions("lib/pysixd/pose_error.py", ["adi", "add", "re", "te"], env={"spatial": spatial, "transform_pts_Rt": mf["transform_pts_Rt"],
                                                                      "misc": types_ns(transform_pts_Rt=mf["transform_pts_Rt"])})
    pts = (rng.standard_normal((700, 3)) * np.array([0.05, 0.04, 0.08])).astype(np.float32)
    Re, Rg = tf.random_rotation_matrix(rng.random(3))[:3, :3], tf.random_rotation_matrix(rng.random(3))[:3, :3]
    te_, tg = rng.uniform(-0.1, 0.1, (3, 1)) + np.array([[0], [0], [0.9]]), rng.uniform(-0.1, 0.1, (3, 1)) + np.array([[0], [0], [0.9]])
    out.update(adi_pts=pts, adi_Re=Re, adi_te=te_, adi_Rg=Rg, adi_tg=tg, adi_val=np.float64(pf["adi"](Re, te_, Rg, tg, pts.astype(np.float64))))
    out["add_val"] = np.float64(pf["add"](Re, te_, Rg, tg, pts.astype(np.float64)))
    out["re_val"] = np.float64(pf["re"](Re, Rg))
    out["te_val"] = np.float64(pf["te"](te_, tg))
    import torch

    gf = ref_functions("core/utils/pose_utils.py", ["get_closest_rot"], env={"re": pf["re"], "torch": torch})
    sym = np.stack([tf.rotation_matrix(a, [0, 0, 1])[:3, :3] for a in (np.pi / 2, np.pi, 3 * np.pi / 2)])
    est = Rg.dot(sym[1]).dot(tf.rotation_matrix(0.05, [1, 0, 0])[:3, :3])  # closest to the 180-degree copy
    out.update(gcr_est=est, gcr_gt=Rg, gcr_sym=sym, gcr_out=gf["get_closest_rot"](est, Rg, sym), gcr_out_none=gf["get_closest_rot"](est, Rg, None),
               gcr_out_single=gf["get_closest_rot"](est, Rg, sym[0]))
    # symmetry sets (misc.py:206-254) of a YCB-V-like model_info: two discrete symmetries and one continuous axis
    sf = ref_functions("lib/pysixd/misc.py", ["get_symmetry_transformations"], env={"transform": tf})["get_symmetry_transformations"]
    d1 = np.eye(4)
    d1[:3, :3] = tf.rotation_matrix(np.pi, [0, 1, 0])[:3, :3]
    d1[:3, 3] = [1.0, -2.0, 0.5]
    infos = [{"symmetries_discrete": [d1.reshape(-1).tolist()]},
             {"symmetries_continuous": [{"axis": [0, 0, 1], "offset": [0.5, -1.0, 2.0]}]},
             {"symmetries_discrete": [d1.reshape(-1).tolist()], "symmetries_continuous": [{"axis": [0, 1, 0], "offset": [0, 0, 0]}]},
             {}]
    import json

    out["sym_infos"] = np.array(json.dumps(infos))
    out["sym_step"] = np.float64(0.2)
    for i, info in enumerate(infos):
        tr = sf(info, 0.2)
        out["sym%d_R" % i] = np.stack([t["R"] for t in tr])
        out["sym%d_t" % i] = np.stack([t["t"] for t in tr])
    # batched rigid apply (misc.py:930-949, torch) on float64 inputs; drawn last so that the vectors above keep their values
    bf = ref_functions("lib/pysixd/misc.py", ["transform_pts_batch"])["transform_pts_batch"]
    bp = rng.standard_normal((3, 17, 3))
    bR = np.stack([tf.random_rotation_matrix(rng.random(3))[:3, :3] for _ in range(3)])
    bt = rng.standard_normal((3, 3, 1))
    out.update(tpb_pts=bp, tpb_R=bR, tpb_t=bt, tpb_out=bf(torch.from_numpy(bp), torch.from_numpy(bR), torch.from_numpy(bt)).numpy(),
               tpb_out_not=bf(torch.from_numpy(bp), torch.from_numpy(bR)).numpy())
    np.savez_compressed(os.path.join(GOLD, "metrics_golden.npz"), **out)
    print("metrics_golden.npz adi", out["adi_val"], "symmetry sets", [out["sym%d_R" % i].shape[0] for i in range(len(infos))])


def types_ns(**kw):
    import types

    return types.SimpleNamespace(**kw)


def gen_rows():
    """BOP result rows: GDRN_Evaluator.pose_prediction_to_json (gdrn_evaluator.py:483-513) executed from source with its
    helper to_list (test_utils.py:29-30); the evaluator hook must emit the same dicts."""
    import json
    import types

    fns = ref_functions("core/gdrn_modeling/test_utils.py", ["to_list"])
    to_json = ref_functions("core/gdrn_modeling/gdrn_evaluator.py", ["pose_prediction_to_json"], env=fns)["pose_prediction_to_json"]
    rng = np.random.default_rng(29)
    me = types.SimpleNamespace(cfg=None)
    tf = _load("ref_transform", "lib/pysixd/transform.py")
    cases = []
    for i in range(6):
        pose = np.concatenate([tf.random_rotation_matrix(rng.random(3))[:3, :3],
                               rng.uniform(-0.5, 1.5, (3, 1))], axis=1).astype(np.float32 if i % 2 else np.float64)
        kw = dict(scene_id=str(i + 1), im_id=100 + i, obj_id=i % 3 + 1)
        if i % 3:
            kw["score"] = float(rng.random())
        if i >= 3:
            kw["pose_time"] = float(rng.random())
        out = to_json(me, pose, **kw)
        cases.append(dict(pose=pose.astype(np.float64).tolist(), pose_dtype=str(pose.dtype), kwargs=kw, rows=out))
    with open(os.path.join(GOLD, "rows_golden.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("rows_golden.json", len(cases))


def gen_fps_center():
    """a12: the reference's Python FPS surface executed from source -- fps_utils.farthest_point_sampling
    (core/csrc/fps/fps_utils.py:6-21, its cffi handles replaced by ctypes handles on the reference's own C++ build in
    oracle/_ref) and data_utils.get_fps_and_center (core/utils/data_utils.py:217-226) on top of it."""
    import ctypes
    import types

    from oracle import build as _obuild
    from rdpn6d_b200.synth import fps_cloud

    _obuild()
    cdll = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfps_ref.so"))
    for fn in (cdll.farthest_point_sampling_init_center, cdll.farthest_point_sampling):
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        fn.restype = None
    ffi = types.SimpleNamespace(cast=lambda ctype, addr: ctypes.c_void_p(addr))
    wrap = ref_functions("core/csrc/fps/fps_utils.py", ["farthest_point_sampling"], env={"ffi": ffi, "lib": cdll})
    mod = types.ModuleType("core.csrc.fps.fps_utils")
    mod.farthest_point_sampling = wrap["farthest_point_sampling"]
    saved = {k: sys.modules.get(k) for k in ("core", "core.csrc", "core.csrc.fps", "core.csrc.fps.fps_utils")}
    for k in ("core", "core.csrc", "core.csrc.fps"):
        sys.modules[k] = types.ModuleType(k)
    sys.modules["core.csrc.fps.fps_utils"] = mod
    try:
        get = ref_functions("core/utils/data_utils.py", ["get_fps_and_center"])["get_fps_and_center"]
        out = {}
        for name, cloud in (("f64", fps_cloud(3000, seed=6).astype(np.float64) * 1.000000123),
                            ("f32", fps_cloud(2000, seed=7))):
            out[name + "_pts"] = cloud
            for n in (8, 32):
                out["%s_fps%d_and_center" % (name, n)] = get(cloud, num_fps=n, init_center=True)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    np.savez_compressed(os.path.join(GOLD, "fps_center_golden.npz"), **out)
    print("fps_center_golden.npz", {k: (v.shape, str(v.dtype)) for k, v in out.items() if "and_center" in k})


def gen_sampler():
    """(e): the reference's shard rule -- class InferenceSampler (core/utils/my_distributed_sampler.py:170-199) executed
    from source with a stand-in for detectron2's comm (rank / world size): the index range of every rank."""
    import ast
    import json
    import textwrap
    import types

    rel = "core/utils/my_distributed_sampler.py"
    src = open(os.path.join(REF, rel)).read()
    node = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.ClassDef) and n.name == "InferenceSampler")
    state = {"rank": 0, "world": 1}
    comm = types.SimpleNamespace(get_rank=lambda: state["rank"], get_world_size=lambda: state["world"])
    ns = {"Sampler": object, "comm": comm}
    exec(compile(textwrap.dedent(ast.get_source_segment(src, node)), os.path.join(REF, rel), "exec"), ns)
    cases = []
    for size in (1, 2, 7, 8, 10, 1000, 1024, 8192, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            ranges = []
            for rank in range(world):
                state.update(rank=rank, world=world)
                idx = list(ns["InferenceSampler"](size))
                ranges.append([idx[0], idx[-1] + 1] if idx else None)
            cases.append([size, world, ranges])
    with open(os.path.join(GOLD, "sampler_golden.json"), "w") as f:
        json.dump(cases, f)
    print("sampler_golden.json", len(cases))


def gen_roi_scalars():
    """a1: the loader's ROI scalars -- data_loader.py:477-482 (bbox centre, bw / bh, padded scale clipped to the image)
    executed from their source lines, plus the resize ratio of :488 (out_res / scale)."""
    import json
    import textwrap
    import types

    lines = open(os.path.join(REF, "core/gdrn_modeling/data_loader.py")).read().splitlines()
    block = textwrap.dedent("\n".join(lines[476:482]))  # 1-based lines 477..482
    assert block.startswith("x1, y1, x2, y2 = bbox") and "DZI_PAD_SCALE" in block, block
    assert 'roi_infos["resize_ratio"].append(out_res / scale)' in lines[487]
    rng = np.random.default_rng(31)
    cases = []
    boxes = [rng.uniform(0, 400, 2).tolist() + (rng.uniform(0, 400, 2) + rng.uniform(5, 230, 2)).tolist() for _ in range(12)]
    boxes += [[10.0, 20.0, 10.4, 20.2], [0.0, 0.0, 640.0, 480.0], [100.0, 50.0, 100.0, 300.0], [5.5, 7.25, 600.0, 470.0]]
    for bb in boxes:
        for pad, (im_H, im_W) in ((1.5, (480, 640)), (1.0, (480, 640)), (1.5, (1080, 1920))):
            env = dict(np=np, bbox=np.array(bb), im_H=im_H, im_W=im_W,
                       cfg=types.SimpleNamespace(INPUT=types.SimpleNamespace(DZI_PAD_SCALE=pad)))
            exec(compile(block, "data_loader.py:477-482", "exec"), env)
            cases.append(dict(bbox=bb, pad=pad, im_H=im_H, im_W=im_W, center=[float(v) for v in env["bbox_center"]],
                              scale=float(env["scale"]), resize_ratio=float(64 / env["scale"]), wh=[float(env["bw"]), float(env["bh"])]))
    with open(os.path.join(GOLD, "roi_scalars_golden.json"), "w") as f:
        json.dump(cases, f)
    print("roi_scalars_golden.json", len(cases))


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not mounted at %s" % REF)
    os.makedirs(GOLD, exist_ok=True)
    tf = _load("ref_transform", "lib/pysixd/transform.py")
    du = _load("ref_data_utils", "core/utils/data_utils.py")
    gens = dict(fps=gen_fps, kabsch=lambda: gen_kabsch(tf), affine=lambda: gen_affine(du), region=lambda: gen_region(du),
                pose=lambda: gen_pose(tf), path=gen_path, ransac_roi=lambda: gen_ransac_roi(tf), rows=gen_rows, metrics=gen_metrics, fps_center=gen_fps_center, sampler=gen_sampler, roi_scalars=gen_roi_scalars)
    for name in (sys.argv[1:] or list(gens)):  # python -m oracle.gen_golden [name ...]
        gens[name]()


if __name__ == "__main__":
    main()
