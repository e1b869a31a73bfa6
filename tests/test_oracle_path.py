"""CPU: the oracle's restatements of the path pieces whose reference MODULES cannot be imported here, checked
against tests/golden/path_golden.npz -- outputs of the reference's own function bodies (cut out of the reference
files with `ast` and executed by oracle/gen_golden.py:gen_path)."""
import os

import numpy as np
import pytest

from oracle import pose_oracle as po


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "path_golden.npz"))


def test_gate_and_denormalisation_match_reference(g):
    """gdrn_evaluator.py:89-126: model points = de-normalised residual of the gated pixels, in raster order."""
    off = 0
    for i in range(len(g["gate_n"])):
        coor = np.ascontiguousarray(g["gate_coor"][i].transpose(2, 0, 1))  # HWC -> 3HW
        delta = po.denormalise_residual(coor, g["gate_extent"][i])
        sel = po.gate(g["gate_mask"][i], delta, g["gate_extent"][i], np.ones((64, 64), np.float32), 0.5)
        n = int(g["gate_n"][i])
        assert int(sel.sum()) == n
        mine = delta.transpose(1, 2, 0)[sel]
        assert np.array_equal(mine.view(np.uint32), g["gate_model_points"][off:off + n].view(np.uint32))
        # the rows constructed to sit exactly on / inside the |delta| > 1e-4 * extent rule are out
        assert not sel[:12].any()
        off += n


def test_mask_postprocessing_matches_reference(g):
    raw = g["mask_raw"]
    for b in range(raw.shape[0]):
        l1 = po.out_mask(raw[b, 0], po.MASK_L1)
        assert np.array_equal(l1.view(np.uint32), g["mask_L1"][b, 0].view(np.uint32))
        bce = po.out_mask(raw[b, 0], po.MASK_BCE)
        np.testing.assert_allclose(bce, g["mask_BCE"][b, 0], rtol=3e-7, atol=0)  # torch.sigmoid vs 1/(1+exp(-m)): 2 ulp


def test_backprojection_matches_reference(g):
    mine = po.backproject(g["bp_depth"], g["bp_K"])
    # torch float32 version (misc.py:334-349): same operation order -> bit-exact
    assert np.array_equal(mine.view(np.uint32), g["bp_th"].view(np.uint32))
    # numpy version computes in float64 (range() is int64, K float64): equal to float32 rounding
    np.testing.assert_allclose(mine, g["bp_np"], rtol=2e-7, atol=1e-9)


def test_rigid_apply_and_error_metrics_match_reference(g):
    out = po.transform_pts_Rt(g["rt_pts"], g["rt_R"], g["rt_t"])
    assert np.array_equal(out, g["rt_out"])
    for i in range(8):
        assert po.re(g["err_R"][i], g["err_R"][(i + 1) % 8]) == pytest.approx(float(g["err_re"][i]), abs=1e-12)
        assert po.te(g["err_t"][i], g["err_t"][(i + 1) % 8]) == pytest.approx(float(g["err_te"][i]), abs=1e-15)
    assert float(g["err_re"][0]) == 0.0  # identical rotations (clamp branch)


def test_rot6d_matches_reference(g):
    np.testing.assert_allclose(po.ortho6d_to_mat(g["rot6d_in"]), g["rot6d_out"], rtol=0, atol=3e-7)


@pytest.mark.parametrize("zt", ["REL", "ABS"])
def test_pose_assembly_matches_reference(g, zt):
    """pose_from_pred_centroid_z.py:52-141 + utils.py:39-94 (allo -> ego)."""
    rot, tr = po.pose_from_pred_centroid_z_test(g["assm_rots"], g["assm_cent"], g["assm_z"], g["assm_cams"], g["assm_ctr"],
                                                g["assm_rr"], g["assm_whs"], is_allo=True, z_type=zt)
    assert np.array_equal(tr.view(np.uint32), g["assm_trans_" + zt].view(np.uint32))
    np.testing.assert_allclose(rot, g["assm_rot_" + zt], rtol=0, atol=1.5e-7)


def test_loader_backprojection_matches_reference_lines(g):
    """data_loader.py:530-576 + :625 executed from the source lines: crop intrinsics K' = A K, bilinear depth crop
    / resize_ratio, (x - cx') * d / fx' at the kept pixels.  Depth is bit-exact; X / Y agree to one float32 ulp of
    the coordinate (the reference evaluates this formula in float32 under its pinned numpy 1.23 -- as the oracle
    and the kernels do -- but in float64 under the numpy 2 that generated the vectors)."""
    for i in range(len(g["loader_scales"])):
        c, sc = g["loader_centers"][i], float(g["loader_scales"][i])
        Kp = po.roi_intrinsics(g["loader_K"], c, sc, 256)
        nk = g["loader_newK"][i]
        np.testing.assert_allclose(Kp, [nk[0, 0], nk[1, 1], nk[0, 2], nk[1, 2]], rtol=1e-12)
        d64 = po.roi_crop_depth(g["loader_depth_img"], c, sc, 256, 64)
        q = po.backproject_roi(d64, Kp, depth_div=np.float32(64 / sc), stride=4)
        ref = g["loader_depth_xyz"][i]
        assert np.array_equal(q[2].view(np.uint32), ref[2].view(np.uint32))
        assert (np.abs(q[:2] - ref[:2]) <= 1.2e-7 * ref[2][None]).all()


def test_ransac_loop_rules_match_reference_loop(g):
    """misc.pnp_ransac_custom (misc.py:58-142) executed from source behind a cv2 shim that makes its solver calls
    3D-3D (see oracle/gen_golden.py): the oracle reproduces (a) the inlier count of every iteration -- same sampling
    call, all points scored, strict '<' -- and (b) the iteration at which the adaptive rule (:134-138) stops the loop,
    including the reference's quirk that a sample with w^10 below 1 ulp gives k = -inf and ends the loop at
    min_iter + 1."""
    thr = float(g["ransac_thr"])
    for ci in range(int(g["ransac_cases"])):
        mpts, cpts = g["ransac%d_model" % ci], g["ransac%d_cam" % ci]
        counts_ref, iters = g["ransac%d_counts" % ci], int(g["ransac%d_iters" % ci])
        assert len(counts_ref) == iters
        n = len(mpts)
        np.random.seed(int(g["ransac%d_seed" % ci]))
        counts = []
        for _ in range(iters):
            idx = np.random.choice(n, 10, replace=False)  # misc.py:91
            M = po.kabsch(mpts[idx].T, cpts[idx].T)
            errs = np.linalg.norm(po.transform_pts_Rt(mpts, M[:3, :3], M[:3, 3]) - cpts, axis=1)  # :108-109
            counts.append(int((errs < thr).sum()))  # :111
        assert counts == counts_ref.tolist()
        # the stop rule on a longer sequence: the oracle must stop by itself where the reference loop did
        longer = np.concatenate([counts_ref, np.full(40, counts_ref.max())])
        best, examined = po.select_best(longer, np.ones(len(longer), np.uint8), n, min_inliers=4, adaptive=True,
                                        confidence=0.995, min_iter=10)
        assert examined == iters
        assert best == int(np.argmax(counts_ref[:iters])) or counts_ref[:iters].max() < 4


def test_bop_rows_match_reference_pose_prediction_to_json(golden_dir):
    """The evaluator hook's result rows against GDRN_Evaluator.pose_prediction_to_json (gdrn_evaluator.py:483-513)
    executed from source (oracle/gen_golden.py:gen_rows): same keys, R row-major, t in millimetres, same float values."""
    import json

    from rdpn6d_b200 import evaluator

    cases = json.load(open(os.path.join(golden_dir, "rows_golden.json")))
    assert len(cases) >= 6
    for c in cases:
        pose = np.array(c["pose"], np.float64).astype(c["pose_dtype"])
        assert evaluator.pose_prediction_to_json(pose, **c["kwargs"]) == c["rows"]


def test_metric_helpers_match_reference(golden_dir):
    """misc.backproject_v2 / calc_emb_bp_fast (misc.py:288-371), pose_error.adi (pose_error.py:315-337, cKDTree) and
    pose_utils.get_closest_rot (pose_utils.py:430-454) executed from source (oracle/gen_golden.py:gen_metrics)."""
    m = np.load(os.path.join(golden_dir, "metrics_golden.npz"))
    np.testing.assert_allclose(po.backproject_v2(m["bp_depth"], m["bp_K"]), m["bp_v2"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(po.calc_emb_bp_fast(m["bp_depth"], m["bp_R"], m["bp_T"], m["bp_K"]), m["bp_emb"], rtol=0, atol=1e-12)
    assert abs(po.adi(m["adi_Re"], m["adi_te"], m["adi_Rg"], m["adi_tg"], m["adi_pts"]) - float(m["adi_val"])) <= 1e-12
    assert abs(po.add(m["adi_Re"], m["adi_te"], m["adi_Rg"], m["adi_tg"], m["adi_pts"].astype(np.float64)) - float(m["add_val"])) <= 1e-12
    from rdpn6d_b200 import geometry

    for mod in (po, geometry):  # re / te: 3 x 3 host arithmetic in both
        assert abs(mod.re(m["adi_Re"], m["adi_Rg"]) - float(m["re_val"])) <= 1e-12
        assert abs(mod.te(m["adi_te"], m["adi_tg"]) - float(m["te_val"])) <= 1e-15
    assert geometry.re(m["adi_Rg"], m["adi_Rg"]) == 0.0
    assert np.array_equal(po.get_closest_rot(m["gcr_est"], m["gcr_gt"], m["gcr_sym"]), m["gcr_out"])
    assert np.array_equal(po.get_closest_rot(m["gcr_est"], m["gcr_gt"], None), m["gcr_out_none"])
    assert np.array_equal(po.get_closest_rot(m["gcr_est"], m["gcr_gt"], m["gcr_sym"][0]), m["gcr_out_single"])
    assert not np.array_equal(m["gcr_out"], m["gcr_gt"])  # a symmetric copy really was closer
    # misc.transform_pts_batch (misc.py:930-949): the batched rigid apply, with and without the translation
    np.testing.assert_allclose(po.transform_pts_batch(m["tpb_pts"], m["tpb_R"], m["tpb_t"]), m["tpb_out"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(po.transform_pts_batch(m["tpb_pts"], m["tpb_R"]), m["tpb_out_not"], rtol=0, atol=1e-15)


def test_quat2mat_and_sibling_heads_match_reference(g):
    """pose_from_pred.py:21-58 and pose_from_pred_centroid_z_abs.py:21-92 (test branches) executed from source on rotation
    matrices and on unnormalised quaternions: the oracle's allo -> ego + quat2mat composition gives the same rotations."""
    n = g["pfp_trans"].shape[0]
    for tag, rin in (("mat", g["assm_rots"]), ("quat", g["pfp_quats"])):
        for key, trans in (("pfp", g["pfp_trans"]), ("pfpabs", g["pfpabs_trans_" + tag])):
            for i in range(n):
                R = rin[i].astype(np.float64) if tag == "mat" else po.quat2mat(rin[i])
                ego = po.allocentric_to_egocentric_mat(R, trans[i], np.float32 if tag == "mat" else np.float64)
                np.testing.assert_allclose(ego, g[key + "_rot_" + tag][i], rtol=0, atol=1e-7)  # float32 outputs; the ray is normalised in the dtype of the pose the reference function receives
    cams = g["assm_cams"]
    z = g["assm_z"].reshape(-1)
    c = g["pfpabs_cent"]
    t = np.stack([z * (c[:, 0] - cams[:, 0, 2]) / cams[:, 0, 0], z * (c[:, 1] - cams[:, 1, 2]) / cams[:, 1, 1], z], 1)
    np.testing.assert_allclose(t, g["pfpabs_trans_mat"], rtol=0, atol=1e-7)


def test_symmetry_sets_match_reference(golden_dir):
    """misc.get_symmetry_transformations (misc.py:206-254) executed from source on model_info records with discrete,
    continuous, both and no symmetries: the oracle's restatement and the package's host helper return the same sets."""
    import json

    from rdpn6d_b200 import geometry

    m = np.load(os.path.join(golden_dir, "metrics_golden.npz"))
    infos = json.loads(str(m["sym_infos"]))
    for i, info in enumerate(infos):
        # the oracle follows the reference's operation order (1e-15); the package builds the rotations in one batch (1e-12)
        for fn, tol in ((po.get_symmetry_transformations, 1e-15), (geometry.get_symmetry_transformations, 1e-12)):
            tr = fn(info, float(m["sym_step"]))
            assert len(tr) == m["sym%d_R" % i].shape[0]
            np.testing.assert_allclose(np.stack([t["R"] for t in tr]), m["sym%d_R" % i], rtol=0, atol=tol)
            np.testing.assert_allclose(np.stack([t["t"] for t in tr]), m["sym%d_t" % i], rtol=0, atol=tol)
