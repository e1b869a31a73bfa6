"""CPU: the composite pose-path oracle (oracle/pose_oracle.py) against golden vectors produced by the
reference's own functions (oracle/gen_golden.py) and against analytic properties."""
import os

import numpy as np
import pytest

from oracle import pose_oracle as po
from rdpn6d_b200 import synth


@pytest.fixture(scope="module")
def gold(golden_dir):
    return {k: np.load(os.path.join(golden_dir, k + "_golden.npz")) for k in ("affine", "region", "pose")}


def test_roi_affine_closed_form_matches_reference(gold):
    """a1: the restated closed form (with the reference's float32 point roundings) vs golden matrices from
    core/utils/data_utils.get_affine_transform (cv2.getAffineTransform inside), float64 and float32 inputs."""
    g = gold["affine"]
    for ck, sk, ak in (("centers", "scales", "A256"), ("centers32", "scales32", "A256_32")):
        for i in range(len(g[sk])):
            A = po.roi_affine(g[ck][i], g[sk][i], 256)
            ref = g[ak][i]
            assert np.abs(A - ref).max() <= 1e-12 * np.abs(ref).max()
            K = synth.K_LM
            Kp = po.roi_intrinsics(K, g[ck][i], g[sk][i], 256)
            Kref = np.vstack([ref, [0, 0, 1]]) @ K  # data_loader.py:555-564
            np.testing.assert_allclose(Kp, [Kref[0, 0], Kref[1, 1], Kref[0, 2], Kref[1, 2]], rtol=1e-12)
            # and the naive s = crop/scale form is only ~1e-6 away (why the roundings are restated)
            s = 256.0 / g[sk][i]
            assert abs(A[0, 0] - s) / s < 1e-5


def test_xyz_to_region_matches_reference(gold):
    g = gold["region"]
    for i in range(g["xyz"].shape[0]):
        r, d = po.xyz_to_region(g["xyz"][i], g["fps"][i])
        assert np.array_equal(r, g["region"][i])
        np.testing.assert_allclose(d, g["delta"][i], atol=0)


def test_composite_matches_reference_kabsch_golden(gold):
    """Whole path on 4 ROIs: the oracle (numpy SVD) vs the run where every Kabsch call went through the
    reference's transform.affine_matrix_from_points."""
    g = gold["pose"]
    b = {k: g[k] for k in ("depth", "Kp", "coor", "mask", "extent", "region_idx", "anchors")}
    res = po.pose_solve_batch(b, g["hyp_idx"], float(g["thr"]))
    for i, r in enumerate(res):
        assert r["status"] == g["out_status"][i]
        assert r["n_sel"] == g["out_nsel"][i]
        assert np.array_equal(r["s1"]["sel"], g["out_sel"][i])
        assert np.array_equal(r["s1"]["cam"], g["out_cam"][i])
        assert np.array_equal(r["valid"], g["out_valid"][i])
        np.testing.assert_allclose(r["Rt_hyp"], g["out_Rt_hyp"][i], atol=2e-6)
        # counts may differ only where a hypothesis pose differs in its last float32 bit
        same = (r["Rt_hyp"] == g["out_Rt_hyp"][i]).all(axis=1)
        assert np.array_equal(r["counts"][same], g["out_counts"][i][same])
        assert same.mean() > 0.95
        assert r["best_h"] == g["out_best_h"][i]
        assert r["n_inl"] == g["out_ninl"][i]
        assert np.array_equal(r["inlier_mask"], g["out_inlier_mask"][i])
        assert po.re_rad_small(r["pose"][:, :3], g["out_pose"][i][:, :3]) < 1e-6
        assert po.te(r["pose"][:, 3], g["out_pose"][i][:, 3]) < 1e-7


def test_composite_recovers_ground_truth():
    b = synth.make_batch(6, H=128, seed=77)
    res = po.pose_solve_batch(b, b["hyp_idx"], 0.005)
    for i, r in enumerate(res):
        assert r["status"] == po.STATUS_OK
        assert po.re_rad_small(r["pose"][:, :3], b["gt_pose"][i][:, :3]) < 0.01
        assert po.te(r["pose"][:, 3], b["gt_pose"][i][:, 3]) < 0.001


def test_dense_mode_recovers_ground_truth():
    b = synth.make_batch(3, H=64, seed=5, dense=True)
    res = po.pose_solve_batch(b, b["hyp_idx"], 0.005)
    for i, r in enumerate(res):
        assert r["status"] == po.STATUS_OK
        assert po.re_rad_small(r["pose"][:, :3], b["gt_pose"][i][:, :3]) < 0.01


def test_backprojection_is_float32_and_matches_formula():
    rng = np.random.default_rng(0)
    d = rng.uniform(0.5, 1.5, (64, 64)).astype(np.float32)
    Kp = np.array([700.123, 701.5, 130.25, 126.75])
    q = po.backproject_roi(d, Kp)
    assert q.dtype == np.float32
    x64 = (4.0 * np.arange(64)[None, :] - np.float32(Kp[2])) * d.astype(np.float64) / np.float32(Kp[0])
    np.testing.assert_allclose(q[0], x64, rtol=3e-7)
    q2 = po.backproject_roi(d, Kp, depth_div=0.25)
    np.testing.assert_allclose(q2[2], d / np.float32(0.25), rtol=0)
    full = po.backproject(d, np.array([[500.0, 0, 32], [0, 500, 32], [0, 0, 1]]))
    assert full.shape == (64, 64, 3) and full.dtype == np.float32


def test_gate_rules():
    ext = np.array([0.1, 0.2, 0.3], np.float32)
    delta = np.zeros((3, 2, 2), np.float32)
    delta[:, 0, 0] = [0.01, 0.01, 0.01]
    delta[:, 0, 1] = [0.01, 0.01, 0.00002]  # below 1e-4 * 0.3
    delta[:, 1, 0] = [0.01, 0.01, 0.01]
    delta[:, 1, 1] = [0.01, 0.01, 0.01]
    mp = np.array([[0.9, 0.9], [0.5, 0.9]], np.float32)  # 0.5 is not > 0.5 (strict)
    z = np.array([[1.0, 1.0], [1.0, 0.0]], np.float32)  # last pixel has no depth
    sel = po.gate(mp, delta, ext, z)
    assert sel.tolist() == [[True, False], [False, False]]


def test_flat_mask_selects_nothing_and_reports_few_points():
    b = synth.make_batch(1, H=16, seed=3)
    b["mask"][:] = 0.7
    r = po.pose_solve_batch(b, b["hyp_idx"], 0.005)[0]
    assert r["n_sel"] == 0 and r["status"] == po.STATUS_FEW_POINTS
    assert (r["pose"] == -100).all()


def test_select_best_rules():
    counts = np.array([3, 10, 10, 12, 12, 2], np.int32)
    valid = np.ones(6, np.uint8)
    assert po.select_best(counts, valid, 100)[0] == 3  # strictly greater -> earliest maximum
    assert po.select_best(np.array([3, 3, 2]), np.ones(3), 100)[0] == -1  # below the >= 4 rule
    valid[3] = 0
    assert po.select_best(counts, valid, 100)[0] == 4
    # adaptive stop: 11 hypotheses with zero inliers -> k = -inf -> stop once i_ransac > min_iter (10)
    c = np.zeros(20, np.int32)
    c[15] = 50
    best, examined = po.select_best(c, np.ones(20), 100, adaptive=True)
    assert best == -1 and examined == 11


def test_sq_cut_equivalence():
    rng = np.random.default_rng(1)
    for thr in [0.005, 0.01, 1e-3, 0.123456]:
        cut = po.sq_cut(thr)
        x = np.concatenate([np.float32(thr) ** 2 * (1 + rng.uniform(-1e-6, 1e-6, 2000)), [cut, np.nextafter(cut, np.float32(0))]]).astype(np.float32)
        assert np.array_equal(np.sqrt(x) < np.float32(thr), x < cut)


def test_pose_assembly_allo_ego_roundtrip():
    """a9: allo->ego leaves R unchanged on the optical axis and rotates by the ray angle elsewhere."""
    R = po.axangle2mat([0.3, -0.5, 0.8], 0.7)
    np.testing.assert_allclose(po.allocentric_to_egocentric_mat(R, [0, 0, 1.0]), R, atol=1e-15)
    t = np.array([0.2, -0.1, 0.9])
    Re = po.allocentric_to_egocentric_mat(R, t)
    ang = np.arccos(t[2] / np.linalg.norm(t))
    assert abs(po.re_rad_small(Re, R) - ang) < 3e-7  # the reference normalises the object ray in float32 (path_golden)
    m = po.ortho6d_to_mat(np.array([[1, 0, 0, 0, 1, 0], [0.5, 0.1, -0.3, 0.2, 0.9, 0.4]], np.float32))
    np.testing.assert_allclose(m[0], np.eye(3), atol=1e-7)
    np.testing.assert_allclose(m[1] @ m[1].T, np.eye(3), atol=1e-6)
    rot, tr = po.pose_from_pred_centroid_z_test(m, np.array([[0.1, -0.2], [0, 0]], np.float32), np.array([[1.5], [2.0]], np.float32),
                                                synth.K_LM[None].repeat(2, 0), np.array([[300, 200], [320, 240]], np.float32),
                                                np.array([0.5, 0.4], np.float32), np.array([[50, 60], [70, 80]], np.float32))
    assert rot.shape == (2, 3, 3) and tr.shape == (2, 3)
    np.testing.assert_allclose(tr[:, 2], [0.75, 0.8], rtol=1e-6)


def test_roi_crop_restatement_matches_cv2():
    """f1: the numpy restatement of OpenCV's fixed-point bilinear warpAffine equals cv2 bit for bit at the
    pixels the loader keeps (incl. a crop hanging over the image border)."""
    pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    depth = rng.uniform(0.4, 1.6, (120, 160)).astype(np.float32)
    depth[30:60, 70:110] = 0
    for t in range(4):
        c = np.array([rng.uniform(10, 150), rng.uniform(10, 110)], np.float32)
        s = np.float32(rng.uniform(20, 160))
        if t == 0:
            c, s = np.array([3, 3], np.float32), np.float32(80)
        assert np.array_equal(po.roi_crop_depth(depth, c, s), po.roi_crop_depth_cv2(depth, c, s))


def test_sample_triplets_stream_properties():
    """The internal hypothesis sampling (hyp_idx == NULL): deterministic, drawn from the gated pixels only,
    different per ROI / seed, roughly uniform; empty gate -> all -1."""
    rng = np.random.default_rng(0)
    sel = rng.random(4096) < 0.1
    a = po.sample_triplets(sel, 256, 7, 3)
    assert a.shape == (256, 3) and a.dtype == np.int32
    assert np.array_equal(a, po.sample_triplets(sel, 256, 7, 3))
    assert sel[a.reshape(-1)].all()
    assert not np.array_equal(a, po.sample_triplets(sel, 256, 8, 3))
    assert not np.array_equal(a, po.sample_triplets(sel, 256, 7, 4))
    # the first H' hypotheses of a longer draw are the shorter draw (counter = 3 h + v)
    assert np.array_equal(a[:64], po.sample_triplets(sel, 64, 7, 3))
    big = po.sample_triplets(sel, 20000, 1, 0).reshape(-1)
    g = np.nonzero(sel)[0]
    counts = np.bincount(np.searchsorted(g, big), minlength=len(g))
    assert counts.min() > 0.5 * counts.mean() and counts.max() < 1.6 * counts.mean()
    assert (po.sample_triplets(np.zeros(4096, bool), 8, 1, 0) == -1).all()
    # known answer: pins the hash itself (fmix32 chain documented in include/rdpn6d_b200.h)
    one = np.zeros(4096, bool)
    one[[5, 77, 1000, 4095]] = True
    assert po.sample_triplets(one, 2, 42, 9).tolist() == [[77, 1000, 77], [5, 77, 77]]


def test_s10_hypotheses_reproduce_reference_ransac_loop(gold):
    """oracle/gen_golden.py:gen_ransac_roi ran misc.pnp_ransac_custom (misc.py:58-142) from source on the gated pairs of
    four synthetic ROIs (10 pairs per sample, misc.py:72,91).  The oracle on the same planes and pixel sets
    (hypothesis_poses with S = 10, float32 scoring) reproduces the loop's inlier count of every iteration up to
    boundary ties of the float32 scoring, and its adaptive rule stops where the loop stopped."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ransac_roi_golden.npz"))
    thr = float(g["thr"])
    exact = total = 0
    for r in range(4):
        c = po.correspondences(g["depth"][r], g["Kp"][r], g["coor"][r], g["mask"][r], g["extent"][r], g["region_idx"][r],
                               g["anchors"][r])
        n_it = int(g["iters"][r])
        hyp = g["hyp_idx"][r].copy()
        Rt, valid = po.hypothesis_poses(c["obj"], c["cam"], c["sel"], hyp)
        assert valid[:n_it].all() and not valid[n_it:].any()
        pix = np.nonzero(c["sel"])[0]
        counts = po.score_hypotheses(c["obj"][:, pix].T, c["cam"][:, pix].T, Rt, valid, thr)
        d = counts[:n_it].astype(int) - g["counts"][r, :n_it]
        assert np.abs(d).max() <= 2, (r, d)
        exact += int((d == 0).sum())
        total += n_it
        if n_it < 20:  # the loop stopped by its adaptive rule: with valid samples appended the oracle stops there too
            hyp[n_it:] = hyp[0]
            Rt, valid = po.hypothesis_poses(c["obj"], c["cam"], c["sel"], hyp)
            counts = po.score_hypotheses(c["obj"][:, pix].T, c["cam"][:, pix].T, Rt, valid, thr)
            _, examined = po.select_best(counts, valid, len(pix), 4, True, 0.995, 10)
            assert examined == n_it
    assert exact >= 0.97 * total


def test_min_mean_err_rule_reproduces_the_reference_loops_return_value():
    """What misc.pnp_ransac_custom RETURNS (misc.py:113-132, 139-142: the pose with the lowest mean error over all points
    among the sample fits and the refits of the hypotheses that raised the inlier count) is stored in the golden as
    ret_pose.  The oracle's select_rule="min_mean_err" on the same planes and pixel sets (10 pairs per sample, adaptive
    stop) returns that pose: <= 1e-6 rad / 1e-6 m (float32 hypothesis poses and scoring against the float64 loop)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ransac_roi_golden.npz"))
    b = {k: g[k] for k in ("depth", "Kp", "coor", "mask", "extent", "region_idx", "anchors")}
    res = po.pose_solve_batch(b, g["hyp_idx"], float(g["thr"]), select_rule="min_mean_err", adaptive=True, confidence=0.995,
                              min_iter=10)
    for r, o in enumerate(res):
        assert o["status"] == 0 and o["best_h"] < int(g["iters"][r])
        assert po.re_rad_small(o["pose"][:, :3], g["ret_pose"][r][:, :3]) <= 1e-6, r
        assert po.te(o["pose"][:, 3], g["ret_pose"][r][:, 3]) <= 1e-6, r
    # the rule is not the most-inliers rule: on these ROIs they pick different hypotheses at least once
    res_mi = po.pose_solve_batch(b, g["hyp_idx"], float(g["thr"]), adaptive=True, confidence=0.995, min_iter=10)
    assert any(o["best_h"] != m["best_h"] or not np.array_equal(o["pose"], m["pose"]) for o, m in zip(res, res_mi))


def test_s_pair_validity_rules():
    b = synth.make_batch(1, H=8, seed=5)
    c = po.correspondences(b["depth"][0], b["Kp"][0], b["coor"][0], b["mask"][0], b["extent"][0], b["region_idx"][0],
                           b["anchors"][0])
    pix = np.nonzero(c["sel"])[0]
    off = np.nonzero(~c["sel"])[0]
    good = pix[np.linspace(0, len(pix) - 1, 6).astype(int)]
    hyp = np.stack([good, np.r_[good[:5], good[0]], np.r_[good[:5], off[0]], np.r_[good[:5], -1]]).astype(np.int32)
    Rt, valid = po.hypothesis_poses(c["obj"], c["cam"], c["sel"], hyp)
    assert valid.tolist() == [1, 0, 0, 0]  # ok | repeated pixel | ungated pixel | out of range
    M = po.kabsch(c["obj"][:, good].astype(np.float64), c["cam"][:, good].astype(np.float64))
    assert np.allclose(Rt[0].reshape(3, 4), M[:3, :4], atol=1e-6)
    assert po.sample_triplets(c["sel"], 5, 1, 0, sample_size=7).shape == (5, 7)


def test_roi_scalars_match_loader_lines(golden_dir):
    """a1: data_loader.py:477-482 + :488 executed from their source lines (oracle/gen_golden.py:gen_roi_scalars), incl.
    boxes thinner than a pixel (bw = bh = 1) and boxes whose padded scale is clipped to the image."""
    import json

    cases = json.load(open(os.path.join(golden_dir, "roi_scalars_golden.json")))
    assert len(cases) >= 40
    for c in cases:
        center, scale, rr, wh = po.roi_scalars(c["bbox"], c["im_H"], c["im_W"], dzi_pad_scale=c["pad"])
        assert center.tolist() == c["center"] and scale == c["scale"] and rr == c["resize_ratio"] and wh.tolist() == c["wh"]
