"""GPU: the kernels against outputs of the REFERENCE's own function bodies (tests/golden/path_golden.npz, produced by
oracle/gen_golden.py:gen_path): gate + de-normalisation, mask post-processing, back-projection, rot6d and the
test-time pose assembly."""
import os

import numpy as np
import pytest
import torch

from rdpn6d_b200 import geometry, pose_from_pred, pose_solver

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "path_golden.npz"))


def _cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_s1_gate_matches_reference_gate(cuda, g):
    """gdrn_evaluator.py:89-126 executed from source: same selected pixels, and the dense-mode object point of the
    S1 kernel is exactly the reference's de-normalised model point."""
    G = len(g["gate_n"])
    coor = g["gate_coor"].transpose(0, 3, 1, 2)  # [G,3,64,64]
    depth = np.ones((G, 64, 64), np.float32)  # the 3D-3D gate additionally wants a depth; give every pixel one
    Kp = np.tile(np.array([[500.0, 500.0, 128.0, 128.0]], np.float32), (G, 1))
    s1 = pose_solver.correspond(_cu(depth), _cu(Kp), _cu(coor[:, 0]), _cu(coor[:, 1]), _cu(coor[:, 2]), _cu(g["gate_mask"]),
                                _cu(g["gate_extent"]), mask_mode="raw", want_obj=True)
    sel = s1["sel"].reshape(G, 64, 64).cpu().numpy().astype(bool)
    obj = s1["obj"].reshape(G, 3, 64, 64).cpu().numpy()
    off = 0
    for i in range(G):
        n = int(g["gate_n"][i])
        assert int(sel[i].sum()) == n == int(s1["n_sel"][i])
        mine = obj[i].transpose(1, 2, 0)[sel[i]]
        assert np.array_equal(mine.view(np.uint32), g["gate_model_points"][off:off + n].view(np.uint32))
        off += n


def test_s1_mask_probability_matches_reference_get_out_mask(cuda, g):
    raw = g["mask_raw"][:, 0]
    B = raw.shape[0]
    depth = np.ones((B, 64, 64), np.float32)
    Kp = np.tile(np.array([[500.0, 500.0, 128.0, 128.0]], np.float32), (B, 1))
    half = np.full((B, 64, 64), 0.7, np.float32)
    ext = np.full((B, 3), 0.1, np.float32)
    for mode, key, exact in (("l1", "mask_L1", True), ("bce", "mask_BCE", False)):
        s1 = pose_solver.correspond(_cu(depth), _cu(Kp), _cu(half), _cu(half), _cu(half), _cu(raw), _cu(ext), mask_mode=mode)
        w = s1["w"].reshape(B, 64, 64).cpu().numpy()
        if exact:
            assert np.array_equal(w.view(np.uint32), g[key][:, 0].view(np.uint32))
        else:
            np.testing.assert_allclose(w, g[key][:, 0], rtol=3e-7, atol=0)  # torch.sigmoid vs 1 / (1 + expf(-m))
        assert np.array_equal(s1["sel"].reshape(B, 64, 64).cpu().numpy().astype(bool), g[key][:, 0] > 0.5)


def test_backproject_matches_reference_backproject_th(cuda, g):
    out = geometry.backproject_th(_cu(g["bp_depth"]), _cu(g["bp_K"].astype(np.float32))).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), g["bp_th"].view(np.uint32))


@pytest.mark.parametrize("zt", ["REL", "ABS"])
def test_pose_assembly_matches_reference(cuda, g, zt):
    rot, tr = pose_from_pred.pose_from_pred_centroid_z(_cu(g["assm_rots"]), _cu(g["assm_cent"]), _cu(g["assm_z"]),
                                                       _cu(g["assm_cams"]), _cu(g["assm_ctr"]), _cu(g["assm_rr"]),
                                                       _cu(g["assm_whs"]), is_allo=True, z_type=zt)
    assert np.array_equal(tr.cpu().numpy().view(np.uint32), g["assm_trans_" + zt].view(np.uint32))
    np.testing.assert_allclose(rot.cpu().numpy(), g["assm_rot_" + zt], rtol=0, atol=2e-6)


def test_rot6d_matches_reference(cuda, g):
    n = g["rot6d_in"].shape[0]
    z = torch.ones(n, 1, device="cuda")
    K = torch.eye(3, device="cuda")[None].repeat(n, 1, 1) * 500
    K[:, 2, 2] = 1
    rot, _ = pose_from_pred.pose_from_pred_centroid_z(_cu(g["rot6d_in"]), torch.zeros(n, 2, device="cuda"), z, K,
                                                      torch.zeros(n, 2, device="cuda"), torch.ones(n, device="cuda"),
                                                      torch.ones(n, 2, device="cuda"), is_allo=False, z_type="ABS")
    np.testing.assert_allclose(rot.cpu().numpy(), g["rot6d_out"], rtol=0, atol=5e-7)


def test_loader_backprojection_matches_reference_lines(cuda, g):
    """data_loader.py:530-576 + :625 executed from the source lines vs rdpn_roi_intrinsics + rdpn_roi_crop_depth +
    the S1 kernel (dense mode: cam = back-projected point).  Tolerance as in tests/test_oracle_path.py."""
    n = len(g["loader_scales"])
    K = _cu(np.tile(g["loader_K"][None], (n, 1, 1)))
    ctr, sc = _cu(g["loader_centers"]), _cu(g["loader_scales"])
    Kp = geometry.roi_intrinsics(K, ctr, sc)
    d64 = geometry.roi_crop_depth(_cu(g["loader_depth_img"])[None], ctr, sc)
    rr = 64.0 / sc
    half = torch.full((n, 64, 64), 0.7, device="cuda")
    s1 = pose_solver.correspond(d64, Kp, half, half, half, torch.ones(n, 64, 64, device="cuda"), torch.ones(n, 3, device="cuda"),
                                depth_div=rr, mask_mode="raw")
    cam = s1["cam"].reshape(n, 3, 64, 64).cpu().numpy()
    ref = g["loader_depth_xyz"]
    assert np.array_equal(cam[:, 2].view(np.uint32), ref[:, 2].view(np.uint32))
    assert (np.abs(cam[:, :2] - ref[:, :2]) <= 1.2e-7 * ref[:, 2][:, None]).all()


def test_sibling_heads_match_reference(cuda, g):
    """pose_from_pred (pose_from_pred.py:21-58) and pose_from_pred_centroid_z_abs (:21-92), test branches executed from
    source: rotation matrices and unnormalised quaternions, the whole batch in one kernel."""
    for tag, rin in (("mat", g["assm_rots"]), ("quat", g["pfp_quats"])):
        rot, tr = pose_from_pred.pose_from_pred(_cu(rin), _cu(g["pfp_trans"]), is_allo=True, is_train=False)
        np.testing.assert_allclose(rot.cpu().numpy(), g["pfp_rot_" + tag], rtol=0, atol=2e-6)
        assert np.array_equal(tr.cpu().numpy(), g["pfp_trans_" + tag])
        rot, tr = pose_from_pred.pose_from_pred_centroid_z_abs(_cu(rin), _cu(g["pfpabs_cent"]), _cu(g["assm_z"]), _cu(g["assm_cams"]),
                                                               is_allo=True, is_train=False)
        np.testing.assert_allclose(rot.cpu().numpy(), g["pfpabs_rot_" + tag], rtol=0, atol=2e-6)
        np.testing.assert_allclose(tr.cpu().numpy(), g["pfpabs_trans_" + tag], rtol=0, atol=1e-7)


def test_metric_helpers_match_reference_gpu(cuda, golden_dir):
    """geometry.backproject_v2 / calc_emb_bp_fast / adi / get_closest_rot against the reference functions run from source."""
    import os

    from rdpn6d_b200 import geometry

    m = np.load(os.path.join(golden_dir, "metrics_golden.npz"))
    d = _cu(m["bp_depth"])
    np.testing.assert_allclose(geometry.backproject_v2(d, m["bp_K"]).cpu().numpy(), m["bp_v2"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(geometry.calc_emb_bp_fast(d, m["bp_R"], m["bp_T"], m["bp_K"]).cpu().numpy(), m["bp_emb"], rtol=0, atol=1e-12)
    for _ in range(2):  # the scratch ticket is left at zero: a second call gives the same bits
        v = geometry.adi(m["adi_Re"], m["adi_te"], m["adi_Rg"], m["adi_tg"], _cu(m["adi_pts"]))
        assert abs(v - float(m["adi_val"])) <= 1e-12
        v = geometry.add(m["adi_Re"], m["adi_te"], m["adi_Rg"], m["adi_tg"], _cu(m["adi_pts"]))
        assert abs(v - float(m["add_val"])) <= 1e-12
    assert np.array_equal(geometry.get_closest_rot(m["gcr_est"], m["gcr_gt"], m["gcr_sym"]), m["gcr_out"])
    assert np.array_equal(geometry.get_closest_rot(m["gcr_est"], m["gcr_gt"], None), m["gcr_out_none"])
