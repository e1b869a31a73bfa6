"""GPU: batched Kabsch (B4), pose assembly (B3), evaluator hook (B2) and the NCCL gather (e)."""
import os

import numpy as np
import pytest
import torch

from oracle import pose_oracle as po
from rdpn6d_b200 import distributed as D
from rdpn6d_b200 import evaluator, geometry, pose_from_pred, pose_solver, synth

pytestmark = pytest.mark.gpu


def test_kabsch_vs_reference_golden(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "kabsch_golden.npz"))
    for name in g["case_names"]:
        a, c, M, sc = g[f"{name}_v0"], g[f"{name}_v1"], g[f"{name}_M"], bool(g[f"{name}_scale"])
        a32, c32 = a.astype(np.float32), c.astype(np.float32)
        Mo = po.kabsch(a32, c32, scale=sc)  # oracle on the float32-rounded inputs the kernel sees
        Mg, s = geometry.kabsch(torch.from_numpy(a32.T.copy())[None].cuda(), torch.from_numpy(c32.T.copy())[None].cuda(), scale=sc)
        Mg = Mg[0].cpu().numpy().astype(np.float64)
        lin_o, lin_g = Mo[:3, :3], Mg[:, :3]
        if sc:
            so = np.cbrt(np.linalg.det(lin_o))
            assert abs(float(s[0]) - so) < 1e-6 * so
            lin_o, lin_g = lin_o / so, lin_g / float(s[0])
        assert po.re_rad_small(lin_g, lin_o) < 1e-5, name
        assert po.te(Mg[:, 3], Mo[:3, 3]) < 1e-6, name
        if name != "planar":
            np.testing.assert_allclose(Mg, M[:3, :4], atol=5e-6)  # and close to the float64-input reference result
        sup = geometry.superimposition_matrix(torch.from_numpy(a32).cuda(), torch.from_numpy(c32).cuda(), scale=sc).cpu().numpy()
        np.testing.assert_allclose(sup[:3, :4], Mg, atol=1e-6)


def test_kabsch_batched_weighted_and_errors(cuda):
    rng = np.random.default_rng(0)
    B, N = 37, 777
    a = rng.uniform(-0.2, 0.2, (B, N, 3)).astype(np.float32)
    c = rng.uniform(-0.2, 0.2, (B, N, 3)).astype(np.float32) + 0.8
    w = rng.uniform(0.1, 1, (B, N)).astype(np.float32)
    M, _ = geometry.kabsch(torch.from_numpy(a).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(w).cuda())
    M = M.cpu().numpy()
    for i in range(0, B, 6):
        Mo = po.kabsch(a[i].T, c[i].T, w=w[i])
        assert po.re_rad_small(M[i][:, :3], Mo[:3, :3]) < 1e-5
        assert po.te(M[i][:, 3], Mo[:3, 3]) < 1e-6
    with pytest.raises(ValueError):
        geometry.kabsch(torch.zeros(1, 2, 3, device="cuda"), torch.zeros(1, 2, 3, device="cuda"))  # transform.py:917-918


def test_pose_from_pred_centroid_z(cuda):
    rng = np.random.default_rng(1)
    B = 64
    r6 = rng.standard_normal((B, 6)).astype(np.float32)
    Rm = po.ortho6d_to_mat(r6)
    cen = rng.uniform(-0.3, 0.3, (B, 2)).astype(np.float32)
    z = rng.uniform(0.5, 3, (B, 1)).astype(np.float32)
    K = np.repeat(synth.K_LM[None].astype(np.float32), B, 0)
    ctr = rng.uniform(100, 500, (B, 2)).astype(np.float32)
    rr = rng.uniform(0.2, 0.9, B).astype(np.float32)
    wh = rng.uniform(30, 200, (B, 2)).astype(np.float32)
    cen[0] = 0
    ctr[0] = [K[0, 0, 2], K[0, 1, 2]]  # object exactly on the optical axis -> angle 0 branch (utils.py:66)
    tc = lambda x: torch.from_numpy(x).cuda()
    for is_allo in (True, False):
        for zt in ("REL", "ABS"):
            ro, to = po.pose_from_pred_centroid_z_test(Rm, cen, z, K, ctr, rr, wh, is_allo=is_allo, z_type=zt)
            rg, tg = pose_from_pred.pose_from_pred_centroid_z(tc(Rm), tc(cen), tc(z), tc(K), tc(ctr), tc(rr), tc(wh),
                                                              is_allo=is_allo, z_type=zt, is_train=False)
            assert rg.is_cuda  # documented deviation: stays on the device (reference returns a CPU tensor)
            assert np.array_equal(tg.cpu().numpy().view(np.uint32), to.view(np.uint32))  # translation bit-exact
            np.testing.assert_allclose(rg.cpu().numpy(), ro, atol=2e-6)
    r2, _ = pose_from_pred.pose_from_pred_centroid_z(tc(r6), tc(cen), tc(z), tc(K), tc(ctr), tc(rr), tc(wh), is_allo=False)
    np.testing.assert_allclose(r2.cpu().numpy(), Rm, atol=1e-6)  # rot6d path (rot_reps.py:34-49)
    with pytest.raises(ValueError):
        pose_from_pred.pose_from_pred_centroid_z(tc(Rm), tc(cen), tc(z), tc(K), tc(ctr), tc(rr), tc(wh), z_type="X")
    with pytest.raises(NotImplementedError):
        pose_from_pred.pose_from_pred_centroid_z(tc(Rm), tc(cen), tc(z), tc(K), tc(ctr), tc(rr), tc(wh), is_train=True)


def test_evaluator_hook_end_to_end(cuda):
    """GpuRansacKabsch.process on reference-shaped inputs (2 images x 3 ROIs) recovers the poses and
    emits BOP rows (R row-major, t in mm)."""
    b = synth.make_batch(6, H=8, seed=31)
    R_ = 32
    region = torch.full((6, R_ + 1, 64, 64), -5.0)
    region.scatter_(1, (torch.from_numpy(b["region_idx"]).long() + 1)[:, None], 5.0)
    q = np.stack([po.backproject_roi(b["depth"][i], b["Kp"][i], depth_div=b["resize_ratio"][i]) for i in range(6)])
    coord2d = np.concatenate([q, np.zeros((6, 2, 64, 64), np.float32)], 1)  # data_loader.py:624-625 layout
    inputs, outputs = [], []
    for im in range(2):
        sl = slice(3 * im, 3 * im + 3)
        inputs.append(dict(roi_img=[0, 1, 2], roi_coord_2d=torch.from_numpy(coord2d[sl]), cam=torch.from_numpy(b["K"][sl]),
                           roi_extent=torch.from_numpy(b["extent"][sl]), bbox_center=torch.from_numpy(b["bbox_center"][sl]),
                           scale=torch.from_numpy(b["scale"][sl]), resize_ratio=torch.from_numpy(b["resize_ratio"][sl]),
                           roi_cls=[0, 1, 2], score=[1.0, 1.0, 1.0], scene_im_id=["2/%d" % (10 + im)] * 3,
                           fps=torch.from_numpy(b["anchors"][sl])))
        outputs.append({"time": 0.0})
    tc = lambda x: torch.from_numpy(x).cuda()
    out_dict = dict(coor_x=tc(b["coor"][:, 0:1]), coor_y=tc(b["coor"][:, 1:2]), coor_z=tc(b["coor"][:, 2:3]),
                    mask=tc(b["mask"][:, None]), region=region.cuda(), trans=tc(b["gt_pose"][:, :, 3].astype(np.float32)),
                    rot=torch.from_numpy(b["gt_pose"][:, :, :3].astype(np.float32)))
    ev = evaluator.GpuRansacKabsch(num_hyp=128, inlier_thr=0.005)
    rows = ev.process(inputs, outputs, out_dict)
    assert len(rows) == 6 and len(ev._predictions) == 6
    for i, r in enumerate(rows):
        R = np.array(r["R"]).reshape(3, 3)
        t = np.array(r["t"]) / 1000.0
        assert r["scene_id"] == "2" and r["im_id"] in (10, 11) and r["obj_id"] == (i % 3) + 1
        assert po.re_rad_small(R, b["gt_pose"][i][:, :3]) < 0.03
        assert po.te(t, b["gt_pose"][i][:, 3]) < 0.003
        assert r["time"] > 0
    # the reference loop's sampling (10 pairs per sample, misc.py:72) and early stop (misc.py:134-138) through the hook
    ev10 = evaluator.GpuRansacKabsch(num_hyp=128, inlier_thr=0.005, sample_size=10, adaptive=True)
    rows10 = ev10.process(inputs, [{"time": 0.0}, {"time": 0.0}], out_dict)
    assert len(rows10) == 6
    for i, r in enumerate(rows10):
        assert po.re_rad_small(np.array(r["R"]).reshape(3, 3), b["gt_pose"][i][:, :3]) < 0.03
        assert po.te(np.array(r["t"]) / 1000.0, b["gt_pose"][i][:, 3]) < 0.003


def test_gather_rows_single_gpu_identity(cuda):
    x = torch.randn(7, 16, device="cuda")
    assert D.gather_rows(x, 7) is x


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_sharded_solve_two_gpus(cuda):
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(root, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "NCCL_GATHER_OK" in r.stdout


def _coor_feat_torch(coor_x, coor_y, coor_z, roi_coord_2d, region, fps, mask, region_attention, mask_attention, mask_mode):
    """The reference's own op sequence: GDRN.py:199-222 + conv_pnp_net.py:128-136 + model_utils.py:24-42."""
    import torch.nn.functional as F

    coor_feat = torch.cat([coor_x, coor_y, coor_z], dim=1)
    coor_feat = torch.cat([coor_feat, roi_coord_2d], dim=1)
    region_softmax = F.softmax(region[:, 1:, :, :], dim=1)
    am = torch.argmax(region_softmax.reshape(region_softmax.shape[0], region_softmax.shape[1], -1), dim=1).unsqueeze(2)
    region_fps = torch.gather(fps.unsqueeze(1).expand(-1, am.shape[1], -1, -1), 2, am.unsqueeze(3).expand(-1, -1, -1, 3))
    region_fps = region_fps.squeeze(2).reshape(region_fps.shape[0], 64, 64, 3).permute(0, 3, 1, 2)
    coor_feat = torch.cat([coor_feat, region_fps], dim=1)
    x = torch.cat([coor_feat, region_softmax], dim=1) if region_attention else coor_feat
    if mask_attention != "none":
        bs = mask.shape[0]
        if mask_mode == "l1":
            mmax = torch.max(mask.view(bs, -1), dim=-1)[0].view(bs, 1, 1, 1)
            mmin = torch.min(mask.view(bs, -1), dim=-1)[0].view(bs, 1, 1, 1)
            mp = (mask - mmin) / (mmax - mmin)
        else:
            mp = torch.sigmoid(mask)
        x = x * mp if mask_attention == "mul" else torch.cat([x, mp], dim=1)
    return x


@pytest.mark.parametrize("R,ra,ma,mm", [(32, True, "mul", "l1"), (64, True, "mul", "l1"), (32, False, "none", "l1"),
                                        (20, True, "concat", "bce"), (32, True, "mul", "bce")])
def test_coor_feat_matches_reference_ops(cuda, R, ra, ma, mm):
    """f2: fused correspondence-feature assembly vs the reference's torch op sequence (43 channels at R=32)."""
    g = torch.Generator(device="cuda").manual_seed(R)
    B = 6
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    cx, cy, cz = torch.rand(B, 1, 64, 64, device="cuda", generator=g), torch.rand(B, 1, 64, 64, device="cuda", generator=g), torch.rand(B, 1, 64, 64, device="cuda", generator=g)
    c2d, region, fps, mask = rnd(B, 5, 64, 64), 3 * rnd(B, R + 1, 64, 64), 0.1 * rnd(B, R, 3), rnd(B, 1, 64, 64)
    region[:, 7] = region[:, 3]  # exact ties between two regions: the first maximum must win
    out = geometry.coor_feat(cx, cy, cz, c2d, region, fps, mask, mask_mode=mm, region_attention=ra, mask_attention=ma)
    ref = _coor_feat_torch(cx, cy, cz, c2d, region, fps, mask, ra, ma, mm)
    assert out.shape == ref.shape
    if R == 32 and ra and ma == "mul":
        assert out.shape[1] == 43  # nIn of ConvPnPNet (conv_pnp_net.py:73)
    # the anchor channels must be bit-identical except where the reference's softmax rounding merged a near-tie
    agree = (out[:, 8:11] == ref[:, 8:11]).all(dim=1).float().mean()
    assert agree > 0.999
    torch.testing.assert_close(out[:, :8], ref[:, :8], rtol=2e-6, atol=1e-7)
    if ra:
        torch.testing.assert_close(out[:, 11:11 + R], ref[:, 11:11 + R], rtol=1e-5, atol=2e-7)
    if ma == "concat":
        torch.testing.assert_close(out[:, -1], ref[:, -1], rtol=2e-6, atol=1e-7)


def test_xyz_to_region_vs_reference_golden(cuda, golden_dir):
    """f4: nearest-anchor region ids and residuals against golden outputs of the reference's
    core/utils/data_utils.xyz_to_region (scipy cdist + argmin)."""
    g = np.load(os.path.join(golden_dir, "region_golden.npz"))
    xyz32 = g["xyz"].astype(np.float32)
    fps32 = g["fps"].astype(np.float32)
    reg, delta = geometry.xyz_to_region(torch.from_numpy(xyz32).cuda(), torch.from_numpy(fps32).cuda())
    for i in range(xyz32.shape[0]):
        r_o, d_o = po.xyz_to_region(xyz32[i], fps32[i])  # oracle on the float32 inputs the kernel sees
        assert np.array_equal(reg[i].cpu().numpy(), r_o)
        assert np.array_equal(delta[i].cpu().numpy(), d_o)
        # and against the reference's own float64 run: ids identical except float32-rounding near-ties
        assert (reg[i].cpu().numpy() == g["region"][i]).mean() > 0.999
        np.testing.assert_allclose(delta[i].cpu().numpy()[reg[i].cpu().numpy() == g["region"][i]],
                                   g["delta"][i][reg[i].cpu().numpy() == g["region"][i]], atol=1e-7)
    assert (reg[:, :10] == 0).all()  # background rows


def test_fps_for_models_prefix_property(cuda):
    """tools/*/..._compute_fps.py loop: FPS(n) is a prefix of FPS(n_max), so one run serves every sample count."""
    from oracle.fps import get_fps_and_center
    from rdpn6d_b200 import fps_utils

    clouds = {1: synth.fps_cloud(4000, seed=1).astype(np.float64), 5: synth.fps_cloud(9000, seed=2).astype(np.float64)}
    out = fps_utils.fps_and_center_for_models(clouds, nums_fps=(2, 8, 32))
    for oid, pts in clouds.items():
        for n in (2, 8, 32):
            np.testing.assert_array_equal(out[str(oid)][f"fps{n}_and_center"], get_fps_and_center(pts, n))
