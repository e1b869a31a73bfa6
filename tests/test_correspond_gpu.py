"""GPU: stage S1 (fused back-projection + residual + mask gate) bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from oracle import pose_oracle as po
from rdpn6d_b200 import geometry, pose_solver, synth

pytestmark = pytest.mark.gpu


def _to_cuda(b):
    return {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in b.items()}


def _s1_oracle(b, i, mask_mode, depth_div=None, mask_thr=0.5):
    return po.correspondences(b["depth"][i], b["Kp"][i], b["coor"][i], b["mask"][i], b["extent"][i],
                              None if b["region_idx"] is None else b["region_idx"][i],
                              None if b["anchors"] is None else b["anchors"][i],
                              None if depth_div is None else depth_div[i], mask_mode=mask_mode, mask_thr=mask_thr)


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("mask_mode", [po.MASK_L1, po.MASK_RAW])
def test_s1_bit_exact(cuda, dense, mask_mode):
    b = synth.make_batch(12, H=8, seed=321, dense=dense, occlusion_max=0.5)
    g = _to_cuda(b)
    out = pose_solver.correspond(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"],
                                 g["extent"], g["region_idx"], g["anchors"], mask_mode=mask_mode)
    for i in range(12):
        o = _s1_oracle(b, i, mask_mode)
        assert np.array_equal(out["cam"][i].cpu().numpy().view(np.uint32), o["cam"].view(np.uint32)), i
        assert np.array_equal(out["obj"][i].cpu().numpy().view(np.uint32), o["obj"].view(np.uint32)), i
        assert np.array_equal(out["w"][i].cpu().numpy().view(np.uint32), o["w"].view(np.uint32)), i
        assert np.array_equal(out["sel"][i].cpu().numpy().astype(bool), o["sel"]), i
        assert int(out["n_sel"][i]) == int(o["sel"].sum())


def test_s1_depth_div_and_threshold(cuda):
    b = synth.make_batch(4, H=8, seed=9)
    g = _to_cuda(b)
    out = pose_solver.correspond(g["depth"], g["Kp"], g["coor"][:, 0:1], g["coor"][:, 1:2], g["coor"][:, 2:3],
                                 g["mask"][:, None], g["extent"], g["region_idx"], g["anchors"],
                                 depth_div=g["resize_ratio"], mask_thr=0.7)
    for i in range(4):
        o = _s1_oracle(b, i, po.MASK_L1, depth_div=b["resize_ratio"], mask_thr=0.7)
        assert np.array_equal(out["cam"][i].cpu().numpy().view(np.uint32), o["cam"].view(np.uint32))
        assert np.array_equal(out["sel"][i].cpu().numpy().astype(bool), o["sel"])


def test_s1_bce_mask_mode(cuda):
    """Sigmoid uses expf on both sides: gate identical away from the threshold, weights within 2 ulp."""
    b = synth.make_batch(3, H=8, seed=10)
    b["mask"] = ((b["mask"] - 0.5) * 8).astype(np.float32)  # logits
    g = _to_cuda(b)
    out = pose_solver.correspond(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"],
                                 g["extent"], g["region_idx"], g["anchors"], mask_mode="bce")
    for i in range(3):
        o = _s1_oracle(b, i, po.MASK_BCE)
        np.testing.assert_allclose(out["w"][i].cpu().numpy(), o["w"], rtol=3e-7)
        far = np.abs(o["w"] - 0.5) > 1e-5
        assert np.array_equal(out["sel"][i].cpu().numpy().astype(bool)[far], o["sel"][far])


def test_s1_flat_mask_and_empty_depth(cuda):
    b = synth.make_batch(2, H=8, seed=12)
    b["mask"][0] = 0.3  # flat -> (m-min)/(max-min) = NaN -> nothing selected (engine_utils.py:123-128, no eps)
    b["depth"][1] = 0.0  # no depth anywhere
    g = _to_cuda(b)
    out = pose_solver.correspond(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"],
                                 g["extent"], g["region_idx"], g["anchors"])
    assert out["n_sel"].tolist() == [0, 0]
    assert torch.isnan(out["w"][0]).all()


def test_roi_intrinsics_bit_exact_vs_oracle_and_reference_golden(cuda, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "affine_golden.npz"))
    c32, s32 = g["centers32"].astype(np.float32), g["scales32"].astype(np.float32)
    K = np.repeat(synth.K_LM[None].astype(np.float32), len(s32), 0)
    Kp = geometry.roi_intrinsics(torch.from_numpy(K).cuda(), torch.from_numpy(c32).cuda(), torch.from_numpy(s32).cuda()).cpu().numpy()
    for i in range(len(s32)):
        o = po.roi_intrinsics(K[i], c32[i], s32[i]).astype(np.float32)
        assert np.array_equal(Kp[i].view(np.uint32), o.view(np.uint32))
        Kref = np.vstack([g["A256_32"][i], [0, 0, 1]]) @ K[i].astype(np.float64)  # data_loader.py:555-564 on the reference affine
        ref = np.array([Kref[0, 0], Kref[1, 1], Kref[0, 2], Kref[1, 2]])
        np.testing.assert_allclose(Kp[i], ref, rtol=2e-7)


def test_generic_backproject_bit_exact(cuda):
    rng = np.random.default_rng(3)
    d = rng.uniform(0, 2, (3, 48, 80)).astype(np.float32)
    K = np.array([[500.5, 0, 41.25], [0, 499.75, 23.5], [0, 0, 1]], np.float32)
    out = geometry.backproject_th(torch.from_numpy(d).cuda(), torch.from_numpy(K).cuda()).cpu().numpy()
    for i in range(3):
        assert np.array_equal(out[i].view(np.uint32), po.backproject(d[i], K).view(np.uint32))
    one = geometry.backproject_th(torch.from_numpy(d[0]).cuda(), torch.from_numpy(K).cuda())
    assert one.shape == (48, 80, 3)


def test_region_argmax_matches_torch_and_oracle(cuda):
    g = torch.Generator(device="cuda").manual_seed(0)
    reg = torch.randn(5, 33, 64, 64, device="cuda", generator=g)
    reg[:, 5] = reg[:, 9]  # ties -> first maximum wins
    idx = geometry.region_argmax(reg)
    ref = torch.argmax(torch.softmax(reg[:, 1:], dim=1), dim=1)  # GDRN.py:206-209
    ref_logit = torch.argmax(reg[:, 1:], dim=1)
    assert torch.equal(idx.long(), ref_logit)
    assert (idx.long() == ref).float().mean() > 0.999  # softmax rounding can merge near-ties
    assert np.array_equal(idx[0].cpu().numpy(), po.region_argmax(reg[0].cpu().numpy()))


def test_roi_crop_depth_matches_cv2_warpaffine(cuda):
    """f1: GPU ROI crop sampled at the kept pixels vs the loader's cv2.warpAffine(...)[::4, ::4]
    (data_loader.py:532-535, 625).  OpenCV's fixed-point coordinates are reproduced, so the only slack is
    float rounding of the 4-tap blend (OpenCV's SIMD path may fuse multiply-adds)."""
    pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    depth = rng.uniform(0.4, 1.6, (2, 480, 640)).astype(np.float32)
    depth[:, 100:200, 300:400] = 0.0  # holes
    B = 24
    centers = np.stack([rng.uniform(40, 600, B), rng.uniform(40, 440, B)], 1).astype(np.float32)
    scales = rng.uniform(40, 640, B).astype(np.float32)
    centers[0] = [5, 5]  # crop hangs over the image border -> zero taps
    scales[0] = 200
    idx = (np.arange(B) % 2).astype(np.int32)
    out = geometry.roi_crop_depth(torch.from_numpy(depth).cuda(), torch.from_numpy(centers).cuda(), torch.from_numpy(scales).cuda(),
                                  torch.from_numpy(idx).cuda()).cpu().numpy()
    for b in range(B):
        ref = po.roi_crop_depth_cv2(depth[idx[b]], centers[b], scales[b])
        assert ref.shape == (64, 64)
        assert np.array_equal(out[b].view(np.uint32), ref.view(np.uint32)), b  # bit-identical to cv2 4.13
    assert np.array_equal(out[3], po.roi_crop_depth(depth[idx[3]], centers[3], scales[3]))  # and to the restatement
