"""GPU: the fused solver (csrc/pose_solve.cu through the C ABI) against the oracle and the golden run.

Bars (BASELINE.json north_star): bit-exact inlier masks and counts on identical inputs and hypothesis
index sets (boundary ties documented below), rotation within 1e-5 rad, translation within 1e-3 mm.

Boundary ties: the kernel derives each hypothesis pose with a closed-form FP64 solve, the oracle with
numpy's SVD as the reference does; both round once to FP32.  The two FP64 results agree to ~1e-13, so
in rare cases one FP32 pose element differs by 1 ulp and a point whose residual sits within 1 ulp of the
threshold may flip.  The tests therefore demand exact count equality on every hypothesis whose FP32 pose
is bit-identical, and allow |delta count| <= 2 on the (rare) others; the winning hypothesis, its inlier
mask and count must always match exactly.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import pose_oracle as po
from rdpn6d_b200 import _lib, pose_solver, synth

pytestmark = pytest.mark.gpu
ROT_TOL_RAD = 1e-5
TRANS_TOL_M = 1e-6  # 1e-3 mm
THR = 0.005


def _to_cuda(b):
    return {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in b.items()}


def _solve(g, **kw):
    kw.setdefault("inlier_thr", THR)
    args = (g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"], g["extent"], g["hyp_idx"])
    kwargs = dict(region_idx=g.get("region_idx"), anchors=g.get("anchors"), t_net=g.get("t_net"))
    res = pose_solver.PoseSolver(want_inlier_mask=True, want_hyp=True, **kw)(*args, **kwargs)
    # The production call (no per-hypothesis diagnostics) must give bit-identical results to the diagnostic call.
    fast = pose_solver.PoseSolver(want_inlier_mask=True, want_hyp=False, **kw)(*args, **kwargs)
    assert torch.equal(fast.pose.view(torch.int32), res.pose.view(torch.int32))
    assert torch.equal(fast.best_h, res.best_h) and torch.equal(fast.n_inliers, res.n_inliers)
    assert torch.equal(fast.status, res.status) and torch.equal(fast.n_sel, res.n_sel)
    assert torch.equal(fast.inlier_mask, res.inlier_mask)
    # ... and the two implementations (three-kernel pipeline / fused kernel) agree: integers and hypothesis poses bit for
    # bit, the refit pose to FP32 rounding (its FP64 sums run in a different order)
    if g.get("region_idx") is not None:
        assert _lib.lib().rdpn_pose_solve_workspace_bytes(1, 8, 32, 0) > 0
        fused = pose_solver.PoseSolver(want_inlier_mask=True, want_hyp=True, pipeline="fused", **kw)(*args, **kwargs)
        split = pose_solver.PoseSolver(want_inlier_mask=True, want_hyp=True, pipeline="split", chunk_rois=5, **kw)(*args, **kwargs)
        for other in (fused, split):
            for name in ("best_h", "n_inliers", "status", "n_sel", "inlier_mask", "hyp_counts"):
                assert torch.equal(getattr(other, name), getattr(res, name)), name
            assert torch.equal(other.hyp_poses.view(torch.int32), res.hyp_poses.view(torch.int32))
            ok = res.status == 0
            assert torch.allclose(other.pose[ok], res.pose[ok], rtol=0, atol=2e-6)
            assert torch.equal(other.pose[~ok], res.pose[~ok])
    return res


def _compare(res, ores, b, check_counts=True):
    B = len(ores)
    pose = res.pose.cpu().numpy()
    stats = dict(pose_mismatch=0, hyp=0)
    for i in range(B):
        o = ores[i]
        assert int(res.status[i]) == o["status"], i
        assert int(res.n_sel[i]) == o["n_sel"], i
        assert int(res.best_h[i]) == o["best_h"], i
        assert int(res.n_inliers[i]) == o["n_inl"], i
        if check_counts:
            hp = res.hyp_poses[i].reshape(-1, 12).cpu().numpy()
            cnt = res.hyp_counts[i].cpu().numpy()
            if o["n_sel"] >= 4:
                np.testing.assert_allclose(hp, o["Rt_hyp"], atol=2e-6)
                same = (hp.view(np.uint32) == o["Rt_hyp"].view(np.uint32)).all(axis=1)
                assert np.array_equal(cnt[same], o["counts"][same]), i
                assert np.abs(cnt[~same].astype(int) - o["counts"][~same]).max(initial=0) <= 2
                stats["pose_mismatch"] += int((~same).sum())
                stats["hyp"] += len(same)
        assert np.array_equal(res.inlier_mask[i].reshape(-1).cpu().numpy(), o["inlier_mask"]), i
        if o["status"] in (po.STATUS_OK, po.STATUS_T_SANITY):
            assert po.re_rad_small(pose[i][:, :3], o["pose"][:, :3]) <= ROT_TOL_RAD, i
            assert po.te(pose[i][:, 3], o["pose"][:, 3]) <= TRANS_TOL_M, i
        else:
            assert (pose[i] == -100).all()
    if stats["hyp"]:
        assert stats["pose_mismatch"] <= 0.001 * stats["hyp"] + 1  # FP32 hypothesis poses are ~always bit-identical
    # the kernel-written gather rows are exactly the individual outputs
    rows = res.rows16()
    expect = torch.cat([res.pose.reshape(B, 12), res.n_inliers.float()[:, None], res.status.float()[:, None],
                        res.n_sel.float()[:, None], res.best_h.float()[:, None]], dim=1)
    assert rows.shape == (B, 16) and torch.equal(rows, expect)
    return stats


def test_golden_batch_with_reference_kabsch(cuda, golden_dir):
    """4 ROIs whose expected outputs were produced with the reference's affine_matrix_from_points."""
    g = np.load(os.path.join(golden_dir, "pose_golden.npz"))
    b = {k: g[k] for k in ("depth", "Kp", "coor", "mask", "extent", "region_idx", "anchors", "hyp_idx")}
    res = _solve(_to_cuda(b), inlier_thr=float(g["thr"]))
    pose = res.pose.cpu().numpy()
    for i in range(4):
        assert int(res.status[i]) == g["out_status"][i]
        assert int(res.n_sel[i]) == g["out_nsel"][i]
        assert int(res.best_h[i]) == g["out_best_h"][i]
        assert int(res.n_inliers[i]) == g["out_ninl"][i]
        assert np.array_equal(res.inlier_mask[i].reshape(-1).cpu().numpy(), g["out_inlier_mask"][i])
        hp = res.hyp_poses[i].reshape(-1, 12).cpu().numpy()
        same = (hp.view(np.uint32) == g["out_Rt_hyp"][i].view(np.uint32)).all(axis=1)
        assert np.array_equal(res.hyp_counts[i].cpu().numpy()[same], g["out_counts"][i][same])
        assert po.re_rad_small(pose[i][:, :3], g["out_pose"][i][:, :3]) <= ROT_TOL_RAD
        assert po.te(pose[i][:, 3], g["out_pose"][i][:, 3]) <= TRANS_TOL_M


def test_config0_lm13_32_rois_vs_oracle(cuda):
    """BASELINE configs[0]: 32 synthetic 64x64 ROIs, H=256 -- full parity against the CPU oracle."""
    b = synth.make_batch(32, H=256, seed=20260101)
    ores = po.pose_solve_batch(b, b["hyp_idx"], THR)
    res = _solve(_to_cuda(b))
    _compare(res, ores, b)
    # and the estimate is the right pose
    pose = res.pose.cpu().numpy()
    for i in range(32):
        assert po.re_rad_small(pose[i][:, :3], b["gt_pose"][i][:, :3]) < 0.02
        assert po.te(pose[i][:, 3], b["gt_pose"][i][:, 3]) < 0.002


@pytest.mark.parametrize("H", [16, 64, 100, 256, 512])
def test_hypothesis_counts_various_H(cuda, H):
    b = synth.make_batch(6, H=H, seed=100 + H, occlusion_max=0.4)
    ores = po.pose_solve_batch(b, b["hyp_idx"], THR)
    _compare(_solve(_to_cuda(b)), ores, b)


def test_dense_mode(cuda):
    b = synth.make_batch(8, H=128, seed=55, dense=True)
    ores = po.pose_solve_batch(b, b["hyp_idx"], THR)
    _compare(_solve(_to_cuda(b)), ores, b)


def test_symmetric_objects_lmo_like_64_regions(cuda):
    models = synth.make_models(6, num_regions=64, seed=3, n_symmetric=3)
    b = synth.make_batch(12, models=models, H=128, seed=56, K=synth.K_YCBV, occlusion_max=0.6)
    ores = po.pose_solve_batch(b, b["hyp_idx"], THR)
    _compare(_solve(_to_cuda(b)), ores, b)


@pytest.mark.parametrize("kw", [dict(weighted=True), dict(refit_iters=3), dict(with_scale=True),
                                dict(adaptive=True), dict(adaptive=True, min_iter=3, confidence=0.9),
                                dict(mask_thr=0.7, mask_mode=po.MASK_RAW), dict(min_inliers=50, min_pts=10)])
def test_solver_options(cuda, kw):
    b = synth.make_batch(6, H=96, seed=77)
    okw = dict(kw)
    okw["scale"] = okw.pop("with_scale", False)
    ores = po.pose_solve_batch(b, b["hyp_idx"], THR, **okw)
    res = _solve(_to_cuda(b), **kw)
    _compare(res, ores, b)
    if kw.get("with_scale"):
        for i in range(6):
            assert abs(float(res.scale[i]) - ores[i]["scale"]) < 1e-5


def test_edge_cases_status_codes(cuda):
    b = synth.make_batch(6, H=32, seed=88)
    b["mask"][0] = 0.4  # flat mask -> NaN -> 0 selected -> FEW_POINTS
    b["depth"][1] = 0  # no depth -> FEW_POINTS
    b["hyp_idx"][2] = b["hyp_idx"][2][:, :1]  # every triplet is one pixel three times -> degenerate -> NO_CONSENSUS
    b["hyp_idx"][3] = 4095  # background pixel (not gated) -> invalid hypotheses -> NO_CONSENSUS
    b["hyp_idx"][4, ::2] = -7  # out-of-range indices are rejected, the rest still work
    keep = np.zeros((64, 64), bool)
    keep[30:32, 30:31] = True  # 2 pixels only
    b["mask"][5] = np.where(keep, 0.9, 0.1)
    ores = po.pose_solve_batch(b, b["hyp_idx"], THR)
    res = _solve(_to_cuda(b))
    _compare(res, ores, b)
    st = res.status.cpu().tolist()
    assert st[0] == st[1] == st[5] == pose_solver.STATUS_FEW_POINTS
    assert st[2] == st[3] == pose_solver.STATUS_NO_CONSENSUS
    assert st[4] == pose_solver.STATUS_OK


def test_translation_sanity_fallback(cuda):
    b = synth.make_batch(4, H=64, seed=99)
    t_net = b["gt_pose"][:, :, 3].astype(np.float32)
    t_net[1] += np.array([0, 0, 1.5], np.float32)  # > 1 m away -> status 2, translation replaced by t_net
    b["t_net"] = t_net
    ores = po.pose_solve_batch(b, b["hyp_idx"], THR, t_net=t_net)
    res = _solve(_to_cuda(b))
    _compare(res, ores, b, check_counts=False)
    assert res.status.cpu().tolist() == [0, 2, 0, 0]
    np.testing.assert_array_equal(res.pose[1, :, 3].cpu().numpy(), t_net[1])


@pytest.mark.parametrize("H,R", [(600, 130), (1, 3), (129, 255), (2048, 32)])
def test_pipeline_odd_sizes_match_oracle(cuda, H, R):
    """The three-kernel pipeline at sizes off its fast path: hypothesis counts that are not a multiple of 128 (last
    scoring pass of 1-4 per lane, more than one full pass), one hypothesis, region counts up to the 255 the format
    allows (8 buckets per lane in the bucket scan), B not a multiple of the warps per CTA -- against the oracle and the
    fused kernel (_solve runs both)."""
    models = synth.make_models(3, R, seed=11)
    b = synth.make_batch(7, models=models, H=H, seed=1000 + H, occlusion_max=0.3)
    res = _solve(_to_cuda(b))
    _compare(res, po.pose_solve_batch(b, b["hyp_idx"], THR), b)


def test_full_size_lmo_1024_properties(cuda):
    """BASELINE configs[1] at full size: size-independent properties instead of a full oracle run
    (determinism, ROI-order equivariance, ground-truth recovery) + an oracle spot check."""
    models = synth.make_models(8, 64, seed=1)  # the bench's LM-O workload: 64 anchors per object (bench.py WORKLOADS["lmo"])
    base = synth.make_batch(128, models=models, H=256, seed=20260101, occlusion_max=0.6)
    b = synth.tile_batch(base, 1024)
    g = _to_cuda(b)
    r1 = _solve(g)
    p1 = r1.pose.clone()
    n1 = r1.n_inliers.clone()
    m1 = r1.inlier_mask.clone()
    r2 = _solve(g)
    assert torch.equal(p1, r2.pose) and torch.equal(n1, r2.n_inliers) and torch.equal(m1, r2.inlier_mask)  # deterministic
    assert torch.equal(p1[:128], p1[128:256]) and torch.equal(p1[:128], p1[896:])  # same ROI -> same bits, any slot
    perm = torch.randperm(1024, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    gp = {k: (None if v is None else v[perm].contiguous()) for k, v in g.items()}
    r3 = _solve(gp)
    assert torch.equal(r3.pose, p1[perm]) and torch.equal(r3.n_inliers, n1[perm])  # equivariant to ROI order
    ok = (r1.status == 0).cpu().numpy()
    assert ok.mean() > 0.9
    pose = p1.cpu().numpy()
    errs = [po.re_rad_small(pose[i][:, :3], b["gt_pose"][i][:, :3]) for i in range(128) if ok[i]]
    assert np.median(errs) < 0.01
    ores = po.pose_solve_batch(base, base["hyp_idx"], THR)  # EVERY unique ROI of the config against the oracle
    for i in range(128):
        assert int(r1.status[i]) == ores[i]["status"], i
        assert int(n1[i]) == ores[i]["n_inl"] and int(r1.best_h[i]) == ores[i]["best_h"], i
        assert np.array_equal(m1[i].reshape(-1).cpu().numpy(), ores[i]["inlier_mask"]), i
        if ores[i]["status"] == 0:
            assert po.re_rad_small(pose[i][:, :3], ores[i]["pose"][:, :3]) <= ROT_TOL_RAD
            assert po.te(pose[i][:, 3], ores[i]["pose"][:, 3]) <= TRANS_TOL_M


def test_full_size_ycbv_8192_incl_symmetric(cuda):
    """BASELINE configs[2]: 21 objects (5 with symmetric geometry), 8192 ROIs, YCB-V intrinsics (the bench's headline
    workload).  Properties at full size + oracle parity on EVERY unique ROI, unweighted and with the weighted refit the
    bench runs."""
    models = synth.make_models(21, 32, seed=7, n_symmetric=5)
    base = synth.make_batch(84, models=models, H=256, seed=777, K=synth.K_YCBV, occlusion_max=0.5)
    b = synth.tile_batch(base, 8192)
    g = _to_cuda(b)
    r = _solve(g)
    assert torch.equal(r.pose[:84], r.pose[84 * 96:84 * 97])  # any slot, same bits
    assert float((r.status == 0).float().mean()) > 0.95
    for weighted in (False, True):
        if weighted:
            r = _solve(g, weighted=True)
        ores = po.pose_solve_batch(base, base["hyp_idx"], THR, weighted=weighted)
        pose = r.pose.cpu().numpy()
        for i in range(84):
            assert int(r.status[i]) == ores[i]["status"], i
            assert int(r.n_inliers[i]) == ores[i]["n_inl"] and int(r.best_h[i]) == ores[i]["best_h"], i
            assert np.array_equal(r.inlier_mask[i].reshape(-1).cpu().numpy(), ores[i]["inlier_mask"]), i
            if ores[i]["status"] == 0:
                assert po.re_rad_small(pose[i][:, :3], ores[i]["pose"][:, :3]) <= ROT_TOL_RAD, i
                assert po.te(pose[i][:, 3], ores[i]["pose"][:, 3]) <= TRANS_TOL_M, i


def test_mp6d_scale_65536_rois_shard_invariance(cuda):
    """BASELINE configs[4]: 65 536 ROIs (5.6 GB of maps).  Size-independent properties on ONE GPU: every tile copy of
    the 64 base ROIs gives the same bits, and solving the job in the contiguous shards of an 8-rank run
    (my_distributed_sampler.py:189-192) reproduces the single-call rows -- with explicit triplets and with the
    kernel-drawn ones (roi_base keeps the stream independent of the sharding)."""
    from rdpn6d_b200 import distributed

    B, U, H = 65536, 64, 64
    models = synth.make_models(20, 32, seed=9, n_symmetric=4)
    base = synth.make_batch(U, models=models, H=H, seed=4242, K=synth.K_YCBV, occlusion_max=0.5)
    g = {k: (None if v is None else torch.from_numpy(v).cuda().repeat(B // U, *([1] * (v.ndim - 1)))) for k, v in base.items()}
    args = [g["depth"], g["Kp"], g["coor"][:, 0].contiguous(), g["coor"][:, 1].contiguous(), g["coor"][:, 2].contiguous(),
            g["mask"], g["extent"]]
    solver = pose_solver.PoseSolver(inlier_thr=THR, num_hyp=H, seed=5)
    full = solver(*args, g["hyp_idx"], region_idx=g["region_idx"], anchors=g["anchors"])
    rows = full.rows16().clone()
    assert rows.shape == (B, 16)
    assert torch.equal(rows.view(B // U, U, 16)[0].expand(B // U, U, 16), rows.view(B // U, U, 16))
    assert float((full.status == 0).float().mean()) > 0.9
    auto_rows = solver(*args, None, region_idx=g["region_idx"], anchors=g["anchors"]).rows16().clone()
    assert not torch.equal(auto_rows[:U], auto_rows[U:2 * U])  # every ROI draws its own triplets
    for rank in (0, 3, 7):
        b0, b1 = distributed.shard_range(B, rank, 8)
        sl = slice(b0, b1)
        part = solver(*[a[sl] for a in args], g["hyp_idx"][sl], region_idx=g["region_idx"][sl], anchors=g["anchors"][sl])
        assert torch.equal(part.rows16(), rows[sl])
        part = solver(*[a[sl] for a in args], None, region_idx=g["region_idx"][sl], anchors=g["anchors"][sl], roi_base=b0)
        assert torch.equal(part.rows16(), auto_rows[sl])


def test_many_gated_points_multi_chunk(cuda):
    """More than 1024 gated pixels per ROI: the staging/scoring loop runs several chunks and the refit
    re-gathers its slots.  Big objects filling the crop, no dropout, no outliers."""
    models = [synth.ObjectModel("box", [0.12, 0.12, 0.12], 32, np.random.default_rng(0)),
              synth.ObjectModel("ellipsoid", [0.125, 0.125, 0.12], 32, np.random.default_rng(1))]
    b = synth.make_batch(6, models=models, H=64, seed=3, dzi_pad_scale=1.0, mask_dropout=0.0, outlier_frac=0.02)
    ores = po.pose_solve_batch(b, b["hyp_idx"], THR)
    assert max(o["n_sel"] for o in ores) > 1024
    _compare(_solve(_to_cuda(b)), ores, b)
    bd = synth.make_batch(4, models=models, H=64, seed=4, dzi_pad_scale=1.0, mask_dropout=0.0, outlier_frac=0.02, dense=True)
    od = po.pose_solve_batch(bd, bd["hyp_idx"], THR)
    assert max(o["n_sel"] for o in od) > 512
    _compare(_solve(_to_cuda(bd)), od, bd)


def test_host_buffer_plugin_call_matches_device_call(cuda):
    """rdpn_pose_solve_host: HOST pointers in/out through the C ABI (chunked, two streams)."""
    L = _lib.lib()
    B, H = 600, 64  # spans three pipeline chunks (256 ROIs each)
    base = synth.make_batch(40, H=H, seed=7)
    b = synth.tile_batch(base, B)
    dev = _solve(_to_cuda(b))
    ctx = ctypes.c_void_p()
    _lib.check(L.rdpn_ctx_create(0, ctypes.byref(ctx)), "ctx_create")
    try:
        cx, cy, cz = [np.ascontiguousarray(b["coor"][:, c]) for c in range(3)]
        inp = _lib.RoiInputs(depth=b["depth"].ctypes.data, Kp=b["Kp"].ctypes.data, depth_div=None, coor_x=cx.ctypes.data,
                             coor_y=cy.ctypes.data, coor_z=cz.ctypes.data, mask=b["mask"].ctypes.data,
                             extent=b["extent"].ctypes.data, region_idx=b["region_idx"].ctypes.data,
                             anchors=b["anchors"].ctypes.data, num_regions=32, mask_mode=1, mask_thr=0.5, B=B)
        prm = _lib.SolveParams(inlier_thr=THR, num_hyp=H, min_pts=4, min_inliers=4, weighted=0, refit_iters=1,
                               with_scale=0, adaptive=0, confidence=0.995, min_iter=10)
        pose = np.zeros((B, 12), np.float32)
        ninl = np.zeros(B, np.int32)
        status = np.zeros(B, np.int32)
        best = np.zeros(B, np.int32)
        imask = np.zeros((B, 4096), np.uint8)
        out = _lib.SolveOutputs(pose=pose.ctypes.data, n_inliers=ninl.ctypes.data, status=status.ctypes.data,
                                best_h=best.ctypes.data, inlier_mask=imask.ctypes.data)
        _lib.check(L.rdpn_pose_solve_host(ctx, ctypes.byref(inp), b["hyp_idx"].ctypes.data, None, ctypes.byref(prm),
                                          ctypes.byref(out)), "pose_solve_host")
    finally:
        L.rdpn_ctx_destroy(ctx)
    assert np.array_equal(pose.view(np.uint32), dev.pose.reshape(B, 12).cpu().numpy().view(np.uint32))
    assert np.array_equal(ninl, dev.n_inliers.cpu().numpy()) and np.array_equal(status, dev.status.cpu().numpy())
    assert np.array_equal(best, dev.best_h.cpu().numpy())
    assert np.array_equal(imask, dev.inlier_mask.reshape(B, -1).cpu().numpy())


def _host_call(L, ctx, t, B, H, R, mask_mode=1, dense=False, pinned_out=False):
    """rdpn_pose_solve_host on a dict of (pinned or pageable) torch CPU tensors; returns numpy outputs."""
    inp = _lib.RoiInputs(depth=t["depth"].data_ptr(), Kp=t["Kp"].data_ptr(), depth_div=None, coor_x=t["cx"].data_ptr(),
                         coor_y=t["cy"].data_ptr(), coor_z=t["cz"].data_ptr(), mask=t["mask"].data_ptr(),
                         extent=t["extent"].data_ptr(), region_idx=None if dense else t["region_idx"].data_ptr(),
                         anchors=None if dense else t["anchors"].data_ptr(), num_regions=R, mask_mode=mask_mode,
                         mask_thr=0.5, B=B)
    prm = _lib.SolveParams(inlier_thr=THR, num_hyp=H, min_pts=4, min_inliers=4, weighted=0, refit_iters=1,
                           with_scale=0, adaptive=0, confidence=0.995, min_iter=10)
    o = {"pose": np.zeros((B, 12), np.float32), "ninl": np.zeros(B, np.int32), "status": np.zeros(B, np.int32),
         "best": np.zeros(B, np.int32), "nsel": np.zeros(B, np.int32), "imask": np.zeros((B, 4096), np.uint8)}
    keep = None
    if pinned_out:  # mapped result buffers and no per-pixel diagnostics: the kernel writes them directly
        keep = {k: torch.from_numpy(v).pin_memory() for k, v in o.items() if k != "imask"}
        o = {k: v.numpy() for k, v in keep.items()}
    out = _lib.SolveOutputs(pose=o["pose"].ctypes.data, n_inliers=o["ninl"].ctypes.data, status=o["status"].ctypes.data,
                            best_h=o["best"].ctypes.data, n_sel=o["nsel"].ctypes.data,
                            inlier_mask=None if pinned_out else o["imask"].ctypes.data)
    _lib.check(L.rdpn_pose_solve_host(ctx, ctypes.byref(inp), t["hyp_idx"].data_ptr(), None, ctypes.byref(prm),
                                      ctypes.byref(out)), "pose_solve_host")
    return {k: v.copy() for k, v in o.items()}


def _host_tensors(b, pinned):
    t = {k: torch.from_numpy(np.ascontiguousarray(b[k])) for k in ("depth", "Kp", "mask", "extent", "region_idx", "anchors",
                                                                   "hyp_idx") if b.get(k) is not None}
    for c, name in enumerate(("cx", "cy", "cz")):
        t[name] = torch.from_numpy(np.ascontiguousarray(b["coor"][:, c]))
    return {k: v.pin_memory() for k, v in t.items()} if pinned else t


@pytest.mark.parametrize("mask_mode", [1, 2, 0])
def test_gated_pull_transfer_is_bit_identical_to_full_copy(cuda, mask_mode):
    """Pinned host buffers: only the mask plane is copied, depth / coor / region ids are fetched over PCIe for
    the pixel groups whose mask test passes.  Every output must equal the full-copy strategy bit for bit, for
    every fetch granularity, and fewer bytes must cross the bus."""
    L = _lib.lib()
    B, H = 300, 64
    b = synth.tile_batch(synth.make_batch(30, H=H, seed=11, occlusion_max=0.5), B)
    if mask_mode == 2:  # logits: sigmoid(m) > 0.5 <=> m > 0
        b["mask"] = ((b["mask"] - 0.5) * 8).astype(np.float32)
    t = _host_tensors(b, pinned=True)
    ctx = ctypes.c_void_p()
    _lib.check(L.rdpn_ctx_create(0, ctypes.byref(ctx)), "ctx_create")
    try:
        _lib.check(L.rdpn_ctx_set_option(ctx, _lib.OPT_COUNT_BYTES, 1), "opt")
        _lib.check(L.rdpn_ctx_set_option(ctx, _lib.OPT_CHUNK_ROIS, 128), "opt")
        _lib.check(L.rdpn_ctx_set_option(ctx, _lib.OPT_TRANSFER, _lib.TRANSFER_COPY), "opt")
        ref = _host_call(L, ctx, t, B, H, 32, mask_mode)
        assert L.rdpn_ctx_last_transfer(ctx) == _lib.TRANSFER_COPY
        full_bytes = L.rdpn_ctx_last_h2d_bytes(ctx)
        assert full_bytes >= B * (5 * 16384 + 4096)
        assert (ref["status"] == 0).mean() > 0.8
        _lib.check(L.rdpn_ctx_set_option(ctx, _lib.OPT_TRANSFER, _lib.TRANSFER_AUTO), "opt")
        for gran in (1, 2, 8, 16):
            _lib.check(L.rdpn_ctx_set_option(ctx, _lib.OPT_PULL_GRANULARITY, gran), "opt")
            got = _host_call(L, ctx, t, B, H, 32, mask_mode)
            assert L.rdpn_ctx_last_transfer(ctx) == _lib.TRANSFER_PULL
            for k in ref:
                assert np.array_equal(got[k].view(np.uint8), ref[k].view(np.uint8)), (gran, k)
            pulled = L.rdpn_ctx_last_h2d_bytes(ctx)
            assert pulled < 0.75 * full_bytes, (gran, pulled, full_bytes)
            direct = _host_call(L, ctx, t, B, H, 32, mask_mode, pinned_out=True)
            for k in direct:
                assert np.array_equal(direct[k].view(np.uint8), ref[k].view(np.uint8)), (gran, k, "direct")
        # pageable buffers cannot be pulled: AUTO copies, an explicit PULL is refused
        tp = _host_tensors(b, pinned=False)
        got = _host_call(L, ctx, tp, B, H, 32, mask_mode)
        assert L.rdpn_ctx_last_transfer(ctx) == _lib.TRANSFER_COPY
        assert np.array_equal(got["pose"].view(np.uint32), ref["pose"].view(np.uint32))
        _lib.check(L.rdpn_ctx_set_option(ctx, _lib.OPT_TRANSFER, _lib.TRANSFER_PULL), "opt")
        with pytest.raises(RuntimeError):
            _host_call(L, ctx, tp, B, H, 32, mask_mode)
    finally:
        L.rdpn_ctx_destroy(ctx)


def test_gated_pull_dense_mode(cuda):
    L = _lib.lib()
    B, H = 64, 32
    b = synth.make_batch(B, H=H, seed=5, dense=True)
    t = _host_tensors(b, True)
    ctx = ctypes.c_void_p()
    _lib.check(L.rdpn_ctx_create(0, ctypes.byref(ctx)), "ctx_create")
    try:
        _lib.check(L.rdpn_ctx_set_option(ctx, _lib.OPT_TRANSFER, _lib.TRANSFER_COPY), "opt")
        ref = _host_call(L, ctx, t, B, H, 0, 1, dense=True)
        _lib.check(L.rdpn_ctx_set_option(ctx, _lib.OPT_TRANSFER, _lib.TRANSFER_PULL), "opt")
        got = _host_call(L, ctx, t, B, H, 0, 1, dense=True)
        for k in ref:
            assert np.array_equal(got[k].view(np.uint8), ref[k].view(np.uint8)), k
    finally:
        L.rdpn_ctx_destroy(ctx)


def test_host_pose_solver_wrapper(cuda):
    """Python face of the host-buffer call: pinned CPU tensors in, pinned CPU tensors out, gated pull underneath."""
    B, H = 96, 64
    b = synth.make_batch(B, H=H, seed=21, occlusion_max=0.4)
    dev = _solve(_to_cuda(b))
    t = {k: (None if v is None else torch.from_numpy(np.ascontiguousarray(v)).pin_memory()) for k, v in b.items()}
    cx, cy, cz = [t["coor"][:, c].contiguous().pin_memory() for c in range(3)]
    hs = pose_solver.HostPoseSolver(inlier_thr=THR, count_bytes=True)
    res = hs(t["depth"], t["Kp"], cx, cy, cz, t["mask"], t["extent"], t["hyp_idx"], t["region_idx"], t["anchors"])
    assert hs.last_transfer == "pull" and 0 < hs.last_h2d_bytes < B * (5 * 16384 + 4096)
    assert torch.equal(res.pose.view(torch.int32), dev.pose.cpu().view(torch.int32))
    assert torch.equal(res.n_inliers, dev.n_inliers.cpu()) and torch.equal(res.status, dev.status.cpu())
    assert torch.equal(res.best_h, dev.best_h.cpu()) and torch.equal(res.n_sel, dev.n_sel.cpu())
    # pageable tensors: same results through the full copy
    hc = pose_solver.HostPoseSolver(inlier_thr=THR, pin_outputs=False, want_inlier_mask=True)
    u = {k: (None if v is None else torch.from_numpy(np.ascontiguousarray(v))) for k, v in b.items()}
    res2 = hc(u["depth"], u["Kp"], u["coor"][:, 0], u["coor"][:, 1], u["coor"][:, 2], u["mask"], u["extent"], u["hyp_idx"],
              u["region_idx"], u["anchors"])
    assert hc.last_transfer == "copy"
    assert torch.equal(res2.pose.view(torch.int32), dev.pose.cpu().view(torch.int32))
    assert torch.equal(res2.inlier_mask, dev.inlier_mask.cpu())
    # the deployment split of the reference: head outputs (coor, mask, region ids) already on the GPU, the loader's
    # depth / intrinsics / extents / anchors and the hypothesis triplets in pinned host memory
    g = _to_cuda(b)
    res3 = hs(t["depth"], t["Kp"], g["coor"][:, 0].contiguous(), g["coor"][:, 1].contiguous(), g["coor"][:, 2].contiguous(),
              g["mask"], t["extent"], t["hyp_idx"], g["region_idx"], t["anchors"])
    assert hs.last_transfer == "pull" and 0 < hs.last_h2d_bytes < B * (16384 + H * 12 + 32 * 12 + 28 + 1)
    assert torch.equal(res3.pose.view(torch.int32), dev.pose.cpu().view(torch.int32))
    assert torch.equal(res3.n_inliers, dev.n_inliers.cpu()) and torch.equal(res3.best_h, dev.best_h.cpu())
    # everything on the device: nothing moves, results still land in the pinned CPU tensors
    res4 = hs(g["depth"], g["Kp"], g["coor"][:, 0].contiguous(), g["coor"][:, 1].contiguous(), g["coor"][:, 2].contiguous(),
              g["mask"], g["extent"], g["hyp_idx"], g["region_idx"], g["anchors"])
    assert hs.last_h2d_bytes == 0
    assert torch.equal(res4.pose.view(torch.int32), dev.pose.cpu().view(torch.int32))
    hs.close()
    hc.close()


def test_cpu_tensors_are_rejected_loudly(cuda):
    b = synth.make_batch(2, H=8, seed=1)
    t = {k: (None if v is None else torch.from_numpy(v)) for k, v in b.items()}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pose_solver.pose_solve(t["depth"], t["Kp"], t["coor"][:, 0], t["coor"][:, 1], t["coor"][:, 2], t["mask"],
                               t["extent"], t["hyp_idx"], t["region_idx"], t["anchors"])


def test_sample_hypotheses_draws_gated_pixels(cuda):
    b = synth.make_batch(5, H=8, seed=2)
    g = _to_cuda(b)
    s1 = pose_solver.correspond(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"],
                                g["extent"], g["region_idx"], g["anchors"])
    hyp = pose_solver.sample_hypotheses(s1["sel"], 64, generator=torch.Generator(device="cuda").manual_seed(0))
    assert hyp.shape == (5, 64, 3) and hyp.dtype == torch.int32
    picked = torch.gather(s1["sel"].long(), 1, hyp.reshape(5, -1).long())
    assert bool(picked.all())


def test_internal_sampling_equals_explicit_triplets_from_the_oracle(cuda):
    """hyp_idx=None: the kernel draws the triplets itself (counter-based stream documented in the header).  The
    oracle's sample_triplets is the same arithmetic on the oracle's own gate: feeding its triplets back explicitly
    must give bit-identical results -- for the device call, for the chunked host call and for a sharded call with
    roi_base."""
    B, H, seed = 300, 96, 1234
    b = synth.tile_batch(synth.make_batch(30, H=8, seed=31, occlusion_max=0.5), B)
    g = _to_cuda(b)
    args = (g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"], g["extent"])
    kw = dict(region_idx=g["region_idx"], anchors=g["anchors"])
    solver = pose_solver.PoseSolver(inlier_thr=THR, num_hyp=H, seed=seed, want_inlier_mask=True, want_hyp=True)
    auto = solver(*args, None, **kw)
    auto = {k: getattr(auto, k).clone() for k in ("pose", "n_inliers", "status", "best_h", "n_sel", "inlier_mask", "hyp_counts")}
    # the oracle's gate and sampling
    hyp = np.zeros((B, H, 3), np.int32)
    for i in range(B):
        c = po.correspondences(b["depth"][i], b["Kp"][i], b["coor"][i], b["mask"][i], b["extent"][i], b["region_idx"][i],
                               b["anchors"][i])
        hyp[i] = po.sample_triplets(c["sel"], H, seed, i)
    expl = solver(*args, torch.from_numpy(hyp).cuda(), **kw)
    for k, v in auto.items():
        assert torch.equal(v, getattr(expl, k)), k
    assert float((auto["status"] == 0).float().mean()) > 0.8
    # a different seed draws different triplets
    other = pose_solver.PoseSolver(inlier_thr=THR, num_hyp=H, seed=seed + 1, want_hyp=True)(*args, None, **kw)
    assert not torch.equal(other.hyp_counts, auto["hyp_counts"])
    # shard [100, 300) with roi_base = 100 reproduces the tail of the full call
    sl = slice(100, 300)
    part = pose_solver.PoseSolver(inlier_thr=THR, num_hyp=H, seed=seed)(*[a[sl] for a in args], None,
                                                                       region_idx=g["region_idx"][sl], anchors=g["anchors"][sl],
                                                                       roi_base=100)
    assert torch.equal(part.pose.view(torch.int32), auto["pose"][sl].view(torch.int32))
    assert torch.equal(part.best_h, auto["best_h"][sl])
    # host call, chunks of 64 ROIs: chunking must not change the stream
    t = {k: (None if v is None else torch.from_numpy(np.ascontiguousarray(v)).pin_memory()) for k, v in b.items()}
    hs = pose_solver.HostPoseSolver(inlier_thr=THR, num_hyp=H, seed=seed, chunk_rois=64)
    res = hs(t["depth"], t["Kp"], t["coor"][:, 0].contiguous().pin_memory(), t["coor"][:, 1].contiguous().pin_memory(),
             t["coor"][:, 2].contiguous().pin_memory(), t["mask"], t["extent"], None, t["region_idx"], t["anchors"])
    assert torch.equal(res.pose.view(torch.int32), auto["pose"].cpu().view(torch.int32))
    assert torch.equal(res.best_h, auto["best_h"].cpu())
    hs.close()


def _ransac_roi_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ransac_roi_golden.npz"))
    b = {k: g[k] for k in ("depth", "Kp", "coor", "mask", "extent", "region_idx", "anchors")}
    return g, b


def test_s10_samples_reproduce_the_reference_ransac_loop(cuda, golden_dir):
    """misc.pnp_ransac_custom (misc.py:58-142) was run from source on the correspondences of four synthetic ROIs
    (oracle/gen_golden.py:gen_ransac_roi): 10 pairs per sample (misc.py:72,91), the reference's Kabsch on every solve,
    every point scored in float64, strict '<'.  The solver, fed the same planes and the same pixel sets
    (sample_size = 10), must reproduce the loop's inlier count of every iteration -- against the oracle under the
    usual bit-exactness rules, against the float64 reference loop up to boundary ties of the FP32 scoring."""
    g, b = _ransac_roi_golden(golden_dir)
    hyp, ref, iters, thr = g["hyp_idx"], g["counts"], g["iters"], float(g["thr"])
    assert hyp.shape[2] == 10
    ores = po.pose_solve_batch(b, hyp, thr)
    res = _solve(_to_cuda({**b, "hyp_idx": hyp}), inlier_thr=thr)
    cnt = res.hyp_counts.cpu().numpy()
    hp = res.hyp_poses.reshape(4, -1, 12).cpu().numpy()
    exact = total = 0
    for r in range(4):
        o, n_it = ores[r], int(iters[r])
        assert int(res.status[r]) == o["status"] and int(res.n_sel[r]) == o["n_sel"]
        assert int(res.best_h[r]) == o["best_h"] and int(res.n_inliers[r]) == o["n_inl"]
        assert np.array_equal(res.inlier_mask[r].reshape(-1).cpu().numpy(), o["inlier_mask"])
        assert po.re_rad_small(res.pose[r].cpu().numpy()[:, :3], o["pose"][:, :3]) <= ROT_TOL_RAD
        assert po.te(res.pose[r].cpu().numpy()[:, 3], o["pose"][:, 3]) <= TRANS_TOL_M
        assert o["valid"][:n_it].all() and not o["valid"][n_it:].any()
        np.testing.assert_allclose(hp[r], o["Rt_hyp"], atol=2e-6)
        same = (hp[r].view(np.uint32) == o["Rt_hyp"].view(np.uint32)).all(axis=1)
        assert np.array_equal(cnt[r][same], o["counts"][same])
        assert np.abs(cnt[r][~same].astype(int) - o["counts"][~same]).max(initial=0) <= 2
        assert same.mean() > 0.9
        # the reference loop itself
        d = cnt[r, :n_it].astype(int) - ref[r, :n_it]
        assert np.abs(d).max() <= 2, (r, d)
        assert (cnt[r, n_it:] == 0).all()
        exact += int((d == 0).sum())
        total += n_it
    assert exact >= 0.97 * total, (exact, total)


def test_s10_adaptive_stop_matches_the_reference_loop(cuda, golden_dir):
    """The reference loop stopped by itself after iters[r] iterations on the two clean ROIs (misc.py:134-138).  With
    further valid samples appended, the solver in adaptive mode must stop at the same iteration: its winner is the best
    of the first iters[r] hypotheses although a later one scores more."""
    g, b = _ransac_roi_golden(golden_dir)
    hyp, ref, iters, thr = g["hyp_idx"].copy(), g["counts"], g["iters"], float(g["thr"])
    for r in range(2):
        n_it = int(iters[r])
        assert n_it < 20
        top = int(np.argmax(ref[r, :n_it]))
        hyp[r, n_it:] = hyp[r, top]  # valid samples behind the stop
        # make a later hypothesis the global best by a margin: refine it to the ROI's consensus (oracle refit pose is
        # not expressible as a sample, so reuse the top sample: ties never displace the earlier winner, misc.py:121)
    sel = slice(0, 2)
    bb = {k: v[sel] for k, v in b.items()}
    ores = po.pose_solve_batch(bb, hyp[sel], thr, adaptive=True, confidence=0.995, min_iter=10)
    res = _solve(_to_cuda({**bb, "hyp_idx": hyp[sel]}), inlier_thr=thr, adaptive=True, confidence=0.995, min_iter=10)
    for r in range(2):
        n_it = int(iters[r])
        counts = ores[r]["counts"]
        _, examined = po.select_best(counts, ores[r]["valid"], ores[r]["n_sel"], 4, True, 0.995, 10)
        assert examined == n_it, (examined, n_it)  # the oracle stops where the reference loop stopped
        assert int(res.best_h[r]) == ores[r]["best_h"] < n_it
        assert int(res.n_inliers[r]) == ores[r]["n_inl"]


def test_min_mean_err_rule_returns_the_reference_loops_pose(cuda, golden_dir):
    """select_rule="min_mean_err": the pose misc.pnp_ransac_custom RETURNS (misc.py:113-132, 139-142), run from source on
    four whole ROIs and stored in the golden (ret_pose).  Same planes, same pixel sets (sample_size = 10), adaptive stop:
    the kernels return that pose within the north_star tolerance, and agree with the oracle's restatement of the rule
    on the source hypothesis and the best count."""
    g, b = _ransac_roi_golden(golden_dir)
    thr = float(g["thr"])
    kw = dict(select_rule="min_mean_err", adaptive=True, confidence=0.995, min_iter=10)
    ores = po.pose_solve_batch(b, g["hyp_idx"], thr, **kw)
    gd = _to_cuda({**b, "hyp_idx": g["hyp_idx"]})
    solver = pose_solver.PoseSolver(inlier_thr=thr, want_inlier_mask=True, **kw)
    res = solver(gd["depth"], gd["Kp"], gd["coor"][:, 0], gd["coor"][:, 1], gd["coor"][:, 2], gd["mask"], gd["extent"], gd["hyp_idx"],
                 region_idx=gd["region_idx"], anchors=gd["anchors"])
    pose = res.pose.cpu().numpy().astype(np.float64)
    for r, o in enumerate(ores):
        assert int(res.status[r]) == o["status"] == 0
        assert int(res.best_h[r]) == o["best_h"] and int(res.n_inliers[r]) == o["n_inl"], r
        assert np.array_equal(res.inlier_mask[r].reshape(-1).cpu().numpy(), o["inlier_mask"]), r
        assert po.re_rad_small(pose[r][:, :3], o["pose"][:, :3]) <= ROT_TOL_RAD and po.te(pose[r][:, 3], o["pose"][:, 3]) <= TRANS_TOL_M
        assert po.re_rad_small(pose[r][:, :3], g["ret_pose"][r][:, :3]) <= ROT_TOL_RAD, r   # the reference function's return value
        assert po.te(pose[r][:, 3], g["ret_pose"][r][:, 3]) <= TRANS_TOL_M, r
    # ... and on an ordinary batch (3-pair samples, no adaptive stop, weighted refit) against the oracle
    bb = synth.make_batch(24, H=128, seed=99, occlusion_max=0.5)
    o2 = po.pose_solve_batch(bb, bb["hyp_idx"], THR, select_rule="min_mean_err", weighted=True)
    g2 = _to_cuda(bb)
    r2 = pose_solver.PoseSolver(inlier_thr=THR, select_rule="min_mean_err", weighted=True)(
        g2["depth"], g2["Kp"], g2["coor"][:, 0], g2["coor"][:, 1], g2["coor"][:, 2], g2["mask"], g2["extent"], g2["hyp_idx"],
        region_idx=g2["region_idx"], anchors=g2["anchors"])
    p2 = r2.pose.cpu().numpy().astype(np.float64)
    for i, o in enumerate(o2):
        assert int(r2.status[i]) == o["status"], i
        if o["status"] == 0:
            assert int(r2.best_h[i]) == o["best_h"] and int(r2.n_inliers[i]) == o["n_inl"], i
            assert po.re_rad_small(p2[i][:, :3], o["pose"][:, :3]) <= ROT_TOL_RAD and po.te(p2[i][:, 3], o["pose"][:, 3]) <= TRANS_TOL_M, i


@pytest.mark.parametrize("S", [4, 10, 16])
def test_internal_sampling_with_larger_samples(cuda, S):
    """hyp_idx=None with sample_size S: kernel-drawn samples == oracle.sample_triplets(sample_size=S) fed back explicitly,
    on the device call and on the chunked host call.  Samples are drawn WITHOUT replacement (a repeated pixel is
    re-drawn from the same counter stream, as np.random.choice(replace=False) at misc.py:91), so repeats are rare; a
    sample that still repeats a pixel is an invalid hypothesis."""
    B, H, seed = 64, 64, 77
    b = synth.tile_batch(synth.make_batch(16, H=8, seed=41, occlusion_max=0.4), B)
    g = _to_cuda(b)
    args = (g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"], g["extent"])
    kw = dict(region_idx=g["region_idx"], anchors=g["anchors"])
    solver = pose_solver.PoseSolver(inlier_thr=THR, num_hyp=H, seed=seed, sample_size=S, want_inlier_mask=True, want_hyp=True)
    auto = solver(*args, None, **kw)
    auto = {k: getattr(auto, k).clone() for k in ("pose", "n_inliers", "status", "best_h", "n_sel", "inlier_mask", "hyp_counts")}
    hyp = np.zeros((B, H, S), np.int32)
    ndup = 0
    for i in range(B):
        c = po.correspondences(b["depth"][i], b["Kp"][i], b["coor"][i], b["mask"][i], b["extent"][i], b["region_idx"][i],
                               b["anchors"][i])
        hyp[i] = po.sample_triplets(c["sel"], H, seed, i, sample_size=S)
        srt = np.sort(hyp[i], axis=1)
        dup = (srt[:, 1:] == srt[:, :-1]).any(axis=1)
        ndup += int(dup.sum())
        assert (auto["hyp_counts"][i].cpu().numpy()[dup] == 0).all()  # repeated pixel: invalid, scores nothing
    assert ndup <= 2  # re-draws: practically every sample is usable (independent draws lost ~45 / n of the 10-pair samples)
    expl = solver(*args, torch.from_numpy(hyp).cuda(), **kw)
    for k, v in auto.items():
        assert torch.equal(v, getattr(expl, k)), k
    assert float((auto["status"] == 0).float().mean()) > 0.8
    # oracle parity on the explicit samples
    ores = po.pose_solve_batch({k: v[:16] for k, v in b.items() if v is not None}, hyp[:16], THR)
    for i in range(16):
        assert int(expl.best_h[i]) == ores[i]["best_h"] and int(expl.n_inliers[i]) == ores[i]["n_inl"]
        assert np.array_equal(expl.inlier_mask[i].reshape(-1).cpu().numpy(), ores[i]["inlier_mask"])
    # host call (pinned buffers, gated pull, chunks of 32 ROIs): explicit [B,H,S] samples and kernel-drawn ones
    t = {k: (None if v is None else torch.from_numpy(np.ascontiguousarray(v)).pin_memory()) for k, v in b.items()}
    hs = pose_solver.HostPoseSolver(inlier_thr=THR, num_hyp=H, seed=seed, sample_size=S, chunk_rois=32)
    hargs = (t["depth"], t["Kp"], t["coor"][:, 0].contiguous().pin_memory(), t["coor"][:, 1].contiguous().pin_memory(),
             t["coor"][:, 2].contiguous().pin_memory(), t["mask"], t["extent"])
    for h_in in (None, torch.from_numpy(hyp).pin_memory()):
        res = hs(*hargs, h_in, t["region_idx"], t["anchors"])
        assert torch.equal(res.pose.view(torch.int32), auto["pose"].cpu().view(torch.int32))
        assert torch.equal(res.best_h, auto["best_h"].cpu())
    hs.close()


def test_sample_size_out_of_range_is_rejected(cuda):
    b = synth.make_batch(2, H=8, seed=3)
    g = _to_cuda(b)
    with pytest.raises(ValueError):
        pose_solver.PoseSolver(sample_size=2)
    with pytest.raises(ValueError):
        pose_solver.PoseSolver()(g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"], g["extent"],
                                 torch.zeros(2, 8, 17, dtype=torch.int32, device="cuda"), region_idx=g["region_idx"],
                                 anchors=g["anchors"])
    L = _lib.lib()
    prm = _lib.SolveParams(inlier_thr=THR, num_hyp=8, min_pts=4, min_inliers=4, refit_iters=1, sample_size=17)
    inp = _lib.RoiInputs(depth=16, Kp=16, coor_x=16, coor_y=16, coor_z=16, mask=16, extent=16, B=1, mask_mode=1)
    out = _lib.SolveOutputs(pose=16, n_inliers=16, status=16)
    assert L.rdpn_pose_solve(ctypes.byref(inp), None, None, ctypes.byref(prm), ctypes.byref(out), None) == -1


def test_host_call_submit_wait_pipeline(cuda):
    """rdpn_pose_solve_host_submit / rdpn_ctx_wait: calls in flight together (own inputs, own outputs) give the results
    of the synchronous call, in any wait order, also when the ticket ring wraps (more than 8 submissions)."""
    B, H = 300, 32
    batches = [synth.tile_batch(synth.make_batch(20, H=H, seed=50 + i, occlusion_max=0.4), B) for i in range(3)]
    hs = pose_solver.HostPoseSolver(inlier_thr=THR, chunk_rois=64)

    def pinned(b):
        t = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items() if v is not None}
        cs = [t["coor"][:, c].contiguous().pin_memory() for c in range(3)]
        return (t["depth"], t["Kp"], cs[0], cs[1], cs[2], t["mask"], t["extent"], t["hyp_idx"], t["region_idx"], t["anchors"])

    args = [pinned(b) for b in batches]
    want = []
    for a in args:
        r = hs(*a)
        want.append((r.pose.clone(), r.best_h.clone(), r.n_inliers.clone(), r.status.clone()))
    plans = [hs.plan(*a, private_outputs=True) for a in args]
    outs = [p() for p in plans]  # the plans' own result objects
    assert len({o.pose.data_ptr() for o in outs}) == 3
    for order in ((0, 1, 2), (2, 0, 1)):
        for o in outs:  # results of a previous round must not leak into this one
            o.pose.zero_()
            o.best_h.zero_()
        tickets = [p.submit() for p in plans]
        for i in order:
            r = plans[i].wait(tickets[i])
            assert torch.equal(r.pose.view(torch.int32), want[i][0].view(torch.int32)), (order, i)
            assert torch.equal(r.best_h, want[i][1]) and torch.equal(r.n_inliers, want[i][2])
            assert torch.equal(r.status, want[i][3])
    # depth-2 loop over 20 steps (ring of 8 tickets wraps): every step's result is checked after its wait
    prev = None
    for step in range(20):
        i = step % 3
        if prev is not None and prev[0] == i:  # a plan must not be resubmitted while it is in flight
            plans[prev[0]].wait(prev[1])
            prev = None
        tk = plans[i].submit()
        if prev is not None:
            r = plans[prev[0]].wait(prev[1])
            assert torch.equal(r.pose.view(torch.int32), want[prev[0]][0].view(torch.int32)), step
        prev = (i, tk)
    r = plans[prev[0]].wait(prev[1])
    assert torch.equal(r.pose.view(torch.int32), want[prev[0]][0].view(torch.int32))
    # the synchronous call still reports its own bytes after submitted calls (counter stepped over)
    hs.set_option(_lib.OPT_COUNT_BYTES, 1)
    hs(*args[0])
    n0 = hs.last_h2d_bytes
    tk = plans[1].submit()
    plans[1].wait(tk)
    hs(*args[0])
    assert hs.last_h2d_bytes == n0 > 0
    L = _lib.lib()
    assert L.rdpn_ctx_wait(hs._ctx, 99) == -1
    hs.close()


@pytest.mark.parametrize("S", [5, 10])
def test_dense_mode_with_larger_samples(cuda, S):
    """Dense mode (no anchors: the object point is the de-normalised coordinate itself) with S pairs per hypothesis:
    the DENSE instantiation of the S > 3 path against the oracle, on samples drawn by the oracle's stream."""
    b = synth.make_batch(6, H=8, seed=58, dense=True)
    H = 96
    hyp = np.zeros((6, H, S), np.int32)
    for i in range(6):
        c = po.correspondences(b["depth"][i], b["Kp"][i], b["coor"][i], b["mask"][i], b["extent"][i])
        hyp[i] = po.sample_triplets(c["sel"], H, 9, i, sample_size=S)
    ores = po.pose_solve_batch(b, hyp, THR)
    g = _to_cuda({**b, "hyp_idx": hyp})
    res = _solve(g)
    cnt = res.hyp_counts.cpu().numpy()
    hp = res.hyp_poses.reshape(6, H, 12).cpu().numpy()
    for i, o in enumerate(ores):
        assert int(res.status[i]) == o["status"] and int(res.n_sel[i]) == o["n_sel"]
        assert int(res.best_h[i]) == o["best_h"] and int(res.n_inliers[i]) == o["n_inl"]
        assert np.array_equal(res.inlier_mask[i].reshape(-1).cpu().numpy(), o["inlier_mask"])
        np.testing.assert_allclose(hp[i], o["Rt_hyp"], atol=2e-6)
        same = (hp[i].view(np.uint32) == o["Rt_hyp"].view(np.uint32)).all(axis=1)
        assert np.array_equal(cnt[i][same], o["counts"][same])
        assert np.abs(cnt[i][~same].astype(int) - o["counts"][~same]).max(initial=0) <= 2
        assert same.mean() > 0.9
        assert po.re_rad_small(res.pose[i].cpu().numpy()[:, :3], o["pose"][:, :3]) <= ROT_TOL_RAD
        assert po.te(res.pose[i].cpu().numpy()[:, 3], o["pose"][:, 3]) <= TRANS_TOL_M
    # kernel-drawn samples are the same samples
    auto = pose_solver.PoseSolver(inlier_thr=THR, num_hyp=H, seed=9, sample_size=S, want_hyp=True)(
        g["depth"], g["Kp"], g["coor"][:, 0], g["coor"][:, 1], g["coor"][:, 2], g["mask"], g["extent"], None)
    assert torch.equal(auto.hyp_counts, res.hyp_counts) and torch.equal(auto.best_h, res.best_h)
    assert torch.equal(auto.pose.view(torch.int32), res.pose.view(torch.int32))
