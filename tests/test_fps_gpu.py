"""GPU: FPS kernel (csrc/fps.cu through the C ABI) bit-exact against the reference-produced golden
indices, the C restatement and -- when oracle/_ref travelled -- the reference's own build."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import libfps_ref
from oracle.fps import fps_indices_port, fps_indices_reference
from rdpn6d_b200 import _lib, fps_utils
from rdpn6d_b200.synth import fps_cloud

pytestmark = pytest.mark.gpu
SMALL = ["gauss_2000", "lattice_1000", "dups_40", "n_eq_k_50", "n_lt_k_10", "cube_8", "single_1"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "fps_golden.npz"))


def _gpu_idx(pts, k, **kw):
    t = torch.from_numpy(np.ascontiguousarray(pts, np.float32)).cuda()
    return fps_utils.fps_indices(t, k, **kw).cpu().numpy()


@pytest.mark.parametrize("name", SMALL)
def test_golden_small(cuda, gold, name):
    pts, idx = gold[name + "_pts"], gold[name + "_idx"]
    assert np.array_equal(_gpu_idx(pts, len(idx)), idx)


@pytest.mark.parametrize("n,k,seed", [(200_000, 64, 3), (1_000_000, 8, 0), (1_000_000, 64, 0), (1_000_000, 512, 0)])
def test_golden_full_size(cuda, gold, n, k, seed):
    """BASELINE config 4: 1M-point cloud to 8/64/512 keypoints, bit-exact vs the reference C++."""
    idx = gold[f"seeded_{n}_{k}_{seed}_idx"]
    assert np.array_equal(_gpu_idx(fps_cloud(n, seed=seed), k), idx)


@pytest.mark.parametrize("n,k", [(1, 1), (2, 5), (31, 31), (513, 40), (1025, 7), (4097, 100), (70_000, 33), (300_001, 16)])
def test_vs_port_random_sizes(cuda, n, k):
    rng = np.random.default_rng(n * 31 + k)
    p = rng.standard_normal((n, 3)).astype(np.float32)
    assert np.array_equal(_gpu_idx(p, k), fps_indices_port(p, k))


def test_ties_duplicates_lattice(cuda):
    rng = np.random.default_rng(5)
    lat = rng.integers(0, 5, (20_000, 3)).astype(np.float32)  # 125 distinct positions, massive ties
    assert np.array_equal(_gpu_idx(lat, 300), fps_indices_port(lat, 300))


def test_from_index_variant(cuda):
    p = fps_cloud(50_000, seed=8)
    for start in (0, 7, 49_999):
        assert np.array_equal(_gpu_idx(p, 20, init_center=False, start=start), fps_indices_port(p, 20, start=start))


@pytest.mark.skipif(libfps_ref() is None, reason="reference FPS build (oracle/_ref) not present on this box")
def test_vs_live_reference_build(cuda):
    p = fps_cloud(123_457, seed=11)
    assert np.array_equal(_gpu_idx(p, 48), fps_indices_reference(p, 48))


def test_python_surface_numpy_and_tensor(cuda, gold):
    """fps_utils.farthest_point_sampling keeps the reference wrapper's contract (fps_utils.py:6-21)."""
    pts, idx = gold["gauss_2000_pts"], gold["gauss_2000_idx"]
    out = fps_utils.farthest_point_sampling(pts.astype(np.float64), 64, init_center=True)
    assert out.dtype == np.float32 and out.shape == (64, 3)
    assert np.array_equal(out, pts[idx])
    out_t = fps_utils.farthest_point_sampling(torch.from_numpy(pts).cuda(), 64, init_center=True)
    assert out_t.is_cuda and np.array_equal(out_t.cpu().numpy(), pts[idx])
    rnd = fps_utils.farthest_point_sampling(pts, 8, init_center=False)  # random start: valid rows of pts
    assert rnd.shape == (8, 3)
    fc = fps_utils.get_fps_and_center(pts.astype(np.float64), 8)
    assert fc.shape == (9, 3) and np.allclose(fc[-1], pts.astype(np.float64).mean(0))
    fct = fps_utils.get_fps_and_center(torch.from_numpy(pts).cuda(), 8)
    assert np.allclose(fct.cpu().numpy(), fc, atol=1e-9)


def test_c_abi_dropin_symbols_host_pointers(cuda, gold):
    """The exact ext.h entry points: host pointers in, indices out (what the reference's cffi binds)."""
    L = _lib.lib()
    pts = np.ascontiguousarray(gold["lattice_1000_pts"], np.float32)
    idx = np.zeros(200, np.int32)
    L.farthest_point_sampling_init_center(pts.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p), len(pts), 200)
    assert np.array_equal(idx, gold["lattice_1000_idx"])


def test_bad_arguments_fail_loudly(cuda):
    L = _lib.lib()
    t = torch.zeros(10, 3, device="cuda")
    idx = torch.zeros(4, dtype=torch.int32, device="cuda")
    assert L.rdpn_fps_init_center(t.data_ptr(), idx.data_ptr(), 10, 4, None, 0, None) == -3  # workspace
    with pytest.raises(AssertionError):
        fps_utils.fps_indices(torch.zeros(10, 2, device="cuda"), 2)


def test_get_fps_and_center_matches_reference_python_surface(cuda, golden_dir):
    """The numpy face of fps_utils.get_fps_and_center against the reference's own Python surface run from source
    (tests/golden/fps_center_golden.npz): same rows, same dtype."""
    g = np.load(os.path.join(golden_dir, "fps_center_golden.npz"))
    for name in ("f64", "f32"):
        for n in (8, 32):
            want = g["%s_fps%d_and_center" % (name, n)]
            got = fps_utils.get_fps_and_center(g[name + "_pts"], n)
            assert got.dtype == want.dtype and got.shape == want.shape
            assert np.array_equal(got, want)


def test_cluster_path_sizes_and_cooperative_path_agree(cuda, monkeypatch):
    """<= 65 536 points run in one thread-block cluster (candidates pushed through DSMEM); the same clouds through the cooperative grid
    (RDPN_FPS_NO_CLUSTER) and the C restatement give the same indices, bit for bit."""
    for n, k in ((100, 10), (5_000, 32), (8_192, 64), (8_193, 64), (16_385, 24), (50_000, 64), (32_768, 40), (32_769, 40), (65_536, 16), (65_537, 16)):
        p = np.random.default_rng(n + k).standard_normal((n, 3)).astype(np.float32)
        a = _gpu_idx(p, k)
        assert np.array_equal(a, fps_indices_port(p, k)), (n, k)
        monkeypatch.setenv("RDPN_FPS_NO_CLUSTER", "1")
        assert np.array_equal(_gpu_idx(p, k), a), (n, k)
        monkeypatch.delenv("RDPN_FPS_NO_CLUSTER")
        monkeypatch.setenv("RDPN_FPS_EXCHANGE", "barrier")  # the cluster-barrier exchange the push exchange replaced
        assert np.array_equal(_gpu_idx(p, k), a), (n, k)
        monkeypatch.delenv("RDPN_FPS_EXCHANGE")
        for ct in ("128", "256", "512"):  # every CTA width of the cluster kernel
            monkeypatch.setenv("RDPN_FPS_CLUSTER_THREADS", ct)
            assert np.array_equal(_gpu_idx(p, k), a), (n, k, ct)
            monkeypatch.delenv("RDPN_FPS_CLUSTER_THREADS")


def test_batched_objects_one_launch(cuda):
    """rdpn_fps_batch: every object of a model set in one launch (tools/lm/1_compute_fps.py:26-35 loops object by
    object) equals the per-object runs -- init_center and explicit starts, ragged sizes incl. tiny clouds."""
    rng = np.random.default_rng(11)
    sizes = [37, 5_000, 12_345, 1, 64_000, 9_000, 513, 65_536]
    clouds = [rng.standard_normal((n, 3)).astype(np.float32) * rng.uniform(0.05, 0.2, 3).astype(np.float32) for n in sizes]
    dev = [torch.from_numpy(c).cuda() for c in clouds]
    before = _lib.launch_count()
    idx = fps_utils.fps_indices_batch(dev, 48).cpu().numpy()
    assert _lib.launch_count() - before == 1
    for c, row in zip(clouds, idx):
        assert np.array_equal(row, fps_indices_port(c, 48))
    starts = [0, 17, 12_344, 0, 63_999, 5, 512, 65_535]
    idx = fps_utils.fps_indices_batch(dev, 20, starts=starts).cpu().numpy()
    for c, row, s0 in zip(clouds, idx, starts):
        assert np.array_equal(row, fps_indices_port(c, 20, start=s0))


def test_streaming_path_beyond_register_limit(cuda):
    """More than 148 x 512 x 16 points: cloud and running minima streamed from global memory, same indices."""
    p = fps_cloud(1_300_000, seed=4)
    assert np.array_equal(_gpu_idx(p, 12), fps_indices_port(p, 12))
    assert np.array_equal(_gpu_idx(p, 5, init_center=False, start=1_299_999), fps_indices_port(p, 5, start=1_299_999))


def test_center_row_is_deterministic(cuda):
    """get_fps_and_center's mean row: fixed-order FP64 reduction -- identical bits run after run, equal to numpy."""
    t = torch.from_numpy(fps_cloud(300_000, seed=2)).cuda()
    a = fps_utils.get_fps_and_center(t, 8)
    for _ in range(5):
        assert torch.equal(fps_utils.get_fps_and_center(t, 8), a)
    ref = t.double().mean(0).cpu().numpy()
    np.testing.assert_allclose(a[-1].cpu().numpy(), ref, rtol=0, atol=1e-12)
