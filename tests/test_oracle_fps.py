"""CPU: the FPS restatement (oracle/fps_oracle.c) against the reference's own C++ build and the
golden indices that build produced (tests/golden/fps_golden.npz, oracle/gen_golden.py)."""
import os

import numpy as np
import pytest

from oracle import libfps_ref
from oracle.fps import fps_indices_port, fps_indices_reference, get_fps_and_center
from rdpn6d_b200.synth import fps_cloud

SMALL = ["gauss_2000", "lattice_1000", "dups_40", "n_eq_k_50", "n_lt_k_10", "cube_8", "single_1"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "fps_golden.npz"))


@pytest.mark.parametrize("name", SMALL)
def test_port_matches_golden_small(gold, name):
    pts, idx = gold[name + "_pts"], gold[name + "_idx"]
    assert np.array_equal(fps_indices_port(pts, len(idx)), idx)


@pytest.mark.parametrize("n,k,seed", [(200_000, 64, 3), (1_000_000, 8, 0), (1_000_000, 64, 0)])
def test_port_matches_golden_seeded(gold, n, k, seed):
    idx = gold[f"seeded_{n}_{k}_{seed}_idx"]
    assert np.array_equal(fps_indices_port(fps_cloud(n, seed=seed), k), idx)


def test_reference_edge_semantics(gold):
    # all-equal distances -> index order; exhausted clouds repeat index 0; sn > pn pads with 0
    assert list(gold["cube_8_idx"]) == list(range(8))
    d = gold["dups_40_idx"]
    assert (d[10:] == 0).all()
    assert (gold["n_lt_k_10_idx"][10:] == 0).all()
    assert list(gold["single_1_idx"]) == [0, 0, 0]


@pytest.mark.skipif(libfps_ref() is None, reason="reference FPS build not available")
def test_port_matches_live_reference():
    rng = np.random.default_rng(123)
    for n, k in [(1, 1), (2, 5), (17, 17), (333, 40), (5000, 128)]:
        p = rng.standard_normal((n, 3)).astype(np.float32)
        assert np.array_equal(fps_indices_port(p, k), fps_indices_reference(p, k))
    lat = rng.integers(0, 4, (500, 3)).astype(np.float32)  # heavy ties and duplicates
    assert np.array_equal(fps_indices_port(lat, 100), fps_indices_reference(lat, 100))


def test_fps_invariants():
    p = fps_cloud(3000, seed=9)
    idx = fps_indices_port(p, 50)
    ctr = 0.5 * (p.max(0) + p.min(0))
    d0 = ((p - ctr) ** 2).sum(1)
    assert d0[idx[0]] == d0.max()  # first pick is farthest from the bbox centre
    assert len(set(idx.tolist())) == 50
    md = np.full(len(p), np.inf)
    picked = []
    for i in idx[:-1]:
        md = np.minimum(md, ((p - p[i]) ** 2).sum(1))
        picked.append(md.max())
    assert all(a >= b - 1e-12 for a, b in zip(picked, picked[1:]))  # selected distances never increase


def test_from_index_variant():
    p = fps_cloud(500, seed=2)
    idx = fps_indices_port(p, 10, start=7)
    assert idx[0] == 7
    d = ((p - p[7]) ** 2).sum(1)
    assert idx[1] == int(np.argmax(d))


def test_get_fps_and_center_shape():
    p = fps_cloud(400, seed=4).astype(np.float64)
    out = get_fps_and_center(p, 8)
    assert out.shape == (9, 3)
    np.testing.assert_allclose(out[-1], p.mean(0))


def test_get_fps_and_center_matches_reference_python_surface(golden_dir):
    """a12: fps_utils.farthest_point_sampling (core/csrc/fps/fps_utils.py:6-21) + data_utils.get_fps_and_center
    (core/utils/data_utils.py:217-226) executed from source on the reference's own C++ build
    (oracle/gen_golden.py:gen_fps_center): same rows, same dtype (float64 input -> float32-rounded samples in a float64
    array with a float64 centre; float32 input -> float32)."""
    g = np.load(os.path.join(golden_dir, "fps_center_golden.npz"))
    for name in ("f64", "f32"):
        for n in (8, 32):
            want = g["%s_fps%d_and_center" % (name, n)]
            got = get_fps_and_center(g[name + "_pts"], n)
            assert got.dtype == want.dtype and got.shape == want.shape
            assert np.array_equal(got, want)
