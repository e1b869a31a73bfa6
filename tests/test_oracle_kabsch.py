"""CPU: the Kabsch/Umeyama restatement (oracle/pose_oracle.py:kabsch) against golden matrices produced
by the reference's lib/pysixd/transform.py (affine_matrix_from_points / superimposition_matrix), plus the
property checks SURVEY section 4 asks for (recover random (R,t); det<0 reflection case)."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import pose_oracle as po


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "kabsch_golden.npz"))


def test_doctest_literal_pins_reference_module(gold):
    # transform.py:893-898 -- the literal the reference's own doctest expects
    expect = np.array([[0.14549, 0.00062, 675.50008], [0.00048, 0.14094, 53.24971], [0.0, 0.0, 1.0]])
    np.testing.assert_allclose(gold["doctest_M"], expect, atol=6e-6)


def test_restatement_matches_reference_golden(gold):
    for name in gold["case_names"]:
        v0, v1, M, sc = gold[f"{name}_v0"], gold[f"{name}_v1"], gold[f"{name}_M"], bool(gold[f"{name}_scale"])
        mine = po.kabsch(v0, v1, scale=sc)
        np.testing.assert_allclose(mine, M, atol=1e-12, err_msg=str(name))
        np.testing.assert_allclose(po.superimposition_matrix(v0, v1, scale=sc), gold[f"{name}_Msup"], atol=1e-12)


def test_reflection_case_is_proper_rotation(gold):
    M = po.kabsch(gold["reflect_v0"], gold["reflect_v1"])
    assert np.linalg.det(M[:3, :3]) > 0.999


def test_wrong_shapes_raise():
    with pytest.raises(ValueError):
        po.kabsch(np.zeros((3, 2)), np.zeros((3, 2)))
    with pytest.raises(ValueError):
        po.kabsch(np.zeros((3, 5)), np.zeros((3, 6)))


@settings(max_examples=50, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(3, 200))
def test_recovers_random_rigid_motion(seed, n):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    t = rng.uniform(-1, 1, 3)
    a = rng.uniform(-0.3, 0.3, (3, n))
    if np.linalg.matrix_rank(a - a.mean(1, keepdims=True), tol=1e-6) < 2:
        return
    M = po.kabsch(a, R @ a + t[:, None])
    if np.linalg.matrix_rank(a - a.mean(1, keepdims=True), tol=1e-6) == 3 or n == 3:
        np.testing.assert_allclose(M[:3, :3] @ a + M[:3, 3:4], R @ a + t[:, None], atol=1e-9)


def test_weighted_equals_repeated_points():
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (3, 6))
    c = rng.uniform(-1, 1, (3, 6))
    w = np.array([1, 2, 1, 3, 1, 1.0])
    rep = np.repeat(np.arange(6), w.astype(int))
    np.testing.assert_allclose(po.kabsch(a, c, w=w), po.kabsch(a[:, rep], c[:, rep]), atol=1e-12)


def test_re_te_metrics():
    R = po.axangle2mat([0, 0, 1], 0.25)
    assert abs(po.re(R, np.eye(3)) - np.rad2deg(0.25)) < 1e-9
    assert abs(po.re_rad_small(R, np.eye(3)) - 0.25) < 1e-12
    assert abs(po.re_rad_small(po.axangle2mat([1, 2, 3], 1e-7), np.eye(3)) - 1e-7) < 1e-12
    assert po.te(np.array([1.0, 2, 3]), np.array([1.0, 2, 5])) == 2.0
