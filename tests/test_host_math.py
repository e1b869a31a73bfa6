"""CPU: the FP64 device math of csrc/kabsch_math.cuh (closed-form 3-pair Kabsch, Horn rotation via 4x4
Jacobi, triangle validity) compiled for the HOST with g++ and compared with the oracle (numpy SVD
Kabsch following lib/pysixd/transform.py:940-951) and with the reference-produced golden matrices."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import pose_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r"""
#include <cmath>
#define __device__
#define __forceinline__ inline
#define __noinline__
static inline double __dsub_rn(double a,double b){return a-b;}
static inline double __dadd_rn(double a,double b){return a+b;}
static inline double __dmul_rn(double a,double b){return a*b;}
static inline double rsqrt(double x){return 1.0/std::sqrt(x);}
#include "kabsch_math_host.cuh"
extern "C" {
void host_kabsch3(const double* a, const double* c, double* Rt){ rdpn::kabsch3((const double(*)[3])a,(const double(*)[3])c,Rt);}
void host_rot_from_cov(const double* S, double ga, double gb, double* R){ rdpn::rotation_from_cov(S,ga,gb,R);}
void host_rot_from_cov_jacobi(const double* S, double* R){ rdpn::rotation_from_cov_jacobi(S,R);}
int host_triangle_ok(const double* p0,const double*p1,const double*p2){return rdpn::triangle_ok(p0,p1,p2);}
}
"""
dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def hm():
    d = tempfile.mkdtemp(prefix="rdpn_hm_")
    src = open(os.path.join(ROOT, "rdpn6d_b200", "csrc", "kabsch_math.cuh")).read().replace("#include <cuda_runtime.h>", "")
    open(os.path.join(d, "kabsch_math_host.cuh"), "w").write(src)
    open(os.path.join(d, "shim.cpp"), "w").write(SHIM)
    so = os.path.join(d, "libhm.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", os.path.join(d, "shim.cpp"), "-o", so])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(dp)


def test_kabsch3_matches_svd_oracle(hm):
    rng = np.random.default_rng(1)
    worst = 0.0
    for i in range(1500):
        a = rng.uniform(-0.1, 0.1, (3, 3))
        c = rng.uniform(-0.1, 0.1, (3, 3)) + 0.5
        Rt = np.zeros(12)
        hm.host_kabsch3(_p(a), _p(c), _p(Rt))
        M = po.kabsch(a.T, c.T)
        worst = max(worst, np.abs(Rt.reshape(3, 4) - M[:3, :4]).max())
    assert worst < 1e-11


def test_kabsch3_on_reference_golden_flipped_triangle(hm, golden_dir):
    g = np.load(os.path.join(golden_dir, "kabsch_golden.npz"))
    for name in ("rigid0", "reflect_tri"):
        a = np.ascontiguousarray(g[name + "_v0"].T)
        c = np.ascontiguousarray(g[name + "_v1"].T)
        Rt = np.zeros(12)
        hm.host_kabsch3(_p(a), _p(c), _p(Rt))
        np.testing.assert_allclose(Rt.reshape(3, 4), g[name + "_M"][:3, :4], atol=1e-11)


def test_rotation_from_cov_matches_reference_golden(hm, golden_dir):
    g = np.load(os.path.join(golden_dir, "kabsch_golden.npz"))
    for name in g["case_names"]:
        if bool(g[f"{name}_scale"]):
            continue
        a, c, M = g[f"{name}_v0"], g[f"{name}_v1"], g[f"{name}_M"]
        ac = a - a.mean(1, keepdims=True)
        cc = c - c.mean(1, keepdims=True)
        S = np.ascontiguousarray(cc @ ac.T)
        for fn in ("qcp", "jacobi"):
            R = np.zeros(9)
            if fn == "qcp":
                hm.host_rot_from_cov(_p(S), ctypes.c_double((ac * ac).sum()), ctypes.c_double((cc * cc).sum()), _p(R))
            else:
                hm.host_rot_from_cov_jacobi(_p(S), _p(R))
            np.testing.assert_allclose(R.reshape(3, 3), M[:3, :3], atol=1e-10, err_msg=str(name) + fn)


def test_rotation_from_cov_random_incl_reflection_and_planar(hm):
    rng = np.random.default_rng(2)
    worst = 0.0
    for i in range(1500):
        n = int(rng.integers(3, 50))
        a = rng.uniform(-0.1, 0.1, (3, n))
        if i % 3 == 0:
            c = np.diag([1, 1, -1.0]) @ a + rng.normal(0, 1e-3, (3, n))
        elif i % 3 == 1:
            c = rng.uniform(-1, 1, (3, n))
        else:
            a[2] = 0
            c = po.kabsch(rng.uniform(-1, 1, (3, 5)), rng.uniform(-1, 1, (3, 5)))[:3, :3] @ a + 0.3
        M = po.kabsch(a, c)
        ac, cc = a - a.mean(1, keepdims=True), c - c.mean(1, keepdims=True)
        S = np.ascontiguousarray(cc @ ac.T)
        R = np.zeros(9)
        hm.host_rot_from_cov(_p(S), ctypes.c_double((ac * ac).sum()), ctypes.c_double((cc * cc).sum()), _p(R))
        worst = max(worst, np.abs(R.reshape(3, 3) - M[:3, :3]).max())
        R2 = np.zeros(9)
        hm.host_rot_from_cov_jacobi(_p(S), _p(R2))
        worst = max(worst, np.abs(R2.reshape(3, 3) - M[:3, :3]).max())
    assert worst < 1e-9


def test_rotation_from_cov_degenerate_inputs_do_not_produce_nan(hm):
    """Collinear points / zero covariance: the rotation is not unique, but the result must be a finite
    proper rotation (Jacobi fallback)."""
    rng = np.random.default_rng(4)
    for kind in ("collinear", "zero", "identical"):
        if kind == "collinear":
            tt = rng.uniform(-1, 1, 20)
            a = np.outer([0.3, -0.2, 0.9], tt)
            c = np.outer([0.1, 0.8, 0.2], tt)
        elif kind == "zero":
            a = np.zeros((3, 5)); c = np.zeros((3, 5))
        else:
            a = rng.uniform(-1, 1, (3, 10)); c = a.copy()
        ac, cc = a - a.mean(1, keepdims=True), c - c.mean(1, keepdims=True)
        S = np.ascontiguousarray(cc @ ac.T)
        R = np.zeros(9)
        hm.host_rot_from_cov(_p(S), ctypes.c_double((ac * ac).sum()), ctypes.c_double((cc * cc).sum()), _p(R))
        R = R.reshape(3, 3)
        assert np.isfinite(R).all()
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-9)
        assert abs(np.linalg.det(R) - 1) < 1e-9
        if kind == "identical":
            np.testing.assert_allclose(R, np.eye(3), atol=1e-9)


def test_triangle_rule_is_bitwise_the_oracles(hm):
    rng = np.random.default_rng(3)
    p = rng.uniform(-1, 1, (4000, 3, 3))
    p[::5, 2] = p[::5, 0] + (p[::5, 1] - p[::5, 0]) * rng.uniform(0.2, 2, (800, 1)) + rng.normal(0, 1e-4, (800, 3))  # near-collinear
    p[::7, 1] = p[::7, 0]  # duplicate vertex
    ok_o = po._triangle_ok(p[:, 0], p[:, 1], p[:, 2])
    ok_h = np.array([hm.host_triangle_ok(_p(np.ascontiguousarray(q[0])), _p(np.ascontiguousarray(q[1])), _p(np.ascontiguousarray(q[2]))) for q in p], bool)
    assert np.array_equal(ok_o, ok_h)
    assert 0 < ok_o.sum() < len(p)
