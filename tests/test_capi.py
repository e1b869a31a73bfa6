"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/rdpn6d_b200.h
declares; argument validation runs without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

from rdpn6d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rdpn6d_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
    src = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n.startswith("rdpn_") or n.startswith("farthest_point")))


def test_header_declares_the_expected_surface():
    names = _declared_functions()
    for must in ("farthest_point_sampling", "farthest_point_sampling_init_center", "rdpn_pose_solve",
                 "rdpn_correspond", "rdpn_fps_init_center", "rdpn_kabsch", "rdpn_pose_solve_host"):
        assert must in names
    assert len(names) >= 20


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in _declared_functions():
        assert hasattr(L, name), name
        assert name in _lib.SIGNATURES, "python binding missing for " + name


def test_struct_layouts_match_header_sizes():
    # x86-64 SysV: 10 pointers + 4 x 4-byte scalars; 10 x 4-byte; 9 pointers
    assert ctypes.sizeof(_lib.RoiInputs) == 10 * 8 + 16
    assert ctypes.sizeof(_lib.SolveParams) == 64
    assert ctypes.sizeof(_lib.SolveOutputs) == 10 * 8


def test_version_and_error_strings():
    L = _lib.lib()
    assert L.rdpn_version() == 200
    assert L.rdpn_error_string(0) == b"success"
    assert b"aligned" in L.rdpn_error_string(-2)
    assert L.rdpn_fps_workspace_bytes(512) >= 514 * 12 + 24
    assert L.rdpn_fps_workspace_bytes(512) % 256 == 0


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.rdpn_pose_solve(None, None, None, None, None, None) == -1
    inp = _lib.RoiInputs()
    assert L.rdpn_correspond(ctypes.byref(inp), None, None, None, None, None, None) == -1
    assert L.rdpn_kabsch(None, None, None, 10, 0, None, None, 1, None) == -1
    assert L.rdpn_kabsch(1, 1, None, 2, 0, 1, None, 1, None) == -1  # n < 3 (transform.py:917-918)
    assert L.rdpn_fps_init_center(None, None, 10, 2, None, 0, None) == -1
    assert L.rdpn_region_argmax(None, 32, None, 1, None) == -1
    # misaligned ROI planes are rejected, not silently copied
    inp = _lib.RoiInputs(depth=20, Kp=16, coor_x=16, coor_y=16, coor_z=16, mask=16, extent=16, B=1, mask_mode=1)
    assert L.rdpn_correspond(ctypes.byref(inp), 16, 16, 16, 16, 16, None) == -2


def test_sass_contains_bulk_tma():
    """The ROI staging must be the TMA bulk copy (UBLKCP in SASS), not a plain load loop."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass
    assert "sm_100a" in sass or "SM100" in sass.upper()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rdpn6d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_struct_field_offsets_match_the_header(tmp_path):
    """The ctypes mirrors in rdpn6d_b200/_lib.py against offsetof() of the C structs, compiled from the header with gcc:
    catches a field added or reordered on one side only."""
    import subprocess

    structs = {"rdpn_roi_inputs": _lib.RoiInputs, "rdpn_solve_params": _lib.SolveParams, "rdpn_solve_outputs": _lib.SolveOutputs}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "rdpn6d_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0;", "}"]
    src = tmp_path / "offsets.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "offsets"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call(["gcc", "-I", inc, str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_stale_library_is_an_error_not_a_silent_fallback(monkeypatch):
    """when the sources are newer than the shipped .so and the rebuild fails, loading must raise
    (a stale binary would silently run old kernels); RDPN_ALLOW_STALE_LIB=1 is the explicit opt-in."""
    import warnings

    from rdpn6d_b200 import build as _build

    def boom(*a, **k):
        raise RuntimeError("nvcc not found")

    monkeypatch.setattr(_build, "build", boom)
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.delenv("RDPN_ALLOW_STALE_LIB", raising=False)
    with pytest.raises(RuntimeError, match="rebuild failed"):
        _lib.lib()
    monkeypatch.setenv("RDPN_ALLOW_STALE_LIB", "1")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        L = _lib.lib()
    assert L.rdpn_version() >= 200 and any("STALE" in str(x.message) for x in w)
    monkeypatch.setattr(_lib, "_lib", None)  # the next test loads normally again


def test_ce_mask_mode_resolves_to_the_argmaxed_plane():
    """engine_utils.get_out_mask, MASK_LOSS_TYPE == "CE" (engine_utils.py:131-132): torch.argmax over the class channels.
    The tensor entries turn mask_mode="ce" into that plane in 'raw' mode (the gate's `> 0.5` then keeps class 1); the
    host-buffer entry, which only sees planes, says what to pass instead."""
    import pytest
    import torch

    from rdpn6d_b200 import pose_solver as ps

    x = torch.randn(3, 2, 64, 64)
    plane, mode = ps._resolve_mask(x, "ce")
    assert mode == ps.MASK_RAW and plane.dtype == torch.float32 and plane.shape == (3, 64, 64)
    assert torch.equal(plane, torch.argmax(x, dim=1, keepdim=True)[:, 0].float())
    assert torch.equal(plane > 0.5, x[:, 1] > x[:, 0])
    same, mode = ps._resolve_mask(x[:, 0], "L1")
    assert same is not plane and mode == ps.MASK_L1
    with pytest.raises(ValueError):
        ps._mask_mode("ce")


def test_bench_reference_arm_contract_and_no_cpu_fallback_of_the_gpu_arm():
    """bench.py --impl reference (the CPU implementation of the path on the host cores) prints ONE JSON line with the
    contract's keys; the default arm refuses to run without a CUDA device instead of falling back to the CPU."""
    import json
    import subprocess
    import sys

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ROI poses/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("ROI poses/sec") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline"],
                           capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr
