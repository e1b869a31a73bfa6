"""Build librdpn6d_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m rdpn6d_b200.build            # incremental
    python -m rdpn6d_b200.build --force    # rebuild
    python -m rdpn6d_b200.build --ptxas    # print registers / spills / shared memory per kernel
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librdpn6d_b200.so")
SOURCES = ["fps.cu", "correspond.cu", "pose_solve.cu", "solve_pipe.cu", "geometry.cu", "roi_crop.cu", "coor_feat.cu",
           "host_api.cu"]


def _headers():
    """every csrc/*.cuh + the public header: editing any of them makes the library stale"""
    import glob

    return sorted(os.path.basename(h) for h in glob.glob(os.path.join(CSRC, "*.cuh"))) + [
        os.path.join("..", "..", "include", "rdpn6d_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + _headers()] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, ptxas=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library (file-locked: ranks may race)."""
    import fcntl

    if not (force or ptxas or _stale()):
        return LIB
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not (force or ptxas or _stale()):  # another process built it while we waited
                return LIB
            return _build_locked(ptxas, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(ptxas, verbose):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(CSRC, os.path.basename(s)[:-3] + ".o")
        extra = os.environ.get("RDPN_NVCC_EXTRA", "").split()  # e.g. -DRDPN_SOLVE_CTAS=3 for tuning runs
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if ptxas else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if ptxas or verbose:
            print(out)
    tmp = LIB + ".tmp.%d" % os.getpid()
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs + ["-ldl"]  # dl: NVTX 3
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, ptxas="--ptxas" in sys.argv, verbose=True))
