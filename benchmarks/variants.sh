#!/bin/bash
# Tuning helper: build the library once per value of a compile-time knob (here, in the container) ...
#   bash benchmarks/variants.sh build RDPN_SCORE_SPLIT 0 1 2 3
# ... and time every variant on the GPU box in one call:
#   gpurun -- 'bash benchmarks/variants.sh run "python benchmarks/kernels.py --only solve --no-cpu"'
# Variants live in gpurun_variants/ (git-ignored *.so, shipped to the box with the snapshot).
set -u
cd "$(dirname "$0")/.."
V=gpurun_variants
if [ "$1" = build ]; then
    knob=$2; shift 2
    mkdir -p $V
    for val in "$@"; do
        RDPN_NVCC_EXTRA="-D$knob=$val" python -m rdpn6d_b200.build --force > /dev/null || exit 1
        cp rdpn6d_b200/librdpn6d_b200.so $V/lib_${knob}_$val.so
    done
    python -m rdpn6d_b200.build --force > /dev/null   # leave the default build in place
else
    cmd=$2
    cp rdpn6d_b200/librdpn6d_b200.so /tmp/lib_default.so
    for f in $V/lib_*.so; do
        echo "== $f"
        cp $f rdpn6d_b200/librdpn6d_b200.so
        eval "$cmd"
    done
    cp /tmp/lib_default.so rdpn6d_b200/librdpn6d_b200.so
fi
