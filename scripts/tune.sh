#!/bin/bash
# Kernel tuning harness.  Here (no GPU):   scripts/tune.sh build NAME "-DGRB_X=1 ..."   compiles a variant of the
# library into gorender_b200/lib/variants/NAME.so.  On the GPU box:  scripts/tune.sh run [config]  times every
# variant (and the default build) with scripts/kernel_times.py and checks a few parity scenes against the oracle.
set -e
cd "$(dirname "$0")/.."
case "$1" in
build)
    make -s -C gorender_b200/csrc VARIANT="$2" EXTRA="$3" 2>&1 | grep -E "error|raster_kernelILb0|setup_kernelILb0ELb0" -A2 | grep -E "error|registers|spill" || true
    ;;
run)
    cfg="${2:-c3}"
    for lib in gorender_b200/lib/libgorender_b200.so gorender_b200/lib/variants/*.so; do
        [ -f "$lib" ] || continue
        echo "== $lib"
        GORENDER_B200_LIB="$PWD/$lib" timeout 120 python scripts/kernel_times.py "$cfg" --reps 20 2>&1 | tail -1
        GORENDER_B200_LIB="$PWD/$lib" timeout 200 python -m pytest tests/test_parity_gpu.py -q -x -k "framebuffer_bit_exact or c5_pose" 2>&1 | tail -1
    done
    ;;
esac
