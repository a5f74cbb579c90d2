#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests: memcheck by default, TOOL=racecheck|initcheck|synccheck for the others.
#   scripts/sanitize.sh ["pytest -k expression"]        (on a B200 box; 10-50x slower than a plain run — keep the selection small)
TOOL=${TOOL:-memcheck}
SEL=${1:-"wgrad_rs or linear_tc or test_inorm or test_upsample"}
OUT=gpurun_out/sanitize_$TOOL.log; mkdir -p gpurun_out
timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $TOOL --error-exitcode 77 --launch-timeout 120 \
    python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > $OUT 2>&1
rc=$?
echo "compute-sanitizer --tool $TOOL rc=$rc (77 = errors reported)"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|Uninitialized|hazard" $OUT | tail -12
