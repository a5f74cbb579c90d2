#!/bin/bash
OUT=gpurun_out/r2q; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "test_conv3d" > $OUT/conv_tests.log 2>&1; echo "conv tests rc=$?" > $OUT/summary.txt
grep -E "passed|failed|FAILED|Error" $OUT/conv_tests.log | tail -12 >> $OUT/summary.txt
cat $OUT/summary.txt
MODES=0,1 timeout 600 python scripts/bench_wgrad.py 2>&1 | tee $OUT/wgrad_diag4.txt
echo "modes 2 3"; MODES=2,3 timeout 600 python scripts/bench_wgrad.py 2>&1 | head -6
