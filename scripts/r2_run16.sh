#!/bin/bash
OUT=gpurun_out/r2p; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "test_conv3d" > $OUT/conv_tests.log 2>&1; echo "conv tests rc=$?" > $OUT/summary.txt
grep -E "passed|failed|FAILED|Error" $OUT/conv_tests.log | tail -12 >> $OUT/summary.txt
timeout 600 python scripts/bench_wgrad.py > $OUT/wgrad.txt 2>&1; echo "bench_wgrad rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt; cat $OUT/wgrad.txt
