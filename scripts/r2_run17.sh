#!/bin/bash
OUT=gpurun_out/r2q; mkdir -p $OUT
PB_WG_RS_MIN_VOX=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "test_conv3d" > $OUT/conv_tests.log 2>&1; echo "conv tests rc=$?" > $OUT/summary.txt
grep -E "passed|failed|FAILED|Error" $OUT/conv_tests.log | tail -30 >> $OUT/summary.txt
cat $OUT/summary.txt
MODES=0,1 timeout 300 python scripts/bench_wgrad.py 2>&1 | cut -c1-110 | tee $OUT/wgrad_i4.txt
timeout 200 python scripts/probe_wgrad_issuer.py
