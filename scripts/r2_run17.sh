#!/bin/bash
OUT=gpurun_out/r2q; mkdir -p $OUT
PB_WG_RS_MIN_VOX=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "test_conv3d" > $OUT/conv_tests.log 2>&1; echo "conv tests rc=$?" > $OUT/summary.txt
grep -E "passed|failed|FAILED|Error" $OUT/conv_tests.log | tail -30 >> $OUT/summary.txt
cat $OUT/summary.txt
export PB_WG_RS_MIN_VOX=0 MODES=1 REPS=5
for cfg in "ROWS=8" "ROWS=12" "ROWS=16" "ROWS=8 PB_WG_RS_PREFETCH=2" "ROWS=8 PB_WG_RS_ND=5" "ROWS=8 PB_WG_RS_ND=8" "ROWS=8 PB_WG_RS_ND=10" "ROWS=8 PB_WG_RS_REGIONS=1"; do
echo "== $cfg"; env PB_WG_RS_$cfg timeout 300 python scripts/bench_wgrad.py 2>&1 | head -10 | cut -c1-70
done
