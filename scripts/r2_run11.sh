#!/bin/bash
OUT=gpurun_out/r2k; mkdir -p $OUT
timeout 300 python scripts/debug_two_forwards.py > $OUT/two_fwd.log 2>&1; echo "two_forwards rc=$?" > $OUT/summary.txt; grep -v "Warn\|warn" $OUT/two_fwd.log | tail -5 >> $OUT/summary.txt
PB_BATCH_WEIGHTS=0 timeout 300 python scripts/debug_two_forwards.py > $OUT/two_fwd_b0.log 2>&1; echo "two_forwards(batch0) rc=$?" >> $OUT/summary.txt; grep -v "Warn\|warn" $OUT/two_fwd_b0.log | tail -5 >> $OUT/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "_k1_ or parameter_layout or single_modality" > $OUT/kern.log 2>&1; echo "kernels(k1) rc=$?" >> $OUT/summary.txt; tail -4 $OUT/kern.log >> $OUT/summary.txt
cat $OUT/summary.txt
