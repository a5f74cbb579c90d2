#!/bin/bash
OUT=gpurun_out/r2r; mkdir -p $OUT; rm -f $OUT/probe2.txt
for cfg in "64 72 1 0" "64 72 3 0" "64 72 1 1" "64 72 3 1" "64 32 3 0" "128 64 3 0"; do set -- $cfg
echo "M=$1 N=$2 ALT=$3 ONE=$4" >> $OUT/probe2.txt
PB_WG_RS_M=$1 PB_WG_RS_N=$2 PB_WG_RS_ALT=$3 PB_WG_RS_ONE=$4 MODES=1,4 REPS=3 timeout 300 python scripts/bench_wgrad.py 2>&1 | head -1 >> $OUT/probe2.txt
done
cat $OUT/probe2.txt
