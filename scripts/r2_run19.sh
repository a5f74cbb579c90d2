#!/bin/bash
OUT=gpurun_out/r2s; mkdir -p $OUT
MODES=1 REPS=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3_wgrad_rs --launch-skip 0 -c 1 -o $OUT/rs -f python scripts/bench_wgrad.py > $OUT/ncu.log 2>&1
ncu -i $OUT/rs.ncu-rep --page raw --csv > $OUT/rs.raw.csv 2>/dev/null
ncu -i $OUT/rs.ncu-rep --page details --csv > $OUT/rs.details.csv 2>/dev/null
ncu -i $OUT/rs.ncu-rep --page source --csv --print-source=sass > $OUT/rs.source.csv 2>/dev/null
rm -f $OUT/rs.ncu-rep
ls -la $OUT; tail -5 $OUT/ncu.log
