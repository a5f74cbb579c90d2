#!/bin/bash
# ncu --set full capture of the row-stacked weight gradient (first launch of scripts/bench_wgrad.py = c16->8, 80^3, n10) + launch list of one bench step
OUT=gpurun_out/r2s; mkdir -p $OUT
MODES=1 REPS=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3_wgrad_rs --launch-skip 0 -c 1 -o $OUT/rs -f python scripts/bench_wgrad.py > $OUT/ncu.log 2>&1
ncu -i $OUT/rs.ncu-rep --page raw --csv > $OUT/rs.raw.csv 2>/dev/null
ncu -i $OUT/rs.ncu-rep --page details --csv > $OUT/rs.details.csv 2>/dev/null
rm -f $OUT/rs.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $OUT/launch_bench.log 2>&1
ls -la $OUT; tail -3 $OUT/ncu.log; wc -l $OUT/launches.csv
