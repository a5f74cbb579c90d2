#!/bin/bash
# round-2 launch list of one eager step (same command as round 1) for profiles/r02_final_launches.csv
OUT=gpurun_out/r2w; mkdir -p $OUT
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/final_launches.csv python scripts/profile_step.py > $OUT/final_launches.log 2>&1; echo "launch_list rc=$?"
wc -l $OUT/final_launches.csv
