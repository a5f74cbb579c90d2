#!/bin/bash
# usage: scripts/ncu_capture.sh <name> <kernel-regex> <launch-skip> <count>
# One `ncu --set full` capture of scripts/profile_step.py, exported to CSV on the box (raw metrics + per-source-line
# hot spots); the .ncu-rep itself is deleted so that gpurun_out/ stays small.
name=$1; regex=$2; skip=${3:-0}; cnt=${4:-2}
out=gpurun_out/$name
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$regex" \
    --launch-skip "$skip" -c "$cnt" -o "$out" -f python scripts/profile_step.py > "$out.log" 2>&1
ncu -i "$out.ncu-rep" --page raw --csv > "$out.raw.csv" 2>/dev/null
ncu -i "$out.ncu-rep" --page details --csv > "$out.details.csv" 2>/dev/null
ncu -i "$out.ncu-rep" --page source --csv --print-source=cuda > "$out.source.csv" 2>/dev/null
rm -f "$out.ncu-rep"
ls -la "$out".* | awk '{print $5, $9}'
