#!/bin/bash
OUT=gpurun_out/r2u; mkdir -p $OUT
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "trajectory_50" -s > $OUT/traj.log 2>&1; echo "trajectory rc=$?"; grep -E "trajectory|passed|failed|assert|Error" $OUT/traj.log | tail -8
