#!/bin/bash
OUT=gpurun_out/r2u; mkdir -p $OUT
timeout 900 python -m pytest tests/test_mmformer_gpu.py -m gpu -q -p no:cacheprovider -k "cold_scratch" -s > $OUT/cold.log 2>&1; echo "rc=$?"; grep -E "graph vs eager|passed|failed|Error|assert" $OUT/cold.log | tail -6 | cut -c1-400
