#!/bin/bash
OUT=gpurun_out/r2u; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "wgrad_rs" > $OUT/wgrs.log 2>&1; echo "wgrad_rs tests rc=$?"; tail -15 $OUT/wgrs.log
