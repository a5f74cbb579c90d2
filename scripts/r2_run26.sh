#!/bin/bash
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_token_path_gpu.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30 | cut -c1-250 | tee gpurun_out/r2y/token2.log
timeout 300 python scripts/bench_attn.py 2>&1 | tail -60 | tee gpurun_out/r2y/bench_attn_fused.txt
