#!/bin/bash
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_token_path_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 | cut -c1-250 | tee gpurun_out/r2y/token.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k linear 2>&1 | tail -5 | cut -c1-250
