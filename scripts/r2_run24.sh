#!/bin/bash
mkdir -p gpurun_out/r2y
PB_DUMP_KERNELS=gpurun_out/r2y/mm128_kernels.txt timeout 600 python bench.py --model mmformer --size 128 --batch 1 --no-cpu-baseline --no-extras --steps 8 > gpurun_out/r2y/mm128.json 2>gpurun_out/r2y/mm128.err
PB_DUMP_KERNELS=gpurun_out/r2y/mm80_kernels.txt timeout 600 python bench.py --model mmformer --no-cpu-baseline --no-extras --steps 8 > gpurun_out/r2y/mm80.json 2>gpurun_out/r2y/mm80.err
grep -E "attn|layernorm|linear" gpurun_out/r2y/mm128_kernels.txt
