#!/bin/bash
mkdir -p gpurun_out/r2y
PB_DUMP_KERNELS=gpurun_out/r2y/rfnet_kernels.txt timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 8 > gpurun_out/r2y/rfnet.json 2>gpurun_out/r2y/rfnet.err
head -70 gpurun_out/r2y/rfnet_kernels.txt
