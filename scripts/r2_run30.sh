#!/bin/bash
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "c2+0_2 or c4+0_4 or c1+0_8" 2>&1 | tail -4 | cut -c1-250
for a in 0 1 0 1; do
PB_SMALL_WGRAD_ASYNC=$a PB_DUMP_KERNELS=gpurun_out/r2y/rf_$a.txt timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('async=$a', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['families_ms_per_step'].get('conv3d_small_wgrad'))" | tee -a gpurun_out/r2y/small_ab.txt
done
grep small_wgrad gpurun_out/r2y/rf_0.txt gpurun_out/r2y/rf_1.txt | tee -a gpurun_out/r2y/small_ab.txt
