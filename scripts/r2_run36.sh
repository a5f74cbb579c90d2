#!/bin/bash
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "_s2_" 2>&1 | tail -6 | cut -c1-300
for a in 1 2; do
PB_DUMP_KERNELS=gpurun_out/r2y/ds3_$a.txt timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); f=d['roofline']['families_ms_per_step']; print('run $a', d['ms_per_step'], d['e2e']['ms_per_step'], f.get('conv3d_dgrad'))" | tee -a gpurun_out/r2y/ds3_ab.txt
done
grep "conv3d_dgrad |" gpurun_out/r2y/ds3_1.txt | tee -a gpurun_out/r2y/ds3_ab.txt
