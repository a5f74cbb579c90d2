#!/bin/bash
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "proto or criterion or fp32_check or bf16" 2>&1 | tail -4 | cut -c1-250
for a in 1 2; do
PB_DUMP_KERNELS=gpurun_out/r2y/pr_$a.txt timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); f=d['roofline']['families_ms_per_step']; print('run $a', d['ms_per_step'], d['e2e']['ms_per_step'], f.get('proto_bwd1'), f.get('proto_fwd'))" | tee -a gpurun_out/r2y/pr_ab.txt
done
