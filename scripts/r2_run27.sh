#!/bin/bash
mkdir -p gpurun_out/r2y
timeout 900 python -m pytest tests/test_mmformer_gpu.py tests/test_token_path_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | cut -c1-250 | tee gpurun_out/r2y/mmformer3.log
for a in 0 1 0 1; do
  PB_ATTN_TC=$a timeout 600 python bench.py --model mmformer --size 128 --batch 1 --no-cpu-baseline --no-extras --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('attn_tc=$a mmformer 128^3', d['ms_per_step'], d['value'], 'last_loss', d['e2e']['last_loss'])" | tee -a gpurun_out/r2y/ab3.log
done
for a in 0 1; do
  PB_ATTN_TC=$a timeout 600 python bench.py --model mmformer --no-cpu-baseline --no-extras --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('attn_tc=$a mmformer 80^3', d['ms_per_step'], d['value'], 'last_loss', d['e2e']['last_loss'])" | tee -a gpurun_out/r2y/ab3.log
done
timeout 300 python scripts/bench_attn.py 2>&1 | tail -60 > gpurun_out/r2y/bench_attn_fused.txt
