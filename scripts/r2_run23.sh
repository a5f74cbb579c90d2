#!/bin/bash
# mmFormer with the token path on our attention / LayerNorm kernels: parity tests, then the step A/B against the library attention
mkdir -p gpurun_out/r2y
timeout 900 python -m pytest tests/test_mmformer_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 | cut -c1-250 | tee gpurun_out/r2y/mmformer.log
for a in 0 1 0 1; do
  PB_ATTN_TC=$a timeout 600 python bench.py --model mmformer --no-cpu-baseline --no-extras --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('attn_tc=$a mmformer 80^3', d['ms_per_step'], d['value'], 'last_loss', d['e2e']['last_loss'])" | tee -a gpurun_out/r2y/ab.log
done
for a in 0 1; do
  PB_ATTN_TC=$a timeout 600 python bench.py --model mmformer --size 128 --batch 1 --no-cpu-baseline --no-extras --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('attn_tc=$a mmformer 128^3', d['ms_per_step'], d['value'], 'last_loss', d['e2e']['last_loss'])" | tee -a gpurun_out/r2y/ab.log
done
