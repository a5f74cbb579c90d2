#!/bin/bash
mkdir -p gpurun_out/r2y
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "stay_on_tcgen05 or c64+64_64_k3 or c128+0_64_k3" 2>&1 | tail -25 | cut -c1-300
for i in 1 2; do
timeout 600 python bench.py --model mmformer --size 128 --batch 1 --no-cpu-baseline --no-extras --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mmformer 128^3', d['ms_per_step'], d['value'], 'last_loss', d['e2e']['last_loss'])" | tee -a gpurun_out/r2y/ab2.log
done
