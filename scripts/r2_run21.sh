#!/bin/bash
OUT=gpurun_out/r2u; mkdir -p $OUT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_mmformer_gpu.py -m gpu -q -p no:cacheprovider > $OUT/k.log 2>&1; echo "kernel+mmformer tests rc=$?"; tail -2 $OUT/k.log
timeout 300 python scripts/bench_gemm.py 2>&1 | tail -10 | tee $OUT/gemm_shapes.txt
for v in 1 0; do
echo "PB_LINEAR_TC=$v"
PB_LINEAR_TC=$v timeout 600 python bench.py --model mmformer --size 128 --batch 1 --no-cpu-baseline --no-extras --steps 8 2> $OUT/mm128_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('128^3 B=1', d['ms_per_step'], d['value'])"
done
