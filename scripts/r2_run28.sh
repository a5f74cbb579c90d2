#!/bin/bash
# channel_stats with 8 loads in flight: unit test, per-shape timing before / after, mmFormer step before / after
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "channel_stats" 2>&1 | tail -4 | cut -c1-250
cp passion_b200/libpassion_b200.so build/libCur.so
for v in before Cur; do
cp build/lib$v.so passion_b200/libpassion_b200.so 2>/dev/null || cp build/lib_before.so passion_b200/libpassion_b200.so
echo "== $v"; timeout 300 python scripts/bench_stats.py 2>&1 | tail -12
done | tee gpurun_out/r2y/stats_ab.txt
for v in before Cur before Cur; do
cp build/lib$v.so passion_b200/libpassion_b200.so 2>/dev/null || cp build/lib_before.so passion_b200/libpassion_b200.so
timeout 600 python bench.py --model mmformer --size 128 --batch 1 --no-cpu-baseline --no-extras --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v mmformer 128^3', d['ms_per_step'], d['value'])" | tee -a gpurun_out/r2y/stats_ab.txt
done
cp build/libCur.so passion_b200/libpassion_b200.so
