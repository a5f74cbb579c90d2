#!/bin/bash
# round-2 GPU call 1: root-cause check of the round-1 mmFormer failure (cross-stream block reuse), full suite, poison run, bench
OUT=gpurun_out/r2a; mkdir -p $OUT
T="tests/test_augment_gpu.py tests/test_kernels_gpu.py tests/test_mmformer_gpu.py::test_fp32_check_mode"
echo "== A: old behaviour (no record_stream, point upsample), in-suite order" > $OUT/summary.txt
PB_SIDE_RECORD=0 PB_UP_POINT=1 timeout 400 python -m pytest $T -q -p no:cacheprovider > $OUT/A_old.log 2>&1; echo "A rc=$?" >> $OUT/summary.txt; tail -3 $OUT/A_old.log >> $OUT/summary.txt
echo "== B: record_stream fix, point upsample" >> $OUT/summary.txt
PB_SIDE_RECORD=1 PB_UP_POINT=1 timeout 400 python -m pytest $T -q -p no:cacheprovider > $OUT/B_fix.log 2>&1; echo "B rc=$?" >> $OUT/summary.txt; tail -3 $OUT/B_fix.log >> $OUT/summary.txt
echo "== C: no record_stream but single stream" >> $OUT/summary.txt
PB_SIDE_RECORD=0 PB_SEP_STREAM=0 PB_UP_POINT=1 timeout 400 python -m pytest $T -q -p no:cacheprovider > $OUT/C_nosep.log 2>&1; echo "C rc=$?" >> $OUT/summary.txt; tail -3 $OUT/C_nosep.log >> $OUT/summary.txt
echo "== D: full suite, defaults (fix + row-walking fp32 upsample)" >> $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > $OUT/D_full.log 2>&1; echo "D rc=$?" >> $OUT/summary.txt; tail -5 $OUT/D_full.log >> $OUT/summary.txt
echo "== E: full suite with poisoned allocator" >> $OUT/summary.txt
PB_POISON=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/E_poison.log 2>&1; echo "E rc=$?" >> $OUT/summary.txt; tail -5 $OUT/E_poison.log >> $OUT/summary.txt
echo "== F: smoke" >> $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/F_smoke.log 2>&1; echo "F rc=$?" >> $OUT/summary.txt; tail -3 $OUT/F_smoke.log >> $OUT/summary.txt
echo "== G: bench" >> $OUT/summary.txt
timeout 600 python bench.py > $OUT/G_bench.json 2> $OUT/G_bench.err; echo "G rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
head -c 3000 $OUT/G_bench.json
