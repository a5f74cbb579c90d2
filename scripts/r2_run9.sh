#!/bin/bash
OUT=gpurun_out/r2i; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $OUT/full.log 2>&1; echo "full suite rc=$?" > $OUT/summary.txt; grep -E "passed|failed|FAILED|Error" $OUT/full.log | tail -12 >> $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt; tail -2 $OUT/smoke.log >> $OUT/summary.txt
PB_DUMP_KERNELS=$OUT/kernels.txt timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
cat $OUT/bench.json | head -c 6000
echo
cat $OUT/bench_ref.json
