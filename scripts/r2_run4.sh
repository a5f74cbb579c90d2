#!/bin/bash
OUT=gpurun_out/r2d; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > $OUT/full.log 2>&1; echo "full suite rc=$?" > $OUT/summary.txt; grep -E "passed|failed|idt64|FAILED|Error" $OUT/full.log | tail -12 >> $OUT/summary.txt
PB_DUMP_KERNELS=$OUT/kernels_sp1.txt timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_sp1.json 2> $OUT/bench_sp1.err; echo "bench sparse1 rc=$?" >> $OUT/summary.txt
PB_SPARSE_SINGLES=0 PB_DUMP_KERNELS=$OUT/kernels_sp0.txt timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_sp0.json 2> $OUT/bench_sp0.err; echo "bench sparse0 rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
python - <<'P'
import json
for t in ("sp1","sp0"):
    try:
        d=json.loads(open(f"gpurun_out/r2d/bench_{t}.json").read()); print(t, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["families_ms_per_step"])
    except Exception as e: print(t, "ERR", e)
P
tail -5 $OUT/bench_sp1.err
