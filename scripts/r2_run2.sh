#!/bin/bash
OUT=gpurun_out/r2b; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "conv3d" > $OUT/kern.log 2>&1; echo "kernels rc=$?" > $OUT/summary.txt; tail -15 $OUT/kern.log >> $OUT/summary.txt
timeout 600 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -k "bf16 or side" >> $OUT/model.log 2>&1; echo "model rc=$?" >> $OUT/summary.txt; tail -5 $OUT/model.log >> $OUT/summary.txt
PB_DUMP_KERNELS=$OUT/kernels_kws1.txt timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_kws1.json 2> $OUT/bench_kws1.err; echo "bench kws1 rc=$?" >> $OUT/summary.txt
PB_TC_KWS=0 PB_DUMP_KERNELS=$OUT/kernels_kws0.txt timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_kws0.json 2> $OUT/bench_kws0.err; echo "bench kws0 rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
python - <<'P'
import json
for t in ("kws1","kws0"):
    try:
        d=json.loads(open(f"gpurun_out/r2b/bench_{t}.json").read()); print(t, d["ms_per_step"], d["roofline"]["families_ms_per_step"])
    except Exception as e: print(t, "ERR", e)
P
head -30 $OUT/kernels_kws1.txt
