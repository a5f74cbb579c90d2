#!/bin/bash
OUT=gpurun_out/r2b; mkdir -p $OUT
PB_BATCH_WEIGHTS=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "conv3d" > $OUT/kern_b0.log 2>&1; echo "kernels(batch0) rc=$?" > $OUT/summary.txt; tail -4 $OUT/kern_b0.log >> $OUT/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider > $OUT/kern_b1.log 2>&1; echo "kernels(batch1) rc=$?" >> $OUT/summary.txt; tail -4 $OUT/kern_b1.log >> $OUT/summary.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_model_full_gpu.py tests/test_predict_gpu.py tests/test_mmformer_gpu.py -q -p no:cacheprovider -s > $OUT/model.log 2>&1; echo "model rc=$?" >> $OUT/summary.txt; grep -E "passed|failed|idt64" $OUT/model.log | tail -8 >> $OUT/summary.txt
PB_DUMP_KERNELS=$OUT/kernels_kws1.txt timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_kws1.json 2> $OUT/bench_kws1.err; echo "bench kws1 rc=$?" >> $OUT/summary.txt
PB_TC_KWS=0 PB_DUMP_KERNELS=$OUT/kernels_kws0.txt timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_kws0.json 2> $OUT/bench_kws0.err; echo "bench kws0 rc=$?" >> $OUT/summary.txt
PB_BATCH_WEIGHTS=0 timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_b0.json 2> $OUT/bench_b0.err; echo "bench batch0 rc=$?" >> $OUT/summary.txt
PB_PW2=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "_k1_" > $OUT/kern_pw2.log 2>&1; echo "kernels(pw2) rc=$?" >> $OUT/summary.txt; tail -3 $OUT/kern_pw2.log >> $OUT/summary.txt
PB_PW2=1 timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_pw2.json 2> $OUT/bench_pw2.err; echo "bench pw2 rc=$?" >> $OUT/summary.txt
timeout 300 python scripts/profile_glue.py > $OUT/glue.txt 2> $OUT/glue.err; echo "glue rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
python - <<'P'
import json
for t in ("kws1","kws0","b0","pw2"):
    try:
        d=json.loads(open(f"gpurun_out/r2b/bench_{t}.json").read()); print(t, d["ms_per_step"], d["gpu_launches"], d["roofline"]["families_ms_per_step"])
    except Exception as e: print(t, "ERR", e)
P
head -24 $OUT/kernels_kws1.txt
