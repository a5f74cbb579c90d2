#!/bin/bash
OUT=gpurun_out/r2c; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "conv3d" > $OUT/kern.log 2>&1; echo "kernels rc=$?" > $OUT/summary.txt; tail -4 $OUT/kern.log >> $OUT/summary.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_model_full_gpu.py -q -p no:cacheprovider -s -k "bf16 or side" > $OUT/model.log 2>&1; echo "model rc=$?" >> $OUT/summary.txt; grep -E "passed|failed|idt64" $OUT/model.log | tail -8 >> $OUT/summary.txt
PB_DUMP_KERNELS=$OUT/kernels_kws1.txt timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_kws1.json 2> $OUT/bench_kws1.err; echo "bench kws1 rc=$?" >> $OUT/summary.txt
PB_TC_KWS=0 PB_DUMP_KERNELS=$OUT/kernels_kws0.txt timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_kws0.json 2> $OUT/bench_kws0.err; echo "bench kws0 rc=$?" >> $OUT/summary.txt
bash scripts/ncu_capture.sh r2c/ncu_kws "conv3_tc_kws" 0 4 > $OUT/ncu_kws.txt 2>&1
PB_TC_KWS=0 bash scripts/ncu_capture.sh r2c/ncu_old "conv3_tc_kernel" 0 4 > $OUT/ncu_old.txt 2>&1
cat $OUT/summary.txt
python - <<'P'
import json
for t in ("kws1","kws0"):
    try:
        d=json.loads(open(f"gpurun_out/r2c/bench_{t}.json").read()); print(t, d["ms_per_step"], d["gpu_launches"], d["roofline"]["families_ms_per_step"])
    except Exception as e: print(t, "ERR", e)
P
grep "conv3d_fwd_tc\|conv3d_dgrad_tc" $OUT/kernels_kws1.txt | head -12
