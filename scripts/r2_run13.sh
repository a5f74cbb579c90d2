#!/bin/bash
OUT=gpurun_out/r2m; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider > $OUT/kern.log 2>&1; echo "kernels rc=$?" > $OUT/summary.txt; tail -6 $OUT/kern.log >> $OUT/summary.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_model_full_gpu.py tests/test_mmformer_gpu.py tests/test_predict_gpu.py tests/test_augment_gpu.py -q -p no:cacheprovider > $OUT/model.log 2>&1; echo "model rc=$?" >> $OUT/summary.txt; tail -4 $OUT/model.log >> $OUT/summary.txt
for m in 1 0; do
PB_C1_TC=$m PB_DUMP_KERNELS=$OUT/kernels_c1$m.txt timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 16 > $OUT/bench_c1$m.json 2> $OUT/bench_c1$m.err; echo "bench c1tc$m rc=$?" >> $OUT/summary.txt
done
cat $OUT/summary.txt
python - <<'P'
import json
for t in ("c11","c10"):
    try:
        d=json.loads(open(f"gpurun_out/r2m/bench_{t}.json").read()); f=d["roofline"]["families_ms_per_step"]; print(t, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]//16, {k:f[k] for k in f if "conv1" in k})
    except Exception as e: print(t, "ERR", e)
P
grep "^conv1_fwd\|^conv1_dgrad" $OUT/kernels_c11.txt | head -16
