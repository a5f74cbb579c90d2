#!/bin/bash
OUT=gpurun_out/r2g; mkdir -p $OUT
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "conv3d" > $OUT/kern.log 2>&1; echo "kernels rc=$?" > $OUT/summary.txt; tail -4 $OUT/kern.log >> $OUT/summary.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_model_full_gpu.py tests/test_mmformer_gpu.py tests/test_predict_gpu.py -q -p no:cacheprovider -k "bf16 or side or sweep or predict" > $OUT/model.log 2>&1; echo "model rc=$?" >> $OUT/summary.txt; tail -3 $OUT/model.log >> $OUT/summary.txt
for m in 2 1 0; do
PB_TC_TMA=$m PB_DUMP_KERNELS=$OUT/kernels_tma$m.txt timeout 600 python bench.py --no-cpu-baseline --steps 16 > $OUT/bench_tma$m.json 2> $OUT/bench_tma$m.err; echo "bench tma$m rc=$?" >> $OUT/summary.txt
done
cat $OUT/summary.txt
python - <<'P'
import json
for t in ("tma2","tma1","tma0"):
    try:
        d=json.loads(open(f"gpurun_out/r2g/bench_{t}.json").read()); f=d["roofline"]["families_ms_per_step"]; print(t, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]//16, {k:f[k] for k in f if "tc" in k or "fold" in k})
    except Exception as e: print(t, "ERR", e)
P
grep "conv3d_fwd_tc" $OUT/kernels_tma2.txt | head -6; echo; grep "conv3d_fwd_tc" $OUT/kernels_tma1.txt | head -6
