#!/bin/bash
OUT=gpurun_out/r2h; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/full.log 2>&1; echo "full suite rc=$?" > $OUT/summary.txt; grep -E "passed|failed|FAILED|Error" $OUT/full.log | tail -12 >> $OUT/summary.txt
for m in 1 0; do
PB_FUSED_LOSS=$m PB_DUMP_KERNELS=$OUT/kernels_fl$m.txt timeout 600 python bench.py --no-cpu-baseline --steps 16 > $OUT/bench_fl$m.json 2> $OUT/bench_fl$m.err; echo "bench fused_loss$m rc=$?" >> $OUT/summary.txt
done
cat $OUT/summary.txt
python - <<'P'
import json
for t in ("fl1","fl0"):
    try:
        d=json.loads(open(f"gpurun_out/r2h/bench_{t}.json").read()); f=d["roofline"]["families_ms_per_step"]; print(t, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]//16, {k:f[k] for k in f if any(w in k for w in ("softmax","kl","cedice","logit","upsample"))})
    except Exception as e: print(t, "ERR", e)
P
