#!/bin/bash
OUT=gpurun_out/r2e; mkdir -p $OUT
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_model_full_gpu.py -q -p no:cacheprovider > $OUT/model.log 2>&1; echo "model rc=$?" > $OUT/summary.txt; tail -2 $OUT/model.log >> $OUT/summary.txt
for sp in 0 1 2 4; do
PB_SPARSE_SINGLES=$sp timeout 600 python bench.py --no-cpu-baseline --steps 16 > $OUT/bench_sp$sp.json 2> $OUT/bench_sp$sp.err; echo "bench sparse$sp rc=$?" >> $OUT/summary.txt
done
PB_PW2=1 timeout 600 python bench.py --no-cpu-baseline --steps 16 > $OUT/bench_pw2.json 2> $OUT/bench_pw2.err; echo "bench pw2 rc=$?" >> $OUT/summary.txt
PB_SEP_STREAM=0 timeout 600 python bench.py --no-cpu-baseline --steps 16 > $OUT/bench_nosep.json 2> $OUT/bench_nosep.err; echo "bench nosep rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
python - <<'P'
import json
for t in ("sp0","sp1","sp2","sp4","pw2","nosep"):
    try:
        d=json.loads(open(f"gpurun_out/r2e/bench_{t}.json").read()); f=d["roofline"]["families_ms_per_step"]; print(t, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]//16, {k:f[k] for k in ("conv1_fwd","conv1_dgrad","conv1_wgrad","conv1_wgrad_tc","masked_stack_fwd","masked_stack_bwd")})
    except Exception as e: print(t, "ERR", e)
P
