#!/bin/bash
# sweep of the InstanceNorm traversal orders (PB_INORM_ORDER bit mask): kernel parity + step time per setting
mkdir -p gpurun_out; OUT=gpurun_out; : > $OUT/sweep.txt
for o in ${ORDERS:-0 1 4 5 2 3 7}; do
  PB_INORM_ORDER=$o timeout 120 python -m pytest tests/test_kernels_gpu.py -q -k inorm > $OUT/inorm_test_$o.log 2>&1; rc=$?
  PB_INORM_ORDER=$o timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_o$o.json 2> $OUT/bench_o$o.err
  python - <<PY >> $OUT/sweep.txt
import json
try:
    d = json.load(open("$OUT/bench_o$o.json")); f = d["roofline"]["families_ms_per_step"]
    print("order $o test_rc $rc ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "resident", (d.get("e2e_resident_cases") or {}).get("ms_per_step"), "in_fwd", f.get("inorm_lrelu_fwd"), "in_bwd", f.get("inorm_lrelu_bwd"), "loss", d["e2e"]["last_loss"])
except Exception as e:
    print("order $o test_rc $rc bench failed", e)
PY
done
cat $OUT/sweep.txt
