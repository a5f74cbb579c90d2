#!/bin/bash
OUT=gpurun_out/r2n; mkdir -p $OUT
PB_DUMP_KERNELS=$OUT/kernels.txt timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 16 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" > $OUT/summary.txt
timeout 300 python scripts/profile_glue.py > $OUT/glue.txt 2> $OUT/glue.err; echo "glue rc=$?" >> $OUT/summary.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2n/bench.json").read()); print(d["ms_per_step"], d["gpu_launches"]//16, d["roofline"]["families_ms_per_step"])
P
