#!/bin/bash
OUT=gpurun_out/r2t; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $OUT/full.log 2>&1; echo "full suite rc=$?" > $OUT/summary.txt; grep -E "passed|failed|FAILED|Error" $OUT/full.log | tail -12 >> $OUT/summary.txt
PB_DUMP_KERNELS=$OUT/kernels.txt timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 16 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/summary.txt
cat $OUT/summary.txt
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2t/bench.json").read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]//16, d["roofline"]["families_ms_per_step"])
P
MODES=0,1 timeout 300 python scripts/bench_wgrad.py > $OUT/wgrad_classes.txt 2>&1; tail -1 $OUT/wgrad_classes.txt
PB_WG_RS_MIN_VOX=0 PB_WG_RS_RU=4 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "test_conv3d" 2>&1 | tail -1
