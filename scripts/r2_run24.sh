#!/bin/bash
OUT=gpurun_out/r2x; mkdir -p $OUT
T0=$(date +%s)
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default bench rc=$? t=$(( $(date +%s) - T0 ))s"
T1=$(date +%s)
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "reference arm rc=$? t=$(( $(date +%s) - T1 ))s"
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2x/bench_default.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","vs_baseline")}, d["e2e"]["value"], d.get("cpu_baseline"), d.get("parity"))
r=d["roofline"]; print({k:v for k,v in r.items() if k not in ("families_ms_per_step",)})
print(d.get("e2e_resident_cases")); print({k:v for k,v in d.items() if k.startswith("config") or k in ("configs3","configs4","extras")})
d=json.loads(open("gpurun_out/r2x/bench_ref.json").read().strip().splitlines()[-1]); print(d)
P
