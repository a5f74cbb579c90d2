#!/bin/bash
# End-of-round check on one B200 (round 2, after the token-path / 128^3 / small-weight-gradient work): full GPU suite, smoke(),
# default bench line (+ the reference arm), ncu launch list of one RFNet step.
OUT=gpurun_out/r2final; mkdir -p $OUT; : > $OUT/summary.txt
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > $OUT/gpu_tests.log 2>&1; echo "gpu_suite rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
PB_DUMP_KERNELS=$OUT/kernels.txt timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench_ref rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py > $OUT/launches.log 2>&1; echo "launch_list rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
python scripts/summarize_ncu.py launches $OUT/launches.csv > $OUT/launch_list_summary.txt 2>&1
cat $OUT/summary.txt; tail -3 $OUT/gpu_tests.log; tail -2 $OUT/smoke.log; head -c 1500 $OUT/bench.json; echo; head -c 600 $OUT/bench_ref.json; echo; head -12 $OUT/launch_list_summary.txt
