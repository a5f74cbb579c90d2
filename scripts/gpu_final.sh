#!/bin/bash
# Final round check on one B200: full GPU suite, the default bench line, the ncu launch list and two --set full captures.
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/final_summary.txt
T0=$(date +%s)
timeout 200 python -m pytest tests -m gpu -q --maxfail=5 > $OUT/final_gpu_tests.log 2>&1; echo "gpu_suite rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/final_summary.txt
PB_DUMP_KERNELS=$OUT/final_kernels.txt timeout 150 python bench.py > $OUT/final_bench.json 2> $OUT/final_bench.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/final_summary.txt
timeout 120 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/final_launches.csv python scripts/profile_step.py > $OUT/final_launches.log 2>&1; echo "launch_list rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/final_summary.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct
timeout 90 ncu --profile-from-start off --metrics $M --clock-control none -k regex:"up_fwd_kernel|bwd_reduce_kernel|bwd_apply_kernel|conv_dgrad_pairs_kernel|conv3_small_wgrad_kernel|conv_wgrad_kernel" --csv --log-file $OUT/final_kernels_ncu.csv python scripts/profile_step.py > $OUT/final_ncu.log 2>&1; echo "ncu_kernels rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/final_summary.txt
cat $OUT/final_summary.txt; tail -2 $OUT/final_gpu_tests.log; head -c 600 $OUT/final_bench.json
