#!/bin/bash
OUT=gpurun_out/r2u; mkdir -p $OUT
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "test_conv3d" 2>&1 | tail -2
for r in 1 2; do
PB_DUMP_KERNELS=$OUT/kernels1.txt timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); f=d['roofline']['families_ms_per_step']; print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:f[k] for k in ('conv3d_fwd_tc','conv3d_dgrad_tc','conv1_fwd_tc','conv1_dgrad_tc')})"
done
grep "^conv3d_fwd_tc\|^conv3d_dgrad_tc" $OUT/kernels1.txt | head -6
timeout 200 python scripts/probe_conv_tc.py 2>&1 | tail -3 | cut -c1-260
