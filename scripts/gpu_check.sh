#!/bin/bash
# One gpurun call (single B200): parity of the device sample pipeline, the bench line, a short --device_aug training run,
# an ncu capture of the augment kernel, then as much of the full GPU suite as the time budget allows.
# usage: gpurun --timeout 840 -- 'bash scripts/gpu_check.sh 780'
BUDGET=${1:-780}
T0=$(date +%s)
left() { echo $(( BUDGET - ( $(date +%s) - T0 ) )); }
mkdir -p gpurun_out
OUT=gpurun_out
: > $OUT/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/smi.txt 2>&1
timeout 300 python -m pytest tests/test_augment_gpu.py -q --maxfail=8 > $OUT/aug_tests.log 2>&1; echo "aug_tests rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
timeout 400 python bench.py --steps 8 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
timeout 150 python train.py --use_passion --batch_size 2 --synthetic --num_epochs 1 --iters_per_epoch 6 --device_aug --savepath /tmp/train_out > $OUT/train_aug.log 2>&1; echo "train_device_aug rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
timeout 150 ncu --set full --clock-control none -k regex:augment_kernel -c 2 --csv --page raw --log-file $OUT/ncu_augment.csv python -m pytest tests/test_augment_gpu.py -q -k full_size > $OUT/ncu.log 2>&1; echo "ncu rc=$? t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
L=$(left)
if [ "$L" -gt 60 ]; then
  timeout $L python -m pytest tests -m gpu -q --maxfail=5 --durations=12 --deselect tests/test_augment_gpu.py > $OUT/gpu_tests.log 2>&1; echo "gpu_suite rc=$? (124 = ran out of the call's time budget) t=$(( $(date +%s) - T0 ))" >> $OUT/summary.txt
fi
cat $OUT/summary.txt
tail -3 $OUT/aug_tests.log
head -c 1500 $OUT/bench.json
tail -5 $OUT/gpu_tests.log 2>/dev/null
