#!/bin/bash
OUT=gpurun_out/r2j; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/ddp_equivalence.py --size 32 > $OUT/ddp_eq.log 2>&1; echo "ddp_equivalence rc=$?" > $OUT/summary.txt
grep "ddp_equivalence\|^  " $OUT/ddp_eq.log | grep -v "Warn\|warn\|File\|raise\|^    \|r_l =" | head -14 >> $OUT/summary.txt
cat $OUT/summary.txt
