#!/bin/bash
OUT=gpurun_out/r2l; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 scripts/debug_ddp_sink.py > $OUT/dbg.log 2>&1; echo "rc=$?"
grep "before\|after" $OUT/dbg.log
