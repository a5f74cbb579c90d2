#!/bin/bash
OUT=gpurun_out/r2v; mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench2.json 2> $OUT/bench2.err; echo "bench2 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2v/bench2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/ddp_equivalence.py > $OUT/ddp_eq.txt 2>&1; echo "ddp_equivalence rc=$?"; tail -12 $OUT/ddp_eq.txt
