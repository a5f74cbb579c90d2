#!/bin/bash
OUT=gpurun_out/r2z; mkdir -p $OUT
N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 16 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench$N.json 2> $OUT/bench$N.err; echo "bench$N rc=$?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2z/bench$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
