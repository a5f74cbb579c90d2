#!/bin/bash
# 2 x B200: the DDP equivalence check and the bench line with the end-of-round code
mkdir -p gpurun_out/r2final
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_equivalence.py > gpurun_out/r2final/ddp_equivalence.log 2>&1; echo "ddp_equivalence rc=$?"; tail -4 gpurun_out/r2final/ddp_equivalence.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2final/bench_2gpu.json 2> gpurun_out/r2final/bench_2gpu.err; echo "bench2 rc=$?"; head -c 400 gpurun_out/r2final/bench_2gpu.json
