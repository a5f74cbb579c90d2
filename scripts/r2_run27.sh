#!/bin/bash
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 train.py --use_passion --model rfnet --batch_size 2 --synthetic --num_epochs 1 --iters_per_epoch 8 --savepath /tmp/rf2 2>&1 | grep -E "Iter 8/8|rp_epoch|Error|error" | cut -c25-190
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 train.py --use_passion --batch_size 2 --synthetic --num_epochs 1 --iters_per_epoch 8 --savepath /tmp/mm2 2>&1 | grep -E "Iter 8/8|rp_epoch|Error|error" | cut -c25-190
