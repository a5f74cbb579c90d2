#!/bin/bash
timeout 300 python train.py --use_passion --batch_size 2 --synthetic --num_epochs 2 --iters_per_epoch 8 --savepath /tmp/mm 2>&1 | grep -E "Iter 8/8|rp_epoch" | cut -c25-190
timeout 300 python train.py --use_passion --model rfnet --batch_size 2 --synthetic --num_epochs 1 --iters_per_epoch 8 --savepath /tmp/rf 2>&1 | grep -E "Iter 8/8|rp_epoch" | cut -c25-190
timeout 300 python train.py --use_passion --batch_size 2 --synthetic --num_epochs 1 --iters_per_epoch 8 --device_aug --savepath /tmp/mmd 2>&1 | grep -E "Iter 8/8" | cut -c25-170
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_mmformer_gpu.py tests/test_augment_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
