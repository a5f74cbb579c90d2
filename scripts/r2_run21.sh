#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "linear" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_mmformer_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -1
timeout 300 python train.py --use_passion --batch_size 2 --synthetic --num_epochs 1 --iters_per_epoch 6 --savepath /tmp/mm 2>&1 | grep -E "Iter 6/6" | cut -c25-170
