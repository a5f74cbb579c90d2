"""passion_b200 — B200-native (sm_100a) implementation of the PASSION training hot path.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/passion_b200.h)
  _lib.py          ctypes loader (raises if the library is missing — no fallback)
  ops.py           torch.autograd wrappers over the C ABI
  models/rfnet.py  drop-in for the reference models/rfnet.py (same names, forward contract)
  criterions.py    drop-in for the reference utils/criterions.py (*_bs signatures)
  train_step.py    the per-step loss mix / preference update of the reference train.py
  ddp.py           one-process-per-GPU data parallel (NCCL bucketed gradient all-reduce)
"""
__version__ = "0.1.0"
