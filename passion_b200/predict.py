"""Sliding-window inference (reference utils/predict.py:144-218) on the B200-native RFNet.

Two entry points:

* `predict_volume(model, x, mask, patch_size)` — the reference loop restated: overlapping windows (50 % overlap,
  predict.py:181-195), `pred += model(x_window, mask)` (:209-214), divide by the overlap count (:215), argmax (:218).
* `predict_all_masks(model, x, masks, patch_size)` — the 15-mask sweep of train.py:589-604 as one engine
  (SURVEY.md §8 f-2): each encoder only sees its own modality channel (rfnet.py:222-225, 234-237), so its output for a
  PRESENT modality is the same for every mask and is zeroed otherwise (rfnet.py:239-242).  The four encoders therefore
  run once per window (4 encoder passes instead of 60) and the fused decoder runs once on a batch of 15 masked
  copies; overlap-add, division and argmax stay on the device.  Same kernels and the same per-sample arithmetic as 15
  calls of `predict_volume`; the only difference is the summation order of the float64 statistics atomics, i.e. the
  probabilities agree to ~1e-5 in fp32 (tests/test_predict_gpu.py) and the label maps differ only at bf16 near-ties.
"""
import numpy as np
import torch

from . import ops

# datasets_nii.py:27-30 / train.py:42-45
MASKS_TEST = [[False, False, False, True], [False, True, False, False], [False, False, True, False], [True, False, False, False],
              [False, True, False, True], [False, True, True, False], [True, False, True, False], [False, False, True, True],
              [True, False, False, True], [True, True, False, False], [True, True, True, False], [True, False, True, True],
              [True, True, False, True], [False, True, True, True], [True, True, True, True]]


def window_origins(size, patch, overlap=0.5):
    """predict.py:181-195 for one axis."""
    step = int(patch * (1 - overlap))
    cnt = int(np.ceil((size - patch) / (patch * (1 - overlap))))
    return [i * step for i in range(cnt)] + [size - patch]


def _windows(shape, patch):
    H, W, Z = shape
    return [(h, w, z) for h in window_origins(H, patch) for w in window_origins(W, patch) for z in window_origins(Z, patch)]


def _overlap_count(shape, patch, device):
    weight = torch.zeros((1, 1) + tuple(shape), dtype=torch.float32, device=device)
    for h, w, z in _windows(shape, patch):
        weight[:, :, h:h + patch, w:w + patch, z:z + patch] += 1.0                     # predict.py:198-203
    return weight


@torch.no_grad()
def predict_volume(model, x, mask, patch_size=80):
    """x [B,4,H,W,Z] float32 (CUDA), mask [B,4] bool -> (labels int64 [B,H,W,Z], averaged probabilities [B,C,H,W,Z])."""
    was, was_mode = model.is_training, model.training
    model.is_training = False
    model.eval()                                    # reference utils/predict.py:154 (mmFormer's dropout must be off)
    B = x.shape[0]
    shape = tuple(x.shape[2:])
    weight = _overlap_count(shape, patch_size, x.device)
    pred = None
    for h, w, z in _windows(shape, patch_size):
        part = model(x[:, :, h:h + patch_size, w:w + patch_size, z:z + patch_size].contiguous(), mask)
        if pred is None:
            pred = torch.zeros((B, part.shape[1]) + shape, dtype=torch.float32, device=x.device)
        pred[:, :, h:h + patch_size, w:w + patch_size, z:z + patch_size] += part
    pred = pred / weight
    model.is_training = was
    model.train(was_mode)
    return torch.argmax(pred, dim=1), pred


@torch.no_grad()
def predict_all_masks(model, x, masks=None, patch_size=80):
    """x [1,4,H,W,Z] float32 (CUDA); masks: list of [4] bool patterns (default: the 15 test masks).
    Returns labels int64 [M,H,W,Z] and averaged probabilities [M,C,H,W,Z], M = len(masks)."""
    if x.shape[0] != 1:
        raise ValueError("predict_all_masks handles one volume at a time (the reference test loader uses batch 1)")
    if model.mask_type == 'pdt':
        raise ValueError("the shared-encoder sweep relies on the idt masking order (rfnet.py:232-242)")
    masks = MASKS_TEST if masks is None else masks
    was_mode = model.training
    model.eval()                                    # reference utils/predict.py:154
    mt = torch.tensor(masks, dtype=torch.bool, device=x.device)                        # [M,4]
    M = mt.shape[0]
    shape = tuple(x.shape[2:])
    weight = _overlap_count(shape, patch_size, x.device)
    pred = torch.zeros((M, model.num_cls) + shape, dtype=torch.float32, device=x.device)
    all_present = torch.ones((1, 4), dtype=torch.bool, device=x.device)
    ms = mt.to(torch.float32)[:, None, :]                                              # [M, B=1, 4] pass masks
    for h, w, z in _windows(shape, patch_size):
        xw = x[:, :, h:h + patch_size, w:w + patch_size, z:z + patch_size].contiguous()
        ops.begin_step(x.device)
        enc = model._features(xw, all_present)                                         # encoders once
        ys = model._masked(enc, ms)                                                    # 15 masked copies
        logits, _, _ = model.decoder_fuse.run(*ys)                                     # one batch-15 decoder pass
        prob = ops.softmax4(logits).permute(0, 4, 1, 2, 3)                             # [M,C,p,p,p]
        pred[:, :, h:h + patch_size, w:w + patch_size, z:z + patch_size] += prob
    pred = pred / weight
    model.train(was_mode)
    return torch.argmax(pred, dim=1), pred
