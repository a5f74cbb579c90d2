"""Evaluation scores of the inference sweep (reference utils/predict.py:82-128 `softmax_output_dice_class4`, the
`AverageMeter` of :131-143 and the per-mask loop of train.py:589-604), computed from exact integer confusion counts.

The reference thresholds the label maps into float masks and reduces each of them separately (18 full-volume reductions
per call, and per mask of the 15-mask sweep).  Here one `bincount` per label map gives the 4x4 confusion matrix
(prediction x target); every Dice variant is then the reference's own float32 expression on those counts, which is
bit-identical to the reference as long as the counts stay below 2^24 (a 240x240x155 volume has 8.9 M voxels).
Plain tensor arithmetic on whatever device the label maps live on (no custom kernel involved).
"""
import torch

EPS = 1e-8                                                   # predict.py:83
CLASS_EVALUATION = ('whole', 'core', 'enhancing', 'enhancing_postpro')      # predict.py:162
CLASS_SEPARATE = ('ncr_net', 'edema', 'enhancing')                          # predict.py:163
POSTPRO_MIN_VOXELS = 500                                     # predict.py:108


def confusion_counts(pred, target, num_cls=4):
    """pred, target: integer label maps [B, ...] (same shape) -> int64 [B, num_cls (pred), num_cls (target)]."""
    if pred.shape != target.shape:
        raise ValueError(f"label maps differ in shape: {tuple(pred.shape)} vs {tuple(target.shape)}")
    B = pred.shape[0]
    idx = pred.reshape(B, -1).long() * num_cls + target.reshape(B, -1).long()
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= num_cls * num_cls):
        raise ValueError("labels outside [0, num_cls)")
    idx = idx + (torch.arange(B, device=idx.device) * (num_cls * num_cls))[:, None]
    return torch.bincount(idx.reshape(-1), minlength=B * num_cls * num_cls).view(B, num_cls, num_cls)


def _dice(inter, o, t):
    """predict.py:88-90 on counts: (sum 2 o t + eps) / (sum o + sum t + eps), all float32."""
    inter, o, t = inter.to(torch.float32), o.to(torch.float32), t.to(torch.float32)
    return (2 * inter + EPS) / (o + t + EPS)


def dice_class4(pred, target):
    """softmax_output_dice_class4 (predict.py:82-128).  pred, target: label maps [B,H,W,Z] ->
    (dice_separate float32 [B,3] = ncr_net, edema, enhancing; dice_evaluate float32 [B,4] = whole, core, enhancing,
    enhancing_postpro).  Tensors stay on the input's device (the reference returns numpy arrays)."""
    cm = confusion_counts(pred, target, 4)                   # [B, pred, target]
    o = cm.sum(2)                                            # voxels predicted as class c
    t = cm.sum(1)                                            # voxels labelled class c
    d = [_dice(cm[:, c, c], o[:, c], t[:, c]) for c in (1, 2, 3)]
    # post-processing (:107-115): drop the enhancing prediction when fewer than 500 voxels IN THE WHOLE BATCH carry it
    keep = (o[:, 3].sum() >= POSTPRO_MIN_VOXELS).to(cm.dtype)
    d_post = _dice(cm[:, 3, 3] * keep, o[:, 3] * keep, t[:, 3])
    whole = _dice(cm[:, 1:, 1:].sum((1, 2)), o[:, 1:].sum(1), t[:, 1:].sum(1))
    core_i = cm[:, 1, 1] + cm[:, 1, 3] + cm[:, 3, 1] + cm[:, 3, 3]
    core = _dice(core_i, o[:, 1] + o[:, 3], t[:, 1] + t[:, 3])
    return torch.stack(d, 1), torch.stack((whole, core, d[2], d_post), 1)


class AverageMeter:
    """predict.py:131-143."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def evaluate_all_masks(model, x, target, patch_size=80, masks=None, mask_names=None):
    """One test case through the 15-mask sweep (predict.predict_all_masks) and the reference's Dice scores per mask.
    x [1,4,H,W,Z] float32 (CUDA), target [1,H,W,Z] integer labels -> {mask name: float32 [4] (whole, core, enhancing,
    enhancing_postpro)} in the reference's evaluation order (train.py:589-604 walks masks_test[::-1])."""
    from . import predict
    masks = predict.MASKS_TEST if masks is None else masks
    names = mask_names if mask_names is not None else [str(i) for i in range(len(masks))]
    if hasattr(model, "_features") and getattr(model, "mask_type", "idt") != "pdt":
        labels, _ = predict.predict_all_masks(model, x, masks, patch_size)             # shared-encoder sweep (RFNet)
    else:                                                                              # one sliding-window pass per mask
        labels = torch.cat([predict.predict_volume(model, x, torch.tensor([m], dtype=torch.bool, device=x.device), patch_size)[0]
                            for m in masks])
    tgt = target.expand(labels.shape[0], *target.shape[1:]).to(labels.device)
    # the reference scores one mask per call (batch 1), so its < 500-voxel post-processing rule applies per mask
    rows = [dice_class4(labels[i:i + 1], tgt[i:i + 1])[1][0] for i in range(labels.shape[0])]
    return {names[i]: rows[i] for i in reversed(range(len(masks)))}
