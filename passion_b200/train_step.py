"""Per-step loss mix, per-epoch relative-preference update and LR schedule of the reference training
loop (train.py:228-280, :299-308, :325-335; utils/lr_scheduler.py:15-17), on device tensors and
without host synchronisation (the reference loops over samples in Python with implicit syncs).
"""
import numpy as np
import torch

from . import criterions


def loss_mix(outputs, target, mask, imb_beta, modal_weight, *, mask_type="idt", warmup=False, num_cls=4,
             rp_allreduce=None):
    """outputs = Model.forward tuple.  Returns (loss, parts).  `rp_allreduce`, if given, is applied to the
    4-float rp_iter before thresholding (rp_mask is a GLOBAL-batch statistic, train.py:265-268)."""
    fuse_pred, prm_bs, sep_bs, kl_bs, proto_bs, dist_bs = outputs
    ce_bs, dice_bs = criterions.ce_dice_bs(fuse_pred, target, num_cls=num_cls)                 # one pass for both terms
    fuse_loss = (ce_bs + dice_bs).sum()          # :228-229
    prm_loss = prm_bs.sum()
    if mask_type == "pdt":
        sep_m, kl_m, proto_m, dist_m = sep_bs.sum(0), kl_bs.sum(0), proto_bs.sum(0), dist_bs.sum(0)
        rp_iter = (dist_bs / dist_bs.mean(1, keepdim=True) - 1).sum(0)                        # :239-241
        mw = torch.ones_like(imb_beta)
    else:
        fm = mask.to(torch.float32)
        sep_m, kl_m = (sep_bs * fm).sum(0), (kl_bs * fm).sum(0)                               # :260-263
        proto_m, dist_m = (proto_bs * fm).sum(0), (dist_bs * fm).sum(0)
        avg = dist_bs.sum(1, keepdim=True) / fm.sum(1, keepdim=True)
        # 0/0 -> NaN for a sample whose single present modality equals the fused path; NaN > 0 is
        # False, so such a sample switches rp_mask off for the whole batch — reference behaviour (:265-268)
        rp_iter = (fm * (dist_bs / avg - 1)).sum(0)
        mw = modal_weight
    if rp_allreduce is not None:
        rp_iter = rp_allreduce(rp_iter)
    rp_mask = (rp_iter > 0).to(torch.float32)
    kl_loss = (imb_beta * mw * kl_m).sum()
    proto_loss = (rp_mask * mw * proto_m).sum()
    if warmup:                                                                                # :275-277
        sep_loss = (imb_beta * mw * sep_m).sum()
        loss = fuse_loss * 0.0 + sep_loss + prm_loss * 0.0 + kl_loss * 0.0 + proto_loss * 0.0
    else:
        sep_loss = (rp_mask * imb_beta * mw * sep_m).sum()
        loss = fuse_loss + sep_loss + prm_loss + kl_loss * 0.5 + proto_loss * 0.1            # :280
    parts = dict(fuse=fuse_loss, prm=prm_loss, sep=sep_loss, kl=kl_loss, proto=proto_loss, sep_m=sep_m,
                 kl_m=kl_m, proto_m=proto_m, dist_m=dist_m, rp_iter=rp_iter, rp_mask=rp_mask)
    return loss, parts


def loss_mix_baseline(outputs, target, mask, *, mask_type="idt", warmup=False, num_cls=4):
    """The non-PASSION loop's loss (train.py:410-437): fuse + sum_m sep_m + prm, no preference gating."""
    fuse_pred, prm_bs, sep_bs = outputs
    ce_bs, dice_bs = criterions.ce_dice_bs(fuse_pred, target, num_cls=num_cls)                 # one pass for both terms
    fuse_loss = (ce_bs + dice_bs).sum()
    prm_loss = prm_bs.sum()
    sep_m = sep_bs.sum(0) if mask_type == "pdt" else (sep_bs * mask.to(torch.float32)).sum(0)
    sep_loss = sep_m.sum()
    loss = fuse_loss * 0.0 + sep_loss + prm_loss * 0.0 if warmup else fuse_loss + sep_loss + prm_loss
    return loss, dict(fuse=fuse_loss, prm=prm_loss, sep=sep_loss, sep_m=sep_m)


def preference_update(imb_beta, epoch_dist_m, eta, epoch, eta_ext=1.5):
    """train.py:325-335 (non-warm-up branch), same order of operations, on CPU tensors."""
    avg = sum(epoch_dist_m) / 4.0
    rp_epoch = (avg - epoch_dist_m) / avg
    if epoch % 100 == 0:
        eta = eta * eta_ext
    beta = imb_beta.cpu() - eta * rp_epoch
    beta = torch.clamp(beta, min=0.1, max=4.0)
    beta = 2 * beta / (sum(beta ** 2) ** 0.5)
    return beta, eta, rp_epoch


def poly_lr(base_lr, epoch, num_epochs):
    """utils/lr_scheduler.py:15-17."""
    return round(base_lr * np.power(1 - np.float32(epoch) / np.float32(num_epochs), 0.9), 8)
