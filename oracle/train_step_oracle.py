"""Oracle (TEST INFRASTRUCTURE): restatement of the per-step loss mix, the per-epoch
relative-preference update and the LR schedule of the reference training loop.

Follows /root/reference/code/train.py:
  :228-229   fused-prediction CE + Dice
  :233-253   'pdt' mix          :258-280   'idt' mix
  :299-308   per-epoch dist accumulation
  :325-335   imb_beta update
and /root/reference/code/utils/lr_scheduler.py:15-17 (poly schedule).
"""
import numpy as np
import torch

from . import criterions_oracle as crit


def loss_mix(outputs, target, mask, imb_beta, modal_weight, *, mask_type="idt",
             warmup=False, num_cls=4):
    """outputs = Model.forward tuple (softmax(fuse), prm[B,1], sep[B,4], kl[B,4], proto[B,4], dist[B,4]).
    Returns (loss, parts dict).  Mirrors train.py:228-280 term by term."""
    fuse_pred, prm_bs, sep_bs, kl_bs, proto_bs, dist_bs = outputs
    B = fuse_pred.shape[0]
    fuse_loss = (crit.softmax_weighted_loss_bs(fuse_pred, target, num_cls)
                 + crit.dice_loss_bs(fuse_pred, target, num_cls)).sum()
    prm_loss = prm_bs.sum()
    rp_iter = torch.zeros(4)
    if mask_type == "pdt":
        sep_m, kl_m, proto_m, dist_m = sep_bs.sum(0), kl_bs.sum(0), proto_bs.sum(0), dist_bs.sum(0)
        for b in range(B):
            rp_iter = rp_iter + (dist_bs[b] / dist_bs[b].mean() - 1)
        mw = torch.ones(4)
    else:
        fm = mask.float()
        sep_m, kl_m = (sep_bs * fm).sum(0), (kl_bs * fm).sum(0)
        proto_m, dist_m = (proto_bs * fm).sum(0), (dist_bs * fm).sum(0)
        for b in range(B):
            avg = dist_bs[b].sum() / fm[b].sum()
            rp_iter = rp_iter + fm[b] * (dist_bs[b] / avg - 1)
        mw = modal_weight
    rp_mask = (rp_iter > 0).float()
    kl_loss = (imb_beta * mw * kl_m).sum()
    proto_loss = (rp_mask * mw * proto_m).sum()
    if warmup:                                               # train.py:275-277
        sep_loss = (imb_beta * mw * sep_m).sum()
        loss = fuse_loss * 0.0 + sep_loss + prm_loss * 0.0 + kl_loss * 0.0 + proto_loss * 0.0
    else:
        sep_loss = (rp_mask * imb_beta * mw * sep_m).sum()
        loss = fuse_loss + sep_loss + prm_loss + kl_loss * 0.5 + proto_loss * 0.1
    parts = dict(fuse=fuse_loss, prm=prm_loss, sep=sep_loss, kl=kl_loss, proto=proto_loss,
                 sep_m=sep_m, kl_m=kl_m, proto_m=proto_m, dist_m=dist_m, rp_iter=rp_iter, rp_mask=rp_mask)
    return loss, parts


def loss_mix_baseline(outputs, target, mask, *, mask_type="idt", num_cls=4):
    """train.py:410-437 — the non-PASSION loss: fuse + sum_m sep_m + prm."""
    fuse_pred, prm_bs, sep_bs = outputs
    fuse_loss = (crit.softmax_weighted_loss_bs(fuse_pred, target, num_cls)
                 + crit.dice_loss_bs(fuse_pred, target, num_cls)).sum()
    sep_m = sep_bs.sum(0) if mask_type == "pdt" else (sep_bs * mask.float()).sum(0)
    return fuse_loss + sep_m.sum() + prm_bs.sum(), dict(fuse=fuse_loss, prm=prm_bs.sum(), sep=sep_m.sum(), sep_m=sep_m)


def preference_update(imb_beta, epoch_dist_m, eta, epoch, eta_ext=1.5):
    """train.py:325-335 (the non-warm-up branch).  Returns (new imb_beta, new eta, rp_epoch)."""
    avg = epoch_dist_m.sum() / 4.0
    rp_epoch = (avg - epoch_dist_m) / avg
    if epoch % 100 == 0:
        eta = eta * eta_ext
    beta = torch.clamp(imb_beta - eta * rp_epoch, min=0.1, max=4.0)
    beta = 2 * beta / (beta ** 2).sum() ** 0.5
    return beta, eta, rp_epoch


def poly_lr(base_lr, epoch, num_epochs):
    """lr_scheduler.py:15-17."""
    return round(base_lr * np.power(1 - np.float32(epoch) / np.float32(num_epochs), 0.9), 8)
