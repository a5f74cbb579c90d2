"""Oracle (TEST INFRASTRUCTURE): PyTorch-CPU restatement of the four per-sample
("_bs") PASSION loss functions.

Follows /root/reference/code/utils/criterions.py:
  dice_loss_bs               :25-38
  softmax_weighted_loss_bs   :59-76
  temp_kl_loss_bs            :92-103
  prototype_passion_loss_bs  :144-180   (dead code :153,160,165-173 dropped)
All return [B,1] float32.  `target` is the one-hot [B,C,D,H,W] tensor (float64 in
the reference data loader, cast to float32 first thing — criterions.py:26,60,145).
"""
import torch
import torch.nn.functional as F

CLAMP_MIN = 0.005          # criterions.py:69,98-99


def dice_loss_bs(output, target, num_cls=4, eps=1e-7, up_op=None):
    target = target.float()
    if up_op:
        output = up_op(output)
    dims = (2, 3, 4)
    num = (output * target).sum(dims)                     # [B,C]
    den = output.sum(dims) + target.sum(dims) + eps
    dice = (2.0 * num / den).sum(1)                       # sum over classes, in class order
    return (1.0 - dice / num_cls).unsqueeze(1)


def softmax_weighted_loss_bs(output, target, num_cls=4, up_op=None):
    target = target.float()
    if up_op:
        output = up_op(output)
    tot = target.sum((1, 2, 3, 4))                        # [B]
    wgt = 1.0 - target.sum((2, 3, 4)) / tot[:, None]      # [B,C]   (:67)
    ce = -(wgt[:, :, None, None, None] * target * torch.log(torch.clamp(output, CLAMP_MIN, 1.0))).sum(1)
    return ce.mean((1, 2, 3)).unsqueeze(1)


def temp_kl_loss_bs(logit_s, logit_t, temp=1.0, up_op=None):
    ps = F.softmax(logit_s / temp, 1)
    pt = F.softmax(logit_t / temp, 1)
    if up_op:
        ps, pt = up_op(ps), up_op(pt)
    ps = torch.clamp(ps, CLAMP_MIN, 1.0)
    pt = torch.clamp(pt, CLAMP_MIN, 1.0)
    kl = temp * temp * pt * (torch.log(pt) - torch.log(ps))
    return kl.mean((1, 2, 3, 4)).unsqueeze(1)


def prototype_passion_loss_bs(feature_s, feature_t, target, num_cls=4, eps=1e-5):
    """Class i contributes only when EVERY sample of the (local) batch contains it (:157)."""
    target = target.float()
    sims_s, sims_t = [], []
    for i in range(num_cls):
        ti = target[:, i]                                  # [B,D,H,W]
        cnt = ti.sum((1, 2, 3))
        if bool((cnt > 0).all()):
            den = cnt[:, None] + eps
            proto_s = (feature_s * ti[:, None]).sum((2, 3, 4)) / den     # [B,C]
            proto_t = (feature_t * ti[:, None]).sum((2, 3, 4)) / den
            sims_s.append(F.cosine_similarity(feature_s, proto_s[:, :, None, None, None], dim=1, eps=eps))
            sims_t.append(F.cosine_similarity(feature_t, proto_t[:, :, None, None, None], dim=1, eps=eps))
    s = torch.stack(sims_s, 1)
    t = torch.stack(sims_t, 1)
    proto = ((s - t) ** 2).mean((1, 2, 3, 4)).unsqueeze(1)
    dist = torch.sqrt((s - t) ** 2).mean((1, 2, 3, 4)).unsqueeze(1)
    return proto, dist
