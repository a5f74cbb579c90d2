"""PASSION loss library with the reference's signatures (utils/criterions.py) on CUDA tensors.

Public, reference-compatible entry points (all return [B,1] float32, as the reference):
    dice_loss_bs(output, target, num_cls, eps, up_op)              criterions.py:25-38
    softmax_weighted_loss_bs(output, target, num_cls, up_op)       criterions.py:59-76
    temp_kl_loss_bs(logit_s, logit_t, target, num_cls, temp, up_op) criterions.py:92-103
    prototype_passion_loss_bs(feature_s, feature_t, target, logit_s, logit_t, num_cls, temp, up_op)  :144-180
`output`/`logit_*`/`feature_*` are [B,C,D,H,W]; `target` is the one-hot [B,num_cls,D,H,W] tensor.
`up_op` is either None or an integer scale factor / nn.Upsample-like object with .scale_factor.

The model does not go through these wrappers: it calls the channels-last (`*_cl`) kernels below
directly on its internal [N,D,H,W,C] tensors, batched over the five decoder passes.
"""
import torch

from . import ops

CLAMP_MIN = 0.005        # criterions.py:69, 98-99
__all__ = ["dice_loss_bs", "softmax_weighted_loss_bs", "temp_kl_loss_bs", "prototype_passion_loss_bs"]


def _scale_of(up_op):
    if up_op is None:
        return 1
    if isinstance(up_op, int):
        return up_op
    sf = getattr(up_op, "scale_factor", None)
    if sf is None:          # nn.Identity (rfnet.py:207)
        return 1
    return int(sf)


def _to_cl(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


def up_probs(p, scale):
    """trilinear align_corners up-sampling of a cl fp32 tensor by an integer factor (own kernel)."""
    return p if scale == 1 else ops.upsample(p, scale)


# ------------------------------------------------------------------ channels-last internals
def target_stats(target_cl):
    """per-sample class voxel counts [B,C] and CE class weights 1 - count/total (criterions.py:67)."""
    cnt = target_cl.sum((1, 2, 3))
    return cnt, 1.0 - cnt / cnt.sum(1, keepdim=True)


def cedice_cl(prob, target_cl, cnt, wgt, eps=1e-7):
    """prob [P,B,D,H,W,C] (any leading pass dim) fp32 at label resolution; target_cl [B,D,H,W,C].
    Returns (ce [P,B], dice [P,B]) following criterions.py:25-38 and :59-76."""
    t = target_cl[None]
    dims = (2, 3, 4)
    num = (prob * t).sum(dims)                                     # [P,B,C]
    den = prob.sum(dims) + cnt[None] + eps
    dice = 1.0 - (2.0 * num / den).sum(-1) / prob.shape[-1]
    logp = torch.log(torch.clamp(prob, CLAMP_MIN, 1.0))
    voxels = prob.shape[2] * prob.shape[3] * prob.shape[4]
    ce = -((logp * t).sum(dims) * wgt[None]).sum(-1) / voxels
    return ce, dice


def kl_cl(ps, pt, temp):
    """ps [P,B,D,H,W,C], pt [B,D,H,W,C] (already soft-maxed at temperature and up-sampled).  [P,B]."""
    ps = torch.clamp(ps, CLAMP_MIN, 1.0)
    pt = torch.clamp(pt, CLAMP_MIN, 1.0)[None]
    kl = temp * temp * pt * (torch.log(pt) - torch.log(ps))
    return kl.mean((2, 3, 4, 5))


def _cos(f, proto, eps):
    """F.cosine_similarity(f, proto[..., None], dim=channel, eps) for cl tensors. f [..., V, C], proto [..., 1, C]."""
    fn = f.norm(dim=-1).clamp_min(eps)
    pn = proto.norm(dim=-1).clamp_min(eps)
    return (f * proto).sum(-1) / (fn * pn)


def proto_cl(fs, ft, target_cl, cnt, eps=1e-5):
    """fs [P,B,V,C] student features, ft [B,V,C] teacher features (detached), target_cl [B,V,K] one-hot.
    Returns proto [P,B], dist [P,B] (criterions.py:144-180; class used iff present in every local sample)."""
    present = (cnt > 0).all(0).to(fs.dtype)                        # [K]  (:157) — stays on device, no sync
    n_present = present.sum()
    den = cnt + eps                                                # [B,K]
    proto_s = torch.einsum("pbvc,bvk->pbkc", fs, target_cl) / den[None, :, :, None]
    proto_t = torch.einsum("bvc,bvk->bkc", ft, target_cl) / den[:, :, None]
    V = fs.shape[2]
    se = torch.zeros(fs.shape[:2], dtype=fs.dtype, device=fs.device)
    ab = torch.zeros_like(se)
    for k in range(target_cl.shape[-1]):
        s = _cos(fs, proto_s[:, :, k:k + 1, :], eps)              # [P,B,V]
        t = _cos(ft, proto_t[:, k:k + 1, :], eps)[None]
        d = s - t
        se = se + present[k] * (d * d).sum(-1)
        ab = ab + present[k] * d.abs().sum(-1)
    return se / (n_present * V), ab / (n_present * V)


# ------------------------------------------------------------------ reference-signature wrappers
def dice_loss_bs(output, target, num_cls=5, eps=1e-7, up_op=None):
    p = up_probs(_to_cl(output.float()), _scale_of(up_op))
    t = _to_cl(target.float())
    cnt, wgt = target_stats(t)
    return cedice_cl(p[None], t, cnt, wgt, eps)[1][0].unsqueeze(1)


def softmax_weighted_loss_bs(output, target, num_cls=5, up_op=None):
    p = up_probs(_to_cl(output.float()), _scale_of(up_op))
    t = _to_cl(target.float())
    cnt, wgt = target_stats(t)
    return cedice_cl(p[None], t, cnt, wgt)[0][0].unsqueeze(1)


def temp_kl_loss_bs(logit_s, logit_t, target=None, num_cls=5, temp=1.0, up_op=None):
    s = _scale_of(up_op)
    ps = up_probs(torch.softmax(_to_cl(logit_s.float()) / temp, -1), s)
    pt = up_probs(torch.softmax(_to_cl(logit_t.float()) / temp, -1), s)
    return kl_cl(ps[None], pt, temp)[0].unsqueeze(1)


def prototype_passion_loss_bs(feature_s, feature_t, target, logit_s=None, logit_t=None, num_cls=5, temp=1.0, up_op=None):
    fs = _to_cl(feature_s.float())
    ft = _to_cl(feature_t.float())
    t = _to_cl(target.float())
    B, C = fs.shape[0], fs.shape[-1]
    cnt, _ = target_stats(t)
    proto, dist = proto_cl(fs.reshape(1, B, -1, C), ft.reshape(B, -1, C), t.reshape(B, -1, t.shape[-1]), cnt)
    return proto[0].unsqueeze(1), dist[0].unsqueeze(1)
