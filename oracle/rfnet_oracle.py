"""Oracle (TEST INFRASTRUCTURE): functional PyTorch-CPU fp32 restatement of the
RFNet backbone + in-forward PASSION loss assembly.

Follows (reference, /root/reference/code):
  models/blocks.py:354-370   general_conv3d  -> conv_in_lrelu()
  models/rfnet.py:15-48      Encoder         -> encoder()
  models/rfnet.py:50-89      Decoder_sep     -> decoder_sep()
  models/rfnet.py:91-152     Decoder_fuse    -> decoder_fuse()
  models/blocks.py:396-464   prm_generator_{laststage_,}pk -> prm_generator()
  models/blocks.py:495-531   modal_fusion / region_fusion
  models/blocks.py:582-626   region_aware_modal_fusion -> rfm()
  models/rfnet.py:154-174    MaskModal / MaskModal_NoCat -> mask_modal()
  models/rfnet.py:217-403    Model.forward   -> forward()

Parameters come in as a flat dict with the reference's state_dict names, so a
reference checkpoint drives this file unchanged.  Nothing here is shipped as
product code; the CUDA path in passion_b200/ never imports it.
"""
import torch
import torch.nn.functional as F

from . import criterions_oracle as crit

NUM_CLS = 4
MODALS = ("flair", "t1ce", "t1", "t2")


# ----------------------------------------------------------------------------- blocks
def conv_in_lrelu(P, name, x, k=3, stride=1, pad_mode="reflect"):
    """blocks.py:354-370: Conv3d(bias) -> InstanceNorm3d(affine=False, eps=1e-5) -> LeakyReLU(0.2)."""
    w, b = P[name + ".conv.weight"], P[name + ".conv.bias"]
    if k == 3:
        x = F.pad(x, (1,) * 6, mode=pad_mode if pad_mode != "zeros" else "constant")
    y = F.conv3d(x, w, b, stride=stride)
    y = F.instance_norm(y, eps=1e-5)
    return F.leaky_relu(y, 0.2)


def up2(x):
    return F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)


def mask_modal(x, mask):
    """rfnet.py:154-174 — zero the missing modalities of [B,K,C,D,H,W] (bool mask [B,K])."""
    return x * mask.to(x.dtype)[:, :, None, None, None, None]


def encoder(P, pre, x):
    """rfnet.py:36-48."""
    feats = []
    for lvl in (1, 2, 3, 4):
        x = conv_in_lrelu(P, f"{pre}.e{lvl}_c1", x, stride=1 if lvl == 1 else 2)
        x = x + conv_in_lrelu(P, f"{pre}.e{lvl}_c3", conv_in_lrelu(P, f"{pre}.e{lvl}_c2", x))
        feats.append(x)
    return feats


def decoder_sep(P, x1, x2, x3, x4, pre="decoder_sep"):
    """rfnet.py:72-89 — returns softmax probabilities."""
    de = conv_in_lrelu(P, f"{pre}.d3_c1", up2(x4))
    de = conv_in_lrelu(P, f"{pre}.d3_out", conv_in_lrelu(P, f"{pre}.d3_c2", torch.cat((de, x3), 1)), k=1)
    de = conv_in_lrelu(P, f"{pre}.d2_c1", up2(de))
    de = conv_in_lrelu(P, f"{pre}.d2_out", conv_in_lrelu(P, f"{pre}.d2_c2", torch.cat((de, x2), 1)), k=1)
    de = conv_in_lrelu(P, f"{pre}.d1_c1", up2(de))
    de = conv_in_lrelu(P, f"{pre}.d1_out", conv_in_lrelu(P, f"{pre}.d1_c2", torch.cat((de, x1), 1)), k=1)
    logits = F.conv3d(de, P[f"{pre}.seg_layer.weight"], P[f"{pre}.seg_layer.bias"])
    return F.softmax(logits, 1)


def prm_generator(P, pre, x, mask, upper=None):
    """blocks.py:409-416 (last stage, upper=None) / 457-464: PRM logits [B,4,...]."""
    B, K, C = x.shape[:3]
    y = mask_modal(x, mask).reshape(B, K * C, *x.shape[3:])
    e = conv_in_lrelu(P, f"{pre}.embedding_layer.0", y, k=1)
    e = conv_in_lrelu(P, f"{pre}.embedding_layer.1", e, k=3)
    e = conv_in_lrelu(P, f"{pre}.embedding_layer.2", e, k=1)
    if upper is not None:
        e = torch.cat((upper, e), 1)
    h = conv_in_lrelu(P, f"{pre}.prm_layer.0", e, k=1)
    return F.conv3d(h, P[f"{pre}.prm_layer.1.weight"], P[f"{pre}.prm_layer.1.bias"])


def rfm(P, pre, x, prm, mask):
    """blocks.py:597-626 with modal_fusion :504-517 and region_fusion :528-531.

    x [B,K,C,...] encoder features, prm [B,4,...] detached class probabilities.
    Written without the reference's [B,K,cls,C,...] broadcast product; algebra identical.
    """
    B, K, C = x.shape[:3]
    sp = x.shape[3:]
    y = mask_modal(x, mask)
    region = []
    for i in range(NUM_CLS):
        p = prm[:, i]                                        # [B,...]
        yp = y * p[:, None, None]                            # [B,K,C,...]  (blocks.py:602-609)
        prm_avg = p.mean(dim=(1, 2, 3)) + 1e-7               # [B]          (:507)
        feat_avg = yp.mean(dim=(3, 4, 5)) / prm_avg[:, None, None]   # [B,K,C] (:508)
        feat = torch.cat((feat_avg.reshape(B, K * C), prm_avg[:, None]), 1)   # (:510-511)
        w0, b0 = P[f"{pre}.modal_fusion.{i}.weight_layer.0.weight"], P[f"{pre}.modal_fusion.{i}.weight_layer.0.bias"]
        w2, b2 = P[f"{pre}.modal_fusion.{i}.weight_layer.2.weight"], P[f"{pre}.modal_fusion.{i}.weight_layer.2.bias"]
        h = F.leaky_relu(feat @ w0.reshape(w0.shape[0], -1).t() + b0, 0.2)
        gate = torch.sigmoid(h @ w2.reshape(w2.shape[0], -1).t() + b2)       # [B,K]  (:512-513)
        region.append((yp * gate[:, :, None, None, None, None]).sum(1))      # [B,C,...] (:516)
    r = torch.stack(region, 1).reshape(B, NUM_CLS * C, *sp)
    r = conv_in_lrelu(P, f"{pre}.region_fusion.fusion_layer.0", r, k=1)
    r = conv_in_lrelu(P, f"{pre}.region_fusion.fusion_layer.1", r, k=3)
    r = conv_in_lrelu(P, f"{pre}.region_fusion.fusion_layer.2", r, k=1)
    s = y.reshape(B, K * C, *sp)
    s = conv_in_lrelu(P, f"{pre}.short_cut.0", s, k=1)
    s = conv_in_lrelu(P, f"{pre}.short_cut.1", s, k=3)
    s = conv_in_lrelu(P, f"{pre}.short_cut.2", s, k=1)
    return torch.cat((r, s), 1)                              # (:625)


def decoder_fuse(P, x1, x2, x3, x4, mask, pre="decoder_fuse"):
    """rfnet.py:126-152 — returns logits, (prm1..4 logits), (de_x1..4)."""
    prm4 = prm_generator(P, f"{pre}.prm_generator4", x4, mask)
    de4 = rfm(P, f"{pre}.RFM4", x4, F.softmax(prm4, 1).detach(), mask)
    de4 = conv_in_lrelu(P, f"{pre}.d3_c1", up2(de4))

    prm3 = prm_generator(P, f"{pre}.prm_generator3", x3, mask, upper=de4)
    de3 = rfm(P, f"{pre}.RFM3", x3, F.softmax(prm3, 1).detach(), mask)
    de3 = conv_in_lrelu(P, f"{pre}.d3_out", conv_in_lrelu(P, f"{pre}.d3_c2", torch.cat((de3, de4), 1)), k=1)
    de3 = conv_in_lrelu(P, f"{pre}.d2_c1", up2(de3))

    prm2 = prm_generator(P, f"{pre}.prm_generator2", x2, mask, upper=de3)
    de2 = rfm(P, f"{pre}.RFM2", x2, F.softmax(prm2, 1).detach(), mask)
    de2 = conv_in_lrelu(P, f"{pre}.d2_out", conv_in_lrelu(P, f"{pre}.d2_c2", torch.cat((de2, de3), 1)), k=1)
    de2 = conv_in_lrelu(P, f"{pre}.d1_c1", up2(de2))

    prm1 = prm_generator(P, f"{pre}.prm_generator1", x1, mask, upper=de2)
    de1 = rfm(P, f"{pre}.RFM1", x1, F.softmax(prm1, 1).detach(), mask)
    de1 = conv_in_lrelu(P, f"{pre}.d1_out", conv_in_lrelu(P, f"{pre}.d1_c2", torch.cat((de1, de2), 1)), k=1)

    logits = F.conv3d(de1, P[f"{pre}.seg_layer.weight"], P[f"{pre}.seg_layer.bias"])
    return logits, (prm1, prm2, prm3, prm4), (de1, de2, de3, de4)


UP_SCALES = (1, 2, 4, 8)   # rfnet.py:207-211


def _up(scale):
    if scale == 1:
        return None
    return lambda t: F.interpolate(t, scale_factor=scale, mode="trilinear", align_corners=True)


# ----------------------------------------------------------------------------- model
def forward(P, x, mask, target=None, temp=1.0, *, is_training=True, use_passion=True,
            mask_type="idt", return_internals=False):
    """rfnet.py:217-403.  x [B,4,D,H,W] f32, mask [B,4] bool, target one-hot [B,4,D,H,W]."""
    B = x.shape[0]
    idt = mask_type != "pdt"
    if idt:
        x = x * mask.to(x.dtype)[:, :, None, None, None]                    # :232-233
    enc = [encoder(P, f"{m}_encoder", x[:, i:i + 1]) for i, m in enumerate(MODALS)]   # :234-237
    xs = [torch.stack([enc[m][l] for m in range(4)], 1) for l in range(4)]  # [B,4,C,...]
    if idt:
        xs = [mask_modal(t, mask) for t in xs]                              # :239-242
    fuse_pred, preds, de_f = decoder_fuse(P, *xs, mask)                     # :244
    if not is_training:
        return F.softmax(fuse_pred, 1)                                      # :403

    sep_preds = [decoder_sep(P, *enc[m]) for m in range(4)]                 # :254-257 (un-masked encoder feats)
    if idt:
        sep_preds = [p * mask[:, m].to(p.dtype)[:, None, None, None, None] for m, p in enumerate(sep_preds)]  # :259-260
    eye = torch.eye(4, dtype=torch.bool)
    masks_mod = [eye[m][None].repeat(B, 1) for m in range(4)]               # :262-265
    fmask = mask.to(torch.float32) if idt else torch.ones(B, 4)

    prm_loss = torch.zeros(B, 1)
    sep_loss = torch.zeros(B, 4)
    w = 1.0
    for prm_pred, s in zip(preds, UP_SCALES):                               # :284-288 / :384-387
        w /= 2.0
        pr = F.softmax(prm_pred, 1)
        prm_loss = prm_loss + w * crit.softmax_weighted_loss_bs(pr, target, NUM_CLS, up_op=_up(s)) \
                            + w * crit.dice_loss_bs(pr, target, NUM_CLS, up_op=_up(s))
    for m in range(4):
        e = masks_mod[m].float() * fmask
        sep_loss = sep_loss + e * (crit.softmax_weighted_loss_bs(sep_preds[m], target, NUM_CLS)
                                   + crit.dice_loss_bs(sep_preds[m], target, NUM_CLS))
    if not use_passion:
        return F.softmax(fuse_pred, 1), prm_loss, sep_loss                  # :402

    kl_loss = torch.zeros(B, 4)
    proto_loss = torch.zeros(B, 4)
    dist = torch.zeros(B, 4)
    internals = {"fuse_logits": fuse_pred, "prm_logits": preds, "de_f": de_f, "enc": enc, "mod": []}
    for m in range(4):                                                      # :269-275, :336-377
        fp_m, preds_m, de_m = decoder_fuse(P, *xs, masks_mod[m])
        internals["mod"].append((fp_m, preds_m, de_m))
        e = masks_mod[m].float() * fmask
        pl, dm = crit.prototype_passion_loss_bs(de_m[0], de_f[0].detach(), target, NUM_CLS)
        proto_loss = proto_loss + e * pl
        dist = dist + e * dm
        kl_loss = kl_loss + e * crit.temp_kl_loss_bs(fp_m, fuse_pred.detach(), temp)
        w = 1.0
        for pt, ps, s in zip(preds, preds_m, UP_SCALES):
            w /= 2.0
            kl_loss = kl_loss + e * w * crit.temp_kl_loss_bs(ps, pt.detach(), temp, up_op=_up(s))
    out = (F.softmax(fuse_pred, 1), prm_loss, sep_loss, kl_loss, proto_loss, dist)   # :379
    if return_internals:
        return out, internals
    return out
