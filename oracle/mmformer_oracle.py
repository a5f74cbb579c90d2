"""Oracle (TEST INFRASTRUCTURE): functional PyTorch-CPU fp32 restatement of the mmFormer-style backbone with the
in-forward PASSION assembly (BASELINE.json configs[3]).

Follows (reference, /root/reference/code):
  models/blocks.py:300-316      general_conv3d_prenorm (IN -> LeakyReLU -> Conv, zero padding by default) -> prenorm_conv()
  models/blocks.py:533-542      fusion_prenorm                                                       -> fusion_prenorm()
  models/mmformer.py:24-64      Encoder (5 levels, first conv bare)                                   -> encoder()
  models/mmformer.py:66-114     Decoder_sep                                                           -> decoder_sep()
  models/mmformer.py:116-189    Decoder_fuse (deep-supervision heads seg_d1..4)                       -> decoder_fuse()
  models/mmformer.py:192-313    SelfAttention / FeedForward / Transformer (dropout disabled here)      -> transformer()
  models/mmformer.py:381-659    Model.forward incl. the T2-path quirk at :522 (x5 masked with masks_mod2)
Dropout (p = 0.1 in the reference's Transformer) is the only stochastic piece; parity runs disable it on both sides.
"""
import torch
import torch.nn.functional as F

from . import criterions_oracle as crit

NUM_CLS = 4
MODALS = ("flair", "t1ce", "t1", "t2")
TDIM = 512
HEADS = 8
UP_SCALES = (2, 4, 8, 16)          # mmformer.py:366-370


def prenorm_conv(P, name, x, k=3, stride=1, pad_mode="reflect"):
    """blocks.py:312-316: InstanceNorm -> LeakyReLU(0.2) -> Conv3d(bias)."""
    x = F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2)
    if k == 3:
        x = F.pad(x, (1,) * 6, mode="reflect" if pad_mode == "reflect" else "constant")
    return F.conv3d(x, P[name + ".conv.weight"], P[name + ".conv.bias"], stride=stride)


def up(x, s=2):
    return F.interpolate(x, scale_factor=s, mode="trilinear", align_corners=True)


def mask_flat(x, mask):
    """mmformer.py:316-326 MaskModal: [B,K,C,...] -> zero missing modalities -> [B,K*C,...]."""
    B, K = x.shape[:2]
    y = x * mask.to(x.dtype).view(B, K, *([1] * (x.dim() - 2)))
    return y.reshape(B, -1, *x.shape[3:])


def encoder(P, pre, x):
    x = F.conv3d(F.pad(x, (1,) * 6, mode="reflect"), P[f"{pre}.e1_c1.weight"], P[f"{pre}.e1_c1.bias"])
    feats = []
    for lvl in (1, 2, 3, 4, 5):
        if lvl > 1:
            x = prenorm_conv(P, f"{pre}.e{lvl}_c1", x, stride=2)
        x = x + prenorm_conv(P, f"{pre}.e{lvl}_c3", prenorm_conv(P, f"{pre}.e{lvl}_c2", x))
        feats.append(x)
    return feats


def decoder_sep(P, x1, x2, x3, x4, x5, pre="decoder_sep"):
    de = prenorm_conv(P, f"{pre}.d4_c1", up(x5))
    for lvl, skip in ((4, x4), (3, x3), (2, x2), (1, x1)):
        de = prenorm_conv(P, f"{pre}.d{lvl}_out", prenorm_conv(P, f"{pre}.d{lvl}_c2", torch.cat((de, skip), 1)), k=1)
        if lvl > 1:
            de = prenorm_conv(P, f"{pre}.d{lvl - 1}_c1", up(de))
    logits = F.conv3d(de, P[f"{pre}.seg_layer.weight"], P[f"{pre}.seg_layer.bias"])
    return F.softmax(logits, 1)


def fusion_prenorm(P, pre, x):
    x = prenorm_conv(P, f"{pre}.fusion_layer.0", x, k=1)
    x = prenorm_conv(P, f"{pre}.fusion_layer.1", x, k=3, pad_mode="zeros")
    return prenorm_conv(P, f"{pre}.fusion_layer.2", x, k=1)


def decoder_fuse(P, x1, x2, x3, x4, x5, pre="decoder_fuse"):
    f5 = fusion_prenorm(P, f"{pre}.RFM5", x5)
    preds, feats = [], [f5]
    preds.append(F.conv3d(f5, P[f"{pre}.seg_d4.weight"], P[f"{pre}.seg_d4.bias"]))
    de = prenorm_conv(P, f"{pre}.d4_c1", up(f5))
    f = None
    for lvl, xl in ((4, x4), (3, x3), (2, x2), (1, x1)):
        r = fusion_prenorm(P, f"{pre}.RFM{lvl}", xl)
        f = prenorm_conv(P, f"{pre}.d{lvl}_out", prenorm_conv(P, f"{pre}.d{lvl}_c2", torch.cat((r, de), 1)), k=1)
        feats.append(f)
        if lvl > 1:
            preds.append(F.conv3d(f, P[f"{pre}.seg_d{lvl - 1}.weight"], P[f"{pre}.seg_d{lvl - 1}.bias"]))
            de = prenorm_conv(P, f"{pre}.d{lvl - 1}_c1", up(f))
    logits = F.conv3d(f, P[f"{pre}.seg_layer.weight"], P[f"{pre}.seg_layer.bias"])
    # reference order: (pred1, pred2, pred3, pred4) = heads at levels 2,3,4,5 ; (de_x1_f .. de_x5_f)
    return logits, tuple(reversed(preds)), tuple(reversed(feats))


def transformer(P, pre, x, pos):
    """mmformer.py:282-313 with depth 1 and dropout off: x = x + pos; x = x + attn(LN(x)); x = x + ffn(LN(x))."""
    x = x + pos
    a = f"{pre}.cross_attention_list.0.fn"
    h = F.layer_norm(x, (TDIM,), P[f"{a}.norm.weight"], P[f"{a}.norm.bias"])
    B, N, C = h.shape
    qkv = F.linear(h, P[f"{a}.fn.qkv.weight"]).reshape(B, N, 3, HEADS, C // HEADS).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * (C // HEADS) ** -0.5).softmax(-1)
    h = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(h, P[f"{a}.fn.proj.weight"], P[f"{a}.fn.proj.bias"])
    f = f"{pre}.cross_ffn_list.0.fn"
    h = F.layer_norm(x, (TDIM,), P[f"{f}.norm.weight"], P[f"{f}.norm.bias"])
    h = F.linear(F.gelu(F.linear(h, P[f"{f}.fn.net.0.weight"], P[f"{f}.fn.net.0.bias"])), P[f"{f}.fn.net.3.weight"], P[f"{f}.fn.net.3.bias"])
    return x + h


def _tokens(t):
    """[B,C,p,p,p] -> [B, p^3, C]  (mmformer.py:417)."""
    return t.permute(0, 2, 3, 4, 1).reshape(t.shape[0], -1, t.shape[1])


def _inter(P, intra_masked, pos_all, p):
    """inter-modal transformer + decode conv on the four (masked) intra-modal maps (mmformer.py:433-441)."""
    B = intra_masked[0].shape[0]
    tok = torch.cat([_tokens(t) for t in intra_masked], 1)                       # [B, 4 p^3, 512], modality-major
    out = transformer(P, "multimodal_transformer", tok, pos_all)
    vol = out.reshape(B, p, p, p, TDIM * 4).permute(0, 4, 1, 2, 3)               # raw view, exactly as the reference (:440)
    return F.conv3d(vol, P["multimodal_decode_conv.weight"], P["multimodal_decode_conv.bias"])


def _up(scale):
    return lambda t: F.interpolate(t, scale_factor=scale, mode="trilinear", align_corners=True)


def forward(P, x, mask, target=None, temp=1.0, *, is_training=True, use_passion=True, mask_type="idt"):
    """mmformer.py:381-659, 'idt' masking (the 'pdt' branch of the reference reads an undefined x5 and is not restated)."""
    assert mask_type != "pdt"
    B = x.shape[0]
    fm = mask.to(x.dtype)
    x = x * fm[:, :, None, None, None]                                           # :397-398
    enc = [encoder(P, f"{m}_encoder", x[:, i:i + 1]) for i, m in enumerate(MODALS)]
    # masked per-modality features (the reference re-chunks the masked stack, :406-416)
    feat = [[enc[m][l] * fm[:, m].view(B, 1, 1, 1, 1) for m in range(4)] for l in range(5)]
    xs = [torch.cat(feat[l], 1) for l in range(4)]                               # x1..x4 [B,4C,...]
    p = feat[4][0].shape[-1]
    intra = []
    for i, m in enumerate(MODALS):                                               # IntraFormer :417-431
        tok = _tokens(F.conv3d(feat[4][i], P[f"{m}_encode_conv.weight"], P[f"{m}_encode_conv.bias"]))
        out = transformer(P, f"{m}_transformer", tok, P[f"{m}_pos"])
        intra.append(out.reshape(B, p, p, p, TDIM).permute(0, 4, 1, 2, 3))
    intra = [t * fm[:, i].view(B, 1, 1, 1, 1) for i, t in enumerate(intra)]       # :433
    pos_all = torch.cat([P[f"{m}_pos"] for m in MODALS], 1)
    x5 = _inter(P, intra, pos_all, p)
    fuse_pred, preds, de_f = decoder_fuse(P, *xs, x5)                            # :443
    if not is_training:
        return F.softmax(fuse_pred, 1)

    sep_preds = [decoder_sep(P, *[feat[l][m] for l in range(5)]) for m in range(4)]
    sep_preds = [q * fm[:, m].view(B, 1, 1, 1, 1) for m, q in enumerate(sep_preds)]
    prm_loss = torch.zeros(B, 1)
    sep_loss = torch.zeros(B, 4)
    w = 1.0
    for prm_pred, s in zip(preds, UP_SCALES):
        w /= 2.0
        pr = F.softmax(prm_pred, 1)
        prm_loss = prm_loss + w * (crit.softmax_weighted_loss_bs(pr, target, NUM_CLS, up_op=_up(s))
                                   + crit.dice_loss_bs(pr, target, NUM_CLS, up_op=_up(s)))
    eye = torch.eye(4)
    for m in range(4):
        e = eye[m][None] * fm
        sep_loss = sep_loss + e * (crit.softmax_weighted_loss_bs(sep_preds[m], target, NUM_CLS)
                                   + crit.dice_loss_bs(sep_preds[m], target, NUM_CLS))
    if not use_passion:
        return F.softmax(fuse_pred, 1), prm_loss, sep_loss

    kl_loss = torch.zeros(B, 4)
    proto_loss = torch.zeros(B, 4)
    dist = torch.zeros(B, 4)
    for m in range(4):
        mm = eye[m]
        xs_m = [torch.cat([feat[l][k] * mm[k] for k in range(4)], 1) for l in range(4)]
        m5 = eye[2] if m == 3 else mm                                            # reference quirk, mmformer.py:522
        x5_m = _inter(P, [intra[k] * m5[k] for k in range(4)], pos_all, p)
        fp_m, preds_m, de_m = decoder_fuse(P, *xs_m, x5_m)
        e = eye[m][None] * fm
        pl, dm = crit.prototype_passion_loss_bs(de_m[0], de_f[0].detach(), target, NUM_CLS)
        proto_loss = proto_loss + e * pl
        dist = dist + e * dm
        kl_loss = kl_loss + e * crit.temp_kl_loss_bs(fp_m, fuse_pred.detach(), temp)
        w = 1.0
        for pt, ps, s in zip(preds, preds_m, UP_SCALES):
            w /= 2.0
            kl_loss = kl_loss + e * w * crit.temp_kl_loss_bs(ps, pt.detach(), temp, up_op=_up(s))
    return F.softmax(fuse_pred, 1), prm_loss, sep_loss, kl_loss, proto_loss, dist
