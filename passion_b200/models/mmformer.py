"""mmFormer-style backbone + in-forward PASSION loss assembly on the passion_b200 CUDA kernels.

Drop-in for the reference `models/mmformer.py`: same constructor, same attributes poked from the training script
(`is_training`, `use_passion`, `mask_type`; train.py:91-92,212), same `forward(x, mask, target=None, temp=1.0)` and
return tuples (mmformer.py:447, :586, :659), same state_dict names and shapes.

B200-first structure (same ideas as models/rfnet.py):
  * activations are channels-last [N,D,H,W,C] in `compute_dtype` (bf16 by default, fp32 = check mode);
  * the four modality encoders (mmformer.py:399-402) run as ONE grouped launch per layer (4 weight groups, batch 4B);
  * the five decoder_fuse passes (full mask + four single-modality masks, mmformer.py:443,489-530) and the five
    inter-modal transformer passes run as ONE pass at batch 5B; the four decoder_sep passes run at batch 4B;
  * the pre-norm block (blocks.py:300-316: InstanceNorm -> LeakyReLU -> Conv+bias) is channel_stats + the fused
    IN/LeakyReLU kernel + the conv kernels (tcgen05 for the 3x3x3 layers in bf16) with the bias in the epilogue;
    torch.cat feeding a pre-norm conv is never materialised (InstanceNorm is per channel, the conv reads two sources);
  * channels-last makes every token reshape of the reference a free view: [B,C,p,p,p] -> tokens is `view`, and the
    reference's raw `view(B,p,p,p,4*512)` of the inter-modal tokens (mmformer.py:440) is the cl tensor itself.
The token path (LayerNorm, attention, MLP: ~125-500 tokens of width 512) is plain library GEMMs through torch
(cuBLAS / SDPA); it is not the hot path and costs well under a millisecond per step.
The nn.Conv3d / nn.Linear objects below are parameter containers only; their torch forward is never called.
"""
import weakref

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import criterions as crit
from .. import ops
from . import rfnet as _rf

basic_dims = 8
transformer_basic_dims = 512
mlp_dim = 4096
num_heads = 8
depth = 1
num_modals = 4
patch_size = 5                     # tokens per axis = input size / 16 (80^3 crops -> 5)
MODALS = ("flair", "t1ce", "t1", "t2")
UP_SCALES = (2, 4, 8, 16)          # mmformer.py:366-370


# ------------------------------------------------------------------------------------ conv blocks
class general_conv3d_prenorm(nn.Module):
    """Parameter container for blocks.py:300-316 (InstanceNorm -> LeakyReLU(0.2) -> Conv3d with bias)."""

    def __init__(self, in_ch, out_ch, k_size=3, stride=1, padding=1, pad_type='zeros'):
        super().__init__()
        self.conv = nn.Conv3d(in_ch, out_ch, kernel_size=k_size, stride=stride, padding=padding,
                              padding_mode=pad_type, bias=True)
        self.k_size, self.stride, self.pad_type = k_size, stride, pad_type

    def run(self, x0, x1=None):
        """x0 / x1: a cl tensor, or a (tensor, stats) pair when its producer already accumulated the InstanceNorm sums.
        Returns (y, stats of y): the conv epilogue accumulates the sums the NEXT pre-norm block needs."""
        a0 = ops.prenorm(*_pair(x0))
        a1 = ops.prenorm(*_pair(x1)) if x1 is not None else None
        return ops.conv3d_ref(a0, [self.conv.weight], [self.conv.bias], a1, ksize=self.k_size, stride=self.stride,
                              pad_mode=self.pad_type, want_stats=True)


def _pair(a):
    return a if isinstance(a, tuple) else (a, None)


def _plain_conv1(conv, x):
    """nn.Conv3d 1x1x1 head with bias -> logits (mmformer.py:90, 135-139)."""
    y, _ = ops.conv3d_ref(x, [conv.weight], [conv.bias], ksize=1, pad_mode="zeros")
    return y


class Encoder(nn.Module):
    """mmformer.py:24-64: five levels, the very first conv has no norm in front of it."""

    def __init__(self):
        super().__init__()
        b = basic_dims
        self.e1_c1 = nn.Conv3d(1, b, kernel_size=3, stride=1, padding=1, padding_mode='reflect', bias=True)
        self.e1_c2 = general_conv3d_prenorm(b, b, pad_type='reflect')
        self.e1_c3 = general_conv3d_prenorm(b, b, pad_type='reflect')
        for lvl in (2, 3, 4, 5):
            c = b * 2 ** (lvl - 1)
            setattr(self, f"e{lvl}_c1", general_conv3d_prenorm(c // 2, c, stride=2, pad_type='reflect'))
            setattr(self, f"e{lvl}_c2", general_conv3d_prenorm(c, c, pad_type='reflect'))
            setattr(self, f"e{lvl}_c3", general_conv3d_prenorm(c, c, pad_type='reflect'))


def _run_encoders(encoders, x):
    """Four Encoders as one grouped pass.  x [4B,D,H,W,1] ordered modality-major."""

    def gconv(name, t, stride=1, norm=True):
        """t: tensor or (tensor, stats); returns (y, stats of y)."""
        convs = [getattr(e, name) for e in encoders]
        convs = [c.conv if norm else c for c in convs]
        a = ops.prenorm(*_pair(t)) if norm else _pair(t)[0]
        return ops.conv3d_ref(a, [c.weight for c in convs], [c.bias for c in convs], ksize=3, stride=stride,
                              pad_mode="reflect", want_stats=True)

    feats = []
    x = gconv("e1_c1", x, norm=False)                              # (tensor, stats)
    for lvl in (1, 2, 3, 4, 5):
        if lvl > 1:
            x = gconv(f"e{lvl}_c1", x, stride=2)
        t = x[0] + gconv(f"e{lvl}_c3", gconv(f"e{lvl}_c2", x))[0]   # residual sum: its statistics need their own pass,
        x = (t, ops.channel_stats(t))                              # shared by every consumer of this level's features
        feats.append(x)
    return feats


class Decoder_sep(nn.Module):
    """mmformer.py:66-114.  Returns LOGITS (cl); the caller applies the softmax."""

    def __init__(self, num_cls=4):
        super().__init__()
        b = basic_dims
        for lvl in (4, 3, 2, 1):
            c = b * 2 ** (lvl - 1)
            setattr(self, f"d{lvl}_c1", general_conv3d_prenorm(c * 2, c, pad_type='reflect'))
            setattr(self, f"d{lvl}_c2", general_conv3d_prenorm(c * 2, c, pad_type='reflect'))
            setattr(self, f"d{lvl}_out", general_conv3d_prenorm(c, c, k_size=1, padding=0, pad_type='reflect'))
        self.seg_layer = nn.Conv3d(b, num_cls, kernel_size=1, stride=1, padding=0, bias=True)

    def run(self, x1, x2, x3, x4, x5):
        de = _pair(x5)[0]
        for lvl, skip in ((4, x4), (3, x3), (2, x2), (1, x1)):
            de = getattr(self, f"d{lvl}_c1").run(ops.upsample(de))
            de = getattr(self, f"d{lvl}_out").run(getattr(self, f"d{lvl}_c2").run(de, skip))[0]  # cat((de, skip))
        return _plain_conv1(self.seg_layer, de)


class fusion_prenorm(nn.Module):
    """blocks.py:533-542 (zero padding: general_conv3d_prenorm's default)."""

    def __init__(self, in_channel=64, num_cls=4):
        super().__init__()
        self.fusion_layer = nn.Sequential(general_conv3d_prenorm(in_channel * num_cls, in_channel, k_size=1, padding=0),
                                          general_conv3d_prenorm(in_channel, in_channel, k_size=3, padding=1),
                                          general_conv3d_prenorm(in_channel, in_channel, k_size=1, padding=0))

    def run(self, x):
        for m in self.fusion_layer:
            x = m.run(x)
        return x                                                   # (tensor, stats)


class Decoder_fuse(nn.Module):
    """mmformer.py:116-189, batched over decoder passes."""

    def __init__(self, num_cls=4):
        super().__init__()
        b = basic_dims
        for lvl in (4, 3, 2, 1):
            c = b * 2 ** (lvl - 1)
            setattr(self, f"d{lvl}_c1", general_conv3d_prenorm(c * 2, c, pad_type='reflect'))
            setattr(self, f"d{lvl}_c2", general_conv3d_prenorm(c * 2, c, pad_type='reflect'))
            setattr(self, f"d{lvl}_out", general_conv3d_prenorm(c, c, k_size=1, padding=0, pad_type='reflect'))
        for lvl in (4, 3, 2, 1):
            setattr(self, f"seg_d{lvl}", nn.Conv3d(b * 2 ** lvl, num_cls, kernel_size=1, stride=1, padding=0, bias=True))
        self.seg_layer = nn.Conv3d(b, num_cls, kernel_size=1, stride=1, padding=0, bias=True)
        for lvl in (5, 4, 3, 2, 1):
            setattr(self, f"RFM{lvl}", fusion_prenorm(in_channel=b * 2 ** (lvl - 1), num_cls=num_cls))

    def run(self, x1, x2, x3, x4, x5):
        """x_l [N,...,4*C_l] masked encoder features (x5: the inter-modal transformer output).
        Returns logits, (pred1..4) deep-supervision logits at levels 2..5, (de_x1_f..de_x5_f), all cl."""
        f = self.RFM5.run(x5)[0]
        preds, feats = [_plain_conv1(self.seg_d4, f)], [f]
        for lvl, xl in ((4, x4), (3, x3), (2, x2), (1, x1)):
            de = getattr(self, f"d{lvl}_c1").run(ops.upsample(f))
            r = getattr(self, f"RFM{lvl}").run(xl)
            f = getattr(self, f"d{lvl}_out").run(getattr(self, f"d{lvl}_c2").run(r, de))[0]       # cat((RFM, de))
            feats.append(f)
            if lvl > 1:
                preds.append(_plain_conv1(getattr(self, f"seg_d{lvl - 1}"), f))
        logits = _plain_conv1(self.seg_layer, f)
        return logits, tuple(reversed(preds)), tuple(reversed(feats))


# ------------------------------------------------------------------------------------ token path (library GEMMs)
class SelfAttention(nn.Module):
    def __init__(self, dim, heads=8, qkv_bias=False, dropout_rate=0.0):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(dropout_rate)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(dropout_rate)


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class PreNormDrop(nn.Module):
    def __init__(self, dim, dropout_rate, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.fn = fn


class GELU(nn.Module):
    def forward(self, x):
        return F.gelu(x)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout_rate):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), GELU(), nn.Dropout(p=dropout_rate),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(p=dropout_rate))


class Transformer(nn.Module):
    """mmformer.py:282-313: per layer  x += pos;  x += drop(attn(LN(x)));  x += ffn(LN(x))."""

    def __init__(self, embedding_dim, depth, heads, mlp_dim, dropout_rate=0.1, n_levels=1, n_points=4):
        super().__init__()
        self.depth, self.dropout_rate = depth, dropout_rate
        self.cross_attention_list = nn.ModuleList(
            [Residual(PreNormDrop(embedding_dim, dropout_rate,
                                  SelfAttention(embedding_dim, heads=heads, dropout_rate=dropout_rate))) for _ in range(depth)])
        self.cross_ffn_list = nn.ModuleList(
            [Residual(PreNorm(embedding_dim, FeedForward(embedding_dim, mlp_dim, dropout_rate))) for _ in range(depth)])

    def run(self, x, pos):
        """x [N, T, dim] in the compute dtype, pos [1, T, dim] fp32 parameter."""
        dt = x.dtype
        p = self.dropout_rate if self.training else 0.0

        def lin(t, layer):
            # nn.Linear of qkv / proj / FFN (mmformer.py:201,214,270-276): tcgen05 GEMM in bf16 (ops.linear), library GEMM in the fp32 check mode
            return ops.linear(t, layer.weight, layer.bias)

        def ln(t, layer):
            # nn.LayerNorm of PreNorm / PreNormDrop (mmformer.py:233-250): csrc/attn.cu
            return ops.layer_norm(t, layer.weight, layer.bias, layer.eps)

        for j in range(self.depth):
            x = x + pos.to(dt)
            pn = self.cross_attention_list[j].fn
            sa = pn.fn
            N, T, C = x.shape
            qkv = lin(ln(x, pn.norm), sa.qkv).view(N, T, 3, sa.num_heads, C // sa.num_heads)
            # softmax(q k^T / sqrt(d)) v with attention dropout (mmformer.py:203-213): batched tcgen05 GEMMs + row softmax in bf16
            h = F.dropout(lin(ops.attention(qkv, p), sa.proj), p, self.training)
            x = x + F.dropout(h, p, self.training)
            pf = self.cross_ffn_list[j].fn
            net = pf.fn.net
            h = F.dropout(F.gelu(lin(ln(x, pf.norm), net[0])), p, self.training)
            x = x + F.dropout(lin(h, net[3]), p, self.training)
        return x


class MaskModal(nn.Module):
    """kept for attribute compatibility (mmformer.py:316-326); masking is a scale in this implementation."""

    def forward(self, x, mask):
        B, K = x.shape[:2]
        return (x * mask.to(x.dtype).view(B, K, *([1] * (x.dim() - 2)))).reshape(B, -1, *x.shape[3:])


# ------------------------------------------------------------------------------------ model
class Model(nn.Module):
    def __init__(self, num_cls=4):
        super().__init__()
        td, ps = transformer_basic_dims, patch_size
        self.flair_encoder = Encoder()
        self.t1ce_encoder = Encoder()
        self.t1_encoder = Encoder()
        self.t2_encoder = Encoder()
        for m in MODALS:                                                    # IntraFormer, mmformer.py:337-351
            setattr(self, f"{m}_encode_conv", nn.Conv3d(basic_dims * 16, td, kernel_size=1, stride=1, padding=0))
        for m in MODALS:
            setattr(self, f"{m}_pos", nn.Parameter(torch.zeros(1, ps ** 3, td)))
        for m in MODALS:
            setattr(self, f"{m}_transformer", Transformer(td, depth, num_heads, mlp_dim))
        self.multimodal_transformer = Transformer(td, depth, num_heads, mlp_dim, n_levels=num_modals)
        self.multimodal_decode_conv = nn.Conv3d(td * num_modals, basic_dims * 16 * num_modals, kernel_size=1, padding=0)
        self.decoder_fuse = Decoder_fuse(num_cls=num_cls)
        self.decoder_sep = Decoder_sep(num_cls=num_cls)
        self.masker = MaskModal()

        self.mask_type = 'idt'
        self.use_passion = False
        self.is_training = False
        self.num_cls = num_cls
        self.compute_dtype = torch.bfloat16          # torch.float32 = check mode
        self.last = {}

        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                torch.nn.init.kaiming_normal_(m.weight)                     # mmformer.py:377-379

    # ------------------------------------------------------------------ pieces
    def _encoders(self):
        return (self.flair_encoder, self.t1ce_encoder, self.t1_encoder, self.t2_encoder)

    def _intra(self, f5, fm):
        """f5 [4,B,T,128] masked level-5 features -> masked intra-modal tokens [B,4,T,512] (mmformer.py:417-433)."""
        out = []
        for i, m in enumerate(MODALS):
            conv = getattr(self, f"{m}_encode_conv")
            tok = F.linear(f5[i], conv.weight.flatten(1).to(f5.dtype), conv.bias.to(f5.dtype))
            out.append(getattr(self, f"{m}_transformer").run(tok, getattr(self, f"{m}_pos")))
        intra = torch.stack(out, 1)                                          # [B,4,T,512]
        return intra * fm.to(intra.dtype)[:, :, None, None]

    def _inter(self, intra, ms, p):
        """intra [B,4,T,512], pass masks ms [P,B,4] -> x5 [P*B,p,p,p,512] cl (mmformer.py:433-441)."""
        P, B = ms.shape[:2]
        T = intra.shape[2]
        tok = (intra[None] * ms.to(intra.dtype)[:, :, :, None, None]).reshape(P * B, 4 * T, -1)
        pos = torch.cat([getattr(self, f"{m}_pos") for m in MODALS], 1)
        out = self.multimodal_transformer.run(tok, pos)
        conv = self.multimodal_decode_conv
        # the reference views the [B, 4T, 512] tokens as [B,p,p,p,2048] without un-interleaving the modalities
        vol = out.reshape(P * B, p, p, p, -1)
        return F.linear(vol, conv.weight.flatten(1).to(vol.dtype), conv.bias.to(vol.dtype)).contiguous()

    # ------------------------------------------------------------------ forward
    def forward(self, x, mask, target=None, temp=1.0):
        if not x.is_cuda:
            raise RuntimeError("passion_b200.models.mmformer.Model runs on CUDA only (no CPU fallback)")
        ops.begin_step(x.device)
        if self.mask_type == 'pdt':
            raise NotImplementedError("the reference's 'pdt' branch of mmformer.Model.forward reads x5 before it is "
                                      "defined (mmformer.py:449-470); only 'idt' masking is defined for this backbone")
        B = x.shape[0]
        dt = self.compute_dtype
        dev = x.device
        fm = mask.to(torch.float32)
        e = fm.t().contiguous()                                               # [4(m),B]
        xin = x.to(torch.float32).permute(1, 0, 2, 3, 4) * e[:, :, None, None, None]     # mmformer.py:397-398
        xe = xin.reshape(4 * B, *x.shape[2:], 1).to(dt).contiguous()
        enc_s = _run_encoders(self._encoders(), xe)                           # 5 levels of ([4B,d,h,w,C], sums), modality-major
        enc = [f for f, _ in enc_s]
        escale = e.reshape(4 * B, 1, 1, 1, 1).to(dt)
        # masked per-modality features (:406-416); masks are 0/1, so their InstanceNorm sums are the masked sums
        feat = [(f * escale, st * e.reshape(4 * B, 1, 1).double()) for f, st in enc_s]
        p = enc[4].shape[1]
        intra = self._intra(feat[4][0].view(4, B, p ** 3, -1), fm)

        train_passion = self.is_training and self.use_passion
        eye = torch.eye(4, device=dev, dtype=torch.float32)
        if train_passion:
            single = eye[:, None, :].expand(4, B, 4) * fm[None]
            ms = torch.cat((fm[None], single), 0)                             # [5,B,4]
            # mmformer.py:522: the T2 pass masks the transformer branch with masks_mod2 (not mod3)
            ms5 = torch.cat((fm[None], single[:3], single[2:3]), 0)
        else:
            ms = ms5 = fm[None]
        P = ms.shape[0]
        # masks are 0/1, so masking the already-masked features again (mask & pass mask) is the product of the two
        ys = []
        for f, st in enc_s[:4]:
            C = f.shape[-1]
            st_p = st.view(4, B, C, 2).permute(1, 0, 2, 3)[None] * ms.double()[:, :, :, None, None]      # [P,B,4,C,2]
            ys.append((ops.masked_stack(f, ms), st_p.reshape(P * B, 4 * C, 2).contiguous()))
        sep_logits = None
        if self.is_training and _rf.SEP_STREAM:
            # decoder_sep only needs the masked encoder features: second stream, as in models/rfnet.py
            main = torch.cuda.current_stream(dev)
            side = _rf._side_stream(dev)
            side.wait_stream(main)
            _rf._lend_to_stream(feat, side)              # `feat` is referenced by nothing but decoder_sep's saved tensors
            with torch.cuda.stream(side):
                sep_logits = self.decoder_sep.run(*feat)
                sep_logits.record_stream(main)
        x5 = self._inter(intra, ms5, p)
        logits, preds, des = self.decoder_fuse.run(*ys, x5)
        D, H, W = logits.shape[1:4]
        fuse_logits = logits.view(P, B, D, H, W, -1)
        self.last = {"fuse_logits": fuse_logits, "prm_logits": preds, "de_f": des, "passes": P, "enc": enc, "x5": x5}
        if not self.is_training:
            return ops.softmax4(fuse_logits[0]).permute(0, 4, 1, 2, 3)         # [B,C,D,H,W]

        if sep_logits is None:
            sep_logits = self.decoder_sep.run(*feat)                          # [4B,D,H,W,C], modality-major
        else:
            torch.cuda.current_stream(dev).wait_stream(_rf._side_stream(dev))
        self.last["sep_logits"] = sep_logits
        labels, cnt, wgt = crit.label_stats(target)
        V = D * H * W
        # one fused pass over the fused-decoder logits (prediction, its CE / Dice sums, KL of the single-modality passes) and one
        # over the four decoder_sep predictions; mmformer.py:480-483 zeroes a missing modality's probabilities, but its loss is
        # multiplied by the same 0 below, so that product over the volume is skipped (see models/rfnet.py)
        ce_f, dice_f, kl, probs = crit.logit_losses(logits, labels, cnt, wgt, P, 0, temp, want_probs=True)
        fuse_prob = probs.permute(0, 4, 1, 2, 3)                              # [B,C,D,H,W]
        fuse_prob._pb_ce_dice = (ce_f, dice_f, weakref.ref(target))
        ce, dice, _, _ = crit.logit_losses(sep_logits, labels, cnt, wgt, 4, 1)
        sep_loss = (e * (ce + dice).view(4, B)).t()                           # [B,4]

        prm_loss = torch.zeros(B, device=dev)
        wl = 1.0
        for prm, s in zip(preds, UP_SCALES):                                  # mmformer.py:564-571, 628-650
            wl /= 2.0
            pr = prm.view(P, B, *prm.shape[1:])
            ce, dice = crit.cedice(crit.up_probs(ops.softmax4(pr[0]), s), labels, cnt, wgt)
            prm_loss = prm_loss + wl * (ce + dice)
            if train_passion:
                ps_l = crit.up_probs(ops.softmax4(pr[1:].reshape(4 * B, *prm.shape[1:]), temp), s)
                pt_l = crit.up_probs(ops.softmax4(pr[0].detach(), temp), s)
                kl = kl + wl * crit.kl(ps_l, pt_l, temp)
        if not self.use_passion:
            return fuse_prob, prm_loss[:, None], sep_loss                     # mmformer.py:586

        kl = kl.view(4, B)
        de1 = des[0].view(P, B, V, -1)
        proto, dist = crit.proto(de1[1:].reshape(4 * B, V, -1), de1[0].detach(), labels.view(B, V), cnt)
        proto, dist = proto.view(4, B), dist.view(4, B)
        if not (self.training and self.multimodal_transformer.dropout_rate > 0):
            # Without dropout a sample whose ONLY present modality is m makes pass 1+m identical to pass 0, and the
            # reference gets kl = proto = dist = 0 exactly (see models/rfnet.py).  Not so for m = T2: its transformer
            # branch is masked with masks_mod2 (mmformer.py:522), so that pass differs from pass 0.
            ident = (mask.to(torch.bool) & (fm.sum(1, keepdim=True) == 1)).t().clone()
            ident[3] = False
            zero = torch.zeros((), device=dev)
            kl, proto, dist = (torch.where(ident, zero, t) for t in (kl, proto, dist))
        return fuse_prob, prm_loss[:, None], sep_loss, (e * kl).t(), (e * proto).t(), (e * dist).t()      # :659
