"""RFNet backbone + in-forward PASSION loss assembly on the passion_b200 CUDA kernels.

Drop-in for the reference `models/rfnet.py`: same constructor, same attributes poked from the
training script (`is_training`, `use_passion`, `mask_type`; train.py:91-92,212), same
`forward(x, mask, target=None, temp=1.0)` and return tuples (rfnet.py:379, :402, :403), same
state_dict names and shapes, so reference checkpoints load both ways.

What is different underneath (B200-first):
  * activations are channels-last [N,D,H,W,C] in `compute_dtype` (bf16 by default, fp32 = check mode);
  * the four modality encoders (rfnet.py:234-237) run as ONE grouped launch per layer (4 weight groups);
  * the five decoder_fuse passes (full mask + four single-modality masks, rfnet.py:244,269-275) share
    weights and only use per-sample statistics, so they run as ONE pass at batch 5B; likewise the four
    decoder_sep passes (rfnet.py:254-257) run at batch 4B;
  * torch.cat feeding a conv is never materialised (two-source conv kernel), conv biases that feed an
    InstanceNorm are dropped (they cancel exactly), missing-modality masking is a per-(sample, modality)
    scale;
  * the class-presence gate of the prototype loss (criterions.py:157) is evaluated on the device, so
    the whole step is free of host synchronisation and CUDA-graph capturable.
The nn.Conv3d / nn.Sequential objects below are parameter containers only (names, shapes, init);
their torch forward is never called.
"""
import os
import weakref

import numpy as np
import torch
import torch.nn as nn

from .. import criterions as crit
from .. import ops

basic_dims = 8
num_modals = 4
SEP_STREAM = os.environ.get("PB_SEP_STREAM", "1") != "0"     # run decoder_sep concurrently with decoder_fuse (measured: -0.9 ms/step)
SIDE_RECORD = os.environ.get("PB_SIDE_RECORD", "1") != "0"   # debugging switch for _lend_to_stream (keep on)
# single-modality passes without their 3/4-zero stacks, at the N finest levels (the coarse levels are launch-bound: splitting
# their batch into dense + single parts costs more launches than the bytes it saves)
SPARSE_SINGLES = int(os.environ.get("PB_SPARSE_SINGLES", "2"))
_side = {}


def _side_stream(device):
    st = _side.get(device)
    if st is None:
        st = _side[device] = torch.cuda.Stream(device=device)
        ops.register_side_stream(device, st)
    return st


_consts = {}


def _level_weights(device):
    """[1, 1/2, 1/4, 1/8, 1/16]: weight of the full-resolution term and of the four PRM levels (rfnet.py:285-288)."""
    t = _consts.get(device)
    if t is None:
        t = _consts[device] = torch.tensor([1.0, 0.5, 0.25, 0.125, 0.0625], device=device)
    return t


def _lend_to_stream(tensors, stream):
    """Tensors allocated on the current stream are about to be read by kernels on `stream`, forward AND (as tensors saved
    for backward) by the backward of those ops, which autograd replays on `stream` too.  The caching allocator only knows
    the allocating stream: without this, a block whose last reference dies while a side-stream kernel is still reading it
    goes straight back to the main stream's pool and the next main-stream allocation overwrites it (the round-1 mmFormer
    t1_encoder gradient error: the masked level features were only referenced by decoder_sep's saved tensors)."""
    if not SIDE_RECORD:
        return
    for t in tensors:
        if isinstance(t, (tuple, list)):
            _lend_to_stream(t, stream)
        elif isinstance(t, torch.Tensor) and t.is_cuda:
            t.record_stream(stream)


class general_conv3d(nn.Module):
    """Parameter container for reference blocks.py:354-370 (conv -> InstanceNorm -> LeakyReLU)."""

    def __init__(self, in_ch, out_ch, k_size=3, stride=1, padding=1, pad_type='reflect'):
        super().__init__()
        self.conv = nn.Conv3d(in_ch, out_ch, kernel_size=k_size, stride=stride, padding=padding,
                              padding_mode=pad_type, bias=True)
        self.k_size, self.stride, self.pad_type = k_size, stride, pad_type

    def run(self, x0, x1=None, res=None):
        y = ops.conv_in_lrelu_ref(x0, [self.conv.weight], x1=x1, ksize=self.k_size, stride=self.stride,
                                  pad_mode=self.pad_type, res=res)
        # the bias cancels in InstanceNorm: give it an exact-zero gradient (keeps optimizer state shape)
        return _ZeroGradTouch.apply(y, self.conv.bias)

    def run_stack(self, y, enc=None):
        """A conv over the 4-modality stack: `y` [Nd,...,4C] holds the dense passes; `enc` [4B,...,C] (modality-major encoder
        output), if given, stands for four further single-modality passes (pass m: modality m in slot m, zeros elsewhere).
        For those only the m-th cin-quarter of the weight matters, so they run as ONE grouped conv on `enc` itself — a quarter
        of the bytes and FLOPs, and their 3/4-zero stacks are never built.  Returns the passes concatenated along the batch."""
        out = self.run(y)
        if enc is None:
            return out
        single = ops.conv_in_lrelu_ref(enc, [self.conv.weight], ksize=self.k_size, stride=self.stride, pad_mode=self.pad_type,
                                       slices=4)
        return torch.cat((out, single), 0)


class _ZeroGradTouch(torch.autograd.Function):
    """Identity on `y` that reports an all-zero gradient for `b` (a parameter with no influence on y)."""

    @staticmethod
    def forward(ctx, y, b):
        ctx.shape, ctx.meta = b.shape, (b.dtype, b.device)
        return y.view_as(y)

    @staticmethod
    def backward(ctx, dy):
        return dy, ops.zero_grad_like(ctx.shape, *ctx.meta)


def _plain_conv1(conv, x):
    """plain nn.Conv3d 1x1x1 head with bias (rfnet.py:69,107; blocks.py:407,455) -> logits."""
    y, _ = ops.conv3d_ref(x, [conv.weight], [conv.bias], ksize=1, pad_mode="zeros")
    return y


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        b = basic_dims
        cin = 1
        for lvl in (1, 2, 3, 4):
            c = b * 2 ** (lvl - 1)
            setattr(self, f"e{lvl}_c1", general_conv3d(cin, c, stride=1 if lvl == 1 else 2))
            setattr(self, f"e{lvl}_c2", general_conv3d(c, c))
            setattr(self, f"e{lvl}_c3", general_conv3d(c, c))
            cin = c


def _run_encoders(encoders, x):
    """Four Encoders (rfnet.py:36-48) as one grouped pass.  x [4B,D,H,W,1] ordered modality-major."""
    feats = []
    for lvl in (1, 2, 3, 4):
        def gw(name):
            return [getattr(e, name).conv.weight for e in encoders]
        stride = 1 if lvl == 1 else 2
        x = ops.conv_in_lrelu_ref(x, gw(f"e{lvl}_c1"), stride=stride)
        t = ops.conv_in_lrelu_ref(x, gw(f"e{lvl}_c2"))
        x = ops.conv_in_lrelu_ref(t, gw(f"e{lvl}_c3"), res=x)             # x + c3(c2(x))
        for e in encoders:                                                 # zero grads for the cancelled biases
            for nm in ("c1", "c2", "c3"):
                x = _ZeroGradTouch.apply(x, getattr(e, f"e{lvl}_{nm}").conv.bias)
        feats.append(x)
    return feats


class _DecoderConvs(nn.Module):
    def __init__(self, num_cls):
        super().__init__()
        b = basic_dims
        for lvl, c in ((3, b * 4), (2, b * 2), (1, b)):
            setattr(self, f"d{lvl}_c1", general_conv3d(c * 2, c))
            setattr(self, f"d{lvl}_c2", general_conv3d(c * 2, c))
            setattr(self, f"d{lvl}_out", general_conv3d(c, c, k_size=1, padding=0))
        self.seg_layer = nn.Conv3d(b, num_cls, kernel_size=1, stride=1, padding=0, bias=True)


class Decoder_sep(_DecoderConvs):
    """rfnet.py:50-89.  Returns LOGITS (cl); the caller applies the softmax."""

    def run(self, x1, x2, x3, x4):
        de = self.d3_c1.run(ops.upsample(x4))
        de = self.d3_out.run(self.d3_c2.run(de, x3))
        de = self.d2_c1.run(ops.upsample(de))
        de = self.d2_out.run(self.d2_c2.run(de, x2))
        de = self.d1_c1.run(ops.upsample(de))
        de = self.d1_out.run(self.d1_c2.run(de, x1))
        return _plain_conv1(self.seg_layer, de)


class modal_fusion(nn.Module):
    """Parameter container for blocks.py:495-503."""

    def __init__(self, in_channel):
        super().__init__()
        self.weight_layer = nn.Sequential(nn.Conv3d(4 * in_channel + 1, 128, 1, padding=0, bias=True),
                                          nn.LeakyReLU(negative_slope=0.2, inplace=True),
                                          nn.Conv3d(128, 4, 1, padding=0, bias=True))


class region_fusion(nn.Module):
    def __init__(self, in_channel, num_cls):
        super().__init__()
        self.fusion_layer = nn.Sequential(general_conv3d(in_channel * num_cls, in_channel, k_size=1, padding=0),
                                          general_conv3d(in_channel, in_channel, k_size=3, padding=1),
                                          general_conv3d(in_channel, in_channel // 2, k_size=1, padding=0))


def _run_seq(seq, x):
    for m in seq:
        x = m.run(x)
    return x


def _run_stack_seq(seq, y, enc):
    """first layer on the 4-modality stack (dense passes `y` + single-modality passes `enc`), the rest on the joint batch"""
    return _run_seq(list(seq)[1:], seq[0].run_stack(y, enc))


class region_aware_modal_fusion(nn.Module):
    """blocks.py:582-626."""

    def __init__(self, in_channel, num_cls=4):
        super().__init__()
        self.modal_fusion = nn.ModuleList([modal_fusion(in_channel) for _ in range(num_cls)])
        self.region_fusion = region_fusion(in_channel, num_cls)
        self.short_cut = nn.Sequential(general_conv3d(in_channel * 4, in_channel, k_size=1, padding=0),
                                       general_conv3d(in_channel, in_channel, k_size=3, padding=1),
                                       general_conv3d(in_channel, in_channel // 2, k_size=1, padding=0))

    def run(self, y, prm, enc=None):
        """y [Nd,D,H,W,4C] masked features of the dense passes (channel = modality*C + c); prm [N,D,H,W,4] fp32 detached
        probs of ALL passes; enc [4B,D,H,W,C]: four further single-modality passes (see general_conv3d.run_stack)."""
        mf = self.modal_fusion
        gp = ([m.weight_layer[0].weight for m in mf] + [m.weight_layer[0].bias for m in mf]
              + [m.weight_layer[2].weight for m in mf] + [m.weight_layer[2].bias for m in mf])      # the 16 gate-MLP parameters
        nd = y.shape[0]
        fl = self.region_fusion.fusion_layer
        r = fl[0].run(ops.rfm_region(y, prm[:nd], gp))
        if enc is not None:
            # the class-weighted mix of a single-modality pass reads C channels instead of 4C; its [4B,...,4C] region tensor
            # goes through the same 1x1x1 conv (weights shared with the dense passes) and only then joins the batch
            r = torch.cat((r, fl[0].run(ops.rfm_region_single(enc, prm[nd:].contiguous(), gp, enc.shape[0] // 4))), 0)
        r = _run_seq(list(fl)[1:], r)
        s = _run_stack_seq(self.short_cut, y, enc)
        return torch.cat((r, s), -1)


class prm_generator_pk(nn.Module):
    """blocks.py:396-416 (laststage: no upper feature) and :443-464."""

    def __init__(self, in_channel, num_cls=4, laststage=False):
        super().__init__()
        q = in_channel // 4
        self.embedding_layer = nn.Sequential(general_conv3d(in_channel * 4, q, k_size=1, padding=0),
                                             general_conv3d(q, q, k_size=3, padding=1),
                                             general_conv3d(q, in_channel, k_size=1, padding=0))
        self.prm_layer = nn.Sequential(general_conv3d(in_channel if laststage else in_channel * 2, 16, k_size=1, padding=0),
                                       nn.Conv3d(16, num_cls, kernel_size=1, padding=0, stride=1, bias=True))

    def run(self, y, upper=None, enc=None):
        e = _run_stack_seq(self.embedding_layer, y, enc)
        h = self.prm_layer[0].run(e) if upper is None else self.prm_layer[0].run(upper, e)   # cat((x1, emb))
        return _plain_conv1(self.prm_layer[1], h)


class Decoder_fuse(_DecoderConvs):
    """rfnet.py:91-152, batched over decoder passes."""

    def __init__(self, num_cls=4):
        super().__init__(num_cls)
        b = basic_dims
        self.RFM4 = region_aware_modal_fusion(b * 8, num_cls)
        self.RFM3 = region_aware_modal_fusion(b * 4, num_cls)
        self.RFM2 = region_aware_modal_fusion(b * 2, num_cls)
        self.RFM1 = region_aware_modal_fusion(b * 1, num_cls)
        self.prm_generator4 = prm_generator_pk(b * 8, num_cls, laststage=True)
        self.prm_generator3 = prm_generator_pk(b * 4, num_cls)
        self.prm_generator2 = prm_generator_pk(b * 2, num_cls)
        self.prm_generator1 = prm_generator_pk(b * 1, num_cls)

    @staticmethod
    def _probs(logits):
        """detached class probabilities of the PRM logits (rfnet.py:128: `prm_pred.detach()` after the softmax of blocks.py:414)"""
        with torch.no_grad():
            return ops.softmax4(logits)

    def run(self, y1, y2, y3, y4, enc=None):
        """y_l [Nd,D_l,H_l,W_l,4*C_l] masked encoder features of the dense passes; enc = the four levels of the modality-major
        encoder output [4B,...,C_l] when four single-modality passes follow the dense ones in the batch (PASSION training:
        Nd = B, N = 5B).  Returns logits, (prm1..4), (de1..4) of all N passes, all cl."""
        e1, e2, e3, e4 = enc if enc is not None else (None,) * 4
        prm4 = self.prm_generator4.run(y4, None, e4)
        de4 = self.RFM4.run(y4, self._probs(prm4), e4)
        de4 = self.d3_c1.run(ops.upsample(de4))

        prm3 = self.prm_generator3.run(y3, de4, e3)
        de3 = self.RFM3.run(y3, self._probs(prm3), e3)
        de3 = self.d3_out.run(self.d3_c2.run(de3, de4))
        de3 = self.d2_c1.run(ops.upsample(de3))

        prm2 = self.prm_generator2.run(y2, de3, e2)
        de2 = self.RFM2.run(y2, self._probs(prm2), e2)
        de2 = self.d2_out.run(self.d2_c2.run(de2, de3))
        de2 = self.d1_c1.run(ops.upsample(de2))

        prm1 = self.prm_generator1.run(y1, de2, e1)
        de1 = self.RFM1.run(y1, self._probs(prm1), e1)
        de1 = self.d1_out.run(self.d1_c2.run(de1, de2))

        logits = _plain_conv1(self.seg_layer, de1)
        return logits, (prm1, prm2, prm3, prm4), (de1, de2, de3, de4)


class MaskModal(nn.Module):
    """kept for state/attribute compatibility (rfnet.py:154-163); masking is a scale in this implementation."""

    def forward(self, x, mask):
        B, K = x.shape[:2]
        return (x * mask.to(x.dtype).view(B, K, *([1] * (x.dim() - 2)))).reshape(B, -1, *x.shape[3:])


class MaskModal_NoCat(nn.Module):
    def forward(self, x, mask):
        B, K = x.shape[:2]
        return x * mask.to(x.dtype).view(B, K, *([1] * (x.dim() - 2)))


UP_SCALES = (1, 2, 4, 8)          # rfnet.py:207-211


class Model(nn.Module):
    def __init__(self, num_cls=4):
        super().__init__()
        self.flair_encoder = Encoder()
        self.t1ce_encoder = Encoder()
        self.t1_encoder = Encoder()
        self.t2_encoder = Encoder()
        self.decoder_fuse = Decoder_fuse(num_cls=num_cls)
        self.decoder_sep = Decoder_sep(num_cls=num_cls)
        self.masker = MaskModal()
        self.masker_nocat = MaskModal_NoCat()

        self.is_training = False
        self.use_passion = False
        self.mask_type = 'idt'
        self.num_cls = num_cls
        self.compute_dtype = torch.bfloat16          # torch.float32 = check mode
        self.last = {}                               # prediction tuples of the last forward (parity tests)

        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                torch.nn.init.kaiming_normal_(m.weight)          # rfnet.py:213-215

    # ------------------------------------------------------------------ feature extraction
    def _features(self, x, mask):
        """-> enc: 4 levels of [4B,d,h,w,C], modality-major."""
        B = x.shape[0]
        dt = self.compute_dtype
        idt = self.mask_type != 'pdt'
        fm = mask.to(torch.float32)
        xin = x.to(torch.float32).permute(1, 0, 2, 3, 4)                       # [4,B,D,H,W]
        if idt:
            xin = xin * fm.t()[:, :, None, None, None]                         # rfnet.py:232-233
        xe = xin.reshape(4 * B, *x.shape[2:], 1).to(dt).contiguous()
        encs = (self.flair_encoder, self.t1ce_encoder, self.t1_encoder, self.t2_encoder)
        return _run_encoders(encs, xe)

    @staticmethod
    def _masked(enc, ms):
        """enc: per-level [4B,d,h,w,C] (modality-major) x pass masks ms [P,B,4] -> per-level [P*B,d,h,w,4C]."""
        return [ops.masked_stack(f, ms) for f in enc]

    # ------------------------------------------------------------------ forward
    def forward(self, x, mask, target=None, temp=1.0):
        if not x.is_cuda:
            raise RuntimeError("passion_b200.models.rfnet.Model runs on CUDA only (no CPU fallback)")
        ops.begin_step(x.device)
        B = x.shape[0]
        idt = self.mask_type != 'pdt'
        enc = self._features(x, mask)
        fm = mask.to(torch.float32)
        train_passion = self.is_training and self.use_passion
        eye = torch.eye(4, device=x.device, dtype=torch.float32)
        if train_passion:
            single = eye[:, None, :].expand(4, B, 4)                           # masks_mod0..3 (rfnet.py:262-265)
            eff = fm if idt else torch.ones_like(fm)
            # features are already masked by `mask` in idt mode (rfnet.py:239-242); the single-modality
            # passes then see mask & masks_mod_m
            ms = torch.cat((eff[None], single * eff[None]), 0)                 # [5,B,4]
        else:
            ms = (fm if idt else torch.ones_like(fm))[None]
        # pdt: decoder_fuse still masks inside PRM/RFM with the pass mask (blocks.py:409-413,597-600)
        if not idt:
            ms = ms.clone()
            ms[0] = fm
            if train_passion:
                ms[1:] = single
        P = ms.shape[0]
        sep_logits = None
        if self.is_training and SEP_STREAM:
            # decoder_sep only depends on the encoder outputs: run it on a second stream so that its many small launches at
            # the coarse levels (grids of 20-80 CTAs) fill the SMs that decoder_fuse's leave idle; autograd replays the
            # backward of each op on the stream of its forward, so the overlap carries over to the backward pass
            main = torch.cuda.current_stream(x.device)
            side = _side_stream(x.device)
            side.wait_stream(main)
            _lend_to_stream(enc, side)
            with torch.cuda.stream(side):
                sep_logits = self.decoder_sep.run(*enc)
                sep_logits.record_stream(main)
        if train_passion and SPARSE_SINGLES > 0:
            # passes 1..4 see ONE modality each (ms[1+m] = e_m * mask): they read the encoder output itself (already masked in
            # idt mode, unmasked in pdt mode — exactly what ms[1:] selects); only the full-mask pass needs a 4-modality stack
            sparse = [lvl < SPARSE_SINGLES for lvl in range(4)]
            ys = [ops.masked_stack(f, ms[:1] if sp else ms) for f, sp in zip(enc, sparse)]
            logits, prms, des = self.decoder_fuse.run(*ys, enc=[f if sp else None for f, sp in zip(enc, sparse)])
        else:
            ys = self._masked(enc, ms)
            logits, prms, des = self.decoder_fuse.run(*ys)
        D, H, W = logits.shape[1:4]
        fuse_logits = logits.view(P, B, D, H, W, -1)
        self.last = {"fuse_logits": fuse_logits, "prm_logits": prms, "de_f": des, "passes": P, "enc": enc}
        if not self.is_training:
            return ops.softmax4(fuse_logits[0]).permute(0, 4, 1, 2, 3)                   # [B,C,D,H,W]

        if sep_logits is None:
            sep_logits = self.decoder_sep.run(*enc)                           # [4B,D,H,W,C], modality-major
        else:
            torch.cuda.current_stream(x.device).wait_stream(_side_stream(x.device))
        self.last["sep_logits"] = sep_logits
        e = (fm if idt else torch.ones_like(fm)).t()                          # [4(m),B]
        labels, cnt, wgt = crit.label_stats(target)
        V = D * H * W

        # ONE pass over the fused-decoder logits: softmax of the full-mask pass (the returned prediction), its CE / Dice sums (the
        # step's fuse loss, train.py:228-229 — handed to criterions.ce_dice_bs through the tensor) and, with PASSION, the KL of the
        # four single-modality passes against it; the same pass over the level-1 PRM logits; a four-pass supervised variant over
        # the decoder_sep predictions.  No probability tensor is materialised at the labels' resolution.
        # (rfnet.py:259-260 multiplies the probabilities of a MISSING modality by 0 before the loss; that sample's loss is then
        # multiplied by the same 0 below and again in train.py:260: e * f(p) == e * f(e * p) for e in {0, 1}.)
        s_fuse, kl_sums, probs = ops.logit_loss(logits, labels, P, 0, temp, want_probs=True)
        fuse_prob = probs.permute(0, 4, 1, 2, 3)                                         # [B,C,D,H,W]
        s_sep, _, _ = ops.logit_loss(sep_logits, labels, 4, 1)
        sums, kls = [s_fuse, s_sep], [kl_sums]                # A / L / E sums of 1 + 4 + 4 supervised predictions per sample
        for prm, s in zip(prms, UP_SCALES):                   # prm loss (rfnet.py:284-288) and the PRM part of the KL (:340-344 ...)
            if s == 1:
                s_l, kl_l, _ = ops.logit_loss(prm, labels, P, 0, temp)
            else:
                pr = prm.view(P, B, *prm.shape[1:])
                s_l = ops.cedice_sums(crit.up_probs(ops.softmax4(pr[0]), s), labels)
                kl_l = None
                if train_passion:
                    ps_l = crit.up_probs(ops.softmax4(pr[1:].reshape(4 * B, *prm.shape[1:]), temp), s)
                    pt_l = crit.up_probs(ops.softmax4(pr[0].detach(), temp), s)
                    kl_l = ops.kl_sums(ps_l, pt_l.detach())
            sums.append(s_l)
            kls.append(kl_l)
        # the small arithmetic once for all nine predictions instead of once per prediction
        ce, dice = crit.ce_dice_from_sums(torch.cat(sums), cnt, wgt, V, B)               # [9B]: fuse | sep x4 | prm x4
        cd = (ce + dice).view(9, B)
        fuse_prob._pb_ce_dice = (ce[:B], dice[:B], weakref.ref(target))
        sep_loss = (e * cd[1:5]).t()                                                     # [B,4]   (rfnet.py:336 ...)
        lvl_w = _level_weights(x.device)                      # 1, 1/2, 1/4, 1/8, 1/16 (created once per device: no copy in a capture)
        prm_loss = (lvl_w[1:, None] * cd[5:9]).sum(0)
        if train_passion:                                     # T^2 * mean over (voxels, classes), criterions.py:98-102
            kl = (lvl_w[:, None] * torch.stack(kls)).sum(0) * (temp * temp / (V * 4))    # [4B]
        if not self.use_passion:
            return fuse_prob, prm_loss[:, None], sep_loss                     # rfnet.py:402

        # ---- PASSION terms: single-modality passes 1..4 vs the detached full-mask pass 0
        kl = kl.view(4, B)
        de1 = des[0].view(P, B, V, -1)
        proto, dist = crit.proto(de1[1:].reshape(4 * B, V, -1), de1[0].detach(), labels.view(B, V), cnt)
        proto, dist = proto.view(4, B), dist.view(4, B)
        # A sample whose ONLY present modality is m makes pass 1+m identical to pass 0 (same inputs, same
        # weights): the reference then gets kl = proto = dist = 0 EXACTLY (bit-identical passes), which is what
        # turns rp_iter into 0/0 = NaN in train.py:265-268.  Reproduce the exact zeros structurally.
        ident = (mask.to(torch.bool) & (fm.sum(1, keepdim=True) == 1)).t()   # [4(m),B]
        zero = torch.zeros((), device=x.device)
        kl, proto, dist = (torch.where(ident, zero, t) for t in (kl, proto, dist))
        kl_loss = (e * kl).t()
        proto_loss = (e * proto).t()
        dist_out = (e * dist).t()
        return fuse_prob, prm_loss[:, None], sep_loss, kl_loss, proto_loss, dist_out     # rfnet.py:379
