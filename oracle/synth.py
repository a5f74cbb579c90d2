"""Oracle (TEST INFRASTRUCTURE): deterministic synthetic weights / inputs.

numpy.RandomState (MT19937) is stable across numpy versions, so the same seed gives
the same tensors here, on the GPU box and in the committed golden fixtures — no
dependence on torch's RNG stream.  Distribution follows the reference initialisation
(rfnet.py:213-215: kaiming_normal_ on every Conv3d weight, i.e. N(0, 2/fan_in);
biases keep nn.Conv3d's default U(-1/sqrt(fan_in), 1/sqrt(fan_in))).

Synthetic batch follows SURVEY.md §8(d): x ~ N(0,1) [B,4,S,S,S]; labels 'U' = i.i.d.
uniform over 4 classes, 'S' = nested-sphere phantom; masks drawn from the mr2468 table.
"""
from collections import OrderedDict

import numpy as np
import torch

from .masks import MASK_ARRAY

BASIC = 8          # rfnet.py:11


def _gc(d, name, cin, cout, k):
    d[f"{name}.conv.weight"] = (cout, cin, k, k, k)
    d[f"{name}.conv.bias"] = (cout,)


def rfnet_param_shapes(num_cls=4):
    """state_dict names/shapes of rfnet.Model (rfnet.py:176-215), in registration order."""
    d = OrderedDict()
    b = BASIC
    for m in ("flair", "t1ce", "t1", "t2"):
        pre = f"{m}_encoder"
        cin = 1
        for lvl in (1, 2, 3, 4):
            c = b * 2 ** (lvl - 1)
            _gc(d, f"{pre}.e{lvl}_c1", cin, c, 3)
            _gc(d, f"{pre}.e{lvl}_c2", c, c, 3)
            _gc(d, f"{pre}.e{lvl}_c3", c, c, 3)
            cin = c

    def dec_convs(pre):
        for lvl, c in ((3, b * 4), (2, b * 2), (1, b)):
            _gc(d, f"{pre}.d{lvl}_c1", c * 2, c, 3)
            _gc(d, f"{pre}.d{lvl}_c2", c * 2, c, 3)
            _gc(d, f"{pre}.d{lvl}_out", c, c, 1)
        d[f"{pre}.seg_layer.weight"] = (num_cls, b, 1, 1, 1)
        d[f"{pre}.seg_layer.bias"] = (num_cls,)

    dec_convs("decoder_fuse")
    for lvl in (4, 3, 2, 1):
        c = b * 2 ** (lvl - 1)
        pre = f"decoder_fuse.RFM{lvl}"
        for i in range(num_cls):
            d[f"{pre}.modal_fusion.{i}.weight_layer.0.weight"] = (128, 4 * c + 1, 1, 1, 1)
            d[f"{pre}.modal_fusion.{i}.weight_layer.0.bias"] = (128,)
            d[f"{pre}.modal_fusion.{i}.weight_layer.2.weight"] = (4, 128, 1, 1, 1)
            d[f"{pre}.modal_fusion.{i}.weight_layer.2.bias"] = (4,)
        _gc(d, f"{pre}.region_fusion.fusion_layer.0", c * num_cls, c, 1)
        _gc(d, f"{pre}.region_fusion.fusion_layer.1", c, c, 3)
        _gc(d, f"{pre}.region_fusion.fusion_layer.2", c, c // 2, 1)
        _gc(d, f"{pre}.short_cut.0", c * 4, c, 1)
        _gc(d, f"{pre}.short_cut.1", c, c, 3)
        _gc(d, f"{pre}.short_cut.2", c, c // 2, 1)
    for lvl in (4, 3, 2, 1):
        c = b * 2 ** (lvl - 1)
        pre = f"decoder_fuse.prm_generator{lvl}"
        _gc(d, f"{pre}.embedding_layer.0", c * 4, c // 4, 1)
        _gc(d, f"{pre}.embedding_layer.1", c // 4, c // 4, 3)
        _gc(d, f"{pre}.embedding_layer.2", c // 4, c, 1)
        _gc(d, f"{pre}.prm_layer.0", c if lvl == 4 else 2 * c, 16, 1)
        d[f"{pre}.prm_layer.1.weight"] = (num_cls, 16, 1, 1, 1)
        d[f"{pre}.prm_layer.1.bias"] = (num_cls,)
    dec_convs("decoder_sep")
    return d


def mmformer_param_shapes(num_cls=4, patch=5, basic=8, tdim=512, mlp=4096):
    """state_dict names/shapes of mmformer.Model (mmformer.py:24-189, 329-379) in registration order
    (a module's own Parameters — the four *_pos — come before its sub-modules)."""
    d = OrderedDict()
    mods = ("flair", "t1ce", "t1", "t2")
    for m in mods:
        d[f"{m}_pos"] = (1, patch ** 3, tdim)
    for m in mods:
        pre = f"{m}_encoder"
        d[f"{pre}.e1_c1.weight"] = (basic, 1, 3, 3, 3)
        d[f"{pre}.e1_c1.bias"] = (basic,)
        _gc(d, f"{pre}.e1_c2", basic, basic, 3)
        _gc(d, f"{pre}.e1_c3", basic, basic, 3)
        for lvl in (2, 3, 4, 5):
            c = basic * 2 ** (lvl - 1)
            _gc(d, f"{pre}.e{lvl}_c1", c // 2, c, 3)
            _gc(d, f"{pre}.e{lvl}_c2", c, c, 3)
            _gc(d, f"{pre}.e{lvl}_c3", c, c, 3)
    for m in mods:
        d[f"{m}_encode_conv.weight"] = (tdim, basic * 16, 1, 1, 1)
        d[f"{m}_encode_conv.bias"] = (tdim,)

    def transformer(pre):
        a = f"{pre}.cross_attention_list.0.fn"
        d[f"{a}.norm.weight"] = (tdim,)
        d[f"{a}.norm.bias"] = (tdim,)
        d[f"{a}.fn.qkv.weight"] = (3 * tdim, tdim)
        d[f"{a}.fn.proj.weight"] = (tdim, tdim)
        d[f"{a}.fn.proj.bias"] = (tdim,)
        f = f"{pre}.cross_ffn_list.0.fn"
        d[f"{f}.norm.weight"] = (tdim,)
        d[f"{f}.norm.bias"] = (tdim,)
        d[f"{f}.fn.net.0.weight"] = (mlp, tdim)
        d[f"{f}.fn.net.0.bias"] = (mlp,)
        d[f"{f}.fn.net.3.weight"] = (tdim, mlp)
        d[f"{f}.fn.net.3.bias"] = (tdim,)

    for m in mods:
        transformer(f"{m}_transformer")
    transformer("multimodal_transformer")
    d["multimodal_decode_conv.weight"] = (basic * 16 * 4, tdim * 4, 1, 1, 1)
    d["multimodal_decode_conv.bias"] = (basic * 16 * 4,)

    def dec_convs(pre):
        for lvl in (4, 3, 2, 1):
            c = basic * 2 ** (lvl - 1)
            _gc(d, f"{pre}.d{lvl}_c1", c * 2, c, 3)
            _gc(d, f"{pre}.d{lvl}_c2", c * 2, c, 3)
            _gc(d, f"{pre}.d{lvl}_out", c, c, 1)

    dec_convs("decoder_fuse")
    for lvl in (4, 3, 2, 1):
        d[f"decoder_fuse.seg_d{lvl}.weight"] = (num_cls, basic * 2 ** lvl, 1, 1, 1)
        d[f"decoder_fuse.seg_d{lvl}.bias"] = (num_cls,)
    d["decoder_fuse.seg_layer.weight"] = (num_cls, basic, 1, 1, 1)
    d["decoder_fuse.seg_layer.bias"] = (num_cls,)
    for lvl in (5, 4, 3, 2, 1):
        c = basic * 2 ** (lvl - 1)
        pre = f"decoder_fuse.RFM{lvl}.fusion_layer"
        _gc(d, f"{pre}.0", c * num_cls, c, 1)
        _gc(d, f"{pre}.1", c, c, 3)
        _gc(d, f"{pre}.2", c, c, 1)
    dec_convs("decoder_sep")
    d["decoder_sep.seg_layer.weight"] = (num_cls, basic, 1, 1, 1)
    d["decoder_sep.seg_layer.bias"] = (num_cls,)
    return d


def make_state_dict(seed=1037, shapes=None, bias_scale=1.0):
    """Kaiming-normal weights / uniform biases from numpy RandomState(seed), fp32 torch tensors.
    `shapes` = OrderedDict name -> shape (default: the RFNet table).  Rules by name/rank: conv / linear weights
    N(0, 2/fan_in); their biases U(+-1/sqrt(fan_in)); 1-D "weight" (LayerNorm gain) 1 + 0.1 N; LayerNorm bias 0.1 N;
    "*_pos" position embeddings 0.02 N (zeros in the reference — randomised so that the tests exercise them)."""
    rs = np.random.RandomState(seed)
    shapes = shapes or rfnet_param_shapes()
    sd = OrderedDict()
    fan = {}
    for name, shp in shapes.items():
        shp = tuple(shp)
        if name.endswith("_pos"):
            sd[name] = torch.from_numpy((0.02 * rs.standard_normal(shp)).astype(np.float32))
        elif name.endswith("weight") and len(shp) == 1:
            fan[name[:-len("weight")]] = None
            sd[name] = torch.from_numpy((1.0 + 0.1 * rs.standard_normal(shp)).astype(np.float32))
        elif name.endswith("weight"):
            fan_in = int(np.prod(shp[1:]))
            fan[name[:-len("weight")]] = fan_in
            w = rs.standard_normal(shp).astype(np.float32) * np.float32(np.sqrt(2.0 / fan_in))
            sd[name] = torch.from_numpy(w)
        else:
            fan_in = fan.get(name[:-len("bias")])
            if fan_in is None:
                sd[name] = torch.from_numpy((0.1 * rs.standard_normal(shp)).astype(np.float32))
            else:
                bound = bias_scale / np.sqrt(fan_in)
                sd[name] = torch.from_numpy(rs.uniform(-bound, bound, shp).astype(np.float32))
    return sd


def phantom_labels(rs, B, S):
    """Nested-sphere phantom with BraTS-like class fractions (≈97/1/1.5/0.5 %)."""
    g = np.stack(np.meshgrid(*(np.arange(S),) * 3, indexing="ij"), 0).astype(np.float32)
    out = np.zeros((B, S, S, S), np.int64)
    for b in range(B):
        c = rs.uniform(0.35 * S, 0.65 * S, 3).astype(np.float32)
        r = np.sqrt(((g - c[:, None, None, None]) ** 2).sum(0))
        r3, r2, r1 = 0.105 * S, 0.165 * S, 0.195 * S
        lab = np.zeros((S, S, S), np.int64)
        lab[r < r1] = 2          # edema shell
        lab[r < r2] = 1          # necrotic / non-enhancing
        lab[r < r3] = 3          # enhancing core
        out[b] = lab
    return out


def mr2468_mask_ids():
    """mask_id column of the reference's Brats2020_imb_split_mr2468.csv (golden copy under tests/golden)."""
    import csv
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "tests", "golden", "Brats2020_imb_split_mr2468.csv")
    with open(path) as f:
        return [int(r["mask_id"]) for r in csv.DictReader(f)]


def make_batch(B, S, seed=1037, labels="U", mask_ids=None):
    """Synthetic batch: x f32 [B,4,S,S,S], target one-hot float64 [B,4,S,S,S], mask bool [B,4]."""
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((B, 4, S, S, S)).astype(np.float32)
    if labels == "U":
        y = rs.randint(0, 4, (B, S, S, S))
    else:
        y = phantom_labels(rs, B, S)
    if mask_ids is None:
        table = mr2468_mask_ids()
        mask_ids = [table[i] for i in rs.randint(0, len(table), B)]
    mask = MASK_ARRAY[np.asarray(mask_ids)]
    target = np.eye(4)[y].transpose(0, 4, 1, 2, 3)           # float64 one-hot, as datasets_nii.py:150-153
    return (torch.from_numpy(x), torch.from_numpy(np.ascontiguousarray(target)),
            torch.from_numpy(mask.copy()), torch.from_numpy(y))
