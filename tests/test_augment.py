"""CPU tests of the training-sample pipeline: the oracle against the reference's golden outputs, the
numpy restatement of scipy's nearest-neighbour rotation against scipy itself, and the product's HOST logic (draw order,
rotation matrices, parameter-block packing) against the oracle.  The kernel itself is checked in test_augment_gpu.py."""
import ctypes
import glob
import os
import random

import numpy as np
import pytest
import torch

from oracle import augment_oracle as ao
from passion_b200 import _lib, data

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[len("augment_"):-4] for p in glob.glob(os.path.join(GOLD, "augment_*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLD, f"augment_{name}.npz"))
    vol, seg = ao.synth_volume(int(z["vseed"]), tuple(int(v) for v in z["vshape"]))
    return z, vol, seg


def test_fixtures_present():
    assert len(CASES) >= 6


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("use_scipy", [True, False])
def test_oracle_matches_reference_golden(name, use_scipy):
    """oracle (scipy rotate / plain-numpy rotate) == the unmodified reference transforms, bit for bit."""
    z, vol, seg = load_case(name)
    size = tuple(int(v) for v in z["size"])
    p = ao.sample(vol.shape[:3], size, random.Random(int(z["py_seed"])), np.random.RandomState(int(z["np_seed"])))
    assert p["start"] == z["start"].tolist() and tuple(p["axes"]) == tuple(z["axes"].tolist())
    assert p["angle"] == int(z["angle"]) and p["flip"] == z["flip"].tolist()
    x, y, yo = ao.apply(vol, seg, p, use_scipy=use_scipy)
    assert x.dtype == np.float32 and np.array_equal(x, z["x"])
    assert np.array_equal(y, z["y"].astype(np.int64))
    assert yo.dtype == np.float64 and np.array_equal(yo.argmax(0), y) and np.all(yo.sum(0) == 1.0)


def test_rotate_nearest_matches_scipy():
    from scipy.ndimage import rotate
    rs = np.random.RandomState(0)
    x = rs.standard_normal((20, 16, 24)).astype(np.float32)
    y = rs.randint(0, 4, (20, 16, 24)).astype(np.uint8)
    for axes in ao.ROT_AXES:
        for ang in range(-10, 10):
            kw = dict(axes=axes, reshape=False, order=0, mode="constant", cval=-1)
            assert np.array_equal(rotate(x, ang, **kw), ao.rotate_nearest(x, ang, axes))
            assert np.array_equal(rotate(y, ang, **kw), ao.rotate_nearest(y, ang, axes))


def test_rotation_table_matches_scipy():
    from scipy import special
    for a in range(-45, 46):
        m, off = data.rotation_params(a, 80, 64)
        c, s = special.cosdg(a), special.sindg(a)
        assert m[0, 0] == c and m[0, 1] == s and m[1, 0] == -s and m[1, 1] == c
        center = (np.array([80, 64]) - 1) / 2
        assert np.array_equal(off, center - np.array([[c, s], [-s, c]]) @ center)
    with pytest.raises(ValueError):
        data.rotation_params(46, 8, 8)


@pytest.mark.parametrize("seeds", [(5, 7), (11, 1037), (0, 0)])
def test_sampler_draw_order(seeds):
    """AugmentSampler consumes the generators exactly like the reference (same values, same generator state after)."""
    shape, size = (40, 44, 36), (16, 12, 20)
    r1, n1 = random.Random(seeds[0]), np.random.RandomState(seeds[1])
    r2, n2 = random.Random(seeds[0]), np.random.RandomState(seeds[1])
    smp = data.AugmentSampler(size, py_rng=r1, np_rng=n1)
    for _ in range(3):
        a, b = smp.sample(shape), ao.sample(shape, size, r2, n2)
        assert a["start"] == b["start"] and tuple(a["axes"]) == tuple(b["axes"]) and a["angle"] == b["angle"] and a["flip"] == b["flip"]
        assert np.array_equal(a["shift"], b["shift"].reshape(size[0], 4)) and np.array_equal(a["scale"], b["scale"].reshape(size[0], 4))
    assert r1.random() == r2.random() and n1.rand() == n2.rand()
    with pytest.raises(ValueError):
        smp.sample((10, 44, 36))


def test_sampler_defaults_to_global_generators():
    random.seed(3); np.random.seed(4)
    a = data.AugmentSampler((16, 16, 16)).sample((40, 44, 36))
    random.seed(3); np.random.seed(4)
    b = ao.sample((40, 44, 36), (16, 16, 16))
    assert a["start"] == b["start"] and a["angle"] == b["angle"] and np.array_equal(a["scale"], b["scale"].reshape(16, 4))


class _HostCases:
    def __init__(self, vol, seg):
        self.vols, self.segs = [torch.from_numpy(vol)], [torch.from_numpy(seg)]


def _emulate_kernel(block, B, size, cases):
    """numpy transcription of csrc/augment.cu reading the packed parameter block (records + factor tables)."""
    s0, s1, s2 = size
    rec = ctypes.sizeof(_lib.AugmentSample)
    recs = (_lib.AugmentSample * B).from_buffer(block, 0)
    fac = block[B * rec:B * rec + 2 * B * s0 * 32].view(np.float64).reshape(2, B, s0, 4)
    xs, ls = [], []
    for b in range(B):
        r = recs[b]
        vol, seg = cases.vols[0].numpy(), cases.segs[0].numpy()
        assert r.vol == cases.vols[0].data_ptr() and r.seg == cases.segs[0].data_ptr() and tuple(r.shape) == vol.shape[:3]
        i, j, k = np.meshgrid(np.arange(s0), np.arange(s1), np.arange(s2), indexing="ij")
        p = [s0 - 1 - i if r.flip[0] else i, s1 - 1 - j if r.flip[1] else j, s2 - 1 - k if r.flip[2] else k]
        a0, a1 = r.rot_axes[0], r.rot_axes[1]
        n = (s0, s1, s2)
        o0, o1 = p[a0].astype(np.float64), p[a1].astype(np.float64)
        c0 = (o0 * r.rot_m[0] + o1 * r.rot_m[1]) + r.rot_off[0]
        c1 = (o0 * r.rot_m[2] + o1 * r.rot_m[3]) + r.rot_off[1]
        inb = (c0 >= 0) & (c0 <= n[a0] - 1) & (c1 >= 0) & (c1 <= n[a1] - 1)
        q = list(p)
        q[a0] = np.where(inb, np.floor(c0 + 0.5), 0).astype(np.int64)
        q[a1] = np.where(inb, np.floor(c1 + 0.5), 0).astype(np.int64)
        v = vol[r.start[0] + q[0], r.start[1] + q[1], r.start[2] + q[2]]            # [s0,s1,s2,4]
        v = np.where(inb[..., None], v, np.float32(-1))
        lab = np.where(inb, seg[r.start[0] + q[0], r.start[1] + q[1], r.start[2] + q[2]], 0).astype(np.uint8)
        x = (v.astype(np.float64) * fac[0, b][p[0]] + fac[1, b][p[0]]).astype(np.float32)
        xs.append(x.transpose(3, 0, 1, 2))
        ls.append(lab)
    del recs
    return np.stack(xs), np.stack(ls)


@pytest.mark.parametrize("name", CASES)
def test_packed_block_reproduces_golden(name):
    """pack_batch (product host code) + a numpy transcription of the kernel's arithmetic == the reference's output:
    pins the record layout, the rotation matrices / offsets and the index conventions without a GPU."""
    z, vol, seg = load_case(name)
    size = tuple(int(v) for v in z["size"])
    smp = data.AugmentSampler(size, py_rng=random.Random(int(z["py_seed"])), np_rng=np.random.RandomState(int(z["np_seed"])))
    p = smp.sample(vol.shape[:3])
    cases = _HostCases(vol, seg)
    rec = ctypes.sizeof(_lib.AugmentSample)
    host = torch.zeros(rec + 2 * size[0] * 32, dtype=torch.uint8)
    data.pack_batch(host, cases, [0], [p], size)
    x, lab = _emulate_kernel(host.numpy(), 1, size, cases)
    assert np.array_equal(x[0], z["x"]) and np.array_equal(lab[0], z["y"])


def test_pack_rejects_bad_crops():
    vol, seg = ao.synth_volume(1, (20, 20, 20))
    cases = _HostCases(vol, seg)
    p = dict(start=[8, 0, 0], axes=(1, 0), angle=0, flip=[False] * 3, shift=np.zeros((16, 4)), scale=np.ones((16, 4)))
    host = torch.zeros(ctypes.sizeof(_lib.AugmentSample) + 2 * 16 * 32, dtype=torch.uint8)
    with pytest.raises(ValueError):
        data.pack_batch(host, cases, [0], [p], (16, 16, 16))


def test_device_augment_refuses_cpu():
    with pytest.raises(RuntimeError):
        data.DeviceAugment("cpu")


def test_label_map_target_equals_onehot_stats():
    """criterions.label_stats: the uint8 label map gives the same labels / counts / class weights as the one-hot target."""
    from passion_b200 import criterions as crit
    rs = np.random.RandomState(0)
    y = rs.randint(0, 4, (2, 6, 5, 4))
    y[1][y[1] == 3] = 0                                                   # a class absent from one sample
    onehot = torch.from_numpy(np.ascontiguousarray(np.eye(4)[y].transpose(0, 4, 1, 2, 3)))
    la, ca, wa = crit.label_stats(onehot)
    lb, cb, wb = crit.label_stats(torch.from_numpy(y.astype(np.uint8)))
    assert torch.equal(la, lb) and torch.equal(ca, cb) and torch.equal(wa, wb)
