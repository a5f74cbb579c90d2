"""Device-side training-sample pipeline.

The reference builds every training item on the CPU inside DataLoader workers (train.py:122-128): np.load of the
preprocessed case, options.py:50's transform chain (numpy / scipy, data/transforms.py), the float64 one-hot target
(data/datasets_nii.py:150-153, 33 MB per batch of two), then `.cuda()`.  Here the preprocessed cases stay RESIDENT in HBM
(`ResidentCases`: 143 MB per 240x240x155x4 float32 case, 31 GB for the 219 BraTS2020 training cases — a sixth of one
B200), the host only draws the random parameters — in the reference's order, from the same generators, so a seeded run
sees the same crops / angles / flips / factors (`AugmentSampler`) — and ONE kernel launch per batch (`DeviceAugment`,
C ABI pb_augment_batch) produces Model.forward's inputs bit-exactly: x float32 [B,4,S,S,S] and the labels as a uint8 map
[B,S,S,S] (1 MB instead of the 33 MB one-hot; `Model.forward` / `criterions.*_bs` accept either form as `target`).
Per step the host sends ~10 KB of parameters instead of 49 MB of tensors.

No CPU fallback: DeviceAugment raises without the CUDA library / on host tensors.
"""
import ctypes
import os
import random as _pyrandom

import numpy as np
import torch

from . import _lib

ROT_AXES = ((1, 0), (2, 1), (2, 0))              # data/transforms.py:90, axes of [H, W, Z]

# scipy.special.cosdg / sindg of 0..45 degrees (what scipy.ndimage.rotate builds its matrix from; 2 of the 46 pairs
# differ from libm's cos/sin of the radian angle in the last bit, which can move a nearest-neighbour pick).
# Generated with scipy 1.18.1 (`[special.cosdg(a).hex() for a in range(46)]`); tests/test_augment.py re-checks them.
_COSDG = tuple(float.fromhex(h) for h in (
    "0x1.0000000000000p+0", "0x1.ffec097f5af8ap-1", "0x1.ffb0278bf0567p-1", "0x1.ff4c5ed12e61dp-1", "0x1.fec0b7170fff6p-1",
    "0x1.fe0d3b41815a2p-1", "0x1.fd31f94f867c6p-1", "0x1.fc2f025a23e8bp-1", "0x1.fb046a930947ap-1", "0x1.f9b24942fe45cp-1",
    "0x1.f838b8c811c17p-1", "0x1.f697d6938b6c2p-1", "0x1.f4cfc327a0080p-1", "0x1.f2e0a214e870fp-1", "0x1.f0ca99f79ba25p-1",
    "0x1.ee8dd4748bf15p-1", "0x1.ec2a7e35e7b80p-1", "0x1.e9a0c6e7bdb1fp-1", "0x1.e6f0e134454ffp-1", "0x1.e41b02bfeb4cbp-1",
    "0x1.e11f642522d1cp-1", "0x1.ddfe40effb805p-1", "0x1.dab7d7997cb58p-1", "0x1.d74c6982c666fp-1", "0x1.d3bc3aeff7f95p-1",
    "0x1.d0079302dd767p-1", "0x1.cc2ebbb5638cap-1", "0x1.c83201d3d2c6dp-1", "0x1.c411b4f6d2708p-1", "0x1.bfce277d339c6p-1",
    "0x1.bb67ae8584cabp-1", "0x1.b6dea1e76eadep-1", "0x1.b2335c2cda945p-1", "0x1.ad663a8ae2fdcp-1", "0x1.a8779cda8eea4p-1",
    "0x1.a367e59158747p-1", "0x1.9e3779b97f4a8p-1", "0x1.98e6c0ea27a14p-1", "0x1.9376253f463d1p-1", "0x1.8de613515a328p-1",
    "0x1.8836fa2cf503ap-1", "0x1.82694b4a11c37p-1", "0x1.7c7d7a833bec2p-1", "0x1.7673fe0c86982p-1", "0x1.704d4e6a54d39p-1",
    "0x1.6a09e667f3bccp-1"))
_SINDG = tuple(float.fromhex(h) for h in (
    "0x0.0p+0", "0x1.1df0b2b89dd1ep-6", "0x1.1de58c9f7dc27p-5", "0x1.acbc748efc90ep-5", "0x1.1db8f6d6a5128p-4",
    "0x1.64fd6b8c28102p-4", "0x1.ac2609b3c576cp-4", "0x1.f32d44c4f62d3p-4", "0x1.1d06c968d9e19p-3", "0x1.4060b67a85375p-3",
    "0x1.63a1a7e0b7389p-3", "0x1.86c6ddd76624fp-3", "0x1.a9cd9ac4258f6p-3", "0x1.ccb3236cdc675p-3", "0x1.ef74bf2e4b91dp-3",
    "0x1.0907dc1930690p-2", "0x1.1a40add328e29p-2", "0x1.2b637cf83d5c8p-2", "0x1.3c6ef372fe94fp-2", "0x1.4d61bd000cddcp-2",
    "0x1.5e3a8748a0bf5p-2", "0x1.6ef801fced33cp-2", "0x1.7f98deee59681p-2", "0x1.901bd2298ffabp-2", "0x1.a07f921061ad1p-2",
    "0x1.b0c2d77379853p-2", "0x1.c0e45dabe05c8p-2", "0x1.d0e2e2b44de00p-2", "0x1.e0bd274245079p-2", "0x1.f071eedefa0edp-2",
    "0x1.fffffffffffffp-2", "0x1.07b3120fddf13p-1", "0x1.0f5193eacdd2ap-1", "0x1.16daed770771dp-1", "0x1.1e4e88411fd13p-1",
    "0x1.25abcf87c4978p-1", "0x1.2cf2304755a5ep-1", "0x1.342119455beb6p-1", "0x1.3b37fb1bdc939p-1", "0x1.4236484487abdp-1",
    "0x1.491b7523c161cp-1", "0x1.4fe6f81384fd4p-1", "0x1.5698496e20bd8p-1", "0x1.5d2ee398c9c2bp-1", "0x1.63aa430e07310p-1",
    "0x1.6a09e667f3bcdp-1"))


def rotation_params(angle, n0, n1):
    """The 2x2 matrix and offset scipy.ndimage.rotate(reshape=False) hands to its geometric transform for an integer
    `angle` (degrees, |angle| <= 45) in a plane of extent (n0, n1): M = [[c, s], [-s, c]], offset = centre - M centre."""
    a = int(angle)
    if abs(a) > 45:
        raise ValueError("rotation_params: |angle| must be <= 45 degrees")
    c, s = _COSDG[abs(a)], (_SINDG[abs(a)] if a >= 0 else -_SINDG[abs(a)])
    m = np.array([[c, s], [-s, c]])
    center = (np.array([n0, n1]) - 1) / 2
    return m, center - m @ center


class AugmentSampler:
    """The random draws of one training item in the reference's order (Compose.sample, then the draws that
    RandomIntensityChange.tf makes for the image):  3 x random.randint (crop origin), np choice (rotation plane),
    np randint (angle), 3 x np choice (flips), np uniform shift [S0,4], np uniform scale [S0,4].
    `py_rng` / `np_rng` default to the GLOBAL `random` / `numpy.random` the reference draws from, so seeding those as
    train.py does reproduces its stream; pass random.Random(seed) / np.random.RandomState(seed) for private streams."""

    def __init__(self, size=(80, 80, 80), angle_spectrum=10, shift=0.1, scale=0.1, py_rng=None, np_rng=None):
        self.size = tuple(int(s) for s in size)
        self.angle_spectrum, self.shift, self.scale = int(angle_spectrum), float(shift), float(scale)
        self.py_rng = py_rng if py_rng is not None else _pyrandom
        self.np_rng = np_rng if np_rng is not None else np.random

    def sample(self, shape):
        """shape = (H, W, Z) of the case -> dict(start, axes, angle, flip, shift [S0,4], scale [S0,4])."""
        size = self.size
        if any(s < n for s, n in zip(shape, size)):
            raise ValueError(f"case of shape {tuple(shape)} is smaller than the crop {size}")
        start = [self.py_rng.randint(0, s - n) for n, s in zip(size, shape)]
        axes = ROT_AXES[self.np_rng.choice(list(range(len(ROT_AXES))))]
        angle = int(self.np_rng.randint(-self.angle_spectrum, self.angle_spectrum))
        flip = [bool(self.np_rng.choice([True, False])) for _ in range(3)]
        shift = self.np_rng.uniform(-self.shift, self.shift, size=[1, size[0], 1, 1, 4])
        scale = self.np_rng.uniform(1.0 - self.scale, 1.0 + self.scale, size=[1, size[0], 1, 1, 4])
        return dict(start=start, axes=tuple(axes), angle=angle, flip=flip, shift=shift.reshape(size[0], 4),
                    scale=scale.reshape(size[0], 4))


class ResidentCases:
    """Preprocessed cases kept in HBM: vol float32 [H,W,Z,4] and seg uint8 [H,W,Z] per case (the layout of
    `<root>/vol/<name>_vol.npy`, `<root>/seg/<name>_seg.npy`, preprocessing/preprocess_brats.py:71-84)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.vols, self.segs, self.names = [], [], []

    def add(self, vol, seg, name=None):
        vol = torch.as_tensor(np.ascontiguousarray(vol) if isinstance(vol, np.ndarray) else vol)
        seg = torch.as_tensor(np.ascontiguousarray(seg) if isinstance(seg, np.ndarray) else seg)
        if vol.dtype != torch.float32 or vol.dim() != 4 or vol.shape[-1] != 4:
            raise ValueError("vol must be float32 [H, W, Z, 4]")
        if seg.dtype != torch.uint8 or tuple(seg.shape) != tuple(vol.shape[:3]):
            raise ValueError("seg must be uint8 [H, W, Z] matching vol")
        self.vols.append(vol.to(self.device).contiguous())
        self.segs.append(seg.to(self.device).contiguous())
        self.names.append(name if name is not None else str(len(self.names)))
        return len(self.vols) - 1

    def add_files(self, root, names):
        for n in names:
            self.add(np.load(os.path.join(root, "vol", n + "_vol.npy")),
                     np.load(os.path.join(root, "seg", n + "_seg.npy")).astype(np.uint8), n)
        return self

    def __len__(self):
        return len(self.vols)

    def nbytes(self):
        return sum(v.numel() * 4 + s.numel() for v, s in zip(self.vols, self.segs))


def pack_batch(host, cases, case_ids, params, size):
    """Writes one batch's parameter block into the uint8 tensor `host` (pinned in DeviceAugment; any CPU tensor works):
    B pb_augment_sample records, then the float64 factor tables scale [B,S0,4] and shift [B,S0,4]."""
    B, (s0, s1, s2) = len(case_ids), size
    rec = ctypes.sizeof(_lib.AugmentSample)
    buf = host.numpy()
    recs = (_lib.AugmentSample * B).from_buffer(buf, 0)
    fac = buf[B * rec:B * rec + 2 * B * s0 * 4 * 8].view(np.float64).reshape(2, B, s0, 4)
    n = (s0, s1, s2)
    for b, (cid, p) in enumerate(zip(case_ids, params)):
        vol, seg = cases.vols[cid], cases.segs[cid]
        if vol.data_ptr() % 16:
            raise RuntimeError("case volume is not 16-byte aligned")
        H, W, Z = vol.shape[:3]
        st = [int(v) for v in p["start"]]
        if any(o < 0 or o + m > e for o, m, e in zip(st, n, (H, W, Z))):
            raise ValueError(f"crop {st}+{n} leaves the case of shape {(H, W, Z)}")
        a0, a1 = sorted(int(a) for a in p["axes"])
        m, off = rotation_params(p["angle"], n[a0], n[a1])
        r = recs[b]
        r.vol, r.seg = vol.data_ptr(), seg.data_ptr()
        r.shape[:] = (H, W, Z)
        r.start[:] = st
        r.flip[:] = [int(bool(f)) for f in p["flip"]]
        r.rot_axes[:] = (a0, a1)
        r.rot_m[:] = (m[0, 0], m[0, 1], m[1, 0], m[1, 1])
        r.rot_off[:] = (off[0], off[1])
        fac[0, b] = np.asarray(p["scale"], np.float64).reshape(s0, 4)
        fac[1, b] = np.asarray(p["shift"], np.float64).reshape(s0, 4)
    del recs


class DeviceAugment:
    """One pb_augment_batch launch per batch.  `slots` parameter buffers (pinned host + device) are reused in turn, so a
    call never waits unless the launch that last used its slot (`slots` calls ago) is still in flight."""

    def __init__(self, device, size=(80, 80, 80), batch=2, want_onehot=False, slots=4):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("passion_b200.data.DeviceAugment runs on CUDA only (no CPU fallback)")
        self.lib = _lib.load()
        self.size = tuple(int(s) for s in size)
        self.batch, self.want_onehot = int(batch), bool(want_onehot)
        self.rec = ctypes.sizeof(_lib.AugmentSample)
        self.fac = self.size[0] * 4 * 8
        nbytes = self.batch * (self.rec + 2 * self.fac)
        self.host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.dev = [torch.empty(nbytes, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        self.done = [None] * slots
        self.k = 0

    def param_bytes(self):
        return self.batch * (self.rec + 2 * self.fac)

    def __call__(self, cases, case_ids, params, out=None):
        """cases: ResidentCases; case_ids / params: one entry per sample of the batch (params from AugmentSampler.sample).
        Returns (x float32 [B,4,S0,S1,S2], labels uint8 [B,S0,S1,S2], onehot float64 [B,4,S0,S1,S2] or None); `out` may
        hand in preallocated (x, labels, onehot) tensors (CUDA-graph static inputs)."""
        B, (s0, s1, s2) = self.batch, self.size
        if len(case_ids) != B or len(params) != B:
            raise ValueError(f"expected {B} samples")
        slot = self.k % len(self.host)
        self.k += 1
        if self.done[slot] is not None:
            self.done[slot].synchronize()               # the copy that last read this pinned buffer has finished
        pack_batch(self.host[slot], cases, case_ids, params, self.size)
        dev = self.dev[slot]
        dev.copy_(self.host[slot], non_blocking=True)
        if out is None:
            x = torch.empty((B, 4, s0, s1, s2), dtype=torch.float32, device=self.device)
            labels = torch.empty((B, s0, s1, s2), dtype=torch.uint8, device=self.device)
            onehot = torch.empty((B, 4, s0, s1, s2), dtype=torch.float64, device=self.device) if self.want_onehot else None
        else:
            x, labels, onehot = out
            for t, shp, dt in ((x, (B, 4, s0, s1, s2), torch.float32), (labels, (B, s0, s1, s2), torch.uint8),
                               (onehot, (B, 4, s0, s1, s2), torch.float64)):
                if t is not None and (tuple(t.shape) != shp or t.dtype != dt or not t.is_cuda or not t.is_contiguous()):
                    raise ValueError("DeviceAugment: bad output tensor")
        base = dev.data_ptr()
        scale = base + B * self.rec
        vp = ctypes.c_void_p
        rc = self.lib.pb_augment_batch(vp(base), vp(scale), vp(scale + B * self.fac), B, s0, s1, s2, vp(x.data_ptr()),
                                       vp(labels.data_ptr()) if labels is not None else None,
                                       vp(onehot.data_ptr()) if onehot is not None else None,
                                       vp(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(rc, "augment_batch")
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.done[slot] = ev
        return x, labels, onehot
