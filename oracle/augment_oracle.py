"""Oracle (TEST INFRASTRUCTURE): the reference's training-sample pipeline, restated on numpy + scipy.

Follows (reference, /root/reference/code):
  options.py:50                   Compose([RandCrop3D((80,80,80)), RandomRotion(10), RandomIntensityChange((0.1,0.1)),
                                           RandomFlip(0), NumpyType((np.float32, np.int64))])
  data/transforms.py:14-36        Base.__call__: sample() once per item, then tf(x, k=0), tf(y, k=1)
  data/transforms.py:407-418      RandCrop3D.sample        3 x random.randint(0, s - size)          (python `random`)
  data/transforms.py:86-120       RandomRotion             np.random.choice(3) -> axes; np.random.randint(-a, a) -> angle;
                                                           scipy.ndimage.rotate(order=0, mode='constant', cval=-1, reshape=False)
  data/transforms.py:133-155      RandomIntensityChange    x only: np.random.uniform shift then scale, size [1, H, 1, 1, C]
  data/transforms.py:217-240      RandomFlip               3 x np.random.choice([True, False]); np.flip on axes 1, 2, 3
  data/transforms.py:378-390      NumpyType                x -> float32, y -> int64
  data/datasets_nii.py:141-160    x[None], y[None]; transpose to [C,H,W,Z]; one-hot np.eye(4)[y] float64 -> [4,H,W,Z]
Third-party arithmetic: scipy.ndimage.rotate (reference pin scipy==1.8.1, requirements.txt:51; this image has
scipy 1.18.1 — same nearest-neighbour geometric transform).  `apply()` calls scipy itself; `rotate_nearest()` is the
plain-numpy restatement of that call which the CUDA kernel implements, pinned bit-exactly against scipy in
tests/test_augment.py.  The whole module is pinned against the UNMODIFIED reference transforms by
oracle/gen_golden_augment.py (fixtures tests/golden/augment_*.npz).
"""
import random as _pyrandom

import numpy as np

ROT_AXES = [(1, 0), (2, 1), (2, 0)]          # transforms.py:90 (indices into [H, W, Z])


def sample(shape, size, py_rng=_pyrandom, np_rng=np.random, angle_spectrum=10, shift=0.1, scale=0.1):
    """All random draws of one item, in the reference's order.  `py_rng` / `np_rng` default to the global generators the
    reference uses; pass random.Random(seed) / np.random.RandomState(seed) for private streams (same algorithms)."""
    start = [py_rng.randint(0, s - i) for i, s in zip(size, shape)]                      # RandCrop3D.sample
    axes = ROT_AXES[np_rng.choice(list(range(len(ROT_AXES))))]                            # RandomRotion.sample
    angle = int(np_rng.randint(-angle_spectrum, angle_spectrum))
    flip = [bool(np_rng.choice([True, False])) for _ in range(3)]                         # RandomFlip.sample
    sh = np_rng.uniform(-shift, shift, size=[1, size[0], 1, 1, 4])                        # RandomIntensityChange.tf (k = 0)
    sc = np_rng.uniform(1.0 - scale, 1.0 + scale, size=[1, size[0], 1, 1, 4])
    return dict(start=start, size=list(size), axes=axes, angle=angle, flip=flip, shift=sh, scale=sc)


def rotate_nearest(vol, angle, axes, cval=-1):
    """Plain-numpy restatement of scipy.ndimage.rotate(vol, angle, axes, reshape=False, order=0, mode='constant', cval):
    in = M out + offset in float64 (products and sums in scipy's order, no fused multiply-add), a coordinate strictly
    outside [0, n-1] gives cval, else the voxel floor(in + 0.5).  Unsigned outputs clamp a negative cval to 0."""
    from scipy import special
    a0, a1 = sorted(axes)
    n0, n1 = vol.shape[a0], vol.shape[a1]
    c, s = special.cosdg(angle), special.sindg(angle)
    m = np.array([[c, s], [-s, c]])
    center = (np.array([n0, n1]) - 1) / 2
    off = center - m @ center
    o0, o1 = np.meshgrid(np.arange(n0, dtype=np.float64), np.arange(n1, dtype=np.float64), indexing="ij")
    c0 = (o0 * m[0, 0] + o1 * m[0, 1]) + off[0]
    c1 = (o0 * m[1, 0] + o1 * m[1, 1]) + off[1]
    inb = (c0 >= 0) & (c0 <= n0 - 1) & (c1 >= 0) & (c1 <= n1 - 1)
    i0 = np.floor(c0 + 0.5).astype(np.int64).clip(0, n0 - 1)
    i1 = np.floor(c1 + 0.5).astype(np.int64).clip(0, n1 - 1)
    v = np.moveaxis(vol, (a0, a1), (0, 1))
    fill = 0 if (vol.dtype.kind == "u" and cval < 0) else cval
    out = np.where(inb.reshape(inb.shape + (1,) * (v.ndim - 2)), v[i0, i1], np.asarray(fill, dtype=vol.dtype))
    return np.moveaxis(out, (0, 1), (a0, a1))


def apply(vol, seg, p, num_cls=4, use_scipy=True):
    """vol [H,W,Z,4] float32, seg [H,W,Z] uint8, p = sample(...) -> x float32 [4,S0,S1,S2], y int64 [S0,S1,S2],
    one-hot float64 [4,S0,S1,S2]  (what Brats_loadall_train_nii_idt.__getitem__ returns, datasets_nii.py:141-160)."""
    from scipy.ndimage import rotate
    s, n = p["start"], p["size"]
    x = vol[s[0]:s[0] + n[0], s[1]:s[1] + n[1], s[2]:s[2] + n[2]].copy()
    y = seg[s[0]:s[0] + n[0], s[1]:s[1] + n[1], s[2]:s[2] + n[2]].copy()
    if use_scipy:
        x = np.stack([rotate(x[..., c], p["angle"], axes=p["axes"], reshape=False, order=0, mode="constant", cval=-1)
                      for c in range(x.shape[-1])], -1)
        y = rotate(y, p["angle"], axes=p["axes"], reshape=False, order=0, mode="constant", cval=-1)
    else:
        x = rotate_nearest(x, p["angle"], p["axes"])
        y = rotate_nearest(y, p["angle"], p["axes"])
    x = x[None] * p["scale"] + p["shift"]                      # float32 * float64 -> float64
    y = y[None]
    for ax, f in zip((1, 2, 3), p["flip"]):
        if f:
            x, y = np.flip(x, ax), np.flip(y, ax)
    x = x.astype(np.float32)
    y = y.astype(np.int64)
    xo = np.ascontiguousarray(x.transpose(0, 4, 1, 2, 3))[0]
    yo = np.ascontiguousarray(np.eye(num_cls)[y.reshape(-1)].reshape(*y.shape, -1).transpose(0, 4, 1, 2, 3))[0]
    return xo, y[0], yo


def synth_volume(seed, shape=(40, 44, 36)):
    """Seeded stand-in for a preprocessed BraTS case: vol [H,W,Z,4] float32, seg [H,W,Z] uint8 in {0..3}."""
    rs = np.random.RandomState(seed)
    vol = rs.standard_normal(tuple(shape) + (4,)).astype(np.float32)
    seg = rs.randint(0, 4, shape).astype(np.uint8)
    return vol, seg
