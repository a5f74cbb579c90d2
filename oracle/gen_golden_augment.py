"""Oracle (TEST INFRASTRUCTURE): golden fixtures of the reference's training-sample pipeline.  Runs ONLY in the
build container (needs /root/reference).

Imports the UNMODIFIED reference `data.transforms` (with the `collections.Sequence` alias that Python >= 3.10 removed,
SURVEY.md §8c), evaluates options.py:50's transform string at a small crop size on seeded synthetic volumes, repeats
datasets_nii.py:141-160 (transposes + one-hot), and
  1. asserts that oracle/augment_oracle.py reproduces it BIT-EXACTLY (draws, x, y, one-hot), with scipy and with the
     plain-numpy rotation;
  2. writes the reference's outputs to tests/golden/augment_<case>.npz (inputs are regenerated from the seed).
Usage:  python -m oracle.gen_golden_augment
"""
import collections
import collections.abc
import os
import random
import sys

import numpy as np

REF = "/root/reference/code"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# name: (volume seed, volume shape, crop size, python-random seed, numpy seed)
CASES = {
    "a": (1, (40, 44, 36), (16, 16, 16), 5, 7),
    "b": (2, (40, 44, 36), (16, 16, 16), 11, 1037),
    "c": (3, (30, 26, 34), (24, 16, 20), 3, 99),          # non-cubic crop: exercises every axis pair with unequal extents
    "d": (4, (40, 44, 36), (16, 16, 16), 8, 2024),
    "e": (5, (24, 24, 24), (24, 24, 24), 1, 4),           # crop == volume (start 0)
    "f": (6, (40, 44, 36), (16, 16, 16), 21, 65),
}


def reference_item(vol, seg, size, py_seed, np_seed):
    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from data import transforms as T
    ns = {k: getattr(T, k) for k in dir(T)}
    ns["np"] = np
    text = ("Compose([RandCrop3D((80,80,80)), RandomRotion(10), RandomIntensityChange((0.1,0.1)), RandomFlip(0), "
            "NumpyType((np.float32, np.int64)),])").replace("(80,80,80)", str(tuple(size)))          # options.py:50
    tf = eval(text, ns)
    random.seed(py_seed)
    np.random.seed(np_seed)
    x, y = vol[None, ...].copy(), seg[None, ...].copy()                       # datasets_nii.py:141-146
    x, y = tf([x, y])
    x = np.ascontiguousarray(x.transpose(0, 4, 1, 2, 3))
    _, H, W, Z = np.shape(y)
    yl = y
    y = np.reshape(y, (-1))
    yo = np.reshape(np.eye(4)[y], (1, H, W, Z, -1))
    yo = np.ascontiguousarray(yo.transpose(0, 4, 1, 2, 3))
    crop, rot, flip = tf.ops[0], tf.ops[1], tf.ops[3]
    draws = dict(start=[s.start for s in crop.buffer[1:]], axes=tuple(rot.axes_buffer), angle=int(rot.angle_buffer),
                 flip=[bool(flip.x_buffer), bool(flip.y_buffer), bool(flip.z_buffer)])
    return x[0], yl[0], yo[0], draws


def main():
    from oracle import augment_oracle as ao
    os.makedirs(GOLD, exist_ok=True)
    for name, (vseed, vshape, size, py_seed, np_seed) in CASES.items():
        vol, seg = ao.synth_volume(vseed, vshape)
        rx, ry, ryo, draws = reference_item(vol, seg, size, py_seed, np_seed)
        p = ao.sample(vshape, size, random.Random(py_seed), np.random.RandomState(np_seed))
        assert p["start"] == draws["start"] and tuple(p["axes"]) == draws["axes"] and p["angle"] == draws["angle"] \
            and p["flip"] == draws["flip"], (name, p, draws)
        for use_scipy in (True, False):
            ox, oy, oyo = ao.apply(vol, seg, p, use_scipy=use_scipy)
            assert ox.dtype == rx.dtype == np.float32 and oy.dtype == ry.dtype == np.int64 and oyo.dtype == ryo.dtype == np.float64
            assert np.array_equal(ox, rx) and np.array_equal(oy, ry) and np.array_equal(oyo, ryo), (name, use_scipy)
        np.savez_compressed(os.path.join(GOLD, f"augment_{name}.npz"), vseed=vseed, vshape=vshape, size=size, py_seed=py_seed,
                            np_seed=np_seed, start=draws["start"], axes=draws["axes"], angle=draws["angle"], flip=draws["flip"],
                            x=rx, y=ry.astype(np.uint8))
        print(f"{name}: start {draws['start']} axes {draws['axes']} angle {draws['angle']} flip {draws['flip']} "
              f"filled voxels {(ry != oy).sum()} | oracle == reference (bit-exact)")


if __name__ == "__main__":
    main()
