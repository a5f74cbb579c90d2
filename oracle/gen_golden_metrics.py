"""Oracle (TEST INFRASTRUCTURE): golden fixtures of the reference's Dice scoring.  Runs ONLY in the build container.

utils/predict.py cannot be imported here (it pulls in nibabel / medpy, which are not installed), so the UNMODIFIED source
of `softmax_output_dice_class4` (predict.py:82-128) is cut out of the file by its AST position and executed on seeded
label maps; its outputs go to tests/golden/metrics_dice.npz (inputs are regenerated from the seeds).
Usage:  python -m oracle.gen_golden_metrics
"""
import ast
import os

import numpy as np
import torch

REF = "/root/reference/code/utils/predict.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name: (seed, batch, shape, class probabilities of the prediction, of the target)
CASES = {
    "uniform": (1, 2, (24, 20, 28), (.25, .25, .25, .25), (.25, .25, .25, .25)),
    "brats_like": (2, 1, (48, 48, 40), (.95, .015, .03, .005), (.955, .012, .028, .005)),
    "few_enhancing": (3, 1, (32, 32, 32), (.97, .015, .012, .003), (.96, .02, .015, .005)),       # < 500 ET voxels predicted
    "no_tumour_pred": (4, 1, (16, 16, 16), (1., 0., 0., 0.), (.9, .04, .04, .02)),
    "empty_both": (5, 2, (8, 8, 8), (1., 0., 0., 0.), (1., 0., 0., 0.)),
}


def label_maps(seed, batch, shape, pp, pt):
    rs = np.random.RandomState(seed)
    pred = rs.choice(4, size=(batch,) + tuple(shape), p=pp)
    agree = rs.rand(batch, *shape) < 0.7                       # correlated with the prediction, like a real segmentation
    target = np.where(agree, pred, rs.choice(4, size=(batch,) + tuple(shape), p=pt))
    return pred.astype(np.int64), target.astype(np.int64)


def reference_fn():
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "softmax_output_dice_class4")
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns["softmax_output_dice_class4"]


def main():
    fn = reference_fn()
    out = {}
    for name, cfg in CASES.items():
        pred, target = label_maps(*cfg)
        sep, ev = fn(torch.from_numpy(pred), torch.from_numpy(target))
        out[name + "_separate"], out[name + "_evaluate"] = sep, ev
        print(name, ev)
    np.savez(os.path.join(ROOT, "tests", "golden", "metrics_dice.npz"), **out)


if __name__ == "__main__":
    main()
