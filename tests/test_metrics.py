"""Dice scoring of the inference sweep (passion_b200/metrics.py) against the UNMODIFIED reference function
`softmax_output_dice_class4` (utils/predict.py:82-128; fixtures by oracle/gen_golden_metrics.py) — bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle.gen_golden_metrics import CASES, label_maps
from passion_b200 import metrics

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_dice.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_dice_matches_reference_bit_exact(name):
    pred, target = label_maps(*CASES[name])
    sep, ev = metrics.dice_class4(torch.from_numpy(pred), torch.from_numpy(target))
    assert sep.dtype == torch.float32 and ev.dtype == torch.float32
    assert np.array_equal(sep.numpy(), GOLD[name + "_separate"])
    assert np.array_equal(ev.numpy(), GOLD[name + "_evaluate"])


def test_uint8_label_maps_give_the_same_scores():
    pred, target = label_maps(*CASES["brats_like"])
    a = metrics.dice_class4(torch.from_numpy(pred), torch.from_numpy(target))
    b = metrics.dice_class4(torch.from_numpy(pred.astype(np.uint8)), torch.from_numpy(target.astype(np.uint8)))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_confusion_counts():
    pred, target = label_maps(7, 3, (9, 7, 5), (.4, .3, .2, .1), (.25, .25, .25, .25))
    cm = metrics.confusion_counts(torch.from_numpy(pred), torch.from_numpy(target))
    assert cm.shape == (3, 4, 4) and int(cm.sum()) == pred.size
    for b in range(3):
        for i in range(4):
            for j in range(4):
                assert int(cm[b, i, j]) == int(((pred[b] == i) & (target[b] == j)).sum())
    with pytest.raises(ValueError):
        metrics.confusion_counts(torch.full((1, 4), 4), torch.zeros((1, 4), dtype=torch.long))
    with pytest.raises(ValueError):
        metrics.confusion_counts(torch.zeros((1, 4)), torch.zeros((1, 5)))


def test_average_meter():
    m = metrics.AverageMeter()
    m.update(np.array([1.0, 3.0]))
    m.update(np.array([3.0, 5.0]), n=3)
    assert np.allclose(m.avg, [2.5, 4.5]) and m.count == 4 and np.allclose(m.val, [3.0, 5.0])


def test_evaluate_all_masks_scores_each_mask_like_a_separate_reference_call(monkeypatch):
    """The sweep scores 15 label maps at once; each row must equal the reference function applied to that mask alone
    (its < 500-voxel post-processing rule is per call), in the reference's reversed mask order."""
    from passion_b200 import predict
    rs = np.random.RandomState(0)
    shape = (20, 24, 16)
    target = rs.choice(4, size=(1,) + shape, p=(.9, .04, .04, .02))
    maps = np.stack([np.where(rs.rand(*shape) < 0.5 + 0.03 * m, target[0], rs.choice(4, size=shape, p=(.7, .1, .1, .1)))
                     for m in range(15)])
    maps[3][maps[3] == 3] = 0                                            # one mask predicts no enhancing tumour at all
    monkeypatch.setattr(predict, "predict_all_masks", lambda model, x, masks, patch: (torch.from_numpy(maps), None))
    names = [f"m{i}" for i in range(15)]
    res = metrics.evaluate_all_masks(None, None, torch.from_numpy(target), mask_names=names)
    assert list(res) == names[::-1]
    for i, n in enumerate(names):
        _, ev = metrics.dice_class4(torch.from_numpy(maps[i:i + 1]), torch.from_numpy(target))
        assert torch.equal(res[n], ev[0])
