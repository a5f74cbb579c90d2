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
    class SharedEncoderModel:                                            # has the shared-encoder API -> predict_all_masks
        mask_type = "idt"

        def _features(self):
            pass
    res = metrics.evaluate_all_masks(SharedEncoderModel(), None, torch.from_numpy(target), mask_names=names)
    assert list(res) == names[::-1]
    for i, n in enumerate(names):
        _, ev = metrics.dice_class4(torch.from_numpy(maps[i:i + 1]), torch.from_numpy(target))
        assert torch.equal(res[n], ev[0])


def test_evaluate_all_masks_falls_back_to_one_sweep_per_mask(monkeypatch):
    """A model without the shared-encoder API (mmFormer) is scored through predict_volume, one mask at a time."""
    from passion_b200 import predict
    rs = np.random.RandomState(1)
    shape = (8, 8, 8)
    target = rs.randint(0, 4, (1,) + shape)
    calls = []

    def fake_volume(model, x, mask, patch):
        calls.append(mask.tolist())
        return torch.from_numpy(rs.randint(0, 4, (1,) + shape)), None
    monkeypatch.setattr(predict, "predict_volume", fake_volume)
    res = metrics.evaluate_all_masks(object(), torch.zeros(1), torch.from_numpy(target))
    assert len(calls) == 15 and calls[0] == [predict.MASKS_TEST[0]] and len(res) == 15


def test_eval_report_layout(tmp_path):
    """eval.py writes the reference's CSV (header, masks in reversed table order, one row per case) and averages like
    the reference's nested AverageMeters (mean over the cases per mask, then mean over the masks)."""
    import csv
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("pb_eval", os.path.join(os.path.dirname(os.path.dirname(__file__)), "eval.py"))
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    rs = np.random.RandomState(0)
    names = ["caseA", "caseB", "caseC"]
    scores = {m: rs.rand(3, 4).astype(np.float32) for m in ev.MASK_NAME}
    path = str(tmp_path / "report.csv")
    per_mask, overall = ev.write_report(path, names, scores)
    rows = list(csv.reader(open(path)))
    assert rows[0][:4] == ["WT Dice", "TC Dice", "ET Dice", "ETPro Dice"] and len(rows) == 1 + 15 * 4
    assert rows[1] == ["flairt1cet1t2"] and rows[5] == ["t1cet1t2"] and rows[-4] == ["t2"]
    assert np.allclose([float(v) for v in rows[2][:4]], scores["flairt1cet1t2"][0]) and rows[2][4:] == [""] * 4
    meter = metrics.AverageMeter()
    for m in ev.MASK_NAME[::-1]:
        inner = metrics.AverageMeter()
        for k in range(3):
            inner.update(scores[m][k].astype(np.float64))
        assert np.allclose(inner.avg, per_mask[m])
        meter.update(inner.avg)
    assert np.allclose(meter.avg, overall)
    args = ev.args_parser(["--model", "rfnet_passion", "--synthetic", "2"])
    cases = list(ev.test_cases(args))
    assert len(cases) == 2 and cases[0][1].shape == (1, 4, 96, 96, 88) and cases[0][2].dtype == np.uint8
