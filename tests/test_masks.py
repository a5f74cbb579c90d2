"""CPU, bit-exact: mask table, imbalanced-missing-rate generator (known-answer = the reference's shipped
CSV), one-hot labels, sliding-window origins."""
import csv
import os

import numpy as np

from oracle import masks

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _read_csv():
    with open(os.path.join(GOLD, "Brats2020_imb_split_mr2468.csv")) as f:
        return list(csv.DictReader(f))


def test_generate_imb_mr_reproduces_shipped_csv():
    with open(os.path.join(GOLD, "brats2020_train.txt")) as f:
        names = sorted(l.strip() for l in f if l.strip())
    assert len(names) == 219
    rows = masks.generate_imb_mr(names, p=(0.2, 0.4, 0.6, 0.8), seed=1037)
    gold = _read_csv()
    assert len(rows) == len(gold) == 219
    for (name, mid, pat, pos), g in zip(rows, gold):
        assert name == g["data_name"]
        assert mid == int(g["mask_id"])
        assert pat == eval(g["mask"].replace("np.True_", "True").replace("np.False_", "False"))
        assert pos == eval(g["pos_mask_ids"])
    counts = np.sum([r[2] for r in rows], 0)           # flair, t1ce, t1, t2
    assert counts.tolist() == [90, 135, 184, 43]        # generate_imb_mr.py:187 "2468: 184 135 90 43" (t1 t1c flair t2)


def test_mask_table_and_ids():
    assert masks.MASK_ARRAY.shape == (15, 4) and len(masks.MASK_NAMES) == 15
    assert len({tuple(r) for r in masks.MASK_ARRAY.tolist()}) == 15
    for g in _read_csv():
        pat = eval(g["mask"])
        assert masks.MASK_ARRAY[int(g["mask_id"])].tolist() == pat
        assert masks.possible_mask_ids(pat) == eval(g["pos_mask_ids"])
    assert masks.select_mask_id("idt", csv_mask_id=7).tolist() == [7]
    rs = np.random.RandomState(0)
    assert masks.select_mask_id("idt_drop", pos_mask_ids=[1, 2, 5], rng=rs)[0] in (1, 2, 5)
    assert 0 <= masks.select_mask_id("pdt", rng=rs)[0] < 15


def test_one_hot_bit_exact():
    rs = np.random.RandomState(1)
    y = rs.randint(0, 4, (1, 5, 6, 7))
    oh = masks.one_hot(y)
    assert oh.dtype == np.float64 and oh.shape == (4, 5, 6, 7)
    assert np.array_equal(oh.argmax(0), y[0]) and np.array_equal(oh.sum(0), np.ones((5, 6, 7)))


def test_window_origins():
    assert masks.window_origins(240, 80) == [0, 40, 80, 120, 160]
    assert masks.window_origins(155, 80) == [0, 40, 75]
    assert masks.window_origins(240, 128) == [0, 64, 112]
    assert masks.window_origins(155, 128) == [0, 27]
    assert masks.window_origins(80, 80) == [0]


def test_sliding_window_argmax_first_max_tiebreak():
    x = np.zeros((1, 4, 20, 20, 12), np.float32)

    def prob_fn(win):
        return np.full((1, 3) + win.shape[2:], 1.0 / 3, np.float32)
    lab = masks.sliding_window_argmax(prob_fn, x, 8)
    assert lab.shape == (1, 20, 20, 12) and (lab == 0).all()
