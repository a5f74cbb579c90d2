"""Oracle (TEST INFRASTRUCTURE): the bit-exact integer/bool pieces of the data path.

Follows (reference, /root/reference/code):
  data/datasets_nii.py:27-30      mask table (also train.py:42-45)      -> MASK_ARRAY
  data/datasets_nii.py:130-139    mask selection per mask_type          -> select_mask_id()
  data/datasets_nii.py:150-153    one-hot float64 [C,D,H,W]             -> one_hot()
  preprocessing/generate_imb_mr.py:18-283  imbalanced missing-rate table -> generate_imb_mr()
  utils/predict.py:181-218        sliding-window origins / argmax       -> window_origins(), sliding_window_argmax()
Pure numpy; checked against the reference's shipped CSV tables in tests/test_masks.py.
"""

import numpy as np

MASK_ARRAY = np.array([
    [False, False, False, True], [False, True, False, False], [False, False, True, False],
    [True, False, False, False], [False, True, False, True], [False, True, True, False],
    [True, False, True, False], [False, False, True, True], [True, False, False, True],
    [True, True, False, False], [True, True, True, False], [True, False, True, True],
    [True, True, False, True], [False, True, True, True], [True, True, True, True]])
MASK_NAMES = ['t2', 't1c', 't1', 'flair', 't1cet2', 't1cet1', 'flairt1', 't1t2', 'flairt2', 'flairt1ce',
              'flairt1cet1', 'flairt1t2', 'flairt1cet2', 't1cet1t2', 'flairt1cet1t2']   # train.py:47-50


def mask_id_of(pattern):
    """Index of a [flair,t1ce,t1,t2] bool pattern in MASK_ARRAY."""
    hit = np.where((MASK_ARRAY == np.asarray(pattern, bool)).all(1))[0]
    return int(hit[0])


def possible_mask_ids(pattern):
    """All table rows whose present set is a subset of `pattern` (the CSV's pos_mask_ids)."""
    pat = np.asarray(pattern, bool)
    return [i for i in range(15) if not (MASK_ARRAY[i] & ~pat).any()]


def select_mask_id(mask_type, csv_mask_id=None, pos_mask_ids=None, rng=np.random):
    """datasets_nii.py:134-139."""
    if mask_type == 'idt':
        return np.array([csv_mask_id])
    if mask_type == 'idt_drop':
        return rng.choice(pos_mask_ids, 1)
    if mask_type == 'pdt':
        return rng.choice(15, 1)
    raise ValueError(mask_type)


def one_hot(labels, num_cls=4):
    """datasets_nii.py:148-153: int labels [1,H,W,Z] -> float64 one-hot [num_cls,H,W,Z]."""
    _, H, W, Z = labels.shape
    yo = np.eye(num_cls)[labels.reshape(-1)].reshape(1, H, W, Z, -1)
    return np.ascontiguousarray(yo.transpose(0, 4, 1, 2, 3))[0]


def generate_imb_mr(names, p=(0.2, 0.4, 0.6, 0.8), seed=1037):
    """generate_imb_mr.py restated.  `names` = sorted training ids.  p = missing rates of
    (t1, t1c, flair, t2).  Returns rows (name, mask_id, [flair,t1c,t1,t2], pos_mask_ids)."""
    rs = np.random.RandomState(seed)
    n = len(names)
    cols = [rs.rand(n) > p[j] for j in range(4)]             # t1, t1c, flair, t2 (:41-44); overwritten below
    count = 0
    # block order of the reference (:50-170), as (t1, t1c, flair, t2):
    order = ["TTTT", "TTFT", "TTTF", "TTFF", "TFTT", "TFTF", "TFFT", "TFFF",
             "FTTT", "FTFT", "FTTF", "FTFF", "FFTT", "FFTF", "FFFT"]
    for pat in order:
        v = n
        for j, ch in enumerate(pat):
            v = v * ((1 - p[j]) if ch == "T" else p[j])
        k = int(v)
        k = k if k > 0 else k + 1
        for j, ch in enumerate(pat):
            cols[j][count:count + k] = (ch == "T")
        count += k
    for j in range(4):
        cols[j][count:] = False
    state = rs.get_state()                                   # :188-195: same permutation for all four
    for j in range(4):
        rs.set_state(state)
        rs.shuffle(cols[j])
    t1, t1c, flair, t2 = cols
    rows = []
    for i in range(n):
        while not (t1[i] or t1c[i] or flair[i] or t2[i]):    # :212-218
            t1[i] = rs.rand(1) > p[0]
            t1c[i] = rs.rand(1) > p[1]
            flair[i] = rs.rand(1) > p[2]
            t2[i] = rs.rand(1) > p[3]
        pat = [bool(flair[i]), bool(t1c[i]), bool(t1[i]), bool(t2[i])]
        rows.append((names[i], mask_id_of(pat), pat, possible_mask_ids(pat)))
    return rows


def window_origins(size, patch, overlap=0.5):
    """predict.py:181-195 for one axis."""
    step = int(patch * (1 - overlap))
    cnt = int(np.ceil((size - patch) / (patch * (1 - overlap))))
    return [i * step for i in range(cnt)] + [size - patch]


def sliding_window_argmax(prob_fn, x, patch):
    """predict.py:197-218.  prob_fn(x_window [B,4,p,p,p]) -> probs [B,C,p,p,p] (numpy).
    Returns int64 labels [B,H,W,Z] (first-max tie-break, like torch.argmax)."""
    B, _, H, W, Z = x.shape
    hs, ws, zs = window_origins(H, patch), window_origins(W, patch), window_origins(Z, patch)
    weight = np.zeros((1, 1, H, W, Z), np.float32)
    pred = None
    for h in hs:
        for w in ws:
            for z in zs:
                weight[:, :, h:h + patch, w:w + patch, z:z + patch] += 1.0
    for h in hs:
        for w in ws:
            for z in zs:
                part = prob_fn(x[:, :, h:h + patch, w:w + patch, z:z + patch])
                if pred is None:
                    pred = np.zeros((B, part.shape[1], H, W, Z), np.float32)
                pred[:, :, h:h + patch, w:w + patch, z:z + patch] += part
    pred = pred / weight
    return pred.argmax(1)
