#!/usr/bin/env python
"""Evaluation entry point with the reference's flags and report format (code/eval.py:29-183), on the B200-native path.

    python eval.py --model rfnet --resume outputs/.../model_last.pth --savepath outputs/eval
    python eval.py --model rfnet --synthetic 2 --savepath /tmp/eval            # no dataset / checkpoint needed

The reference walks the 15 missing-modality masks in reverse table order and, for each of them, loops over the whole
test set with a sliding window (eval.py:159-175 -> utils/predict.py:144-258): 15 passes over every case.  Here every case
is read once and all 15 masks are swept together (passion_b200.predict.predict_all_masks: encoders once per window, the
fused decoder on a batch of 15 masked copies); the Dice scores are the reference's own expressions on exact confusion
counts (passion_b200.metrics.dice_class4, bit-exact against predict.py:82-128).  The CSV keeps the reference's layout —
header, then per mask (reversed order) its name and one row per case — with the four HD95 columns left empty: the
Hausdorff distance comes from medpy on the CPU in the reference (predict.py:22-80) and is outside this hot path.
"""
import argparse
import csv
import logging
import os

import numpy as np
import torch

from passion_b200 import metrics, ops
from passion_b200.models import build_model
from passion_b200.predict import MASKS_TEST

MASK_NAME = ['t2', 't1c', 't1', 'flair', 't1cet2', 't1cet1', 'flairt1', 't1t2', 'flairt2', 'flairt1ce',
             'flairt1cet1', 'flairt1t2', 'flairt1cet2', 't1cet1t2', 'flairt1cet1t2']                 # eval.py:76-80
HEADER = ['WT Dice', 'TC Dice', 'ET Dice', 'ETPro Dice', 'WT HD95', 'TC HD95', 'ET HD95' 'ETPro HD95']   # eval.py:157 (sic)


def args_parser(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument('--model', default='rfnet', type=str, help='rfnet | mmformer (also accepted: rfnet_passion, mmformer_passion)')
    p.add_argument('-batch_size', '--batch_size', default=1, type=int)
    p.add_argument('--dataname', default='BraTS/BRATS2020', type=str)
    p.add_argument('--datapath', default='BraTS/BRATS2020_Training_none_npy', type=str)
    p.add_argument('--savepath', default='outputs/eval', type=str)
    p.add_argument('--resume', default=None, type=str)
    p.add_argument('--mask_type', default='idt', type=str)
    p.add_argument('--seed', default=1037, type=int)
    # additions (not in the reference)
    p.add_argument('--datarootPath', default=None, type=str, help='dataset root (default: ./datasets)')
    p.add_argument('--synthetic', default=0, type=int, help='evaluate N synthetic cases instead of <datapath>/test.txt')
    p.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    p.add_argument('--patch_size', default=80, type=int, help='sliding-window edge (utils/predict.py:20)')
    return p.parse_args(argv)


def test_cases(args):
    """Yields (name, x float32 [1,4,H,W,Z], y uint8 [1,H,W,Z]) like Brats_loadall_test_nii (datasets_nii.py:165-205)."""
    if args.synthetic:
        rs = np.random.RandomState(args.seed)
        for i in range(args.synthetic):
            shape = (96, 96, 88)
            yield f'synthetic_{i}', rs.standard_normal((1, 4) + shape).astype(np.float32), rs.randint(0, 4, (1,) + shape).astype(np.uint8)
        return
    root = args.datarootPath or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'datasets')
    data = os.path.abspath(os.path.join(root, args.datapath))
    test_file = os.path.join(data, 'test1.txt' if args.dataname == 'BraTS/BRATS2018' else 'test.txt')     # eval.py:131-138
    with open(test_file) as f:
        names = sorted(i.strip() for i in f.readlines())
    for name in names:
        x = np.load(os.path.join(data, 'vol', name + '_vol.npy'))                      # [H,W,Z,4]
        y = np.load(os.path.join(data, 'seg', name + '_seg.npy')).astype(np.uint8)
        yield name, np.ascontiguousarray(x.transpose(3, 0, 1, 2))[None].astype(np.float32), y[None]


def write_report(csv_name, names, scores):
    """scores: {mask name: float array [n cases, 4]}.  Writes the reference's CSV (eval.py:155-175, predict.py:236-241) and
    returns ({mask name: mean over the cases}, mean over the masks of those means) — AverageMeter semantics."""
    per_mask = {}
    with open(csv_name, 'a+', newline='') as f:
        w = csv.writer(f)
        w.writerow(HEADER)
        for mname in MASK_NAME[::-1]:
            w.writerow([mname])
            for k in range(len(names)):
                w.writerow([float(v) for v in scores[mname][k]] + [''] * 4)
            per_mask[mname] = np.asarray(scores[mname], np.float64).mean(0)
    overall = np.mean([per_mask[m] for m in MASK_NAME[::-1]], 0)
    return per_mask, overall


def evaluate(model, cases, patch_size=80, device='cuda'):
    """-> (names, {mask name: [n cases, 4] Dice (whole, core, enhancing, enhancing_postpro)})."""
    names, scores = [], {m: [] for m in MASK_NAME}
    for name, x, y in cases:
        res = metrics.evaluate_all_masks(model, torch.from_numpy(x).to(device), torch.from_numpy(y).to(device),
                                         patch_size=patch_size, masks=MASKS_TEST, mask_names=MASK_NAME)
        ops.check_tc_errors()                          # once per case: a tcgen05 pipeline time-out must not pass silently
        names.append(name)
        msg = []
        for m in MASK_NAME[::-1]:
            s = res[m].float().cpu().numpy()
            scores[m].append(s)
            msg.append('%s %.4f/%.4f/%.4f' % (m, s[0], s[1], s[2]))
        logging.info('Subject %d %s  DSC whole/core/enhancing: %s', len(names), name, ', '.join(msg[:3]) + ', ...')
    return names, scores


def main(argv=None):
    args = args_parser(argv)
    if not torch.cuda.is_available():
        raise SystemExit('eval.py: no CUDA device — the B200-native path has no CPU fallback')
    os.makedirs(args.savepath, exist_ok=True)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s %(message)s')
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    name = args.model.replace('_passion', '')
    dev = torch.device('cuda', 0)
    model = build_model(name, num_cls=4, crop=args.patch_size).to(dev)
    model.compute_dtype = torch.float32 if args.dtype == 'f32' else torch.bfloat16
    model.mask_type = args.mask_type
    if args.resume is not None:                                                        # eval.py:151-153
        ck = torch.load(args.resume, map_location=dev)
        sd = {k[len('module.'):] if k.startswith('module.') else k: v for k, v in ck['state_dict'].items()}
        model.load_state_dict(sd)
        logging.info('last epoch: %d', ck.get('epoch', -1) + 1)
    model.is_training = False
    model.eval()                                                                       # reference utils/predict.py:154
    names, scores = evaluate(model, test_cases(args), args.patch_size, dev)
    csv_name = os.path.join(args.savepath, f'{args.model}.csv')
    per_mask, overall = write_report(csv_name, names, scores)
    for m in MASK_NAME[::-1]:
        logging.info('%s Average scores: DSC: %s', m, ', '.join('%s: %.4f' % kv for kv in zip(metrics.CLASS_EVALUATION, per_mask[m])))
    logging.info('Avg Dice scores: %s', overall)
    return overall


if __name__ == '__main__':
    main()
