#!/usr/bin/env python
"""Inference sweep of BASELINE.json configs[4]: all 15 missing-modality subsets, sliding window over a BraTS-sized
240x240x155 volume (reference code/utils/predict.py:144-218 + code/train.py:589-604), 1 GPU.

    python scripts/bench_infer.py [--patch 128] [--reps 3] [--cpu-windows 1]

Prints one JSON line: volumes/s for (a) the one-engine sweep `predict_all_masks` (encoders once per window, the fused
decoder on a batch of 15 masked copies), (b) 15 separate `predict_volume` calls on the same kernels (the reference's
loop structure), and (c) the CPU port (oracle/, fp32, all host threads) timed on a bounded number of windows and scaled.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passion_b200 import ops                                     # noqa: E402
from passion_b200.models import rfnet                            # noqa: E402
from passion_b200.predict import MASKS_TEST, _windows, predict_all_masks, predict_volume   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-windows", type=int, default=1)
    ap.add_argument("--shape", type=int, nargs=3, default=(240, 240, 155))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(1037)
    model = rfnet.Model(num_cls=4).to(dev)
    model.compute_dtype = torch.bfloat16
    rs = np.random.RandomState(7)
    x = torch.from_numpy(rs.standard_normal((1, 4) + tuple(args.shape)).astype(np.float32)).to(dev)
    nwin = len(_windows(tuple(args.shape), args.patch))

    def timed(fn):
        fn()                                                    # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps, out

    ms_sweep, (lab_s, _) = timed(lambda: predict_all_masks(model, x, patch_size=args.patch))

    def loop15():
        labs = []
        for m in MASKS_TEST:
            lab, _ = predict_volume(model, x, torch.tensor([m], device=dev), patch_size=args.patch)
            labs.append(lab)
        return torch.cat(labs), None
    ms_loop, (lab_l, _) = timed(loop15)
    agree = float((lab_s == lab_l).float().mean())
    ops.check_tc_errors()

    # CPU port on a bounded sample: `cpu_windows` windows x 15 masks, scaled to the whole volume
    from oracle import rfnet_oracle, synth
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_state_dict(1037)
    xc = x[:, :, :args.patch, :args.patch, :args.patch].float().cpu()
    with torch.no_grad():
        rfnet_oracle.forward(sd, xc[:, :, :32, :32, :32], torch.tensor([MASKS_TEST[-1]]), is_training=False)     # library warm-up
        t0 = time.time()
        n_cpu = 0
        for _ in range(args.cpu_windows):
            for m in MASKS_TEST[:3]:                            # 3 of the 15 masks per window keep the sample bounded
                rfnet_oracle.forward(sd, xc, torch.tensor([m]), is_training=False)
                n_cpu += 1
        dt = time.time() - t0
    cpu_s_per_volume = dt / n_cpu * nwin * 15
    out = {"metric": "15-mask inference sweep, volumes/s (240x240x155, sliding window %d^3, 50%% overlap)" % args.patch,
           "windows_per_volume": nwin,
           "sweep_engine": {"ms_per_volume": round(ms_sweep, 1), "volumes_per_s": round(1e3 / ms_sweep, 3)},
           "loop_of_15": {"ms_per_volume": round(ms_loop, 1), "volumes_per_s": round(1e3 / ms_loop, 3)},
           "label_agreement_sweep_vs_loop": agree,
           "cpu_port": {"s_per_volume": round(cpu_s_per_volume, 1), "volumes_per_s": round(1.0 / cpu_s_per_volume, 6),
                        "cores": torch.get_num_threads(),
                        "sample": f"{n_cpu} window forwards of {args.patch}^3 (fp32, oracle/), scaled by {nwin} windows x 15 masks"},
           "dtype": "bf16", "data": "synthetic"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
