"""Study behind DESIGN.md §2 "Round-off can move a single encoder's gradient by percents" (not a test; CPU only, ~2 min):

    python tests/study_grad_discontinuity.py

Runs the mmFormer+PASSION oracle of fixture `idtU` in float64, in float32, in float32 on inputs perturbed by 1e-7 (a stand-in
for another fp32 summation order) and in float64 on inputs perturbed by 2e-6 (the probes of test_fp32_check_mode), and prints
the rel-L2 error of a few weight gradients against the float64 run.  Output of 2026-10-17 (torch 2.11 CPU, 8 threads):

    columns: t1_encoder.e1_c1.weight  t1_encoder.e1_c2.conv.weight  t1_encoder.e5_c3.conv.weight  t1ce_encoder.e1_c2.conv.weight
             flair_encoder.e1_c2.conv.weight  decoder_fuse.d1_c2.conv.weight
    fp32 oracle           1.21e-03 9.84e-04 9.74e-04 4.13e-03 7.62e-04 1.52e-04
    fp32 + 1e-7 input #0  1.21e-03 1.00e-03 1.03e-03 8.58e-04 7.52e-04 5.76e-05
    fp32 + 1e-7 input #1  8.00e-04 7.08e-04 7.24e-04 6.38e-04 5.84e-04 1.49e-04
    fp32 + 1e-7 input #2  8.73e-04 7.49e-04 8.27e-04 4.10e-03 7.41e-04 1.48e-04
    fp32 + 1e-7 input #3  1.48e-03 1.12e-03 9.91e-04 4.03e-03 5.82e-04 1.54e-04
    fp32 + 1e-7 input #4  9.16e-04 8.03e-04 8.26e-04 4.02e-03 7.75e-04 7.14e-05
    fp32 + 1e-7 input #5  9.79e-04 8.00e-04 7.24e-04 3.97e-03 5.66e-04 4.62e-05
    fp64 + 2e-6 input #0  2.37e-03 2.02e-03 1.95e-03 5.87e-03 1.62e-03 1.99e-04
    fp64 + 2e-6 input #1  2.85e-03 2.44e-03 2.50e-03 2.33e-03 3.64e-03 4.94e-04
    fp64 + 2e-6 input #2  2.65e-03 2.29e-03 1.83e-03 2.20e-02 2.61e-03 1.60e-04

The t1ce column is bimodal (6-9e-4 or 4e-3) under perturbations of 1e-7 and jumps to 2.2e-2 in one float64 probe: discrete events
(a LeakyReLU / clamp branch flipping behind an InstanceNorm over a handful of voxels) that reach one modality encoder only.
On the B200 the same kind of event appeared when the fp32 up-sampling forward got another summation order: 8 % on t1_encoder,
everything else within bounds; gone again with the old order (csrc/upsample.cu keeps the point-wise kernel for fp32).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import test_mmformer_gpu as T
    from oracle import synth
    from oracle.masks import mask_id_of
    z = np.load(os.path.join(T.GOLD, "mmformer_passion_idtU.npz"), allow_pickle=True)
    B, S = int(z["B"]), int(z["S"])
    x, target, mask, _ = synth.make_batch(B, S, seed=int(z["seed"]), labels=str(z["labels_kind"]),
                                          mask_ids=[mask_id_of(m) for m in z["mask"]])
    sd = synth.make_state_dict(2051, synth.mmformer_param_shapes(patch=2))

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    g64 = T._oracle(sd, x, target, mask, z, torch.float64)[2]
    keys = ["t1_encoder.e1_c1.weight", "t1_encoder.e1_c2.conv.weight", "t1_encoder.e5_c3.conv.weight",
            "t1ce_encoder.e1_c2.conv.weight", "flair_encoder.e1_c2.conv.weight", "decoder_fuse.d1_c2.conv.weight"]
    print("columns:", "  ".join(keys))

    def row(g):
        return " ".join("%.2e" % rel(g[k], g64[k]) for k in keys)
    print("fp32 oracle          ", row(T._oracle(sd, x, target, mask, z, torch.float32)[2]))
    for seed in range(6):
        g = torch.Generator().manual_seed(100 + seed)
        xp = (x.double() * (1 + 1e-7 * torch.randn(x.shape, generator=g, dtype=torch.float64))).float()
        print("fp32 + 1e-7 input #%d " % seed, row(T._oracle(sd, xp, target, mask, z, torch.float32)[2]))
    for seed in range(3):
        g = torch.Generator().manual_seed(seed)
        xp = x.double() * (1 + 2e-6 * torch.randn(x.shape, generator=g, dtype=torch.float64))
        print("fp64 + 2e-6 input #%d " % seed, row(T._oracle(sd, xp, target, mask, z, torch.float64)[2]))


if __name__ == "__main__":
    main()
