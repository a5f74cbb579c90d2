"""Oracle (TEST INFRASTRUCTURE): golden fixture AT THE SCALE THE BENCHMARK RUNS (round-2 item: parity was only pinned at
16^3-32^3, where the coarsest InstanceNorms see 8-27 voxels).  Runs ONLY in the build container (needs /root/reference).

Case `idt64`: RFNet + PASSION, B = 2, 4x64^3 crops (coarsest level 8^3 = 512 voxels per InstanceNorm), the first two masks
bench.py draws from the mr2468 table, temp 4, labels 'U'.  80^3 (the bench crop) is not used for the gradient fixture only
because the float64 "exact gradient" runs of a B = 2 batch need ~75 GB of host memory at 80^3; bench.py's `parity` block
covers the forward quantities at 80^3 on the box itself.

Written to tests/golden/rfnet_passion_idt64.npz:
  * the UNMODIFIED reference's outputs (fuse_prob sub-sampled ::4, argmax sub-sampled ::2, the five per-sample loss tensors,
    step loss, rp_iter) and gradient summaries (per-parameter L2 norm + fixed random projection), fp32 CPU;
  * from the float64 oracle: the same summaries ("exact"), the per-parameter and global SENSITIVITY of the gradient to a
    2e-6 relative input perturbation (2 probes) and the fp32-oracle-vs-float64 error — the calibration the GPU test uses;
  * the bf16-storage noise floor of the network at this size (oracle with activations rounded to bf16 on the CPU).
Usage: python -m oracle.gen_golden_full      (≈ 10 minutes on 8 cores, ≈ 40 GB of host memory)
"""
import os
import sys
import time

import numpy as np
import torch

from .gen_golden import GOLD, grad_summary, import_reference, reference_mix

CASE = dict(name="idt64", B=2, S=64, labels="U", temp=4.0, seed=1037)


def main():
    sys.path.insert(0, os.path.dirname(GOLD.rstrip("/")).rsplit("/tests", 1)[0])
    import bench
    from oracle import rfnet_oracle, synth, train_step_oracle
    ref_rfnet, ref_crit = import_reference()
    B, S, temp = CASE["B"], CASE["S"], CASE["temp"]
    mask_ids = bench.mask_ids_for(B)
    x, target, mask, _ = synth.make_batch(B, S, seed=CASE["seed"], labels=CASE["labels"], mask_ids=mask_ids)
    sd = synth.make_state_dict(1037)
    imb_beta = torch.tensor([1.1, 0.9, 1.3, 0.7])
    modal_weight = bench.modal_weight()

    t0 = time.time()
    ref_model = ref_rfnet.Model(num_cls=4)
    ref_model.load_state_dict(sd)
    ref_model.is_training, ref_model.use_passion, ref_model.mask_type = True, True, "idt"
    outs = ref_model(x, mask, target=target, temp=temp)
    loss, rp_iter = reference_mix(ref_crit, outs, target, mask, imb_beta, modal_weight, "idt")
    loss.backward()
    ref_g = {k: p.grad.detach().clone() for k, p in ref_model.named_parameters() if p.grad is not None}
    print(f"reference fp32: loss {float(loss):.6f}  rp_iter {rp_iter.tolist()}  ({time.time() - t0:.0f} s)", flush=True)
    names_out = ["fuse_prob", "prm_loss", "sep_loss", "kl_loss", "proto_loss", "dist"]
    save = {"B": B, "S": S, "temp": temp, "seed": CASE["seed"], "mask": mask.numpy(), "mask_ids": np.array(mask_ids),
            "labels_kind": CASE["labels"], "mask_type": "idt", "use_passion": True, "imb_beta": imb_beta.numpy(),
            "modal_weight": modal_weight.numpy(), "loss": float(loss), "rp_iter": rp_iter.detach().numpy()}
    for n, r in zip(names_out, outs):
        save[n] = r.detach().numpy()
    save["fuse_argmax_s2"] = outs[0].argmax(1)[:, ::2, ::2, ::2].numpy().astype(np.int8)
    top2 = torch.topk(outs[0].detach(), 2, dim=1).values
    save["fuse_gap_s2"] = (top2[:, 0] - top2[:, 1])[:, ::2, ::2, ::2].numpy().astype(np.float16)
    save["fuse_prob"] = save["fuse_prob"][:, :, ::4, ::4, ::4].copy()
    names, norms, projs = grad_summary(ref_g)
    save["grad_names"], save["grad_norms"], save["grad_projs"] = np.array(names), norms, projs
    del ref_model, outs, loss

    def oracle_run(xin, dtype):
        P = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
        o = rfnet_oracle.forward(P, xin.to(dtype), mask, target, temp)
        l, _ = train_step_oracle.loss_mix(o, target, mask, imb_beta.to(dtype), modal_weight.to(dtype))
        l.backward()
        return [t.detach() for t in o], float(l), {k: p.grad for k, p in P.items()}

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    t0 = time.time()
    o32, l32, g32 = oracle_run(x, torch.float32)
    worst = max(rel(g32[k], ref_g[k]) for k in names if float(ref_g[k].norm()) > 1e-5)
    print(f"oracle fp32: loss {l32:.6f}; worst per-tensor gradient deviation from the reference {worst:.2e} ({time.time() - t0:.0f} s)", flush=True)
    assert abs(l32 - save["loss"]) < 1e-4 * abs(save["loss"]) and worst < 5e-3
    t0 = time.time()
    o64, l64, g64 = oracle_run(x, torch.float64)
    print(f"oracle fp64: loss {l64:.8f} ({time.time() - t0:.0f} s)", flush=True)
    keys = [k for k in names]
    probes = []
    for seed in range(2):
        g = torch.Generator().manual_seed(seed)
        xp = x.double() * (1 + 2e-6 * torch.randn(x.shape, generator=g, dtype=torch.float64))
        probes.append(oracle_run(xp, torch.float64)[2])
        print(f"probe {seed} done", flush=True)
    cat = lambda d: torch.cat([d[k].flatten().double() for k in keys if not k.endswith(".conv.bias")])
    save["x_grad_norms"] = np.array([float(g64[k].norm()) for k in keys])
    _, _, save["x_grad_projs"] = grad_summary({k: g64[k] for k in keys})
    save["sens"] = np.array([max(rel(p[k], g64[k]) for p in probes) for k in keys])
    save["sens_global"] = max(rel(cat(p), cat(g64)) for p in probes)
    save["fp32_oracle_err"] = np.array([rel(g32[k], g64[k]) for k in keys])
    save["fp32_oracle_err_global"] = rel(cat(g32), cat(g64))
    print(f"float64 sensitivity to 2e-6 input noise: global {save['sens_global']:.2e}, worst tensor {save['sens'].max():.2e}; "
          f"fp32 oracle vs float64: global {save['fp32_oracle_err_global']:.2e}", flush=True)
    del probes, g64, o64

    # bf16-storage noise floor (scripts/bf16_sim.py's experiment at this size): activations + raw conv outputs rounded to bf16
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD.rstrip("/")).rsplit("/tests", 1)[0], "scripts"))
    import bf16_sim
    orig = rfnet_oracle.conv_in_lrelu
    rfnet_oracle.conv_in_lrelu = bf16_sim.patched
    bf16_sim.MODE.update(dict(y=True, a=True))
    try:
        ob, lb, gb = oracle_run(x, torch.float32)
    finally:
        rfnet_oracle.conv_in_lrelu = orig
    save["bf16_sim_prob_rel"] = rel(ob[0], o32[0])
    save["bf16_sim_grad_rel"] = rel(cat(gb), cat(g32))
    save["bf16_sim_loss_rel"] = abs(lb - l32) / abs(l32)
    save["bf16_sim_out_rel"] = np.array([rel(a, b) for a, b in zip(ob, o32)])
    print(f"bf16-storage simulation on the CPU oracle: prob rel-L2 {save['bf16_sim_prob_rel']:.2e}, gradient rel-L2 "
          f"{save['bf16_sim_grad_rel']:.2e}, loss {save['bf16_sim_loss_rel']:.2e}", flush=True)
    np.savez_compressed(os.path.join(GOLD, f"rfnet_passion_{CASE['name']}.npz"), **save)
    print("written", os.path.join(GOLD, f"rfnet_passion_{CASE['name']}.npz"))


if __name__ == "__main__":
    main()
