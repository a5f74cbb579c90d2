"""CPU baseline: the UNMODIFIED reference model and criterions (baseline/_ref/code, staged by install_ref.py; in the build
container /root/reference/code directly) driven through one training step exactly as the reference training loop does.

What is the reference's own code here: models.rfnet.Model (construction, Kaiming init, forward incl. the in-forward PASSION
losses), utils.criterions.* (CE / Dice of the fused prediction).  What is restated, because code/train.py cannot be imported
(it parses argv and imports nibabel / medpy at module level): the ~20 lines of per-step loss mix, train.py:228-229 and
:258-280 ('idt'), and the optimizer construction, train.py:94-96.  `.cuda()` is shimmed to the identity (criterions.py:153
hard-codes it) — the one-line shim SURVEY.md §8(c) describes.

    python baseline/run_cpu_reference.py [--batch 1] [--size 80] [--steps 3]
"""
import argparse
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_code_dir():
    for cand in (os.path.join(HERE, "_ref", "code"), "/root/reference/code"):
        if os.path.exists(os.path.join(cand, "models", "rfnet.py")):
            return cand
    return None


def available():
    return reference_code_dir() is not None


def import_reference():
    import torch
    code = reference_code_dir()
    if code is None:
        raise RuntimeError("reference modules not staged: run `python baseline/install_ref.py` where /root/reference exists")
    if code not in sys.path:
        sys.path.insert(0, code)
    torch.Tensor.cuda = lambda self, *a, **k: self          # CPU shim (criterions.py:153)
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.") or m == "utils" or m.startswith("utils.")]:
        mod = sys.modules[name]
        if not (getattr(mod, "__file__", None) or "").startswith(code):
            del sys.modules[name]                            # a same-named package from elsewhere must not shadow it
    from models import rfnet as ref_rfnet
    from utils import criterions as ref_crit
    return ref_rfnet, ref_crit, code


def step_fn(x, target, mask, imb_beta, modal_weight, temp=4.0, seed=1037, state_dict=None, threads=None):
    """-> (step, model): step() runs forward + loss mix + backward + AdamW(amsgrad) once and returns (loss, outs)."""
    import torch
    ref_rfnet, ref_crit, _ = import_reference()
    torch.set_num_threads(threads or os.cpu_count())
    torch.manual_seed(seed)
    model = ref_rfnet.Model(num_cls=4)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    model.is_training, model.use_passion, model.mask_type = True, True, "idt"        # train.py:91-92,212
    model.train()
    params = [{"params": model.parameters(), "lr": 2e-4, "weight_decay": 1e-4}]      # train.py:94-96
    opt = torch.optim.AdamW(params, betas=(0.9, 0.999), eps=1e-08, amsgrad=True)
    fm = mask

    def step():
        outs = model(x, mask, target=target, temp=temp)                               # train.py:222
        fuse_pred, prm_bs, sep_bs, kl_bs, proto_bs, dist_bs = outs
        fuse = (ref_crit.softmax_weighted_loss_bs(fuse_pred, target, num_cls=4)
                + ref_crit.dice_loss_bs(fuse_pred, target, num_cls=4)).sum()          # :228-229
        prm = prm_bs.sum()
        sep_m, kl_m, proto_m = (sep_bs * fm).sum(0), (kl_bs * fm).sum(0), (proto_bs * fm).sum(0)     # :258-263
        rp_iter = torch.zeros(4)
        for bs in range(fuse_pred.size(0)):                                           # :265-267
            rp_iter += fm[bs] * (dist_bs[bs] / (sum(dist_bs[bs]) / sum(fm[bs])) - 1)
        rp_mask = rp_iter > 0                                                          # :268
        kl = (imb_beta * modal_weight * kl_m).sum()
        proto = (rp_mask * modal_weight * proto_m).sum()
        sep = (rp_mask * imb_beta * modal_weight * sep_m).sum()
        loss = fuse + sep + prm + kl * 0.5 + proto * 0.1                               # :280
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss), outs
    return step, model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--size", type=int, default=80)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    import torch
    import bench
    from oracle import synth
    x, target, mask, _ = synth.make_batch(args.batch, args.size, seed=1037, labels="U", mask_ids=bench.mask_ids_for(args.batch))
    step, _ = step_fn(x, target, mask, torch.ones(4), bench.modal_weight())
    t0 = time.time(); l0, _ = step(); t_first = time.time() - t0
    ts = []
    for _ in range(args.steps):
        t0 = time.time(); loss, _ = step(); ts.append(time.time() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"reference CPU step: B={args.batch} 4x{args.size}^3 fp32, {torch.get_num_threads()} threads of {os.cpu_count()} cores: "
          f"first {t_first:.2f} s, median {med:.2f} s/step = {args.batch / med:.4f} samples/s; loss {l0:.6f} -> {loss:.6f}")


if __name__ == "__main__":
    main()
