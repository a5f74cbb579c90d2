"""Developer probe: compare gradients w.r.t. intermediate activations (CUDA fp32 path vs fp64 oracle)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_model_gpu as T    # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    from oracle import rfnet_oracle, train_step_oracle
    case = os.environ.get("CASE", "idtS24")
    z, model, sd, x, target, mask = T._setup(case, torch.float32)
    model.use_passion = True
    # ---- CUDA
    dev = "cuda"
    from passion_b200.train_step import loss_mix
    outs = model(x.to(dev), mask.to(dev), target=target.to(dev), temp=float(z["temp"]))
    last = model.last
    keep = {"fuse_logits": last["fuse_logits"]}
    for l in range(4):
        keep[f"prm{l+1}"] = last["prm_logits"][l]
        keep[f"de{l+1}"] = last["de_f"][l]
    keep["sep_prob"] = torch.softmax(last["sep_logits"].float(), -1)
    for i, e in enumerate(last["enc"]):
        keep[f"enc{i+1}"] = e
    for t in keep.values():
        t.retain_grad()
    loss, _ = loss_mix(outs, target.to(dev), mask.to(dev), torch.from_numpy(z["imb_beta"]).to(dev),
                       torch.from_numpy(z["modal_weight"]).to(dev), mask_type=str(z["mask_type"]))
    loss.backward()
    # ---- oracle fp64
    dt = torch.float64
    P = {k: v.clone().to(dt).requires_grad_(True) for k, v in sd.items()}
    o_outs, it = rfnet_oracle.forward(P, x.to(dt), mask, target, float(z["temp"]), mask_type=str(z["mask_type"]), return_internals=True)
    okeep = {"fuse_logits": it["fuse_logits"]}
    for l in range(4):
        okeep[f"prm{l+1}"] = it["prm_logits"][l]
        okeep[f"de{l+1}"] = it["de_f"][l]
    for m in range(4):
        okeep[f"mod{m}_fuse_logits"] = it["mod"][m][0]
        okeep[f"mod{m}_de1"] = it["mod"][m][2][0]
    for lvl in range(4):
        for m in range(4):
            okeep[f"enc{lvl+1}_m{m}"] = it["enc"][m][lvl]
    for t in okeep.values():
        t.retain_grad()
    o_loss, _ = train_step_oracle.loss_mix(o_outs, target, mask, torch.from_numpy(z["imb_beta"]).to(dt),
                                           torch.from_numpy(z["modal_weight"]).to(dt), mask_type=str(z["mask_type"]))
    o_loss.backward()
    B = x.shape[0]
    Pn = last["passes"]
    print(f"case {case}: loss {float(loss):.7f} vs {float(o_loss):.7f}")

    def cl2nc(t):
        return t.permute(0, 4, 1, 2, 3)
    g = keep["fuse_logits"].grad          # [P,B,D,H,W,C]
    print("d fuse_logits pass0:", rel(cl2nc(g[0]), okeep["fuse_logits"].grad))
    for m in range(4):
        print(f"d fuse_logits mod{m}:", rel(cl2nc(g[1 + m]), okeep[f"mod{m}_fuse_logits"].grad) if okeep[f"mod{m}_fuse_logits"].grad is not None else None)
    for l in range(4):
        gp = keep[f"prm{l+1}"].grad
        gp = gp.view(Pn, B, *gp.shape[1:])
        print(f"d prm{l+1} pass0:", rel(cl2nc(gp[0]), okeep[f"prm{l+1}"].grad))
        gd = keep[f"de{l+1}"].grad
        gd = gd.view(Pn, B, *gd.shape[1:])
        print(f"d de{l+1} pass0:", rel(cl2nc(gd[0]), okeep[f"de{l+1}"].grad))
    gd = keep["de1"].grad.view(Pn, B, *keep["de1"].shape[1:])
    for m in range(4):
        og = okeep[f"mod{m}_de1"].grad
        print(f"d de1 mod{m}:", rel(cl2nc(gd[1 + m]), og) if og is not None and float(og.norm()) > 0 else "zero")
    for lvl in range(4):
        ge = keep[f"enc{lvl+1}"].grad     # [4B,...] modality-major
        ge = ge.view(4, B, *ge.shape[1:])
        for m in range(4):
            og = okeep[f"enc{lvl+1}_m{m}"].grad
            if og is None or float(og.norm()) == 0:
                continue
            print(f"d enc{lvl+1} m{m}: rel {rel(cl2nc(ge[m]), og):.3e}  |g| {float(og.norm()):.3e}")


if __name__ == "__main__":
    main()
