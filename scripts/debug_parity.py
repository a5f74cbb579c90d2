"""Developer probe: per-tensor parity table of the CUDA path vs the CPU oracle (fp32 and bf16)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_model_gpu as T    # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    from oracle import rfnet_oracle
    case = os.environ.get("CASE", "idtS24")
    for dt in (torch.float32, torch.bfloat16):
        z, model, sd, x, target, mask = T._setup(case, dt)
        outs, loss, parts = T._cuda_step(model, x, target, mask, z)
        o_outs, o_loss, o_grads = T._oracle(sd, x, target, mask, z)
        print(f"==== {case} {dt}: loss {float(loss):.6f} oracle {float(o_loss):.6f}")
        for n, a, b in zip(["fuse_prob", "prm", "sep", "kl", "proto", "dist"], outs, o_outs):
            print(f"  out {n:10s} rel {rel(a, b):.3e}")
        # internals of the full-mask pass
        with torch.no_grad():
            _, internals = rfnet_oracle.forward(sd, x, mask, target, float(z["temp"]), mask_type=str(z["mask_type"]),
                                                return_internals=True)
        B = x.shape[0]
        last = model.last
        print(f"  fuse_logits pass0 rel {rel(last['fuse_logits'][0].permute(0, 4, 1, 2, 3), internals['fuse_logits']):.3e}")
        for l in range(4):
            prm = last["prm_logits"][l]
            prm0 = prm.view(last["passes"], B, *prm.shape[1:])[0].permute(0, 4, 1, 2, 3)
            de = last["de_f"][l]
            de0 = de.view(last["passes"], B, *de.shape[1:])[0].permute(0, 4, 1, 2, 3)
            print(f"  level {l + 1}: prm_logits rel {rel(prm0, internals['prm_logits'][l]):.3e}   de_f rel {rel(de0, internals['de_f'][l]):.3e}")
        for m in range(4):
            fl = last["fuse_logits"][1 + m].permute(0, 4, 1, 2, 3)
            print(f"  mod {m}: fuse_logits rel {rel(fl, internals['mod'][m][0]):.3e}")
        rows = []
        for k, p in model.named_parameters():
            if k.endswith(".conv.bias"):
                continue
            go = o_grads[k]
            rows.append((rel(p.grad, go), float(go.norm()), k))
        rows.sort(reverse=True)
        print("  worst grads:")
        for r, n, k in rows[:25]:
            print(f"    {r:.3e}  |g|={n:.3e}  {k}")
        flat = torch.cat([p.grad.flatten().cpu() for k, p in model.named_parameters() if not k.endswith('.conv.bias')])
        flat_o = torch.cat([o_grads[k].flatten() for k, p in model.named_parameters() if not k.endswith('.conv.bias')])
        print(f"  global grad rel-L2 {rel(flat, flat_o):.3e}; median per-tensor {np.median([r for r, _, _ in rows]):.3e}")


if __name__ == "__main__":
    main()
