"""CPU experiment behind DESIGN.md "Conditioning": what bf16 STORAGE of activations costs this network,
independent of any kernel.  Rounds the raw conv outputs and the activations of the ORACLE to bf16
(straight-through in backward) and compares logits / gradients with the unrounded fp32 oracle; also measures
the float64 oracle's sensitivity to a tiny input perturbation.      python scripts/bf16_sim.py [S]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rfnet_oracle as ro, synth, train_step_oracle as ts   # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


class RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


MODE = {"y": False, "a": False}


def patched(P, name, x, k=3, stride=1, pad_mode="reflect"):
    w = P[name + ".conv.weight"]
    xx = F.pad(x, (1,) * 6, mode=pad_mode) if k == 3 else x
    y = F.conv3d(xx, w, None, stride=stride)
    if MODE["y"]:
        y = RoundSTE.apply(y)
    y = F.leaky_relu(F.instance_norm(y, eps=1e-5), 0.2)
    if MODE["a"]:
        y = RoundSTE.apply(y)
    return y


def run(sd, x, target, mask, dtype=torch.float32):
    beta = torch.tensor([1.1, 0.9, 1.3, 0.7], dtype=dtype)
    mw = torch.tensor([219 / 90., 219 / 135., 219 / 184., 219 / 43.], dtype=dtype)
    P = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    outs, it = ro.forward(P, x.to(dtype), mask, target, 2.0, return_internals=True)
    loss, _ = ts.loss_mix(outs, target, mask, beta, mw)
    loss.backward()
    g = torch.cat([P[k].grad.flatten() for k in P if not k.endswith('.conv.bias')])
    return it["fuse_logits"].detach(), g


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    sd = synth.make_state_dict(1037)
    x, target, mask, _ = synth.make_batch(1, S, seed=3, labels="S", mask_ids=[11])
    base = run(sd, x, target, mask)
    orig = ro.conv_in_lrelu
    ro.conv_in_lrelu = patched
    for tag, mode in (("bf16 y+a", dict(y=True, a=True)), ("bf16 a only", dict(y=False, a=True))):
        MODE.update(mode)
        o = run(sd, x, target, mask)
        print(f"S={S} {tag:12s}: logits rel-L2 {rel(o[0], base[0]):.2e}   weight-grad rel-L2 {rel(o[1], base[1]):.2e}")
    ro.conv_in_lrelu = orig
    b64 = run(sd, x, target, mask, torch.float64)
    print(f"S={S} fp32 oracle vs fp64 oracle: logits {rel(base[0], b64[0]):.2e}  weight-grad {rel(base[1], b64[1]):.2e}")
    torch.manual_seed(0)
    for eps in (1e-6, 1e-5):
        xp = x.double() * (1 + eps * torch.randn_like(x.double()))
        o = run(sd, xp, target, mask, torch.float64)
        print(f"S={S} fp64 oracle, input perturbed by {eps:g}: logits {rel(o[0], b64[0]):.2e}  weight-grad {rel(o[1], b64[1]):.2e}")


if __name__ == "__main__":
    main()
