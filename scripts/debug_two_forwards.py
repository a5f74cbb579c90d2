"""Two forwards through one model followed by ONE backward (the chunk loop of scripts/ddp_equivalence.py) against two separate
forward/backward rounds: isolates the accumulation of conv weight gradients over several uses of a parameter."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from passion_b200.models import rfnet
from passion_b200.train_step import loss_mix

dev = torch.device("cuda", 0)
B, S = 2, 32
x, target, mask, _ = synth.make_batch(2 * B, S, seed=77, labels="U", mask_ids=[10, 7, 12, 5])
sd = synth.make_state_dict(1037)
beta = torch.tensor([1.1, 0.9, 1.3, 0.7]).to(dev)
mw = torch.tensor([2.4, 1.6, 1.2, 5.1]).to(dev)

def fresh():
    m = rfnet.Model(4).to(dev); m.load_state_dict(sd); m.compute_dtype = torch.float32
    m.is_training, m.use_passion, m.mask_type = True, True, "idt"
    return m

def flat(m):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten().double() for p in m.parameters()])

ma = fresh()
gs = []
for c in range(2):
    sc = slice(c * B, (c + 1) * B)
    ma.zero_grad(set_to_none=True)
    t = target[sc].to(dev)
    outs = ma(x[sc].to(dev), mask[sc].to(dev), target=t, temp=4.0)
    l, _ = loss_mix(outs, t, mask[sc].to(dev), beta, mw)
    l.backward()
    gs.append(flat(ma).clone())
g_sep = gs[0] + gs[1]
mb = fresh()
tot = 0.0
for c in range(2):
    sc = slice(c * B, (c + 1) * B)
    t = target[sc].to(dev)
    outs = mb(x[sc].to(dev), mask[sc].to(dev), target=t, temp=4.0)
    l, _ = loss_mix(outs, t, mask[sc].to(dev), beta, mw)
    tot = tot + l
tot.backward()
g_joint = flat(mb)
print("joint vs separate rel-L2:", float((g_joint - g_sep).norm() / g_sep.norm()))
names = [n for n, _ in mb.named_parameters()]
off = 0
bad = []
for n, p in mb.named_parameters():
    k = p.numel()
    a, b = g_joint[off:off + k], g_sep[off:off + k]
    r = float((a - b).norm() / b.norm().clamp_min(1e-30))
    if r > 1e-4 and float(b.norm()) > 1e-8:
        bad.append((n, round(r, 4), float(a.norm()), float(b.norm())))
    off += k
print(len(bad), "tensors differ; first:", bad[:12])
