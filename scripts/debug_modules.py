"""Developer probe: per-level input-gradient error of Decoder_sep (fp32 check mode) against the oracle in fp32 AND in float64."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, rfnet_oracle as O
from passion_b200 import ops
from passion_b200.models import rfnet

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())
cl = lambda t: t.permute(0, 2, 3, 4, 1).contiguous()
nc = lambda t: t.permute(0, 4, 1, 2, 3)
sd = synth.make_state_dict(1037)
model = rfnet.Model(4).cuda(); model.load_state_dict(sd); model.compute_dtype = torch.float32
g = torch.Generator().manual_seed(5)
xs = [torch.randn(2, 8 * 2 ** i, 16 >> i, 16 >> i, 16 >> i, generator=g) for i in range(4)]
w = torch.randn(2, 4, 16, 16, 16, generator=g)
ops.begin_step(torch.device("cuda", 0))
xc = [cl(t).cuda().requires_grad_(True) for t in xs]
prob = torch.softmax(nc(model.decoder_sep.run(*xc)).float(), 1)
(prob * w.cuda()).sum().backward()
out = {}
for dt in (torch.float32, torch.float64):
    P = {k: v.to(dt) if v.is_floating_point() else v for k, v in sd.items()}
    xr = [t.detach().clone().to(dt).requires_grad_(True) for t in xs]
    ref = O.decoder_sep(P, *xr)
    (ref * w.to(dt)).sum().backward()
    out[dt] = (ref, [t.grad for t in xr])
print("prob: cuda vs f32 %.2e, cuda vs f64 %.2e, f32 vs f64 %.2e" % (rel(prob, out[torch.float32][0]), rel(prob, out[torch.float64][0]),
      rel(out[torch.float32][0], out[torch.float64][0])))
for i in range(4):
    a = nc(xc[i].grad)
    print("level %d grad: cuda vs f32 %.2e, cuda vs f64 %.2e, f32 oracle vs f64 %.2e" % (i + 1, rel(a, out[torch.float32][1][i]),
          rel(a, out[torch.float64][1][i]), rel(out[torch.float32][1][i], out[torch.float64][1][i])))
