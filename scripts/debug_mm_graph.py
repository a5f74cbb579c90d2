"""Debug: mmFormer + PASSION step, eager vs CUDA-graph replay with dropout off, same weights and batch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from passion_b200.engine import Trainer
from passion_b200.models import build_model
from passion_b200 import ops

def run(use_graph, steps=6, train_mode=False, vary=False):
    torch.manual_seed(1037)
    dev = torch.device("cuda", 0)
    model = build_model("mmformer", num_cls=4, crop=80).to(dev)
    if not train_mode:
        model.eval()                  # dropout off; Trainer sets is_training itself
    tr = Trainer(model, lr=2e-4, weight_decay=1e-4, temp=4.0, mask_type="idt", use_passion=True, modal_weight=bench.modal_weight(), use_graph=use_graph)
    hb = bench.synth_host_batches(0, 1, 2, 80)
    bs = [tuple(t.to(dev) for t in h) for h in hb]
    out = []
    for i in range(steps):
        loss, _ = tr.step(*bs[i % len(bs) if vary else 0])
        out.append(float(loss))
    ops.check_tc_errors()
    return out

print("eager", run(False))
print("graph", run(True, steps=6))
print("graph, dropout on, varying batches", run(True, steps=14, train_mode=True, vary=True))
print("eager, dropout on, varying batches", run(False, steps=8, train_mode=True, vary=True))
