"""Where do the torch fill / copy / add launches of one eager RFNet+PASSION step come from?  (torch.profiler with Python stacks)
    python scripts/profile_fills.py > gpurun_out/fills.txt"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from passion_b200.engine import Trainer  # noqa: E402
from passion_b200.models import build_model  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1037)
    model = build_model("rfnet", num_cls=4, crop=80).to(dev)
    tr = Trainer(model, lr=2e-4, weight_decay=1e-4, temp=4.0, mask_type="idt", use_passion=True, modal_weight=bench.modal_weight(), use_graph=False)
    host = bench.synth_host_batches(0, 1, 2, 80)
    b = tuple(t.to(dev) for t in host[0])
    for _ in range(3):
        tr.step(*b)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
        tr.step(*b)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.name in ("aten::fill_", "aten::zero_", "aten::copy_", "aten::add_", "aten::add", "aten::mul", "aten::cat", "aten::sum", "aten::to"):
            if e.self_device_time_total <= 0 and e.device_time_total <= 0:
                continue
            frames = [f for f in (e.stack or []) if "passion_b200" in f or "autograd" in f or "bench" in f][:2]
            key = (e.name, str(e.input_shapes)[:60], " <- ".join(f.split("/")[-1][:70] for f in frames))
            agg[key][0] += 1
            agg[key][1] += e.device_time_total / 1e3
    rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
    for (name, shp, where), (cnt, ms) in rows[:60]:
        print(f"x{cnt:<4d} {ms:7.3f} ms  {name:12s} {shp:60s} {where}")


if __name__ == "__main__":
    main()
