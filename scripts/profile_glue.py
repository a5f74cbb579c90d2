"""Which torch (non-library) kernels run inside one eager RFNet+PASSION step, by aten op and input shape (torch.profiler).
    python scripts/profile_glue.py > gpurun_out/glue.txt"""
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
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        tr.step(*b)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages(group_by_input_shape=True):
        if e.self_device_time_total > 0 and (e.key.startswith("aten::") or "Backward" in e.key):
            rows.append((e.self_device_time_total / 1e3, e.count, e.key, str(e.input_shapes)[:150]))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"torch ops with device time in one eager step: {tot:.3f} ms total")
    for ms, cnt, key, shp in rows[:70]:
        print(f"{ms:8.3f} ms  x{cnt:<4d} {key:40s} {shp}")


if __name__ == "__main__":
    main()
