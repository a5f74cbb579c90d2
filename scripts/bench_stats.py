"""Developer probe: ops.channel_stats on the level shapes of the pre-norm backbone (mmFormer, 128^3 and 80^3 crops), CUDA-event timed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passion_b200 import ops                              # noqa: E402

for (n, s, c) in [(5, 128, 8), (4, 128, 8), (5, 128, 16), (5, 64, 16), (5, 64, 32), (5, 32, 64), (5, 16, 128), (10, 80, 8), (10, 80, 16), (10, 40, 32)]:
    x = torch.randn(n, s, s, s, c, device="cuda").bfloat16()
    for _ in range(3):
        ops.channel_stats(x)
    ops.begin_step(torch.device("cuda", 0))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.channel_stats(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"channel_stats n{n} {s}^3 c{c}: {ms:.4f} ms  {x.numel() * 2 / ms / 1e6:.0f} GB/s")
    ops.begin_step(torch.device("cuda", 0))
