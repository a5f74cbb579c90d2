"""Developer probe: the attention of the mmFormer token path (ops.attention: batched tcgen05 GEMMs + row softmax) against the library's
fused attention, forward + backward, at the shapes of the 80^3 and 128^3 crops; per-kernel times from CUDA events recorded while the
GPU is kept behind the CPU (a long dummy kernel is queued first, so no event pair spans a launch gap)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passion_b200 import ops                              # noqa: E402

SHAPES = [(10, 500, 8, 64, 0.1), (8, 125, 8, 64, 0.1), (5, 2048, 8, 64, 0.1), (4, 512, 8, 64, 0.1), (5, 2048, 8, 64, 0.0)]


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    busy = torch.randn(8192, 8192, device="cuda")
    for (N, T, H, d, p) in SHAPES:
        qkv = torch.randn(N, T, 3, H, d, device="cuda").bfloat16().requires_grad_(True)
        go = torch.randn(N, T, H * d, device="cuda").bfloat16()

        def ours():
            qkv.grad = None
            ops.attention(qkv, p).backward(go)

        def lib():
            qkv.grad = None
            t = qkv.permute(2, 0, 3, 1, 4)
            h = torch.nn.functional.scaled_dot_product_attention(t[0], t[1], t[2], dropout_p=p)
            h.transpose(1, 2).reshape(N, T, H * d).backward(go)

        t_ours, t_lib = timed(ours), timed(lib)
        timer = ops.KernelTimer()
        for _ in range(3):
            busy @ busy
            busy @ busy
            ops.TIMER = timer
            ours()
            ops.TIMER = None
        torch.cuda.synchronize()
        summ = timer.summary()
        print(f"N{N} T{T} H{H} d{d} p{p}: ours {t_ours:.3f} ms, library {t_lib:.3f} ms (fwd + bwd)")
        for (name, key), v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"    {name:18s} {key:24s} {v['ms'] / v['calls']:.4f} ms  {v['bytes'] / v['ms'] / 1e6:7.0f} GB/s  {v['flops'] / v['ms'] / 1e9:6.1f} TFLOP/s")


if __name__ == "__main__":
    main()
