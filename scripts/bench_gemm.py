"""Developer probe: the mmFormer token-path GEMM shapes on pb_gemm_tc against torch (cuBLAS), 50 back-to-back launches each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passion_b200 import ops                                # noqa: E402

SHAPES = [(250, 1536, 512), (250, 512, 512), (250, 4096, 512), (250, 512, 4096), (1000, 1536, 512), (1000, 4096, 512), (1000, 512, 4096),
          (128, 1536, 512), (1024, 4096, 512)]


def timeit(fn, reps=20):
    """GPU time per call: `reps` calls captured into one CUDA graph (the Python / ctypes launch cost would dominate otherwise)."""
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


def main():
    dev = "cuda"
    for (M, N, K) in SHAPES:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        dy = torch.randn(M, N, device=dev).bfloat16()
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        dw = torch.empty(N, K, device=dev, dtype=torch.float32)
        t_f = timeit(lambda: ops._gemm_tc("f", x, w, None, y, M, N, K, K, K, True, True))
        t_d = timeit(lambda: ops._gemm_tc("d", dy, w, None, dx, M, K, N, N, K, True, False))
        t_w = timeit(lambda: ops._gemm_tc("w", dy, x, None, dw, N, K, M, N, K, False, False))
        c_f = timeit(lambda: torch.matmul(x, w.t()))
        c_d = timeit(lambda: torch.matmul(dy, w))
        c_w = timeit(lambda: torch.matmul(dy.t(), x))
        print(f"M{M} N{N} K{K}: fwd {t_f:6.1f} us (cuBLAS {c_f:6.1f})  dgrad {t_d:6.1f} ({c_d:6.1f})  wgrad {t_w:6.1f} ({c_w:6.1f})", flush=True)
        ops.begin_step(dev)


if __name__ == "__main__":
    main()
