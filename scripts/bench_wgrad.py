"""Developer probe: the 3x3x3 weight-gradient classes of one RFNet step (B = 2, 80^3), old kh-stacked kernels (PB_WG_RS=0)
against the row-stacked kernel (PB_WG_RS=1) through the same C-ABI entry point — results compared, launches timed with CUDA
events (inputs 80-250 MB per launch at the fine levels, i.e. larger than L2; the coarse levels are L2-resident in the real step too)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passion_b200 import _lib                              # noqa: E402
from passion_b200._lib import PB_BF16, PB_PAD_REFLECT, PB_PAD_ZERO, ConvDesc   # noqa: E402

MODES = [int(v) for v in os.environ.get("MODES", "0,1").split(",")]
CLASSES = [  # c0, c1, cout, S, n, groups, launches per step
    (16, 0, 8, 80, 10, 1, 2), (16, 0, 8, 80, 8, 1, 2), (8, 0, 8, 80, 10, 1, 2), (8, 0, 8, 80, 8, 4, 2),
    (32, 0, 16, 40, 10, 1, 2), (32, 0, 16, 40, 8, 1, 2), (16, 0, 16, 40, 10, 1, 2), (16, 0, 16, 40, 8, 4, 2),
    (64, 0, 32, 20, 10, 1, 2), (64, 0, 32, 20, 8, 1, 2), (32, 0, 32, 20, 10, 1, 2), (32, 0, 32, 20, 8, 4, 2),
    (64, 0, 64, 10, 10, 1, 2), (64, 0, 64, 10, 8, 4, 2), (8, 8, 8, 80, 2, 1, 0), (16, 16, 16, 40, 2, 1, 0),
]


def main():
    lib = _lib.load()
    dev = "cuda"
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    total = {m: 0.0 for m in MODES}
    reps = int(os.environ.get("REPS", 5))
    for (c0, c1, cout, S, n, groups, per_step) in CLASSES:
        for pad in (PB_PAD_REFLECT, PB_PAD_ZERO):
            if pad == PB_PAD_ZERO and not (S == 40 and n == 10):
                continue
            g = torch.Generator(device="cpu").manual_seed(c0 * 100 + cout + S)
            x0 = torch.randn(n, S, S, S, c0, generator=g).to(dev).bfloat16()
            x1 = torch.randn(n, S, S, S, c1, generator=g).to(dev).bfloat16() if c1 else None
            dy = torch.randn(n, S, S, S, cout, generator=g).to(dev).bfloat16()
            d = ConvDesc(dtype=PB_BF16, n=n, di=S, hi=S, wi=S, dout=S, ho=S, wo=S, c0=c0, c1=c1, cout=cout, ksize=3, stride=1,
                         pad_mode=pad, groups=groups)
            res, ms = {}, {}
            for mode in MODES:
                os.environ["PB_WG_RS"] = str(mode)
                dw = torch.zeros(groups, 27, c0 + c1, cout, device=dev)
                args = (ctypes.byref(d), ctypes.c_void_p(x0.data_ptr()), ctypes.c_void_p(x1.data_ptr()) if c1 else None,
                        ctypes.c_void_p(dy.data_ptr()), ctypes.c_void_p(dw.data_ptr()), ctypes.c_void_p(err.data_ptr()), st)
                rc = lib.pb_conv3d_wgrad_tc(*args)
                assert rc == 0, (rc, _lib.last_error())
                torch.cuda.synchronize()
                res[mode] = dw.clone()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    lib.pb_conv3d_wgrad_tc(*args)
                e1.record()
                torch.cuda.synchronize()
                ms[mode] = e0.elapsed_time(e1) / reps
                total[mode] += ms[mode] * per_step
            relerr = float((res[MODES[-1]] - res[MODES[0]]).norm() / res[MODES[0]].norm())
            nbytes = n * S ** 3 * (c0 + c1 + cout) * 2
            print("c%d+%d->%d %d^3 n%d g%d %s | %s | last mode: %.0f GB/s, %.1f TFLOP/s | new vs old rel %.2e | err flag %d" % (
                c0, c1, cout, S, n, groups, "reflect" if pad == PB_PAD_REFLECT else "zeros",
                "  ".join("m%d %.3f ms" % (m, ms[m]) for m in MODES), nbytes / ms[MODES[-1]] / 1e6,
                2 * 27 * (c0 + c1) * cout * n * S ** 3 / ms[MODES[-1]] / 1e9, relerr, int(err.item())), flush=True)
    print("per step (RFNet classes): " + ", ".join("mode %d: %.3f ms" % (m, total[m]) for m in MODES))


if __name__ == "__main__":
    main()
