"""Developer probe: where issuer thread 0 of conv3_wgrad_rs_kernel spends its cycles (PB_WG_RS=5)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passion_b200 import _lib
from passion_b200._lib import PB_BF16, PB_PAD_REFLECT, ConvDesc
lib = _lib.load()
os.environ["PB_WG_RS"] = "5"
for (c0, cout, S, n) in [(16, 8, 80, 10), (8, 8, 80, 10), (32, 16, 40, 10)]:
    x = torch.randn(n, S, S, S, c0, device="cuda").bfloat16(); dy = torch.randn(n, S, S, S, cout, device="cuda").bfloat16()
    dw = torch.zeros(1, 27, c0, cout, device="cuda"); err = torch.zeros(1, dtype=torch.int32, device="cuda")
    d = ConvDesc(dtype=PB_BF16, n=n, di=S, hi=S, wi=S, dout=S, ho=S, wo=S, c0=c0, c1=0, cout=cout, ksize=3, stride=1, pad_mode=PB_PAD_REFLECT, groups=1)
    out = (ctypes.c_ulonglong * 8)()
    lib.pb_wgrad_rs_debug(out)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        lib.pb_conv3d_wgrad_tc(ctypes.byref(d), ctypes.c_void_p(x.data_ptr()), None, ctypes.c_void_p(dy.data_ptr()), ctypes.c_void_p(dw.data_ptr()), ctypes.c_void_p(err.data_ptr()), st)
    torch.cuda.synchronize()
    lib.pb_wgrad_rs_debug(out)
    v = list(out); ct = max(v[6], 1)
    print(f"c{c0}->{cout} {S}^3 n{n}: per CTA-launch: total {v[5]/ct:.0f} cyc, steps {v[4]/ct:.1f}, wait dy at item start {v[0]/ct:.0f}, wait dy in steps {v[1]/ct:.0f}, wait x {v[2]/ct:.0f}, issue {v[3]/ct:.0f}")
