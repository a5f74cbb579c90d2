"""Developer probe: where the MMA-issuing thread of the kw-stacked 3x3x3 kernel spends its cycles (PB_TC_PROBE=1): forward (reflect padding,
cp.async producers) and data gradient (full correlation, TMA producer) of the Cout = 8 classes at 80^3."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PB_TC_PROBE"] = "1"
from passion_b200 import _lib, ops
lib = _lib.load()
lib.pb_conv3d_tc_debug.argtypes = [ctypes.c_void_p]

def report(tag):
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 8)()
    lib.pb_conv3d_tc_debug(out)
    v = list(out); ct = max(v[5], 1)
    if v[5]:
        print(f"{tag}: per CTA: total {v[4]/ct:.0f} cyc, plane-steps {v[3]/ct:.1f}, wait input planes {v[0]/ct:.0f}, wait free accumulator block {v[1]/ct:.0f}, "
              f"issue {v[2]/ct:.0f}; epilogue thread 0 waiting for a full block (all CTAs) {v[6]/ct:.0f}")

for (cin, cout, S, n) in [(16, 8, 80, 10), (8, 8, 80, 10)]:
    x = torch.randn(n, S, S, S, cin, device="cuda").bfloat16().requires_grad_(True)
    w = (torch.randn(1, 27, cin, cout, device="cuda") / (27 * cin) ** 0.5).requires_grad_(True)
    out = (ctypes.c_ulonglong * 8)(); lib.pb_conv3d_tc_debug(out)
    for _ in range(2):
        y, _ = ops.conv3d(x, w, None, None, ksize=3, stride=1, pad_mode="reflect", groups=1, want_stats=True)
    report(f"c{cin}->{cout} {S}^3 n{n} forward x2")
    gy = torch.randn_like(y)
    os.environ["PB_WGRAD_TC"] = "1"
    y.backward(gy)
    report(f"c{cin}->{cout} {S}^3 n{n} data gradient x1")
