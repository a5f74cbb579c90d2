"""GPU parity tests of the individual CUDA kernels (through the C ABI / ops.py) against plain PyTorch
float64 references of the same op.  fp32 storage must agree to ~1e-5, bf16 storage to bf16 rounding."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def to_cl(t):           # NCDHW -> NDHWC contiguous
    return t.permute(0, 2, 3, 4, 1).contiguous()


def to_nc(t):
    return t.permute(0, 4, 1, 2, 3)


TOL = {torch.float32: 2e-5, torch.bfloat16: 8e-3}

CONV_CASES = [
    # c0, c1, cout, k, stride, pad, groups, n, (d,h,w), bias
    (1, 0, 8, 3, 1, "reflect", 4, 8, (12, 10, 14), False),
    (8, 0, 8, 3, 1, "reflect", 1, 2, (10, 12, 9), False),
    (8, 0, 16, 3, 2, "reflect", 4, 4, (12, 8, 10), False),
    (16, 16, 8, 3, 1, "reflect", 1, 3, (9, 10, 11), False),
    (32, 32, 16, 3, 1, "reflect", 1, 2, (6, 5, 7), False),
    (64, 0, 64, 3, 1, "reflect", 1, 2, (5, 6, 4), False),
    (64, 64, 32, 3, 1, "reflect", 1, 1, (4, 4, 4), False),
    (32, 0, 64, 3, 2, "reflect", 2, 2, (4, 6, 4), False),
    (2, 0, 2, 3, 1, "reflect", 1, 2, (8, 8, 8), False),
    (4, 0, 4, 3, 1, "reflect", 1, 2, (7, 8, 9), False),
    (8, 0, 8, 3, 1, "zeros", 1, 2, (6, 7, 8), False),
    (16, 0, 16, 3, 1, "reflect", 1, 1, (2, 2, 2), False),
    (32, 0, 2, 1, 1, "zeros", 1, 2, (8, 8, 8), False),
    (256, 0, 16, 1, 1, "zeros", 1, 2, (3, 4, 5), False),
    (16, 0, 4, 1, 1, "zeros", 1, 2, (6, 6, 6), True),
    (8, 8, 16, 1, 1, "zeros", 1, 3, (5, 6, 7), False),
    (2, 0, 8, 1, 1, "zeros", 1, 2, (6, 6, 6), False),
    (128, 0, 32, 1, 1, "zeros", 1, 2, (4, 4, 4), False),
    # larger volumes: several q-tiles / depth chunks of the tcgen05 path (bf16) and of the wgrad tiling
    (16, 0, 8, 3, 1, "reflect", 1, 3, (24, 20, 22), False),
    (8, 0, 8, 3, 1, "reflect", 4, 4, (18, 16, 20), False),
    (8, 8, 8, 3, 1, "reflect", 1, 2, (17, 13, 19), False),
    (32, 0, 32, 3, 1, "reflect", 1, 2, (10, 12, 10), False),
    (16, 16, 16, 3, 1, "zeros", 1, 2, (12, 10, 14), False),
    (64, 0, 32, 3, 1, "reflect", 1, 2, (8, 10, 6), False),
    (16, 0, 64, 3, 1, "reflect", 2, 2, (6, 8, 10), False),
    # biased 3x3 (pre-norm blocks of the mmFormer backbone) and wide channels
    (16, 0, 16, 3, 1, "reflect", 1, 2, (10, 9, 12), True),
    (8, 8, 8, 3, 1, "zeros", 2, 4, (8, 10, 9), True),
    (128, 0, 64, 3, 1, "reflect", 1, 1, (4, 4, 4), True),
    (64, 0, 128, 3, 2, "reflect", 1, 1, (4, 4, 4), True),
    (128, 128, 64, 3, 1, "reflect", 1, 1, (5, 5, 5), True),
    (128, 0, 128, 3, 1, "zeros", 1, 2, (5, 5, 5), True),
    (512, 0, 128, 1, 1, "zeros", 1, 2, (5, 5, 5), True),
    (32, 0, 8, 1, 1, "zeros", 1, 2, (6, 7, 8), True),
    # 1x1x1 on the TMA-fed tcgen05 GEMM: two K blocks and two Cout tiles, weight groups, ragged last tile, several super-tiles
    (256, 0, 64, 1, 1, "zeros", 1, 2, (6, 6, 6), True),
    (16, 0, 16, 1, 1, "zeros", 4, 4, (8, 9, 10), False),
    (8, 0, 8, 1, 1, "zeros", 1, 2, (20, 18, 22), False),
    # 128-wide planes (128^3 crops): producer budgets of the tcgen05 kernels
    (16, 0, 8, 3, 1, "reflect", 1, 1, (4, 6, 128), False),
    (8, 0, 8, 3, 1, "reflect", 1, 1, (3, 5, 128), True),
    (16, 0, 16, 3, 1, "zeros", 1, 1, (3, 4, 128), True),
    (64, 0, 32, 3, 1, "reflect", 1, 1, (4, 5, 32), True),
    # 128 input channels on planes wider than RFNet's 10^3: mmFormer's 64 + 64 -> 64 decoder conv at 16^3 (128^3 crop) and a W = 30 plane
    (64, 64, 64, 3, 1, "reflect", 1, 1, (16, 16, 16), True),
    (128, 0, 64, 3, 1, "zeros", 1, 1, (3, 20, 30), True),
    # stride 2 with odd extents and with zero padding (parity-class data gradient, double-buffered weight gradient: ragged tiles)
    (8, 0, 16, 3, 2, "zeros", 1, 2, (7, 9, 6), False),
    (16, 0, 32, 3, 2, "reflect", 2, 2, (9, 7, 11), True),
    (8, 0, 16, 3, 2, "reflect", 1, 1, (20, 18, 22), False),
    # very few channels: shared-memory tiled kernels (several tiles in d and in the plane, ragged edges, zero padding)
    (2, 0, 2, 3, 1, "reflect", 1, 2, (20, 18, 22), False),
    (2, 0, 2, 3, 1, "zeros", 1, 2, (6, 7, 8), True),
    (4, 0, 4, 3, 1, "reflect", 1, 1, (10, 33, 9), True),
    (4, 0, 4, 3, 1, "zeros", 1, 2, (9, 5, 6), False),
    (1, 0, 8, 3, 1, "reflect", 4, 4, (17, 20, 24), True),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "c%d+%d_%d_k%d_s%d_%s_g%d" % c[:7])
def test_conv3d(lib_built, case, dtype):
    from passion_b200 import ops
    c0, c1, cout, k, stride, pad, groups, n, (d, h, w), has_bias = case
    g = torch.Generator(device="cpu").manual_seed(hash(case[:7]) % 2 ** 31)
    dev = "cuda"
    cin = c0 + c1
    x = torch.randn(n, cin, d, h, w, generator=g).to(dev)
    wt = (torch.randn(groups, cout, cin, k, k, k, generator=g) / (cin * k ** 3) ** 0.5).to(dev)
    bias = torch.randn(groups, cout, generator=g).to(dev) if has_bias else None
    xq = x.to(dtype)                                   # the kernel sees dtype-rounded activations
    # ---- float64 reference on the same (rounded) inputs
    xr = xq.double().requires_grad_(True)
    # the tcgen05 path multiplies bf16 weights (fp32 accumulate); give the reference the same rounded operands
    uses_tc = (ops._tc_eligible(dtype, k, stride, c0, c1, cout)
               or ops._tc1_eligible(dtype, k, stride, c0, c1, cout, x.shape[2] * x.shape[3] * x.shape[4]))
    wr = (wt.to(torch.bfloat16) if uses_tc else wt).double().requires_grad_(True)
    ys = []
    npg = n // groups
    for gi in range(groups):
        xi = xr[gi * npg:(gi + 1) * npg]
        if k == 3:
            xi = F.pad(xi, (1,) * 6, mode="reflect" if pad == "reflect" else "constant")
        ys.append(F.conv3d(xi, wr[gi], None if bias is None else bias[gi].double(), stride=stride))
    yr = torch.cat(ys, 0)
    gy = torch.randn(yr.shape, generator=g).to(dev).to(dtype)
    yr.backward(gy.double())
    # ---- kernel path
    x_cl = to_cl(xq)
    x0 = x_cl[..., :c0].contiguous().requires_grad_(True)
    x1 = x_cl[..., c0:].contiguous().requires_grad_(True) if c1 else None
    wk = torch.stack([ops.kernel_layout(wt[gi]) for gi in range(groups)]).contiguous().requires_grad_(True)
    bk = bias.clone().requires_grad_(True) if has_bias else None
    y, stats = ops.conv3d(x0, wk, bk, x1, ksize=k, stride=stride, pad_mode=pad, groups=groups, want_stats=True)
    y.backward(to_cl(gy))
    tol = TOL[dtype]
    assert rel(to_nc(y), yr) < tol
    # statistics come from the fp32 accumulators (before storage rounding)
    assert rel(stats[..., 0], yr.sum((2, 3, 4))) < 1e-4 + (0 if dtype == torch.float32 else 0)
    assert rel(stats[..., 1], (yr * yr).sum((2, 3, 4))) < 1e-4
    dxr = to_cl(xr.grad)
    assert rel(x0.grad, dxr[..., :c0]) < tol
    if c1:
        assert rel(x1.grad, dxr[..., c0:]) < tol
    dwr = torch.stack([ops.kernel_layout(wr.grad[gi]) for gi in range(groups)])
    assert rel(wk.grad, dwr) < (5e-5 if dtype == torch.float32 else tol)
    if has_bias:
        npg_ = n // groups
        ref_db = torch.stack([gy.double()[gi * npg_:(gi + 1) * npg_].sum((0, 2, 3, 4)) for gi in range(groups)])
        assert rel(bk.grad, ref_db) < 1e-3
    ops.check_tc_errors()


# (c0, c1, cout, plane size): the stride-1 3x3x3 classes of RFNet at 80^3 and of mmFormer at 128^3, on the plane sizes they run at
ROUTED_CLASSES = [(8, 0, 8, 80), (8, 8, 8, 80), (16, 0, 16, 40), (16, 16, 16, 40), (32, 0, 32, 20), (32, 32, 32, 20), (64, 0, 64, 10),
                  (8, 0, 8, 128), (8, 8, 8, 128), (16, 0, 16, 64), (16, 16, 16, 64), (32, 0, 32, 32), (32, 32, 32, 32), (64, 0, 64, 16),
                  (64, 64, 64, 16), (128, 0, 128, 8)]


@pytest.mark.parametrize("case", ROUTED_CLASSES, ids=lambda c: "c%d+%d_%d_s%d" % c)
def test_conv3d_bf16_classes_stay_on_tcgen05(lib_built, case):
    """No silent fall-back: every stride-1 3x3x3 bf16 class of the two backbones, at the plane size it has in the benchmark
    configurations, must run its forward, data gradient and weight gradient on the tcgen05 kernels.  (ops falls back to the FFMA
    kernels when a tcgen05 launch reports PB_EUNSUPPORTED; the 64 + 64 -> 64 conv of mmFormer at 16^3 did so for a whole round
    because its plane exceeded the producer budget sized on RFNet's 10^3 — 3.7 ms of a 36 ms step.)"""
    from passion_b200 import ops
    c0, c1, cout, S = case
    g = torch.Generator(device="cpu").manual_seed(S + cout)
    d = min(S, 4)
    x0 = torch.randn(1, d, S, S, c0, generator=g).cuda().bfloat16().requires_grad_(True)
    x1 = torch.randn(1, d, S, S, c1, generator=g).cuda().bfloat16().requires_grad_(True) if c1 else None
    wk = (torch.randn(1, 27, c0 + c1, cout, generator=g) / (27 * (c0 + c1)) ** 0.5).cuda().requires_grad_(True)
    timer = ops.KernelTimer()
    ops.TIMER = timer
    try:
        y, _ = ops.conv3d(x0, wk, None, x1, ksize=3, stride=1, pad_mode="reflect")
        y.backward(torch.randn(y.shape, generator=g).cuda().bfloat16())
        torch.cuda.synchronize()
    finally:
        ops.TIMER = None
    names = {r[0] for r in timer.records}
    generic = names & {"conv3d_fwd", "conv3d_dgrad", "conv3d_wgrad", "conv3d_small_fwd", "conv3d_small_dgrad", "conv3d_small_wgrad"}
    if cout > 64:
        generic -= {"conv3d_wgrad"}            # the tcgen05 weight gradients hold Cout <= 64 rows (documented limit; 8^3 volumes only)
    assert not generic, (names, generic)
    assert {"conv3d_fwd_tc", "conv3d_dgrad_tc"} <= names
    ops.check_tc_errors()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", [(8, 0, 16, 3, 1, "reflect", 4, 4, (9, 8, 10), True), (16, 16, 8, 3, 1, "reflect", 1, 2, (8, 9, 10), False),
                                  (32, 0, 4, 1, 1, "zeros", 1, 2, (5, 6, 7), True), (8, 0, 16, 3, 2, "reflect", 4, 4, (8, 8, 6), True),
                                  (1, 0, 8, 3, 1, "reflect", 4, 4, (6, 7, 8), True), (64, 0, 32, 3, 1, "zeros", 1, 1, (4, 5, 6), True)],
                         ids=lambda c: "c%d+%d_%d_k%d_s%d_%s_g%d" % c[:7])
def test_conv3d_on_parameter_layout(lib_built, case, dtype):
    """ops.conv3d_ref (weights as nn.Conv3d stores them; one gather / one scatter launch for the layout work) must equal
    ops.conv3d on the kernel-layout copies bit for bit: same kernels, same operands."""
    from passion_b200 import ops
    c0, c1, cout, k, stride, pad, groups, n, (d, h, w), has_bias = case
    g = torch.Generator(device="cpu").manual_seed(7 + c0 + cout)
    cin = c0 + c1
    x = torch.randn(n, d, h, w, cin, generator=g).cuda().to(dtype)
    ws = [(torch.randn(cout, cin, k, k, k, generator=g) / (cin * k ** 3) ** 0.5).cuda().requires_grad_(True) for _ in range(groups)]
    bs = [torch.randn(cout, generator=g).cuda().requires_grad_(True) for _ in range(groups)] if has_bias else None
    gy = None
    res = []
    for ref_path in (True, False):
        x0 = x[..., :c0].contiguous().requires_grad_(True)
        x1 = x[..., c0:].contiguous().requires_grad_(True) if c1 else None
        for t in ws + (bs or []):
            t.grad = None
        if ref_path:
            y, st = ops.conv3d_ref(x0, ws, bs, x1, ksize=k, stride=stride, pad_mode=pad, want_stats=True)
        else:
            wk = torch.stack([ops.kernel_layout(wi) for wi in ws])
            bk = torch.stack(bs) if has_bias else None
            y, st = ops.conv3d(x0, wk, bk, x1, ksize=k, stride=stride, pad_mode=pad, groups=groups, want_stats=True)
        if gy is None:
            gy = torch.randn(y.shape, generator=g).cuda().to(dtype)
        y.backward(gy)
        res.append((y.detach(), st, x0.grad, None if x1 is None else x1.grad, [t.grad.clone() for t in ws],
                    [t.grad.clone() for t in bs] if has_bias else []))
    a, b = res
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
    assert rel(a[1], b[1]) < 1e-12                       # float64 atomics: order may differ
    if c1:
        assert torch.equal(a[3], b[3])
    for ga, gb in zip(a[4], b[4]):
        assert rel(ga, gb) < 1e-6                        # weight-gradient kernels accumulate with atomics
    for ga, gb in zip(a[5], b[5]):
        assert rel(ga, gb) < (1e-6 if dtype == torch.float32 else 2e-3)      # float64 sums vs torch's fp32 reduction
    ops.check_tc_errors()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,P,B,shape", [(8, 5, 2, (6, 5, 7)), (16, 1, 3, (4, 4, 4)), (64, 15, 1, (2, 3, 2))])
def test_masked_stack(lib_built, C, P, B, shape, dtype):
    """MaskModal for P passes in one launch (rfnet.py:154-163): out[p*B+b, ..., m*C+c] = enc[m*B+b, ..., c] * ms[p,b,m]."""
    from passion_b200 import ops
    g = torch.Generator().manual_seed(C + P)
    enc = torch.randn(4 * B, *shape, C, generator=g).cuda().to(dtype).requires_grad_(True)
    ms = (torch.rand(P, B, 4, generator=g) > 0.4).float().cuda()
    out = ops.masked_stack(enc, ms)
    ref_in = enc.detach().double().requires_grad_(True)
    st = ref_in.view(4, B, *shape, C).permute(1, 2, 3, 4, 0, 5)                       # [B,d,h,w,4,C]
    ref = (st[None] * ms.double().view(P, B, 1, 1, 1, 4, 1)).reshape(P * B, *shape, 4 * C)
    assert torch.equal(out.double(), ref.detach())                                   # masks are 0/1: exact
    gy = torch.randn(out.shape, generator=g).cuda().to(dtype)
    out.backward(gy)
    ref.backward(gy.double())
    assert rel(enc.grad, ref_in.grad) < (1e-6 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("c,shape,with_res", [(8, (12, 10, 14), True), (16, (6, 7, 8), False), (2, (8, 8, 8), False),
                                              (4, (5, 5, 5), True), (64, (2, 2, 2), False), (32, (4, 5, 6), True),
                                              (1, (6, 6, 6), False), (8, (40, 44, 36), True), (16, (30, 20, 28), False)])
def test_inorm_lrelu(lib_built, c, shape, with_res, dtype):
    from passion_b200 import ops
    g = torch.Generator().manual_seed(c * 131 + shape[0])
    n = 3
    y = (torch.randn(n, c, *shape, generator=g) * 2 + 0.7).cuda().to(dtype)
    res = torch.randn(n, c, *shape, generator=g).cuda().to(dtype) if with_res else None
    gy = torch.randn(n, c, *shape, generator=g).cuda().to(dtype)
    yr = y.double().requires_grad_(True)
    out_r = F.leaky_relu(F.instance_norm(yr, eps=1e-5), 0.2)
    if with_res:
        rr = res.double().requires_grad_(True)
        out_r = out_r + rr
    out_r.backward(gy.double())
    y_cl = to_cl(y).requires_grad_(True)
    res_cl = to_cl(res).requires_grad_(True) if with_res else None
    yd = y.double()
    stats = torch.stack((yd.sum((2, 3, 4)), (yd * yd).sum((2, 3, 4))), -1).contiguous()
    voxels = shape[0] * shape[1] * shape[2]
    mr = ops.inorm_finalize(stats, voxels)
    out = ops._InormLrelu.apply(y_cl, mr, res_cl)
    out.backward(to_cl(gy))
    tol = TOL[dtype]
    assert rel(to_nc(out), out_r) < tol
    assert rel(to_nc(y_cl.grad), yr.grad) < (1e-4 if dtype == torch.float32 else 2e-2)
    if with_res:
        assert rel(to_nc(res_cl.grad), rr.grad) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("c,shape,scale", [(8, (5, 6, 7), 2), (64, (2, 2, 2), 2), (4, (3, 4, 5), 4), (4, (2, 3, 2), 8),
                                           (16, (10, 10, 10), 2), (2, (4, 4, 4), 2), (32, (5, 5, 5), 2)])
def test_upsample(lib_built, c, shape, scale, dtype):
    from passion_b200 import ops
    g = torch.Generator().manual_seed(c + scale)
    x = torch.randn(2, c, *shape, generator=g).cuda().to(dtype)
    xr = x.double().requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=scale, mode="trilinear", align_corners=True)
    gy = torch.randn(yr.shape, generator=g).cuda().to(dtype)
    yr.backward(gy.double())
    x_cl = to_cl(x).requires_grad_(True)
    y = ops.upsample(x_cl, scale)
    y.backward(to_cl(gy))
    tol = TOL[dtype]
    assert rel(to_nc(y), yr) < tol
    assert rel(to_nc(x_cl.grad), xr.grad) < tol


def _gate_params(w0, b0, w2, b2):
    """stacked [4, ...] gate-MLP tensors -> the 16 parameters in the reference's nn.Conv3d shapes (ops._RfmRegion order)"""
    mk = lambda t: t.clone().requires_grad_(True)
    return ([mk(w0[i][:, :, None, None, None]) for i in range(4)] + [mk(b0[i]) for i in range(4)]
            + [mk(w2[i][:, :, None, None, None]) for i in range(4)] + [mk(b2[i]) for i in range(4)])


def _gate_grads(params):
    g = [t.grad for t in params]
    return (torch.stack([t.flatten(1) for t in g[0:4]]), torch.stack(g[4:8]), torch.stack([t.flatten(1) for t in g[8:12]]),
            torch.stack(g[12:16]))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("c,shape", [(8, (6, 7, 8)), (16, (4, 5, 6)), (64, (2, 2, 2)), (32, (3, 4, 3))])
def test_rfm_region(lib_built, c, shape, dtype):
    from passion_b200 import ops
    g = torch.Generator().manual_seed(c)
    n = 3
    kc = 4 * c
    y = torch.randn(n, *shape, kc, generator=g).cuda().to(dtype)
    y[0, ..., c:2 * c] = 0                                            # a missing modality
    p = torch.softmax(torch.randn(n, *shape, 4, generator=g), -1).cuda()
    w0 = (torch.randn(4, 128, kc + 1, generator=g) / kc ** 0.5).cuda()
    b0 = (torch.randn(4, 128, generator=g) * 0.1).cuda()
    w2 = (torch.randn(4, 4, 128, generator=g) / 128 ** 0.5).cuda()
    b2 = (torch.randn(4, 4, generator=g) * 0.1).cuda()
    gr = torch.randn(n, *shape, kc, generator=g).cuda().to(dtype)

    def reference(y, w0, b0, w2, b2):
        V = shape[0] * shape[1] * shape[2]
        yv = y.reshape(n, V, 4, c)
        pv = p.double().reshape(n, V, 4)
        outs = []
        for i in range(4):
            yp = yv * pv[:, :, i, None, None]                         # [n,V,4,c]
            prm_avg = pv[:, :, i].mean(1) + 1e-7
            feat = torch.cat(((yp.mean(1) / prm_avg[:, None, None]).reshape(n, kc), prm_avg[:, None]), 1)
            hdn = F.leaky_relu(feat @ w0[i].t() + b0[i], 0.2)
            gate = torch.sigmoid(hdn @ w2[i].t() + b2[i])             # [n,4]
            outs.append((yp * gate[:, None, :, None]).sum(2))         # [n,V,c]
        return torch.stack(outs, 2).reshape(n, *shape, kc)

    refs = [t.double().requires_grad_(True) for t in (y, w0, b0, w2, b2)]
    rr = reference(*refs)
    rr.backward(gr.double())
    yin = y.clone().requires_grad_(True)
    params = _gate_params(w0, b0, w2, b2)
    r = ops.rfm_region(yin, p, params)
    r.backward(gr)
    tol = TOL[dtype]
    assert rel(r, rr) < tol
    assert rel(yin.grad, refs[0].grad) < tol
    for a, b in zip(_gate_grads(params), refs[1:]):
        assert rel(a, b.grad) < (2e-4 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("c,shape", [(8, (6, 7, 8)), (16, (4, 5, 6)), (64, (2, 2, 2))])
def test_single_modality_passes_equal_their_stacks(lib_built, c, shape, dtype):
    """The four single-modality decoder passes computed on the modality-major encoder output itself (ops.rfm_region_single and
    the cin-sliced grouped 1x1x1 conv, conv3d_ref(..., slices=4)) must equal the same ops on their explicitly built
    [4B,...,4C] stacks (modality m in slot m, zeros elsewhere): values, input gradients and parameter gradients."""
    from passion_b200 import ops
    g = torch.Generator().manual_seed(100 + c)
    B, kc = 2, 4 * c
    enc = torch.randn(4 * B, *shape, c, generator=g).cuda().to(dtype)
    enc[1 * B + 0] = 0                                                # a missing modality of sample 0
    p = torch.softmax(torch.randn(4 * B, *shape, 4, generator=g), -1).cuda()
    w0 = (torch.randn(4, 128, kc + 1, generator=g) / kc ** 0.5).cuda()
    b0 = (torch.randn(4, 128, generator=g) * 0.1).cuda()
    w2 = (torch.randn(4, 4, 128, generator=g) / 128 ** 0.5).cuda()
    b2 = (torch.randn(4, 4, generator=g) * 0.1).cuda()
    wc = (torch.randn(c, kc, 1, 1, 1, generator=g) / kc ** 0.5).cuda()
    gr = torch.randn(4 * B, *shape, kc, generator=g).cuda().to(dtype)
    gc = torch.randn(4 * B, *shape, c, generator=g).cuda().to(dtype)
    ms = torch.eye(4)[:, None, :].expand(4, B, 4).contiguous().cuda()          # pass m: modality m only

    def run(single):
        e = enc.clone().requires_grad_(True)
        gp = _gate_params(w0, b0, w2, b2)
        wcp = wc.clone().requires_grad_(True)
        if single:
            r = ops.rfm_region_single(e, p, gp, B)
            y, st = ops.conv3d_ref(e, [wcp], ksize=1, pad_mode="zeros", want_stats=True, slices=4)
        else:
            stack = ops.masked_stack(e, ms)
            r = ops.rfm_region(stack, p, gp, B)
            y, st = ops.conv3d_ref(stack, [wcp], ksize=1, pad_mode="zeros", want_stats=True)
        (r.float() * gr.float()).sum().backward(retain_graph=True)
        (y.float() * gc.float()).sum().backward()
        return r.detach(), y.detach(), st.clone(), e.grad, list(_gate_grads(gp)) + [wcp.grad.clone()]

    a, b = run(True), run(False)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel(a[0], b[0]) < tol and rel(a[1], b[1]) < tol
    assert rel(a[2], b[2]) < 1e-4 if dtype == torch.float32 else rel(a[2], b[2]) < 1e-2
    assert rel(a[3], b[3]) < (1e-5 if dtype == torch.float32 else 2e-2)
    for ga, gb in zip(a[4], b[4]):
        assert rel(ga, gb) < (2e-4 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("passes,mode", [(5, 0), (1, 0), (4, 1)])
def test_fused_logit_loss_equals_unfused(lib_built, passes, mode, dtype):
    """ops.logit_loss (one pass over the logits: softmax, CE / Dice sums, temperature KL against the detached first pass) against
    the separate softmax4 / cedice_sums / kl_sums kernels: sums, returned probabilities and d/d logits."""
    from passion_b200 import ops
    g = torch.Generator().manual_seed(10 * passes + mode)
    B, shape, temp = 2, (6, 7, 8), 4.0
    logits = (3 * torch.randn(passes * B, *shape, 4, generator=g)).cuda().to(dtype)
    labels = torch.randint(0, 4, (B, *shape), generator=g).to(torch.uint8).cuda()
    n_ce = B if mode == 0 else passes * B
    w_ce = torch.randn(n_ce, 3, 4, generator=g).cuda()
    w_kl = torch.randn((passes - 1) * B, generator=g).cuda()
    w_p = torch.randn(B, *shape, 4, generator=g).cuda()

    def fused(use_probs):
        l = logits.clone().requires_grad_(True)
        ce, kl, probs = ops.logit_loss(l, labels, passes, mode, temp, want_probs=True)
        loss = (ce * w_ce).sum() + ((kl * w_kl).sum() if kl.numel() else 0.0) + ((probs * w_p).sum() if use_probs else 0.0)
        loss.backward()
        return ce.detach(), kl.detach(), probs.detach(), l.grad

    def unfused(use_probs):
        l = logits.clone().requires_grad_(True)
        lv = l.view(passes, B, *shape, 4)
        p0 = ops.softmax4(lv[0])
        if mode == 0:
            ce = ops.cedice_sums(p0, labels)
            kl = (ops.kl_sums(ops.softmax4(lv[1:].reshape((passes - 1) * B, *shape, 4), temp), ops.softmax4(lv[0].detach(), temp))
                  if passes > 1 else torch.zeros(0, device="cuda"))
        else:
            ce = ops.cedice_sums(ops.softmax4(l), labels)
            kl = torch.zeros(0, device="cuda")
        loss = (ce * w_ce).sum() + ((kl * w_kl).sum() if kl.numel() else 0.0) + ((p0 * w_p).sum() if use_probs else 0.0)
        loss.backward()
        return ce.detach(), kl.detach(), p0.detach(), l.grad

    for use_probs in (False, True):
        a, b = fused(use_probs), unfused(use_probs)
        assert rel(a[0], b[0]) < 1e-5 and rel(a[2], b[2]) < 1e-6
        if a[1].numel():
            assert rel(a[1], b[1]) < 1e-4
        assert rel(a[3], b[3]) < (1e-5 if dtype == torch.float32 else 1e-2)


# ------------------------------------------------------------------------------------------------ loss kernels
def _onehot_target(labels, num_cls=4):
    return F.one_hot(labels.long(), num_cls).permute(0, 4, 1, 2, 3).double()


@pytest.mark.parametrize("scale", [1, 2, 4])
def test_criterions_match_oracle(lib_built, scale):
    """The reference-signature loss wrappers (fused kernels) against the CPU oracle's restatement, values and
    gradients, including the trilinear up_op and the batched n % B label pairing."""
    from oracle import criterions_oracle as oc
    from passion_b200 import criterions as crit
    g = torch.Generator().manual_seed(scale)
    B, S = 2, 16
    s = S // scale
    labels = torch.randint(0, 4, (B, S, S, S), generator=g)
    labels[1][labels[1] == 3] = 0                                     # class 3 absent from sample 1 -> presence gate
    target = _onehot_target(labels)
    logit_s = torch.randn(B, 4, s, s, s, generator=g) * 3
    logit_t = torch.randn(B, 4, s, s, s, generator=g) * 3
    up = (lambda t: F.interpolate(t, scale_factor=scale, mode="trilinear", align_corners=True)) if scale > 1 else None
    for name in ("dice", "ce", "kl"):
        a = logit_s.clone().requires_grad_(True)
        ac = logit_s.cuda().requires_grad_(True)
        if name == "dice":
            ref = oc.dice_loss_bs(torch.softmax(a, 1), target, 4, up_op=up)
            out = crit.dice_loss_bs(torch.softmax(ac, 1), target.cuda(), num_cls=4, up_op=scale)
        elif name == "ce":
            ref = oc.softmax_weighted_loss_bs(torch.softmax(a, 1), target, 4, up_op=up)
            out = crit.softmax_weighted_loss_bs(torch.softmax(ac, 1), target.cuda(), num_cls=4, up_op=scale)
        else:
            ref = oc.temp_kl_loss_bs(a, logit_t, 4.0, up_op=up)
            out = crit.temp_kl_loss_bs(ac, logit_t.cuda(), target.cuda(), num_cls=4, temp=4.0, up_op=scale)
        assert out.shape == ref.shape == (B, 1)
        assert rel(out, ref) < 1e-5, name
        wv = torch.tensor([[0.7], [1.3]])
        (ref * wv).sum().backward()
        (out * wv.cuda()).sum().backward()
        assert rel(ac.grad, a.grad) < 1e-4, name


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_prototype_loss_matches_oracle(lib_built, dtype):
    from oracle import criterions_oracle as oc
    from passion_b200 import criterions as crit
    g = torch.Generator().manual_seed(5)
    B, S, C = 2, 12, 8
    labels = torch.randint(0, 4, (B, S, S, S), generator=g)
    labels[0][labels[0] == 2] = 1                                     # class 2 absent from sample 0
    target = _onehot_target(labels)
    fs = torch.randn(B, C, S, S, S, generator=g).to(dtype)
    ft = (fs.float() + 0.5 * torch.randn(B, C, S, S, S, generator=g)).to(dtype)
    a = fs.float().clone().requires_grad_(True)
    ref_p, ref_d = oc.prototype_passion_loss_bs(a, ft.float(), target, 4)
    ac = fs.cuda().requires_grad_(True)
    out_p, out_d = crit.prototype_passion_loss_bs(ac, ft.cuda(), target.cuda(), None, None, num_cls=4)
    assert rel(out_p, ref_p) < 1e-4 and rel(out_d, ref_d) < 1e-4
    wv = torch.tensor([[0.7], [1.3]])
    (ref_p * wv).sum().backward()
    (out_p * wv.cuda()).sum().backward()
    assert rel(ac.grad, a.grad) < (1e-4 if dtype == torch.float32 else 1e-2)


LINEAR_CASES = [  # tokens M, out features N, in features K, bias
    (250, 1536, 512, False),      # qkv of an intra-modal transformer at 80^3 (B = 2, 125 tokens)
    (1000, 512, 512, True),       # proj of the inter-modal transformer
    (1000, 4096, 512, True),      # FFN up
    (250, 512, 4096, True),       # FFN down
    (128, 128, 64, True),         # exactly one tile, one K block
    (77, 24, 40, True),           # ragged everywhere (rows, columns and K beyond the tensor = TMA zero fill)
    (300, 200, 136, False),
]


@pytest.mark.parametrize("case", LINEAR_CASES, ids=lambda c: "m%d_n%d_k%d_b%d" % c)
def test_linear_tc(lib_built, case):
    """Token-path GEMM (csrc/gemm_tc.cu) behind ops.linear: forward, data gradient and weight gradient against float64 matmuls of
    the same bf16-rounded operands."""
    from passion_b200 import ops
    M, N, K, has_bias = case
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, generator=g).cuda().bfloat16().requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().requires_grad_(True)
    b = torch.randn(N, generator=g).cuda().requires_grad_(True) if has_bias else None
    assert ops.linear_tc_eligible(x, w)
    y = ops.linear(x, w, b)
    gy = torch.randn(M, N, generator=g).cuda().bfloat16()
    y.backward(gy)
    xr = x.detach().double().requires_grad_(True)
    wr = w.detach().bfloat16().double().requires_grad_(True)
    br = b.detach().double().requires_grad_(True) if has_bias else None
    yr = xr @ wr.t() + (br if has_bias else 0)
    yr.backward(gy.double())
    assert rel(y, yr) < 8e-3
    assert rel(x.grad, xr.grad) < 8e-3
    assert rel(w.grad, wr.grad) < 1e-5          # fp32 accumulators stored as fp32
    if has_bias:
        assert rel(b.grad, br.grad) < 1e-5
    ops.check_tc_errors()


def test_linear_tc_3d_and_fallback(lib_built):
    from passion_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(2, 125, 512, generator=g).cuda()
    w = (torch.randn(1536, 512, generator=g) / 512 ** 0.5).cuda()
    y16 = ops.linear(x.bfloat16(), w)
    y32 = ops.linear(x, w)                                  # fp32 check mode: library GEMM
    assert y16.shape == (2, 125, 1536) and y16.dtype == torch.bfloat16 and y32.dtype == torch.float32
    assert rel(y16, y32) < 1.5e-2
    ops.check_tc_errors()


WGRAD_RS_CASES = [
    # c0, c1, cout, pad, groups, n, (d, h, w): every unit width of the row-stacked weight gradient (csrc/conv3d_wgrad_rs.cu) — 8-channel
    # plain rows, 16-channel SWIZZLE_32B rows, 32-channel SWIZZLE_64B rows, several chunk groups / output chunks per launch, two sources
    # that straddle a unit, zero padding, ragged strips and depth chunks, W = 128 and W not a multiple of 16 (K tail = zero fill)
    (8, 0, 8, "reflect", 1, 2, (9, 11, 16)),
    (16, 0, 8, "reflect", 1, 2, (7, 19, 20)),
    (32, 0, 16, "reflect", 1, 2, (6, 9, 12)),
    (32, 0, 32, "zeros", 2, 4, (5, 7, 10)),
    (64, 0, 32, "reflect", 1, 1, (5, 6, 9)),
    (64, 64, 64, "reflect", 1, 1, (4, 5, 6)),
    (8, 16, 16, "reflect", 1, 2, (6, 8, 11)),
    (16, 16, 8, "zeros", 1, 1, (18, 17, 40)),
    (24, 0, 8, "reflect", 1, 1, (5, 6, 7)),
    (8, 0, 8, "reflect", 1, 1, (3, 5, 128)),
    (16, 0, 16, "reflect", 4, 4, (20, 21, 22)),
]


@pytest.mark.parametrize("case", WGRAD_RS_CASES, ids=lambda c: "c%d+%d_%d_%s_g%d_n%d" % c[:6])
def test_wgrad_rs_classes(lib_built, case, monkeypatch):
    """The row-stacked tcgen05 weight gradient on every unit width and routing branch, forced on for small volumes
    (PB_WG_RS_MIN_VOX=0; by default volumes below 20^3 stay on the kh-stacked kernels), against a float64 reference of the same
    bf16-rounded operands; and the same launch through the old kernels (PB_WG_RS=0) as a second opinion."""
    from passion_b200 import ops
    c0, c1, cout, pad, groups, n, (d, h, w) = case
    cin = c0 + c1
    g = torch.Generator(device="cpu").manual_seed(cin * 131 + cout * 7 + d * h * w)
    x = torch.randn(n, cin, d, h, w, generator=g).cuda().bfloat16()
    wt = (torch.randn(groups, cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5).cuda()
    gy = torch.randn(n, cout, d, h, w, generator=g).cuda().bfloat16()
    xr = x.double()
    wr = wt.bfloat16().double().requires_grad_(True)
    npg = n // groups
    ys = []
    for gi in range(groups):
        xi = F.pad(xr[gi * npg:(gi + 1) * npg], (1,) * 6, mode="reflect" if pad == "reflect" else "constant")
        ys.append(F.conv3d(xi, wr[gi]))
    torch.cat(ys, 0).backward(gy.double())
    dwr = torch.stack([ops.kernel_layout(wr.grad[gi]) for gi in range(groups)])
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("PB_WG_RS", mode)
        monkeypatch.setenv("PB_WG_RS_MIN_VOX", "0")
        x_cl = to_cl(x)
        x0 = x_cl[..., :c0].contiguous().requires_grad_(True)
        x1 = x_cl[..., c0:].contiguous().requires_grad_(True) if c1 else None
        wk = torch.stack([ops.kernel_layout(wt[gi]) for gi in range(groups)]).contiguous().requires_grad_(True)
        y, _ = ops.conv3d(x0, wk, None, x1, ksize=3, stride=1, pad_mode=pad, groups=groups, want_stats=False)
        y.backward(to_cl(gy))
        got[mode] = wk.grad.clone()
        assert rel(wk.grad, dwr) < 2e-5, (mode, rel(wk.grad, dwr))      # fp32 accumulation of exact bf16 products
    assert rel(got["1"], got["0"]) < 2e-5
    ops.check_tc_errors()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 5, 6, 7, 4), (3, 9, 10, 11, 8), (1, 33, 20, 17, 16), (2, 8, 8, 8, 64), (2, 4, 4, 4, 128),
                                   (1, 2, 2, 2, 512), (5, 24, 24, 24, 8), (2, 6, 5, 3, 2)], ids=lambda s: "x".join(map(str, s)))
def test_channel_stats_and_prenorm(lib_built, dtype, shape):
    """ops.channel_stats (per-(n, c) sum and sum of squares of an arbitrary tensor: the statistics of a PRE-norm block, reference
    models/blocks.py:312-316) against float64, and ops.prenorm (InstanceNorm -> LeakyReLU(0.2)) built on it, forward and backward."""
    from passion_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g) * 1.5 + 0.3).cuda().to(dtype)
    ops.begin_step(torch.device("cuda", 0))
    st = ops.channel_stats(x)
    xd = x.double()
    assert st.dtype == torch.float64 and st.shape == (shape[0], shape[-1], 2)
    assert rel(st[..., 0], xd.sum((1, 2, 3))) < 1e-6 + (0 if dtype == torch.float32 else 1e-5)
    assert rel(st[..., 1], (xd * xd).sum((1, 2, 3))) < 1e-6 + (0 if dtype == torch.float32 else 1e-5)
    xq = x.clone().requires_grad_(True)
    y = ops.prenorm(xq)
    gy = torch.randn(*shape, generator=g).cuda().to(dtype)
    y.backward(gy)
    xr = xd.clone().requires_grad_(True)
    xn = xr.permute(0, 4, 1, 2, 3)
    yr = F.leaky_relu(F.instance_norm(xn, eps=1e-5), 0.2).permute(0, 2, 3, 4, 1)
    yr.backward(gy.double())
    tol = TOL[dtype]
    assert rel(y, yr) < tol
    assert rel(xq.grad, xr.grad) < (tol if dtype == torch.bfloat16 else 1e-4)
