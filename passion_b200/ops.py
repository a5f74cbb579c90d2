"""torch.autograd wrappers around the C-ABI kernels (include/passion_b200.h).

All activations are dense channels-last tensors of shape [N, D, H, W, C] ("cl"), dtype float32
(check mode) or bfloat16.  Every op requires CUDA tensors; there is no eager/CPU fallback.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import PB_BF16, PB_F32, PB_PAD_REFLECT, PB_PAD_ZERO, ConvDesc

LRELU_SLOPE = 0.2     # reference models/blocks.py:355 relufactor
IN_EPS = 1e-5         # nn.InstanceNorm3d default (models/blocks.py:18)


def _dt(t):
    if t.dtype == torch.bfloat16:
        return PB_BF16
    if t.dtype == torch.float32:
        return PB_F32
    raise TypeError(f"passion_b200: unsupported dtype {t.dtype}")


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("passion_b200 ops need CUDA tensors (no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("passion_b200 ops need contiguous tensors")


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class KernelTimer:
    """Optional per-launch CUDA-event timing of the library's kernels (bench.py's live roofline numbers).
    Events are recorded on the launching stream around each C-ABI call."""

    def __init__(self):
        self.records = []

    def run(self, name, key, nbytes, flops, fn):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn()
        e1.record()
        self.records.append((name, key, nbytes, flops, e0, e1))
        return rc

    def summary(self):
        """{(name, key): dict(calls, ms, bytes, flops)} — call after torch.cuda.synchronize()."""
        out = {}
        for name, key, nbytes, flops, e0, e1 in self.records:
            d = out.setdefault((name, key), dict(calls=0, ms=0.0, bytes=0, flops=0))
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["bytes"] += nbytes
            d["flops"] += flops
        return out


TIMER = None          # set to a KernelTimer() to time every launch


def _run(name, key, nbytes, flops, fn, allow_unsupported=False):
    rc = TIMER.run(name, key, nbytes, flops, fn) if TIMER is not None else fn()
    if allow_unsupported and rc == -3:          # PB_EUNSUPPORTED: caller falls back to the generic kernel
        return False
    _lib.check(rc, name)
    return True


def _publish_to_all_streams(device):
    """A buffer that was just created and filled on the CURRENT stream is about to be shared by ops on other streams (the
    models run decoder_sep on a side stream): make the fill visible everywhere.  Only happens while the arenas grow, i.e.
    in the first eager steps; never inside a CUDA-graph capture (the trainer's warm-up steps have sized everything by then)."""
    if device.type == "cuda" and not torch.cuda.is_current_stream_capturing():
        torch.cuda.current_stream(device).synchronize()


class _ZeroScratch:
    """Zero-initialised short-lived scratch (statistics / weight-gradient accumulators / split-K partials that kernels fill with
    atomics and the very next launch consumes).  Slices of one arena per device, handed out linearly; `begin_step()` re-zeroes the
    used part with ONE memset instead of one fill launch per buffer (~350 per training step).

    The arena must not grow inside a step that is being captured into a CUDA graph: the replayed memset would then cover only the part
    used since the last growth, and the accumulators beyond it would start every replay from the previous replay's sums (this is what
    turned the mmFormer step into NaNs after three replays once the split-K workspaces pushed the demand past the initial 8 MB).
    `begin_step()` therefore sizes the arena for the TOTAL demand of the previous step, so the step after a growing one never grows,
    and `generation(device)` lets the trainer verify that a capture saw a stable arena."""
    MIN_BYTES = 8 << 20

    def __init__(self):
        self.arenas = {}          # device -> [tensor, used bytes, bytes requested since begin_step, generation]

    def begin_step(self, device):
        a = self.arenas.get(device)
        if a is None:
            return
        if a[2] > a[0].numel() and not (device.type == "cuda" and torch.cuda.is_current_stream_capturing()):
            # the previous step outgrew the arena on its way: give the coming step one that holds all of it
            self.arenas[device] = [torch.zeros(max(self.MIN_BYTES, 2 * a[2]), dtype=torch.uint8, device=device), 0, 0, a[3] + 1]
            _publish_to_all_streams(device)
            return
        if a[1] > 0:
            a[0][:a[1]].zero_()
        a[1] = 0
        a[2] = 0

    def zeros(self, shape, dtype, device):
        n = 1
        for v in shape:
            n *= int(v)
        nbytes = (n * torch.empty((), dtype=dtype).element_size() + 255) & ~255
        a = self.arenas.get(device)
        if a is None or a[1] + nbytes > a[0].numel():
            size = max(self.MIN_BYTES, 2 * nbytes, 2 * (a[0].numel() if a is not None else 0))
            a = [torch.zeros(size, dtype=torch.uint8, device=device), 0, a[2] if a is not None else 0, (a[3] + 1) if a is not None else 0]
            self.arenas[device] = a
            _publish_to_all_streams(device)
        off = a[1]
        a[1] = off + nbytes
        a[2] += nbytes
        return a[0][off:off + nbytes].view(dtype)[:n].view(shape)

    def generation(self, device):
        a = self.arenas.get(torch.device(device))
        return -1 if a is None else a[3]


_scratch = _ZeroScratch()


def scratch_generation(device):
    """Changes whenever the zero-scratch arena of `device` is re-allocated (see _ZeroScratch)."""
    return _scratch.generation(device)


def begin_step(device):
    """Call once at the start of a forward pass: recycles the zero scratch of the previous step and refreshes the
    kernel-layout copies of all registered conv weights in one launch."""
    device = torch.device(device)
    _scratch.begin_step(device)
    if BATCH_WEIGHTS and device.type == "cuda":
        c = _wcache.get(device)
        if c is not None:
            c.refresh(_lib.load(), device)


def _conv_work(d, esize):
    """algorithmic bytes (each activation tensor touched once) and FLOPs of one conv launch."""
    cin = d.c0 + d.c1
    vin, vout = d.di * d.hi * d.wi, d.dout * d.ho * d.wo
    nbytes = d.n * (vin * cin + vout * d.cout) * esize
    flops = 2 * d.n * vout * d.ksize ** 3 * cin * d.cout
    key = f"c{cin}->{d.cout} k{d.ksize} s{d.stride} {d.dout}x{d.ho}x{d.wo} n{d.n} g{d.groups}"
    return key, nbytes, flops


def _conv_desc(x0, x1, cout, ksize, stride, pad_mode, groups):
    n, di, hi, wi, c0 = x0.shape
    pad = ksize // 2
    osz = lambda i: (i + 2 * pad - ksize) // stride + 1
    d = ConvDesc(dtype=_dt(x0), n=n, di=di, hi=hi, wi=wi, dout=osz(di), ho=osz(hi), wo=osz(wi), c0=c0,
                 c1=0 if x1 is None else x1.shape[-1], cout=cout, ksize=ksize, stride=stride,
                 pad_mode=PB_PAD_REFLECT if pad_mode == "reflect" else PB_PAD_ZERO, groups=groups)
    return d


# ---- tcgen05 implicit-GEMM path (csrc/conv3d_tc.cu) ------------------------------------------------------
TC_ENABLED = os.environ.get("PB_TC", "1") != "0"
WGRAD_TC = os.environ.get("PB_WGRAD_TC", "1") != "0"
WGRAD1_TC = os.environ.get("PB_WGRAD1_TC", "1") != "0"        # 1x1x1 weight gradient on tcgen05
DGRAD_FOLD = os.environ.get("PB_DGRAD_FOLD", "1") != "0"      # reflect-pad data gradient: extended-domain tc pass + fold
UPSAMPLE_SEPARABLE_FROM = int(os.environ.get("PB_UPS_SEP", "2"))     # trilinear adjoint: three 1-D passes from this scale on
_tc_err = {}


def _tc_err_flag(device):
    t = _tc_err.get(device)
    if t is None:
        t = torch.zeros(1, dtype=torch.int32, device=device)
        _tc_err[device] = t
        _publish_to_all_streams(torch.device(device))
    return t


def check_tc_errors():
    """Synchronises and raises if any tcgen05 conv launch reported an internal pipeline time-out."""
    for dev, t in _tc_err.items():
        code = int(t.item())
        if code:
            t.zero_()
            raise RuntimeError(f"passion_b200: conv3d_tc pipeline time-out (code {code}) on {dev}")


def _tc_ntile(cin, cout):
    return int(_lib.load().pb_conv3d_tc_ntile(cin, cout))


def _tc_eligible(dtype, ksize, stride, c0, c1, cout):
    return (TC_ENABLED and dtype == torch.bfloat16 and ksize == 3 and stride == 1 and c0 % 8 == 0 and c1 % 8 == 0
            and cout % 8 == 0 and _tc_ntile(c0 + c1, cout) != 0)


def _tc1_ntile(cin, cout):
    return int(_lib.load().pb_conv1_tc_ntile(cin, cout))


def _tc1_eligible(dtype, ksize, stride, c0, c1, cout, voxels):
    """1x1x1 conv on the TMA-fed tcgen05 GEMM (csrc/conv1_tc.cu)."""
    return (TC_ENABLED and dtype == torch.bfloat16 and ksize == 1 and stride == 1 and c0 % 8 == 0 and c1 % 8 == 0 and c0 >= 8
            and voxels >= 128 and _tc1_ntile(c0 + c1, cout) != 0)


def tc1_weight_image(w, nt):
    """fp32 [G, 1, cin, cout] -> bf16 image [G, cout tiles, cin/8 rounded up to even, nt, 8] (zero padded)."""
    G, _, cin, cout = w.shape
    nch = (cin // 8 + 1) & ~1
    tiles = (cout + nt - 1) // nt
    img = torch.zeros((G, nch * 8, tiles * nt), dtype=torch.float32, device=w.device)
    img[:, :cin, :cout] = w[:, 0]
    return img.view(G, nch, 8, tiles, nt).permute(0, 3, 1, 4, 2).to(torch.bfloat16).contiguous()


def tc_weight_image(w, nt):
    """fp32 [G, 27, cin, cout] -> bf16 image [G, cout tiles, 9, max(2, cin/8), 3 nt, 8] (zero padded); for the classes of
    the kw-stacked kernel (pb_conv3d_tc_kws) [G, cout tiles, 3 (kh), max(2, cin/8), 9 nt (kd = 2,1,0 | kw | co), 8]."""
    G, taps, cin, cout = w.shape
    nchr = cin // 8
    nch = max(2, nchr)
    tiles = (cout + nt - 1) // nt
    img = torch.zeros((G, taps, nch, 8, tiles * nt), dtype=torch.float32, device=w.device)
    img[:, :, :nchr, :, :cout] = w.reshape(G, taps, nchr, 8, cout)
    if _lib.load().pb_conv3d_tc_kws(cin, cout):
        # [G, kd, kh, kw, chunk, 8, tile, nt] -> [G, tile, kh, chunk, kd (flipped), kw, nt, 8]
        img = img.view(G, 3, 3, 3, nch, 8, tiles, nt).flip(1).permute(0, 6, 2, 4, 1, 3, 7, 5)
        return img.to(torch.bfloat16).reshape(G, tiles, 3, nch, 9 * nt, 8).contiguous()
    # kernel layout: [G][tile][9 (kh,kw)][chunk][3 nt rows, kd = 2,1,0][8 channels]
    img = img.view(G, 3, 9, nch, 8, tiles, nt).flip(1).permute(0, 5, 2, 3, 1, 6, 4)
    return img.to(torch.bfloat16).reshape(G, tiles, 9, nch, 3 * nt, 8).contiguous()


def _tc_conv_call(lib, d, x0, x1, w, y0, y1, co0, co1, stats, err, bias=None):
    """Launch the tensor-core implicit GEMM on kernel-layout weights; returns the C status code."""
    cin, cout = d.c0 + d.c1, co0 + co1
    img = tc_weight_image(w, _tc_ntile(cin, cout))
    return lib.pb_conv3d_tc(ctypes.byref(d), _p(x0), _p(x1), _p(img), _p(bias), _p(y0), _p(y1), co0, co1, _p(stats), _p(err), _stream())


SMALL_ENABLED = os.environ.get("PB_SMALL", "1") != "0"       # shared-memory tiled kernels for the C <= 8 3x3x3 classes


def _small_ok(d):
    return (SMALL_ENABLED and d.ksize == 3 and d.stride == 1 and d.c1 == 0
            and bool(_lib.load().pb_conv3d_small_supported(d.c0, d.cout)))


def _conv_fwd_launch(lib, d, x0, x1, get_wk, tc_call, bias, y, stats):
    """Forward launch: the tcgen05 implicit GEMM when `tc_call` is given (and supports the shape), else the FFMA kernel
    with the fp32 kernel-layout weights returned by get_wk()."""
    key, nb, fl = _conv_work(d, x0.element_size())
    done = False
    if tc_call is not None:
        done = _run("conv1_fwd_tc" if d.ksize == 1 else "conv3d_fwd_tc", key, nb, fl, tc_call, allow_unsupported=True)
    if not done and _small_ok(d):
        wk = get_wk()
        done = _run("conv3d_small_fwd", key, nb, fl,
                    lambda: lib.pb_conv3d_small_fwd(ctypes.byref(d), _p(x0), _p(wk), _p(bias), _p(y), _p(stats), _stream()),
                    allow_unsupported=True)
    if not done:
        wk = get_wk()
        _run("conv1_fwd" if d.ksize == 1 else "conv3d_fwd", key, nb, fl,
             lambda: lib.pb_conv3d_fwd(ctypes.byref(d), _p(x0), _p(x1), _p(wk), _p(bias), _p(y), _p(stats), _stream()))


def _dgrad_tc_ok(d, dtype, ksize, stride, pad_mode):
    return (_tc_eligible(dtype, ksize, stride, d.cout, 0, d.c0 + d.c1) and d.c0 % 8 == 0 and d.c1 % 8 == 0
            and (pad_mode != "reflect" or min(d.di, d.hi, d.wi) >= 4))


def _dgrad_tc1_ok(d, dtype, ksize, stride):
    """data gradient of a 1x1x1 conv on the tcgen05 GEMM: dy has cout channels (multiple of 8), dx = c0 | c1 channels"""
    cin = d.c0 + d.c1
    return (TC_ENABLED and dtype == torch.bfloat16 and ksize == 1 and stride == 1 and d.cout % 8 == 0 and d.di * d.hi * d.wi >= 128
            and (d.c1 == 0 or (d.c0 % 8 == 0 and d.c1 % 8 == 0)) and _tc1_ntile(d.cout, cin) != 0)


def _conv_bwd_launch(lib, d, x0, x1, dy, get_wt, get_imgT, pad_mode, need_dx, need_dw):
    """Data and weight gradient launches.  get_imgT() -> bf16 image of the flipped / transposed weights (None = class
    not on the tensor-core path), get_wt() -> fp32 [G][taps][cout][cin] for the FFMA kernels.
    Returns dx0, dx1, dw (kernel layout [G][taps][cin][cout] fp32)."""
    groups = d.groups
    cin = d.c0 + d.c1
    dx0 = dx1 = dw = None
    key, nb, fl = _conv_work(d, x0.element_size())
    if need_dx:
        dx0 = torch.empty_like(x0)
        dx1 = torch.empty_like(x1) if x1 is not None else None
        done = False
        imgT = get_imgT()
        if imgT is not None and d.ksize == 1:
            # data gradient of a 1x1x1 conv = the same tcgen05 GEMM on dy with the transposed weight image
            err = _tc_err_flag(dy.device)
            dd = ConvDesc(dtype=d.dtype, n=d.n, di=d.di, hi=d.hi, wi=d.wi, dout=d.di, ho=d.hi, wo=d.wi, c0=d.cout, c1=0, cout=cin,
                          ksize=1, stride=1, pad_mode=PB_PAD_ZERO, groups=groups)
            done = _run("conv1_dgrad_tc", key, nb, fl,
                        lambda: lib.pb_conv1_tc(ctypes.byref(dd), _p(dy), None, _p(imgT), None, _p(dx0), _p(dx1), d.c0, d.c1, None,
                                                _p(err), _stream()), allow_unsupported=True)
        elif imgT is not None:
            # data gradient = the same implicit GEMM on dy with flipped taps / transposed channels
            err = _tc_err_flag(dy.device)
            if pad_mode == "reflect" and DGRAD_FOLD:
                # "full" correlation on the domain grown by one voxel, then fold the halo back along the reflections
                dd = ConvDesc(dtype=d.dtype, n=d.n, di=d.di, hi=d.hi, wi=d.wi, dout=d.di + 2, ho=d.hi + 2, wo=d.wi + 2,
                              c0=d.cout, c1=0, cout=cin, ksize=3, stride=1, pad_mode=PB_PAD_ZERO, groups=groups)
                ext = torch.empty((d.n, d.di + 2, d.hi + 2, d.wi + 2, cin), dtype=dy.dtype, device=dy.device)
                done = _run("conv3d_dgrad_tc", key, nb, fl,
                            lambda: lib.pb_conv3d_tc_full(ctypes.byref(dd), _p(dy), _p(imgT), _p(dx0), _p(dx1), d.c0, d.c1,
                                                          _p(ext), _p(err), _stream()), allow_unsupported=True)
                if done:
                    _run("reflect_fold", key, 0, 0,
                         lambda: lib.pb_reflect_fold(_p(ext), _p(dx0), _p(dx1), d.n, d.di, d.hi, d.wi, d.c0, d.c1, _stream()))
            else:
                dd = ConvDesc(dtype=d.dtype, n=d.n, di=d.di, hi=d.hi, wi=d.wi, dout=d.di, ho=d.hi, wo=d.wi, c0=d.cout, c1=0,
                              cout=cin, ksize=3, stride=1, pad_mode=PB_PAD_ZERO, groups=groups)
                done = _run("conv3d_dgrad_tc", key, nb, fl,
                            lambda: lib.pb_conv3d_tc(ctypes.byref(dd), _p(dy), None, _p(imgT), None, _p(dx0), _p(dx1), d.c0, d.c1,
                                                     None, _p(err), _stream()), allow_unsupported=True)
                if done and pad_mode == "reflect":
                    wt = get_wt()
                    _run("conv3d_dgrad_fix", key, 0, 0,
                         lambda: lib.pb_conv3d_dgrad_reflect_fix(ctypes.byref(d), _p(dy), _p(wt), _p(dx0), _p(dx1), _stream()))
        if not done and _small_ok(d) and (pad_mode != "reflect" or min(d.di, d.hi, d.wi) >= 4):
            wt = get_wt()
            ext = (torch.empty((d.n, d.di + 2, d.hi + 2, d.wi + 2, cin), dtype=dy.dtype, device=dy.device)
                   if pad_mode == "reflect" else None)
            done = _run("conv3d_small_dgrad", key, nb, fl,
                        lambda: lib.pb_conv3d_small_dgrad(ctypes.byref(d), _p(dy), _p(wt), _p(dx0), _p(ext), _stream()),
                        allow_unsupported=True)
        if not done:
            wt = get_wt()
            _run("conv1_dgrad" if d.ksize == 1 else "conv3d_dgrad", key, nb, fl,
                 lambda: lib.pb_conv3d_dgrad(ctypes.byref(d), _p(dy), _p(wt), _p(dx0), _p(dx1), _stream()))
    if need_dw:
        dw = _scratch.zeros((groups, d.ksize ** 3, cin, d.cout), torch.float32, dy.device)
        done = False
        if (TC_ENABLED and WGRAD_TC and dy.dtype == torch.bfloat16 and d.ksize == 3 and d.stride == 1 and d.c0 % 8 == 0
                and d.c1 % 8 == 0 and d.cout % 8 == 0 and d.cout <= 64):
            err = _tc_err_flag(dy.device)
            done = _run("conv3d_wgrad_tc", key, nb, fl,
                        lambda: lib.pb_conv3d_wgrad_tc(ctypes.byref(d), _p(x0), _p(x1), _p(dy), _p(dw), _p(err), _stream()),
                        allow_unsupported=True)
        if (not done and TC_ENABLED and WGRAD1_TC and dy.dtype == torch.bfloat16 and d.ksize == 1 and d.stride == 1
                and d.c0 % 8 == 0 and d.c1 % 8 == 0 and d.cout % 8 == 0 and d.cout <= 64 and cin <= 256):
            err = _tc_err_flag(dy.device)
            done = _run("conv1_wgrad_tc", key, nb, fl,
                        lambda: lib.pb_conv1_wgrad_tc(ctypes.byref(d), _p(x0), _p(x1), _p(dy), _p(dw), _p(err), _stream()),
                        allow_unsupported=True)
        if not done and _small_ok(d):
            done = _run("conv3d_small_wgrad", key, nb, fl,
                        lambda: lib.pb_conv3d_small_wgrad(ctypes.byref(d), _p(x0), _p(dy), _p(dw), _stream()),
                        allow_unsupported=True)
        if not done:
            _run("conv1_wgrad" if d.ksize == 1 else "conv3d_wgrad", key, nb, fl,
                 lambda: lib.pb_conv3d_wgrad(ctypes.byref(d), _p(x0), _p(x1), _p(dy), _p(dw), _stream()))
    return dx0, dx1, dw


class _Conv3d(torch.autograd.Function):
    """y = conv(cat(x0, x1), w) (+ bias) with KERNEL-layout weights [groups, taps, cin, cout]; optionally also the
    per-(n,c) sum / sum-of-squares of y.  (The models use _Conv3dRef, which takes the parameters as they are.)"""

    @staticmethod
    def forward(ctx, x0, x1, w, bias, ksize, stride, pad_mode, groups, want_stats):
        lib = _lib.load()
        _chk(x0, x1, w, bias)
        assert w.dtype == torch.float32 and w.dim() == 4, "w must be fp32 [groups, taps, cin, cout]"
        cout = w.shape[-1]
        d = _conv_desc(x0, x1, cout, ksize, stride, pad_mode, groups)
        assert w.shape == (groups, ksize ** 3, d.c0 + d.c1, cout), (tuple(w.shape), groups, ksize, d.c0, d.c1, cout)
        y = torch.empty((d.n, d.dout, d.ho, d.wo, cout), dtype=x0.dtype, device=x0.device)
        stats = _scratch.zeros((d.n, cout, 2), torch.float64, x0.device) if want_stats else None
        tc_call = None
        if _tc_eligible(x0.dtype, ksize, stride, d.c0, d.c1, cout):
            err = _tc_err_flag(x0.device)
            tc_call = lambda: _tc_conv_call(lib, d, x0, x1, w, y, None, cout, 0, stats, err, bias)
        elif _tc1_eligible(x0.dtype, ksize, stride, d.c0, d.c1, cout, d.di * d.hi * d.wi):
            err = _tc_err_flag(x0.device)
            img1 = tc1_weight_image(w, _tc1_ntile(d.c0 + d.c1, cout))
            tc_call = lambda: lib.pb_conv1_tc(ctypes.byref(d), _p(x0), _p(x1), _p(img1), _p(bias), _p(y), None, cout, 0, _p(stats),
                                              _p(err), _stream())
        _conv_fwd_launch(lib, d, x0, x1, lambda: w, tc_call, bias, y, stats)
        ctx.save_for_backward(x0, x1, w)
        ctx.cfg = (ksize, stride, pad_mode, groups, bias is not None)
        ctx.set_materialize_grads(False)         # no zero-filled "gradient" of the statistics output (one fill launch per conv otherwise)
        if want_stats:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return y, None

    @staticmethod
    def backward(ctx, dy, _dstats):
        if dy is None:
            return (None,) * 9
        lib = _lib.load()
        x0, x1, w = ctx.saved_tensors
        ksize, stride, pad_mode, groups, has_bias = ctx.cfg
        dy = dy.contiguous()
        d = _conv_desc(x0, x1, w.shape[-1], ksize, stride, pad_mode, groups)
        need_dx = ctx.needs_input_grad[0] or (x1 is not None and ctx.needs_input_grad[1])

        def get_imgT():
            if _dgrad_tc1_ok(d, dy.dtype, ksize, stride):
                return tc1_weight_image(w.transpose(2, 3), _tc1_ntile(d.cout, d.c0 + d.c1))
            if not _dgrad_tc_ok(d, dy.dtype, ksize, stride, pad_mode):
                return None
            return tc_weight_image(w.flip(1).transpose(2, 3), _tc_ntile(d.cout, d.c0 + d.c1))

        dx0, dx1, dw = _conv_bwd_launch(lib, d, x0, x1, dy, lambda: w.transpose(2, 3).contiguous(), get_imgT, pad_mode,
                                        need_dx, ctx.needs_input_grad[2])
        db = None
        if has_bias and ctx.needs_input_grad[3]:
            db = dy.float().reshape(groups, -1, dy.shape[-1]).sum(1)
        if dw is not None:
            dw = dw.clone()                      # the accumulator is recycled scratch
        return dx0, dx1, dw, db, None, None, None, None, None


BATCH_WEIGHTS = os.environ.get("PB_BATCH_WEIGHTS", "1") != "0"   # one weight-prep / one weight-grad-unpack launch per step


class _WeightCache:
    """Kernel-layout copies of every conv layer's parameters, kept in persistent buffers and refreshed by ONE batched launch
    at the start of a step (`begin_step` -> pb_weight_prep_batch) instead of one launch per layer: the weights change once
    per step, after the optimizer.  An entry is registered the first time a layer runs (that call prepares it alone); it is
    used only while the parameters' version counters still equal those seen by the last refresh, so weights modified in
    place outside the step protocol simply take the per-layer path again.  Entries die with their parameters."""

    def __init__(self):
        self.entries = {}            # key -> entry
        self.dirty = False
        self.table = None            # device scratch holding the packed descriptors of the batched launch
        self.descs = None
        self.n = 0

    @staticmethod
    def _versions(e):
        ps = [r() for r in e["refs"]]
        if any(p is None for p in ps):
            return None
        return tuple(p._version for p in ps)

    def lookup(self, key):
        e = self.entries.get(key)
        if e is None:
            return None
        v = self._versions(e)
        if v is None:
            del self.entries[key]
            self.dirty = True
            return None
        return e if v == e["versions"] else False          # False: registered but stale

    def refresh(self, lib, device):
        dead = [k for k, e in self.entries.items() if self._versions(e) is None]
        for k in dead:
            del self.entries[k]
            self.dirty = True
        if not self.entries:
            return
        capturing = torch.cuda.is_current_stream_capturing()
        if self.dirty:
            if capturing:
                return                                     # cannot upload a new table inside a capture: per-layer path
            ents = list(self.entries.values())
            self.n = len(ents)
            self.descs = (_lib.WeightPrepDesc * self.n)()
            for d, e in zip(self.descs, ents):
                ws, bs = e["ws"](), e["bs"]()
                wk, wt, img, imgT, bias = e["outs"]
                S = e["slices"]
                G = S if S else len(ws)
                d.groups, d.cin, d.cout, d.ksize = G, e["cin"], e["cout"], e["ksize"]
                d.wk, d.wt, d.img, d.imgT, d.bias = (t.data_ptr() if t is not None else None for t in (wk, wt, img, imgT, bias))
                d.nt, d.ntT, d.w_cin_stride = e["nt"], e["ntT"], S * e["cin"]
                for g in range(G):
                    d.w[g] = ws[0].data_ptr() + 4 * g * e["cin"] * e["ksize"] ** 3 if S else ws[g].data_ptr()
                    d.b[g] = bs[g].data_ptr() if bias is not None else None
            need = int(lib.pb_weight_batch_table_bytes(self.n))
            if self.table is None or self.table.numel() < need:
                self.table = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=device)
        upload = 1 if self.dirty else 0
        _run("weight_prep", "batch", 0, 0, lambda: lib.pb_weight_prep_batch(self.descs, self.n, _p(self.table), upload, _stream()))
        self.dirty = False
        for e in self.entries.values():
            e["versions"] = self._versions(e)


_wcache = {}


def _wcache_for(device):
    c = _wcache.get(device)
    if c is None:
        c = _wcache[device] = _WeightCache()
    return c


def _weight_prep(lib, ws, bs, cin, cout, ksize, want_wk=False, want_wt=False, nt=0, ntT=0, want_bias=False, slices=0):
    """Parameter-layout weights of G groups -> the layouts asked for.  Served from the step's batched refresh when the layer
    is registered and fresh, else one pb_weight_prep launch (which also registers the layer).
    slices = S > 0: `ws` is ONE parameter [cout, S*cin, k, k, k] and group g reads its cin-slice g."""
    import weakref
    G = slices if slices else len(ws)
    dev = ws[0].device
    cache = _wcache_for(dev) if BATCH_WEIGHTS else None
    key = (tuple(id(w) for w in ws), bool(want_wk), bool(want_wt), nt, ntT, bool(want_bias), slices)
    e = cache.lookup(key) if cache is not None else None
    if e:
        return e["outs"]
    taps = ksize ** 3
    f32 = dict(dtype=torch.float32, device=dev)
    if e is False:                                                   # registered but stale: refresh this layer in place
        ent = cache.entries[key]
        wk, wt, img, imgT, bias = ent["outs"]
    else:
        ent = None
        wk = torch.empty((G, taps, cin, cout), **f32) if want_wk else None
        wt = torch.empty((G, taps, cout, cin), **f32) if want_wt else None
        if ksize == 1:                                               # csrc/conv1_tc.cu: [G][tiles][cin/8 rounded up to even][nt][8]
            img = torch.empty((G, (cout + nt - 1) // nt, (cin // 8 + 1) & ~1, nt, 8), dtype=torch.bfloat16, device=dev) if nt else None
            imgT = torch.empty((G, (cin + ntT - 1) // ntT, (cout // 8 + 1) & ~1, ntT, 8), dtype=torch.bfloat16, device=dev) if ntT else None
        else:
            img = torch.empty((G, (cout + nt - 1) // nt, 9, max(2, cin // 8), 3 * nt, 8), dtype=torch.bfloat16, device=dev) if nt else None
            imgT = torch.empty((G, (cin + ntT - 1) // ntT, 9, max(2, cout // 8), 3 * ntT, 8), dtype=torch.bfloat16, device=dev) if ntT else None
        bias = torch.empty((G, cout), **f32) if want_bias else None
    desc = _lib.WeightPrepDesc(groups=G, cin=cin, cout=cout, ksize=ksize, wk=_p(wk), wt=_p(wt), img=_p(img), nt=nt,
                               imgT=_p(imgT), ntT=ntT, bias=_p(bias), w_cin_stride=slices * cin)
    for g in range(G):
        desc.w[g] = ws[0].data_ptr() + 4 * g * cin * taps if slices else ws[g].data_ptr()
        desc.b[g] = bs[g].data_ptr() if want_bias else None
    _run("weight_prep", f"c{cin}->{cout} k{ksize} g{G}", 0, 0, lambda: lib.pb_weight_prep(ctypes.byref(desc), _stream()))
    if cache is not None and all(isinstance(w, torch.nn.Parameter) for w in ws):
        params = list(ws) + (list(bs) if want_bias else [])
        if ent is None:
            # weak references only: the cache must not keep a dead model's parameters alive
            wrefs, brefs = [weakref.ref(w) for w in ws], [weakref.ref(b) for b in bs] if want_bias else []
            ent = dict(refs=[weakref.ref(p) for p in params], outs=(wk, wt, img, imgT, bias), cin=cin, cout=cout, ksize=ksize,
                       nt=nt, ntT=ntT, slices=slices, ws=lambda r=wrefs: [x() for x in r], bs=lambda r=brefs: [x() for x in r])
            cache.entries[key] = ent
            cache.dirty = True
        ent["versions"] = tuple(p._version for p in params)
    return wk, wt, img, imgT, bias


def _unpack_desc(item, gws, gbs, accumulate, desc=None):
    desc = desc if desc is not None else _lib.WeightUnpackDesc()
    S = item.get("slices", 0)
    G = S if S else len(item["ws"])
    desc.dw, desc.db = item["dw"].data_ptr(), None
    desc.dy_stats = item["st"].data_ptr() if gbs is not None else None
    desc.npg, desc.groups, desc.cin, desc.cout, desc.ksize = item["npg"], G, item["cin"], item["cout"], item["ksize"]
    desc.accumulate, desc.w_cin_stride = accumulate, S * item["cin"]
    for g in range(4):
        if S:
            desc.gw[g] = gws[0].data_ptr() + 4 * g * item["cin"] * item["ksize"] ** 3 if g < G else None
        else:
            desc.gw[g] = gws[g].data_ptr() if g < G else None
        desc.gb[g] = gbs[g].data_ptr() if (g < G and gbs is not None) else None
    return desc


class _GradSink:
    """Weight gradients of the conv layers of one backward pass, scattered into the parameters' .grad by ONE launch
    (pb_weight_grad_unpack_batch) from an end-of-backward callback, instead of one scatter launch per layer.
    A parameter without a .grad gets a persistent per-parameter buffer (pointer-stable, so the packed descriptor table is
    uploaded only when something changed and the launch is CUDA-graph capturable); an existing .grad (DDP bucket views,
    gradient accumulation) is added to."""

    def __init__(self):
        self.pending = []
        self.bufs = {}               # id(param) -> (weakref, buffer)
        self.launches = []           # per launch ordinal within a flush: dict(sig, table, descs)
        self.ordinal = 0

    def _buffer(self, p):
        import weakref
        ent = self.bufs.get(id(p))
        if ent is None or ent[0]() is not p:
            if len(self.bufs) > 4096:
                self.bufs = {k: v for k, v in self.bufs.items() if v[0]() is not None}
            ent = self.bufs[id(p)] = (weakref.ref(p), torch.empty_like(p, memory_format=torch.contiguous_format))
        return ent[1]

    def _target(self, p):
        g = p.grad
        if g is None:
            p.grad = self._buffer(p)
            return p.grad, 0
        if g.dtype == torch.float32 and g.is_contiguous() and g.shape == p.shape:
            return g, 1
        raise RuntimeError("passion_b200: conv parameter .grad must be a contiguous fp32 tensor of the parameter's shape")

    def flush(self, device):
        items, self.pending = self.pending, []
        if not items:
            return
        lib = _lib.load()
        self.ordinal = 0
        # a parameter used by several conv calls of this backward (the dense and the single-modality passes share weights) is
        # written by its r-th use in the r-th launch: no two rows of one launch touch the same gradient
        rounds, uses = [], {}
        for it in items:
            ps = list(it["ws"]) + list(it["bs"])
            r = max(uses.get(id(p), 0) for p in ps)
            for p in ps:
                uses[id(p)] = r + 1
            while len(rounds) <= r:
                rounds.append([])
            rounds[r].append(it)
        for its in rounds:
            rows = []
            for it in its:
                tg = [self._target(w) for w in it["ws"]]
                tb = [self._target(b) for b in it["bs"]] if it["bs"] else None
                acc = {a for _, a in tg} | ({a for _, a in tb} if tb else set())
                if len(acc) != 1:
                    raise RuntimeError("passion_b200: weight and bias .grad of one conv layer must both exist or both be None")
                rows.append((it, [t for t, _ in tg], [t for t, _ in tb] if tb else None, acc.pop()))
            self._launch(lib, rows, device)

    def _launch(self, lib, rows, device):
        if not rows:
            return
        n = len(rows)
        if self.ordinal >= len(self.launches):
            self.launches.append(dict(sig=None, table=None, descs=None))
        slot = self.launches[self.ordinal]
        self.ordinal += 1
        sig = tuple((it["dw"].data_ptr(), it["st"].data_ptr() if gbs is not None else 0, acc, tuple(t.data_ptr() for t in gws),
                     tuple(t.data_ptr() for t in gbs) if gbs is not None else (), it.get("slices", 0)) for it, gws, gbs, acc in rows)
        changed = sig != slot["sig"]
        if changed and torch.cuda.is_current_stream_capturing():
            for it, gws, gbs, acc in rows:                  # a new table cannot be uploaded inside a capture
                desc = _unpack_desc(it, gws, gbs, acc)
                _run("weight_grad_unpack", "layer", 0, 0, lambda: lib.pb_weight_grad_unpack(ctypes.byref(desc), _stream()))
            return
        if changed:
            slot["descs"] = (_lib.WeightUnpackDesc * n)()
            for d, (it, gws, gbs, acc) in zip(slot["descs"], rows):
                _unpack_desc(it, gws, gbs, acc, d)
            need = int(lib.pb_weight_batch_table_bytes(n))
            if slot["table"] is None or slot["table"].numel() < need:
                slot["table"] = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=device)
            slot["sig"] = sig
        _run("weight_grad_unpack", "batch", 0, 0,
             lambda: lib.pb_weight_grad_unpack_batch(slot["descs"], n, _p(slot["table"]), 1 if changed else 0, _stream()))


_sinks = {}


def _queue_weight_grads(device, item):
    sink = _sinks.get(device)
    if sink is None:
        sink = _sinks[device] = _GradSink()
    if not sink.pending:
        torch.autograd.Variable._execution_engine.queue_callback(lambda: flush_weight_grads(device))
    sink.pending.append(item)


def flush_weight_grads(device):
    """Runs at the end of the backward pass that queued conv weight gradients (autograd final callback, on the caller's
    stream, after the engine has joined the streams the backward ran on)."""
    sink = _sinks.get(device)
    if sink is not None and sink.pending:
        for st in _side_streams_of(device):
            torch.cuda.current_stream(device).wait_stream(st)
        sink.flush(device)


_side_stream_registry = {}


def register_side_stream(device, stream):
    """Streams other than the caller's on which conv backward kernels may run (the models' decoder_sep stream)."""
    _side_stream_registry.setdefault(device, [])
    if stream not in _side_stream_registry[device]:
        _side_stream_registry[device].append(stream)


def _side_streams_of(device):
    return _side_stream_registry.get(device, [])


class _Conv3dRef(torch.autograd.Function):
    """Same op as _Conv3d, but on the PARAMETERS as the reference stores them: G weights [cout, cin, k, k, k] (one per
    weight group) and optionally G biases [cout].  All layout work is one gather launch in forward (pb_weight_prep)
    and one scatter launch in backward (pb_weight_grad_unpack) instead of ~10 tensor-op launches per layer."""

    @staticmethod
    def forward(ctx, x0, x1, ksize, stride, pad_mode, want_stats, bias_slices, *params):
        lib = _lib.load()
        has_bias, slices = bias_slices
        nw = len(params) // 2 if has_bias else len(params)               # weight PARAMETERS
        ws, bs = params[:nw], params[nw:]
        G = slices if slices else nw                                     # weight GROUPS over the modality-major batch
        _chk(x0, x1, *params)
        cout, cin = ws[0].shape[0], ws[0].shape[1] // (slices or 1)
        d = _conv_desc(x0, x1, cout, ksize, stride, pad_mode, G)
        assert cin == d.c0 + d.c1 and all(w.dtype == torch.float32 and tuple(w.shape) == (cout, cin * (slices or 1), ksize, ksize, ksize) for w in ws)
        assert not slices or (nw == 1 and not has_bias), "sliced weights: one parameter, no bias"
        need_dx = ctx.needs_input_grad[0] or (x1 is not None and ctx.needs_input_grad[1])
        fwd_tc = _tc_eligible(x0.dtype, ksize, stride, d.c0, d.c1, cout)
        fwd_tc1 = _tc1_eligible(x0.dtype, ksize, stride, d.c0, d.c1, cout, d.di * d.hi * d.wi)
        dgrad_tc1 = need_dx and _dgrad_tc1_ok(d, x0.dtype, ksize, stride)
        dgrad_tc = need_dx and (dgrad_tc1 or _dgrad_tc_ok(d, x0.dtype, ksize, stride, pad_mode))
        want_wt = need_dx and (not dgrad_tc or (pad_mode == "reflect" and not DGRAD_FOLD and not dgrad_tc1))
        nt = _tc_ntile(cin, cout) if fwd_tc else (_tc1_ntile(cin, cout) if fwd_tc1 else 0)
        ntT = (_tc1_ntile(cout, cin) if dgrad_tc1 else _tc_ntile(cout, cin)) if dgrad_tc else 0
        wk, wt, img, imgT, bias = _weight_prep(lib, ws, bs, cin, cout, ksize, want_wk=not (fwd_tc or fwd_tc1), want_wt=want_wt,
                                               nt=nt, ntT=ntT, want_bias=has_bias, slices=slices)
        y = torch.empty((d.n, d.dout, d.ho, d.wo, cout), dtype=x0.dtype, device=x0.device)
        stats = _scratch.zeros((d.n, cout, 2), torch.float64, x0.device) if want_stats else None
        tc_call = None
        if fwd_tc:
            err = _tc_err_flag(x0.device)
            tc_call = lambda: lib.pb_conv3d_tc(ctypes.byref(d), _p(x0), _p(x1), _p(img), _p(bias), _p(y), None, cout, 0,
                                               _p(stats), _p(err), _stream())
        elif fwd_tc1:
            err = _tc_err_flag(x0.device)
            tc_call = lambda: lib.pb_conv1_tc(ctypes.byref(d), _p(x0), _p(x1), _p(img), _p(bias), _p(y), None, cout, 0,
                                              _p(stats), _p(err), _stream())
        _conv_fwd_launch(lib, d, x0, x1, lambda: wk if wk is not None else _weight_prep(lib, ws, bs, cin, cout, ksize, want_wk=True, slices=slices)[0],
                         tc_call, bias, y, stats)
        ctx.save_for_backward(x0, x1, *ws, *bs)
        ctx.bwd_w = (wt, imgT)
        ctx.cfg = (ksize, stride, pad_mode, G, has_bias, nw, slices)
        ctx.set_materialize_grads(False)         # no zero-filled "gradient" of the statistics output (one fill launch per conv otherwise)
        if want_stats:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return y, None

    @staticmethod
    def backward(ctx, dy, _dstats):
        if dy is None:
            return (None,) * (7 + len(ctx.saved_tensors) - 2)
        lib = _lib.load()
        ksize, stride, pad_mode, G, has_bias, nw, slices = ctx.cfg
        x0, x1 = ctx.saved_tensors[:2]
        ws = ctx.saved_tensors[2:2 + nw]
        wt, imgT = ctx.bwd_w
        cout, cin = ws[0].shape[0], ws[0].shape[1] // (slices or 1)
        dy = dy.contiguous()
        d = _conv_desc(x0, x1, cout, ksize, stride, pad_mode, G)
        need_dx = ctx.needs_input_grad[0] or (x1 is not None and ctx.needs_input_grad[1])
        need_dw = any(ctx.needs_input_grad[7:7 + nw])
        need_db = has_bias and any(ctx.needs_input_grad[7 + nw:])
        get_wt = lambda: wt if wt is not None else _weight_prep(lib, ws, (), cin, cout, ksize, want_wt=True, slices=slices)[1]
        dx0, dx1, dw = _conv_bwd_launch(lib, d, x0, x1, dy, get_wt, lambda: imgT, pad_mode, need_dx, need_dw or need_db)
        gws = [None] * nw
        gbs = [None] * nw if has_bias else []
        if need_dw or need_db:
            bs = ctx.saved_tensors[2 + nw:] if has_bias else ()
            st = channel_stats(dy) if need_db else None
            item = dict(ws=ws, bs=bs if need_db else (), dw=dw, st=st, npg=d.n // G, cin=cin, cout=cout, ksize=ksize, slices=slices)
            if (BATCH_WEIGHTS and need_dw and all(w.is_leaf and w.requires_grad for w in ws)
                    and (not has_bias or (need_db and all(b.is_leaf and b.requires_grad for b in bs)))):
                # leaf parameters: their .grad is written by ONE batched scatter launch when this backward pass ends
                _queue_weight_grads(dy.device, item)
            else:
                gws = [torch.empty_like(w) for w in ws]
                if need_db:
                    gbs = [torch.empty((cout,), dtype=torch.float32, device=dy.device) for _ in range(nw)]
                desc = _unpack_desc(item, gws, gbs if need_db else None, 0)
                _run("weight_grad_unpack", f"c{cin}->{cout} k{ksize} g{G}", 0, 0,
                     lambda: lib.pb_weight_grad_unpack(ctypes.byref(desc), _stream()))
        return (dx0, dx1, None, None, None, None, None, *gws, *gbs)


def conv3d_ref(x0, weights, biases=None, x1=None, ksize=3, stride=1, pad_mode="reflect", want_stats=False, slices=0):
    """Conv on the parameters in nn.Conv3d layout: `weights` = list of G tensors [cout, cin, k, k, k] (G weight groups
    over a modality-major batch), `biases` = list of G tensors [cout] or None.  Returns (y, stats or None).
    slices = S: `weights` is ONE tensor [cout, S*cin, k, k, k]; group g of the batch is convolved with its cin-slice g (the
    single-modality decoder passes of a conv over the 4-modality stack: only modality g's quarter of the weight matters)."""
    params = list(weights) + (list(biases) if biases is not None else [])
    return _Conv3dRef.apply(x0, x1, ksize, stride, pad_mode, want_stats, (biases is not None, int(slices)), *params)


def conv3d(x0, w, bias=None, x1=None, ksize=3, stride=1, pad_mode="reflect", groups=1, want_stats=False):
    """w: fp32 kernel layout [groups, ksize^3, cin, cout]; bias: fp32 [groups, cout] or None."""
    return _Conv3d.apply(x0, x1, w, bias, ksize, stride, pad_mode, groups, want_stats)


def kernel_layout(w):
    """reference Conv3d weight [cout, cin, kd, kh, kw] -> [taps, cin, cout] (differentiable)."""
    cout, cin = w.shape[:2]
    return w.permute(2, 3, 4, 1, 0).reshape(-1, cin, cout).contiguous()


def inorm_finalize(stats, voxels, eps=IN_EPS):
    lib = _lib.load()
    n, c, _ = stats.shape
    mr = torch.empty((n, c, 2), dtype=torch.float32, device=stats.device)
    _lib.check(lib.pb_inorm_finalize(_p(stats), _p(mr), n, c, voxels, eps, _stream()), "inorm_finalize")
    return mr


class _InormLrelu(torch.autograd.Function):
    """out = LeakyReLU(InstanceNorm(y)) (+ res).  `mr` = (mean, rstd) of y — or the float64 [N,C,2] sums of y, in which case
    the finalisation runs inside the apply kernel (one launch instead of two); backward is the full InstanceNorm adjoint
    (the dependence of the statistics on y is accounted for)."""

    @staticmethod
    def forward(ctx, y, mr, res):
        lib = _lib.load()
        _chk(y, mr, res)
        n, c = y.shape[0], y.shape[-1]
        voxels = y.numel() // (n * c)
        out = torch.empty_like(y)
        nb = y.numel() * y.element_size() * (3 if res is not None else 2)
        if mr.dtype == torch.float64:
            stats = mr
            mr = torch.empty((n, c, 2), dtype=torch.float32, device=y.device)
            _run("inorm_lrelu_fwd", f"c{c}", nb, 0,
                 lambda: lib.pb_inorm_lrelu_fwd_stats(_dt(y), _p(y), _p(stats), _p(mr), _p(res), _p(out), n, voxels, c, IN_EPS,
                                                      LRELU_SLOPE, _stream()))
        else:
            _run("inorm_lrelu_fwd", f"c{c}", nb, 0,
                 lambda: lib.pb_inorm_lrelu_fwd(_dt(y), _p(y), _p(mr), _p(res), _p(out), n, voxels, c, LRELU_SLOPE, _stream()))
        ctx.save_for_backward(y, mr)
        ctx.has_res = res is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        y, mr = ctx.saved_tensors
        dout = dout.contiguous()
        n, c = y.shape[0], y.shape[-1]
        voxels = y.numel() // (n * c)
        sums = _scratch.zeros((n, c, 2), torch.float64, y.device)
        dy = torch.empty_like(y)
        _run("inorm_lrelu_bwd", f"c{c}", y.numel() * y.element_size() * 3, 0,
             lambda: lib.pb_inorm_lrelu_bwd(_dt(y), _p(dout), _p(y), _p(mr), _p(sums), _p(dy), n, voxels, c, LRELU_SLOPE,
                                            _stream()))
        return dy, None, (dout if ctx.has_res else None)


def conv_in_lrelu(x0, w, x1=None, ksize=3, stride=1, pad_mode="reflect", groups=1, res=None):
    """general_conv3d (reference models/blocks.py:354-370): conv -> InstanceNorm -> LeakyReLU(0.2) (+ res).
    The conv bias is dropped: InstanceNorm(affine=False) cancels it exactly."""
    y, stats = conv3d(x0, w, None, x1, ksize, stride, pad_mode, groups, True)
    return _InormLrelu.apply(y, stats, res)                 # float64 sums: finalised inside the apply kernel


def conv_in_lrelu_ref(x0, weights, x1=None, ksize=3, stride=1, pad_mode="reflect", res=None, slices=0):
    """conv_in_lrelu on parameter-layout weights (list of G tensors, see conv3d_ref)."""
    y, stats = conv3d_ref(x0, weights, None, x1, ksize, stride, pad_mode, True, slices)
    return _InormLrelu.apply(y, stats, res)                 # float64 sums: finalised inside the apply kernel


_zero_arena = {}


def zero_grad_like(shape, dtype, device):
    """A read-only all-zero tensor of the given shape without a launch: a view into a persistent zero buffer (used as the
    gradient of parameters that provably do not influence the output)."""
    n = 1
    for v in shape:
        n *= v
    buf = _zero_arena.get((dtype, device))
    if buf is None or buf.numel() < n:
        buf = torch.zeros(max(n, 1 << 16), dtype=dtype, device=device)
        _zero_arena[(dtype, device)] = buf
        _publish_to_all_streams(torch.device(device))
    return buf[:n].view(shape)


def channel_stats(x):
    """float64 [N, C, 2] (sum, sum of squares over the voxels) of a cl tensor; not differentiable by itself — the
    InstanceNorm adjoint in _InormLrelu accounts for the dependence of the statistics on x."""
    lib = _lib.load()
    x = x.detach()
    _chk(x)
    n, c = x.shape[0], x.shape[-1]
    voxels = x.numel() // (n * c)
    stats = _scratch.zeros((n, c, 2), torch.float64, x.device)
    _run("channel_stats", f"c{c}", x.numel() * x.element_size(), 0,
         lambda: lib.pb_channel_stats(_dt(x), _p(x), _p(stats), n, voxels, c, _stream()))
    return stats


def prenorm(x, stats=None):
    """InstanceNorm3d(affine=False) -> LeakyReLU(0.2) of an arbitrary cl tensor (the first half of
    general_conv3d_prenorm, reference models/blocks.py:312-314).  `stats`: the float64 [N, C, 2] sums of x when the kernel
    that produced x already accumulated them in its epilogue (conv3d_ref(..., want_stats=True)); else one extra pass."""
    return _InormLrelu.apply(x, stats if stats is not None else channel_stats(x), None)


class _Upsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        lib = _lib.load()
        _chk(x)
        n, d, h, w, c = x.shape
        y = torch.empty((n, d * scale, h * scale, w * scale, c), dtype=x.dtype, device=x.device)
        _run("upsample_fwd", f"c{c} x{scale}", (x.numel() + y.numel()) * x.element_size(), 0,
             lambda: lib.pb_upsample_fwd(_dt(x), _p(x), _p(y), n, d, h, w, c, scale, _stream()))
        ctx.shape, ctx.scale = (n, d, h, w, c), scale
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        dy = dy.contiguous()
        n, d, h, w, c = ctx.shape
        s = ctx.scale
        dx = torch.empty((n, d, h, w, c), dtype=dy.dtype, device=dy.device)
        if s < UPSAMPLE_SEPARABLE_FROM:
            _run("upsample_bwd", f"c{c} x{s}", (dx.numel() + dy.numel()) * dy.element_size(), 0,
                 lambda: lib.pb_upsample_bwd(_dt(dy), _p(dy), _p(dx), n, d, h, w, c, s, _stream()))
            return dx, None
        # separable adjoint (W, then H, then D): the intermediates shrink by `s` after every pass
        t1 = torch.empty((n, d * s, h * s, w, c), dtype=dy.dtype, device=dy.device)
        t2 = torch.empty((n, d * s, h, w, c), dtype=dy.dtype, device=dy.device)
        nb = (dy.numel() + 2 * t1.numel() + 2 * t2.numel() + dx.numel()) * dy.element_size()
        _run("upsample_bwd", f"c{c} x{s} W", nb, 0,
             lambda: lib.pb_upsample_bwd_axis(_dt(dy), _p(dy), _p(t1), n * d * s * h * s, w * s, w, c, c, _stream()))
        _run("upsample_bwd", f"c{c} x{s} H", 0, 0,
             lambda: lib.pb_upsample_bwd_axis(_dt(dy), _p(t1), _p(t2), n * d * s, h * s, h, w * c, c, _stream()))
        _run("upsample_bwd", f"c{c} x{s} D", 0, 0,
             lambda: lib.pb_upsample_bwd_axis(_dt(dy), _p(t2), _p(dx), n, d * s, d, h * w * c, c, _stream()))
        return dx, None


def upsample(x, scale=2):
    """nn.Upsample(scale_factor=scale, mode='trilinear', align_corners=True) on a cl tensor."""
    return _Upsample.apply(x, scale)


class _MaskedStack(torch.autograd.Function):
    """out[p*B+b, ..., m*C+c] = enc[m*B+b, ..., c] * ms[p, b, m]  (MaskModal for P decoder passes in one launch)."""

    @staticmethod
    def forward(ctx, enc, ms):
        lib = _lib.load()
        ms = ms.to(torch.float32).contiguous()
        _chk(enc, ms)
        P, B = ms.shape[:2]
        assert ms.shape[2] == 4 and enc.shape[0] == 4 * B
        C = enc.shape[-1]
        V = enc.numel() // (4 * B * C)
        out = torch.empty((P * B,) + tuple(enc.shape[1:-1]) + (4 * C,), dtype=enc.dtype, device=enc.device)
        nb = (enc.numel() + out.numel()) * enc.element_size()
        _run("masked_stack_fwd", f"c{C} p{P}", nb, 0,
             lambda: lib.pb_masked_stack_fwd(_dt(enc), _p(enc), _p(ms), _p(out), P, B, V, C, _stream()))
        ctx.save_for_backward(ms)
        ctx.meta = (tuple(enc.shape), P, B, V, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        (ms,) = ctx.saved_tensors
        shape, P, B, V, C = ctx.meta
        dout = dout.contiguous()
        denc = torch.empty(shape, dtype=dout.dtype, device=dout.device)
        nb = (denc.numel() + dout.numel()) * dout.element_size()
        _run("masked_stack_bwd", f"c{C} p{P}", nb, 0,
             lambda: lib.pb_masked_stack_bwd(_dt(dout), _p(dout), _p(ms), _p(denc), P, B, V, C, _stream()))
        return denc, None


def masked_stack(enc, ms):
    """enc [4B,D,H,W,C] (modality-major), ms [P,B,4] -> [P*B,D,H,W,4C]."""
    return _MaskedStack.apply(enc, ms)


def _ptr_array(ts):
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


class _RfmRegion(torch.autograd.Function):
    """Region-aware modality mixing: y [N,D,H,W,K*C] (masked features, channel = k*C+c), p [N,D,H,W,4] fp32
    (detached class probabilities) -> R [N,D,H,W,4C] (channel = class*C + c).
    params = the 16 modal_fusion gate-MLP parameters as the reference stores them: 4 x weight_layer[0].weight [128,4C+1,1,1,1],
    4 x weight_layer[0].bias [128], 4 x weight_layer[2].weight [4,128,1,1,1], 4 x weight_layer[2].bias [4] (class-major).
    K = 4: the 4-modality stack.  K = 1: y is the modality-major encoder output [4B,...,C] itself — four single-modality
    passes (pass m sees modality m only) without ever building their 3/4-zero stacks.
    Three launches forward (pool, gate MLP, mix) and three backward (gate gradient, gate-MLP adjoint, dy)."""

    @staticmethod
    def forward(ctx, y, p, K, B, *params):
        lib = _lib.load()
        _chk(y, p, *params)
        assert p.dtype == torch.float32 and p.shape[-1] == 4 and K in (1, 4) and len(params) == 16
        n, kc = y.shape[0], y.shape[-1]
        c = kc // K
        voxels = y.numel() // (n * kc)
        dev = y.device
        sums = torch.zeros((n * 4 * kc + n * 4,), dtype=torch.float64, device=dev)       # S | Psum in one zero fill
        S, Ps = sums[:n * 4 * kc].view(n, 4, kc), sums[n * 4 * kc:].view(n, 4)
        _lib.check(lib.pb_rfm_pool(_dt(y), _p(y), _p(p), _p(S), _p(Ps), n, voxels, kc, _stream()), "rfm_pool")
        z1 = torch.empty((n, 4, 128), dtype=torch.float32, device=dev)
        gate = torch.empty((n, 4, K), dtype=torch.float32, device=dev)
        _lib.check(lib.pb_rfm_gate_fwd(_ptr_array(params), _p(S), _p(Ps), _p(z1), _p(gate), n, B, voxels, K, c, _stream()), "rfm_gate_fwd")
        r = torch.empty(y.shape[:-1] + (4 * c,), dtype=y.dtype, device=dev)
        _lib.check(lib.pb_rfm_mix(_dt(y), _p(y), _p(p), _p(gate), _p(r), n, voxels, K, c, _stream()), "rfm_mix")
        ctx.save_for_backward(y, p, S, Ps, z1, gate, *params)
        ctx.meta = (K, B)
        return r

    @staticmethod
    def backward(ctx, dr):
        lib = _lib.load()
        y, p, S, Ps, z1, gate = ctx.saved_tensors[:6]
        params = ctx.saved_tensors[6:]
        K, B = ctx.meta
        dr = dr.contiguous()
        n, kc = y.shape[0], y.shape[-1]
        c = kc // K
        voxels = y.numel() // (n * kc)
        dev = y.device
        dgate = torch.zeros((n, 4, K), dtype=torch.float64, device=dev)
        _lib.check(lib.pb_rfm_mix_bwd_gate(_dt(y), _p(y), _p(p), _p(dr), _p(dgate), n, voxels, K, c, _stream()),
                   "rfm_mix_bwd_gate")
        flat = torch.zeros((sum(t.numel() for t in params),), dtype=torch.float32, device=dev)   # 16 gradient accumulators, one fill
        grads, off = [], 0
        for t in params:
            grads.append(flat[off:off + t.numel()].view(t.shape))
            off += t.numel()
        dS = torch.empty((n, 4, kc), dtype=torch.float32, device=dev)
        _lib.check(lib.pb_rfm_gate_bwd(_ptr_array(params), _ptr_array(grads), _p(S), _p(Ps), _p(z1), _p(dgate), _p(dS), n, B, voxels, K, c,
                                       _stream()), "rfm_gate_bwd")
        dy = torch.empty_like(y)
        _lib.check(lib.pb_rfm_bwd_y(_dt(y), _p(p), _p(gate), _p(dr), _p(dS), _p(dy), n, voxels, K, c, _stream()),
                   "rfm_bwd_y")
        return (dy, None, None, None, *grads)


def rfm_region(y, p, params, B=1):
    """params: the 16 gate-MLP parameters (see _RfmRegion)."""
    return _RfmRegion.apply(y, p, 4, B, *params)


def rfm_region_single(enc, p, params, B):
    """Four single-modality passes on the modality-major encoder output enc [4B,D,H,W,C]; p [4B,D,H,W,4]."""
    return _RfmRegion.apply(enc, p, 1, B, *params)


# ---- PASSION objective kernels (csrc/loss.cu) ----------------------------------------------------------------
class _Softmax4(torch.autograd.Function):
    """probs[..., 4] (fp32) = softmax(logits[..., 4] / temp)."""

    @staticmethod
    def forward(ctx, logits, temp):
        lib = _lib.load()
        _chk(logits)
        assert logits.shape[-1] == 4
        rows = logits.numel() // 4
        probs = torch.empty(logits.shape, dtype=torch.float32, device=logits.device)
        _run("softmax4", "", logits.numel() * (logits.element_size() + 4), 0,
             lambda: lib.pb_softmax4(_dt(logits), _p(logits), _p(probs), rows, 1.0 / temp, _stream()))
        ctx.save_for_backward(probs)
        ctx.meta = (logits.dtype, temp)
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        lib = _lib.load()
        probs, = ctx.saved_tensors
        dtype, temp = ctx.meta
        dprobs = dprobs.contiguous().float()
        dlogits = torch.empty(probs.shape, dtype=dtype, device=probs.device)
        rows = probs.numel() // 4
        _run("softmax4_bwd", "", probs.numel() * (8 + dlogits.element_size()), 0,
             lambda: lib.pb_softmax4_bwd(_dt(dlogits), _p(probs), _p(dprobs), _p(dlogits), rows, 1.0 / temp, _stream()))
        return dlogits, None


def softmax4(logits, temp=1.0):
    return _Softmax4.apply(logits, float(temp))


class _LogitLoss(torch.autograd.Function):
    """Fused logit-level loss pass (csrc/loss.cu logit_loss_*): logits [passes*B, D, H, W, 4], labels uint8 [B, D, H, W].
    mode 0: pass 0 supervised (CE / Dice sums, optionally its probabilities), passes 1.. distilled from pass 0 (KL sums at
    temperature `temp`; the teacher is detached as in the reference).  mode 1: every pass supervised.
    Returns (ce_sums [n,3,4] fp32: A, L, E per supervised sample; kl_sums [(passes-1)*B] fp32; probs [B,D,H,W,4] fp32 or None)."""

    @staticmethod
    def forward(ctx, logits, labels, passes, mode, temp, want_probs):
        lib = _lib.load()
        _chk(logits, labels)
        B = labels.shape[0]
        voxels = labels.numel() // B
        assert logits.shape[-1] == 4 and logits.numel() == passes * B * voxels * 4 and labels.dtype == torch.uint8
        dev = logits.device
        n_ce = B if mode == 0 else passes * B
        ce = torch.zeros((n_ce, 12), dtype=torch.float64, device=dev)
        kl = torch.zeros(((passes - 1) * B,), dtype=torch.float64, device=dev) if mode == 0 and passes > 1 else None
        probs = torch.empty(tuple(labels.shape) + (4,), dtype=torch.float32, device=dev) if want_probs else None
        nb = logits.numel() * logits.element_size() + labels.numel() + (probs.numel() * 4 if want_probs else 0)
        _run("logit_loss_fwd", f"p{passes} m{mode}", nb, 0,
             lambda: lib.pb_logit_loss_fwd(_dt(logits), _p(logits), _p(labels), _p(probs), _p(ce), _p(kl), passes, B, voxels, mode,
                                           1.0 / temp, _stream()))
        ctx.save_for_backward(logits, labels)
        ctx.set_materialize_grads(False)          # the gradient of an unused output arrives as None instead of a dense zero tensor
        ctx.meta = (passes, mode, temp, B, voxels, n_ce)
        klf = kl.float() if kl is not None else torch.zeros((0,), dtype=torch.float32, device=dev)
        if probs is None:
            probs = torch.zeros((0,), dtype=torch.float32, device=dev)
        return ce.float().view(n_ce, 3, 4), klf, probs

    @staticmethod
    def backward(ctx, dce, dkl, dprobs):
        lib = _lib.load()
        logits, labels = ctx.saved_tensors
        passes, mode, temp, B, voxels, n_ce = ctx.meta
        dev = logits.device
        ce_coef = (dce.reshape(n_ce, 12).contiguous().float() if dce is not None
                   else torch.zeros((n_ce, 12), dtype=torch.float32, device=dev))
        kl_coef = None
        if mode == 0 and passes > 1:
            kl_coef = dkl.contiguous().float() if dkl is not None else torch.zeros(((passes - 1) * B,), dtype=torch.float32, device=dev)
        # the returned probabilities normally only leave the graph (Model.forward's first output; the step's CE / Dice of it
        # come from this same pass): a gradient arrives only when a caller differentiates through them
        dp = dprobs.contiguous().float() if dprobs is not None and dprobs.numel() == B * voxels * 4 else None
        dlogits = torch.empty_like(logits)
        _run("logit_loss_bwd", f"p{passes} m{mode}", 2 * logits.numel() * logits.element_size() + labels.numel(), 0,
             lambda: lib.pb_logit_loss_bwd(_dt(logits), _p(logits), _p(labels), _p(ce_coef), _p(kl_coef), _p(dp), _p(dlogits), passes, B,
                                           voxels, mode, 1.0 / temp, _stream()))
        return dlogits, None, None, None, None, None


def logit_loss(logits, labels, passes, mode, temp=1.0, want_probs=False):
    ce, kl, probs = _LogitLoss.apply(logits, labels, passes, mode, float(temp), want_probs)
    return ce, kl, (probs if want_probs else None)


class _CeDiceSums(torch.autograd.Function):
    """probs [N, V.., 4] fp32 at label resolution, labels uint8 [B, V..] -> sums [N, 3, 4] fp32:
    A_c = sum p_c t_c, L_c = sum p_c, E_c = sum t_c log(clamp(p_c, .005, 1)).  Sample n uses labels n % B."""

    @staticmethod
    def forward(ctx, probs, labels):
        lib = _lib.load()
        _chk(probs, labels)
        n, b = probs.shape[0], labels.shape[0]
        voxels = labels.numel() // b
        assert probs.dtype == torch.float32 and probs.numel() == n * voxels * 4 and labels.dtype == torch.uint8
        sums = torch.zeros((n, 12), dtype=torch.float64, device=probs.device)
        _run("cedice_fwd", "", probs.numel() * 4 + n * voxels, 0,
             lambda: lib.pb_cedice_fwd(_p(probs), _p(labels), _p(sums), n, b, voxels, _stream()))
        ctx.save_for_backward(probs, labels)
        return sums.float().view(n, 3, 4)

    @staticmethod
    def backward(ctx, dsums):
        lib = _lib.load()
        probs, labels = ctx.saved_tensors
        n, b = probs.shape[0], labels.shape[0]
        voxels = labels.numel() // b
        coef = dsums.reshape(n, 12).contiguous().float()
        dprobs = torch.empty_like(probs)
        _run("cedice_bwd", "", probs.numel() * 8 + n * voxels, 0,
             lambda: lib.pb_cedice_bwd(_p(probs), _p(labels), _p(coef), _p(dprobs), n, b, voxels, _stream()))
        return dprobs, None


def cedice_sums(probs, labels):
    return _CeDiceSums.apply(probs, labels)


class _KlSums(torch.autograd.Function):
    """ps [N, V.., 4], pt [B, V.., 4] fp32 (already at temperature, same resolution) -> [N] sums of
    clamp(pt) (log clamp(pt) - log clamp(ps)); no gradient to the teacher."""

    @staticmethod
    def forward(ctx, ps, pt):
        lib = _lib.load()
        _chk(ps, pt)
        n, b = ps.shape[0], pt.shape[0]
        voxels = pt.numel() // (b * 4)
        assert ps.numel() == n * voxels * 4
        sums = torch.zeros(n, dtype=torch.float64, device=ps.device)
        _run("kl_fwd", "", (ps.numel() + pt.numel()) * 4, 0,
             lambda: lib.pb_kl_fwd(_p(ps), _p(pt), _p(sums), n, b, voxels, _stream()))
        ctx.save_for_backward(ps, pt)
        return sums.float()

    @staticmethod
    def backward(ctx, dsums):
        lib = _lib.load()
        ps, pt = ctx.saved_tensors
        n, b = ps.shape[0], pt.shape[0]
        voxels = pt.numel() // (b * 4)
        coef = dsums.contiguous().float()
        dps = torch.empty_like(ps)
        _run("kl_bwd", "", ps.numel() * 12, 0,
             lambda: lib.pb_kl_bwd(_p(ps), _p(pt), _p(coef), _p(dps), n, b, voxels, _stream()))
        return dps, None


def kl_sums(ps, pt):
    return _KlSums.apply(ps, pt.detach())


class _ProtoSums(torch.autograd.Function):
    """fs [N, V, 8] student features, ft [B, V, 8] teacher features (detached), labels uint8 [B, V], cnt [B, 4] class
    voxel counts -> (sum d^2 [N], sum |d| [N]) over the classes present in every sample of the local batch."""

    @staticmethod
    def forward(ctx, fs, ft, labels, cnt, eps):
        lib = _lib.load()
        _chk(fs, ft, labels)
        n, b, c = fs.shape[0], ft.shape[0], fs.shape[-1]
        voxels = labels.numel() // b
        dev = fs.device
        Ps = torch.zeros((n, 4, c), dtype=torch.float64, device=dev)
        Pt = torch.zeros((b, 4, c), dtype=torch.float64, device=dev)
        dt = _dt(fs)
        _run("proto_sums", "s", fs.numel() * fs.element_size(), 0,
             lambda: lib.pb_proto_sums(dt, _p(fs), _p(labels), _p(Ps), n, b, voxels, c, _stream()))
        _run("proto_sums", "t", ft.numel() * ft.element_size(), 0,
             lambda: lib.pb_proto_sums(dt, _p(ft), _p(labels), _p(Pt), b, b, voxels, c, _stream()))
        den = (cnt.float() + eps)                                         # [B,4]   (criterions.py:158-159)
        present = (cnt > 0).all(0).float().contiguous()                   # class used iff present in EVERY sample (:157)
        protos = (Ps.float() / den.repeat(n // b, 1)[..., None]).contiguous()
        protot = (Pt.float() / den[..., None]).contiguous()
        out = torch.zeros((n, 2), dtype=torch.float64, device=dev)
        _run("proto_fwd", "", (fs.numel() + fs.numel()) * fs.element_size(), 0,
             lambda: lib.pb_proto_fwd(dt, _p(fs), _p(ft), _p(protos), _p(protot), _p(present), _p(out), n, b, voxels, c, eps,
                                      _stream()))
        ctx.save_for_backward(fs, ft, labels, protos, protot, present, den)
        ctx.eps = eps
        out = out.float()
        ctx.mark_non_differentiable(present)
        return out[:, 0].contiguous(), out[:, 1].contiguous(), present

    @staticmethod
    def backward(ctx, dse, _dab, _dpresent):
        lib = _lib.load()
        fs, ft, labels, protos, protot, present, den = ctx.saved_tensors
        n, b, c = fs.shape[0], ft.shape[0], fs.shape[-1]
        voxels = labels.numel() // b
        coef = dse.contiguous().float()
        dfs = torch.empty_like(fs)
        dP = torch.zeros((n, 4, c), dtype=torch.float64, device=fs.device)
        dt = _dt(fs)
        _run("proto_bwd1", "", fs.numel() * fs.element_size() * 3, 0,
             lambda: lib.pb_proto_bwd1(dt, _p(fs), _p(ft), _p(protos), _p(protot), _p(present), _p(coef), _p(dfs), _p(dP), n, b,
                                       voxels, c, ctx.eps, _stream()))
        dproto = (dP.float() / den.repeat(n // b, 1)[..., None]).contiguous()
        _run("proto_bwd2", "", fs.numel() * fs.element_size() * 2, 0,
             lambda: lib.pb_proto_bwd2(dt, _p(labels), _p(dproto), _p(dfs), n, b, voxels, c, _stream()))
        return dfs, None, None, None, None


def proto_sums(fs, ft, labels, cnt, eps=1e-5):
    return _ProtoSums.apply(fs, ft.detach(), labels, cnt, eps)


# ------------------------------------------------------------------------------------ token-path GEMMs (mmFormer transformer)
LINEAR_TC = os.environ.get("PB_LINEAR_TC", "1") != "0"


def _gemm_tc(name, a, b, bias, out, M, N, K, lda, ldb, a_kmajor, b_kmajor):
    lib = _lib.load()
    err = _tc_err_flag(a.device)
    nws = int(lib.pb_gemm_tc_workspace_floats(M, N, K))
    ws = _scratch.zeros((nws,), torch.float32, a.device) if nws else None          # split-K partial sums (zero arena)
    _run(name, f"m{M} n{N} k{K}", (M * K + N * K) * 2 + M * N * out.element_size(), 2.0 * M * N * K,
         lambda: lib.pb_gemm_tc(_p(a), _p(b), _p(bias), _p(out), _p(ws), M, N, K, lda, ldb, out.stride(0), int(a_kmajor), int(b_kmajor),
                                int(out.dtype == torch.float32), _p(err), _stream()))


class _LinearTC(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM (csrc/gemm_tc.cu): x [M, K] bf16, W [N, K] fp32 parameter (multiplied as bf16, like the
    conv kernels do), b [N] fp32 or None.  Backward: dx = dy W and dW = dy^T x through the same kernel (operands read in place,
    MN-major where the reduction runs over rows), db = column sums of dy."""

    @staticmethod
    def forward(ctx, x, w, b):
        M, K = x.shape
        N = w.shape[0]
        wq = w.detach().to(torch.bfloat16).contiguous()
        y = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
        _gemm_tc("linear_fwd", x, wq, None if b is None else b.detach().float().contiguous(), y, M, N, K, K, K, True, True)
        ctx.save_for_backward(x, wq)
        ctx.has_bias = b is not None
        ctx.w_dtype = w.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wq = ctx.saved_tensors
        M, K = x.shape
        N = wq.shape[0]
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=torch.bfloat16, device=x.device)
            # D[m][k'] = sum_n dy[m][n] W[n][k']: A = dy K-major, B = W [N][K] read as [reduction = N][rows = K]
            _gemm_tc("linear_dgrad", dy, wq, None, dx, M, K, N, N, K, True, False)
        if ctx.needs_input_grad[1]:
            dw = torch.empty((N, K), dtype=torch.float32, device=x.device)
            # D[n][k'] = sum_m dy[m][n] x[m][k']: both operands read as [reduction = M][rows]
            _gemm_tc("linear_wgrad", dy, x, None, dw, N, K, M, N, K, False, False)
            dw = dw.to(ctx.w_dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(0)
        return dx, dw, db


def linear_tc_eligible(x, weight):
    return (LINEAR_TC and TC_ENABLED and x.is_cuda and x.dtype == torch.bfloat16 and weight.shape[0] % 8 == 0
            and weight.shape[1] % 8 == 0)


def linear(x, weight, bias=None):
    """torch.nn.functional.linear for the transformer token path: bf16 activations run on the tcgen05 GEMM, anything else
    (the fp32 check mode) on the library GEMM."""
    if not linear_tc_eligible(x, weight):
        return torch.nn.functional.linear(x, weight.to(x.dtype), None if bias is None else bias.to(x.dtype))
    shp = x.shape
    y = _LinearTC.apply(x.reshape(-1, shp[-1]).contiguous(), weight, bias)
    return y.view(*shp[:-1], weight.shape[0])


# ------------------------------------------------------------------------------------ attention + LayerNorm of the token path
ATTN_TC = os.environ.get("PB_ATTN_TC", "1") != "0"
ATTN_FUSED = os.environ.get("PB_ATTN_FUSED", "1") != "0"      # scores fused with their softmax (head width <= 64); 0: GEMM + row kernels


def _gemm_tc_batched(name, a, b, out, M, N, K, lda, ldb, ldd, a_kmajor, b_kmajor, nb0, nb1, sa, sb, sd):
    lib = _lib.load()
    err = _tc_err_flag(a.device)
    nb = nb0 * nb1
    _run(name, f"b{nb} m{M} n{N} k{K}", nb * ((M * K + N * K) * 2 + M * N * out.element_size()), 2.0 * nb * M * N * K,
         lambda: lib.pb_gemm_tc_batched(_p(a), _p(b), _p(out), M, N, K, lda, ldb, ldd, int(a_kmajor), int(b_kmajor),
                                        int(out.dtype == torch.float32), nb0, nb1, sa[0], sa[1], sb[0], sb[1], sd[0], sd[1], _p(err),
                                        _stream()))


class _AttentionTC(torch.autograd.Function):
    """softmax(q k^T / sqrt(d)) v per (sample, head) with attention dropout — SelfAttention.forward, reference models/mmformer.py:203-213
    — on the batched tcgen05 GEMM (csrc/gemm_tc.cu) and the row kernels of csrc/attn.cu.  qkv [N, T, 3, H, d] bf16 as the qkv GEMM
    wrote it (the heads are read in place as column blocks of its rows); returns [N, T, H * d] bf16, the layout the proj GEMM reads.
    Scores and their gradient are fp32 [N, H, T, T]; the probabilities are kept as bf16 (P, and with dropout P' = P keep / (1 - p))."""

    @staticmethod
    def forward(ctx, qkv, drop_p):
        lib = _lib.load()
        N, T, _, H, d = qkv.shape
        C = H * d
        dev = qkv.device
        ldp = (T + 7) // 8 * 8
        scale = float(d) ** -0.5
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]            # views: [N, T, H, d], row stride 3C, head stride d
        sq = (T * 3 * C, d)
        p = torch.empty((N, H, T, ldp), dtype=torch.bfloat16, device=dev)
        if drop_p > 0.0:
            pd = torch.empty_like(p)
            seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=dev)      # graph-safe: drawn from torch's CUDA generator
        else:
            pd, seed = None, None
        fused = ATTN_FUSED and d <= 64
        if fused:
            # scores and their softmax in one kernel: no fp32 [N, H, T, T] tensor (csrc/gemm_tc.cu attn_rows_kernel)
            err = _tc_err_flag(dev)
            _run("attn_scores_softmax", f"b{N * H} t{T} d{d}", N * H * T * (2 * d * 2 + ldp * (2 if pd is None else 4)), 4.0 * N * H * T * T * d,
                 lambda: lib.pb_attn_scores_softmax(_p(q), _p(k), _p(p), _p(pd), N, H, T, d, 3 * C, sq[0], sq[1], ldp, scale, float(drop_p),
                                                    _p(seed), _p(err), _stream()))
        else:
            s = torch.empty((N, H, T, T), dtype=torch.float32, device=dev)
            # S[t][u] = sum_c q[t][c] k[u][c]
            _gemm_tc_batched("attn_qk", q, k, s, T, T, d, 3 * C, 3 * C, T, True, True, N, H, sq, sq, (H * T * T, T * T))
            _run("attn_softmax_fwd", f"t{T}", N * H * T * (T * 4 + ldp * (2 if pd is None else 4)), 0.0,
                 lambda: lib.pb_attn_softmax_fwd(_p(s), _p(p), _p(pd), N * H * T, T, ldp, scale, float(drop_p), _p(seed), _stream()))
            del s
        pm = p if pd is None else pd
        o = torch.empty((N, T, C), dtype=torch.bfloat16, device=dev)
        # O[t][c] = sum_u P'[t][u] v[u][c]: B = v read as [reduction = u][rows = c]
        _gemm_tc_batched("attn_pv", pm, v, o, T, d, T, ldp, 3 * C, C, True, False, N, H, (H * T * ldp, T * ldp), sq, (T * C, d))
        ctx.save_for_backward(qkv, p, pm, o)
        ctx.scale, ctx.fused = scale, fused
        return o

    @staticmethod
    def backward(ctx, do):
        lib = _lib.load()
        qkv, p, pm, o = ctx.saved_tensors
        N, T, _, H, d = qkv.shape
        C = H * d
        dev = qkv.device
        ldp = p.shape[-1]
        do = do.contiguous()
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        sq, sp, so = (T * 3 * C, d), (H * T * ldp, T * ldp), (T * C, d)
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]
        # dV[u][c] = sum_t P'[t][u] dO[t][c]: both operands read as [reduction = t][rows]
        _gemm_tc_batched("attn_dv", pm, do, dv, T, d, T, ldp, C, 3 * C, False, False, N, H, sp, so, sq)
        ds = torch.empty((N, H, T, ldp), dtype=torch.bfloat16, device=dev)
        if ctx.fused:
            # dS = scale (P' .* (dO V^T) - P * delta), delta = rowsum(dO .* O): the fp32 dP' never leaves the SM
            err = _tc_err_flag(dev)
            delta = torch.empty((N, H, T), dtype=torch.float32, device=dev)
            _run("attn_delta", f"t{T}", 2 * do.numel() * 2, 0.0, lambda: lib.pb_attn_delta(_p(do), _p(o), _p(delta), N, T, H, d, _stream()))
            _run("attn_dsoftmax", f"b{N * H} t{T} d{d}", N * H * T * (2 * d * 2 + ldp * 6), 2.0 * N * H * T * T * d,
                 lambda: lib.pb_attn_dsoftmax(_p(do), _p(v), _p(p), _p(pm), _p(delta), _p(ds), N, H, T, d, C, so[0], so[1], 3 * C, sq[0], sq[1],
                                              ldp, ctx.scale, _p(err), _stream()))
        else:
            # dP'[t][u] = sum_c dO[t][c] v[u][c]
            dp = torch.empty((N, H, T, T), dtype=torch.float32, device=dev)
            _gemm_tc_batched("attn_dp", do, v, dp, T, T, d, C, 3 * C, T, True, True, N, H, so, sq, (H * T * T, T * T))
            _run("attn_softmax_bwd", f"t{T}", N * H * T * (T * 4 + ldp * 6), 0.0,
                 lambda: lib.pb_attn_softmax_bwd(_p(dp), _p(p), _p(pm), _p(ds), N * H * T, T, ldp, ctx.scale, _stream()))
            del dp
        # dQ[t][c] = sum_u dS[t][u] k[u][c];  dK[u][c] = sum_t dS[t][u] q[t][c]
        _gemm_tc_batched("attn_dq", ds, k, dq, T, d, T, ldp, 3 * C, 3 * C, True, False, N, H, sp, sq, sq)
        _gemm_tc_batched("attn_dk", ds, q, dk, T, d, T, ldp, 3 * C, 3 * C, False, False, N, H, sp, sq, sq)
        return dqkv, None


def attention_tc_eligible(qkv):
    return (ATTN_TC and LINEAR_TC and TC_ENABLED and qkv.is_cuda and qkv.dtype == torch.bfloat16 and qkv.dim() == 5
            and qkv.shape[-1] % 8 == 0 and qkv.shape[0] * qkv.shape[3] <= 65535)


def attention(qkv, drop_p=0.0):
    """qkv [N, T, 3, H, d] -> [N, T, H * d]: bf16 on the tcgen05 path above; otherwise (fp32 check mode) the library's fused attention."""
    if attention_tc_eligible(qkv):
        return _AttentionTC.apply(qkv.contiguous(), float(drop_p))
    N, T, _, H, d = qkv.shape
    t = qkv.permute(2, 0, 3, 1, 4)
    h = torch.nn.functional.scaled_dot_product_attention(t[0], t[1], t[2], dropout_p=drop_p)
    return h.transpose(1, 2).reshape(N, T, H * d)


class _LayerNorm(torch.autograd.Function):
    """nn.LayerNorm over the last dimension (PreNorm / PreNormDrop, reference models/mmformer.py:233-250): csrc/attn.cu, statistics
    and affine parameters in fp32, activations in their storage type."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        lib = _lib.load()
        C = x.shape[-1]
        rows = x.numel() // C
        y = torch.empty_like(x)
        mean = torch.empty((rows,), dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        wf, bf = w.detach().float().contiguous(), b.detach().float().contiguous()
        _run("layernorm_fwd", f"c{C}", 2 * x.numel() * x.element_size(), 0.0,
             lambda: lib.pb_layernorm_fwd(_dt(x), _p(x), _p(wf), _p(bf), _p(y), _p(mean), _p(rstd), rows, C, float(eps), _stream()))
        ctx.save_for_backward(x, wf, mean, rstd)
        ctx.pdt = (w.dtype, b.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, wf, mean, rstd = ctx.saved_tensors
        C = x.shape[-1]
        rows = x.numel() // C
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        acc = _scratch.zeros((2, C), torch.float32, x.device)
        _run("layernorm_bwd", f"c{C}", 3 * x.numel() * x.element_size(), 0.0,
             lambda: lib.pb_layernorm_bwd(_dt(x), _p(dy), _p(x), _p(mean), _p(rstd), _p(wf), _p(dx), _p(acc[0]), _p(acc[1]), rows, C, _stream()))
        # copies: the accumulators live in the zero-scratch arena, which the next step clears
        return dx, acc[0].to(ctx.pdt[0], copy=True), acc[1].to(ctx.pdt[1], copy=True), None


def layer_norm(x, weight, bias, eps=1e-5):
    """F.layer_norm over the last dimension on csrc/attn.cu (fp32 or bf16 activations, C in {256, 512, 1024})."""
    if not (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and x.shape[-1] in (256, 512, 1024)):
        return torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps).to(x.dtype)
    return _LayerNorm.apply(x.contiguous(), weight, bias, eps)
