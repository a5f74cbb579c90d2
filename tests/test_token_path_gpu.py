"""Token path of the mmFormer transformer blocks (reference models/mmformer.py:192-313) on our own kernels: attention
as batched tcgen05 GEMMs + row softmax (ops.attention: csrc/gemm_tc.cu pb_gemm_tc_batched, csrc/attn.cu) and LayerNorm (ops.layer_norm),
forward and backward against float64 restatements of the reference's formulas on the same bf16-rounded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def ref_attention(qkv, mask=None, drop_p=0.0):
    """SelfAttention.forward, mmformer.py:203-213, in float64: softmax(q k^T * d^-0.5) (.* mask / (1 - p)) @ v, heads concatenated."""
    N, T, _, H, d = qkv.shape
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    a = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, -1)
    if mask is not None:
        a = a * mask / (1.0 - drop_p)
    return (a @ v).transpose(1, 2).reshape(N, T, H * d)


# N, T, H, d: 125 / 500 tokens = the 80^3 crop (ragged rows, row stride padded to 128 / 504), 512 / 2048 = the 128^3 crop, one tile and
# less than one tile, head widths other than 64
ATTN_CASES = [(2, 125, 8, 64), (1, 500, 8, 64), (1, 512, 8, 64), (1, 2048, 2, 64), (3, 64, 4, 32), (2, 9, 2, 16), (1, 200, 3, 128),
              (2, 130, 8, 8)]


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "gemm+rows"])
@pytest.mark.parametrize("case", ATTN_CASES, ids=lambda c: "n%d_t%d_h%d_d%d" % c)
def test_attention_tc(lib_built, case, fused, monkeypatch):
    """fused: scores and softmax (forward) / dO V^T and the softmax backward in one kernel, no fp32 [N, H, T, T] tensor (head width <= 64);
    gemm+rows: the same products through pb_gemm_tc_batched with fp32 scores and the row kernels of csrc/attn.cu."""
    from passion_b200 import ops
    N, T, H, d = case
    monkeypatch.setattr(ops, "ATTN_FUSED", fused)
    g = torch.Generator(device="cpu").manual_seed(T * 13 + H * 5 + d)
    qkv = (torch.randn(N, T, 3, H, d, generator=g) * 1.5).cuda().bfloat16().requires_grad_(True)
    assert ops.attention_tc_eligible(qkv)
    o = ops.attention(qkv)
    go = torch.randn(N, T, H * d, generator=g).cuda().bfloat16()
    o.backward(go)
    qr = qkv.detach().double().requires_grad_(True)
    orf = ref_attention(qr)
    orf.backward(go.double())
    assert o.shape == (N, T, H * d) and o.dtype == torch.bfloat16
    assert rel(o, orf) < 8e-3, rel(o, orf)
    for i, name in enumerate("qkv"):
        e = rel(qkv.grad[:, :, i], qr.grad[:, :, i])
        assert e < 1.5e-2, (name, e)              # P and dS are stored as bf16 (2^-9 each), as a fused bf16 attention stores them in registers
    ops.check_tc_errors()


def test_attention_tc_matches_library(lib_built):
    """Same inputs through the library's fused attention (the fp32 check mode's path): both are bf16-P implementations of one formula."""
    from passion_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(77)
    qkv = torch.randn(2, 500, 3, 8, 64, generator=g).cuda().bfloat16()
    o = ops.attention(qkv)
    t = qkv.permute(2, 0, 3, 1, 4)
    lib = torch.nn.functional.scaled_dot_product_attention(t[0], t[1], t[2]).transpose(1, 2).reshape(2, 500, 512)
    ref = ref_attention(qkv.double())
    assert rel(o, ref) < 1.5 * max(rel(lib, ref), 3e-3)
    o32 = ops.attention(qkv.float())                          # not eligible: library path, fp32
    assert o32.dtype == torch.float32 and rel(o32, ref) < 1e-5


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "gemm+rows"])
def test_attention_dropout(lib_built, fused, monkeypatch):
    """Attention dropout (mmformer.py:208, p = 0.1 in train mode): the kept set is a fresh Bernoulli(1 - p) draw per call, kept
    probabilities are scaled by 1 / (1 - p), and forward and backward use the SAME mask (checked against float64 with the mask read
    back from the saved P')."""
    from passion_b200 import ops
    monkeypatch.setattr(ops, "ATTN_FUSED", fused)
    N, T, H, d, p = 2, 250, 8, 64, 0.1
    g = torch.Generator(device="cpu").manual_seed(4)
    qkv = torch.randn(N, T, 3, H, d, generator=g).cuda().bfloat16().requires_grad_(True)
    torch.manual_seed(11)
    o = ops.attention(qkv, p)
    P, Pd = o.grad_fn.saved_tensors[1][..., :T], o.grad_fn.saved_tensors[2][..., :T]
    big = P.float() > 1e-4                                   # entries that cannot round to zero
    kept = (Pd != 0) & big
    frac = float(kept.sum() / big.sum())
    assert abs(frac - (1 - p)) < 5e-3, frac
    assert rel(Pd[kept].float(), P[kept].float() / (1 - p)) < 4e-3
    mask = ((Pd != 0) | ~big).double()
    go = torch.randn(N, T, H * d, generator=g).cuda().bfloat16()
    o.backward(go)
    qr = qkv.detach().double().requires_grad_(True)
    orf = ref_attention(qr, mask, p)
    orf.backward(go.double())
    assert rel(o, orf) < 8e-3
    for i in range(3):
        assert rel(qkv.grad[:, :, i], qr.grad[:, :, i]) < 1.5e-2
    o2 = ops.attention(qkv.detach(), p)                      # next call: another mask
    assert rel(o2, o) > 0.05
    torch.manual_seed(11)
    o3 = ops.attention(qkv.detach(), p)                      # same generator state: same mask
    assert torch.equal(o3, o.detach())
    monkeypatch.setattr(ops, "ATTN_FUSED", not fused)        # both paths draw the same mask from the same seed
    torch.manual_seed(11)
    o4 = ops.attention(qkv.detach(), p)
    assert rel(o4, o) < 4e-3
    ops.check_tc_errors()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 125, 512), (5, 2048, 512), (3, 256), (7, 33, 1024)], ids=lambda s: "x".join(map(str, s)))
def test_layer_norm(lib_built, dtype, shape):
    from passion_b200 import ops
    C = shape[-1]
    g = torch.Generator(device="cpu").manual_seed(C + len(shape))
    x = (torch.randn(*shape, generator=g) * 2 + 0.5).cuda().to(dtype).requires_grad_(True)
    w = (1 + 0.2 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    b = (0.1 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    ops.begin_step(torch.device("cuda", 0))
    y = ops.layer_norm(x, w, b, 1e-5)
    gy = torch.randn(*shape, generator=g).cuda().to(dtype)
    y.backward(gy)
    xr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.layer_norm(xr, (C,), wr, br, 1e-5)
    yr.backward(gy.double())
    tol = 2e-5 if dtype == torch.float32 else 8e-3
    assert y.dtype == dtype
    assert rel(y, yr) < tol
    assert rel(x.grad, xr.grad) < tol
    assert rel(w.grad, wr.grad) < 2e-5 and rel(b.grad, br.grad) < 2e-5


def test_layer_norm_fallback_width(lib_built):
    from passion_b200 import ops
    x = torch.randn(4, 10, 96).cuda()
    w, b = torch.ones(96).cuda(), torch.zeros(96).cuda()
    assert rel(ops.layer_norm(x, w, b), torch.nn.functional.layer_norm(x, (96,), w, b)) < 1e-6
