"""Module-level GPU parity: the four grouped encoders, the separate decoder, the PRM generators (last
stage and with an upper feature) and the region-aware modal fusion block of passion_b200.models.rfnet, each ALONE, in the fp32 check mode,
against the CPU oracle's functions of the same name (oracle/rfnet_oracle.py, which is pinned against the unmodified reference) on the
same synthetic weights and inputs.  The oracle is evaluated in float64 here: its fp32 CPU backward is itself 1e-4 .. 1e-3 away from
float64 on the coarse-level input gradients (scripts/debug_modules.py), the CUDA fp32 path ~1e-6, so float64 is the only reference
that makes a 1e-4 bar meaningful for gradients.  The end-to-end tests
(test_model_gpu.py) exercise the same code inside Model.forward; these localise a failure to one block."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cl(t):            # NCDHW -> NDHWC
    return t.permute(0, 2, 3, 4, 1).contiguous()


def nc(t):            # NDHWC -> NCDHW
    return t.permute(0, 4, 1, 2, 3)


@pytest.fixture(scope="module")
def setup(lib_built):
    from oracle import synth
    from passion_b200.models import rfnet
    sd = synth.make_state_dict(1037)
    model = rfnet.Model(4).cuda()
    model.load_state_dict(sd)
    model.compute_dtype = torch.float32
    P = {k: v.double() if v.is_floating_point() else v.clone() for k, v in sd.items()}
    return model, P


def test_encoders_grouped(setup):
    """_run_encoders: four modality encoders as one grouped pass == oracle.encoder per modality (rfnet.py:36-48)."""
    from oracle import rfnet_oracle as O
    from passion_b200 import ops
    from passion_b200.models import rfnet
    model, P = setup
    B, S = 2, 16
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 4, S, S, S, generator=g)
    ops.begin_step(torch.device("cuda", 0))
    encs = [model.flair_encoder, model.t1ce_encoder, model.t1_encoder, model.t2_encoder]
    xin = x.permute(1, 0, 2, 3, 4).reshape(4 * B, S, S, S, 1).cuda().contiguous()          # modality-major
    feats = rfnet._run_encoders(encs, xin)
    for m, pre in enumerate(("flair_encoder", "t1ce_encoder", "t1_encoder", "t2_encoder")):
        ref = O.encoder(P, pre, x[:, m:m + 1].double())
        for lvl in range(4):
            got = nc(feats[lvl][m * B:(m + 1) * B])
            assert rel(got, ref[lvl]) < 1e-4, (pre, lvl, rel(got, ref[lvl]))


def _levels(B, S, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, 8 * 2 ** i, S >> i, S >> i, S >> i, generator=g) for i in range(4)]


def test_decoder_sep(setup):
    """Decoder_sep.run (logits) + softmax == oracle.decoder_sep (rfnet.py:50-89), with input gradients."""
    from oracle import rfnet_oracle as O
    from passion_b200 import ops
    model, P = setup
    xs = _levels(2, 16, 5)
    ops.begin_step(torch.device("cuda", 0))
    xc = [cl(t).cuda().requires_grad_(True) for t in xs]
    logits = model.decoder_sep.run(*xc)
    prob = torch.softmax(nc(logits).float(), 1)
    xr = [t.double().requires_grad_(True) for t in xs]
    ref = O.decoder_sep(P, *xr)
    assert rel(prob, ref) < 1e-4
    g = torch.Generator().manual_seed(9)
    w = torch.randn(ref.shape, generator=g).double()
    (prob * w.cuda().float()).sum().backward()
    (ref * w).sum().backward()
    for a, b in zip(xc, xr):
        assert rel(nc(a.grad), b.grad) < 1e-4, rel(nc(a.grad), b.grad)


@pytest.mark.parametrize("level", [4, 3])
def test_prm_generator(setup, level):
    """prm_generator_pk.run == oracle.prm_generator: last stage (level 4, no upper feature, blocks.py:396-416) and with the upper
    feature concatenated in front of the embedding (level 3, blocks.py:443-464)."""
    from oracle import rfnet_oracle as O
    from passion_b200 import ops
    model, P = setup
    B, S = 2, 8
    C = 8 * 2 ** (level - 1)
    g = torch.Generator().manual_seed(level)
    x = torch.randn(B, 4, C, S, S, S, generator=g)
    mask = torch.tensor([[True, False, True, True], [False, True, True, False]])
    upper = torch.randn(B, C, S, S, S, generator=g) if level == 3 else None
    ops.begin_step(torch.device("cuda", 0))
    y = cl(O.mask_modal(x, mask).reshape(B, 4 * C, S, S, S)).cuda().requires_grad_(True)
    up = cl(upper).cuda().requires_grad_(True) if upper is not None else None
    mod = getattr(model.decoder_fuse, f"prm_generator{level}")
    logits = mod.run(y, up)
    xr = x.double().requires_grad_(True)
    ur = upper.double().requires_grad_(True) if upper is not None else None
    ref = O.prm_generator(P, f"decoder_fuse.prm_generator{level}", xr, mask, ur)
    assert rel(nc(logits), ref) < 1e-4
    w = torch.randn(ref.shape, generator=g).double()
    (nc(logits).float() * w.cuda().float()).sum().backward()
    (ref * w).sum().backward()
    gy = nc(y.grad).reshape(B, 4, C, S, S, S)
    m6 = mask.view(B, 4, 1, 1, 1, 1)
    assert rel(gy * m6.cuda(), xr.grad * m6) < 1e-4           # the oracle masks inside: compare on the present modalities
    if upper is not None:
        assert rel(nc(up.grad), ur.grad) < 1e-4


def test_rfm_block(setup):
    """region_aware_modal_fusion.run == oracle.rfm (blocks.py:582-626: class-wise masked pooling, gate MLP, gated mix, region-fusion
    convs, short cut), on detached class probabilities."""
    from oracle import rfnet_oracle as O
    from passion_b200 import ops
    model, P = setup
    B, S, C = 2, 8, 32
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 4, C, S, S, S, generator=g)
    prm = torch.softmax(torch.randn(B, 4, S, S, S, generator=g), 1)
    mask = torch.tensor([[True, True, False, True], [True, False, False, False]])
    ops.begin_step(torch.device("cuda", 0))
    y = cl(O.mask_modal(x, mask).reshape(B, 4 * C, S, S, S)).cuda().requires_grad_(True)
    out = model.decoder_fuse.RFM3.run(y, cl(prm).cuda().float().contiguous())
    xr = x.double().requires_grad_(True)
    ref = O.rfm(P, "decoder_fuse.RFM3", xr, prm.double(), mask)
    assert rel(nc(out), ref) < 1e-4
    w = torch.randn(ref.shape, generator=g).double()
    (nc(out).float() * w.cuda().float()).sum().backward()
    (ref * w).sum().backward()
    m6 = mask.view(B, 4, 1, 1, 1, 1)
    gy = nc(y.grad).reshape(B, 4, C, S, S, S)
    assert rel(gy * m6.cuda(), xr.grad * m6) < 1e-4
