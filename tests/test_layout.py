"""CPU: repository contracts — the C-ABI library loads and exports every declared symbol; the product path
never imports the oracle and has no CPU fallback; host-side step logic agrees with the oracle."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib_built):
    from passion_b200 import _lib
    syms = _lib.declared_symbols()
    assert len(syms) >= 15 and "pb_conv3d_fwd" in syms
    for s in syms:
        assert hasattr(lib_built, s), s
    assert lib_built.pb_version() >= 100
    assert lib_built.pb_launch_count() >= 0


def test_sass_is_sm100a(lib_built):
    import subprocess
    from passion_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "passion_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f
    for f in ("train.py", "options.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert not pat.search(open(p).read()), f


def test_ops_refuse_cpu_tensors(lib_built):
    from passion_b200 import ops
    with pytest.raises(RuntimeError):
        ops.upsample(torch.zeros(1, 2, 2, 2, 8), 2)
    with pytest.raises(RuntimeError):
        ops.conv3d(torch.zeros(1, 4, 4, 4, 8), torch.zeros(1, 27, 8, 8))


def test_state_dict_layout_matches_reference_table():
    from oracle import synth
    from passion_b200.models import rfnet
    m = rfnet.Model(4)
    sd = m.state_dict()
    shapes = synth.rfnet_param_shapes()
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    m.load_state_dict(synth.make_state_dict(3))


def test_mmformer_state_dict_layout_matches_reference_table():
    from oracle import synth
    from passion_b200.models import mmformer
    m = mmformer.Model(4)
    sd = m.state_dict()
    shapes = synth.mmformer_param_shapes()
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k


def test_preference_update_and_lr_match_oracle():
    from oracle import train_step_oracle as o
    from passion_b200 import train_step as t
    beta = torch.tensor([1.0, 1.0, 1.0, 1.0])
    eta_o = eta_t = 0.01
    rs = np.random.RandomState(0)
    for epoch in range(0, 205, 17):
        d = torch.from_numpy(rs.uniform(0.1, 0.5, 4).astype(np.float32))
        bo, eta_o, ro = o.preference_update(beta, d, eta_o, epoch)
        bt, eta_t, rt = t.preference_update(beta, d, eta_t, epoch)
        assert torch.allclose(bo, bt, atol=1e-7) and eta_o == eta_t and torch.allclose(ro, rt, atol=1e-7)
        beta = bt
    for e in (0, 1, 150, 299):
        assert o.poly_lr(2e-4, e, 300) == t.poly_lr(2e-4, e, 300)


def test_build_model_sets_token_grid():
    from passion_b200.models import build_model
    m = build_model("mmformer", crop=128)
    assert tuple(m.flair_pos.shape) == (1, 8 ** 3, 512)          # 128 / 16 tokens per axis
    m = build_model("mmformer", crop=80)
    assert tuple(m.t2_pos.shape) == (1, 5 ** 3, 512)             # the reference's patch_size = 5
    with pytest.raises(ValueError):
        build_model("mmformer", crop=72)
    with pytest.raises(ValueError):
        build_model("m2ftrans")
    assert type(build_model("rfnet")).__name__ == "Model"


def test_zero_scratch_arena_hands_out_disjoint_zeroed_slices():
    """ops._ZeroScratch: linear hand-out, one re-zero per step, growth keeps earlier slices valid."""
    from passion_b200 import ops
    sc = ops._ZeroScratch()
    dev = torch.device("cpu")
    a = sc.zeros((3, 5, 2), torch.float64, dev)
    b = sc.zeros((7,), torch.float32, dev)
    assert a.shape == (3, 5, 2) and a.dtype == torch.float64 and b.dtype == torch.float32
    a += 1.0
    b += 2.0
    assert float(a.sum()) == 30.0 and float(b.sum()) == 14.0       # disjoint
    sc.begin_step(dev)
    a2 = sc.zeros((3, 5, 2), torch.float64, dev)
    assert a2.data_ptr() == a.data_ptr() and float(a2.abs().sum()) == 0.0          # recycled and re-zeroed
    big = sc.zeros((sc.MIN_BYTES // 4 + 10,), torch.float32, dev)                  # forces a new arena
    assert float(big.abs().sum()) == 0.0
    a2 += 3.0
    assert float(a2.sum()) == 90.0                                                 # old slice still alive
