"""GPU parity of the RFNet + PASSION training step AT SCALE: B = 2, 4x64^3 crops, bench.py's masks (fixture
tests/golden/rfnet_passion_idt64.npz, written from the UNMODIFIED reference by oracle/gen_golden_full.py together with the
float64 oracle's gradient-sensitivity calibration and the bf16-storage noise floor of the network at this size).

Round 1 pinned parity at 16^3-32^3 only, where the coarsest InstanceNorms see 8-27 voxels; here they see 512, and the
80^3 forward quantities are additionally checked on the box by bench.py's `parity` block.

What the numbers say about BASELINE.json's tolerances (all measured, see the fixture):
  * forward quantities, fp32 check mode: <= 1e-4 holds (asserted below against the reference's own outputs);
  * gradients, fp32: the CPU fp32 oracle itself is 6e-4 (global rel-L2) from the float64 oracle, and a 2e-6 relative input
    perturbation moves the float64 gradient by 2.7e-3: the gradient of a LeakyReLU/clamp network is a sum over ~1e8 units of
    terms that jump when a pre-activation crosses zero, so rounding of relative size eps flips ~eps of them and moves the
    gradient by ~sqrt(eps) — 3e-4 for fp32 round-off, ~1e-1 for bf16 storage (eps = 2^-8).  1e-4 (fp32) / 1e-2 (bf16) on
    gradients is therefore not attainable by ANY implementation that rounds differently from the reference; the bars below
    are the measured floors with stated head-room.
"""
import os

import numpy as np
import pytest
import torch

from test_model_gpu import GOLD, _cuda_step, _is_cancelled_bias, _oracle, rel

pytestmark = pytest.mark.gpu
NAMES = ["fuse_prob", "prm_loss", "sep_loss", "kl_loss", "proto_loss", "dist"]


def _setup64(dtype):
    from oracle import synth
    from passion_b200.models import rfnet
    z = np.load(os.path.join(GOLD, "rfnet_passion_idt64.npz"), allow_pickle=True)
    x, target, mask, _ = synth.make_batch(int(z["B"]), int(z["S"]), seed=int(z["seed"]), labels=str(z["labels_kind"]),
                                          mask_ids=[int(i) for i in z["mask_ids"]])
    assert torch.equal(mask, torch.from_numpy(z["mask"]))
    sd = synth.make_state_dict(1037)
    model = rfnet.Model(num_cls=4).cuda()
    model.load_state_dict(sd)
    model.is_training, model.use_passion, model.mask_type = True, True, "idt"
    model.compute_dtype = dtype
    return z, model, sd, x, target, mask


@pytest.fixture(scope="module")
def oracle64():
    """One fp32 CPU oracle step at 64^3, B = 2 (~10-20 s): full gradients for the rel-L2 comparisons."""
    z, _, sd, x, target, mask = _setup64(torch.float32)
    outs, loss, grads = _oracle(sd, x, target, mask, z)
    return [o.detach() for o in outs], float(loss), grads


def _flat(named, keys):
    return torch.cat([named[k].flatten().double().cpu() for k in keys])


def test_fp32_check_mode_64(lib_built, oracle64):
    z, model, sd, x, target, mask = _setup64(torch.float32)
    outs, loss, parts = _cuda_step(model, x, target, mask, z)
    o_outs, o_loss, o_grads = oracle64
    # forward: against the reference's own outputs (golden) and the oracle
    assert rel(outs[0][:, :, ::4, ::4, ::4], torch.from_numpy(z["fuse_prob"])) < 1e-4
    assert rel(outs[0], o_outs[0]) < 1e-4
    for n, a, b in zip(NAMES[1:], outs[1:], o_outs[1:]):
        assert rel(a, torch.from_numpy(z[n])) < 1e-4, (n, "vs golden", rel(a, torch.from_numpy(z[n])))
        assert rel(a, b) < 1e-4, (n, "vs oracle")
    assert abs(float(loss) - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    assert np.allclose(parts["rp_iter"].detach().cpu().numpy(), z["rp_iter"], atol=1e-3, equal_nan=True)
    # integer output: argmax identical wherever the reference's own top-2 gap is above fp32 evaluation-order noise
    pred = outs[0].argmax(1)[:, ::2, ::2, ::2].cpu().numpy().astype(np.int8)
    mism = pred != z["fuse_argmax_s2"]
    assert mism.sum() <= max(3, 2e-4 * mism.size), int(mism.sum())
    assert (z["fuse_gap_s2"].astype(np.float32)[mism] < 2e-4).all()
    # gradients: per tensor within max(1e-4, 4 x float64 sensitivity to 2e-6 input noise) of the fp32 CPU oracle (which is
    # itself `fp32_oracle_err` from the float64 gradient), and the reference's golden norms to the same level
    names = [str(k) for k in z["grad_names"]]
    sens = dict(zip(names, z["sens"])); err32 = dict(zip(names, z["fp32_oracle_err"])); gn = dict(zip(names, z["grad_norms"]))
    sens_g = float(z["sens_global"])
    params = dict(model.named_parameters())
    keys = [k for k in params if not _is_cancelled_bias(k)]
    scale = float(_flat(o_grads, keys).norm())
    bad, worst = [], 0.0
    for k in keys:
        go = o_grads[k]
        if float(go.norm()) < 1e-7 * scale:
            assert float(params[k].grad.norm()) < 1e-5 * scale, k
            continue
        r = rel(params[k].grad, go)
        worst = max(worst, r)
        if not r < max(1e-4, 4 * sens[k], 4 * sens_g) + err32[k]:
            bad.append((k, r, sens[k]))
        assert abs(float(params[k].grad.double().norm()) - gn[k]) < max(5e-3, 6 * sens_g, 6 * sens[k]) * gn[k], k
    r_g = rel(_flat({k: p.grad for k, p in params.items()}, keys), _flat(o_grads, keys))
    print(f"idt64 fp32: global grad rel-L2 vs fp32 CPU oracle {r_g:.2e} (worst tensor {worst:.2e}); float64 sensitivity "
          f"{sens_g:.2e}, fp32 oracle vs float64 {float(z['fp32_oracle_err_global']):.2e}; fuse_prob rel {rel(outs[0], o_outs[0]):.2e}")
    assert not bad, bad[:8]
    assert r_g < max(1e-4, 4 * sens_g)
    for k in params:
        if _is_cancelled_bias(k):
            assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0


def test_bf16_64(lib_built, oracle64):
    """bf16 storage + tcgen05 operands at the scale the benchmark runs.  Bars = 2x the CPU simulation of bf16 STORAGE alone
    (the oracle with activations rounded to bf16, stored in the fixture: probabilities 1.9e-2, gradients 1.2e-1), i.e. the
    kernels may at most double the error the storage format itself causes; losses meet BASELINE.json's 1e-2 outright."""
    z, model, sd, x, target, mask = _setup64(torch.bfloat16)
    outs, loss, parts = _cuda_step(model, x, target, mask, z)
    o_outs, o_loss, o_grads = oracle64
    params = dict(model.named_parameters())
    keys = [k for k in params if not _is_cancelled_bias(k)]
    r_p = rel(outs[0], o_outs[0])
    r_g = rel(_flat({k: p.grad for k, p in params.items()}, keys), _flat(o_grads, keys))
    r_l = [rel(a, b) for a, b in zip(outs[1:], o_outs[1:])]
    agree = float((outs[0].argmax(1).cpu() == o_outs[0].argmax(1)).float().mean())
    print(f"idt64 bf16: fuse_prob rel-L2 {r_p:.3e} (storage floor {float(z['bf16_sim_prob_rel']):.2e}); global grad rel-L2 {r_g:.3e} "
          f"(floor {float(z['bf16_sim_grad_rel']):.2e}); per-sample losses {[round(v, 5) for v in r_l]}; loss "
          f"{abs(float(loss) - o_loss) / abs(o_loss):.2e}; argmax agreement {agree:.4f}")
    assert r_p < 2 * float(z["bf16_sim_prob_rel"])
    assert r_g < 2 * float(z["bf16_sim_grad_rel"])
    for v in r_l:
        assert v < 1e-2
    assert abs(float(loss) - o_loss) < 1e-3 * abs(o_loss)
    assert agree > 0.97
    from passion_b200 import ops
    ops.check_tc_errors()
