"""CPU: the oracle restatement reproduces the REFERENCE's committed golden outputs (tests/golden/*.npz were
written by oracle/gen_golden.py from the unmodified reference modules)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _projection(name, n):
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % (2 ** 31 - 1)
    return np.random.RandomState(h).randint(0, 2, n).astype(np.float64) * 2 - 1


@pytest.mark.parametrize("case", ["idtU", "idtS", "pdtU", "idtU_nopassion"])
def test_oracle_matches_reference_golden(case):
    from oracle import criterions_oracle as oc
    from oracle import rfnet_oracle, synth, train_step_oracle
    from oracle.masks import mask_id_of
    z = np.load(os.path.join(GOLD, f"rfnet_passion_{case}.npz"), allow_pickle=True)
    B, S = int(z["B"]), int(z["S"])
    ids = [mask_id_of(m) for m in z["mask"]]
    x, target, mask, _ = synth.make_batch(B, S, seed=int(z["seed"]), labels=str(z["labels_kind"]), mask_ids=ids)
    assert np.array_equal(mask.numpy(), z["mask"])
    sd = synth.make_state_dict(1037)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    use_passion, mask_type = bool(z["use_passion"]), str(z["mask_type"])
    outs = rfnet_oracle.forward(P, x, mask, target, float(z["temp"]), use_passion=use_passion, mask_type=mask_type)
    names = ["fuse_prob", "prm_loss", "sep_loss", "kl_loss", "proto_loss", "dist"][:len(outs)]
    for n, o in zip(names, outs):
        assert np.allclose(o.detach().numpy(), z[n], atol=3e-5, rtol=1e-4), n
    if use_passion:
        loss, parts = train_step_oracle.loss_mix(outs, target, mask, torch.from_numpy(z["imb_beta"]),
                                                 torch.from_numpy(z["modal_weight"]), mask_type=mask_type)
        assert np.allclose(parts["rp_iter"].detach().numpy(), z["rp_iter"], atol=1e-4, equal_nan=True)
    else:
        fuse = (oc.softmax_weighted_loss_bs(outs[0], target) + oc.dice_loss_bs(outs[0], target)).sum()
        loss = fuse + outs[1].sum() + (outs[2] * mask).sum()
    assert abs(float(loss) - float(z["loss"])) < 1e-4 * max(1.0, abs(float(z["loss"])))
    loss.backward()
    norms = dict(zip(z["grad_names"], z["grad_norms"]))
    projs = dict(zip(z["grad_names"], z["grad_projs"]))
    for k, p in P.items():
        if norms[k] < 1e-5:          # conv biases feeding an InstanceNorm: rounding noise in the reference
            continue
        g = p.grad.double()
        assert abs(float(g.norm()) - norms[k]) < 3e-4 * norms[k], k
        pr = float((g.reshape(-1).numpy() * _projection(k, g.numel())).sum())
        assert abs(pr - projs[k]) < 3e-4 * norms[k] * np.sqrt(g.numel()) + 1e-7, k
    # inference path + bit-exact argmax
    with torch.no_grad():
        inf = rfnet_oracle.forward(sd, x, mask, is_training=False, mask_type=mask_type)
    assert np.allclose(inf.numpy(), z["infer_prob"], atol=3e-5)
    assert np.array_equal(inf.argmax(1).numpy().astype(np.int8), z["infer_argmax"])


def test_single_modality_sample_gives_nan_rp_iter():
    """Reference quirk pinned by the idtU fixture: a sample whose only present modality makes the single-
    modality path identical to the fused path has dist = 0, so dist/avg = 0/0 = NaN and rp_mask is all-False."""
    z = np.load(os.path.join(GOLD, "rfnet_passion_idtU.npz"), allow_pickle=True)
    assert z["mask"][0].sum() == 1 and np.isnan(z["rp_iter"]).all()


def test_param_table_matches_golden_names():
    from oracle import synth
    z = np.load(os.path.join(GOLD, "rfnet_passion_idtU.npz"), allow_pickle=True)
    shapes = synth.rfnet_param_shapes()
    assert sorted(shapes) == list(z["grad_names"])
    assert sum(int(np.prod(s)) for s in shapes.values()) == 2381792


@pytest.mark.parametrize("case", ["idtU", "idtS_t2only", "idtU_nopassion"])
def test_mmformer_oracle_matches_reference_golden(case):
    """mmFormer backbone (BASELINE.json configs[3]): fixtures from the unmodified reference at 32^3 (patch_size 2, eval())."""
    from oracle import mmformer_oracle, synth, train_step_oracle
    from oracle.masks import mask_id_of
    z = np.load(os.path.join(GOLD, f"mmformer_passion_{case}.npz"), allow_pickle=True)
    B, S = int(z["B"]), int(z["S"])
    ids = [mask_id_of(m) for m in z["mask"]]
    x, target, mask, _ = synth.make_batch(B, S, seed=int(z["seed"]), labels=str(z["labels_kind"]), mask_ids=ids)
    sd = synth.make_state_dict(2051, synth.mmformer_param_shapes(patch=2))
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    use_passion = bool(z["use_passion"])
    outs = mmformer_oracle.forward(P, x, mask, target, float(z["temp"]), use_passion=use_passion)
    names = ["fuse_prob", "prm_loss", "sep_loss", "kl_loss", "proto_loss", "dist"][:len(outs)]
    for n, o in zip(names, outs):
        o = o.detach().numpy()
        if n == "fuse_prob":
            assert np.array_equal(o.argmax(1).astype(np.int8), z["fuse_argmax"])
            o = o[:, :, ::2, ::2, ::2]
        assert np.allclose(o, z[n], atol=3e-5, rtol=1e-4), n
    if use_passion:
        loss, parts = train_step_oracle.loss_mix(outs, target, mask, torch.from_numpy(z["imb_beta"]),
                                                 torch.from_numpy(z["modal_weight"]))
        assert np.allclose(parts["rp_iter"].detach().numpy(), z["rp_iter"], atol=1e-4, equal_nan=True)
    else:
        loss, _ = train_step_oracle.loss_mix_baseline(outs, target, mask)
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-4 * max(1.0, abs(float(z["loss"])))
    loss.backward()
    norms = dict(zip(z["grad_names"], z["grad_norms"]))
    projs = dict(zip(z["grad_names"], z["grad_projs"]))
    for k, p in P.items():
        if norms[k] < 1e-5:
            continue
        g = p.grad.double()
        # two fp32 evaluations of this gradient agree to ~1e-3 (oracle/gen_golden.py main_mmformer measured 1.1e-3 worst)
        assert abs(float(g.norm()) - norms[k]) < 5e-3 * norms[k], k
        pr = float((g.reshape(-1).numpy() * _projection(k, g.numel())).sum())
        assert abs(pr - projs[k]) < 5e-3 * norms[k] * np.sqrt(g.numel()) + 1e-7, k
    with torch.no_grad():
        inf = mmformer_oracle.forward(sd, x, mask, is_training=False)
    assert np.array_equal(inf.argmax(1).numpy().astype(np.int8), z["infer_argmax"])


def test_mmformer_t2_pass_uses_t1_mask_in_the_transformer_branch():
    """Reference quirk (mmformer.py:522) pinned by the fixture: for a sample with only T2 present the T2 pass differs
    from the fused pass, so dist[.,3] != 0 and rp_iter stays finite (unlike RFNet's 0/0)."""
    z = np.load(os.path.join(GOLD, "mmformer_passion_idtS_t2only.npz"), allow_pickle=True)
    assert list(z["mask"][0]) == [False, False, False, True]
    assert z["dist"][0, 3] > 0 and np.isfinite(z["rp_iter"]).all()


def test_mmformer_param_table():
    from oracle import synth
    z = np.load(os.path.join(GOLD, "mmformer_passion_idtU.npz"), allow_pickle=True)
    assert sorted(synth.mmformer_param_shapes(patch=2)) == list(z["grad_names"])
    assert sum(int(np.prod(s)) for s in synth.mmformer_param_shapes().values()) == 35359568 + 4 * (125 - 8) * 512
