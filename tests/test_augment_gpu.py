"""GPU parity of the device-side training-sample pipeline: pb_augment_batch through
passion_b200.data.DeviceAugment against the reference's golden outputs and the oracle — BIT-EXACT (x float32, labels,
one-hot) — and the uint8 label-map target against the float64 one-hot target through the model and the trainer."""
import glob
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[len("augment_"):-4] for p in glob.glob(os.path.join(GOLD, "augment_*.npz")))


def _cases_on_device(vols_segs):
    from passion_b200 import data
    rc = data.ResidentCases("cuda")
    for vol, seg in vols_segs:
        rc.add(vol, seg)
    return rc


@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_golden(name, lib_built):
    from oracle import augment_oracle as ao
    from passion_b200 import data
    z = np.load(os.path.join(GOLD, f"augment_{name}.npz"))
    vol, seg = ao.synth_volume(int(z["vseed"]), tuple(int(v) for v in z["vshape"]))
    size = tuple(int(v) for v in z["size"])
    smp = data.AugmentSampler(size, py_rng=random.Random(int(z["py_seed"])), np_rng=np.random.RandomState(int(z["np_seed"])))
    p = smp.sample(vol.shape[:3])
    rc = _cases_on_device([(vol, seg)])
    aug = data.DeviceAugment("cuda", size=size, batch=1, want_onehot=True)
    x, labels, onehot = aug(rc, [0], [p])
    torch.cuda.synchronize()
    assert x.dtype == torch.float32 and labels.dtype == torch.uint8 and onehot.dtype == torch.float64
    assert np.array_equal(x[0].cpu().numpy(), z["x"]), "image differs from the reference transforms"
    assert np.array_equal(labels[0].cpu().numpy(), z["y"]), "labels differ from the reference transforms"
    ref_onehot = np.eye(4)[z["y"].astype(np.int64)].transpose(3, 0, 1, 2)                # datasets_nii.py:150-153
    assert np.array_equal(onehot[0].cpu().numpy(), ref_onehot)


def test_kernel_matches_oracle_full_size(lib_built):
    """BASELINE crop size (80^3) out of two different non-cubic cases, batch 2, several draws; every slot of the
    parameter ring is used more than once."""
    from oracle import augment_oracle as ao
    from passion_b200 import data
    shapes = [(96, 100, 90), (88, 84, 104)]
    vs = [ao.synth_volume(10 + i, s) for i, s in enumerate(shapes)]
    rc = _cases_on_device(vs)
    size = (80, 80, 80)
    r1, n1 = random.Random(1037), np.random.RandomState(1037)
    smp = data.AugmentSampler(size, py_rng=r1, np_rng=n1)
    aug = data.DeviceAugment("cuda", size=size, batch=2, want_onehot=False, slots=2)
    seen_axes = set()
    for it in range(5):
        ids = [it % 2, (it + 1) % 2]
        ps = [smp.sample(shapes[i]) for i in ids]
        x, labels, onehot = aug(rc, ids, ps)
        assert onehot is None
        xc, lc = x.cpu().numpy(), labels.cpu().numpy()
        for b in range(2):
            seen_axes.add(tuple(ps[b]["axes"]))
            p = dict(ps[b], size=list(size), shift=ps[b]["shift"].reshape(1, 80, 1, 1, 4), scale=ps[b]["scale"].reshape(1, 80, 1, 1, 4))
            ox, oy, _ = ao.apply(*vs[ids[b]], p)
            assert np.array_equal(xc[b], ox), f"iteration {it} sample {b}: image mismatch ({p['axes']}, {p['angle']}, {p['flip']})"
            assert np.array_equal(lc[b], oy.astype(np.uint8)), f"iteration {it} sample {b}: label mismatch"
    assert len(seen_axes) >= 2


def test_identity_draw_is_a_plain_crop(lib_built):
    """angle 0, no flips, scale 1, shift 0 -> exactly the crop (and the transposed layout)."""
    from oracle import augment_oracle as ao
    from passion_b200 import data
    vol, seg = ao.synth_volume(3, (40, 36, 44))
    rc = _cases_on_device([(vol, seg)])
    size = (32, 24, 40)
    p = dict(start=[5, 7, 2], axes=(2, 0), angle=0, flip=[False] * 3, shift=np.zeros((32, 4)), scale=np.ones((32, 4)))
    x, labels, _ = data.DeviceAugment("cuda", size=size, batch=1)(rc, [0], [p])
    crop = vol[5:37, 7:31, 2:42]
    assert np.array_equal(x[0].cpu().numpy(), crop.transpose(3, 0, 1, 2))
    assert np.array_equal(labels[0].cpu().numpy(), seg[5:37, 7:31, 2:42])


def _model(dtype):
    from oracle import synth
    from passion_b200.models import rfnet
    model = rfnet.Model(num_cls=4).cuda()
    model.load_state_dict(synth.make_state_dict(1037))
    model.is_training, model.use_passion, model.mask_type = True, True, "idt"
    model.compute_dtype = dtype
    return model


def test_label_map_target_gives_identical_losses(lib_built):
    """Model.forward + loss_mix with the uint8 label map == with the reference's float64 one-hot target (same kernels, same
    label bytes; the float64-atomic statistics make two runs agree to round-off, not necessarily to the bit)."""
    from oracle import synth
    from passion_b200.train_step import loss_mix
    x, target, mask, _ = synth.make_batch(2, 16, seed=5, labels="S", mask_ids=[10, 12])
    labels = target.argmax(1).to(torch.uint8)
    beta, mw = torch.tensor([1.1, 0.9, 1.3, 0.7]).cuda(), torch.tensor([2.4, 1.6, 1.2, 5.1]).cuda()
    model = _model(torch.float32)
    res = []
    for tgt in (target.cuda(), labels.cuda()):
        model.zero_grad(set_to_none=True)
        outs = model(x.cuda(), mask.cuda(), target=tgt, temp=4.0)
        loss, parts = loss_mix(outs, tgt, mask.cuda(), beta, mw, mask_type="idt")
        loss.backward()
        res.append((loss.detach().clone(), [o.detach().clone() for o in outs],
                    model.decoder_fuse.d1_c2.conv.weight.grad.detach().clone()))
    assert abs(float(res[0][0]) - float(res[1][0])) <= 1e-5 * abs(float(res[0][0]))
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6, equal_nan=True)
    assert float((res[0][2] - res[1][2]).norm() / res[0][2].norm()) < 1e-3


@pytest.mark.parametrize("use_graph", [False, True])
def test_trainer_step_from_resident_cases(lib_built, use_graph):
    """The whole device pipeline end to end: resident cases -> DeviceAugment -> Trainer.step with label-map targets ==
    Trainer.step on the oracle-augmented host batch with one-hot targets (same weights, same draws); eagerly and as a
    replayed CUDA graph (the uint8 target is then a static graph input)."""
    from oracle import augment_oracle as ao
    from passion_b200 import data
    from passion_b200.engine import Trainer
    shapes = [(24, 28, 20), (20, 24, 26)]
    vs = [ao.synth_volume(20 + i, s) for i, s in enumerate(shapes)]
    rc = _cases_on_device(vs)
    size = (16, 16, 16)
    mask = torch.tensor([[True, False, True, True], [False, True, True, False]])
    losses = []
    for mode in ("device", "host"):
        smp = data.AugmentSampler(size, py_rng=random.Random(2), np_rng=np.random.RandomState(3))
        trainer = Trainer(_model(torch.float32), lr=2e-4, modal_weight=torch.tensor([2.4, 1.6, 1.2, 5.1]), use_graph=use_graph)
        aug = data.DeviceAugment("cuda", size=size, batch=2)
        out = []
        for it in range(3):
            ps = [smp.sample(s) for s in shapes]
            if mode == "device":
                x, labels, _ = aug(rc, [0, 1], ps)
                loss, _ = trainer.step(x, labels, mask.cuda())
            else:
                items = [ao.apply(*vs[b], dict(ps[b], size=list(size), shift=ps[b]["shift"].reshape(1, 16, 1, 1, 4),
                                               scale=ps[b]["scale"].reshape(1, 16, 1, 1, 4))) for b in range(2)]
                x = torch.from_numpy(np.stack([i[0] for i in items])).cuda()
                target = torch.from_numpy(np.stack([i[2] for i in items])).cuda()
                loss, _ = trainer.step(x, target, mask.cuda())
            out.append(float(loss))
        losses.append(out)
    # BASELINE.json's per-step loss tolerance (1e-3); the two runs use the same kernels on the same label bytes
    assert np.all(np.isfinite(losses[0])) and np.allclose(losses[0], losses[1], rtol=1e-3, atol=1e-6), losses
