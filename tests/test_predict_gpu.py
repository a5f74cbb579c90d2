"""GPU: sliding-window inference (BASELINE.json configs[4]) against the oracle's restatement of
utils/predict.py:181-218, and the shared-encoder 15-mask sweep against 15 independent passes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(dtype=torch.float32):
    from oracle import synth
    from passion_b200.models import rfnet
    sd = synth.make_state_dict(1037)
    m = rfnet.Model(4).cuda()
    m.load_state_dict(sd)
    m.compute_dtype = dtype
    m.is_training = False
    return m, sd


def test_argmax_operator_bit_exact():
    """first-max-index tie-break, as torch.argmax in predict.py:218"""
    from oracle.masks import sliding_window_argmax
    p = torch.tensor([[0.25, 0.25, 0.25, 0.25], [0.1, 0.4, 0.4, 0.1], [0.0, 0.0, 1.0, 0.0]]).t().reshape(1, 4, 3, 1, 1)
    assert torch.argmax(p.cuda(), dim=1).flatten().tolist() == [0, 1, 2]
    assert p.numpy().argmax(1).flatten().tolist() == [0, 1, 2]


def test_sliding_window_matches_oracle(lib_built):
    from oracle import rfnet_oracle
    from oracle.masks import sliding_window_argmax
    from passion_b200.predict import predict_volume, window_origins
    assert window_origins(240, 80) == [0, 40, 80, 120, 160] and window_origins(155, 128) == [0, 27]
    model, sd = _model()
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.standard_normal((1, 4, 24, 24, 20)).astype(np.float32))
    for pattern in ([True, False, True, True], [False, True, False, False]):
        mask = torch.tensor([pattern])

        def prob_fn(win):
            with torch.no_grad():
                return rfnet_oracle.forward(sd, torch.from_numpy(np.ascontiguousarray(win)), mask, is_training=False).numpy()
        ref = sliding_window_argmax(prob_fn, x.numpy(), 16)
        labels, prob = predict_volume(model, x.cuda(), mask.cuda(), patch_size=16)
        agree = float((labels.cpu().numpy() == ref).mean())
        assert agree >= 0.999, agree          # fp32: identical up to exact near-ties of the averaged probabilities


def test_all_masks_sweep_equals_independent_passes(lib_built):
    from passion_b200.predict import MASKS_TEST, predict_all_masks, predict_volume
    model, _ = _model()
    rs = np.random.RandomState(6)
    x = torch.from_numpy(rs.standard_normal((1, 4, 24, 16, 20)).astype(np.float32)).cuda()
    labels, prob = predict_all_masks(model, x, patch_size=16)
    assert labels.shape == (15, 24, 16, 20)
    for i, pattern in enumerate(MASKS_TEST):
        l1, p1 = predict_volume(model, x, torch.tensor([pattern]).cuda(), patch_size=16)
        # batch-15 and batch-1 launches group their fp32 partial sums differently (grid sizes depend on the batch)
        assert torch.allclose(prob[i], p1[0], atol=1e-4), i
        assert float((labels[i] == l1[0]).float().mean()) >= 0.999, i
