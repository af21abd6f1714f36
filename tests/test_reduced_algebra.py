"""The kernels' reduced-form algebra and hand-derived backward (oracle/reduced_mirror.py)
against autograd of the as-written port, in float64 (formula check) and float32 (what
re-association costs: this sets the fp32 tolerance used by the GPU parity tests)."""
import numpy as np
import pytest
import torch

from fixtures import load_weights
from news_recommendation_model_b200.synthetic import make_batch
from oracle import reference_port as O
from oracle import reduced_mirror as M


def _both(dtype, training, B=6, H=13, C=4, seed=5):
    b = make_batch(B, H, C, seed=seed, user_num=50, variable_history=True)
    p = O.load_params(load_weights('train'), dtype=dtype, user_num=50)
    p['delta'] = torch.from_numpy(np.random.default_rng(1).normal(0, 0.3, 51)).to(dtype)
    pa = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and not k.startswith('bn.running') else v.clone())
          for k, v in p.items()}
    out = O.user_model_forward(pa, b.x_history, b.x_target, b.x_global, training=training, dtype=dtype,
                               update_running_stats=False)
    loss = O.user_model_loss(pa['delta'], b.user_id, out, b.label)
    loss.backward()
    with torch.no_grad():
        r, l, g, _ = M.forward_backward(p, b.x_history, b.x_target, b.x_global, b.user_id, b.label,
                                        training=training, dtype=dtype)
    return pa, out.detach(), loss.detach(), r, l, g


@pytest.mark.parametrize('training', [True, False])
def test_formulas_exact_in_float64(training):
    pa, out, loss, r, l, g = _both(torch.float64, training)
    assert (out - r).abs().max() < 1e-11
    assert abs(loss - l) < 1e-12
    for k, grad in g.items():
        ref = pa[k].grad
        assert (grad - ref).abs().max() <= 1e-10 * max(1.0, ref.abs().max().item()), k


def test_reassociation_error_in_float32():
    pa, out, loss, r, l, g = _both(torch.float32, True, B=16, H=50, C=5)
    assert (out - r).abs().max() < 1e-4
    assert abs(loss - l) < 1e-5
    for k, grad in g.items():
        ref = pa[k].grad
        assert (grad - ref).abs().max() <= 2e-4 * max(1e-3, ref.abs().max().item()), k
