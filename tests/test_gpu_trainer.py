"""FusedTrainStep as a replacement of the train.py:48-49, 66-75 loop: learning-rate schedule, ragged last batch
(DataLoader without drop_last, train.py:40), optimizer checkpoint / resume."""
import numpy as np
import pytest
import torch

import news_recommendation_model_b200 as nrm
from fixtures import load_weights
from news_recommendation_model_b200.synthetic import Batch, make_batch
from oracle import reference_port as O
import parity as P

pytestmark = pytest.mark.gpu


def _fresh(U):
    m = nrm.UserModel(U)
    m.load_state_dict(load_weights('train'), strict=False)
    return m.to('cuda').train()


def _head(b: Batch, n):
    return Batch(*[getattr(b, f)[:n] for f in b.__dataclass_fields__])


def test_lr_schedule_and_ragged_last_batch_follow_the_oracle():
    """Three steps: full batch at lr 1e-3, full batch at lr 0.65e-3 (LambdaLR of train.py:49), then a ragged 37-impression
    tail -- against the oracle driven by torch.optim.Adam with the same schedule."""
    U, B, H, C = 50, 64, 50, 5
    batches = [make_batch(B, H, C, seed=600 + i, user_num=U) for i in range(3)]
    batches[2] = _head(batches[2], 37)
    model = _fresh(U)
    p = O.load_params(load_weights('train'), user_num=U)
    leaves = {k: p[k].requires_grad_(True) for k in O.TRAINABLE_KEYS + ('delta',)}
    opt = torch.optim.Adam(list(leaves.values()), lr=1e-3, weight_decay=1e-5)
    tr = nrm.FusedTrainStep(model, B, H, C, lr=1e-3, weight_decay=1e-5)
    for i, bb in enumerate(batches):
        if i == 1:
            opt.param_groups[0]['lr'] = 0.65e-3
            tr.set_lr(0.65e-3)
        out = O.user_model_forward(p, bb.x_history, bb.x_target, bb.x_global, training=True)
        loss = O.user_model_loss(p['delta'], bb.user_id, out, bb.label)
        loss.backward(); opt.step(); opt.zero_grad()
        got = tr.step(bb.pin()).item()
        assert abs(got - float(loss)) <= 1e-5, (i, got, float(loss))
    P.assert_weights_follow(model.named_parameters(), leaves, 3)
    assert int(model.bn.num_batches_tracked) == int(p['bn.num_batches_tracked'])


def test_optimizer_state_roundtrip_resumes_bit_for_bit():
    U, B, H, C = 50, 32, 20, 5
    batches = [make_batch(B, H, C, seed=700 + i, user_num=U).pin() for i in range(4)]
    a = _fresh(U)
    ta = nrm.FusedTrainStep(a, B, H, C, lr=1e-3, weight_decay=1e-5)
    for bb in batches[:2]:
        ta.step(bb)
    sd_model = {k: v.detach().clone() for k, v in a.state_dict().items()}
    sd_opt = ta.state_dict()
    assert sd_opt['step'] == 2 and abs(sd_opt['lr'] - 1e-3) < 1e-9
    for bb in batches[2:]:
        la = ta.step(bb)
    b = nrm.UserModel(U)
    b.load_state_dict(sd_model)
    b.to('cuda').train()
    tb = nrm.FusedTrainStep(b, B, H, C, lr=5e-2, weight_decay=0.0)      # wrong on purpose: load_state_dict must restore them
    tb.load_state_dict(sd_opt)
    for bb in batches[2:]:
        lb = tb.step(bb)
    assert la.item() == lb.item()
    assert torch.equal(a.flat_parameters().buf, b.flat_parameters().buf)
    assert torch.equal(a.bn.running_var, b.bn.running_var)
