"""ORACLE (test infrastructure, NOT product code).

A CPU restatement, in plain functional PyTorch, of the reference's hot path as
it is written: every function cites the reference lines it follows.  It is only
ever imported by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`; the product path
(`news_recommendation_model_b200`) never imports it and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors.  This
port is pinned instead against the unmodified reference modules imported from
`/root/reference` in the build container: `tests/golden/make_golden.py` runs
both on the same seeded inputs with the shipped checkpoints and commits the
reference's outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks
the port against those fixtures on every host (bit-for-bit on the build
container's torch 2.11 CPU kernels, 1e-6 elsewhere).

The arithmetic itself lives in PyTorch ATen kernels (unpinned by the reference;
container pin torch 2.11.0): `F.linear`, `F.embedding`, `F.gelu` (exact erf),
`F.batch_norm`, `softmax`, `binary_cross_entropy`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

II = 'invariant_interest_model.'

# state_dict keys in reference registration order; shapes in floats
STATE_KEYS: Tuple[Tuple[str, Tuple[int, ...]], ...] = (
    (II + 'category_embedding.0.weight', (3000, 32)),
    (II + 'sentiment_embedding.0.weight', (16, 3)),
    (II + 'sentiment_embedding.0.bias', (16,)),
    (II + 'type_embedding.0.weight', (16, 8)),
    (II + 'w1.weight', (64, 66)),
    (II + 'w1.bias', (64,)),
    (II + 'year_embedding.0.weight', (100, 8)),
    (II + 'month_embedding.0.weight', (13, 8)),
    (II + 'day_embedding.0.weight', (32, 8)),
    (II + 'hour_embedding.0.weight', (24, 8)),
    (II + 'label_attention.mlp.fc1.weight', (64, 256)),
    (II + 'label_attention.mlp.fc1.bias', (64,)),
    (II + 'label_attention.mlp.fc2.weight', (1, 64)),
    (II + 'label_attention.mlp.fc2.bias', (1,)),
    (II + 'text_img_attention.mlp.fc1.weight', (64, 256)),
    (II + 'text_img_attention.mlp.fc1.bias', (64,)),
    (II + 'text_img_attention.mlp.fc2.weight', (1, 64)),
    (II + 'text_img_attention.mlp.fc2.bias', (1,)),
    ('instant_interest_model.out_fc.0.weight', (8, 3)),
    ('instant_interest_model.out_fc.0.bias', (8,)),
    ('bn.weight', (264,)),
    ('bn.bias', (264,)),
    ('bn.running_mean', (264,)),
    ('bn.running_var', (264,)),
    ('bn.num_batches_tracked', ()),
    ('gate.fc1.weight', (66, 264)), ('gate.fc1.bias', (66,)),
    ('gate.fc2.weight', (264, 66)), ('gate.fc2.bias', (264,)),
    ('mlp.fc1.weight', (66, 264)), ('mlp.fc1.bias', (66,)),
    ('mlp.fc2.weight', (264, 66)), ('mlp.fc2.bias', (264,)),
    ('out_mlp.fc1.weight', (66, 264)), ('out_mlp.fc1.bias', (66,)),
    ('out_mlp.fc2.weight', (1, 66)), ('out_mlp.fc2.bias', (1,)),
)
BUFFER_KEYS = ('bn.running_mean', 'bn.running_var', 'bn.num_batches_tracked')
TRAINABLE_KEYS = tuple(k for k, _ in STATE_KEYS if k not in BUFFER_KEYS)

# column groups of a packed row (user_invariant_interest_model.py:14-22,50-56)
_SLICES = (4, 64, 1, 5, 3, 1, 1, 1)


def split_row(x: torch.Tensor, n: int) -> List[torch.Tensor]:
    """`slice_x` (user_invariant_interest_model.py:50-56)."""
    out, start = [], 0
    for width in _SLICES[:n]:
        out.append(x[:, :, start:start + width])
        start += width
    return out


def mlp(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """`MLP.forward` with the default exact-erf GELU (attention_model.py:29-32)."""
    hid = F.gelu(F.linear(x, p[prefix + 'fc1.weight'], p[prefix + 'fc1.bias']))
    return F.linear(hid, p[prefix + 'fc2.weight'], p[prefix + 'fc2.bias'])


def pairwise_attention(p: Params, prefix: str, target: torch.Tensor, history: torch.Tensor) -> torch.Tensor:
    """`PointwiseAttentionExpanded.forward` (attention_model.py:52-97): scores
    [B,C,H,1] = MLP(cat[h, t, t-h, t*h]) for all (candidate, history) pairs; no
    softmax, no mask.  Materialises the [B,C,H,4D] concat exactly as written."""
    if target.dim() == 2:
        target = target.unsqueeze(1)
    B, C, D = target.shape
    H = history.shape[1]
    t = target.unsqueeze(2)
    h = history.unsqueeze(1)
    cat = torch.cat([h.expand(-1, C, -1, -1), t.expand(-1, -1, H, -1), t - h, t * h], dim=-1)
    return mlp(p, prefix + 'mlp.', cat.view(-1, 4 * D)).view(B, C, H, 1)


def feature_embedding(p: Params, cat, sub, sent, typ) -> torch.Tensor:
    """`feature_embedding` (user_invariant_interest_model.py:58-64)."""
    table = p[II + 'category_embedding.0.weight']
    c = F.embedding(cat.reshape(-1, 1).to(torch.int64), table).reshape(cat.shape[0], cat.shape[1], -1)
    s = F.embedding(sub.reshape(-1, 1).to(torch.int64), table).reshape(sub.shape[0], sub.shape[1], sub.shape[2], -1).mean(dim=2)
    both = c + s
    se = F.relu(F.linear(sent.reshape(sent.shape[0] * sent.shape[1], -1),
                         p[II + 'sentiment_embedding.0.weight'], p[II + 'sentiment_embedding.0.bias']))
    se = se.reshape(sent.shape[0], sent.shape[1], -1)
    ty = F.embedding(typ.reshape(-1, 1).to(torch.int64), p[II + 'type_embedding.0.weight']).reshape(typ.shape[0], typ.shape[1], -1)
    return torch.cat((both, se, ty), dim=2)


def time_embedding(p: Params, time: torch.Tensor) -> torch.Tensor:
    """`time_embedding` (user_invariant_interest_model.py:66-71): zeros + year +
    month + day + hour, accumulated in that order."""
    acc = torch.zeros(time.shape[0], time.shape[1], 8, dtype=time.dtype, device=time.device)
    for i, name in enumerate(('year', 'month', 'day', 'hour')):
        idx = time[:, :, i:i + 1].reshape(acc.shape[0] * acc.shape[1], -1).to(torch.int64)
        acc = acc + F.embedding(idx, p[II + name + '_embedding.0.weight']).reshape(acc.shape[0], acc.shape[1], -1)
    return acc


def invariant_interest(p: Params, x_history: torch.Tensor, x_target: torch.Tensor,
                       dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """`UserInvariantInterestModel.forward` (user_invariant_interest_model.py:73-88)."""
    time_h, pca_h, cat_h, sub_h, sent_h, typ_h, read_h, scroll_h = split_row(x_history.to(dtype), 8)
    time_t, pca_t, cat_t, sub_t, sent_t, typ_t = split_row(x_target.to(dtype), 6)
    lab_h = torch.cat((feature_embedding(p, cat_h, sub_h, sent_h, typ_h), time_embedding(p, time_h), read_h, scroll_h), dim=2)
    lab_h = F.linear(lab_h.reshape(-1, lab_h.shape[2]), p[II + 'w1.weight'], p[II + 'w1.bias']).reshape(lab_h.shape[0], lab_h.shape[1], -1)
    lab_t = torch.cat((feature_embedding(p, cat_t, sub_t, sent_t, typ_t), time_embedding(p, time_t)), dim=2)
    ec = torch.cat((lab_t, pca_t), dim=2)
    s_lab = pairwise_attention(p, II + 'label_attention.', lab_t, lab_h)
    s_ti = pairwise_attention(p, II + 'text_img_attention.', pca_t, pca_h)
    pooled_lab = torch.sum(s_lab * lab_h.unsqueeze(1), dim=2)
    pooled_ti = torch.sum(s_ti * pca_h.unsqueeze(1), dim=2)
    return torch.cat((pooled_lab, pooled_ti), dim=2), ec


def instant_interest(p: Params, x_global: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """`UserInstantInterestModel.forward` (user_instant_interest_model.py:20-23)."""
    g = x_global.to(dtype)
    out = F.relu(F.linear(g.reshape(g.shape[0] * g.shape[1], -1),
                          p['instant_interest_model.out_fc.0.weight'], p['instant_interest_model.out_fc.0.bias']))
    return out.reshape(g.shape[0], g.shape[1], -1)


def e_concat(p: Params, x_history, x_target, x_global, dtype=torch.float32) -> torch.Tensor:
    """[eu_H | eu_L | ec] rows, [B,C,264] (user_model.py:28-31)."""
    eu_h, ec = invariant_interest(p, x_history, x_target, dtype)
    eu_l = instant_interest(p, x_global, dtype)
    return torch.cat((eu_h, eu_l, ec), dim=2)


def user_model_forward(p: Params, x_history, x_target, x_global, *, training: bool,
                       dtype=torch.float32, update_running_stats: bool = True) -> torch.Tensor:
    """`UserModel.forward` (user_model.py:27-35) -> logits [B,C].

    training=True uses batch statistics and (like nn.BatchNorm1d, momentum 0.1,
    eps 1e-5) updates `bn.running_mean/var` in `p` in place and increments
    `bn.num_batches_tracked`."""
    e = e_concat(p, x_history, x_target, x_global, dtype)
    B, C, _ = e.shape
    flat = e.reshape(B * C, -1)
    if training and update_running_stats:
        ctx = F.batch_norm(flat, p['bn.running_mean'], p['bn.running_var'], p['bn.weight'], p['bn.bias'],
                           True, 0.1, 1e-5)
        p['bn.num_batches_tracked'] += 1
    elif training:
        ctx = F.batch_norm(flat, None, None, p['bn.weight'], p['bn.bias'], True, 0.1, 1e-5)
    else:
        ctx = F.batch_norm(flat, p['bn.running_mean'], p['bn.running_var'], p['bn.weight'], p['bn.bias'],
                           False, 0.1, 1e-5)
    out = mlp(p, 'mlp.', mlp(p, 'gate.', ctx) * flat)     # gate has no sigmoid; multiplies the raw concat
    return mlp(p, 'out_mlp.', out).reshape(B, C)


def user_model_loss(delta: torch.Tensor, user_id: torch.Tensor, out: torch.Tensor, label: torch.Tensor,
                    alpha: float = 0.95) -> torch.Tensor:
    """`UserModel.loss` (user_model.py:37-43)."""
    y = label.to(out.dtype)
    plain = F.binary_cross_entropy(torch.softmax(out, dim=1), y)
    shift = delta[user_id].unsqueeze(1).repeat(1, label.shape[1])
    personal = F.binary_cross_entropy(torch.softmax(out + shift, dim=1), y)
    return (1 - alpha) * plain + alpha * personal


def adam_step(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, *,
              lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-5) -> None:
    """torch.optim.Adam single-tensor update with coupled L2 (train.py:48,74):
    g += wd*p; m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g;
    p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  In place."""
    g = grad + weight_decay * param
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2_sqrt = math.sqrt(1 - beta2 ** step)
    denom = (v.sqrt() / bc2_sqrt).add_(eps)
    param.addcdiv_(m, denom, value=-(lr / bc1))


# ---------------------------------------------------------------------------
# scoring epilogue and metrics (test.py:45-71,118-132; verify.py:25-42)
# ---------------------------------------------------------------------------

def score_batch(param_sets: Sequence[Params], x_history, x_target, x_global, empty_num: torch.Tensor,
                dtype=torch.float32) -> List[torch.Tensor]:
    """`model_test` inner loop (test.py:48-70) for one batch: trim the trailing
    candidate columns that are padding in every row, average the per-model
    softmax, then re-softmax the rows that still carry pad candidates (the
    reference's quirk).  Returns one 1-D score tensor per impression."""
    trim = int(torch.min(empty_num))
    if trim > 0:
        x_target = x_target[:, 0:-trim]
        x_global = x_global[:, 0:-trim]
        empty_num = empty_num - trim
    out = None
    for p in param_sets:
        s = torch.softmax(user_model_forward(p, x_history, x_target, x_global, training=False, dtype=dtype), dim=1)
        out = s if out is None else out + s
    out = out / len(param_sets)
    rows = []
    for i in range(out.shape[0]):
        z = int(empty_num[i])
        rows.append(torch.softmax(out[i:i + 1, 0:-z], dim=1).squeeze(0) if z > 0 else out[i])
    return rows


def rank_string(scores) -> str:
    """`get_string_of_prediction` ranking (test.py:124-129): 1-based rank of every
    candidate under a stable descending sort."""
    order = sorted(enumerate(list(scores)), key=lambda kv: kv[1], reverse=True)
    rank = ['-1'] * len(order)
    for r, (i, _) in enumerate(order):
        rank[i] = str(r + 1)
    return ','.join(rank)


def auc(labels, scores) -> float:
    """`tool/evaluation.py:3-5` (sklearn roc_auc_score, tie-averaged)."""
    from sklearn.metrics import roc_auc_score
    return float(roc_auc_score(labels, scores))


def load_params(state: Params, dtype=torch.float32, user_num: Optional[int] = None) -> Params:
    """Copy a reference `state_dict` (37 keys, `delta` absent: train.py:95-97)
    into an oracle parameter set; optionally adds a zero `delta[user_num+1]`
    (user_model.py:23)."""
    p = {}
    for k, shape in STATE_KEYS:
        t = state[k].detach().clone()
        assert tuple(t.shape) == shape, (k, tuple(t.shape), shape)
        p[k] = t if k == 'bn.num_batches_tracked' else t.to(dtype)
    if user_num is not None:
        p['delta'] = torch.zeros(user_num + 1, dtype=dtype)
    return p
