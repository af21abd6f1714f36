"""GPU ranking metrics (next row N1) against the reference's metric code: sklearn roc_auc_score per impression
(tool/evaluation.py:3-5), numpy argmax hit (verify.py:32), stable descending rank (test.py:124-127)."""
import numpy as np
import pytest
import torch
from sklearn.metrics import roc_auc_score

import news_recommendation_model_b200 as nrm

pytestmark = pytest.mark.gpu


def _reference(scores, labels, n_valid, k):
    auc, hit, rr, ndcg = [], [], [], []
    for s, y, n in zip(scores, labels, n_valid):
        s, y = s[:n], y[:n]
        auc.append(roc_auc_score(y, s) if 0 < y.sum() < n else np.nan)
        hit.append(float(np.argmax(s) == np.argmax(y)))
        order = sorted(range(n), key=lambda i: s[i], reverse=True)          # stable: ties keep index order
        ranks = [order.index(i) + 1 for i in range(n) if y[i] > 0.5]
        rr.append(1.0 / min(ranks) if ranks else 0.0)
        dcg = sum(1.0 / np.log2(1 + r) for r in ranks if r <= k)
        idcg = sum(1.0 / np.log2(1 + i) for i in range(1, min(len(ranks), k) + 1))
        ndcg.append(dcg / idcg if idcg > 0 else 0.0)
    return np.array(auc), np.array(hit), np.array(rr), np.array(ndcg)


@pytest.mark.parametrize('B,C,ties', [(64, 5, False), (257, 15, True), (33, 100, True), (8, 1, False)])
def test_batch_metrics_match_the_reference_metric_code(B, C, ties):
    rng = np.random.default_rng(B * 7 + C)
    scores = rng.standard_normal((B, C)).astype(np.float32)
    if ties:
        scores = np.round(scores * 2) / 2                  # many exact ties
    n_valid = rng.integers(1, C + 1, B)
    labels = np.zeros((B, C))
    for b in range(B):
        labels[b, rng.integers(0, n_valid[b])] = 1.0
        if ties and n_valid[b] > 3 and b % 3 == 0:
            labels[b, rng.integers(0, n_valid[b])] = 1.0   # sometimes two positives
    k = 5
    m = nrm.metrics.batch_metrics(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(n_valid).cuda(), k=k)
    auc, hit, rr, ndcg = _reference(scores, labels, n_valid, k)
    got = {key: v.cpu().numpy() for key, v in m.items()}
    assert np.array_equal(np.isnan(got['auc']), np.isnan(auc))
    ok = ~np.isnan(auc)
    assert np.abs(got['auc'][ok] - auc[ok]).max(initial=0.0) <= 1e-6
    assert np.array_equal(got['hit'], hit)
    assert np.abs(got['rr'] - rr).max() <= 1e-6
    assert np.abs(got['ndcg'] - ndcg).max() <= 1e-5
    # mean AUC, as train.py:81-88 / verify.py:38-40 report it
    assert abs(float(m['auc'].nanmean()) - np.nanmean(auc)) <= 1e-4 if ok.any() else True


def test_auc_of_model_scores_agrees_with_the_oracle_path():
    """Config-1-sized batch: AUC from our logits vs AUC from the oracle's logits (north_star: agree to 1e-4)."""
    from fixtures import load_weights
    from news_recommendation_model_b200.synthetic import make_batch
    from oracle import reference_port as O
    b = make_batch(64, 50, 5, seed=9, user_num=100)
    model = nrm.UserModel(100); model.load_state_dict(load_weights('train'), strict=False)
    model.to('cuda').eval().set_precision('bf16x3')
    d = b.to('cuda')
    with torch.no_grad():
        out = model(d.x_history, d.x_target, d.x_global)
        ref = O.user_model_forward(O.load_params(load_weights('train'), user_num=100), b.x_history, b.x_target, b.x_global, training=False)
    ours = nrm.metrics.batch_metrics(out, d.label)['auc'].cpu().numpy()
    theirs = np.array([roc_auc_score(b.label[i].numpy(), ref[i].numpy()) for i in range(64)])
    assert np.abs(ours - theirs).max() <= 1e-4
