#!/usr/bin/env python
"""One small training step + scoring epilogue + compact-wire expansion, meant to be run under compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import news_recommendation_model_b200 as nrm
from news_recommendation_model_b200 import wire
from news_recommendation_model_b200.synthetic import make_batch
from fixtures import load_weights

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
m = nrm.UserModel(50)
m.load_state_dict(load_weights('train'), strict=False)
m.to('cuda').train().set_precision(prec)
b = make_batch(7, 70, 9, seed=3, user_num=50, variable_history=True, variable_candidates=True).to('cuda')
opt = nrm.FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-5)
for _ in range(2):
    out = m(b.x_history, b.x_target, b.x_global)
    loss = m.loss(b.user_id, out, b.label)
    loss.backward()
    opt.step(); opt.zero_grad()
m.eval()
with torch.no_grad():
    s, r = nrm.scoring.ensemble_scores([m], b.x_history, b.x_target, b.x_global, b.empty_num)
    txt = nrm.scoring.submission_text(b.impression_id, r, b.empty_num)
t = wire.make_article_table(100)
cb = wire.make_compact_batch(t, 5, 20, 6, variable_history=True)
e = wire.expand(t.to('cuda'), cb.to('cuda'))
mt = nrm.metrics.batch_metrics(s, b.label, (s.shape[1] - b.empty_num))
torch.cuda.synchronize()
print('sanitize_step ok', float(loss), len(txt), float(mt['auc'].nanmean()))
