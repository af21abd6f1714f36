#!/usr/bin/env python
"""Short driver for ncu: a few training-mode forwards (+ optionally backwards) at the bench shape B=1024, H=50, C=5, bf16x3."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import news_recommendation_model_b200 as nrm
from news_recommendation_model_b200.synthetic import make_batch
from fixtures import load_weights

what = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.cuda.set_device(0)
B, H, C = 1024, 50, 5
pool = [make_batch(B, H, C, seed=1234 + i, user_num=1000).to('cuda') for i in range(2)]
model = nrm.UserModel(1000)
model.load_state_dict(load_weights('train'), strict=False)
model.to('cuda').train().set_precision('bf16x3')
for i in range(n):
    b = pool[i % 2]
    if what == 'fwd':
        with torch.no_grad():
            model(b.x_history, b.x_target, b.x_global)
    else:
        out = model(b.x_history, b.x_target, b.x_global)
        model.loss(b.user_id, out, b.label).backward()
        model.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print('done')
