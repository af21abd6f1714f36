#!/usr/bin/env python
"""Small driver for ncu: a few training steps of the drop-in module path on one synthetic batch.
   python tools/profile_step.py [--precision bf16x3] [--batch 1024] [--steps 3]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import news_recommendation_model_b200 as nrm
from fixtures import load_weights
from news_recommendation_model_b200.synthetic import make_batch

ap = argparse.ArgumentParser()
ap.add_argument('--precision', default='bf16x3')
ap.add_argument('--batch', type=int, default=1024)
ap.add_argument('--history', type=int, default=50)
ap.add_argument('--candidates', type=int, default=5)
ap.add_argument('--steps', type=int, default=3)
a = ap.parse_args()

torch.cuda.set_device(0)
model = nrm.UserModel(1000)
model.load_state_dict(load_weights('train'), strict=False)
model.to('cuda').train().set_precision(a.precision)
opt = nrm.FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
b = make_batch(a.batch, a.history, a.candidates, seed=1, user_num=1000).to('cuda')
for i in range(a.steps):
    out = model(b.x_history, b.x_target, b.x_global)
    loss = model.loss(b.user_id, out, b.label)
    loss.backward()
    opt.step()
    opt.zero_grad()
torch.cuda.synchronize()
print('loss', float(loss))
