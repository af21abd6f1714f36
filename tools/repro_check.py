#!/usr/bin/env python
"""Run-to-run reproducibility of one full-size training step (B=1024, H=50, C=5, bf16x3): same inputs, fresh gradients, N runs;
prints which gradient tensors differ between runs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import news_recommendation_model_b200 as nrm
from fixtures import load_weights
from news_recommendation_model_b200.synthetic import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 6
H = int(sys.argv[3]) if len(sys.argv) > 3 else 50
C = int(sys.argv[4]) if len(sys.argv) > 4 else 5
model = nrm.UserModel(1000)
model.load_state_dict(load_weights('train'), strict=False)
model.to('cuda').train().set_precision('bf16x3')
d = make_batch(B, H, C, seed=2024, user_num=1000, variable_history=H > 50).to('cuda')
import ctypes
from news_recommendation_model_b200 import _lib
lib = _lib.load()
FIELDS = ['de', 'dz', 'att_dhid', 'att_sc', 'dtp', 'dxh', 'dxt', 'dxin_h']
def snapshot():
    ws = model._runtime().train_ws
    out = {}
    for f in FIELDS:
        nb = ctypes.c_longlong()
        off = lib.nrm_debug_ws_field(B, H, C, 3, f.encode(), ctypes.byref(nb))
        out[f] = ws[off:off + nb.value].clone()
    return out
snaps = []
runs = []
for i in range(N):
    model.zero_grad(set_to_none=True)
    out = model(d.x_history, d.x_target, d.x_global)
    model.loss(d.user_id, out, d.label).backward()
    torch.cuda.synchronize()
    runs.append({k: v.grad.detach().clone() for k, v in model.named_parameters()})
    snaps.append(snapshot())
k0 = 'invariant_interest_model.w1.weight'
print('w1.weight equal to previous run:', [bool((runs[i][k0] == runs[i - 1][k0]).all().item()) for i in range(1, N)],
      'max|w1.weight| per run', ['%.6e' % runs[i][k0].abs().max().item() for i in range(N)])
for f in FIELDS:
    print(f'  workspace {f:9s} bytes differing from run 0:', [int((snaps[i][f] != snaps[0][f]).sum().item()) for i in range(1, N)])
d0 = snaps[0]['dxh'].view(torch.float32).view(B, H, 64); d1 = snaps[1]['dxh'].view(torch.float32).view(B, H, 64)
idx = (d0 != d1).nonzero()
if idx.numel():
    bs = idx[:, 0].unique(); hs = idx[:, 1].unique(); cs = idx[:, 2].unique()
    print('  dxh diffs: impressions', bs.numel(), bs[:24].tolist(), '\n   history rows', hs.tolist(), '\n   columns', cs.tolist())
    b0 = int(bs[0]); sub = idx[idx[:, 0] == b0]
    print('   impression', b0, ': rows', sub[:, 1].unique().tolist(), 'cols', sub[:, 2].unique().tolist(), 'max rel diff %.2e' % ((d0[b0] - d1[b0]).abs().max() / d0[b0].abs().max()).item())
    print('   units per CTA boundaries (148 CTAs):', [int(B * i // 148) for i in range(0, 12)])
bad = {}
for i in range(1, N):
    for k in runs[0]:
        dlt = (runs[i][k] - runs[0][k]).abs().max().item()
        if dlt != 0.0:
            bad[k] = max(bad.get(k, 0.0), dlt / max(runs[0][k].abs().max().item(), 1e-30))
print(f'B={B} H={H} C={C}', ' '.join(f'{k}={os.environ[k]}' for k in os.environ if k.startswith('NRM_')) or 'default', '->',
      'bitwise reproducible' if not bad else 'DIFFERS: ' + ', '.join(f'{k.split(".")[-2]}.{k.split(".")[-1]} {v:.1e}' for k, v in sorted(bad.items())))
