#!/usr/bin/env python
"""Development check of the row-stacked attention kernels (nrm_attention_rs.cu): eval / train logits against the oracle on a set
of shapes, and the kernel group timings of a B=1024 step.  Run on the GPU box:  timeout 300 python tools/rs_check.py [fwd|train]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import news_recommendation_model_b200 as nrm
from news_recommendation_model_b200 import _lib
from news_recommendation_model_b200.synthetic import make_batch
from fixtures import load_weights
from oracle import reference_port as O
import parity as P

mode = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
torch.cuda.set_device(0)
ok = True
shapes = [(8, 50, 5, {}), (3, 130, 3, dict(variable_history=True)), (5, 64, 19, dict(variable_candidates=True)), (2, 1, 1, {}),
          (7, 13, 4, dict(variable_history=True)), (16, 200, 40, dict(variable_history=True, variable_candidates=True)), (33, 256, 5, dict(variable_history=True))]
for prec in ('bf16x3', 'bf16'):
    for B, H, C, kw in shapes:
        b = make_batch(B, H, C, seed=B * 1000 + H, user_num=40, **kw)
        model, p = P.build_models(load_weights('validation'), 40, precision=prec)
        model.eval()
        d = b.to('cuda')
        with torch.no_grad():
            out = model(d.x_history, d.x_target, d.x_global).cpu()
            ref = O.user_model_forward(p, b.x_history, b.x_target, b.x_global, training=False)
        err = (out - ref).abs().max().item()
        tol = 1e-4 if prec == 'bf16x3' else 3e-2
        flag = 'ok' if err <= tol else 'FAIL'
        ok = ok and err <= tol
        print(f'{prec:7s} eval B={B} H={H} C={C}: max|dlogit| {err:.3e} {flag}', flush=True)
if mode == 'train':
    for B, H, C, kw in shapes[:5]:
        if B * C == 1:
            continue
        b = make_batch(B, H, C, seed=B * 1000 + H, user_num=40, **kw)
        delta0 = torch.from_numpy(np.random.default_rng(3).normal(0, 0.3, 41).astype(np.float32))
        model, p = P.build_models(load_weights('train'), 40, delta0, precision='bf16x3')
        rep = P.compare_step(model, p, b, training=True)
        bad = P.grad_failures(rep)
        good = rep['logits'] <= P.TOL_LOGITS and rep['loss'] <= P.TOL_LOSS and not bad
        ok = ok and good
        print(f'train B={B} H={H} C={C}: logits {rep["logits"]:.2e} loss {rep["loss"]:.2e} bad grads {bad[:3]} {"ok" if good else "FAIL"}', flush=True)
# timing at the bench shape
lib = _lib.load()
B, H, C = 1024, 50, 5
pool = [make_batch(B, H, C, seed=1234 + i, user_num=1000).to('cuda') for i in range(4)]
model, _ = P.build_models(load_weights('train'), 1000, precision='bf16x3')
model.train()
opt = nrm.FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
def step(b):
    out = model(b.x_history, b.x_target, b.x_global)
    loss = model.loss(b.user_id, out, b.label)
    loss.backward(); opt.step(); opt.zero_grad()
for i in range(4):
    step(pool[i % 4])
torch.cuda.synchronize()
lib.nrm_timing_enable(1)
for i in range(8):
    step(pool[i % 4])
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(8192)
lib.nrm_timing_report(buf, 8192)
lib.nrm_timing_enable(0)
for ln in buf.value.decode().strip().splitlines():
    name, cnt, tot = ln.split()[:3]
    print(f'  {name:32s} {float(tot) / 8 * 1e3:8.1f} us')
# whole step: FusedTrainStep, CUDA-graph replay, resident inputs (what bench.py reports as `value`)
host = [make_batch(B, H, C, seed=99 + i, user_num=1000).pin() for i in range(4)]
tr = nrm.FusedTrainStep(model, B, H, C, lr=1e-3, weight_decay=1e-5, nslots=4)
slots = [tr.load(hb) for hb in host]
torch.cuda.synchronize()
for i in range(8):
    tr.run(slots[i % 4])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(40):
    tr.run(slots[i % 4])
e1.record()
torch.cuda.synchronize()
print(f'STEP {e0.elapsed_time(e1) / 40 * 1e3:.1f} us per step (graph replay, resident), {tr.launches_per_step} launches')
print('ALL OK' if ok else 'SOME FAILED')
sys.exit(0 if ok else 1)
