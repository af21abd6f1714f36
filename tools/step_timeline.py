#!/usr/bin/env python
"""Timeline of one eager training step (FusedTrainStep, use_graph=False) from the library's per-group CUDA events: start / end of
every kernel group relative to the first one, so that gaps between kernels and overlap with the side stream are visible."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import news_recommendation_model_b200 as nrm
from news_recommendation_model_b200 import _lib
from news_recommendation_model_b200.synthetic import make_batch
from fixtures import load_weights

torch.cuda.set_device(0)
lib = _lib.load()
B, H, C = 1024, 50, 5
model = nrm.UserModel(1000)
model.load_state_dict(load_weights('train'), strict=False)
model.to('cuda').train().set_precision('bf16x3')
host = [make_batch(B, H, C, seed=99 + i, user_num=1000).pin() for i in range(2)]
tr = nrm.FusedTrainStep(model, B, H, C, lr=1e-3, weight_decay=1e-5, nslots=2, use_graph=False)
slots = [tr.load(hb) for hb in host]
torch.cuda.synchronize()
for i in range(6):
    tr.run(slots[i % 2])
torch.cuda.synchronize()
lib.nrm_timing_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.run(slots[0])
e1.record()
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(16384)
lib.nrm_timing_report(buf, 16384)
lib.nrm_timing_enable(0)
rows = []
for ln in buf.value.decode().strip().splitlines():
    p = ln.split()
    rows.append((float(p[3]) * 1e3, float(p[4]) * 1e3, p[0]))
rows.sort()
print(f'eager step: {e0.elapsed_time(e1) * 1e3:.1f} us')
prev_end = 0.0
for b, e, name in rows:
    print(f'  {b:8.1f} -> {e:8.1f}  ({e - b:6.1f} us)  gap {b - prev_end:6.1f}   {name}')
    prev_end = max(prev_end, e)
