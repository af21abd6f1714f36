#!/usr/bin/env python
"""Per-role wait accounting of the tensor-core head kernels (CTA 0).  Needs a profile build:
     make -C news_recommendation_model_b200/csrc clean && make -C news_recommendation_model_b200/csrc -j8 EXTRA=-DNRM_RS_PROFILE
   then on the GPU box:  python tools/head_roleprof.py [fwd|bwd]"""
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

what = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
torch.cuda.set_device(0)
lib = _lib.load()
B, H, C = 1024, 50, 5
pool = [make_batch(B, H, C, seed=1234 + i, user_num=1000).to('cuda') for i in range(2)]
model = nrm.UserModel(1000)
model.load_state_dict(load_weights('train'), strict=False)
model.to('cuda').train().set_precision('bf16x3')
buf = (ctypes.c_longlong * 32)()
def run(n):
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
run(2)
_lib.check(lib.nrm_debug_headprof(buf), 'nrm_debug_headprof')
N = 4
run(N)
_lib.check(lib.nrm_debug_headprof(buf), 'nrm_debug_headprof')
roles = {0: ('loader', ['ring slot empty']), 1: ('mma issuer', ['operand ready', 'chunk landed']),
         2: ('epilogue warp 0', ['accumulator ready (66-wide)', 'accumulator ready (264-wide)', 'prologue', 'fence + arrive', 'tcgen05.ld'])}
for r, (name, kinds) in roles.items():
    tot = buf[r * 8 + 7] / N
    print(f'{name}: total {tot:.0f} cycles per launch (CTA 0)')
    for k, kn in enumerate(kinds):
        v = buf[r * 8 + k] / N
        print(f'    {kn:32s} {v:9.0f} cycles  {100 * v / max(tot, 1):5.1f} %')
print(f'entry -> tile loop: {buf[24] / N:.0f} cycles;  entry -> exit: {buf[25] / N:.0f} cycles')
