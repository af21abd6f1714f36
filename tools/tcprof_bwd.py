#!/usr/bin/env python
"""Phase cycle breakdown of the tensor-core attention BACKWARD (CTA 0, thread 0): build with
   make -C news_recommendation_model_b200/csrc clean; make ... EXTRA="-DNRM_TC_PROFILE -DNRM_TC_PROFILE_BWD" """
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import news_recommendation_model_b200 as nrm
from news_recommendation_model_b200 import _lib
from fixtures import load_weights
from news_recommendation_model_b200.synthetic import make_batch
lib = _lib.load()
names = ['sync before stage', 'stage history', 'ds product + pair vec', 'build W_c', 'sync + MMA#1 issue + wait', 'epilogue 1 (GELU, dhid)',
         'sync + batch-2 issue + wait', 'epilogue 2 (S^T, dt, Gt)', 'tile end (dH)', 'prologue (weights, TMEM)', 'per-CTA partial sums']
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
model = nrm.UserModel(1000); model.load_state_dict(load_weights('train'), strict=False)
model.to('cuda').train().set_precision(prec)
b = make_batch(1024, 50, 5, seed=1, user_num=1000).to('cuda')
buf = (ctypes.c_longlong * 32)()
for it in range(2):
    model.zero_grad(set_to_none=True)
    out = model(b.x_history, b.x_target, b.x_global)
    lib.nrm_debug_tcprof(buf)                      # clear: forward counters share the slots
    model.loss(b.user_id, out, b.label).backward()
_lib.check(lib.nrm_debug_tcprof(buf), 'tcprof')
for base, who in ((0, 'thread 0 (issues the MMAs)'), (16, 'thread 32')):
    tot = sum(buf[base:base + 16])
    print(f'{prec}: backward CTA 0, branch PROF_BRANCH (default label), {who}: {tot} cycles total')
    for i, n in enumerate(names):
        print(f'   {n:32s} {buf[base + i]:9d}  {100.0 * buf[base + i] / max(tot, 1):5.1f}%')
