#!/usr/bin/env python
"""Phase cycle breakdown of the tensor-core attention forward (needs `make -C .../csrc clean; make EXTRA=-DNRM_TC_PROFILE`)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import news_recommendation_model_b200 as nrm
from news_recommendation_model_b200 import _lib
from fixtures import load_weights
from news_recommendation_model_b200.synthetic import make_batch
lib = _lib.load()
names = ['sync before stage', 'stage history', 'pair vec + sync', 'build W_c', 'fence + sync', 'issue MMAs', 'MMA wait', 'tmem ld + GELU epilogue',
         'pool: park+sync (end of pair)', 'pool MMA + store']
for prec in sys.argv[1:] or ['bf16x3', 'bf16']:
    model = nrm.UserModel(1000); model.load_state_dict(load_weights('train'), strict=False)
    model.to('cuda').eval().set_precision(prec)
    b = make_batch(1024, 50, 5, seed=1, user_num=1000).to('cuda')
    buf = (ctypes.c_longlong * 32)()
    with torch.no_grad():
        model(b.x_history, b.x_target, b.x_global)
        lib.nrm_debug_tcprof(buf)
        model(b.x_history, b.x_target, b.x_global)
    _lib.check(lib.nrm_debug_tcprof(buf), 'tcprof')
    tot = sum(buf[:16])
    print(f'{prec}: CTA 0, both branches, {tot} cycles total')
    for i, n in enumerate(names):
        print(f'   {n:32s} {buf[i]:9d}  {100.0 * buf[i] / max(tot, 1):5.1f}%')
