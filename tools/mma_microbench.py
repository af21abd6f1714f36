#!/usr/bin/env python
"""Cycles per tcgen05.mma / tcgen05.ld on this GPU (one CTA, one issuing thread)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from news_recommendation_model_b200 import _lib
lib = _lib.load()
out = torch.zeros(4, dtype=torch.int64, device='cuda')
names = {0: 'M64 N64 K-major', 1: 'M64 N64 MN-major', 2: 'M128 N64 K-major', 3: 'M64 N8 K-major', 4: 'tcgen05.ld x32', 5: '2 issuers M64 N64'}
for variant in (int(a) for a in (sys.argv[1:] or range(6))):
    for reps, nk in ((1, 4), (1, 12), (8, 4), (8, 12), (64, 4)):
        for _ in range(2):
            _lib.check(lib.nrm_debug_mma_microbench(ctypes.c_void_p(out.data_ptr()), variant, reps, nk, None), 'microbench')
            torch.cuda.synchronize()
        o = out.cpu().tolist()
        n = reps * (nk if variant != 4 else 1) * (2 if variant == 5 else 1)
        print(f'{names[variant]:18s} reps={reps:3d} nk={nk:2d}: issue {o[0]:7d} cyc, done {o[1]:7d} cyc  -> {o[1] / n:7.1f} cyc per op')
