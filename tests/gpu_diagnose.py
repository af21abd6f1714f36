#!/usr/bin/env python
"""Print per-tensor GPU-vs-oracle errors for a few shapes (for reading back from a gpurun log)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from fixtures import load_weights          # noqa: E402
from news_recommendation_model_b200.synthetic import make_batch   # noqa: E402
import parity as P                         # noqa: E402


def main():
    shapes = [(16, 50, 5, {}), (7, 13, 4, dict(variable_history=True)), (3, 130, 3, dict(variable_history=True)),
              (5, 64, 19, dict(variable_candidates=True))]
    for B, H, C, kw in shapes:
        b = make_batch(B, H, C, seed=B * 1000 + H, user_num=40, **kw)
        delta0 = torch.from_numpy(np.random.default_rng(3).normal(0, 0.3, 41).astype(np.float32))
        model, p = P.build_models(load_weights('train'), 40, delta0)
        try:
            rep = P.compare_step(model, p, b, training=True)
            torch.cuda.synchronize()
            print(f'=== B={B} H={H} C={C} {kw}')
            print(P.format_report(rep))
            bad = P.grad_failures(rep)
            print('FAILED tensors:', [k for k, _, _ in bad] if bad else 'none',
                  '| logits ok' if rep['logits'] <= P.TOL_LOGITS else '| LOGITS BAD')
        except Exception as exc:          # keep going so one run shows everything
            print(f'=== B={B} H={H} C={C} {kw}: EXCEPTION {type(exc).__name__}: {exc}')
    sys.stdout.flush()


if __name__ == '__main__':
    main()
