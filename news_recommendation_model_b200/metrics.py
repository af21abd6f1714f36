"""Batched per-impression ranking metrics on the GPU.

The reference scores one impression at a time on the host: `train.py:77-80` calls
`tool/evaluation.py:auc_score` (sklearn `roc_auc_score`) once per sample of every training batch, and
`verify.py:25-37` once per validation impression.  `batch_metrics` does a whole batch in one kernel and
leaves the results on the device:

    m = nrm.metrics.batch_metrics(out, label)                 # out: [B,C] logits or scores on the GPU
    avg_auc = m['auc'].nanmean()                              # replaces the loop at train.py:77-80

AUC matches sklearn (ties averaged); `hit` is verify.py:32's argmax hit; MRR / nDCG@k are additions the
reference does not have (parity unpinned)."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib


def batch_metrics(scores: torch.Tensor, labels: torch.Tensor, n_valid: Optional[torch.Tensor] = None, k: int = 10) -> Dict[str, torch.Tensor]:
    """scores [B,C] (any float dtype, CUDA), labels [B,C] (1 = clicked), n_valid [B] = number of real (non-pad)
    candidates per impression (`C - empty_num` in the reference's records) or None.  Returns float32 [B] tensors
    'auc', 'hit', 'rr' (reciprocal rank), 'ndcg' (nDCG@k)."""
    if scores.dim() != 2 or labels.shape != scores.shape:
        raise ValueError('scores and labels must both be [B, C]')
    if not scores.is_cuda:
        raise _lib.NrmError('batch_metrics runs on CUDA tensors only (no CPU fallback)')
    dev = scores.device
    s = scores.detach().to(torch.float32)
    if s.stride(1) != 1:
        s = s.contiguous()
    y = labels.detach().to(device=dev, dtype=torch.float64)
    if y.stride(1) != 1:
        y = y.contiguous()
    B, C = s.shape
    nv = None
    if n_valid is not None:
        nv = n_valid.detach().to(device=dev, dtype=torch.int32).contiguous()
        if nv.shape != (B,):
            raise ValueError('n_valid must be [B]')
    out = {name: torch.empty(B, dtype=torch.float32, device=dev) for name in ('auc', 'hit', 'rr', 'ndcg')}
    if B == 0:
        return out
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(_lib.load().nrm_batch_metrics(p(s), s.stride(0), p(y), y.stride(0), p(nv), B, C, int(k), p(out['auc']), p(out['hit']),
                                            p(out['rr']), p(out['ndcg']), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
               'nrm_batch_metrics')
    return out


def auc_score(true_list, scores_list) -> float:
    """Drop-in for `tool/evaluation.py:auc_score` on CUDA tensors (one impression)."""
    return float(batch_metrics(scores_list.reshape(1, -1), true_list.reshape(1, -1))['auc'][0])


def list_auc_score(labels: torch.Tensor, scores: torch.Tensor, n_valid: Optional[torch.Tensor] = None) -> float:
    """`tool/evaluation.py:list_auc_score` (mean per-impression AUC) for a padded [B,C] batch."""
    return float(batch_metrics(scores, labels, n_valid)['auc'].nanmean())
