"""Fused Adam (north-star item d): torch.optim.Adam semantics (train.py:48 -- coupled
weight decay, no amsgrad) executed by `nrm_adam_step`.

When every parameter is a view into one flat buffer (a `UserModel` after its first
forward) and every gradient is the matching view of one flat gradient buffer (what the
hand-written backward hands to autograd), the whole model is updated with ONE launch
over the flat range; otherwise one launch per tensor.  State (`exp_avg`, `exp_avg_sq`,
`step`) is kept per tensor as views of flat state buffers so `state_dict()` has the
torch.optim.Adam layout.
"""
from __future__ import annotations

import torch

from . import engine


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError('invalid Adam hyper-parameter')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_scale=grad_scale))
        self._flat_state = {}     # id(group) -> (base_ptr, numel, exp_avg, exp_avg_sq)
        self._plan = {}           # id(group) -> cached layout of the flat fast path (see _step_cached)

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plan.clear()                 # the loaded per-tensor state is re-packed into flat buffers by the next step
        self._flat_state.clear()

    @staticmethod
    def _flat_span(plist):
        """If the tensors tile one contiguous float32 range in address order (gaps < 16 B
        allowed: alignment padding), return (base_ptr, numel_span)."""
        spans = sorted((p.data_ptr(), p.numel()) for p in plist)
        base = spans[0][0]
        end = base
        for ptr, n in spans:
            if ptr < end or ptr - end >= 16 or (ptr - base) % 4:
                return None
            end = ptr + 4 * n
        return base, (end - base) // 4

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            if self._step_cached(group):
                continue
            plist = [p for p in group['params'] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.is_sparse:
                    raise engine._lib.NrmError('FusedAdam handles dense float32 CUDA parameters only')
            b1, b2 = group['betas']
            kw = dict(lr=group['lr'], beta1=b1, beta2=b2, eps=group['eps'], weight_decay=group['weight_decay'],
                      grad_scale=group['grad_scale'])
            pspan = self._flat_span(plist) if len(plist) == len(group['params']) else None
            gspan = self._flat_span([p.grad for p in plist]) if pspan is not None else None
            fused = (pspan is not None and gspan is not None and pspan[1] == gspan[1] and
                     all(p.grad.data_ptr() - gspan[0] == p.data_ptr() - pspan[0] for p in plist))
            if fused:
                self._step_flat(group, plist, pspan, gspan, kw)
                first = min(range(len(plist)), key=lambda i: plist[i].data_ptr())
                gbase = plist[first].grad.data_ptr()
                self._plan[id(group)] = dict(pptrs=[p.data_ptr() for p in plist], goffs=[p.grad.data_ptr() - gbase for p in plist],
                                             first=first, n=pspan[1], pflat=_span_tensor(plist[first], pspan[1]))
            else:
                for p in plist:
                    st = self.state[p]
                    if len(st) == 0:
                        st['step'] = 0
                        st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                        st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['step'] += 1
                    g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                    engine.adam_step(p, g, st['exp_avg'], st['exp_avg_sq'], step=st['step'], **kw)
        return loss

    def _step_cached(self, group) -> bool:
        """Fast path for the steady state of a training loop: same parameter tensors as in the last fused step, and
        gradients that are again views of ONE flat buffer at the same offsets (what the hand-written backward hands to
        autograd every step).  Only pointers are compared; anything else falls back to the full analysis."""
        plan = self._plan.get(id(group))
        if plan is None:
            return False
        params = group['params']
        if len(params) != len(plan['pptrs']):
            return False
        g0 = params[plan['first']].grad
        if g0 is None:
            return False
        gbase = g0.data_ptr()
        goffs = plan['goffs']
        for i, p in enumerate(params):
            g = p.grad
            if g is None or p.data_ptr() != plan['pptrs'][i] or g.data_ptr() - gbase != goffs[i]:
                return False
        fs = self._flat_state[id(group)]
        state = self.state
        step = state[params[0]]['step'] + 1
        for p in params:
            state[p]['step'] = step
        b1, b2 = group['betas']
        engine.adam_step(plan['pflat'], _span_tensor(g0, plan['n']), fs[2], fs[3], step=step, lr=group['lr'], beta1=b1, beta2=b2,
                         eps=group['eps'], weight_decay=group['weight_decay'], grad_scale=group['grad_scale'])
        return True

    def _step_flat(self, group, plist, pspan, gspan, kw):
        base, n = pspan
        key = id(group)
        fs = self._flat_state.get(key)
        dev = plist[0].device
        if fs is None or fs[0] != base or fs[1] != n:
            m = torch.zeros(n, dtype=torch.float32, device=dev)
            v = torch.zeros(n, dtype=torch.float32, device=dev)
            for p in plist:       # carry over any per-tensor state (e.g. after load_state_dict)
                st = self.state[p]
                off = (p.data_ptr() - base) // 4
                mv, vv = m[off:off + p.numel()].view(p.shape), v[off:off + p.numel()].view(p.shape)
                if len(st) != 0:
                    mv.copy_(st['exp_avg']); vv.copy_(st['exp_avg_sq'])
                else:
                    st['step'] = 0
                st['exp_avg'], st['exp_avg_sq'] = mv, vv
            fs = (base, n, m, v)
            self._flat_state[key] = fs
        steps = {self.state[p]['step'] for p in plist}
        if len(steps) != 1:
            raise engine._lib.NrmError('FusedAdam: parameters of one flat group have diverging step counts')
        step = steps.pop() + 1
        for p in plist:
            self.state[p]['step'] = step
        # typed views over the raw spans (no copies): reuse the first tensors' storages
        pflat = _span_tensor(min(plist, key=lambda t: t.data_ptr()), n)
        gflat = _span_tensor(min((p.grad for p in plist), key=lambda t: t.data_ptr()), n)
        engine.adam_step(pflat, gflat, fs[2], fs[3], step=step, **kw)


def _span_tensor(first: torch.Tensor, n: int) -> torch.Tensor:
    """A 1-D view of `n` floats starting at `first`'s first element (same storage)."""
    return torch.as_strided(first.detach(), (n,), (1,), first.storage_offset())
