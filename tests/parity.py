"""Shared GPU-vs-oracle comparison used by the -m gpu tests, smoke() and the diagnostic
script: runs one train.py:69-75 step through the CUDA path and through the oracle port on
the same inputs and reports per-tensor errors."""
from __future__ import annotations

import numpy as np
import torch

import news_recommendation_model_b200 as nrm
from oracle import reference_port as O

# fp32 tolerances (tests/test_reduced_algebra.py measures the re-association floor at
# ~3e-6 on logits and ~1e-5 of max|grad| per tensor between two fp32 evaluations)
TOL_LOGITS = 1e-4          # absolute; logits reach |3|
TOL_LOSS = 1e-5
TOL_GRAD_REL = 2e-4        # of max|grad| of the tensor
TOL_GRAD_ABS = 1e-7        # delta / out_mlp.fc2.bias gradients are pure rounding noise (softmax shift invariance)
NOISE_KEYS = ('delta', 'out_mlp.fc2.bias')


def build_models(weights, user_num, delta0=None, device='cuda', precision='fp32'):
    """precision: 'fp32' = the strict FFMA kernels (callers switch with model.set_precision for the tensor-core paths)."""
    model = nrm.UserModel(user_num).set_precision(precision)
    model.load_state_dict(weights, strict=False)
    p = O.load_params(weights, user_num=user_num)
    if delta0 is not None:
        with torch.no_grad():
            model.delta.copy_(delta0)
        p['delta'] = delta0.clone()
    return model.to(device), p


def oracle_step(p, batch, training=True):
    leaves = {k: p[k].requires_grad_(True) for k in O.TRAINABLE_KEYS + ('delta',)}
    out = O.user_model_forward(p, batch.x_history, batch.x_target, batch.x_global, training=training)
    loss = O.user_model_loss(p['delta'], batch.user_id, out, batch.label)
    loss.backward()
    grads = {k: v.grad.detach().clone() for k, v in leaves.items()}
    return out.detach(), loss.detach(), grads


def cuda_step(model, batch, device='cuda'):
    d = batch.to(device)
    model.zero_grad(set_to_none=True)
    out = model(d.x_history, d.x_target, d.x_global)
    loss = model.loss(d.user_id, out, d.label)
    loss.backward()
    grads = {k: (v.grad.detach().cpu() if v.grad is not None else None) for k, v in model.named_parameters()}
    return out.detach().cpu(), loss.detach().cpu(), grads


def compare_step(model, p, batch, training=True, device='cuda'):
    """-> dict of error numbers; raises nothing (callers assert)."""
    model.train(training)
    out_c, loss_c, g_c = cuda_step(model, batch, device)
    out_o, loss_o, g_o = oracle_step(p, batch, training)
    rep = {'logits': float((out_c - out_o).abs().max()), 'loss': float((loss_c - loss_o).abs()),
           'loss_value': float(loss_o), 'grads': {}}
    for k, go in g_o.items():
        gc = g_c[k]
        scale = float(go.abs().max())
        err = float('inf') if gc is None else float((gc - go).abs().max())
        rep['grads'][k] = (err, scale)
    return rep


def grad_failures(rep):
    bad = []
    for k, (err, scale) in rep['grads'].items():
        tol = TOL_GRAD_ABS if k in NOISE_KEYS else TOL_GRAD_REL * scale + TOL_GRAD_ABS
        if not err <= tol:
            bad.append((k, err, scale))
    return bad


def format_report(rep) -> str:
    lines = [f"logits max|d| {rep['logits']:.3e}   loss |d| {rep['loss']:.3e} (loss {rep['loss_value']:.6f})"]
    for k, (err, scale) in rep['grads'].items():
        lines.append(f"  {k:58s} err {err:.3e}  max|g| {scale:.3e}  rel {err / max(scale, 1e-30):.2e}")
    return '\n'.join(lines)


def assert_weights_follow(named_params, ref, steps, lr=1e-3):
    """Weights after `steps` Adam steps against a reference trajectory.  Adam moves every element by ~lr per step whatever the
    size of its gradient, so an element whose gradient is far below its tensor's largest one (a nearly dead ReLU channel of the
    instant-interest Linear, say) follows the SIGN of fp32 re-association noise and can drift by a few lr between two correct
    implementations.  Hence two bounds: 99.9 % of all elements within 1 % of the steps * lr a weight can travel (at least 2e-5),
    and no element further than a quarter of it.  delta / out_mlp.fc2.bias (pure-noise gradients) are left out."""
    errs, worst = [], ('', 0.0)
    for k, v in named_params:
        if k in NOISE_KEYS:
            continue
        e = (v.detach().cpu().float() - ref[k].detach().float()).abs().flatten()
        errs.append(e)
        if e.numel() and e.max().item() > worst[1]:
            worst = (k, e.max().item())
    errs = torch.cat(errs)
    pick = torch.randperm(errs.numel(), generator=torch.Generator().manual_seed(0))[:200000]
    q999 = torch.quantile(errs[pick].double(), 0.999).item()
    travel = steps * lr
    assert q999 <= max(2e-5, 0.01 * travel), ('99.9 % quantile', q999, worst)
    assert worst[1] <= 0.25 * travel, ('max', worst)
