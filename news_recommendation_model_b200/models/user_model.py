"""`UserModel` -- the drop-in for the reference's top module
(`/root/reference/models/user_model.py:12-43`).

Same constructor (`UserModel(user_num=0)`), same `forward(x_history, x_target, x_global)
-> [B,C]` float32 logits, same `loss(id, out, label, alpha=0.95)`, same parameter names,
shapes and registration order, so `ckpt/ckpt_ebnerd_large_*.pth` load with
`load_state_dict(..., strict=False)` exactly as `test.py:160` / `verify.py:68` do.
The arithmetic runs in libnrm_b200 (hand-written sm_100a CUDA, C ABI in
include/nrm_b200.h); there is no CPU path."""
import torch
import torch.nn as nn

from .. import engine
from ..config import config as model_config
from .attention_model import MLP
from .user_instant_interest_model import UserInstantInterestModel
from .user_invariant_interest_model import UserInvariantInterestModel


class UserModel(nn.Module):
    def __init__(self, user_num=0):
        super().__init__()
        self.invariant_interest_model = UserInvariantInterestModel()
        self.instant_interest_model = UserInstantInterestModel(8)
        width = (sum(self.invariant_interest_model.embed_setting) + model_config['pca_vector']) * 2 \
            + self.instant_interest_model.output_dim
        self.bn = nn.BatchNorm1d(width)
        self.gate = MLP(self.bn.num_features, self.bn.num_features)
        self.mlp = MLP(self.bn.num_features, self.bn.num_features)
        self.out_mlp = MLP(self.bn.num_features, 1)
        self.delta = nn.Parameter(torch.zeros(user_num + 1))
        self.bce_loss = nn.BCELoss()
        self.softmax = nn.Softmax(dim=1)
        # runtime state (flat buffers, workspaces, data-parallel hooks): not part of state_dict
        self._rt = None
        self._dp = None
        # pair products on the tensor cores with hi/lo split operands: meets the fp32 tolerances (tests/test_gpu_tensorcore.py),
        # so the unmodified train.py / test.py get the fast path; 'fp32' selects the strict FFMA kernels
        self.precision = 'bf16x3'

    # ---- runtime plumbing --------------------------------------------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state['_rt'] = None          # device scratch / flat views are rebuilt lazily (test.py:177 pickles models)
        state['_dp'] = None
        return state

    def _runtime(self) -> engine.Runtime:
        if self._rt is None:
            self._rt = engine.Runtime()
        return self._rt

    def _precision_code(self) -> int:
        return engine.PRECISION[self.precision]

    def set_precision(self, precision: str):
        """'bf16x3' (default: tcgen05 tiles, hi/lo split operands, fp32-grade), 'fp32' (strict FFMA) or 'bf16' (tcgen05 tiles)."""
        if precision not in engine.PRECISION:
            raise ValueError(f'precision must be one of {sorted(engine.PRECISION)}')
        self.precision = precision
        return self

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() replace the parameter storages: the flat buffer has to be rebuilt
        rt = self.__dict__.get('_rt')
        if rt is not None:
            rt.flat = None
        return super()._apply(fn, *args, **kwargs)

    def flat_parameters(self) -> engine.FlatParams:
        """(Re)build the flat parameter buffer if .to()/load_state_dict moved the tensors.  Fast path: three sentinel
        tensors (first, middle and last entry of the layout) still point into the buffer -- the full walk over
        named_parameters() costs more host time than the smaller kernels of a step."""
        rt = self._runtime()
        flat = rt.flat
        if flat is not None:
            base = flat.buf.data_ptr()
            first = self.invariant_interest_model.category_embedding[0].weight
            if (self.delta.device == flat.device and first.data_ptr() == base
                    and self.bn.weight.data_ptr() == base + 4 * flat.bn_weight_off
                    and self.delta.data_ptr() == base + 4 * flat.fixed and self.delta.numel() == flat.delta_numel):
                return flat
        named = dict(self.named_parameters())
        if rt.flat is None or rt.flat.device != self.delta.device or not rt.flat.is_current(named):
            if not self.delta.is_cuda:
                raise engine._lib.NrmError('news_recommendation_model_b200 runs on CUDA devices only (no CPU '
                                           'fallback): call model.to("cuda") first')
            rt.flat = engine.FlatParams(named)
        return rt.flat

    # ---- reference API -----------------------------------------------------------------
    def forward(self, x_history, x_target, x_global):
        flat = self.flat_parameters()
        mode = engine.MODE_BN_BATCH_STATS if self.training else engine.MODE_EVAL
        params = flat.params
        if torch.is_grad_enabled() and (self.delta.requires_grad or any(p.requires_grad for p in params)):
            mode |= engine.MODE_KEEP_FOR_BWD
            return engine._ForwardFn.apply(self, x_history, x_target, x_global, mode, *params)
        rt = self._runtime()
        logits, _ = engine.forward_logits(rt, flat, self.bn.running_mean, self.bn.running_var,
                                          self.bn.num_batches_tracked, x_history, x_target, x_global, mode,
                                          self._precision_code(), self._dp)
        return logits

    def loss(self, id, out, label, alpha=0.95):
        self.flat_parameters()
        token = getattr(out, '_nrm_token', None)
        return engine._LossFn.apply(self, token, id, label, alpha, out, self.delta)


def getattr_path(module, dotted):
    obj = module
    for part in dotted.split('.'):
        obj = getattr(obj, part)
    return obj
