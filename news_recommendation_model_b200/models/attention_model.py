"""Parameter containers for the attention / head MLPs.

Mirrors `/root/reference/models/attention_model.py` (MLP :10-32, PointwiseAttention
:34-44, PointwiseAttentionExpanded :47-97): same constructor arguments, same
registration order (fc1, fc2 -- every MLP also instantiates the six parameter-free
activation modules so that seeded initialisation consumes the RNG identically) and
therefore the same state_dict keys.

Inside `UserModel` these modules only OWN the weights: the arithmetic of the two
attention instances (256->64->1 over all candidate x history pairs) and of the three
head MLPs runs in the fused CUDA path (`csrc/nrm_attention.cu`, `csrc/nrm_head.cu`),
which reads the weights from the model's flat parameter buffer.  The `forward`
methods below are the stand-alone API of the reference classes (used by none of its
callers); they express the same math with torch tensor ops for arbitrary widths.
"""
import torch
import torch.nn as nn

_ACTIVATIONS = ('relu', 'gelu', 'tanh', 'sigmoid', 'leaky_relu', 'elu')


class MLP(nn.Module):
    """in -> in//4 -> out with a selectable activation (default exact-erf GELU;
    unknown names fall back to GELU as in attention_model.py:18-27)."""

    def __init__(self, input_dim, output_dim, activation_type='gelu'):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, input_dim // 4)
        self.fc2 = nn.Linear(input_dim // 4, output_dim)
        self.activation_type = activation_type
        table = {'relu': nn.ReLU(), 'gelu': nn.GELU(), 'tanh': nn.Tanh(), 'sigmoid': nn.Sigmoid(),
                 'leaky_relu': nn.LeakyReLU(), 'elu': nn.ELU()}
        self.activation = table.get(str(activation_type).lower(), nn.GELU())

    def forward(self, x):
        return self.fc2(self.activation(self.fc1(x)))


class PointwiseAttention(nn.Module):
    """score = MLP([h, t, t-h, t*h]) for aligned (target, history) rows."""

    def __init__(self, input_dim):
        super().__init__()
        self.mlp = MLP(input_dim * 4, 1)

    def forward(self, target, history):
        return self.mlp(torch.cat([history, target, target - history, target * history], dim=-1))


class PointwiseAttentionExpanded(nn.Module):
    """All-pairs scores [B,C,H,1] between `target` [B,C,D] (or [B,D]) and `history`
    [B,H,D].  Stand-alone form: evaluated in the reduced algebra of DESIGN.md section 3
    (no [B,C,H,4D] concat): fc1([h,t,t-h,t*h]) = h(Wa-Wc)^T + t(Wb+Wc)^T + (t*h)Wd^T + b."""

    def __init__(self, input_dim):
        super().__init__()
        self.mlp = MLP(input_dim * 4, 1)

    def forward(self, target, history):
        if target.dim() == 2:
            target = target.unsqueeze(1)
        d = target.shape[-1]
        w = self.mlp.fc1.weight
        wa, wb, wc, wd = w[:, 0:d], w[:, d:2 * d], w[:, 2 * d:3 * d], w[:, 3 * d:4 * d]
        hp = history @ (wa - wc).t()                                   # [B,H,J]
        tp = target @ (wb + wc).t() + self.mlp.fc1.bias               # [B,C,J]
        pair = torch.einsum('bck,bhk,jk->bchj', target, history, wd)
        hid = self.mlp.activation(pair + hp.unsqueeze(1) + tp.unsqueeze(2))
        return self.mlp.fc2(hid)
