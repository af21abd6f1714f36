"""`UserInstantInterestModel` (reference: models/user_instant_interest_model.py:10-23):
ReLU(Linear(3, output_dim)) on the three per-candidate popularity scalars.  Inside
`UserModel` it is computed by `embed_rows_kernel` (csrc/nrm_embed.cu) while the candidate
rows are decoded; stand-alone it goes through the same encoder entry point."""
import torch
import torch.nn as nn

from .. import engine
from ..config import HIST_COLS, TGT_COLS


class UserInstantInterestModel(nn.Module):
    def __init__(self, output_dim):
        super().__init__()
        self.output_dim = output_dim
        self.out_fc = nn.Sequential(
            nn.Linear(3, output_dim),
            nn.ReLU(),
        )

    def forward(self, x_global):
        if self.output_dim != 8:
            raise NotImplementedError('the CUDA path is built for output_dim=8 (user_model.py:16)')
        B, C = x_global.shape[0], x_global.shape[1]
        xh = torch.zeros(B, 1, HIST_COLS, dtype=torch.float64, device=x_global.device)
        xt = torch.zeros(B, C, TGT_COLS, dtype=torch.float64, device=x_global.device)
        named = {'instant_interest_model.' + k: p for k, p in self.named_parameters()}
        e = engine.standalone_encoder(named, xh, xt, x_global)
        return e[:, :, 128:136]
