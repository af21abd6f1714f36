"""Seeded synthetic EB-NeRD-shaped batches (SURVEY.md §8d recipe).

The reference's ETL (`/root/reference/tool/process_data.py:195-252`) packs every
impression as float64 rows with categorical ids stored as doubles:

    x_history [B,H,80] = [time4 | pca64 | cat1 | sub5 | sent3 | type1 | read_time1 | scroll1]
    x_target  [B,C,78] = the same without the last two columns
    x_global  [B,C,3]  = normalised (total_inviews, total_pageviews, total_read_time)

Padded history rows / padded candidates are all-zero rows (process_data.py:198,
214-222).  This module draws batches with those conventions from a numpy PCG64
stream so that the same seed yields the same bytes on every host.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .config import HIST_COLS, TGT_COLS, GLOBAL_COLS, PCA, SUBCATS

# Per-component scale of the joint text+image PCA vector.  The shipped train
# checkpoint's bn.running_var[200:264] decays from 5.2e-2 to 3.4e-4;
# a geometric ramp between those two end points reproduces that spectrum.
_PCA_VAR = 5.2e-2 * (3.4e-4 / 5.2e-2) ** (np.arange(PCA) / (PCA - 1))
_PCA_STD = np.sqrt(_PCA_VAR)


@dataclass
class Batch:
    """One collated batch, laid out as `torch.utils.data.DataLoader` hands it to
    `train.py:66-71` / `test.py:46-56` (float64 features, int64 ids)."""
    impression_id: torch.Tensor   # [B] int64
    user_id: torch.Tensor         # [B] int64
    x_history: torch.Tensor       # [B,H,80] float64
    x_target: torch.Tensor        # [B,C,78] float64
    x_global: torch.Tensor        # [B,C,3]  float64
    label: torch.Tensor           # [B,C]    float64
    label_id: torch.Tensor        # [B,C]    float64 (-1 on pad candidates)
    empty_num: torch.Tensor       # [B]      int64  (# pad candidates)

    def to(self, device, non_blocking=False):
        return Batch(*[getattr(self, f).to(device, non_blocking=non_blocking)
                       for f in self.__dataclass_fields__])

    def pin(self):
        return Batch(*[getattr(self, f).pin_memory() for f in self.__dataclass_fields__])

    def input_bytes(self) -> int:
        return sum(t.numel() * t.element_size()
                   for t in (self.x_history, self.x_target, self.x_global, self.label, self.user_id))


def _item_rows(rng: np.random.Generator, n: int, cols: int, fp32_exact: bool) -> np.ndarray:
    x = np.zeros((n, cols), dtype=np.float64)
    # time buckets [years, months, days, hours] (tool/normalization.py:31-39)
    x[:, 0] = rng.integers(0, 3, n)
    x[:, 1] = rng.integers(0, 13, n)
    x[:, 2] = rng.integers(0, 31, n)
    x[:, 3] = rng.integers(0, 24, n)
    pca = rng.standard_normal((n, PCA)) * _PCA_STD
    if fp32_exact:
        pca = pca.astype(np.float32).astype(np.float64)
    x[:, 4:4 + PCA] = pca
    x[:, 68] = rng.integers(2, 2976, n)                      # category
    n_sub = rng.integers(0, SUBCATS + 1, n)
    sub = rng.integers(2, 2976, (n, SUBCATS))
    sub[np.arange(SUBCATS)[None, :] >= n_sub[:, None]] = 0     # trailing zeros = pad id
    x[:, 69:74] = sub
    pos = rng.integers(0, 3, n)
    score = rng.random(n)
    if fp32_exact:
        score = score.astype(np.float32).astype(np.float64)
    x[np.arange(n), 74 + pos] = score                        # one-hot position x score
    x[:, 77] = rng.integers(0, 16, n)                        # article type
    if cols == HIST_COLS:
        rs = rng.random((n, 2))
        if fp32_exact:
            rs = rs.astype(np.float32).astype(np.float64)
        x[:, 78:80] = rs                                     # read_time, scroll
    return x


def make_batch(batch: int, history: int, candidates: int, *, seed: int = 1234,
               user_num: int = 1000, variable_history: bool = False,
               variable_candidates: bool = False, fp32_exact: bool = False) -> Batch:
    """Draw one batch.

    variable_history    n_h ~ U{1..H}; rows >= n_h are all-zero (ETL padding).
    variable_candidates n_c ~ clipped lognormal in [5, C] (median ~11), zero padded
                        to C with label_id = -1 and empty_num set.
    fp32_exact          round the continuous columns to fp32-representable values
                        (used for compact golden fixtures).
    """
    rng = np.random.default_rng(seed)
    B, H, C = batch, history, candidates
    xh = _item_rows(rng, B * H, HIST_COLS, fp32_exact).reshape(B, H, HIST_COLS)
    xt = _item_rows(rng, B * C, TGT_COLS, fp32_exact).reshape(B, C, TGT_COLS)
    xg = rng.random((B, C, GLOBAL_COLS)) * 0.05
    if fp32_exact:
        xg = xg.astype(np.float32).astype(np.float64)

    if variable_history:
        n_h = rng.integers(1, H + 1, B)
        xh[np.arange(H)[None, :] >= n_h[:, None]] = 0.0
    if variable_candidates:
        n_c = np.clip(np.round(np.exp(rng.normal(np.log(11.0), 0.6, B))), 5, C).astype(np.int64)
        n_c = np.minimum(n_c, C)
    else:
        n_c = np.full(B, C, dtype=np.int64)
    pad = np.arange(C)[None, :] >= n_c[:, None]
    xt[pad] = 0.0
    xg[pad] = 0.0

    label = np.zeros((B, C), dtype=np.float64)
    label[np.arange(B), (rng.random(B) * n_c).astype(np.int64)] = 1.0
    label_id = rng.integers(9_000_000, 9_900_000, (B, C)).astype(np.float64)
    label_id[pad] = -1.0
    user = rng.integers(0, user_num + 1, B)
    imp = rng.integers(1, 1 << 30, B)
    t = torch.from_numpy
    return Batch(t(imp.astype(np.int64)), t(user.astype(np.int64)), t(xh), t(xt), t(xg),
                 t(label), t(label_id), t(pad.sum(1).astype(np.int64)))
