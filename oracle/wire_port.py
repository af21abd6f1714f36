"""TEST INFRASTRUCTURE — host restatement of the compact-wire expansion (`nrm_expand_compact`), used only by tests/.

The packed layout it rebuilds is the reference ETL's (`tool/process_data.py:195-252`):
    x_history [B,H,80] = [time4 | pca64 | cat | sub5 | sent3 | type | read_time | scroll]   (:206-209)
    x_inview  [B,C,78] = the same without the last two                                      (:232-234)
    x_global  [B,C,3]  = [total_inviews, total_pageviews, total_read_time]                  (:240)
with the time buckets of `tool/normalization.py:31-39`.  The compact format itself has no counterpart in the reference
(parity is pinned through the packed tensors: expand(compact(records)) must equal float32(records))."""
import numpy as np
import torch


def unpack_time(t: np.ndarray) -> np.ndarray:
    t = t.astype(np.int64) & 0xffffffff
    return np.stack([t & 0xfff, (t >> 12) & 0xf, (t >> 16) & 0x1f, (t >> 21) & 0x1f], axis=-1).astype(np.float64)


def expand_reference(rows: np.ndarray, hist_article, hist_time, hist_click, cand_article, cand_time):
    """numpy arrays in, packed float64 (x_history, x_inview, x_global) out."""
    rows = np.asarray(rows, dtype=np.float32)
    ha, ca = np.asarray(hist_article), np.asarray(cand_article)
    ha = np.where((ha < 0) | (ha >= rows.shape[0]), 0, ha)
    ca = np.where((ca < 0) | (ca >= rows.shape[0]), 0, ca)
    B, H = ha.shape
    C = ca.shape[1]
    xh = np.zeros((B, H, 80)); xt = np.zeros((B, C, 78)); xg = np.zeros((B, C, 3))
    xh[:, :, 0:4] = unpack_time(np.asarray(hist_time).view(np.uint32))
    xh[:, :, 4:78] = rows[ha][:, :, 0:74]
    xh[:, :, 78:80] = np.asarray(hist_click, dtype=np.float32)
    xt[:, :, 0:4] = unpack_time(np.asarray(cand_time).view(np.uint32))
    xt[:, :, 4:78] = rows[ca][:, :, 0:74]
    xg[:] = rows[ca][:, :, 74:77]
    return torch.from_numpy(xh), torch.from_numpy(xt), torch.from_numpy(xg)
