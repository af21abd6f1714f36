"""`UserInvariantInterestModel` (reference: models/user_invariant_interest_model.py:11-88):
owns the embedding tables, the w1 projection and the two pairwise attentions.

Registration order below is the reference's (:23-48) so that state_dict keys and seeded
initialisation agree.  The computation is the fused encoder of libnrm_b200
(`nrm_forward_encoder`: csrc/nrm_embed.cu + csrc/nrm_attention.cu)."""
import torch
import torch.nn as nn

from .. import engine
from ..config import config as model_config, GLOBAL_COLS, HIST_COLS, TGT_COLS
from .attention_model import PointwiseAttentionExpanded


class UserInvariantInterestModel(nn.Module):
    def __init__(self, embed_setting=[32, 16, 8, 8]):
        super().__init__()
        self.embed_setting = embed_setting
        n_sent = len(model_config['sentiment_label_dict'])
        # column groups of a packed row: time, pca, category, sub-categories, sentiment,
        # type, read_time, scroll (history rows carry all 8, candidate rows the first 6)
        self.slice_len_list = [4, model_config['pca_vector'], 1, model_config['subcategory_max_num'], n_sent, 1, 1, 1]
        width = sum(embed_setting)
        self.category_embedding = nn.Sequential(nn.Embedding(model_config['category_label_num'], embed_setting[0]))
        self.sentiment_embedding = nn.Sequential(nn.Linear(n_sent, embed_setting[1]), nn.ReLU())
        self.type_embedding = nn.Sequential(nn.Embedding(len(model_config['article_type_dict']), embed_setting[2]))
        self.w1 = nn.Linear(width + 2, width)
        self.year_embedding = nn.Sequential(nn.Embedding(100, embed_setting[3]))
        self.month_embedding = nn.Sequential(nn.Embedding(12 + 1, embed_setting[3]))
        self.day_embedding = nn.Sequential(nn.Embedding(31 + 1, embed_setting[3]))
        self.hour_embedding = nn.Sequential(nn.Embedding(24, embed_setting[3]))
        self.label_attention = PointwiseAttentionExpanded(width)
        self.text_img_attention = PointwiseAttentionExpanded(model_config['pca_vector'])

    def forward(self, x_history, x_target):
        """-> (eu_H [B,C,128], ec [B,C,128]) exactly as :73-88."""
        if list(self.embed_setting) != [32, 16, 8, 8]:
            raise NotImplementedError('the CUDA path is built for embed_setting=[32,16,8,8]')
        B, C = x_target.shape[0], x_target.shape[1]
        xg = torch.zeros(B, C, GLOBAL_COLS, dtype=torch.float64, device=x_target.device)
        named = {'invariant_interest_model.' + k: p for k, p in self.named_parameters()}
        e = engine.standalone_encoder(named, x_history, x_target, xg)
        return e[:, :, 0:128], e[:, :, 136:264]

    # ---- the reference's helper methods (:50-71), same signatures and results -------------------------------------------
    def slice_x(self, x, n):
        """Column groups of packed rows (:50-56): views, no arithmetic."""
        out, start = [], 0
        for i in range(n):
            out.append(x[:, :, start:start + self.slice_len_list[i]])
            start += self.slice_len_list[i]
        return out

    def _embed_rows(self, rows):
        """Run the fused row-embedding kernel on packed candidate rows [B,L,78] (float64) -> x_label_t [B,L,64] =
        [feature_embedding 56 | time_embedding 8] (the `ec` columns 136:200 of e_concat); one dummy history row per
        impression keeps the encoder's shape contract."""
        B = rows.shape[0]
        xh = torch.zeros(B, 1, HIST_COLS, dtype=torch.float64, device=rows.device)
        xg = torch.zeros(B, rows.shape[1], GLOBAL_COLS, dtype=torch.float64, device=rows.device)
        named = {'invariant_interest_model.' + k: p for k, p in self.named_parameters()}
        e = engine.standalone_encoder(named, xh, rows, xg)
        return e[:, :, 136:200]

    def feature_embedding(self, category, sub_category, sentiment, type):
        """(:58-64) category [.,.,1], sub_category [.,.,5], sentiment [.,.,3], type [.,.,1] -> [.,.,56]."""
        B, L = category.shape[0], category.shape[1]
        rows = torch.zeros(B, L, TGT_COLS, dtype=torch.float64, device=category.device)
        rows[:, :, 68:69] = category.to(torch.float64)
        rows[:, :, 69:74] = sub_category.to(torch.float64)
        rows[:, :, 74:77] = sentiment.to(torch.float64)
        rows[:, :, 77:78] = type.to(torch.float64)
        return self._embed_rows(rows)[:, :, 0:56].contiguous()

    def time_embedding(self, time):
        """(:66-71) [years, months, days, hours] buckets [.,.,4] -> [.,.,8] (sum of the four lookups)."""
        B, L = time.shape[0], time.shape[1]
        rows = torch.zeros(B, L, TGT_COLS, dtype=torch.float64, device=time.device)
        rows[:, :, 0:4] = time.to(torch.float64)
        return self._embed_rows(rows)[:, :, 56:64].contiguous()
