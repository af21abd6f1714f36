"""Model hyper-parameters of the EB-NeRD recommender hot path.

Same keys and values as the reference's module-level dict
(`/root/reference/configs/model_config.py:3-33`); the constructors of the
host-side modules read their vocabulary sizes and widths from here exactly as
the reference's do, so `UserModel()` built from this file has the reference's
parameter shapes and `state_dict` keys.
"""

_ARTICLE_TYPES = (
    'article_default', 'article_webtv', 'article_page_nine_girl',
    'article_questions_and_answers', 'article_feature', 'article_opinionen',
    'article_native', 'article_scribblelive', 'article_fullscreen_gallery',
    'article_editorial_production', 'article_standard_feature',
    'article_native_feature', 'article_accordion', 'article_video_standalone',
    'article_image_gallery', 'article_timeline',
)

config = {
    # category ids observed in EB-NeRD live in [2, 2975] (294 distinct)
    'category_label_num': 3000,
    'sentiment_label_dict': {'Negative': 0, 'Neutral': 1, 'Positive': 2},
    'article_type_dict': {name: i for i, name in enumerate(_ARTICLE_TYPES)},

    # ETL normalisers (only used by the offline pre-processor)
    'read_time_norm': 60,
    'scroll_norm': 100,
    'total_views_norm': 1e7,
    'total_read_time_norm': 1e9,

    'pca_vector': 64,
    'subcategory_max_num': 5,
    'history_max_num': 200,
    'inview_max_num': 15,
}

# ---- packed-row geometry (user_invariant_interest_model.py:14-22) ----------
# history row: [time4 | pca64 | cat1 | sub5 | sent3 | type1 | read_time1 | scroll1]
# target  row: the same without the last two columns
TIME_COLS = 4
PCA = config['pca_vector']
SUBCATS = config['subcategory_max_num']
SENT = len(config['sentiment_label_dict'])
HIST_COLS = TIME_COLS + PCA + 1 + SUBCATS + SENT + 1 + 1 + 1   # 80
TGT_COLS = HIST_COLS - 2                                        # 78
GLOBAL_COLS = 3
EMBED_SETTING = (32, 16, 8, 8)
LABEL_DIM = sum(EMBED_SETTING)                                  # 64
E_DIM = (LABEL_DIM + PCA) * 2 + 8                               # 264
assert HIST_COLS == 80 and TGT_COLS == 78 and E_DIM == 264
