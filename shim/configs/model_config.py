"""Shim for the reference's `configs/model_config.py` (same keys and values)."""
from news_recommendation_model_b200.config import config  # noqa: F401
