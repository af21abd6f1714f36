"""Shim for the reference's `models/attention_model.py`: re-exports the B200 implementation."""
from news_recommendation_model_b200.models.attention_model import *  # noqa: F401,F403
