"""Shim for the reference's `models/user_model.py`: re-exports the B200 implementation."""
from news_recommendation_model_b200.models.user_model import *  # noqa: F401,F403
