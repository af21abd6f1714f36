"""Host-side mirror of the reference's `models/` package (same class names, constructor
arguments, forward signatures and state_dict keys), computing through libnrm_b200."""
from .attention_model import MLP, PointwiseAttention, PointwiseAttentionExpanded
from .user_instant_interest_model import UserInstantInterestModel
from .user_invariant_interest_model import UserInvariantInterestModel
from .user_model import UserModel

__all__ = ['MLP', 'PointwiseAttention', 'PointwiseAttentionExpanded', 'UserInstantInterestModel',
           'UserInvariantInterestModel', 'UserModel']
