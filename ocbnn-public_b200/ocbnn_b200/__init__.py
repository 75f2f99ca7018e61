"""ocbnn_b200 — B200-native engine for OC-BNN posterior inference behind the reference's own API."""

from .models import *  # noqa: F401,F403
from .models import __all__ as _model_names

__all__ = list(_model_names)
