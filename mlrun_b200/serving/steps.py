"""convenience re-export (the transforms live in mlrun_b200.feature_store)"""
from ..feature_store.transforms import *  # noqa: F401,F403
