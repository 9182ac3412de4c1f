"""gcbfplus.algo surface (gcbfplus/algo/__init__.py:1-18)."""
from .base import MultiAgentController
from .gcbf_plus import GCBFPlus


def make_algo(algo: str, **kwargs) -> MultiAgentController:
    """gcbfplus/algo/__init__.py:8-18.  Only 'gcbf+' is in the hot-path scope."""
    if algo == "gcbf+":
        return GCBFPlus(**kwargs)
    if algo in ("gcbf", "centralized_cbf", "dec_share_cbf"):
        raise NotImplementedError(f"algo '{algo}' is outside the B200 hot-path scope (SURVEY.md section 2, rows 12/14)")
    raise ValueError(f"Unknown algorithm: {algo}")
