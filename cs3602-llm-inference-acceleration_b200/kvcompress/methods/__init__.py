"""Registry of compression methods — the drop-in plugin surface (reference methods/__init__.py:21-78)."""

from typing import Callable, Dict, List

from .l2_compress import l2_compress
from .fix_size_l2 import fix_size_l2_compress
from .streaming_llm import streaming_llm_compress, evict_for_space
from .recent_only import recent_only_compress
from .h2o_l2 import h2o_l2_compress
from .h2o_attention import h2o_attention_compress, H2OAttentionManager, create_h2o_manager_from_model
from .snapkv_lite import snapkv_lite_compress
from .pyramid_kv import pyramid_kv_compress
from .adaptive_l2 import adaptive_l2_compress

# name -> callable; insertion order is the order list_methods() reports (reference :21-33)
COMPRESS_METHODS: Dict[str, Callable] = {
    "l2_compress": l2_compress,
    "fix_size_l2": fix_size_l2_compress,
    "streaming_llm": streaming_llm_compress,
    "recent_only": recent_only_compress,
    "h2o_l2": h2o_l2_compress,
    "h2o_attention": h2o_attention_compress,
    "snapkv_lite": snapkv_lite_compress,
    "pyramid_kv": pyramid_kv_compress,
    "adaptive_l2": adaptive_l2_compress,
}


def get_compress_fn(method: str) -> Callable:
    """Look a method up by name; unknown names raise ``ValueError`` listing what exists (:36-61)."""
    if method not in COMPRESS_METHODS:
        available = list(COMPRESS_METHODS.keys())
        raise ValueError(f"Unknown method: {method}. Available: {available}")
    return COMPRESS_METHODS[method]


def list_methods() -> List[str]:
    """Registered method names (:64-66)."""
    return list(COMPRESS_METHODS.keys())


def register_method(name: str, fn: Callable) -> None:
    """Add or replace a method: ``fn(past_key_values, **kwargs) -> List[Tuple[Tensor, Tensor]]`` (:69-78)."""
    COMPRESS_METHODS[name] = fn


__all__ = [
    "l2_compress", "fix_size_l2_compress", "streaming_llm_compress", "evict_for_space", "recent_only_compress",
    "h2o_l2_compress", "h2o_attention_compress", "H2OAttentionManager", "create_h2o_manager_from_model",
    "snapkv_lite_compress", "pyramid_kv_compress", "adaptive_l2_compress",
    "get_compress_fn", "list_methods", "register_method", "COMPRESS_METHODS",
]
