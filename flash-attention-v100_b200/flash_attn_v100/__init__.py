"""flash_attn_v100 -- the Python API of ai-bond/flash-attention-v100, served by the B200 build.

Same public names as the reference package (reference flash_attn_v100/__init__.py:1-18).
"""
from .flash_attn_interface import (
    flash_attn_func,
    flash_attn_gpu,
    flash_attn_varlen_func,
    flash_attn_varlen_gpu,
    flash_attn_with_kvcache,
    flash_attn_with_kvcache_gpu,
)

__version__ = "0.1.0+b200"

__all__ = [
    "flash_attn_func", "flash_attn_gpu",
    "flash_attn_varlen_func", "flash_attn_varlen_gpu",
    "flash_attn_with_kvcache", "flash_attn_with_kvcache_gpu",
]
