"""`flash_attn` impersonation shim: lets code that does `import flash_attn` pick up the B200 kernels.

Same role as the reference's flash_attn/ package (reference flash_attn/__init__.py:1-27), which re-exports
its own interface under the upstream package name and reports the upstream version it imitates (:15).
Opt-in: put `flash-attention-v100_b200/shim` on sys.path (ahead of any real flash-attn install).
"""
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from flash_attn_v100 import (  # noqa: E402
    flash_attn_func,
    flash_attn_gpu,
    flash_attn_varlen_func,
    flash_attn_varlen_gpu,
    flash_attn_with_kvcache,
    flash_attn_with_kvcache_gpu,
)

__version__ = "2.8.3"  # the upstream release the reference impersonates

__all__ = [
    "flash_attn_func", "flash_attn_gpu",
    "flash_attn_varlen_func", "flash_attn_varlen_gpu",
    "flash_attn_with_kvcache", "flash_attn_with_kvcache_gpu",
]
