"""The extension module under its upstream name: the reference's build symlinks its .so to
`flash_attn_2_cuda` as well (reference setup.py:149-158, kernel/fused_mha_api.cpp:26-33)."""
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)
from flash_attn_v100_cuda import bwd, fwd, fwd_kvcache, varlen_bwd, varlen_fwd  # noqa: E402,F401
