"""`flash_attn.flash_attn_interface` under the shim (reference flash_attn/flash_attn_interface.py:1-17)."""
from flash_attn_v100.flash_attn_interface import *  # noqa: F401,F403
from flash_attn_v100.flash_attn_interface import __all__  # noqa: F401
