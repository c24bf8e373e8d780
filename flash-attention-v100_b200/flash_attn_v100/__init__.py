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


def install_flash_attn_shim(force: bool = False) -> bool:
    """Make `import flash_attn` (and `flash_attn.flash_attn_interface`, `flash_attn.bert_padding`,
    `flash_attn_2_cuda`) resolve to this build in the running interpreter, without installing anything -- the
    `sys.modules` patch the reference's unsloth demo applies by hand (utils/benchmarks/benchmark_unsloth.py:21-37),
    except that every name maps to its real implementation (the demo aliases all three entry points to
    flash_attn_func and stubs the padding helpers). Returns False, and changes nothing, if a different `flash_attn`
    is already imported and `force` is not set."""
    import importlib.util
    import os
    import sys

    shim_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shim")
    pkg_init = os.path.join(shim_root, "flash_attn", "__init__.py")
    if not os.path.exists(pkg_init):  # installed layout: the shim IS the installed flash_attn package
        import flash_attn  # noqa: F401

        return True
    present = sys.modules.get("flash_attn")
    if present is not None and os.path.abspath(getattr(present, "__file__", "") or "") == os.path.abspath(pkg_init):
        return True
    if present is not None and not force:
        return False
    for name in [n for n in sys.modules if n == "flash_attn" or n.startswith("flash_attn.") or n == "flash_attn_2_cuda"]:
        del sys.modules[name]

    def load(name, path, is_pkg=False):
        spec = importlib.util.spec_from_file_location(
            name, path, submodule_search_locations=[os.path.dirname(path)] if is_pkg else None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    pkg = load("flash_attn", pkg_init, is_pkg=True)
    pkg.flash_attn_interface = load("flash_attn.flash_attn_interface", os.path.join(shim_root, "flash_attn", "flash_attn_interface.py"))
    pkg.bert_padding = load("flash_attn.bert_padding", os.path.join(shim_root, "flash_attn", "bert_padding.py"))
    load("flash_attn_2_cuda", os.path.join(shim_root, "flash_attn_2_cuda.py"))
    return True

__all__ = [
    "flash_attn_func", "flash_attn_gpu",
    "flash_attn_varlen_func", "flash_attn_varlen_gpu",
    "flash_attn_with_kvcache", "flash_attn_with_kvcache_gpu",
    "install_flash_attn_shim",
]
