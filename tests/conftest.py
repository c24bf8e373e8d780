import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-v100_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fa_lib():
    """The C-ABI library; built in-tree on demand (nvcc cross-compiles without a GPU)."""
    lib_path = os.path.join(PKG, "lib", "libfa_b200.so")
    if not os.path.exists(lib_path):
        import __graft_entry__ as g

        g.build()
    import flash_attn_v100_cuda

    return flash_attn_v100_cuda.load_library()
