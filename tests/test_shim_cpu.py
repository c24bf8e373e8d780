"""The opt-in `flash_attn` impersonation shim and its padding helpers (CPU)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "flash-attention-v100_b200", "shim")


def test_shim_imports_in_a_clean_interpreter():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import flash_attn, flash_attn.flash_attn_interface as fi, flash_attn_2_cuda as ext\n"
        "import flash_attn_v100\n"
        "assert flash_attn.__version__ == '2.8.3'\n"
        "assert flash_attn.flash_attn_func is flash_attn_v100.flash_attn_func\n"
        "assert fi.flash_attn_with_kvcache is flash_attn_v100.flash_attn_with_kvcache\n"
        "assert callable(ext.fwd) and callable(ext.varlen_fwd) and callable(ext.fwd_kvcache)\n"
        "print('ok')\n" % SHIM)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_bert_padding_round_trip():
    import importlib.util

    spec = importlib.util.spec_from_file_location("_bp", os.path.join(SHIM, "flash_attn", "bert_padding.py"))
    bp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bp)
    torch.manual_seed(0)
    B, S, H, D = 3, 7, 2, 4
    x = torch.randn(B, S, H, D)
    lens = torch.tensor([7, 2, 5])
    mask = torch.arange(S)[None, :] < lens[:, None]
    tokens, idx, cu, mx, used = bp.unpad_input(x, mask)
    assert tokens.shape == (14, H, D) and mx == 7
    assert cu.dtype == torch.int32 and cu.tolist() == [0, 7, 9, 14] and used.tolist() == [7, 2, 5]
    assert torch.equal(tokens[7:9], x[1, :2])
    back = bp.pad_input(tokens, idx, B, S)
    assert torch.equal(back[mask], x[mask]) and back[~mask].abs().sum() == 0
