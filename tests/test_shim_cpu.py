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


def _load_bp():
    import importlib.util

    spec = importlib.util.spec_from_file_location("_bp2", os.path.join(SHIM, "flash_attn", "bert_padding.py"))
    bp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bp)
    return bp


def test_bert_padding_exports_every_reference_name():
    """reference flash_attn/bert_padding.py:9-147"""
    bp = _load_bp()
    for name in ("IndexFirstAxis", "index_first_axis", "IndexPutFirstAxis", "index_put_first_axis",
                 "IndexFirstAxisResidual", "index_first_axis_residual", "unpad_input",
                 "unpad_input_for_concatenated_sequences", "pad_input"):
        assert hasattr(bp, name), name
    assert bp.index_first_axis == bp.IndexFirstAxis.apply or callable(bp.index_first_axis)


def test_bert_padding_autograd_functions_match_plain_indexing():
    bp = _load_bp()
    torch.manual_seed(1)
    x = torch.randn(10, 3, 4, dtype=torch.float64, requires_grad=True)
    idx = torch.tensor([7, 0, 3, 9])
    g = torch.randn(4, 3, 4, dtype=torch.float64)
    (gx,) = torch.autograd.grad(bp.index_first_axis(x, idx), x, g)
    (gx_ref,) = torch.autograd.grad(x[idx], x, g)
    assert torch.equal(bp.index_first_axis(x, idx), x[idx]) and torch.equal(gx, gx_ref)
    vals = torch.randn(4, 5, dtype=torch.float64, requires_grad=True)
    out = bp.index_put_first_axis(vals, idx, 10)
    ref = torch.zeros(10, 5, dtype=torch.float64).index_put((idx,), vals)
    go = torch.randn(10, 5, dtype=torch.float64)
    assert torch.equal(out, ref)
    assert torch.equal(torch.autograd.grad(out, vals, go)[0], torch.autograd.grad(ref, vals, go)[0])
    # residual variant: gathered rows + the untouched input; gradients of both branches add up
    y = torch.randn(6, 2, dtype=torch.float64, requires_grad=True)
    rows, resid = bp.index_first_axis_residual(y, torch.tensor([1, 4]))
    (gy,) = torch.autograd.grad((rows * 2).sum() + (resid * 3).sum(), y)
    expect = torch.full((6, 2), 3.0, dtype=torch.float64)
    expect[[1, 4]] += 2.0
    assert torch.equal(rows, y[[1, 4]]) and torch.equal(gy, expect)


def test_unpad_input_for_concatenated_sequences():
    """reference flash_attn/bert_padding.py:107-133: rows hold several samples; cu_seqlens has one entry per sample."""
    bp = _load_bp()
    x = torch.arange(3 * 6 * 2, dtype=torch.float32).reshape(3, 6, 2)
    lens = torch.tensor([[2, 3, 0, 0, 0, 0], [3, 2, 0, 0, 0, 0], [6, 0, 0, 0, 0, 0]])
    tokens, idx, cu, mx = bp.unpad_input_for_concatenated_sequences(x, lens)
    assert cu.tolist() == [0, 2, 5, 8, 10, 16] and cu.dtype == torch.int32 and mx == 6
    assert idx.tolist() == [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 12, 13, 14, 15, 16, 17]
    assert torch.equal(tokens, x.reshape(18, 2)[idx])


def test_install_flash_attn_shim_patches_sys_modules_in_a_clean_interpreter():
    """the sys.modules patch of the reference's unsloth demo (utils/benchmarks/benchmark_unsloth.py:21-37)"""
    pkg = os.path.join(ROOT, "flash-attention-v100_b200")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import flash_attn_v100\n"
        "assert flash_attn_v100.install_flash_attn_shim()\n"
        "import flash_attn, flash_attn.flash_attn_interface as fi, flash_attn.bert_padding as bp, flash_attn_2_cuda as ext\n"
        "assert flash_attn.__version__ == '2.8.3'\n"
        "assert fi.flash_attn_varlen_func is flash_attn_v100.flash_attn_varlen_func\n"
        "assert fi.flash_attn_with_kvcache is flash_attn_v100.flash_attn_with_kvcache\n"
        "assert callable(bp.unpad_input) and callable(bp.index_first_axis) and callable(ext.varlen_fwd)\n"
        "print('ok')\n" % pkg)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
