"""CPU-side checks of the drop-in boundary: the header, the ctypes mirror and the built library agree,
every declared symbol is exported, and argument errors come back through the C error convention.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fa_b200.h")

C2CT = {"int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "uint64_t": ctypes.c_uint64}


def _header_fields():
    src = open(HEADER).read()
    body = src[src.index("typedef struct fa_b200_params {"):src.index("} fa_b200_params_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split("{", 1)[1].split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        m = re.match(r"(const )?(\w+)(\s*\*)?\s*(.*)", stmt)
        ctype, is_ptr, names = m.group(2), bool(m.group(3)), m.group(4)
        for name in names.split(","):
            name = name.strip()
            ptr = is_ptr or name.startswith("*")
            fields.append((name.lstrip("* "), ctypes.c_void_p if ptr else C2CT[ctype]))
    return fields


def test_ctypes_struct_mirrors_header(fa_lib):
    import flash_attn_v100_cuda as op

    want = _header_fields()
    got = [(n, t) for n, t in op.FaB200Params._fields_]
    assert [n for n, _ in got] == [n for n, _ in want]
    for (n, t), (_, w) in zip(got, want):
        assert ctypes.sizeof(t) == ctypes.sizeof(w), n


def test_struct_size_matches_compiled_header(tmp_path, fa_lib):
    import flash_attn_v100_cuda as op

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "fa_b200.h"\nint main(void){printf("%zu", sizeof(fa_b200_params_t));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    size = int(subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout)
    assert size == ctypes.sizeof(op.FaB200Params)


def test_library_exports_every_declared_symbol(fa_lib):
    import flash_attn_v100_cuda as op

    declared = re.findall(r"FA_B200_API\s+[\w\s\*]+?\b(fa_b200_\w+)\s*\(", open(HEADER).read())
    assert sorted(declared) == sorted(op.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(fa_lib, name) is not None
    assert fa_lib.fa_b200_abi_version() == 3
    assert fa_lib.fa_b200_launch_count() == 0


def test_argument_errors_use_the_c_error_convention(fa_lib):
    import flash_attn_v100_cuda as op

    p = op.FaB200Params()
    assert fa_lib.fa_b200_fwd(ctypes.byref(p), None) == -1
    assert b"size mismatch" in fa_lib.fa_b200_last_error()
    p.struct_bytes = ctypes.sizeof(p)
    p.dtype = 7
    assert fa_lib.fa_b200_varlen_fwd(ctypes.byref(p), None) == -1
    assert b"dtype" in fa_lib.fa_b200_last_error()
    p.dtype, p.batch, p.num_heads, p.num_heads_k, p.head_dim = 1, 1, 3, 2, 128
    assert fa_lib.fa_b200_kvcache_fwd(ctypes.byref(p), None) == -1
    assert b"divisible" in fa_lib.fa_b200_last_error()
    p.num_heads, p.head_dim = 4, 264
    assert fa_lib.fa_b200_fwd(ctypes.byref(p), None) == -1
    assert b"<= 256" in fa_lib.fa_b200_last_error()  # reference kernel/fused_mha_forward.cu:336
    p.head_dim = 100
    assert fa_lib.fa_b200_fwd(ctypes.byref(p), None) == -1
    assert b"multiple of 8" in fa_lib.fa_b200_last_error()  # reference :335
    assert fa_lib.fa_b200_fwd(None, None) == -1
    assert fa_lib.fa_b200_workspace_bytes(None, 2) == 0


def test_workspace_bytes_for_rotary_q(fa_lib):
    import flash_attn_v100_cuda as op

    p = op.FaB200Params()
    p.struct_bytes = ctypes.sizeof(p)
    p.batch, p.seqlen_q, p.num_heads, p.num_heads_k, p.head_dim, p.seqlen_k = 2, 3, 8, 2, 128, 1024
    assert fa_lib.fa_b200_workspace_bytes(ctypes.byref(p), op.KIND_DENSE) == 0
    p.rotary_dim = 64
    assert fa_lib.fa_b200_workspace_bytes(ctypes.byref(p), op.KIND_KVCACHE) >= 2 * 3 * 8 * 128 * 2


def test_operator_layer_has_the_reference_surface():
    import inspect

    import flash_attn_v100_cuda as op
    import flash_attn_v100 as api

    # reference kernel/fused_mha_api.cpp:19-23 (positional functions) and include/mha.h argument order
    assert list(inspect.signature(op.fwd).parameters) == [
        "q", "k", "v", "out_", "alibi_slopes_", "p_dropout", "softmax_scale", "is_causal", "window_left",
        "window_right", "softcap", "return_softmax", "gen_"]
    assert len(inspect.signature(op.varlen_fwd).parameters) == 22
    assert len(inspect.signature(op.fwd_kvcache).parameters) == 20
    # reference include/mha.h:67-87, 170-195
    assert list(inspect.signature(op.bwd).parameters) == [
        "dout", "q", "k", "v", "out", "softmax_lse", "dq_", "dk_", "dv_", "alibi_slopes_", "p_dropout", "softmax_scale",
        "is_causal", "window_left", "window_right", "softcap", "deterministic", "gen_", "rng_state_"]
    assert list(inspect.signature(op.varlen_bwd).parameters) == [
        "dout", "q", "k", "v", "out", "softmax_lse", "dq_", "dk_", "dv_", "cu_seqlens_q", "cu_seqlens_k", "alibi_slopes_",
        "max_seqlen_q", "max_seqlen_k", "p_dropout", "softmax_scale", "zero_tensors", "is_causal", "window_left",
        "window_right", "softcap", "deterministic", "gen_", "rng_state_"]
    assert issubclass(api.flash_attn_interface.FlashAttnFunc, __import__("torch").autograd.Function)
    assert issubclass(api.flash_attn_interface.FlashAttnVarlenFunc, __import__("torch").autograd.Function)
    # reference flash_attn_v100/flash_attn_interface.py:115-127, 272-289, 323-343
    sig = inspect.signature(api.flash_attn_func)
    assert list(sig.parameters) == ["q", "k", "v", "dropout_p", "softmax_scale", "causal", "window_size", "softcap",
                                    "alibi_slopes", "deterministic", "return_attn_probs"]
    assert sig.parameters["window_size"].default == (-1, -1) and sig.parameters["causal"].default is False
    sig = inspect.signature(api.flash_attn_varlen_func)
    assert list(sig.parameters)[3:7] == ["cu_seqlens_q", "cu_seqlens_k", "max_seqlen_q", "max_seqlen_k"]
    assert list(sig.parameters)[-1] == "block_table"
    sig = inspect.signature(api.flash_attn_with_kvcache)
    assert list(sig.parameters) == ["q", "k_cache", "v_cache", "k", "v", "rotary_cos", "rotary_sin", "cache_seqlens",
                                    "cache_batch_idx", "cache_leftpad", "block_table", "softmax_scale", "causal",
                                    "window_size", "softcap", "rotary_interleaved", "alibi_slopes", "num_splits",
                                    "return_softmax_lse"]
    assert sig.parameters["rotary_interleaved"].default is True
    assert api.flash_attn_gpu is api.flash_attn_func
    assert api.flash_attn_varlen_gpu is api.flash_attn_varlen_func
    assert api.flash_attn_with_kvcache_gpu is api.flash_attn_with_kvcache


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flash-attention-v100_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU or PyTorch fallback", ""), os.path.join(dirpath, f)


def test_operator_layer_fails_loudly_without_cuda():
    import torch

    import flash_attn_v100_cuda as op

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    q = torch.zeros(1, 2, 16, 64, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="must be on CUDA"):
        op.fwd(q, q, q, None, None, 0.0, 1.0, False, -1, -1, 0.0, False, None)


def test_torch_ops_are_registered_with_the_reference_schemas():
    import torch

    import flash_attn_v100_cuda as op

    assert op.register_torch_ops()
    for name in ("fwd", "bwd", "varlen_fwd", "varlen_bwd", "fwd_kvcache"):  # reference fused_mha_api.cpp:308-358
        assert hasattr(torch.ops.flash_attn_v100, name)
    schema = str(torch.ops.flash_attn_v100.fwd.default._schema)
    assert "Tensor(a!) q" in schema and "Generator? gen" in schema and schema.endswith("-> Tensor[]")
    assert len(torch.ops.flash_attn_v100.fwd_kvcache.default._schema.arguments) == 20
    q = torch.zeros(1, 1, 16, 64, dtype=torch.float16)
    with pytest.raises((NotImplementedError, RuntimeError)):  # no CPU kernel behind the op
        torch.ops.flash_attn_v100.fwd(q, q, q, None, None, 0.0, 1.0, False, -1, -1, 0.0, False, None)
