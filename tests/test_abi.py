"""The C-ABI library loads without a GPU and exports every symbol include/brk_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "brk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(brk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    for must in ("brk_create", "brk_gather_rows", "brk_scatter_add_rows", "brk_bpr_fwd_bwd",
                 "brk_adam_dense_keras", "brk_adam_rows", "brk_adagrad_rows", "brk_philox_bpr_negatives"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from binrec_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_header():
    from binrec_b200 import _native
    assert sorted(_native.SIGNATURES) == _declared_symbols()
    assert _native.lib().brk_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    from binrec_b200 import _native
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.BrkError):
        _native.ctx()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "binary-recommendation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_torch_extension_loads_and_registers_the_ops():
    """The PyTorch-extension face of the boundary (csrc_torch/brk_torch.cpp -> libbrk_torch.so): loads without a GPU and
    registers every op under torch.ops.brk; compute calls need a GPU (tests/test_gpu_torchext.py)."""
    import torch
    from binrec_b200 import _torchext as T
    assert os.path.exists(T.LIB_PATH), "build it with __graft_entry__.build()"
    o = T.ops()
    assert o is not None and int(o.abi_version()) == 2
    for name in ("gather_rows", "scatter_add_rows", "philox_bpr_negatives", "bpr_fwd_bwd", "adam_dense_keras", "rows_to_bf16",
                 "score_topk", "topk_merge", "neumf_train_step", "use_ctx"):
        assert hasattr(o, name), name
    import pytest
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        o.gather_rows(torch.zeros(4, 4), torch.zeros(2, dtype=torch.int32))
