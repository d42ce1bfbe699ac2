"""The C-ABI library loads on a CPU-only machine and exports every symbol include/natrium_b200.h declares.
No compute call is made here; creating a context without a GPU must fail loudly (no CPU fallback)."""
import os
import re

import pytest

from natrium_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "natrium_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = _capi.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in natrium_b200.h but not exported"
    assert set(names) == set(_capi.SIGNATURES), set(names) ^ set(_capi.SIGNATURES)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.NatriumB200Error) as ei:
        _capi.Context(0)
    assert ei.value.code == _capi.NB200_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under natrium_b200/ may reference it."""
    pkg = os.path.join(ROOT, "natrium_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "liboracle" not in txt, fn
