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


def test_header_is_plain_c_and_links(tmp_path):
    """include/natrium_b200.h is the FFI surface: it must compile as C99 (no C++ in the signatures) and a C program that
    takes the address of every declared entry point must link against the library."""
    import subprocess
    names = declared_symbols()
    src = tmp_path / "abi_check.c"
    body = "\n".join(f"    p[{i}] = (fn)&{n};" for i, n in enumerate(names))
    src.write_text('#include "natrium_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\nint main(void) {\n'
                   f"    fn p[{len(names)}];\n{body}\n"
                   f"    nb200_collision_params cp; cp.force[2] = 0.0; (void)cp;\n"
                   f'    printf("%d\\n", (int)(sizeof(p) / sizeof(p[0])));\n    return p[0] == 0;\n}}\n')
    exe = tmp_path / "abi_check"
    libdir = os.path.join(ROOT, "natrium_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lnatrium_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert int(out) == len(names)
