"""CPU tests of the drop-in boundary: libpskmer.so builds, loads, and exports exactly the
symbols include/pskmer.h declares; without a GPU the product path fails loudly (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib_path():
    from phenotypeseeker_b200 import build
    return build.build()


def _header_symbols():
    with open(os.path.join(ROOT, "include", "pskmer.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_path):
    declared = _header_symbols()
    assert len(declared) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ps_[a-z0-9_]+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert exported <= set(declared), sorted(exported - set(declared))   # no undeclared entry points


def test_ctypes_binding_covers_the_header(lib_path):
    from phenotypeseeker_b200 import _native
    assert sorted(_native.SIGNATURES) == _header_symbols()
    L = _native.load()
    assert L.ps_version() >= 100


def test_sm100a_sass_present(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_fails_loudly(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phenotypeseeker_b200._native import Context, PsError
    with pytest.raises(PsError):
        Context(0)


def test_null_and_error_paths(lib_path):
    L = ctypes.CDLL(lib_path)
    assert L.ps_begin(None, 16, 2, 1) == -1
    assert L.ps_ctx_create(0, None) == -1
    L.ps_last_error.restype = ctypes.c_char_p
    assert L.ps_last_error(None) is not None


def test_product_code_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "phenotypeseeker_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dp, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, flags=re.M), fn
                assert "kmer_oracle" not in src, fn
