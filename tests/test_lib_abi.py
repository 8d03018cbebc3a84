"""The C-ABI library loads on a CPU-only box and exports exactly what include/svgp_b200.h declares."""
import os
import re

import pytest

from conftest import ROOT
from svgp_vae_b200 import _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "svgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(svgp_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from svgp_vae_b200.build import build_library
        build_library()
    lib = _lib.load()
    declared = _declared()
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.svgp_version() >= 100
    assert lib.svgp_last_error() is not None          # host-only call, no device needed


def _prototypes():
    """name -> list of C parameter types of every prototype in the header (comments stripped)."""
    src = open(os.path.join(ROOT, "include", "svgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(svgp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        params = params.strip()
        plist = [] if params in ("", "void") else [re.sub(r"\s+", " ", q.strip()) for q in params.split(",")]
        out[name] = [re.sub(r"\s*\w+$", "", q) if not q.endswith("*") else q for q in plist]      # drop the parameter name
    return out


def test_ctypes_signatures_match_the_header_prototypes():
    """Every entry of _lib.SIGNATURES has the arity of its prototype and a pointer / 64-bit / int / double in the same places:
    a ctypes call with a stale signature would pass garbage into a kernel launch instead of failing."""
    import ctypes
    kinds = {ctypes.c_void_p: "ptr", ctypes.c_int64: "i64", ctypes.c_int: "int", ctypes.c_double: "f64", ctypes.c_float: "f32"}

    def kind_of_c(t):
        if "*" in t:
            return "ptr"
        t = t.replace("const ", "").strip()
        return {"int64_t": "i64", "int": "int", "double": "f64", "float": "f32"}[t]

    protos = _prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, args in _lib.SIGNATURES.items():
        want = [kind_of_c(t) for t in protos[name]]
        got = ["ptr" if (a not in kinds) else kinds[a] for a in args]           # POINTER(KopStruct) etc. count as pointers
        assert got == want, (name, got, want)


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from svgp_vae_b200 import backend
    old = backend.set_backend_for_tests(None)
    try:
        with pytest.raises(_lib.SvgpLibraryError):
            backend.get_backend()
    finally:
        backend.set_backend_for_tests(old)


def test_no_oracle_import_in_product_sources():
    pkg = os.path.join(ROOT, "svgp_vae_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
