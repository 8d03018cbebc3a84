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
