import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MNIST_FIXTURE = os.path.join(GOLDEN, "mnist_aux.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture
def oracle_backend():
    """Run the product's host logic against the float64 CPU oracle backend (CPU tests only)."""
    from oracle_backend import OracleBackend
    from svgp_vae_b200 import backend
    old = backend.set_backend_for_tests(OracleBackend())
    yield
    backend.set_backend_for_tests(old)


@pytest.fixture(scope="session")
def cuda_backend():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from svgp_vae_b200 import backend
    backend.set_backend_for_tests(None)
    return backend.get_backend()          # raises (fails the test) if libsvgp_b200.so is missing


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().clamp_min(1e-300)
    return ((a - b).abs().max() / den).item()
