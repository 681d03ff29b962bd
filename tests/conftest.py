"""pytest configuration: the `gpu` marker (tests that need a B200) and shared fixtures.

`python -m pytest tests -m "not gpu"` covers the CPU oracle (known-answer tests, the independent
brute-force checker, golden fixtures), the host logic and the C-ABI symbol table; it never launches
a kernel.  `python -m pytest tests -m gpu` holds the parity tests proper: they call the CUDA path
through the C ABI and compare it with the oracle."""
from __future__ import annotations

import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    import oracle_binding

    oracle_binding.build_oracle()
    yield


@pytest.fixture(scope="session")
def gpu_ctx_factory():
    """Factory of C-ABI contexts on cuda:0.  Fails loudly (no skip, no fallback) when the CUDA library or
    the device is missing: a GPU test that cannot reach the kernels must not pass."""
    from curvedspacesim_b200 import binding

    made = []

    def make():
        ctx = binding.Context(0)
        made.append(ctx)
        return ctx

    yield make
    for c in made:
        c.close()
