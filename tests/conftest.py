"""pytest configuration: the `gpu` marker, import paths, shared fixtures.

`-m "not gpu"` (run without a GPU): oracle vs the reference's invariants and golden fixtures,
host logic, C-ABI loading.  `-m gpu` (run on a B200): parity of the CUDA path, through the
C ABI, against the oracle.  Only tests may import oracle/ (it is the checker, never the path).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "direct-nbody_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    O.set_step_cap(20_000_000)  # the reference's step loops can spin forever on bad (eta, state)
    return O


@pytest.fixture(scope="session")
def ctx():
    import nbh
    import nbk
    c = nbk.Context(0)  # raises loudly without libnbk.so or without a B200
    nbh.load_library().nbh_set_step_cap(20_000_000)
    yield c
    c.close()
