import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement of the reference (test infrastructure); built on demand."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def refceed():
    """The unmodified reference library, when oracle/_ref was built (dev container and, shipped, on the GPU box)."""
    from oracle import refceed as R
    if not R.available():
        pytest.skip("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
    return R


def bp_case_key(bp, p, nel, gallery, interlaced):
    return f"bp{bp}_p{p}_n{nel[0]}x{nel[1]}x{nel[2]}_g{int(gallery)}_i{int(interlaced)}"


BP_CASES = [
    (1, 1, (3, 2, 2), False, False), (1, 3, (2, 2, 2), False, False), (3, 1, (2, 3, 2), False, False), (3, 2, (2, 2, 2), False, False),
    (3, 3, (2, 2, 1), False, False), (3, 4, (2, 1, 1), False, False), (3, 6, (1, 1, 2), False, False), (3, 8, (1, 1, 1), False, False),
    (5, 4, (2, 1, 1), False, False), (5, 7, (1, 1, 1), False, False), (2, 2, (2, 2, 1), False, False), (4, 1, (2, 2, 2), False, False),
    (4, 2, (2, 1, 2), False, True), (6, 2, (2, 2, 1), False, False), (6, 3, (1, 2, 1), False, True), (1, 2, (2, 2, 2), True, False),
    (3, 2, (2, 2, 2), True, False), (3, 5, (2, 1, 1), False, False), (3, 7, (1, 1, 1), False, False), (5, 5, (1, 2, 1), False, False),
    (5, 6, (1, 1, 2), False, False), (6, 4, (2, 1, 1), False, False), (6, 6, (1, 1, 1), False, False), (4, 3, (2, 1, 1), False, False),
    (1, 2, (2, 2, 2), False, False), (1, 4, (2, 1, 1), False, False), (1, 5, (1, 1, 2), False, False), (2, 3, (2, 1, 1), False, False),
]
