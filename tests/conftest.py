import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files(sub=""):
    """Fixtures recorded from the unmodified reference.  ``sub="next"``: features added after the last hardware
    session (reset_agent_fixed_duration, MTV distance) — the oracle is pinned on them here on the CPU; their GPU
    parity tests live in ``test_gpu_zz_next.py`` so that they run after the hardware-verified suite."""
    import glob
    return sorted(glob.glob(os.path.join(REPO, "tests", "golden", sub, "*.npz")))


def all_golden_files():
    return golden_files() + golden_files("next")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O
