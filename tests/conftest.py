import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle with the seeded synthetic xxxlarge injected (test infrastructure only)."""
    from oracle.neuralfoil_oracle import Oracle
    from neuralfoil_b200 import synthetic

    extra = {}
    if not (REPO / "neuralfoil_b200" / "nn_weights_and_biases" / "nn-xxxlarge.npz").exists():
        extra["xxxlarge"] = synthetic.synthetic_xxxlarge_params()
    return Oracle(extra_models=extra)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
