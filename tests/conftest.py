import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionfinish(session, exitstatus):
    """Leave the worst parity figures of this session where gpurun brings them back (tests/parity.py: MAXIMA)."""
    import json

    try:
        import parity
    except ImportError:
        return
    if parity.MAXIMA:
        out = REPO / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / "parity_maxima.json").write_text(json.dumps(parity.MAXIMA, indent=1, sort_keys=True))


def _extra_models():
    """Seeded synthetic xxxlarge unless a real nn-xxxlarge.npz has been dropped into the weights dir."""
    from neuralfoil_b200 import synthetic

    if (REPO / "neuralfoil_b200" / "nn_weights_and_biases" / "nn-xxxlarge.npz").exists():
        return {}
    return {"xxxlarge": synthetic.synthetic_xxxlarge_params()}


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle with the seeded synthetic xxxlarge injected (test infrastructure only)."""
    from oracle.neuralfoil_oracle import Oracle

    return Oracle(extra_models=_extra_models())


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


@pytest.fixture(scope="session")
def library():
    """The built C-ABI library (built on demand here; on the GPU box the prebuilt .so is used)."""
    from neuralfoil_b200 import _build, _lib

    if not _lib.LIB_PATH.exists():
        _build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def evaluator(library):
    """The CUDA evaluator on cuda:0 (GPU tests only).  No fallback: raises if no B200 is present."""
    from neuralfoil_b200 import Evaluator

    ev = Evaluator(device=0, extra_models=_extra_models())
    yield ev
    ev.close()
