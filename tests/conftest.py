import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Developer aid (tests/emu/torch_shim.py): KGE_TEST_EMU=1 runs the `-m gpu` files on the CPU emulation of the
# kernels in a container without a GPU.  Never set by the driver; the real GPU run uses DEV = "cuda".
EMU = os.environ.get("KGE_TEST_EMU") == "1"
DEV = "cpu" if EMU else "cuda"
if EMU:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu import torch_shim

    torch_shim.install()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def step_cases():
    return load_golden("step_cases.npz")


@pytest.fixture(scope="session")
def doctest_pins():
    return load_golden("doctest_pins.npz")


@pytest.fixture(scope="session")
def sampler_cases():
    return load_golden("sampler_cases.npz")


@pytest.fixture(scope="session")
def eval_cases():
    return load_golden("eval_cases.npz")


@pytest.fixture(scope="session")
def eval_doctest():
    return load_golden("eval_doctest.npz")


MODELS = ("TransE", "DistMult", "ComplEx", "RotatE")
MODES = ("tail-batch", "head-batch")


def score_tol(ref, rel=1e-4):
    """|Δ| <= rel * max(|ref|, mean|ref|): the '1e-4 relative' bar of BASELINE.json's north_star with
    the scale SURVEY §7 calls for (ComplEx/DistMult/TransE scores cross zero)."""
    ref = np.asarray(ref, dtype=np.float64)
    return rel * np.maximum(np.abs(ref), np.abs(ref).mean() + 1e-30)
