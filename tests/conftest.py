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


# Order of the `-m gpu` run: the files whose kernels have B200 measurements behind them first, the rows
# written after the round's GPU budget was spent next, the tcgen05 pooled-GEMM flow (spin-waits on mbarriers)
# last — so `-x` reports as much as possible before the least-exercised code runs.
_GPU_FILE_ORDER = ("test_gpu_parity", "test_gpu_advice", "test_gpu_fullshape", "test_gpu_cfg5", "test_gpu_cfg1_epoch", "test_gpu_multi",
                   "test_gpu_rows_next", "test_gpu_sharded", "test_gpu_train_byent")


def pytest_collection_modifyitems(config, items):
    def key(item):
        mod = os.path.splitext(os.path.basename(str(getattr(item, "path", None) or item.fspath)))[0]
        rank = _GPU_FILE_ORDER.index(mod) if mod in _GPU_FILE_ORDER else -1  # CPU files keep their place, first
        return (rank, "pooled_gemm" in item.name)

    items.sort(key=key)  # stable: collection order inside each group is kept
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout") and not EMU:
            # a kernel that never returns must not hold the GPU box until the driver's limit: pytest-timeout's
            # thread method ends the process (and with it the CUDA context) from a watchdog thread
            item.add_marker(pytest.mark.timeout(600, method="thread"))


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
