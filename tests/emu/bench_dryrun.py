"""TEST INFRASTRUCTURE (developer aid): run bench.py's own arm end to end — DeviceTrainer loop, the
Pipeline.learn e2e arm, the JSON line — on the CPU emulation of the kernels with a toy config, to catch
plumbing errors in a container without a GPU.  Timings are fake (every CUDA event reports 1 ms).

    python tests/emu/bench_dryrun.py RotatE scatter 0 independent 0
    python tests/emu/bench_dryrun.py RotatE by_entity 0 independent 0
    python tests/emu/bench_dryrun.py ComplEx scatter 0 reference 1      # pooled GEMM flow
    python tests/emu/bench_dryrun.py RotatE scatter 3 independent 0     # 3 virtual row shards
args: model, backward, virtual shards, pool, pooled-gemm flag
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from emu import torch_shim  # noqa: E402

torch_shim.install()
import torch  # noqa: E402

torch_shim._Event.elapsed_time = lambda self, other: 1.0
torch.cuda.set_device = lambda *a, **k: None


def _host(fn):
    def wrapped(*a, **k):
        if k.get("device") is not None and "cuda" in str(k["device"]):
            k["device"] = "cpu"
        return fn(*a, **k)
    return wrapped


for _name in ("tensor", "zeros", "empty", "ones", "full", "arange", "zeros_like", "empty_like", "randn", "as_tensor"):
    setattr(torch, _name, _host(getattr(torch, _name)))

import bench  # noqa: E402

model, backward, vshards, pool, gemm = (sys.argv[1:] + ["RotatE", "scatter", "0", "independent", "0"][len(sys.argv) - 1:])
bench.CONFIGS["tiny"] = ("Toy", model, 300, 5, 3000, 16, 8, 12, 9.0)
bench.cpu_reference_run = lambda *a, **k: {"value": 1.0, "unit": "triples/s", "cores": 1, "kind": "port",
                                           "sample": "stub", "ms_per_step": 1.0}


class Args:
    gpus, steps, warmup, impl, config, no_cpu_baseline, mode = 1, 4, 3, "ours", "tiny", False, None


Args.backward, Args.virtual_shards, Args.pool, Args.pooled_gemm = backward, int(vshards), pool, gemm == "1"
Args.packed_records = Args.merged_backward = False
Args.handshake = None
Args.no_hbm_config = os.environ.get("DRYRUN_HBM_CONFIG") != "1"  # the config-4 leg is too big for the emulation
bench.reference_run = lambda *a, **k: None
sys.exit(bench.run_ours(Args))
