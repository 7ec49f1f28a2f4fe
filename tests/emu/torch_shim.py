"""TEST INFRASTRUCTURE (developer aid, opt-in): run the `-m gpu` test files in a container WITHOUT a GPU.

    KGE_TEST_EMU=1 python -m pytest tests -m gpu -k "not full_size"

swaps libkge_b200.so for the emulation library (tests/emu: the same kernel sources compiled for the
host) underneath the UNCHANGED product package, and teaches torch to pretend CPU tensors are device
tensors, so the whole Python layer — ctypes argument marshalling, autograd functions, models, losses,
sampler, Evaluation, Pipeline, DeviceTrainer — is exercised against the oracle / golden vectors before
a B200 is available.  Nothing here is imported unless KGE_TEST_EMU=1; the driver's `-m gpu` run on the
real box never sees it, and the product keeps refusing to run without CUDA.
"""
import contextlib
import ctypes as C

import torch


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a, **k):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 0.0


class _Stream:
    cuda_stream = 0


def install():
    from mkb_b200 import _native

    from . import build_emu

    import os

    lib = C.CDLL(os.environ.get("KGE_EMU_LIB") or build_emu.build())  # KGE_EMU_LIB: e.g. the sanitizer build
    for name, (res, args) in _native.PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _native._lib = lib
    _native.require_cuda = lambda *tensors: None
    _native.stream_ptr = lambda device=None: None
    torch.cuda.is_available = lambda: True
    torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Event = _Event
    torch.cuda.device_count = lambda: 1
    torch.cuda.is_current_stream_capturing = lambda: False
    torch.Tensor.is_cuda = property(lambda self: True)
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    _orig_to = torch.Tensor.to

    def _to(self, *args, **kwargs):  # .to("cuda") / .to(device="cuda:0") stay on the host
        args = tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) or
                     (isinstance(a, torch.device) and a.type == "cuda") else a for a in args)
        d = kwargs.get("device")
        if (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda"):
            kwargs["device"] = "cpu"
        return _orig_to(self, *args, **kwargs)

    torch.Tensor.to = _to
    _orig_module_to = torch.nn.Module.to

    def _module_to(self, *args, **kwargs):
        args = tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) or
                     (isinstance(a, torch.device) and a.type == "cuda") else a for a in args)
        return _orig_module_to(self, *args, **kwargs)

    torch.nn.Module.to = _module_to
