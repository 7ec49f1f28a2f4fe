"""TEST INFRASTRUCTURE (developer aid): __graft_entry__.smoke() on the CPU emulation of the kernels, to catch
plumbing errors in a container without a GPU.  On the GPU box the driver calls smoke() itself, on cuda:0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from emu import torch_shim  # noqa: E402

torch_shim.install()

import __graft_entry__  # noqa: E402

__graft_entry__.smoke()
