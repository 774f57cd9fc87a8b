"""The PyTorch-extension face of the boundary: `torch.ops.brk.*` custom ops (csrc_torch/brk_torch.cpp, built in-tree as
libbrk_torch.so) over the same C ABI the ctypes binding (_native.py) reaches.  hotpath.py prefers these ops for the
per-call entry points; BRK_BINDING=ctypes forces the ctypes route (the two are tested against each other in
tests/test_gpu_torchext.py).  The ctypes binding remains the GPU-less symbol check and serves the entry points that take
structures of many tables or host arrays."""
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbrk_torch.so")
_state = {"tried": False, "ops": None}


def ops():
    """torch.ops.brk when the extension is built and not disabled, else None."""
    if not _state["tried"]:
        _state["tried"] = True
        if os.environ.get("BRK_BINDING", "torch") != "ctypes" and os.path.exists(LIB_PATH):
            torch.ops.load_library(LIB_PATH)
            from . import _native as N
            if int(torch.ops.brk.abi_version()) != int(N.lib().brk_abi_version()):
                raise N.BrkError("libbrk_torch.so and libbrk_b200.so disagree on the ABI version: rebuild both")
            _state["ops"] = torch.ops.brk
    return _state["ops"]


def share_ctx(device_index, handle):
    o = ops()
    if o is not None:
        o.use_ctx(int(device_index), int(handle))
