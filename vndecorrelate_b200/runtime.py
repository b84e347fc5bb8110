"""Device plumbing between the Python API and the C ABI.

Two kinds of buffers reach the kernels:

* **numpy arrays** go through the ``_host`` entry points of a per-device ``vnd_ctx`` (device arena,
  streams, uploads and downloads inside the call) — this is the drop-in path for the reference's
  numpy-in / numpy-out API and needs nothing but numpy and the shared library;
* **CUDA ``torch.Tensor``s** go through the ``_dev`` entry points on torch's current stream with
  torch-allocated outputs and workspaces — no host round trip, usable inside a larger torch job.

PyTorch is plumbing here (device memory, streams, ``torch.distributed``); all arithmetic is in
``libvnd_b200.so``.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
import weakref

import numpy as np

from . import _native as N

_DT = {np.dtype(np.float32): N.VND_F32, np.dtype(np.float64): N.VND_F64}


def is_torch_tensor(x) -> bool:
    mod = type(x).__module__
    return mod == "torch" or mod.startswith("torch.")


def default_device() -> int:
    if "VND_DEVICE" in os.environ:
        return int(os.environ["VND_DEVICE"])
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    return 0


class HostContext:
    """Owner of one ``vnd_ctx`` (one per device and process)."""

    _instances: dict[int, "HostContext"] = {}
    _lock = threading.Lock()

    def __init__(self, device: int):
        self.device = device
        handle = C.c_void_p()
        N.check(N.lib().vnd_ctx_create(device, C.byref(handle)), "vnd_ctx_create")
        self.handle = handle
        weakref.finalize(self, N.lib().vnd_ctx_destroy, handle)

    @classmethod
    def get(cls, device: int | None = None) -> "HostContext":
        dev = default_device() if device is None else device
        with cls._lock:
            if dev not in cls._instances:
                cls._instances[dev] = HostContext(dev)
            return cls._instances[dev]


# ---------------------------------------------------------------------------------------------
# signal descriptors
# ---------------------------------------------------------------------------------------------


def host_signal(a: np.ndarray, *, mono_as_stereo: bool = False) -> N.SignalStruct:
    """vnd_signal for a numpy array: 2-D ``(frames, channels)`` with any of the dense layouts, or
    1-D (a mono signal; ``mono_as_stereo`` presents it as two channels with stride_c = 0)."""
    if a.dtype not in _DT:
        raise TypeError(f"unsupported dtype {a.dtype}")
    isz = a.dtype.itemsize
    if a.ndim == 1:
        st = a.strides[0] // isz if a.shape[0] > 1 else 1
        return N.SignalStruct(a.ctypes.data, a.shape[0], 2 if mono_as_stereo else 1, _DT[a.dtype], st, 0)
    if a.ndim != 2:
        raise ValueError(f"expected a 1-D or 2-D signal, got shape {a.shape}")
    frames, ch = a.shape
    st = a.strides[0] // isz if frames > 1 else ch
    sc = a.strides[1] // isz if ch > 1 else 1
    return N.SignalStruct(a.ctypes.data, frames, ch, _DT[a.dtype], st, sc)


def dense(a: np.ndarray) -> np.ndarray:
    """``a`` itself if it is one contiguous block in C order or in planar (transposed-contiguous)
    order, else a C-order copy."""
    if a.flags.c_contiguous or (a.ndim == 2 and a.T.flags.c_contiguous):
        return a
    return np.ascontiguousarray(a)


def torch_signal(t, *, mono_as_stereo: bool = False) -> N.SignalStruct:
    import torch

    dt = {torch.float32: N.VND_F32, torch.float64: N.VND_F64}.get(t.dtype)
    if dt is None:
        raise TypeError(f"unsupported dtype {t.dtype}")
    if t.dim() == 1:
        return N.SignalStruct(t.data_ptr(), t.shape[0], 2 if mono_as_stereo else 1, dt, t.stride(0) if t.shape[0] > 1 else 1, 0)
    if t.dim() != 2:
        raise ValueError(f"expected a 1-D or 2-D signal, got shape {tuple(t.shape)}")
    frames, ch = t.shape
    return N.SignalStruct(t.data_ptr(), frames, ch, dt, t.stride(0) if frames > 1 else ch, t.stride(1) if ch > 1 else 1)


def torch_stream_ptr(t) -> int:
    import torch

    return torch.cuda.current_stream(t.device).cuda_stream


def device_program(prog, device):
    """Upload (once per device) a TapProgram for the ``_dev`` entry points; returns the struct and
    keeps the tensors alive on the program object."""
    import torch

    key = str(device)
    if key not in prog._device:
        words = torch.from_numpy(prog.words if prog.words.size else np.zeros(1, np.int32)).to(device)
        offsets = torch.from_numpy(prog.offsets).to(device)
        prog._device[key] = (words, offsets)
    words, offsets = prog._device[key]
    return prog.struct(words.data_ptr(), offsets.data_ptr())


# ---------------------------------------------------------------------------------------------
# pinned host memory (zero-staging transfers for the streaming path)
# ---------------------------------------------------------------------------------------------


class PinnedArray:
    """A numpy view of page-locked host memory from ``vnd_host_alloc``."""

    def __init__(self, shape, dtype=np.float32):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        N.check(N.lib().vnd_host_alloc(self.nbytes, C.byref(ptr)), "vnd_host_alloc")
        self._ptr = ptr
        buf = (C.c_char * max(self.nbytes, 1)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(self, N.lib().vnd_host_free, ptr)
