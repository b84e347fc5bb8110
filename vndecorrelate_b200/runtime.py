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


class _PinnedBlock:
    """One page-locked block of the output pool; goes back to the pool when the last array on it dies."""

    __slots__ = ("ptr", "nbytes", "pid", "__weakref__")

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes, self.pid = ptr, nbytes, os.getpid()

    def __del__(self, _getpid=os.getpid):  # bound at definition: module globals may already be gone at interpreter shutdown
        try:
            if self.pid == _getpid():  # a forked child must not recycle pointers registered in its parent's CUDA context
                _pinned_pool_release(self.ptr, self.nbytes)
        except Exception:  # interpreter shutdown: the pool and the library are being torn down anyway
            pass


_POOL_LOCK = threading.RLock()  # re-entrant: a garbage collection inside the locked region may release another block
_POOL_FREE: dict[int, list[int]] = {}   # block size -> free pointers
_POOL_BYTES = [0]                       # page-locked bytes this pool has allocated (free and in use)
_POOL_CAP = [int(os.environ.get("VND_PINNED_POOL_MB", "512")) << 20]
_POOL_WARNED = [False]


def set_pinned_pool_cap(nbytes: int) -> None:
    """Upper bound of the page-locked memory the result pool may hold (default 512 MB, ``VND_PINNED_POOL_MB``).
    Results that do not fit come back on ordinary memory and are downloaded through the context's staging ring;
    callers that stream GB-sized slabs raise the cap so that results are written by DMA at link speed."""
    with _POOL_LOCK:
        _POOL_CAP[0] = int(nbytes)
    trim_pinned_pool(int(nbytes))


def trim_pinned_pool(keep_bytes: int = 0) -> int:
    """Give free blocks back to the OS (``cudaFreeHost``) until the pool holds at most ``keep_bytes``; returns the
    bytes released.  Called automatically when a request does not fit under the cap."""
    released = 0
    with _POOL_LOCK:
        for size in sorted(_POOL_FREE, reverse=True):
            free = _POOL_FREE[size]
            while free and _POOL_BYTES[0] > keep_bytes:
                ptr = free.pop()
                try:
                    N.lib().vnd_host_free(C.c_void_p(ptr))
                except Exception:  # interpreter shutdown
                    pass
                _POOL_BYTES[0] -= size
                released += size
    return released


def _pinned_pool_release(ptr: int, nbytes: int) -> None:
    try:
        with _POOL_LOCK:
            _POOL_FREE.setdefault(nbytes, []).append(ptr)
    except Exception:  # interpreter shutdown
        pass


def _pinned_pool_after_fork() -> None:
    # the parent's blocks belong to the parent's CUDA context: forget them (nothing is freed in the child)
    _POOL_FREE.clear()
    _POOL_BYTES[0] = 0


if hasattr(os, "register_at_fork"):
    os.register_at_fork(after_in_child=_pinned_pool_after_fork)


def pinned_empty(shape, dtype) -> np.ndarray:
    """``np.empty(shape, dtype)`` on page-locked memory from a recycling pool, so that the device-to-host copy of
    a result is one DMA at link speed instead of a staged copy into freshly mapped pages (which costs more
    than the kernels for the stereo files of BASELINE configs 1 and 2).  The array owns its block: it returns
    to the pool when the array and its views are gone.  Blocks are power-of-two sized; when a request does not
    fit under the cap (``set_pinned_pool_cap``), idle blocks of other sizes are released first, and if it still
    does not fit the result is ordinary ``np.empty`` memory (a ``ResourceWarning`` says so once): correct, but
    downloaded through the staging ring."""
    dt = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = count * dt.itemsize
    if nbytes == 0:
        return np.empty(shape, dtype=dt)
    size = 1 << max(16, (nbytes - 1).bit_length())  # power-of-two blocks of at least 64 KB
    if size > 1 << 30:  # above 1 GB round to 256 MB instead of doubling
        size = -(-nbytes // (256 << 20)) * (256 << 20)
    ptr = None
    with _POOL_LOCK:
        free = _POOL_FREE.get(size)
        if free:
            ptr = free.pop()
        else:
            if _POOL_BYTES[0] + size > _POOL_CAP[0]:
                trim_pinned_pool(max(0, _POOL_CAP[0] - size))
            if _POOL_BYTES[0] + size <= _POOL_CAP[0]:
                _POOL_BYTES[0] += size
            else:
                if not _POOL_WARNED[0]:
                    _POOL_WARNED[0] = True
                    import warnings

                    warnings.warn(f"vndecorrelate_b200: the page-locked result pool is at its cap ({_POOL_CAP[0] >> 20} MB); a {nbytes >> 20} MB "
                                  "result uses pageable memory (slower download). Raise it with runtime.set_pinned_pool_cap().", ResourceWarning)
                return np.empty(shape, dtype=dt)
    if ptr is None:
        p = C.c_void_p()
        try:
            rc = N.lib().vnd_ctx_host_alloc(HostContext.get().handle, size, C.byref(p))  # on the context's device, never on GPU 0 by accident
        except N.VndError:
            rc = -1
        if rc != 0 or not p.value:
            with _POOL_LOCK:
                _POOL_BYTES[0] -= size
            return np.empty(shape, dtype=dt)
        ptr = p.value
    block = _PinnedBlock(ptr, size)
    buf = (C.c_char * size).from_address(ptr)
    buf._vnd_block = block  # the ctypes view keeps the block alive, numpy keeps the ctypes view alive
    return np.frombuffer(buf, dtype=dt, count=count).reshape(shape)


class PinnedArray:
    """A numpy view of page-locked host memory from ``vnd_host_alloc``."""

    def __init__(self, shape, dtype=np.float32):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        N.check(N.lib().vnd_ctx_host_alloc(HostContext.get().handle, self.nbytes, C.byref(ptr)), "vnd_ctx_host_alloc")
        self._ptr = ptr
        buf = (C.c_char * max(self.nbytes, 1)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(self, N.lib().vnd_host_free, ptr)
