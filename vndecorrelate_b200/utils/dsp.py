"""Stereo helpers of the hot path, with the reference's names and in-place contracts
(``src/vndecorrelate/utils/dsp.py``), executed by the CUDA library.

What ``VelvetNoise.decorrelate`` / ``HaasEffect`` / the optimiser touch is here: the M/S helpers, width,
side-channel encode, the normalisers (``rms_normalize`` in every mode, ``peak_normalize``), ``polar_coordinates``,
the dtype/shape helpers and the tap-position maths.  The reference's remaining analysis and plotting helpers
(``cross_correlogram``, ``sine_sweep``, envelopes) are outside the hot path and are not provided.

All in-place functions take numpy arrays (float32 or float64, shape ``(n, 2)``) or CUDA
``torch.Tensor``s and return ``None`` like the reference's.
"""

from __future__ import annotations

import ctypes as C
from enum import StrEnum

import numpy as np

from .. import _native as N
from .. import runtime as R
from ..taps import IDENTITY_ENVELOPE, log_distribution

EPSILON: float = 1e-10  # utils/dsp.py:6

__all__ = [
    "EPSILON", "IDENTITY_ENVELOPE", "NormalizeMode", "LayoutMode", "apply_stereo_width",
    "encode_signal_to_side_channel", "to_float32", "rms_normalize", "mono_to_stereo", "stereo_to_mono",
    "LR_to_MS", "MS_to_LR", "generate_log_distribution", "apply_log_distribution", "uniform_density", "peak_normalize", "polar_coordinates",
    "check_mono", "check_stereo", "check_equal_length",
]


class NormalizeMode(StrEnum):
    STEREO = "stereo"
    DUAL_MONO = "dual_mono"


class LayoutMode(StrEnum):
    LR = "LR"  # Left-Right
    MS = "MS"  # Mid-Side


# ---- shape checks (utils/dsp.py:289-310) -------------------------------------------------------


def check_mono(input_signal) -> None:
    if input_signal.ndim != 1:
        raise ValueError(f"Input shape invalid: Expected shape (num samples,), but got shape {tuple(input_signal.shape)}.")


def check_stereo(input_signal) -> None:
    if input_signal.ndim != 2 or input_signal.shape[1] != 2:
        raise ValueError(f"Input shape invalid: Expected shape (num samples, 2), but got shape {tuple(input_signal.shape)}.")


def check_equal_length(x, y, dim: int = 0) -> None:
    if x.shape[dim] != y.shape[dim]:
        raise ValueError(
            f"Input length mismatch: Expected signals of equal length, but got lengths {x.shape[dim]} and {y.shape[dim]} for dimension {dim}."
        )


# ---- dtype / layout helpers (data movement only) -----------------------------------------------


def to_float32(input_signal):
    """``astype(float32, copy=False)``: integers are NOT rescaled (utils/dsp.py:66-68)."""
    if R.is_torch_tensor(input_signal):
        import torch

        return input_signal.to(torch.float32)
    return input_signal.astype(np.float32, copy=False)


def mono_to_stereo(input_signal):
    check_mono(input_signal)
    if R.is_torch_tensor(input_signal):
        import torch

        return torch.stack((input_signal, input_signal), dim=1)
    return np.column_stack((input_signal, input_signal))


def stereo_to_mono(input_signal):
    check_stereo(input_signal)
    return (input_signal[:, 0] + input_signal[:, 1]) * 0.5


# ---- in-place stereo operations on the device --------------------------------------------------

_OP_LR_TO_MS, _OP_MS_TO_LR, _OP_WIDTH, _OP_ENCODE, _OP_RMS = range(5)


def _float_array(a, name: str):
    if a.dtype not in (np.float32, np.float64):
        raise TypeError(f"{name} must be float32 or float64 for the in-place helpers, got {a.dtype}")
    return a


def _stereo_op(a, op: int, width: float = 0.0, dry=None) -> None:
    if R.is_torch_tensor(a):
        import torch

        sa = R.torch_signal(a)
        sd = None
        if dry is not None:
            dry = dry.to(a.dtype)
            sd = R.torch_signal(dry)
        nbytes = 128 + 8 * (a.shape[0] // 64 + 8)  # op 4 on planar tensors keeps the leaf sums of numpy's pairwise order here
        work = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
        N.check(N.lib().vnd_stereo_op_dev(C.byref(sa), C.byref(sd) if sd is not None else None, op, float(width),
                                          work.data_ptr(), nbytes, R.torch_stream_ptr(a)), "vnd_stereo_op_dev")
        return
    _float_array(a, "signal")
    # C-order and planar (Fortran-ordered) arrays go to the device as they are: rms_normalize sums a signal in the order
    # numpy uses for its layout (sequential along frames for C order, pairwise per column for planar)
    work = a if (R.dense(a) is a and a.flags.writeable) else np.ascontiguousarray(a)
    sa = R.host_signal(work)
    sd = None
    if dry is not None:
        dry = dry if dry.dtype == a.dtype else dry.astype(a.dtype)
        dry = R.dense(dry)
        sd = R.host_signal(dry)
    ctx = R.HostContext.get()
    N.check(N.lib().vnd_stereo_op_host(ctx.handle, C.byref(sa), C.byref(sd) if sd is not None else None, op, float(width)), "vnd_stereo_op_host")
    if work is not a:
        a[...] = work


def LR_to_MS(input_signal) -> None:
    """M = (L + R) / 2, S = (L - R) / 2 in place (utils/dsp.py:124-144)."""
    check_stereo(input_signal)
    _stereo_op(input_signal, _OP_LR_TO_MS)


def MS_to_LR(input_signal) -> None:
    """L = M + S, R = M - S in place (utils/dsp.py:147-167)."""
    check_stereo(input_signal)
    _stereo_op(input_signal, _OP_MS_TO_LR)


def apply_stereo_width(input_signal, width: float) -> None:
    """Scale mid by ``1 - width`` and side by ``width`` in place (utils/dsp.py:21-37)."""
    check_stereo(input_signal)
    _stereo_op(input_signal, _OP_WIDTH, width=width)


def encode_signal_to_side_channel(input_signal, decorrelated_signal) -> None:
    """Overwrite ``decorrelated_signal`` with mid = L+R of ``input_signal`` and side = its own
    (L - R) / 2, decoded back to L/R (utils/dsp.py:40-63)."""
    check_stereo(input_signal)
    check_stereo(decorrelated_signal)
    check_equal_length(input_signal, decorrelated_signal)
    _stereo_op(decorrelated_signal, _OP_ENCODE, dry=input_signal)


def _flat_inplace(fn_host, fn_dev, arrays, flat: bool):
    """Run an in-place device helper on contiguous views of ``arrays`` (numpy: upload / run / download inside the host
    call; CUDA tensors: on torch's stream), copying back when a contiguous temporary had to be made."""
    first = arrays[-1]
    # numpy reduces with axis=None in MEMORY order: a Fortran-ordered (planar) 2-D array is one contiguous run of its
    # transpose, so hand that view over (same memory, same in-place result) instead of a C-order copy
    if R.is_torch_tensor(first):
        import torch

        if flat:
            arrays = [a.t() if (a.dim() == 2 and not a.is_contiguous() and a.t().is_contiguous()) else a for a in arrays]
        first = arrays[-1]
        work = [a if a.is_contiguous() else a.contiguous() for a in arrays]
        nbytes = C.c_size_t()
        N.check(N.lib().vnd_dsp_workspace(max(int(w.numel()) for w in work), C.byref(nbytes)), "vnd_dsp_workspace")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=first.device)
        with torch.cuda.device(first.device):
            fn_dev(work, ws.data_ptr(), nbytes.value, R.torch_stream_ptr(first))
        if work[-1] is not arrays[-1]:
            arrays[-1].copy_(work[-1])
        return
    if flat:
        arrays = [a.T if (a.ndim == 2 and a.flags.f_contiguous and not a.flags.c_contiguous) else a for a in arrays]
    work = [a if a.flags.c_contiguous else np.ascontiguousarray(a) for a in arrays]
    if not work[-1].flags.writeable:
        raise ValueError("output array is read-only")
    fn_host(work)
    if work[-1] is not arrays[-1]:
        arrays[-1][...] = work[-1]


def _sig(a):
    return R.torch_signal(a) if R.is_torch_tensor(a) else R.host_signal(a)


def rms_normalize(input_signal, output_signal, mode: NormalizeMode = NormalizeMode.DUAL_MONO, epsilon: float = EPSILON) -> None:
    """Scale ``output_signal`` in place to the RMS of ``input_signal`` (utils/dsp.py:87-109).

    DUAL_MONO on 2-D signals - what ``VelvetNoise`` uses - gives every channel its own gain from numpy's sequential
    axis-0 sums; STEREO mode and 1-D signals use one gain from statistics over the whole array, summed in numpy's
    pairwise order.  Both orders are reproduced bit for bit on the device."""
    stereo = mode == NormalizeMode.STEREO
    if input_signal.ndim not in (1, 2) or output_signal.ndim not in (1, 2):
        raise ValueError(f"Input shape invalid: Expected shape (num samples,) or (num samples, channels), but got shape {tuple(input_signal.shape)}.")
    in_flat, out_flat = input_signal.ndim == 1 or stereo, output_signal.ndim == 1 or stereo
    if not in_flat and not out_flat:  # 2-D, DUAL_MONO
        if epsilon != EPSILON:
            raise NotImplementedError("rms_normalize in DUAL_MONO mode supports the default epsilon only")
        check_stereo(input_signal)
        check_stereo(output_signal)
        check_equal_length(input_signal, output_signal)
        _stereo_op(output_signal, _OP_RMS, dry=input_signal)
        return
    if in_flat != out_flat:
        raise NotImplementedError("rms_normalize of a 1-D signal against a 2-D signal in DUAL_MONO mode is not provided")
    if input_signal.dtype != output_signal.dtype:
        input_signal = input_signal.to(output_signal.dtype) if R.is_torch_tensor(input_signal) else input_signal.astype(output_signal.dtype)
    if not R.is_torch_tensor(output_signal):
        _float_array(output_signal, "signal")
    lib = N.lib()
    xn, yn = input_signal.ndim, output_signal.ndim

    def host(w):
        N.check(lib.vnd_rms_normalize_host(R.HostContext.get().handle, C.byref(_sig(w[0])), xn, C.byref(_sig(w[1])), yn, int(stereo), float(epsilon)),
                "vnd_rms_normalize_host")

    def dev(w, ws, nbytes, stream):
        N.check(lib.vnd_rms_normalize_dev(C.byref(_sig(w[0])), xn, C.byref(_sig(w[1])), yn, int(stereo), float(epsilon), ws, nbytes, stream),
                "vnd_rms_normalize_dev")

    _flat_inplace(host, dev, [input_signal, output_signal], True)


def peak_normalize(input_signal, mode: NormalizeMode = NormalizeMode.DUAL_MONO, epsilon: float = EPSILON) -> None:
    """Scale ``input_signal`` in place to ``[-1, 1]`` by its peak: one peak for 1-D signals and STEREO mode, one per
    channel in DUAL_MONO mode (utils/dsp.py:71-84)."""
    stereo = mode == NormalizeMode.STEREO
    if input_signal.ndim not in (1, 2):
        raise ValueError(f"Input shape invalid: Expected shape (num samples,) or (num samples, channels), but got shape {tuple(input_signal.shape)}.")
    if not R.is_torch_tensor(input_signal):
        _float_array(input_signal, "signal")
    lib = N.lib()
    nd = input_signal.ndim

    def host(w):
        N.check(lib.vnd_peak_normalize_host(R.HostContext.get().handle, C.byref(_sig(w[0])), nd, int(stereo), float(epsilon)), "vnd_peak_normalize_host")

    def dev(w, ws, nbytes, stream):
        N.check(lib.vnd_peak_normalize_dev(C.byref(_sig(w[0])), nd, int(stereo), float(epsilon), ws, nbytes, stream), "vnd_peak_normalize_dev")

    _flat_inplace(host, dev, [input_signal], nd == 1 or stereo)


def polar_coordinates(left, right, mode: LayoutMode = "MS", semicircular: bool = True, normalize: bool = True, compute_weights: bool = True):
    """Each frame of ``left`` / ``right`` as polar coordinates: ``(radii, thetas, weights)`` or ``(radii, thetas)``
    (utils/dsp.py:374-422).  ``thetas = arctan2(L - R, L + R)`` in MS mode (``arctan2(L, R)`` otherwise), folded onto
    ``[-pi/2, pi/2]`` when ``semicircular``; ``radii = sqrt(L^2 + R^2)``, divided by their maximum when ``normalize``;
    ``weights = radii / (radii.sum() + 1e-10)``.  Radii and weights equal numpy's bit for bit (the sum runs in numpy's
    pairwise order); the angles are within 2 ulp of numpy's ``arctan2``."""
    ms = int(mode == LayoutMode.MS)
    lib = N.lib()
    if R.is_torch_tensor(left):
        import torch

        l, r = left.contiguous(), right.to(left.dtype).contiguous()
        if l.dim() != 1 or l.shape != r.shape or l.dtype not in (torch.float32, torch.float64):
            raise ValueError("left and right must be 1-D float32/float64 tensors of equal length")
        n = l.shape[0]
        rad, th = torch.empty_like(l), torch.empty_like(l)
        w = torch.empty_like(l) if compute_weights else None
        nbytes = C.c_size_t()
        N.check(lib.vnd_dsp_workspace(n, C.byref(nbytes)), "vnd_dsp_workspace")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=l.device)
        with torch.cuda.device(l.device):
            N.check(lib.vnd_polar_dev(l.data_ptr(), r.data_ptr(), n, N.VND_F64 if l.dtype == torch.float64 else N.VND_F32, ms, int(semicircular),
                                      int(normalize), rad.data_ptr(), th.data_ptr(), w.data_ptr() if w is not None else None, ws.data_ptr(),
                                      nbytes.value, R.torch_stream_ptr(l)), "vnd_polar_dev")
        return (rad, th, w) if compute_weights else (rad, th)
    l = np.asarray(left)
    r = np.asarray(right)
    dt = np.result_type(l.dtype, r.dtype)
    if dt not in (np.float32, np.float64):
        dt = np.dtype(np.float64) if dt.itemsize > 4 or dt.kind in "iu" and dt.itemsize >= 4 else np.dtype(np.float32)
    l = np.ascontiguousarray(l, dtype=dt)
    r = np.ascontiguousarray(r, dtype=dt)
    if l.ndim != 1 or l.shape != r.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {l.shape} {r.shape}")
    n = l.shape[0]
    rad, th = np.empty(n, dtype=dt), np.empty(n, dtype=dt)
    w = np.empty(n, dtype=dt) if compute_weights else None
    N.check(lib.vnd_polar_host(R.HostContext.get().handle, l.ctypes.data, r.ctypes.data, n, N.VND_F64 if dt == np.float64 else N.VND_F32, ms,
                               int(semicircular), int(normalize), rad.ctypes.data, th.ctypes.data, w.ctypes.data if w is not None else None),
            "vnd_polar_host")
    return (rad, th, w) if compute_weights else (rad, th)


# ---- tap-position maths (host; utils/dsp.py:170-286) --------------------------------------------


def generate_log_distribution(strength: float, size: int) -> np.ndarray:
    return log_distribution(strength, size)


def apply_log_distribution(randoms, log_distribution, log_impulse_intervals, jitter: float) -> np.ndarray:
    """``round(randoms * max(0, log_distribution * jitter - 1) + log_impulse_intervals)`` as int32."""
    return np.round(randoms * np.fmax(0.0, log_distribution * jitter - 1) + log_impulse_intervals).astype(np.int32)


def uniform_density(randoms, impulse_indexes, impulse_interval: float) -> np.ndarray:
    """``round(index * interval + randoms * (interval - 1))`` as int32 (utils/dsp.py:253-286)."""
    return np.round(impulse_indexes * impulse_interval + randoms * (impulse_interval - 1)).astype(np.int32)
