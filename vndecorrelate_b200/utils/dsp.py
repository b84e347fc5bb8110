"""Stereo helpers of the hot path, with the reference's names and in-place contracts
(``src/vndecorrelate/utils/dsp.py``), executed by the CUDA library.

Only what ``VelvetNoise.decorrelate`` / ``HaasEffect`` / the optimiser touch is here: the M/S
helpers, width, side-channel encode, RMS normalisation, the dtype/shape helpers and the tap-position
maths.  Analysis and plotting helpers of the reference (``cross_correlogram``, ``sine_sweep``,
``peak_normalize`` …) are outside the hot path and are not provided.

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
    "LR_to_MS", "MS_to_LR", "generate_log_distribution", "apply_log_distribution", "uniform_density",
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
        work = torch.empty(256, dtype=torch.uint8, device=a.device)
        N.check(N.lib().vnd_stereo_op_dev(C.byref(sa), C.byref(sd) if sd is not None else None, op, float(width),
                                          work.data_ptr(), 256, R.torch_stream_ptr(a)), "vnd_stereo_op_dev")
        return
    _float_array(a, "signal")
    work = a if (a.flags.c_contiguous and a.flags.writeable) else np.ascontiguousarray(a)
    sa = R.host_signal(work)
    sd = None
    if dry is not None:
        dry = np.ascontiguousarray(dry, dtype=a.dtype)
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


def rms_normalize(input_signal, output_signal, mode: NormalizeMode = NormalizeMode.DUAL_MONO, epsilon: float = EPSILON) -> None:
    """Scale each channel of ``output_signal`` in place to the RMS of the same channel of
    ``input_signal`` (utils/dsp.py:87-109), reproducing numpy's sequential axis-0 summation.

    Only the DUAL_MONO mode on ``(n, 2)`` signals — what ``VelvetNoise`` uses — is on the hot
    path; STEREO mode, 1-D signals and a non-default epsilon are not provided."""
    if mode != NormalizeMode.DUAL_MONO or input_signal.ndim != 2 or output_signal.ndim != 2 or epsilon != EPSILON:
        raise NotImplementedError("only rms_normalize(x, y) in DUAL_MONO mode on 2-D signals is part of the accelerated hot path")
    check_stereo(input_signal)
    check_stereo(output_signal)
    check_equal_length(input_signal, output_signal)
    _stereo_op(output_signal, _OP_RMS, dry=input_signal)


# ---- tap-position maths (host; utils/dsp.py:170-286) --------------------------------------------


def generate_log_distribution(strength: float, size: int) -> np.ndarray:
    return log_distribution(strength, size)


def apply_log_distribution(randoms, log_distribution, log_impulse_intervals, jitter: float) -> np.ndarray:
    """``round(randoms * max(0, log_distribution * jitter - 1) + log_impulse_intervals)`` as int32."""
    return np.round(randoms * np.fmax(0.0, log_distribution * jitter - 1) + log_impulse_intervals).astype(np.int32)


def uniform_density(randoms, impulse_indexes, impulse_interval: float) -> np.ndarray:
    """``round(index * interval + randoms * (interval - 1))`` as int32 (utils/dsp.py:253-286)."""
    return np.round(impulse_indexes * impulse_interval + randoms * (impulse_interval - 1)).astype(np.int32)
