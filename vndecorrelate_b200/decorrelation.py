"""Decorrelators with the reference's API (``src/vndecorrelate/decorrelation.py``) running on the
sm_100a kernels of ``libvnd_b200.so``.

Same names, keyword arguments, defaults, return dtypes/shapes and exceptions as the reference for
the hot path: ``VelvetNoise`` (``decorrelate`` / ``convolve`` / ``FIR`` / ``velvet_noise``),
``HaasEffect``, ``SignalChain`` (``velvet_noise`` / ``haas_effect`` / ``stateless``),
``generate_velvet_noise`` and ``convolve_velvet_noise``.  Inputs may be numpy arrays (numpy comes
back, copies inside the call) or CUDA ``torch.Tensor``s (tensors come back, no host round trip).
A ``[VelvetNoise, HaasEffect(LR)]`` pair in a ``SignalChain`` runs as ONE fused launch sequence.

``WhiteNoise`` (a dense Gaussian FIR, decorrelation.py:670-716) is outside the sparse hot path and
is not provided; ``SignalChain.white_noise`` raises ``NotImplementedError``.
"""

from __future__ import annotations

import ctypes as C
from functools import partial
from typing import Any, Callable, Sequence

import numpy as np

from . import _native as N
from . import runtime as R
from .taps import (
    IDENTITY_ENVELOPE,
    TapProgram,
    TapTable,
    ascending_program,
    generate_dense_fir,
    generate_tap_table,
    segmented_program,
)
from .utils.dsp import LayoutMode, apply_stereo_width, check_equal_length, rms_normalize, to_float32

__all__ = ["Decorrelator", "SignalChain", "HaasEffect", "VelvetNoise", "generate_velvet_noise", "convolve_velvet_noise"]


class Decorrelator:
    """Base of the decorrelators: ``num_outs``, ``width``, ``decorrelate`` and ``__call__``
    (decorrelation.py:40-54)."""

    sample_rate_hz: int
    num_outs: int = 2
    width: float | None = None

    def decorrelate(self, input_signal):  # pragma: no cover - abstract
        raise NotImplementedError

    def __call__(self, input_signal):
        return self.decorrelate(input_signal)


# ------------------------------------------------------------------------------------------------
# low-level runners shared by the classes
# ------------------------------------------------------------------------------------------------


def _is_mode(mode, which: LayoutMode) -> bool:
    return mode == which  # StrEnum compares equal to plain 'LR' / 'MS' strings (decorrelation.py:217,433)


def _fir_compute_array(x: np.ndarray) -> np.ndarray:
    """The array numpy would effectively feed ``f32_buffer -= x[i:]`` with: float32 stays float32,
    float64 stays float64, small integers behave like float32 and 32/64-bit integers like float64
    (numpy type promotion with a float32 accumulator)."""
    if x.dtype in (np.float32, np.float64):
        return x
    if x.dtype == np.float16 or (x.dtype.kind in "iub" and x.dtype.itemsize <= 2):
        return x.astype(np.float32)
    if x.dtype.kind in "iuf":
        return x.astype(np.float64)
    raise TypeError(f"unsupported input dtype {x.dtype}")


def _sparse_fir(x, program: TapProgram, num_outs: int):
    """``(frames, >= num_outs)`` signal -> float32 ``(frames, num_outs)``."""
    lib = N.lib()
    if R.is_torch_tensor(x):
        import torch

        if not x.is_cuda:
            raise N.VndError(N.VND_EUNSUPPORTED, "sparse FIR", "CPU torch tensors are not supported (no CPU fallback); pass a CUDA tensor or a numpy array")
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float32 if (x.dtype == torch.float16 or x.element_size() <= 2) else torch.float64)
        frames = x.shape[0]
        # planar inputs (a transposed (C, n) tensor) get a planar output, so loads and stores coalesce
        planar = x.dim() == 2 and x.shape[1] > 1 and x.stride(0) == 1
        y = torch.empty((num_outs, frames), dtype=torch.float32, device=x.device).t() if planar else torch.empty((frames, num_outs), dtype=torch.float32, device=x.device)
        sx, sy = R.torch_signal(x), R.torch_signal(y)
        with torch.cuda.device(x.device):
            ps = R.device_program(program, x.device)
            N.check(lib.vnd_sparse_fir_dev(C.byref(sx), C.byref(sy), C.byref(ps), R.torch_stream_ptr(x)), "vnd_sparse_fir_dev")
        return y
    xa = R.dense(_fir_compute_array(np.asarray(x)))
    frames = xa.shape[0]
    planar = xa.ndim == 2 and xa.shape[1] > 1 and not xa.flags.c_contiguous
    y = R.pinned_empty((num_outs, frames), np.float32).T if planar else R.pinned_empty((frames, num_outs), np.float32)
    sx, sy = R.host_signal(xa), R.host_signal(y)
    ps = program.host_struct()
    N.check(lib.vnd_sparse_fir_host(R.HostContext.get().handle, C.byref(sx), C.byref(sy), C.byref(ps)), "vnd_sparse_fir_host")
    return y


def _vn_decorrelate(x32, program: TapProgram, *, num_outs: int, ms_encode: bool, width, rms: bool, haas_delay: int = 0,
                    haas_channel: int = 0, out_f64: bool = False):
    """Fused decorrelate on a float32 signal (1-D mono or ``(frames, >= num_outs)``)."""
    lib = N.lib()
    ep = N.EpilogueStruct(int(ms_encode), int(width is not None), float(width) if width is not None else 0.0, int(rms), int(haas_delay), int(haas_channel))
    mono = x32.ndim == 1
    frames = x32.shape[0]
    if R.is_torch_tensor(x32):
        import torch

        if not x32.is_cuda:
            raise N.VndError(N.VND_EUNSUPPORTED, "decorrelate", "CPU torch tensors are not supported (no CPU fallback); pass a CUDA tensor or a numpy array")
        out = torch.empty((frames + haas_delay, num_outs), dtype=torch.float64 if out_f64 else torch.float32, device=x32.device)
        sx = R.torch_signal(x32, mono_as_stereo=mono)
        so = R.torch_signal(out)
        nbytes = C.c_size_t()
        N.check(lib.vnd_vn_decorrelate_workspace(frames, num_outs, C.byref(ep), C.byref(nbytes)), "vnd_vn_decorrelate_workspace")
        with torch.cuda.device(x32.device):
            work = torch.empty(nbytes.value, dtype=torch.uint8, device=x32.device)
            ps = R.device_program(program, x32.device)
            N.check(lib.vnd_vn_decorrelate_dev(C.byref(sx), C.byref(so), C.byref(ps), C.byref(ep), work.data_ptr(), nbytes.value,
                                               R.torch_stream_ptr(x32)), "vnd_vn_decorrelate_dev")
        return out
    xa = R.dense(x32)
    out = R.pinned_empty((frames + haas_delay, num_outs), np.float64 if out_f64 else np.float32)
    sx = R.host_signal(xa, mono_as_stereo=mono)
    so = R.host_signal(out)
    ps = program.host_struct()
    N.check(lib.vnd_vn_decorrelate_host(R.HostContext.get().handle, C.byref(sx), C.byref(so), C.byref(ps), C.byref(ep)), "vnd_vn_decorrelate_host")
    return out


def _haas(x32, *, delay: int, delayed_channel: int, mode_ms: bool, width):
    lib = N.lib()
    mono = x32.ndim == 1
    frames = x32.shape[0]
    args = (int(delay), int(delayed_channel), int(mode_ms), int(mono), int(width is not None), float(width) if width is not None else 0.0)
    if R.is_torch_tensor(x32):
        import torch

        if not x32.is_cuda:
            raise N.VndError(N.VND_EUNSUPPORTED, "haas", "CPU torch tensors are not supported (no CPU fallback)")
        out = torch.empty((frames + delay, 2), dtype=torch.float64, device=x32.device)
        sx, so = R.torch_signal(x32, mono_as_stereo=mono), R.torch_signal(out)
        with torch.cuda.device(x32.device):
            N.check(lib.vnd_haas_dev(C.byref(sx), C.byref(so), *args, R.torch_stream_ptr(x32)), "vnd_haas_dev")
        return out
    xa = R.dense(x32)
    out = R.pinned_empty((frames + delay, 2), np.float64)
    sx, so = R.host_signal(xa, mono_as_stereo=mono), R.host_signal(out)
    N.check(lib.vnd_haas_host(R.HostContext.get().handle, C.byref(sx), C.byref(so), *args), "vnd_haas_host")
    return out


# ------------------------------------------------------------------------------------------------
# Haas effect (decorrelation.py:163-230)
# ------------------------------------------------------------------------------------------------


class HaasEffect(Decorrelator):
    """Delays one channel (L/R, or mid/side in MS mode) by ``delay_time_seconds``; returns float64
    ``(n + round(delay * fs), 2)`` like the reference."""

    def __init__(self, *, sample_rate_hz: int, num_outs: int = 2, width: float | None = None, delayed_channel: int = 0,
                 delay_time_seconds: float = 0.02, mode: LayoutMode = LayoutMode.LR):
        self.sample_rate_hz = sample_rate_hz
        self.num_outs = num_outs
        self.width = width
        self.delayed_channel = delayed_channel
        self.delay_time_seconds = delay_time_seconds
        self.mode = mode

    @property
    def delay_len_samples(self) -> int:
        return round(self.delay_time_seconds * self.sample_rate_hz)  # Python's round-half-even, decorrelation.py:204

    def _check(self, x) -> None:
        if x.ndim not in (1, 2) or (x.ndim == 2 and x.shape[1] != 2):
            raise ValueError(f"could not broadcast input array from shape {tuple(x.shape)} into shape ({x.shape[0]},2)")
        if self.delayed_channel not in (0, 1, -1, -2):
            raise IndexError(f"index {self.delayed_channel} is out of bounds for axis 1 with size 2")

    def haas_delay(self, input_signal):
        """The delay alone (no width), as ``HaasEffect.haas_delay`` (decorrelation.py:202-230)."""
        self._check(input_signal)
        return _haas(to_float32(input_signal), delay=self.delay_len_samples, delayed_channel=self.delayed_channel % 2,
                     mode_ms=_is_mode(self.mode, LayoutMode.MS), width=None)

    def decorrelate(self, input_signal):
        input_signal = to_float32(input_signal if R.is_torch_tensor(input_signal) else np.asarray(input_signal))
        self._check(input_signal)
        return _haas(input_signal, delay=self.delay_len_samples, delayed_channel=self.delayed_channel % 2,
                     mode_ms=_is_mode(self.mode, LayoutMode.MS), width=self.width)


# ------------------------------------------------------------------------------------------------
# Velvet noise (decorrelation.py:326-546)
# ------------------------------------------------------------------------------------------------


class VelvetNoise(Decorrelator):
    """Velvet-noise decorrelator: a sparse FIR of ``num_impulses`` signed unit impulses per channel
    with a segmented decay envelope, applied anti-causally (``y[t] = sum c_k x[t + i_k]``), then
    optional M/S side-channel encode, width and RMS normalisation.

    Constructor arguments, defaults and attribute semantics are the reference's
    (decorrelation.py:355-362): the tap table is generated at construction from ``seed`` and
    regenerated only when ``num_outs``, ``num_impulses`` or the FIR length change; a new
    ``segment_envelope`` is picked up at call time."""

    def __init__(self, *, sample_rate_hz: int, num_outs: int = 2, width: float | None = None, duration_seconds: float = 0.03,
                 num_impulses: int = 30, segment_envelope: Sequence[float] = (0.85, 0.55, 0.35, 0.2),
                 log_distribution_strength: float = 1.0, normalizer: Callable[[Any, Any], None] | None = rms_normalize,
                 filtered_channels: Sequence[int] = (0, 1), mode: LayoutMode = LayoutMode.MS, seed: int | None = None):
        self.sample_rate_hz = sample_rate_hz
        self.num_outs = num_outs
        self.width = width
        self.duration_seconds = duration_seconds
        self.num_impulses = num_impulses
        self.segment_envelope = segment_envelope
        self.log_distribution_strength = log_distribution_strength
        self.normalizer = normalizer
        self.filtered_channels = filtered_channels
        self.mode = mode
        self.seed = seed
        if self.num_impulses >= self.fir_length_samples * 0.2:  # decorrelation.py:382-387
            raise ValueError(
                f"Velvet Noise Filter of length {self.fir_length_samples} with {self.num_impulses} impulses is not sparse. "
                f"(density={self.density:.2f})\n\tnum_impulses must be less than 20% the FIR length in samples."
            )
        if not self.segment_envelope:
            self.segment_envelope = IDENTITY_ENVELOPE
        self._velvet_noise: TapTable = self._generate()
        self._programs: dict = {}

    # ---- properties of the reference ------------------------------------------------------------
    @property
    def density(self) -> float:
        return self.num_impulses / self.duration_seconds

    @property
    def fir_length_samples(self) -> int:
        return int(round(self.sample_rate_hz * self.duration_seconds))

    @property
    def unfiltered_channels(self):
        return filter(lambda i: i not in self.filtered_channels, range(self.num_outs))

    @property
    def velvet_noise(self) -> TapTable:
        """The tap table; regenerated when ``num_outs``, ``num_impulses`` or the FIR length changed
        since it was made (decorrelation.py:368-379)."""
        t = self._velvet_noise
        if self.num_outs != t.num_outs or self.num_impulses != t.num_impluses or self.fir_length_samples != t.fir_length_samples:
            self._velvet_noise = self._generate()
            self._programs.clear()
        return self._velvet_noise

    def _generate(self) -> TapTable:
        return generate_tap_table(
            sample_rate_hz=self.sample_rate_hz, duration_seconds=self.duration_seconds, num_impulses=self.num_impulses,
            num_outs=self.num_outs, num_segments=len(self.segment_envelope), log_distribution_strength=self.log_distribution_strength,
            filtered_channels=self.filtered_channels, seed=self.seed,
        )

    @property
    def FIR(self) -> np.ndarray:
        """Dense float64 ``(fir_length_samples, len(filtered_channels))`` impulse responses
        (decorrelation.py:454-472)."""
        table = self.velvet_noise
        num_filters = len(self.filtered_channels)
        fir = np.zeros((self.fir_length_samples, num_filters))
        chans = np.nonzero(table.filtered)[0]
        if len(chans) != num_filters:
            raise ValueError("setting an array element with a sequence. The requested array has an inhomogeneous shape")
        for col, c in enumerate(chans):
            for s, seg in enumerate(table[c]):
                for indices, sign in ((seg[0], -1), (seg[1], 1)):
                    for i in indices:
                        fir[i, col] = self.segment_envelope[s] * sign
        return fir

    # ---- tap programs ---------------------------------------------------------------------------
    def tap_program(self, frames: int) -> TapProgram:
        """The packed program for a signal of ``frames`` samples (cached; the key includes the
        envelope because it may be replaced after construction, tests/test_decorrelation.py:148-158)."""
        table = self.velvet_noise
        env = self.segment_envelope
        max_index = int(table.index[table.filtered].max()) if table.filtered.any() and table.num_impulses else -1
        key = (min(frames, max_index + 1), tuple(env), env == IDENTITY_ENVELOPE)
        prog = self._programs.get(key)
        if prog is None:
            if len(self._programs) > 8:
                self._programs.clear()
            prog = segmented_program(table, env, frames)
            self._programs[key] = prog
        return prog

    # ---- the hot path ---------------------------------------------------------------------------
    def convolve(self, input_signal):
        """Sparse FIR on each filtered channel, other channels copied; float32 ``(n, num_outs)``
        (decorrelation.py:393-415).  ``input_signal`` must be 2-D with at least ``num_outs`` columns."""
        if input_signal.ndim != 2:
            raise IndexError(f"too many indices for array: array is {input_signal.ndim}-dimensional, but 2 were indexed")
        if input_signal.shape[1] < self.num_outs:
            raise IndexError(f"index {input_signal.shape[1]} is out of bounds for axis 1 with size {input_signal.shape[1]}")
        return _sparse_fir(input_signal, self.tap_program(input_signal.shape[0]), self.num_outs)

    def decorrelate(self, input_signal):
        """cast -> (mono -> stereo) -> FIR -> M/S encode -> width -> normaliser; float32
        ``(n, num_outs)`` (decorrelation.py:417-442)."""
        x = to_float32(input_signal if R.is_torch_tensor(input_signal) else np.asarray(input_signal))
        return self._decorrelate(x, haas=None)

    def _decorrelate(self, x, haas: "HaasEffect | None"):
        if x.ndim == 1:
            in_channels = 2  # mono_to_stereo, passed to the kernel as a stride-0 broadcast
        elif x.ndim == 2:
            in_channels = x.shape[1]
        else:
            raise ValueError(f"Input shape invalid: Expected shape (num samples,) or (num samples, channels), but got shape {tuple(x.shape)}.")
        if in_channels < self.num_outs:
            raise IndexError(f"index {in_channels} is out of bounds for axis 1 with size {in_channels}")
        ms = _is_mode(self.mode, LayoutMode.MS)
        if ms and (in_channels != 2 or self.num_outs != 2):  # check_stereo in encode_signal_to_side_channel
            bad = (x.shape[0], in_channels) if in_channels != 2 else (x.shape[0], self.num_outs)
            raise ValueError(f"Input shape invalid: Expected shape (num samples, 2), but got shape {bad}.")
        if self.width is not None and self.num_outs != 2:
            raise ValueError(f"Input shape invalid: Expected shape (num samples, 2), but got shape {(x.shape[0], self.num_outs)}.")
        fused_norm = self.normalizer is None or self.normalizer is rms_normalize
        if self.normalizer is rms_normalize and in_channels != self.num_outs:
            raise ValueError(f"operands could not be broadcast together with shapes ({x.shape[0]},{self.num_outs}) ({in_channels},)")
        program = self.tap_program(x.shape[0])
        kw = dict(num_outs=self.num_outs, ms_encode=ms, width=self.width)
        if fused_norm:
            if haas is not None:
                return _vn_decorrelate(x, program, rms=self.normalizer is not None, haas_delay=haas.delay_len_samples,
                                       haas_channel=haas.delayed_channel % 2, out_f64=True, **kw)
            return _vn_decorrelate(x, program, rms=self.normalizer is not None, **kw)
        # user-supplied normaliser: run everything else fused, then hand (input, output) to it
        y = _vn_decorrelate(x, program, rms=False, **kw)
        x2 = x if x.ndim == 2 else (np.column_stack((x, x)) if not R.is_torch_tensor(x) else x[:, None].expand(-1, 2))
        self.normalizer(x2, y)
        return y


def _fusable(vn, haas) -> bool:
    return (
        isinstance(vn, VelvetNoise) and isinstance(haas, HaasEffect) and vn.num_outs == 2 and haas.width is None
        and _is_mode(haas.mode, LayoutMode.LR) and (vn.normalizer is None or vn.normalizer is rms_normalize)
        and haas.delayed_channel in (0, 1, -1, -2)
    )


# ------------------------------------------------------------------------------------------------
# Signal chain (decorrelation.py:71-153)
# ------------------------------------------------------------------------------------------------


class SignalChain:
    """Builder of cascaded decorrelators.  Stages are instantiated on the first call unless
    ``lazy=False``.  A ``VelvetNoise`` stage directly followed by an LR-mode ``HaasEffect`` without
    width is executed as one fused device pass (one read and one write per sample when the velvet
    stage has no normaliser; plus the RMS pre-pass otherwise)."""

    def __init__(self, *, sample_rate_hz: int, num_outs: int = 2, lazy: bool = True, _hot: bool = False, _decorrelators=None):
        if _decorrelators is not None:
            raise TypeError(
                "Cannot supply decorrelators directly, use ``SignalChain.velvet_noise``,"
                " ``SignalChain.haas_effect``, or ``SignalChain.stateless``."
            )
        self.sample_rate_hz = sample_rate_hz
        self.num_outs = num_outs
        self.lazy = lazy
        self._hot = _hot or not lazy
        self._decorrelators: list = []

    def haas_effect(self, **kwargs):
        self._add_decorrelator(HaasEffect, **kwargs)
        return self

    def velvet_noise(self, **kwargs):
        self._add_decorrelator(VelvetNoise, **kwargs)
        return self

    def white_noise(self, **kwargs):
        raise NotImplementedError("WhiteNoise is a dense FIR outside the sparse velvet-noise hot path; it is not provided by vndecorrelate_b200")

    def stateless(self, function, *args, **kwargs):
        make = lambda: partial(function, *args, **kwargs)  # noqa: E731
        self._decorrelators.append(make() if self._hot else make)
        return self

    def _add_decorrelator(self, cls, **kwargs) -> None:
        kwargs = self._validate(cls, **kwargs)

        def make():
            if "num_outs" in kwargs:  # the reference passes num_outs twice here (decorrelation.py:115-119)
                raise TypeError(f"{cls.__module__}.{cls.__name__}() got multiple values for keyword argument 'num_outs'")
            return cls(sample_rate_hz=self.sample_rate_hz, num_outs=self.num_outs, **kwargs)

        self._decorrelators.append(make() if self._hot else make)

    def _validate(self, cls, *, sample_rate_hz: int | None = None, **kwargs) -> dict[str, Any]:
        if sample_rate_hz is not None and sample_rate_hz != self.sample_rate_hz:
            raise TypeError(
                f"{sample_rate_hz=} was supplied to {cls} but differs from the sample rate of the enclosing ``SignalChain`` ({self.sample_rate_hz})"
            )
        return kwargs

    def _init_decorrelators(self) -> None:
        if self._hot:
            return
        self._decorrelators = [make() for make in self._decorrelators]
        self._hot = True

    def __call__(self, input_signal):
        self._init_decorrelators()
        stages = self._decorrelators
        sig = input_signal
        i = 0
        while i < len(stages):
            if i + 1 < len(stages) and _fusable(stages[i], stages[i + 1]):
                x = to_float32(sig if R.is_torch_tensor(sig) else np.asarray(sig))
                sig = stages[i]._decorrelate(x, haas=stages[i + 1])
                i += 2
            else:
                sig = stages[i](sig)
                i += 1
        return sig


# ------------------------------------------------------------------------------------------------
# function path (decorrelation.py:549-660)
# ------------------------------------------------------------------------------------------------


def generate_velvet_noise(*, duration_seconds: float, num_impulses: int, num_outs: int = 2, sample_rate_hz: int = 44100,
                          segment_envelope: Sequence[float] = (0.85, 0.55, 0.35, 0.2), log_distribution_strength: float = 1.0,
                          seed: int | None = None) -> np.ndarray:
    """Dense float32 ``(int(duration * fs), num_outs)`` velvet-noise FIR for ``convolve_velvet_noise``."""
    return generate_dense_fir(duration_seconds=duration_seconds, num_impulses=num_impulses, num_outs=num_outs, sample_rate_hz=sample_rate_hz,
                              segment_envelope=segment_envelope, log_distribution_strength=log_distribution_strength, seed=seed)


def convolve_velvet_noise(input_signal, velvet_noise_filters):
    """Stateless sparse FIR: for each non-zero of each FIR column in ascending index order,
    ``y[:n - i] += x[i:] * value``; float32 output with the input's shape (decorrelation.py:630-660)."""
    if input_signal.ndim == 1:  # the reference indexes [:, channel] on the 1-D input (decorrelation.py:650)
        raise IndexError("too many indices for array: array is 1-dimensional, but 2 were indexed")
    fir = velvet_noise_filters.detach().cpu().numpy() if R.is_torch_tensor(velvet_noise_filters) else np.asarray(velvet_noise_filters)
    if input_signal.shape[1] > 1:
        check_equal_length(input_signal, fir, dim=1)
    program = ascending_program(fir, input_signal.shape[0])
    return _sparse_fir(input_signal, program, input_signal.shape[1])
