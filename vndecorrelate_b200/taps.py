"""Velvet-noise tap tables on the host, and their packing into the kernels' tap programs.

Impulse positions and signs come from the same numpy PCG64 draws, in the same order and shapes, as
the reference (``src/vndecorrelate/decorrelation.py:478-546`` for the class path, ``:549-627`` for
``generate_velvet_noise``; position maths ``src/vndecorrelate/utils/dsp.py:170-250``), so a given
seed yields the same filter.  Generation is vectorised over channels (the reference loops), which
keeps a 4096-channel table at a few milliseconds; the per-element float64 operations and their order
are the reference's, so the rounded int32 positions are identical.

A *tap program* is the flat int32 layout ``include/vnd_b200.h`` documents (vnd_tap_program).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from . import _native as N

IDENTITY_ENVELOPE: tuple[float] = (1.0,)  # utils/dsp.py:8


def log_distribution(strength: float, size: int) -> np.ndarray:
    """``generate_log_distribution`` (utils/dsp.py:170-201): ``size + 1`` weights growing as
    ``10 ** (2 * strength * k / size)``, normalised so that strength 0 gives all ones."""
    k = np.arange(size + 1.0) / size
    return (10.0 ** (2.0 * strength * k)) / (100.0 * ((1.0 + (strength * 99.0)) / 100.0))


def _interval_grid(strength: float, num_impulses: int, fir_length: int):
    """Weights and cumulative interval starts scaled to the filter length
    (decorrelation.py:494-506)."""
    w = log_distribution(strength, num_impulses)
    starts = np.cumsum(w)
    if strength == 0.0:
        starts -= 1.0
    starts *= fir_length / starts[-1]
    return w, starts


def _interval_grid_family(strengths, num_impulses: int, fir_length: int):
    """``_interval_grid`` for many strengths at once: ``(K, N + 1)`` weights and interval starts.

    The per-element operations and their order are those of ``_interval_grid`` (elementwise ``**`` and divisions, a
    sequential ``cumsum`` per row), so the rows equal the one-strength results bit for bit; three rows are checked
    against the one-strength function on every call and the loop is used if they ever differ (a numpy build whose
    vector and scalar ``pow`` round differently).  ``tests/test_host_logic.py`` compares 20 000 strengths."""
    ks = np.asarray(strengths, dtype=np.float64)
    K = len(ks)
    col = ks[:, None]
    k = (np.arange(num_impulses + 1.0) / num_impulses)[None, :]
    w = (10.0 ** (2.0 * col * k)) / (100.0 * ((1.0 + (col * 99.0)) / 100.0))
    starts = np.cumsum(w, axis=1)
    starts[ks == 0.0] -= 1.0
    starts *= (fir_length / starts[:, -1])[:, None]
    for i in {0, K // 2, K - 1}:
        w1, s1 = _interval_grid(float(ks[i]), num_impulses, fir_length)
        if not (np.array_equal(w1, w[i]) and np.array_equal(s1, starts[i])):
            grids = [_interval_grid(float(v), num_impulses, fir_length) for v in ks]
            return np.stack([g[0] for g in grids]), np.stack([g[1] for g in grids])
    return w, starts


def _segment_of_impulse(num_impulses: int, num_segments: int) -> np.ndarray:
    """``int(j / (N / S))`` with Python float division (decorrelation.py:540)."""
    return np.array([int(j / (num_impulses / num_segments)) for j in range(num_impulses)], dtype=np.int64)


def _draw(seed, num_impulses: int, num_filters: int):
    """The reference's two ``uniform`` draws (decorrelation.py:488, :510-521)."""
    rng = np.random.default_rng(seed)
    sign_u = rng.uniform(low=0, high=1, size=(num_impulses, num_filters))
    offs_u = rng.uniform(low=0, high=1, size=(num_impulses + 1, num_filters))
    return sign_u, offs_u


class _SegmentView:
    """``taps[channel][segment]``: index 0 -> negative impulse indices, 1 -> positive ones; iterating
    yields ``(indices, '__isub__')`` then ``(indices, '__iadd__')`` like the reference container
    (decorrelation.py:240-271)."""

    __slots__ = ("negative_impulse_indexes", "positive_impulse_indexes")

    def __init__(self, neg, pos):
        self.negative_impulse_indexes = neg
        self.positive_impulse_indexes = pos

    def __getitem__(self, key: int):
        if key == 0:
            return self.negative_impulse_indexes
        if key == 1:
            return self.positive_impulse_indexes
        raise ValueError("Invalid key")

    def __iter__(self):
        return iter(((self.negative_impulse_indexes, "__isub__"), (self.positive_impulse_indexes, "__iadd__")))

    def __eq__(self, other):
        return isinstance(other, _SegmentView) and list(self[0]) == list(other[0]) and list(self[1]) == list(other[1])


class _ChannelView:
    __slots__ = ("segments",)

    def __init__(self, segments):
        self.segments = segments

    def __iter__(self):
        return iter(self.segments)

    def __getitem__(self, key: int):
        return self.segments[key]

    def __len__(self):
        return len(self.segments)

    def __eq__(self, other):
        return isinstance(other, _ChannelView) and self.segments == other.segments


@dataclass(eq=False)
class TapTable:
    """Class-path tap structure for ``num_outs`` output channels.

    ``index[c, j]`` / ``positive[c, j]`` describe impulse ``j`` of channel ``c`` (rows of unfiltered
    channels are unused); ``segment[j]`` is its decay segment.  ``filtered[c]`` is False for channels
    that are copied through (decorrelation.py:399-400, :526-528)."""

    fir_length_samples: int
    num_impulses: int
    num_segments: int
    filtered: np.ndarray  # (num_outs,) bool
    index: np.ndarray  # (num_outs, N) int32
    positive: np.ndarray  # (num_outs, N) bool
    segment: np.ndarray  # (N,) int64
    _views: list | None = field(default=None, repr=False)

    @property
    def num_outs(self) -> int:
        return len(self.filtered)

    @property
    def num_impluses(self) -> int:  # sic — the reference's spelling (decorrelation.py:299)
        return self.num_impulses if self.num_outs and self.filtered[0] else 0

    # --- nested view compatible with the reference's _ParallelVelvetNoise ------------------------
    def _build_views(self):
        if self._views is None:
            views = []
            for c in range(self.num_outs):
                if not self.filtered[c]:
                    views.append([])
                    continue
                segs = []
                for s in range(self.num_segments):
                    m = self.segment == s
                    neg = [i for i in self.index[c, m & ~self.positive[c]]]
                    pos = [i for i in self.index[c, m & self.positive[c]]]
                    segs.append(_SegmentView(neg, pos))
                views.append(_ChannelView(segs))
            self._views = views
        return self._views

    @property
    def output_channels(self):
        return self._build_views()

    def __iter__(self):
        return iter(self._build_views())

    def __getitem__(self, key: int):
        return self._build_views()[key]

    def __eq__(self, other):
        if not isinstance(other, TapTable):
            return NotImplemented
        return (
            self.fir_length_samples == other.fir_length_samples
            and self.num_segments == other.num_segments
            and np.array_equal(self.filtered, other.filtered)
            and np.array_equal(self.index[self.filtered], other.index[other.filtered])
            and np.array_equal(self.positive[self.filtered], other.positive[other.filtered])
        )

    def rows(self) -> np.ndarray:
        """int32 rows ``(channel, segment, index, sign)`` in the reference's iteration order."""
        out = []
        for c, ch in enumerate(self._build_views()):
            for s, seg in enumerate(ch):
                out += [(c, s, int(i), -1) for i in seg[0]]
                out += [(c, s, int(i), 1) for i in seg[1]]
        return np.array(out, dtype=np.int32).reshape(-1, 4)


def generate_tap_table(
    *,
    sample_rate_hz: int,
    duration_seconds: float,
    num_impulses: int,
    num_outs: int,
    num_segments: int,
    log_distribution_strength: float,
    filtered_channels: Sequence[int],
    seed,
) -> TapTable:
    """Tap structure of ``VelvetNoise._generate`` (decorrelation.py:478-546)."""
    fir_length = int(round(sample_rate_hz * duration_seconds))  # decorrelation.py:449-452
    w, starts = _interval_grid(log_distribution_strength, num_impulses, fir_length)
    num_filters = len(filtered_channels)
    sign_u, offs_u = _draw(seed, num_impulses, num_filters)
    jitter = sample_rate_hz / (num_impulses / duration_seconds)  # decorrelation.py:523 with :444-447
    filtered = np.array([c in filtered_channels for c in range(num_outs)], dtype=bool)
    chans = np.nonzero(filtered)[0]
    # The reference indexes the draws by CHANNEL NUMBER (decorrelation.py:531), so a filtered channel
    # whose number is >= len(filtered_channels) is an IndexError there; keep that contract.
    if len(chans) and chans.max() >= num_filters:
        raise IndexError(f"index {int(chans.max())} is out of bounds for axis 1 with size {num_filters}")
    index = np.full((num_outs, num_impulses), -1, dtype=np.int32)
    positive = np.zeros((num_outs, num_impulses), dtype=bool)
    if len(chans):
        u = offs_u[:, chans]  # (N + 1, F)
        pos = np.round(u * np.fmax(0.0, w[:, None] * jitter - 1) + starts[:, None]).astype(np.int32)  # utils/dsp.py:248-250
        index[chans] = pos[:num_impulses].T
        positive[chans] = (np.round(sign_u[:, chans]) == 1.0).T  # 2*round(u) - 1 > 0, decorrelation.py:521
    return TapTable(
        fir_length_samples=fir_length,
        num_impulses=num_impulses,
        num_segments=num_segments,
        filtered=filtered,
        index=index,
        positive=positive,
        segment=_segment_of_impulse(num_impulses, num_segments),
    )


def generate_dense_fir(
    *,
    duration_seconds: float,
    num_impulses: int,
    num_outs: int = 2,
    sample_rate_hz: int = 44100,
    segment_envelope: Sequence[float] = (0.85, 0.55, 0.35, 0.2),
    log_distribution_strength: float = 1.0,
    seed=None,
) -> np.ndarray:
    """``generate_velvet_noise`` (decorrelation.py:549-627): dense fp32 ``(int(dur * fs), num_outs)``
    FIR.  On an index collision the later impulse overwrites the earlier one, as in the reference."""
    fir_length = int(duration_seconds * sample_rate_hz)  # truncation, decorrelation.py:575
    env = tuple(segment_envelope) if len(segment_envelope) else IDENTITY_ENVELOPE
    w, starts = _interval_grid(log_distribution_strength, num_impulses, fir_length)
    sign_u, offs_u = _draw(seed, num_impulses, num_outs)
    jitter = sample_rate_hz / (num_impulses / duration_seconds)
    pos = np.round(offs_u * np.fmax(0.0, w[:, None] * jitter - 1) + starts[:, None]).astype(np.int32)[:num_impulses]
    signs = (2 * np.round(sign_u)) - 1
    seg = _segment_of_impulse(num_impulses, len(env))
    gains = np.array([env[s] for s in seg], dtype=np.float64)
    fir = np.zeros((fir_length, num_outs), dtype=np.float32)
    cols = np.arange(num_outs)
    for j in range(num_impulses):  # impulse order decides who wins a collision
        fir[pos[j], cols] = signs[j] * gains[j]
    return fir


# --------------------------------------------------------------------------------------------------
# tap programs
# --------------------------------------------------------------------------------------------------


@dataclass
class TapProgram:
    """Host copy of a vnd_tap_program plus lazily created device copies."""

    words: np.ndarray  # int32
    offsets: np.ndarray  # int32, channels + 1
    channels: int
    order: int
    apply_gain: int
    halo: int
    max_channel_words: int
    _device: dict = field(default_factory=dict, repr=False)

    def struct(self, words_ptr: int, offsets_ptr: int) -> N.TapProgramStruct:
        return N.TapProgramStruct(words_ptr, offsets_ptr, int(self.words.size), self.channels, self.order, self.apply_gain, self.halo, self.max_channel_words)

    def host_struct(self) -> N.TapProgramStruct:
        return self.struct(self.words.ctypes.data, self.offsets.ctypes.data)

    def slice_channels(self, start: int, stop: int) -> "TapProgram":
        """Program for channels ``[start, stop)`` — what one rank of a channel-sharded job uses.
        The table is always generated once for ALL channels and then sliced: the RNG draw layout
        depends on the total channel count (decorrelation.py:510-520)."""
        lo, hi = int(self.offsets[start]), int(self.offsets[stop])
        offs = (self.offsets[start : stop + 1] - lo).astype(np.int32)
        words = np.ascontiguousarray(self.words[lo:hi])
        sizes = np.diff(offs)
        return TapProgram(words, offs, stop - start, self.order, self.apply_gain, self.halo, int(sizes.max()) if len(sizes) else 0)


def _pack_blocks(offsets_len: np.ndarray):
    offsets = np.zeros(len(offsets_len) + 1, dtype=np.int64)
    np.cumsum(offsets_len, out=offsets[1:])
    if offsets[-1] >= 2**31:
        raise ValueError("tap program too large")
    return offsets


def segmented_program(table: TapTable, envelope: Sequence[float], frames: int) -> TapProgram:
    """Pack a class-path table for a signal of ``frames`` samples (VND_ORDER_SEGMENTED).

    Taps with index >= frames are removed — the reference's slices are empty for them
    (decorrelation.py:404-410).  Segments left without taps are removed as well when their gain is
    finite: they would add ``0 * gain = 0`` to the output, which changes nothing."""
    identity = envelope == IDENTITY_ENVELOPE  # the reference compares with the tuple (decorrelation.py:411)
    S = table.num_segments
    if not identity and len(envelope) < S and table.filtered.any():
        raise IndexError("tuple index out of range")  # envelope[segment_index], decorrelation.py:412
    gains = np.ones(S, dtype=np.float32) if identity else np.array([float(envelope[s]) for s in range(S)], dtype=np.float64).astype(np.float32)
    C, Nimp = table.index.shape
    chans = np.nonzero(table.filtered)[0]
    F = len(chans)
    block_len = np.zeros(C, dtype=np.int64)
    halo = 0
    if F and Nimp:
        idx = table.index[chans]  # (F, N)
        key = table.segment[None, :] * 2 + table.positive[chans]  # iteration order: segment, then neg before pos
        order = np.argsort(key, axis=1, kind="stable")
        idx_s = np.take_along_axis(idx, order, axis=1)
        key_s = np.take_along_axis(key, order, axis=1)
        valid = idx_s < frames
        counts = np.zeros((F, 2 * S), dtype=np.int64)
        rows = np.broadcast_to(np.arange(F)[:, None], idx_s.shape)
        np.add.at(counts, (rows[valid], key_s[valid]), 1)
        counts = counts.reshape(F, S, 2)
        keep = (counts.sum(axis=2) > 0) | ~np.isfinite(gains)[None, :]
        n_seg = keep.sum(axis=1)
        n_tap = valid.sum(axis=1)
        block_len[chans] = 1 + 3 * n_seg + n_tap
        offsets = _pack_blocks(block_len)
        words = np.zeros(int(offsets[-1]), dtype=np.int32)
        base = offsets[chans]
        words[base] = n_seg
        f_i, s_i = np.nonzero(keep)
        rank = (np.cumsum(keep, axis=1) - 1)[f_i, s_i]
        at = base[f_i] + 1 + 3 * rank
        words[at] = counts[f_i, s_i, 0]
        words[at + 1] = counts[f_i, s_i, 1]
        words[at + 2] = gains.view(np.int32)[s_i]
        f_t, j_t = np.nonzero(valid)
        trank = (np.cumsum(valid, axis=1) - 1)[f_t, j_t]
        words[base[f_t] + 1 + 3 * n_seg[f_t] + trank] = idx_s[f_t, j_t]
        if valid.any():
            halo = int(idx_s[valid].max()) + 1
    else:
        block_len[chans] = 1  # filtered channel without impulses: S = 0, output is zero
        offsets = _pack_blocks(block_len)
        words = np.zeros(int(offsets[-1]), dtype=np.int32)
    return TapProgram(words, offsets.astype(np.int32), C, N.ORDER_SEGMENTED, 0 if identity else 1, halo, int(block_len.max()) if C else 0)


def ascending_program(fir: np.ndarray, frames: int) -> TapProgram:
    """Pack a dense FIR ``(m, channels)`` for ``convolve_velvet_noise`` (decorrelation.py:649-658):
    non-zeros in ascending index order with their coefficient.  A float32 FIR keeps float32
    arithmetic for float32 signals (VND_ORDER_ASCENDING); any other dtype makes numpy promote every
    step to float64 (VND_ORDER_ASCENDING_F64)."""
    fir = np.asarray(fir)
    if fir.ndim != 2:
        raise IndexError("too many indices for array: array is 1-dimensional, but 2 were indexed")
    order = N.ORDER_ASCENDING if fir.dtype == np.float32 else N.ORDER_ASCENDING_F64
    m, C = fir.shape
    ii, cc = np.nonzero(fir[: min(m, max(frames, 0))].T != 0)  # channel-major, index ascending
    ii, cc = cc, ii  # (index, channel) after the transpose trick above
    K = np.bincount(cc, minlength=C).astype(np.int64)
    block_len = 1 + 3 * K
    offsets = _pack_blocks(block_len)
    words = np.zeros(int(offsets[-1]), dtype=np.int32)
    words[offsets[:-1]] = K
    coef = fir[ii, cc].astype(np.float64)
    bits = coef.view(np.int64)
    rank = np.arange(len(cc)) - np.repeat(np.cumsum(K) - K, K)
    at = offsets[cc] + 1 + 3 * rank
    words[at] = ii
    words[at + 1] = (bits & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
    words[at + 2] = (bits >> 32).astype(np.int32)
    halo = int(ii.max()) + 1 if len(ii) else 0
    return TapProgram(words, offsets.astype(np.int32), C, order, 0, halo, int(block_len.max()) if C else 0)


def candidate_program(tables: Sequence[TapTable], envelope: Sequence[float], frames: int) -> TapProgram:
    """One SEGMENTED program whose "channels" are the channel-0 filters of many candidate tables —
    the input of the batched objective kernel (optimization.py:260-272 builds one VelvetNoise per
    candidate with ``filtered_channels=(0,)``)."""
    progs = [segmented_program(TapTable(t.fir_length_samples, t.num_impulses, t.num_segments, t.filtered[:1], t.index[:1], t.positive[:1], t.segment), envelope, frames) for t in tables]
    words = np.concatenate([p.words for p in progs]) if progs else np.zeros(0, np.int32)
    sizes = np.array([p.words.size for p in progs], dtype=np.int64)
    offsets = _pack_blocks(sizes)
    return TapProgram(
        np.ascontiguousarray(words, dtype=np.int32), offsets.astype(np.int32), len(progs), N.ORDER_SEGMENTED,
        progs[0].apply_gain if progs else 0, max((p.halo for p in progs), default=0), int(sizes.max()) if len(sizes) else 0,
    )


def kappa_family_program(kappas, *, sample_rate_hz: int, duration_seconds: float, num_impulses: int, envelope: Sequence[float], seed,
                         frames: int) -> TapProgram | None:
    """``candidate_program`` for the family ``optimize_velvet_noise`` sweeps (optimization.py:260-272): stereo
    ``VelvetNoise`` candidates with ``filtered_channels=(0,)`` that differ only in ``log_distribution_strength``.

    With one seed every candidate draws the SAME uniforms (decorrelation.py:488, :510-521), so the signs, the
    decay segments and with them the order of the taps in a program are common to the family; only the interval
    grid (``w``, ``starts``) depends on the strength.  The grid comes from ``_interval_grid_family`` (the single-table
    function's operations per element, verified against it); positions, rounding and packing are done for all
    candidates at once.  Returns None when the shortcut does not apply (a tap at or
    beyond ``frames``, an empty decay segment, no impulses): the caller then packs table by table."""
    kappas = [float(k) for k in kappas]
    K = len(kappas)
    S = len(envelope)
    if K == 0 or num_impulses <= 0 or S == 0:
        return None
    identity = tuple(envelope) == IDENTITY_ENVELOPE
    fir_length = int(round(sample_rate_hz * duration_seconds))
    sign_u, offs_u = _draw(seed, num_impulses, 1)  # one filtered channel: channel 0
    jitter = sample_rate_hz / (num_impulses / duration_seconds)
    w, starts = _interval_grid_family(kappas, num_impulses, fir_length)  # (K, N + 1) each
    pos = np.round(offs_u[:, 0][None, :] * np.fmax(0.0, w * jitter - 1) + starts).astype(np.int32)[:, :num_impulses]
    positive = np.round(sign_u[:, 0]) == 1.0     # (N,)
    segment = _segment_of_impulse(num_impulses, S)
    if int(pos.max()) >= frames or int(pos.min()) < 0:
        return None
    key = segment * 2 + positive                  # iteration order: segment, then negative before positive
    order = np.argsort(key, kind="stable")
    counts = np.zeros(2 * S, dtype=np.int64)
    np.add.at(counts, key, 1)
    counts = counts.reshape(S, 2)
    if (counts.sum(axis=1) == 0).any():
        return None
    gains = np.ones(S, dtype=np.float32) if identity else np.array([float(envelope[s]) for s in range(S)], dtype=np.float64).astype(np.float32)
    head = np.empty(1 + 3 * S, dtype=np.int32)
    head[0] = S
    head[1::3] = counts[:, 0]
    head[2::3] = counts[:, 1]
    head[3::3] = gains.view(np.int32)
    block = 1 + 3 * S + num_impulses
    words = np.empty((K, block), dtype=np.int32)
    words[:, : 1 + 3 * S] = head
    words[:, 1 + 3 * S:] = pos[:, order]
    offsets = (np.arange(K + 1, dtype=np.int64) * block).astype(np.int32)
    return TapProgram(np.ascontiguousarray(words.reshape(-1)), offsets, K, N.ORDER_SEGMENTED, 0 if identity else 1, int(pos.max()) + 1, block)
