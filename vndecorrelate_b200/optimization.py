"""Stereo-image objective and parameter sweeps with the reference's API
(``src/vndecorrelate/optimization.py``), evaluated by batched CUDA kernels.

``grid_scan`` recognises the two candidate families the reference's optimisers build —
``VelvetNoise(filtered_channels=(0,), mode='LR', normalizer=None, log_distribution_strength=k)``
(optimization.py:260-272) and ``HaasEffect(delay_time_seconds=t, mode='LR')`` (:183-190) — and
scores ALL candidates of a family in one kernel launch: the clip is read once per tile and every
candidate is evaluated from shared memory.  Any other decorrelator is run and its output scored by
the same kernels (one candidate).  The kernels return the sums the objective is made of; the scalar
combination below follows the reference's dtype chain (float32 for velvet-noise candidates, float64
for Haas candidates).

Refinement is the bounded Brent minimiser scipy runs for the reference (``minimize_scalar(method="bounded")``),
restated as a coroutine (``_bounded_brent``) so that every local minimum - of every clip - advances in lock-step:
one kernel launch per Brent iteration over all live minimisers instead of one per evaluation.

``optimize_velvet_noise_batch`` runs the whole optimisation for MANY clips (BASELINE config 5): clips are
partitioned over the ranks of ``torch.distributed`` (one process per GPU), the grid scores are all-gathered as one
float32 matrix (the only collective on the data path), and each rank refines the local minima of its own clips.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Sequence

import numpy as np

from . import _native as N
from . import runtime as R
from .decorrelation import Decorrelator, HaasEffect, VelvetNoise, _is_mode
from .taps import IDENTITY_ENVELOPE, TapProgram, candidate_program, kappa_family_program
from .utils.dsp import EPSILON, LayoutMode

__all__ = [
    "symmetry_aware_objective", "grid_scan", "get_local_minima", "optimize_local_minima", "optimize_haas_delay",
    "optimize_velvet_noise", "optimize_velvet_noise_batch", "vn_objective_partials", "haas_objective_partials", "vn_scores_from_partials", "haas_scores_from_partials",
]

_F32_HALF_PI = np.float32(np.pi / 2)
_F32_PI = np.float32(np.pi)


# ------------------------------------------------------------------------------------------------
# kernel calls
# ------------------------------------------------------------------------------------------------


def _planar_clips(clips, dtype=np.float32):
    """``(n_clips, 2, frames)`` contiguous array from a ``(frames, 2)`` signal, a list of such, or an
    already planar 3-D array."""
    if isinstance(clips, np.ndarray) and clips.ndim == 3:
        return np.ascontiguousarray(clips, dtype=dtype)
    if isinstance(clips, np.ndarray) and clips.ndim == 2:
        clips = [clips]
    if isinstance(clips, np.ndarray) and clips.ndim == 1:
        clips = [np.column_stack((clips, clips))]
    frames = clips[0].shape[0]
    out = np.empty((len(clips), 2, frames), dtype=dtype)
    for i, c in enumerate(clips):
        c = np.asarray(c)
        if c.ndim == 1:
            c = np.column_stack((c, c))
        if c.shape != (frames, 2):
            raise ValueError(f"all clips must be (frames, 2) with the same length, got {c.shape}")
        out[i] = c.T
    return out


def vn_objective_partials(clips, program: TapProgram):
    """Per (clip, candidate) partial sums of the objective for velvet-noise candidates.

    ``clips``: float32 ``(n_clips, 2, frames)`` numpy array or CUDA tensor (planar stereo).
    Returns float64 ``(n_clips, n_candidates, 12)`` (layout in ``include/vnd_b200.h``)."""
    lib = N.lib()
    if R.is_torch_tensor(clips):
        import torch

        if not clips.is_cuda or clips.dtype != torch.float32 or clips.dim() != 3 or clips.shape[1] != 2:
            raise ValueError("clips must be a CUDA float32 tensor of shape (n_clips, 2, frames)")
        clips = clips.contiguous()
        n_clips, _, frames = clips.shape
        out = torch.empty((n_clips, program.channels, N.OBJ_SLOTS), dtype=torch.float64, device=clips.device)
        nbytes = C.c_size_t()
        with torch.cuda.device(clips.device):
            N.check(lib.vnd_objective_workspace(frames, n_clips, program.channels, C.byref(nbytes)), "vnd_objective_workspace")
            work = torch.empty(nbytes.value, dtype=torch.uint8, device=clips.device)
            ps = R.device_program(program, clips.device)
            N.check(lib.vnd_vn_objective_batch_dev(clips.data_ptr(), frames, n_clips, 2 * frames, frames, C.byref(ps), out.data_ptr(),
                                                   work.data_ptr(), nbytes.value, R.torch_stream_ptr(clips)), "vnd_vn_objective_batch_dev")
        return out
    clips = _planar_clips(clips)
    n_clips, _, frames = clips.shape
    out = np.empty((n_clips, program.channels, N.OBJ_SLOTS), dtype=np.float64)
    ps = program.host_struct()
    N.check(lib.vnd_vn_objective_batch_host(R.HostContext.get().handle, clips.ctypes.data, frames, n_clips, 2 * frames, frames, C.byref(ps),
                                            out.ctypes.data), "vnd_vn_objective_batch_host")
    return out


def haas_objective_partials(clips, delays: Sequence[int]):
    """Per (clip, delay) partial sums for LR Haas candidates delaying channel 0 (float64
    arithmetic).  ``clips``: float32 or float64 ``(n_clips, 2, frames)``.  Returns float64
    ``(n_clips, n_delays, 8)``."""
    lib = N.lib()
    delays = np.ascontiguousarray(delays, dtype=np.int32)
    if R.is_torch_tensor(clips):
        import torch

        if not clips.is_cuda or clips.dtype not in (torch.float32, torch.float64) or clips.dim() != 3 or clips.shape[1] != 2:
            raise ValueError("clips must be a CUDA float32/float64 tensor of shape (n_clips, 2, frames)")
        clips = clips.contiguous()
        n_clips, _, frames = clips.shape
        out = torch.empty((n_clips, len(delays), N.HAAS_SLOTS), dtype=torch.float64, device=clips.device)
        with torch.cuda.device(clips.device):
            d = torch.from_numpy(delays).to(clips.device)
            N.check(lib.vnd_haas_objective_batch_dev(clips.data_ptr(), N.VND_F64 if clips.dtype == torch.float64 else N.VND_F32, frames, n_clips,
                                                     2 * frames, frames, d.data_ptr(), len(delays), out.data_ptr(), None, 0,
                                                     R.torch_stream_ptr(clips)), "vnd_haas_objective_batch_dev")
        return out
    dt = np.float64 if (isinstance(clips, np.ndarray) and clips.dtype == np.float64) else np.float32
    clips = _planar_clips(clips, dt)
    n_clips, _, frames = clips.shape
    out = np.empty((n_clips, len(delays), N.HAAS_SLOTS), dtype=np.float64)
    N.check(lib.vnd_haas_objective_batch_host(R.HostContext.get().handle, clips.ctypes.data, N.VND_F64 if dt == np.float64 else N.VND_F32, frames,
                                              n_clips, 2 * frames, frames, delays.ctypes.data, len(delays), out.ctypes.data),
            "vnd_haas_objective_batch_host")
    return out


# ------------------------------------------------------------------------------------------------
# scalar combination (optimization.py:71-105 and the helpers :11-43)
# ------------------------------------------------------------------------------------------------


def _max_abs_theta_f32(p: np.ndarray) -> np.float32:
    """max|theta| exactly as the reference's float32 pipeline rounds it, from the two tracked
    frames: a correctly rounded float32 arctan2 (via float64), then the fold of
    utils/dsp.py:405-410 in float32."""
    d_pos, s_pos, d_neg, s_neg = p[6], p[7], p[8], p[9]
    a = np.float32(np.arctan2(np.float64(d_pos), np.float64(s_pos)))
    best = a
    if d_neg > 0.0 or s_neg != 1.0:  # a frame with s < 0 was seen (the tracker starts at 0 / 1)
        t = np.float32(np.arctan2(np.float64(d_neg), -np.float64(s_neg)))
        if t > _F32_HALF_PI:
            t = np.float32(t - _F32_PI)
        best = max(best, np.float32(abs(t)))
    return np.float32(best)


def _max_abs_theta_rows(p: np.ndarray) -> np.ndarray:
    """``_max_abs_theta_f32`` for many rows (the array ``arctan2`` runs the same numpy loop as the scalar call; the
    CPU tests compare the two bit for bit)."""
    d_pos, s_pos, d_neg, s_neg = p[:, 6], p[:, 7], p[:, 8], p[:, 9]
    best = np.arctan2(d_pos, s_pos).astype(np.float32)
    seen = (d_neg > 0.0) | (s_neg != 1.0)  # a frame with s < 0 was seen (the tracker starts at 0 / 1)
    t = np.arctan2(d_neg, -s_neg).astype(np.float32)
    t = np.where(t > _F32_HALF_PI, (t - _F32_PI).astype(np.float32), t)
    cand = np.abs(t)
    return np.where(seen & (cand > best), cand, best).astype(np.float32)  # Python's max(best, cand): cand only if greater


def _vn_score(p: np.ndarray, angle_limit, lambda_mean, lambda_skew, lambda_correlation, lambda_penalty) -> np.float32:
    denom = np.float32(np.float32(p[0]) + np.float32(EPSILON))  # radii.sum() + EPSILON, float32 (utils/dsp.py:418)
    spread = float(np.float32(p[2] / np.float64(denom)))  # optimization.py:19-21
    cen = float(np.float32(p[1] / np.float64(denom)))  # :24-26
    m3 = float(np.float32(p[3] / np.float64(denom)))
    skew = m3 / (max(spread, EPSILON) ** 1.5)  # :29-38
    nl = np.float32(np.float32(np.sqrt(p[5])) + np.float32(EPSILON))  # :11-16, both channels over ||L||
    corr = np.float32(p[4] / (np.float64(nl) * np.float64(nl)))
    exceed = max(0.0, float(np.float32(_max_abs_theta_f32(p) - np.float32(angle_limit))))  # :41-43
    mean_theta_penalty = lambda_mean * cen**2
    skewness_penalty = lambda_skew * skew**2
    lr_correlation_penalty = lambda_correlation * corr**2  # np.float32
    constraint_penalty = lambda_penalty * exceed**2
    objective = spread - mean_theta_penalty - skewness_penalty - lr_correlation_penalty - constraint_penalty
    return -objective  # np.float32, like the reference for float32 decorrelator output


def _haas_score(p: np.ndarray, angle_limit, lambda_mean, lambda_skew, lambda_correlation, lambda_penalty) -> np.float64:
    denom = p[0] + EPSILON
    spread = float(p[2] / denom)
    cen = float(p[1] / denom)
    skew = float(p[3] / denom) / (max(spread, EPSILON) ** 1.5)
    nl = np.sqrt(p[6]) + EPSILON
    corr = np.float64(p[5] / (nl * nl))
    exceed = max(0.0, float(p[4] - angle_limit))
    objective = spread - lambda_mean * cen**2 - lambda_skew * skew**2 - lambda_correlation * corr**2 - lambda_penalty * exceed**2
    return np.float64(-objective)


def _vn_scores_rows(flat: np.ndarray, angle_limit, lambda_mean, lambda_skew, lambda_correlation, lambda_penalty) -> np.ndarray:
    """``_vn_score`` for many rows at once: the same operations in the same dtypes, element by element (IEEE basic
    operations give the same bits in an array as on scalars); every ``**`` stays the scalar call it is in the one-row
    chain, so that no vector math library can round them
    differently from the one-row chain.  ``tests/test_host_logic.py`` asserts bit equality with ``_vn_score``."""
    p = np.asarray(flat, dtype=np.float64)
    eps32 = np.float32(EPSILON)
    denom = (p[:, 0].astype(np.float32) + eps32).astype(np.float64)          # float32 sum, widened for the divisions
    spread = (p[:, 2] / denom).astype(np.float32).astype(np.float64)         # float(np.float32(...)) per row
    cen = (p[:, 1] / denom).astype(np.float32).astype(np.float64)
    m3 = (p[:, 3] / denom).astype(np.float32).astype(np.float64)
    skew = m3 / np.array([m ** 1.5 for m in np.maximum(spread, EPSILON).tolist()], dtype=np.float64)
    nl = (np.sqrt(p[:, 5]).astype(np.float32) + eps32).astype(np.float64)
    corr = (p[:, 4] / (nl * nl)).astype(np.float32)
    max_theta = _max_abs_theta_rows(p)
    exceed = np.maximum(0.0, (max_theta - np.float32(angle_limit)).astype(np.float64))
    # every power stays the scalar call of the one-row chain (float ** 2 is C pow, np.float32 ** 2 is powf: neither is
    # guaranteed to round like x * x, which is what an array power would compute)
    sq = lambda v: np.array([t ** 2 for t in v.tolist()], dtype=np.float64)  # noqa: E731  (Python floats)
    partial = spread - lambda_mean * sq(cen) - lambda_skew * sq(skew)       # Python-float arithmetic in the one-row chain
    corr2 = np.array([c ** 2 for c in corr], dtype=np.float32)               # np.float32 scalars
    lr_pen = (np.float32(lambda_correlation) * corr2).astype(np.float32)    # weak Python scalar times float32
    objective = partial.astype(np.float32) - lr_pen                          # float minus np.float32 -> np.float32
    objective = objective - (lambda_penalty * sq(exceed)).astype(np.float32)
    return -objective


def vn_scores_from_partials(partials, **kw) -> np.ndarray:
    p = partials.cpu().numpy() if R.is_torch_tensor(partials) else np.asarray(partials)
    flat = p.reshape(-1, N.OBJ_SLOTS)
    if len(flat) == 0:
        return np.zeros(p.shape[:-1], dtype=np.float32)
    return _vn_scores_rows(flat, **kw).astype(np.float32).reshape(p.shape[:-1])


def haas_scores_from_partials(partials, **kw) -> np.ndarray:
    p = partials.cpu().numpy() if R.is_torch_tensor(partials) else np.asarray(partials)
    flat = p.reshape(-1, N.HAAS_SLOTS)
    return np.array([_haas_score(row, **kw) for row in flat], dtype=np.float64).reshape(p.shape[:-1])


# ------------------------------------------------------------------------------------------------
# candidate recognition
# ------------------------------------------------------------------------------------------------


def _is_vn_candidate(d) -> bool:
    return (
        isinstance(d, VelvetNoise) and d.num_outs == 2 and d.width is None and d.normalizer is None
        and tuple(d.filtered_channels) == (0,) and _is_mode(d.mode, LayoutMode.LR)
    )


def _is_haas_candidate(d) -> bool:
    return isinstance(d, HaasEffect) and d.width is None and _is_mode(d.mode, LayoutMode.LR) and d.delayed_channel == 0


def _as_stereo_f32(input_signal):
    x = np.asarray(input_signal).astype(np.float32, copy=False)
    if x.ndim == 1:
        x = np.column_stack((x, x))
    if x.ndim != 2 or x.shape[1] < 2:
        raise ValueError(f"Input shape invalid: Expected shape (num samples, 2), but got shape {x.shape}.")
    return x[:, :2]


def _vn_family_program(decorrelators: Sequence[VelvetNoise], frames: int) -> TapProgram | None:
    env = decorrelators[0].segment_envelope
    if any(d.segment_envelope != env for d in decorrelators):
        return None
    return candidate_program([d.velvet_noise for d in decorrelators], env, frames)


_IDENTITY_PROGRAM_WORDS = np.array([1, 0, 1, np.float32(1.0).view(np.int32), 0], dtype=np.int32)  # one +1 tap at index 0


def _score_signal(y, kw) -> Any:
    """Objective of an already decorrelated stereo signal (generic decorrelators)."""
    y = y.detach().cpu().numpy() if R.is_torch_tensor(y) else np.asarray(y)
    if y.dtype == np.float32:
        prog = TapProgram(_IDENTITY_PROGRAM_WORDS, np.array([0, 5], dtype=np.int32), 1, N.ORDER_SEGMENTED, 0, 1, 5)
        return vn_scores_from_partials(vn_objective_partials(y[:, :2], prog), **kw)[0, 0]
    return haas_scores_from_partials(haas_objective_partials(y[:, :2].astype(np.float64, copy=False), [0]), **kw)[0, 0]


# ------------------------------------------------------------------------------------------------
# public API (optimization.py:46-310)
# ------------------------------------------------------------------------------------------------


def symmetry_aware_objective(input_signal, decorrelator: Decorrelator, *, angle_limit: float, lambda_mean: float, lambda_skew: float,
                             lambda_correlation: float, lambda_penalty: float):
    """Value to minimise: ``-(E_w[th^2] - l1*E_w[th]^2 - l2*skew^2 - l3*r^2 - lp*exceed^2)``
    for ``decorrelator`` applied to ``input_signal`` (optimization.py:46-105)."""
    kw = dict(angle_limit=angle_limit, lambda_mean=lambda_mean, lambda_skew=lambda_skew, lambda_correlation=lambda_correlation,
              lambda_penalty=lambda_penalty)
    return _scan(input_signal, [decorrelator], kw)[0]


def _scan(input_signal, decorrelators, kw) -> np.ndarray:
    if len(decorrelators) == 0:
        return np.array([])
    if len(input_signal) == 0:  # the reference takes max(|theta|) of the frames (optimization.py:41-43): numpy's own error
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    if all(_is_vn_candidate(d) for d in decorrelators):
        x = _as_stereo_f32(input_signal)
        prog = _vn_family_program(decorrelators, x.shape[0])
        if prog is not None:
            return vn_scores_from_partials(vn_objective_partials(x, prog), **kw)[0]
    if all(_is_haas_candidate(d) for d in decorrelators):
        x = _as_stereo_f32(input_signal)
        delays = [d.delay_len_samples for d in decorrelators]
        return haas_scores_from_partials(haas_objective_partials(x, delays), **kw)[0]
    return np.array([_score_signal(d.decorrelate(input_signal), kw) for d in decorrelators])


def grid_scan(input_signal, decorrelators: list[Decorrelator], **kwargs) -> np.ndarray:
    """Scores of all ``decorrelators`` on ``input_signal`` (optimization.py:108-117), one launch per
    candidate family."""
    print("Starting Grid Scan")
    return _scan(input_signal, list(decorrelators), kwargs)


def get_local_minima(scores, grid_size: int) -> list[int]:
    """Strict interior local minima, or ``[argmin]`` if there are none (optimization.py:120-128)."""
    local_minima = [i for i in range(1, grid_size - 1) if scores[i] < scores[i - 1] and scores[i] < scores[i + 1]]
    if not local_minima:
        return [int(np.argmin(scores))]
    return local_minima


def optimize_local_minima(local_minima: list[int], scalars, grid_size: int, scalar_objective: Callable[[float], float]):
    """Bounded Brent (``xatol=1e-4``) between the grid neighbours of each local minimum; the first
    strictly best result wins (optimization.py:131-155)."""
    from scipy.optimize import minimize_scalar

    best_scalar, best_score = 0.0, np.inf
    print("Starting Local Minima optimization")
    for i in local_minima:
        low = scalars[max(0, i - 1)]
        high = scalars[min(grid_size - 1, i + 1)]
        result = minimize_scalar(fun=scalar_objective, bounds=(low, high), method="bounded", options={"xatol": 1e-4})
        if result.fun < best_score:
            best_score = result.fun
            best_scalar = result.x
    return best_scalar


# ------------------------------------------------------------------------------------------------
# The two functions below (``_bounded_brent`` and, in array form, ``_BrentBatch``) restate
# ``scipy.optimize._optimize._minimize_scalar_bounded`` statement for statement, which SURVEY.md section 8(f).1 asks for so
# that every comparison of the refinement comes out as in the reference.  That routine is part of SciPy:
#
#   Copyright (c) 2001-2002 Enthought, Inc. 2003, SciPy Developers.  All rights reserved.
#
#   Redistribution and use in source and binary forms, with or without modification, are permitted provided that the
#   following conditions are met:
#   1. Redistributions of source code must retain the above copyright notice, this list of conditions and the following
#      disclaimer.
#   2. Redistributions in binary form must reproduce the above copyright notice, this list of conditions and the
#      following disclaimer in the documentation and/or other materials provided with the distribution.
#   3. Neither the name of the copyright holder nor the names of its contributors may be used to endorse or promote
#      products derived from this software without specific prior written permission.
#
#   THIS SOFTWARE IS PROVIDED BY THE COPYRIGHT HOLDERS AND CONTRIBUTORS "AS IS" AND ANY EXPRESS OR IMPLIED WARRANTIES,
#   INCLUDING, BUT NOT LIMITED TO, THE IMPLIED WARRANTIES OF MERCHANTABILITY AND FITNESS FOR A PARTICULAR PURPOSE ARE
#   DISCLAIMED. IN NO EVENT SHALL THE COPYRIGHT OWNER OR CONTRIBUTORS BE LIABLE FOR ANY DIRECT, INDIRECT, INCIDENTAL,
#   SPECIAL, EXEMPLARY, OR CONSEQUENTIAL DAMAGES (INCLUDING, BUT NOT LIMITED TO, PROCUREMENT OF SUBSTITUTE GOODS OR
#   SERVICES; LOSS OF USE, DATA, OR PROFITS; OR BUSINESS INTERRUPTION) HOWEVER CAUSED AND ON ANY THEORY OF LIABILITY,
#   WHETHER IN CONTRACT, STRICT LIABILITY, OR TORT (INCLUDING NEGLIGENCE OR OTHERWISE) ARISING IN ANY WAY OUT OF THE USE
#   OF THIS SOFTWARE, EVEN IF ADVISED OF THE POSSIBILITY OF SUCH DAMAGE.
#
# (BSD 3-Clause; the algorithm is Forsythe, Malcolm and Moler's ``fmin``.)
# ------------------------------------------------------------------------------------------------
def _bounded_brent(x1, x2, xatol: float, maxiter: int = 500):
    """Brent's bounded scalar minimiser as a coroutine: yields the next abscissa, is sent the function value, and
    returns an ``OptimizeResult`` when it has converged.

    This is the algorithm behind ``scipy.optimize.minimize_scalar(method="bounded")`` - the reference's refinement
    step (optimization.py:144-149) - restated statement for statement after ``_minimize_scalar_bounded`` of the
    pinned scipy (1.16 in the reference's lock file, 1.18 here; the routine is Forsythe, Malcolm and Moler's
    ``fmin``), with the same expressions on the same operand types, so that numpy's promotion rules give every
    intermediate the dtype it has there (the objective returns ``np.float32``) and every comparison comes out the
    same.  ``tests/test_host_logic.py`` pins it against scipy itself (abscissae, values, evaluation counts).
    Written as a coroutine so that hundreds of minimisers advance in lock-step without a thread each."""
    from math import sqrt

    from scipy.optimize import OptimizeResult

    if not (np.isfinite(x1) and np.isfinite(x2)):
        raise ValueError("Optimization bounds must be finite scalars.")
    if x1 > x2:
        raise ValueError("The lower bound exceeds the upper bound.")
    flag = 0
    sqrt_eps = sqrt(2.2e-16)
    golden_mean = 0.5 * (3.0 - sqrt(5.0))
    a, b = x1, x2
    fulc = a + golden_mean * (b - a)
    nfc, xf = fulc, fulc
    rat = e = 0.0
    x = xf
    fx = yield x
    num = 1
    fu = np.inf
    ffulc = fnfc = fx
    xm = 0.5 * (a + b)
    tol1 = sqrt_eps * np.abs(xf) + xatol / 3.0
    tol2 = 2.0 * tol1
    while np.abs(xf - xm) > (tol2 - 0.5 * (b - a)):
        golden = 1
        if np.abs(e) > tol1:  # try a parabolic step
            golden = 0
            r = (xf - nfc) * (fx - ffulc)
            q = (xf - fulc) * (fx - fnfc)
            p = (xf - fulc) * q - (xf - nfc) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = np.abs(q)
            r = e
            e = rat
            if (np.abs(p) < np.abs(0.5 * q * r)) and (p > q * (a - xf)) and (p < q * (b - xf)):
                rat = (p + 0.0) / q
                x = xf + rat
                if ((x - a) < tol2) or ((b - x) < tol2):
                    si = np.sign(xm - xf) + ((xm - xf) == 0)
                    rat = tol1 * si
            else:
                golden = 1
        if golden:  # golden-section step
            if xf >= xm:
                e = a - xf
            else:
                e = b - xf
            rat = golden_mean * e
        si = np.sign(rat) + (rat == 0)
        x = xf + si * np.maximum(np.abs(rat), tol1)
        fu = yield x
        num += 1
        if fu <= fx:
            if x >= xf:
                a = xf
            else:
                b = xf
            fulc, ffulc = nfc, fnfc
            nfc, fnfc = xf, fx
            xf, fx = x, fu
        else:
            if x < xf:
                a = x
            else:
                b = x
            if (fu <= fnfc) or (nfc == xf):
                fulc, ffulc = nfc, fnfc
                nfc, fnfc = x, fu
            elif (fu <= ffulc) or (fulc == xf) or (fulc == nfc):
                fulc, ffulc = x, fu
        xm = 0.5 * (a + b)
        tol1 = sqrt_eps * np.abs(xf) + xatol / 3.0
        tol2 = 2.0 * tol1
        if num >= maxiter:
            flag = 1
            break
    if np.isnan(xf) or np.isnan(fx) or np.isnan(fu):
        flag = 2
    return OptimizeResult(fun=np.asarray(fx)[()], status=flag, success=(flag == 0), x=xf, nfev=num, nit=num)  # numpy scalar, as minimize_scalar returns it


def lockstep_minimize(bounds: Sequence[tuple[float, float]], batch_objective: Callable[..., Sequence[Any]], *, xatol: float = 1e-4,
                      with_ids: bool = False):
    """Run the bounded Brent minimiser on many intervals at once, in lock-step.

    Every interval gets its own minimiser (``_bounded_brent``: the reference's refinement step,
    optimization.py:144-149, comparison for comparison).  In every round the pending abscissae of all live
    minimisers are evaluated with ONE call of ``batch_objective`` (one kernel launch over all of them) and handed
    back.  Returns the ``OptimizeResult`` list in input order.  The sequence of abscissae each minimiser sees
    depends only on its own function values, so the results equal those of the one-at-a-time loop whenever
    ``batch_objective`` returns the values the scalar objective would.  With ``with_ids`` the callback is called as
    ``batch_objective(abscissae, ids)`` (``ids`` = positions in ``bounds``), which is how the multi-clip optimiser
    knows the clip each abscissa belongs to."""
    n = len(bounds)
    results: list[Any] = [None] * n
    live: dict[int, Any] = {}
    pending: dict[int, Any] = {}
    for i, (lo, hi) in enumerate(bounds):
        gen = _bounded_brent(lo, hi, xatol)
        live[i] = gen
        pending[i] = next(gen)  # the first abscissa (a minimiser always evaluates at least once)
    while live:
        ids = sorted(live)
        xs = [float(pending[i]) for i in ids]
        vals = batch_objective(xs, ids) if with_ids else batch_objective(xs)
        for i, v in zip(ids, vals):
            try:
                pending[i] = live[i].send(v)
            except StopIteration as done:
                results[i] = done.value
                del live[i], pending[i]
    return results


class _BrentBatch:
    """``_bounded_brent`` for MANY intervals at once, as array arithmetic: one state vector per variable of scipy's
    ``_minimize_scalar_bounded``, every statement of its loop applied under a mask to the minimisers that take it.

    The element-wise operations (+, -, *, /, abs, sign, comparisons, maximum) are IEEE basic operations, identical in
    an array and on scalars, and the dtypes are the scalar code's: abscissae and step sizes float64, function values
    the objective's own dtype (float32 for velvet-noise candidates), so that ``(xf - nfc) * (fx - ffulc)`` is a float64
    product of a float64 and an exactly widened float32 difference, as under numpy's scalar promotion.  Every minimiser
    therefore sees the abscissae, takes the branches and returns the values of the coroutine (and so of scipy);
    ``tests/test_host_logic.py`` compares them abscissa by abscissa on thousands of float32 objectives.  A Brent round
    over 20 000 minimisers costs about a millisecond here instead of 0.1 s of Python generator switches."""

    def __init__(self, lows, highs, xatol: float, maxiter: int = 500):
        from math import sqrt

        a = np.asarray(lows, dtype=np.float64).copy()
        b = np.asarray(highs, dtype=np.float64).copy()
        if not (np.all(np.isfinite(a)) and np.all(np.isfinite(b))):
            raise ValueError("Optimization bounds must be finite scalars.")
        if np.any(a > b):
            raise ValueError("The lower bound exceeds the upper bound.")
        self.n = len(a)
        self.xatol, self.maxiter = xatol, maxiter
        self.sqrt_eps = sqrt(2.2e-16)
        self.golden_mean = 0.5 * (3.0 - sqrt(5.0))
        self.a, self.b = a, b
        self.fulc = a + self.golden_mean * (b - a)
        self.nfc = self.fulc.copy()
        self.xf = self.fulc.copy()
        self.rat = np.zeros(self.n)
        self.e = np.zeros(self.n)
        self.x = self.xf.copy()
        self.num = np.zeros(self.n, dtype=np.int64)
        self.flag = np.zeros(self.n, dtype=np.int64)
        self.live = np.ones(self.n, dtype=bool)
        self.started = False
        self.fx = self.ffulc = self.fnfc = self.fu = None

    def pending(self):
        """Indices of the live minimisers and the abscissa each wants evaluated."""
        idx = np.nonzero(self.live)[0]
        return idx, self.x[idx]

    def _propose(self, m):
        """The body of scipy's ``while`` loop up to the evaluation, for the minimisers in mask ``m`` (all of which
        passed the loop condition)."""
        a, b, xf, nfc, fulc, fx, ffulc, fnfc, xm, tol1, tol2 = (self.a, self.b, self.xf, self.nfc, self.fulc, self.fx, self.ffulc, self.fnfc, self.xm,
                                                                self.tol1, self.tol2)
        golden = m.copy()
        para = m & (np.abs(self.e) > tol1)
        if para.any():
            i = np.nonzero(para)[0]
            r = (xf[i] - nfc[i]) * (fx[i] - ffulc[i])
            q = (xf[i] - fulc[i]) * (fx[i] - fnfc[i])
            p = (xf[i] - fulc[i]) * q - (xf[i] - nfc[i]) * r
            q = 2.0 * (q - r)
            p = np.where(q > 0.0, -p, p)
            q = np.abs(q)
            r = self.e[i]
            self.e[i] = self.rat[i]
            ok = (np.abs(p) < np.abs(0.5 * q * r)) & (p > q * (a[i] - xf[i])) & (p < q * (b[i] - xf[i]))
            golden[i] = ~ok
            if ok.any():
                j = i[ok]
                with np.errstate(divide="ignore", invalid="ignore"):
                    rat = (p[ok] + 0.0) / q[ok]
                x = xf[j] + rat
                near = ((x - a[j]) < tol2[j]) | ((b[j] - x) < tol2[j])
                d = xm[j] - xf[j]
                si = np.sign(d) + (d == 0)
                self.rat[j] = np.where(near, tol1[j] * si, rat)
        if golden.any():
            i = np.nonzero(golden)[0]
            self.e[i] = np.where(xf[i] >= xm[i], a[i] - xf[i], b[i] - xf[i])
            self.rat[i] = self.golden_mean * self.e[i]
        i = np.nonzero(m)[0]
        si = np.sign(self.rat[i]) + (self.rat[i] == 0)
        self.x[i] = xf[i] + si * np.maximum(np.abs(self.rat[i]), tol1[i])

    def _close_or_propose(self, m):
        """Loop condition for the minimisers in ``m``: the converged ones finish, the others get their next abscissa."""
        go = m & (np.abs(self.xf - self.xm) > (self.tol2 - 0.5 * (self.b - self.a)))
        self.live[m & ~go] = False
        if go.any():
            self._propose(go)

    def feed(self, idx, values):
        """Function values (array in the objective's dtype) for the abscissae handed out by ``pending``."""
        values = np.asarray(values)
        if not self.started:  # the first evaluation of every minimiser (scipy evaluates before its loop)
            assert len(idx) == self.n
            self.started = True
            self.fx = values.copy()
            self.ffulc = values.copy()
            self.fnfc = values.copy()
            self.fu = np.full(self.n, np.inf, dtype=values.dtype)
            self.num[:] = 1
            self.xm = 0.5 * (self.a + self.b)
            self.tol1 = self.sqrt_eps * np.abs(self.xf) + self.xatol / 3.0
            self.tol2 = 2.0 * self.tol1
            self._close_or_propose(self.live.copy())
            return
        a, b, x, xf = self.a, self.b, self.x, self.xf
        fu = values
        self.fu[idx] = fu
        self.num[idx] += 1
        better = fu <= self.fx[idx]
        ib, iw = idx[better], idx[~better]
        if len(ib):  # fu <= fx
            right = x[ib] >= xf[ib]
            a[ib[right]] = xf[ib[right]]
            b[ib[~right]] = xf[ib[~right]]
            self.fulc[ib], self.ffulc[ib] = self.nfc[ib], self.fnfc[ib]
            self.nfc[ib], self.fnfc[ib] = xf[ib], self.fx[ib]
            self.xf[ib], self.fx[ib] = x[ib], fu[better]
        if len(iw):
            fw = fu[~better]
            left = x[iw] < xf[iw]
            a[iw[left]] = x[iw[left]]
            b[iw[~left]] = x[iw[~left]]
            c1 = (fw <= self.fnfc[iw]) | (self.nfc[iw] == xf[iw])
            c2 = ~c1 & ((fw <= self.ffulc[iw]) | (self.fulc[iw] == xf[iw]) | (self.fulc[iw] == self.nfc[iw]))
            i1, i2 = iw[c1], iw[c2]
            self.fulc[i1], self.ffulc[i1] = self.nfc[i1], self.fnfc[i1]
            self.nfc[i1], self.fnfc[i1] = x[i1], fw[c1]
            self.fulc[i2], self.ffulc[i2] = x[i2], fw[c2]
        self.xm[idx] = 0.5 * (a[idx] + b[idx])
        self.tol1[idx] = self.sqrt_eps * np.abs(self.xf[idx]) + self.xatol / 3.0
        self.tol2[idx] = 2.0 * self.tol1[idx]
        m = np.zeros(self.n, dtype=bool)
        m[idx] = True
        over = m & (self.num >= self.maxiter)
        self.flag[over] = 1
        self.live[over] = False
        self._close_or_propose(m & ~over)

    def results(self):
        """(x, fun, nfev) arrays; ``fun`` in the objective's dtype, like ``OptimizeResult.fun``."""
        return self.xf.copy(), self.fx.copy(), self.num.copy()


def lockstep_minimize_arrays(lows, highs, batch_objective: Callable[..., Any], *, xatol: float = 1e-4):
    """``lockstep_minimize`` on the array engine: ``batch_objective(abscissae, ids)`` (numpy arrays) returns the values
    as an array in the objective's dtype.  Returns ``(x, fun, nfev)`` arrays in input order."""
    eng = _BrentBatch(lows, highs, xatol)
    while True:
        idx, xs = eng.pending()
        if len(idx) == 0:
            break
        eng.feed(idx, batch_objective(xs, idx))
    return eng.results()


def optimize_local_minima_batched(local_minima: list[int], scalars, grid_size: int, batch_objective: Callable[[list[float]], Sequence[Any]]):
    """``optimize_local_minima`` with all minima refined in lock-step (one launch per Brent iteration
    instead of one per evaluation); same bounds, same ``xatol``, same first-strictly-best rule
    (optimization.py:131-155).  ``batch_objective(list of abscissae)`` returns their values."""
    print("Starting Local Minima optimization")
    lows = [scalars[max(0, i - 1)] for i in local_minima]
    highs = [scalars[min(grid_size - 1, i + 1)] for i in local_minima]
    xs, funs, _ = lockstep_minimize_arrays(lows, highs, lambda x, ids: np.asarray(batch_objective([float(v) for v in x])), xatol=1e-4)
    best_scalar, best_score = 0.0, np.inf
    for x, fun in zip(xs, funs):
        if fun < best_score:
            best_score = fun
            best_scalar = x
    return best_scalar


def _device_clip(x: np.ndarray):
    """The clip as a resident planar CUDA tensor (uploaded once per optimisation) when torch is there;
    otherwise the numpy array itself (the host entry points then upload it on every call)."""
    try:
        import torch

        if torch.cuda.is_available():
            return torch.from_numpy(np.ascontiguousarray(x.T)[None]).to(f"cuda:{R.default_device()}")
    except ImportError:
        pass
    return x


def optimize_haas_delay(*, input_signal, sample_rate_hz: int, max_delay_seconds: int, grid_size: int = 400, angle_limit: float = np.pi / 4,
                        lambda_mean: float = 5.0, lambda_skew: float = 2.0, lambda_correlation: float = 15.0, lambda_penalty: float = 1e3) -> float:
    """Optimised ``delay_time_seconds`` in ``[0, max_delay_seconds]`` (optimization.py:158-227)."""
    kw = dict(angle_limit=angle_limit, lambda_mean=lambda_mean, lambda_skew=lambda_skew, lambda_correlation=lambda_correlation,
              lambda_penalty=lambda_penalty)
    taus = np.linspace(0.0, max_delay_seconds, grid_size)
    candidates = [HaasEffect(sample_rate_hz=sample_rate_hz, delay_time_seconds=tau, mode="LR") for tau in taus]
    scores = grid_scan(input_signal, candidates, **kw)
    local_minima = get_local_minima(scores, grid_size)
    clip = _device_clip(_as_stereo_f32(input_signal))

    def batch(ts):
        delays = [HaasEffect(sample_rate_hz=sample_rate_hz, delay_time_seconds=t, mode="LR").delay_len_samples for t in ts]
        p = haas_objective_partials(clip, delays)
        p = p.cpu().numpy() if R.is_torch_tensor(p) else p
        return list(haas_scores_from_partials(p, **kw)[0])

    return optimize_local_minima_batched(local_minima, taus, grid_size, batch)


def optimize_velvet_noise(*, input_signal, sample_rate_hz: int, duration_seconds: float, num_impulses: int, seed: int = 1, grid_size: int = 400,
                          angle_limit: float = np.pi / 4, lambda_mean: float = 5.0, lambda_skew: float = 2.0, lambda_correlation: float = 15.0,
                          lambda_penalty: float = 1e3) -> float:
    """Optimised ``log_distribution_strength`` in ``[0, 1]`` (optimization.py:230-310)."""
    kw = dict(angle_limit=angle_limit, lambda_mean=lambda_mean, lambda_skew=lambda_skew, lambda_correlation=lambda_correlation,
              lambda_penalty=lambda_penalty)

    def candidate(kappa):
        return VelvetNoise(sample_rate_hz=sample_rate_hz, duration_seconds=duration_seconds, num_impulses=num_impulses,
                           log_distribution_strength=kappa, normalizer=None, filtered_channels=(0,), mode="LR", seed=seed)

    kappas = np.linspace(0.0, 1.0, grid_size)
    x = _as_stereo_f32(input_signal)
    clip = _device_clip(x)
    envelope = candidate(0.0).segment_envelope

    def family_program(ks):
        """The candidates of the sweep differ only in the strength and share their random draws: their tap
        programs are packed in one vectorised pass (taps.kappa_family_program), table by table only when
        that shortcut does not apply."""
        prog = kappa_family_program(ks, sample_rate_hz=sample_rate_hz, duration_seconds=duration_seconds, num_impulses=num_impulses,
                                    envelope=envelope, seed=seed, frames=x.shape[0]) if seed is not None else None
        return prog if prog is not None else _vn_family_program([candidate(k) for k in ks], x.shape[0])

    def batch(ks):
        p = vn_objective_partials(clip, family_program(ks))
        p = p.cpu().numpy() if R.is_torch_tensor(p) else p
        return list(vn_scores_from_partials(p, **kw)[0])

    print("Starting Grid Scan")  # grid_scan's message (optimization.py:112)
    scores = np.asarray(batch(kappas))
    local_minima = get_local_minima(scores, grid_size)

    return optimize_local_minima_batched(local_minima, kappas, grid_size, batch)


# ------------------------------------------------------------------------------------------------
# many clips at once (BASELINE config 5): clips sharded over ranks, one all-gather of the scores
# ------------------------------------------------------------------------------------------------


class _ClipBank:
    """The rank's clips resident on its GPU as one planar ``(n, 2, frames)`` float32 tensor, and the evaluation of
    ragged (clip, strengths) requests on them: ONE tap program for all requested strengths, uploaded once, one
    launch per clip on a sub-range of its candidates, one download of all partial sums, one vectorised scoring pass."""

    def __init__(self, clips, family: Callable[[Sequence[float]], TapProgram], kw: dict):
        import torch

        if R.is_torch_tensor(clips):
            if not clips.is_cuda or clips.dtype != torch.float32 or clips.dim() != 3 or clips.shape[1] != 2:
                raise ValueError("clips must be a CUDA float32 tensor of shape (n_clips, 2, frames)")
            self.clips = clips.contiguous()
        else:
            host = _planar_clips(clips)
            self.clips = torch.from_numpy(host).to(f"cuda:{R.default_device()}")
        self.n, _, self.frames = self.clips.shape
        self.device = self.clips.device
        self.family = family
        self.kw = kw
        self.evaluations = 0
        self.launches = 0
        self._work = None

    def _workspace(self, n_cand: int):
        import torch

        nbytes = C.c_size_t()
        N.check(N.lib().vnd_objective_workspace(self.frames, 1, n_cand, C.byref(nbytes)), "vnd_objective_workspace")
        if self._work is None or self._work.numel() < nbytes.value:
            self._work = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        return self._work, nbytes.value

    def partials(self, requests: Sequence[Sequence[float]], clip_ids: Sequence[int] | None = None):
        """``requests[i]`` = strengths to evaluate on clip ``clip_ids[i]`` (default: clip ``i``); returns the float64
        partial sums ``(total, OBJ_SLOTS)`` in request order (a CUDA tensor) and the per-request counts.  Everything is
        enqueued on torch's current stream; nothing here waits for the device."""
        import torch

        if clip_ids is None:
            clip_ids = range(len(requests))
        counts = [len(r) for r in requests]
        total = sum(counts)
        out = torch.empty((total, N.OBJ_SLOTS), dtype=torch.float64, device=self.device)
        if total == 0:
            return out, counts
        flat = [float(k) for r in requests for k in r]
        shared = (len(requests) == self.n and len(set(counts)) == 1 and list(clip_ids) == list(range(self.n))
                  and all(list(r) == list(requests[0]) for r in requests[1:]))
        lib = N.lib()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if shared:  # the grid stage: every clip scores the same strengths -> one launch over clips x candidates
                prog = self.family(flat[: counts[0]])
                ps = R.device_program(prog, self.device)
                nbytes = C.c_size_t()
                N.check(lib.vnd_objective_workspace(self.frames, self.n, prog.channels, C.byref(nbytes)), "vnd_objective_workspace")
                work = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
                N.check(lib.vnd_vn_objective_batch_dev(self.clips.data_ptr(), self.frames, self.n, 2 * self.frames, self.frames, C.byref(ps),
                                                       out.data_ptr(), work.data_ptr(), nbytes.value, stream), "vnd_vn_objective_batch_dev")
                self.launches += 1
            else:
                prog = self.family(flat)
                ps = R.device_program(prog, self.device)
                work, wbytes = self._workspace(max(counts))
                base = 0
                for i, n_i in zip(clip_ids, counts):
                    if n_i == 0:
                        continue
                    sub = N.TapProgramStruct(ps.words, ps.offsets + 4 * base, ps.n_words, n_i, ps.order, ps.apply_gain, ps.halo, ps.max_channel_words)
                    N.check(lib.vnd_vn_objective_batch_dev(self.clips.data_ptr() + 4 * i * 2 * self.frames, self.frames, 1, 2 * self.frames, self.frames,
                                                           C.byref(sub), out.data_ptr() + 8 * N.OBJ_SLOTS * base, work.data_ptr(), wbytes, stream),
                            "vnd_vn_objective_batch_dev")
                    self.launches += 1
                    base += n_i
        self.evaluations += total
        return out, counts

    def scores(self, requests: Sequence[Sequence[float]]) -> list[np.ndarray]:
        p, counts = self.partials(requests)
        flat = vn_scores_from_partials(p.cpu().numpy(), **self.kw) if p.shape[0] else np.zeros(0, dtype=np.float32)
        out, base = [], 0
        for n_i in counts:
            out.append(flat[base: base + n_i])
            base += n_i
        return out

    def submit(self, requests: Sequence[Sequence[float]], clip_ids: Sequence[int]):
        """Enqueue the evaluation of ``requests`` on the clips ``clip_ids`` and the download of its partial sums into
        page-locked memory; returns a handle for ``collect``.  The host is free until then: the refinement keeps two such
        batches in flight, so that scoring and the Brent step of one run under the kernels of the other."""
        import torch

        p, _ = self.partials(requests, clip_ids)
        host = torch.empty(p.shape, dtype=p.dtype, pin_memory=True)
        host.copy_(p, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        return host, done, p  # p is kept alive until the copy has run

    def collect(self, handle) -> np.ndarray:
        """Scores (flat float32 array, request order) of a submitted batch."""
        host, done, _ = handle
        done.synchronize()
        return vn_scores_from_partials(host.numpy(), **self.kw) if host.shape[0] else np.zeros(0, dtype=np.float32)


def optimize_velvet_noise_batch(*, input_signals=None, sample_rate_hz: int, duration_seconds: float, num_impulses: int, seed: int = 1,
                                grid_size: int = 400, angle_limit: float = np.pi / 4, lambda_mean: float = 5.0, lambda_skew: float = 2.0,
                                lambda_correlation: float = 15.0, lambda_penalty: float = 1e3, group=None, details: bool = False,
                                local_signals=None, total_clips: int | None = None, _bank_factory=None):
    """``optimize_velvet_noise`` (optimization.py:230-310) for MANY clips of equal length: returns the float64 array of
    optimised ``log_distribution_strength`` values, element ``i`` equal to what ``optimize_velvet_noise(input_signal=
    input_signals[i], ...)`` returns.

    ``input_signals``: a sequence of ``(frames, 2)`` arrays, a planar ``(n_clips, 2, frames)`` float32 array, or a CUDA
    tensor of that shape.  With ``torch.distributed`` initialised (one process per GPU) the clips are partitioned in
    contiguous blocks over the ranks of ``group`` (``sharding.block_range``); every rank passes the SAME
    ``input_signals`` (only its block is read and uploaded) - or, when the clips already live on the ranks, its own
    block as ``local_signals`` (e.g. a resident CUDA tensor) together with ``total_clips``.  Data path per rank: grid scores of its clips (one launch over clips x grid) -> ONE all-gather of the
    float32 score matrix so that every rank holds all rows (``get_local_minima`` needs both neighbours of every grid
    point, optimization.py:120-128) -> lock-step Brent refinement of the local minima of its own clips -> one
    all-gather of the refined strengths.  ``details=True`` also returns the score matrix, the local-minima sets and
    evaluation counts."""
    from . import sharding as S

    kw = dict(angle_limit=angle_limit, lambda_mean=lambda_mean, lambda_skew=lambda_skew, lambda_correlation=lambda_correlation,
              lambda_penalty=lambda_penalty)
    rank, world = S._world(group)
    if local_signals is not None:
        if total_clips is None:
            raise ValueError("local_signals needs total_clips (the number of clips over all ranks)")
        n_clips = int(total_clips)
        lo, hi = S.block_range(n_clips, rank, world)
        local = local_signals
        n_local = local.shape[0] if (R.is_torch_tensor(local) or (isinstance(local, np.ndarray) and local.ndim == 3)) else len(local)
        if n_local != hi - lo:
            raise ValueError(f"rank {rank} of {world} owns clips [{lo}, {hi}) but local_signals holds {n_local}")
    else:
        if input_signals is None:
            raise ValueError("pass input_signals (all clips) or local_signals (this rank's block)")
        n_clips = input_signals.shape[0] if (R.is_torch_tensor(input_signals) or (isinstance(input_signals, np.ndarray) and input_signals.ndim == 3)) \
            else len(input_signals)
        lo, hi = S.block_range(n_clips, rank, world)
        local = input_signals[lo:hi]

    def candidate(kappa):
        return VelvetNoise(sample_rate_hz=sample_rate_hz, duration_seconds=duration_seconds, num_impulses=num_impulses,
                           log_distribution_strength=kappa, normalizer=None, filtered_channels=(0,), mode="LR", seed=seed)

    envelope = candidate(0.0).segment_envelope
    frames_box: list[int] = []

    def family(ks):
        frames = frames_box[0]
        prog = kappa_family_program(ks, sample_rate_hz=sample_rate_hz, duration_seconds=duration_seconds, num_impulses=num_impulses,
                                    envelope=envelope, seed=seed, frames=frames) if seed is not None else None
        return prog if prog is not None else _vn_family_program([candidate(k) for k in ks], frames)

    kappas = np.linspace(0.0, 1.0, grid_size)
    if hi > lo:
        bank = (_bank_factory or _ClipBank)(local, family, kw)
        frames_box.append(bank.frames)
        print("Starting Grid Scan")
        local_scores = np.stack(bank.scores([kappas] * (hi - lo))).astype(np.float32)
    else:
        bank = None
        local_scores = np.zeros((0, grid_size), dtype=np.float32)

    counts = [S.block_range(n_clips, r, world)[1] - S.block_range(n_clips, r, world)[0] for r in range(world)]
    device = bank.device if (bank is not None and getattr(bank, "device", None) is not None) else S.collective_device(group)
    scores = S.all_gather_rows(local_scores, counts, device=device, group=group)  # every rank: (n_clips, grid)

    minima = [get_local_minima(row, grid_size) for row in scores]
    best = np.zeros(hi - lo, dtype=np.float64)
    if hi > lo:
        print("Starting Local Minima optimization")
        lows, highs, owner = [], [], []
        for ci in range(lo, hi):
            for i in minima[ci]:
                lows.append(kappas[max(0, i - 1)])
                highs.append(kappas[min(grid_size - 1, i + 1)])
                owner.append(ci - lo)
        owner = np.asarray(owner, dtype=np.int64)

        def requests_of(xs, ids, clips_of_lane):  # ids ascend and the minimisers are listed clip by clip: a clip's requests are one contiguous run
            own = owner[ids]
            cuts = np.searchsorted(own, np.asarray(list(clips_of_lane) + [clips_of_lane[-1] + 1]))
            return [xs[cuts[j]: cuts[j + 1]] for j in range(len(clips_of_lane))]

        n_local = hi - lo
        if hasattr(bank, "submit") and n_local >= 2:
            # Two half-batches of clips in flight: while the kernels of one half run, the host scores the other half,
            # advances its minimisers and packs their next programs.  A minimiser only ever sees its own values, so
            # the results do not depend on the split.
            halves = [list(range(0, n_local // 2)), list(range(n_local // 2, n_local))]
            lows, highs = np.asarray(lows, dtype=np.float64), np.asarray(highs, dtype=np.float64)
            members = [np.nonzero(np.isin(owner, h))[0] for h in halves]
            engines = [_BrentBatch(lows[m], highs[m], 1e-4) for m in members]
            handles: list[Any] = [None, None]

            def launch(h):
                idx, xs_h = engines[h].pending()
                if len(idx) == 0:
                    handles[h] = None
                    return
                handles[h] = (idx, bank.submit(requests_of(xs_h, members[h][idx], halves[h]), halves[h]))

            for h in (0, 1):
                launch(h)
            while handles[0] is not None or handles[1] is not None:
                for h in (0, 1):
                    if handles[h] is None:
                        continue
                    idx, handle = handles[h]
                    engines[h].feed(idx, bank.collect(handle))
                    launch(h)
            xs = np.empty(len(owner), dtype=np.float64)
            funs = np.empty(len(owner), dtype=np.float32)
            for h in (0, 1):
                x_h, f_h, _ = engines[h].results()
                xs[members[h]] = x_h
                funs[members[h]] = f_h
        else:
            def batch(xs, ids):
                vals = bank.scores(requests_of(xs, ids, list(range(n_local))))
                return np.concatenate(vals) if len(vals) else np.zeros(0, dtype=np.float32)

            xs, funs, _ = lockstep_minimize_arrays(lows, highs, batch, xatol=1e-4)
        best_score = [np.inf] * (hi - lo)
        for x, fun, c in zip(xs, funs, owner):  # first strictly best minimum of each clip (optimization.py:150-153)
            if fun < best_score[c]:
                best_score[c] = fun
                best[c] = x
    refined = S.all_gather_rows(best.reshape(-1, 1), counts, device=device, group=group).reshape(-1)
    if details:
        info = {"scores": scores, "local_minima": minima, "argmin": [int(np.argmin(r)) for r in scores],
                "evaluations_local": bank.evaluations if bank is not None else 0, "launches_local": bank.launches if bank is not None else 0,
                "clips_local": (lo, hi)}
        return refined, info
    return refined
