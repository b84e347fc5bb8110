"""Round-2 golden fixtures, generated from the UNMODIFIED reference (build container only: reads /root/reference):

    python tests/golden/make_golden_r02.py [--skip-cfg5]

* ``tables_random.json`` - 240 random (sample rate, duration, impulses, strength, seed, outputs, segments) tuples with
  the sha256 of the reference's class-path tap rows and of ``generate_velvet_noise``'s dense FIR (or the exception the
  reference raises): pins the vectorised tap generation far beyond the three named configs.
* ``cfg5_clip0.json`` - BASELINE config 5 at its real shape for one clip: the reference's float32 scores of all 1024
  strengths on the 30 s synthetic clip 0, its argmin, its local-minima set and the refined strength
  ``optimize_velvet_noise`` returns (about 10 core-minutes; fanned out over the host cores, minimum by minimum).
* ``dsp_extra.npz`` / ``dsp_extra.json`` - ``rms_normalize`` in STEREO mode and on 1-D signals, ``peak_normalize`` and
  ``polar_coordinates`` cases (the helpers round 1 left out).

Nothing here is imported by the product.
"""

from __future__ import annotations

import hashlib
import io
import json
import multiprocessing as mp
import os
import sys
from contextlib import redirect_stdout

import numpy as np

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))

from vndecorrelate.decorrelation import VelvetNoise, generate_velvet_noise  # noqa: E402
from vndecorrelate.optimization import get_local_minima, symmetry_aware_objective  # noqa: E402
from vndecorrelate.utils.dsp import NormalizeMode, peak_normalize, polar_coordinates, rms_normalize  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OKW = dict(angle_limit=np.pi / 4, lambda_mean=5.0, lambda_skew=2.0, lambda_correlation=15.0, lambda_penalty=1e3)
FS, DUR, NIMP, GRID, FRAMES = 48000, 0.03, 30, 1024, 1_440_000


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def rows(vn: VelvetNoise) -> np.ndarray:
    out = []
    for ch, seq in enumerate(vn.velvet_noise):
        for si, seg in enumerate(seq):
            out += [(ch, si, int(i), -1) for i in seg[0]]
            out += [(ch, si, int(i), 1) for i in seg[1]]
    return np.array(out, dtype=np.int32).reshape(-1, 4)


def coloured_clip(index: int, frames: int) -> np.ndarray:
    from scipy.signal import lfilter

    g = np.random.default_rng(1000 + index)
    m = lfilter([0.02], [1, -0.98], g.standard_normal(frames))
    s = 0.3 * lfilter([0.02], [1, -0.98], g.standard_normal(frames))
    x = np.column_stack((m + s, m - s))
    return (x / np.max(np.abs(x)) * 0.5).astype(np.float32)


# ------------------------------------------------------------------------------------------------ random tap tables
def random_tables(n: int = 240):
    rng = np.random.default_rng(20261018)
    envelopes = [(1.0,), (0.9, -0.5), (0.85, 0.55, 0.35, 0.2), (1.0, 0.8, 0.6, 0.4, 0.2)]
    out = []
    while len(out) < n:
        fs = int(rng.choice([8000, 22050, 44100, 48000, 96000]))
        dur = float(np.round(rng.uniform(0.004, 0.3), 4))
        fir_len = int(round(fs * dur))
        n_max = int(min(400, 0.19 * fir_len))
        if n_max < 1:
            continue
        nimp = int(rng.integers(1, n_max + 1))
        kappa = float(rng.choice([0.0, 1.0, rng.uniform(0.0, 1.0)]))
        seed = int(rng.integers(0, 2**31 - 1))
        num_outs = int(rng.choice([1, 2, 3, 8]))
        n_filtered = int(rng.integers(1, num_outs + 1))
        env = envelopes[int(rng.integers(0, len(envelopes)))]
        params = dict(sample_rate_hz=fs, duration_seconds=dur, num_impulses=nimp, log_distribution_strength=kappa, seed=seed, num_outs=num_outs,
                      filtered_channels=list(range(n_filtered)), segment_envelope=list(env))
        rec = {"params": params}
        try:
            vn = VelvetNoise(sample_rate_hz=fs, duration_seconds=dur, num_impulses=nimp, log_distribution_strength=kappa, seed=seed, num_outs=num_outs,
                             filtered_channels=tuple(range(n_filtered)), segment_envelope=env, mode="LR", normalizer=None)
            r = rows(vn)
            rec["class_rows_sha256"] = sha(r)
            rec["class_rows_count"] = int(r.shape[0])
            rec["class_max_index"] = int(r[:, 2].max()) if len(r) else -1
        except Exception as exc:  # the reference's own error is the contract
            rec["class_error"] = type(exc).__name__
        try:
            fir = generate_velvet_noise(duration_seconds=dur, num_impulses=nimp, num_outs=num_outs, sample_rate_hz=fs, segment_envelope=env,
                                        log_distribution_strength=kappa, seed=seed)
            rec["dense_sha256"] = sha(fir)
            rec["dense_shape"] = list(fir.shape)
        except Exception as exc:
            rec["dense_error"] = type(exc).__name__
        out.append(rec)
    json.dump(out, open(os.path.join(HERE, "tables_random.json"), "w"), indent=0)
    print("tables_random.json:", len(out), "tuples,", sum("class_error" in r for r in out), "class errors,", sum("dense_error" in r for r in out), "dense errors")


# ------------------------------------------------------------------------------------------------ config 5, one clip
_CLIP = None


def _candidate(kappa):
    return VelvetNoise(sample_rate_hz=FS, duration_seconds=DUR, num_impulses=NIMP, log_distribution_strength=float(kappa), normalizer=None,
                       filtered_channels=(0,), mode="LR", seed=1)


def _score_chunk(kappas):
    global _CLIP
    if _CLIP is None:
        _CLIP = coloured_clip(0, FRAMES)
    return [float(quiet(symmetry_aware_objective, _CLIP, _candidate(k), **OKW)) for k in kappas]


def _refine_one(bounds):
    """scipy's bounded Brent on one interval, exactly as optimize_local_minima calls it (optimization.py:144-149)."""
    from scipy.optimize import minimize_scalar

    global _CLIP
    if _CLIP is None:
        _CLIP = coloured_clip(0, FRAMES)
    res = minimize_scalar(fun=lambda k: quiet(symmetry_aware_objective, _CLIP, _candidate(k), **OKW), bounds=bounds, method="bounded",
                          options={"xatol": 1e-4})
    return float(res.x), float(res.fun), int(res.nfev)


def cfg5_clip0():
    kappas = np.linspace(0.0, 1.0, GRID)
    workers = os.cpu_count() or 1
    with mp.get_context("fork").Pool(workers) as pool:
        chunks = np.array_split(kappas, workers * 8)
        scores = np.array([s for part in pool.map(_score_chunk, chunks) for s in part], dtype=np.float32)  # the objective is np.float32
        minima = get_local_minima(scores, GRID)
        bounds = [(float(kappas[max(0, i - 1)]), float(kappas[min(GRID - 1, i + 1)])) for i in minima]
        refined = pool.map(_refine_one, bounds, chunksize=4)
    best_x, best_f = 0.0, np.inf  # first strictly best (optimization.py:150-153)
    for x, f, _ in refined:
        if f < best_f:
            best_f, best_x = f, x
    clip = coloured_clip(0, FRAMES)
    rec = {"fs": FS, "duration_seconds": DUR, "num_impulses": NIMP, "grid_size": GRID, "frames": FRAMES, "clip_index": 0, "input_sha256": sha(clip),
           "scores": [float(s) for s in scores], "argmin": int(np.argmin(scores)), "local_minima": [int(i) for i in minima],
           "kappa": best_x, "best_score": float(best_f), "evaluations": int(GRID + sum(n for _, _, n in refined)),
           "refined": [[x, f] for x, f, _ in refined]}
    json.dump(rec, open(os.path.join(HERE, "cfg5_clip0.json"), "w"))
    print("cfg5_clip0.json: argmin", rec["argmin"], "minima", len(minima), "kappa", best_x, "evaluations", rec["evaluations"])


# ------------------------------------------------------------------------------------------------ dsp helpers
def dsp_inputs(kind: str, seed: int, n: int, dtype):
    """The seeded inputs of one dsp case (the tests regenerate them with this very function)."""
    rng = np.random.default_rng(seed)
    if kind == "polar_coordinates":
        return (rng.standard_normal(n) * 0.3).astype(dtype), (rng.standard_normal(n) * 0.3).astype(dtype)
    return (rng.standard_normal((n, 2)) * 0.3).astype(dtype), (rng.standard_normal((n, 2)) * 0.1).astype(dtype)


def dsp_extra():
    """Inputs are regenerated from seeds; outputs are stored as sha256 (what must match bit for bit) plus a strided
    sample of the values (what is compared with a tolerance), so the fixture stays small."""
    arrays, manifest = {}, []

    def add(kind, params, n, dtype, seed, outputs):
        k = len(manifest)
        rec = {"id": k, "kind": kind, "params": params, "n": n, "dtype": np.dtype(dtype).name, "seed": seed, "sha256": {}, "stride": max(1, n // 64)}
        for name, a in outputs.items():
            rec["sha256"][name] = sha(a)
            arrays[f"c{k}_{name}"] = np.ascontiguousarray(a[:: rec["stride"]])
        manifest.append(rec)

    seed = 5000
    for n in (1, 7, 100, 129, 1000, 4097, 65536, 100003):
        for dtype in (np.float32, np.float64):
            seed += 1
            x, y = dsp_inputs("rms_normalize", seed, n, dtype)
            for mode in (NormalizeMode.STEREO, NormalizeMode.DUAL_MONO):
                out = y.copy()
                rms_normalize(x, out, mode=mode)
                add("rms_normalize", {"mode": str(mode), "ndim": 2}, n, dtype, seed, {"y": out})
            out = y[:, 0].copy()
            rms_normalize(x[:, 0].copy(), out)
            add("rms_normalize", {"mode": "dual_mono", "ndim": 1}, n, dtype, seed, {"y": out})
            for mode in (NormalizeMode.STEREO, NormalizeMode.DUAL_MONO):
                out = y.copy()
                peak_normalize(out, mode=mode)
                add("peak_normalize", {"mode": str(mode), "ndim": 2}, n, dtype, seed, {"y": out})
            out = y[:, 1].copy()
            peak_normalize(out)
            add("peak_normalize", {"mode": "dual_mono", "ndim": 1}, n, dtype, seed, {"y": out})
    for n in (1, 5, 1000, 48000, 250007):
        for dtype in (np.float32, np.float64):
            seed += 1
            l, r = dsp_inputs("polar_coordinates", seed, n, dtype)
            for mode in ("MS", "LR"):
                for semi in (True, False):
                    for norm in (True, False):
                        rad, th, w = polar_coordinates(l, r, mode=mode, semicircular=semi, normalize=norm)
                        add("polar_coordinates", {"mode": mode, "semicircular": semi, "normalize": norm}, n, dtype, seed,
                            {"radii": rad, "thetas": th, "weights": w})
    np.savez_compressed(os.path.join(HERE, "dsp_extra.npz"), **arrays)
    json.dump(manifest, open(os.path.join(HERE, "dsp_extra.json"), "w"))
    print("dsp_extra:", len(manifest), "cases")


if __name__ == "__main__":
    random_tables()
    dsp_extra()
    if "--skip-cfg5" not in sys.argv:
        cfg5_clip0()
