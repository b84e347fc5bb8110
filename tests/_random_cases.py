"""Seeded random ``VelvetNoise.decorrelate`` cases shared by the golden generator (reference), the CPU test (oracle vs the
reference's hashes) and the GPU test (library vs oracle): same generator, same seed, same order."""

from __future__ import annotations

import numpy as np

SEED = 20261019
COUNT = 120


def random_vn_case(rng):
    fs = int(rng.choice([8000, 16000, 22050, 44100, 48000, 96000]))
    dur = float(rng.uniform(0.002, 0.06))
    n_imp = int(rng.integers(1, 49))
    kappa = float(rng.choice([0.0, 0.25, 0.5, 1.0, rng.uniform(0, 1)]))
    n_env = int(rng.integers(1, 6))
    env = tuple(float(v) for v in rng.choice([1.0, 0.85, 0.55, 0.35, 0.2, -0.5, 0.0, 2.0], size=n_env))
    mode = str(rng.choice(["MS", "LR"]))
    width = None if rng.random() < 0.5 else float(rng.uniform(0, 1))
    rms = bool(rng.random() < 0.6)
    fc = (0, 1) if rng.random() < 0.7 else (0,)
    frames = int(rng.choice([1, 2, 7, 33, 500, int(rng.integers(1000, 120000))]))
    mono = bool(rng.random() < 0.3)
    dtype = str(rng.choice(["float32", "float64", "int16"]))
    seed = int(rng.integers(0, 10000))
    shape = (frames,) if mono else (frames, 2)
    if dtype == "int16":
        x = rng.integers(-20000, 20000, size=shape).astype(np.int16)
    else:
        x = (rng.standard_normal(shape) * 0.3).astype(dtype)
    return dict(fs=fs, dur=dur, n_imp=n_imp, kappa=kappa, env=env, mode=mode, width=width, rms=rms, fc=fc, seed=seed), x


def cases():
    rng = np.random.default_rng(SEED)
    for i in range(COUNT):
        p, x = random_vn_case(rng)
        yield i, p, x


def vn_kwargs(p):
    kw = dict(sample_rate_hz=p["fs"], duration_seconds=p["dur"], num_impulses=p["n_imp"], log_distribution_strength=p["kappa"],
              segment_envelope=p["env"], mode=p["mode"], width=p["width"], filtered_channels=p["fc"], seed=p["seed"])
    if not p["rms"]:
        kw["normalizer"] = None
    return kw


def oracle_output(O, p, x):
    taps = O.class_taps(sample_rate_hz=p["fs"], duration_seconds=p["dur"], num_impulses=p["n_imp"], num_outs=2, num_segments=len(p["env"]),
                        log_distribution_strength=p["kappa"], filtered_channels=p["fc"], seed=p["seed"])
    return O.vn_decorrelate(x, taps, envelope=p["env"], num_outs=2, ms_mode=p["mode"] == "MS", width=p["width"],
                            normalizer="rms" if p["rms"] else None)


# ---------------------------------------------------------------------------------------------------------------------
# more families: multichannel ``convolve`` (any layout), ``SignalChain`` velvet noise + Haas, the function path
# ---------------------------------------------------------------------------------------------------------------------
COUNT_MORE = 60


def random_convolve_case(rng):
    fs = int(rng.choice([16000, 44100, 48000, 96000]))
    dur = float(rng.uniform(0.002, 0.05))
    n_imp = int(rng.integers(1, 40))
    kappa = float(rng.choice([0.0, 0.5, 1.0, rng.uniform(0, 1)]))
    env = tuple(float(v) for v in rng.choice([1.0, 0.85, 0.55, 0.35, 0.2, -0.5], size=int(rng.integers(1, 5))))
    num_outs = int(rng.choice([1, 2, 3, 5, 8]))
    k = int(rng.integers(1, num_outs + 1))
    frames = int(rng.choice([3, 40, 700, int(rng.integers(1000, 60000))]))
    extra = int(rng.random() < 0.3)  # an input channel the filter does not use
    dtype = str(rng.choice(["float32", "float64"]))
    order = str(rng.choice(["C", "F"]))
    seed = int(rng.integers(0, 10000))
    x = np.asarray((rng.standard_normal((frames, num_outs + extra)) * 0.3).astype(dtype), order=order)
    return dict(fs=fs, dur=dur, n_imp=n_imp, kappa=kappa, env=env, num_outs=num_outs, fc=tuple(range(k)), seed=seed), x


def convolve_kwargs(p):
    return dict(sample_rate_hz=p["fs"], duration_seconds=p["dur"], num_impulses=p["n_imp"], log_distribution_strength=p["kappa"],
                segment_envelope=p["env"], num_outs=p["num_outs"], filtered_channels=p["fc"], mode="LR", normalizer=None, seed=p["seed"])


def oracle_convolve(O, p, x):
    taps = O.class_taps(sample_rate_hz=p["fs"], duration_seconds=p["dur"], num_impulses=p["n_imp"], num_outs=p["num_outs"],
                        num_segments=len(p["env"]), log_distribution_strength=p["kappa"], filtered_channels=p["fc"], seed=p["seed"])
    return O.fir_class_order(np.ascontiguousarray(x), taps, p["env"], p["num_outs"])


def random_chain_case(rng):
    fs = int(rng.choice([22050, 44100, 48000]))
    dur = float(rng.uniform(0.005, 0.04))
    n_imp = int(rng.integers(2, 40))
    kappa = float(rng.choice([0.0, 1.0, rng.uniform(0, 1)]))
    vn_mode = str(rng.choice(["MS", "LR"]))
    vn_width = None if rng.random() < 0.6 else float(rng.uniform(0, 1))
    delay = float(rng.choice([0.0, 0.0005, 0.02, rng.uniform(0, 0.03)]))
    delayed = int(rng.integers(0, 2))
    h_mode = str(rng.choice(["LR", "MS"]))
    h_width = None if rng.random() < 0.6 else float(rng.uniform(0, 1))
    frames = int(rng.choice([5, 300, int(rng.integers(1000, 80000))]))
    mono = bool(rng.random() < 0.3)
    dtype = str(rng.choice(["float32", "int16"]))
    seed = int(rng.integers(0, 10000))
    shape = (frames,) if mono else (frames, 2)
    x = rng.integers(-20000, 20000, size=shape).astype(np.int16) if dtype == "int16" else (rng.standard_normal(shape) * 0.3).astype(np.float32)
    return dict(fs=fs, dur=dur, n_imp=n_imp, kappa=kappa, vn_mode=vn_mode, vn_width=vn_width, delay=delay, delayed=delayed, h_mode=h_mode,
                h_width=h_width, seed=seed), x


def run_chain(SignalChain, p, x):
    chain = (SignalChain(sample_rate_hz=p["fs"])
             .velvet_noise(duration_seconds=p["dur"], num_impulses=p["n_imp"], log_distribution_strength=p["kappa"], mode=p["vn_mode"],
                           width=p["vn_width"], seed=p["seed"])
             .haas_effect(delay_time_seconds=p["delay"], delayed_channel=p["delayed"], mode=p["h_mode"], width=p["h_width"]))
    return chain(x)


def oracle_chain(O, p, x):
    taps = O.class_taps(sample_rate_hz=p["fs"], duration_seconds=p["dur"], num_impulses=p["n_imp"], log_distribution_strength=p["kappa"], seed=p["seed"])
    y = O.vn_decorrelate(x, taps, ms_mode=p["vn_mode"] == "MS", width=p["vn_width"])
    return O.haas(y, sample_rate_hz=p["fs"], delay_time_seconds=p["delay"], delayed_channel=p["delayed"], ms_mode=p["h_mode"] == "MS", width=p["h_width"])


def random_function_case(rng):
    fs = int(rng.choice([16000, 44100, 48000]))
    dur = float(rng.uniform(0.003, 0.05))
    n_imp = int(rng.integers(1, 40))
    kappa = float(rng.choice([0.0, 1.0, rng.uniform(0, 1)]))
    env = tuple(float(v) for v in rng.choice([1.0, 0.85, 0.55, 0.35, 0.2, -0.5], size=int(rng.integers(0, 5))))
    num_outs = int(rng.choice([1, 2, 4]))
    frames = int(rng.choice([4, 200, int(rng.integers(1000, 60000))]))
    dtype = str(rng.choice(["float32", "float64"]))
    seed = int(rng.integers(0, 10000))
    x = (rng.standard_normal((frames, num_outs)) * 0.3).astype(dtype)
    return dict(fs=fs, dur=dur, n_imp=n_imp, kappa=kappa, env=env, num_outs=num_outs, seed=seed), x


def function_kwargs(p):
    return dict(duration_seconds=p["dur"], num_impulses=p["n_imp"], num_outs=p["num_outs"], sample_rate_hz=p["fs"], segment_envelope=p["env"],
                log_distribution_strength=p["kappa"], seed=p["seed"])


def oracle_function(O, p, x):
    fir = O.dense_fir(**function_kwargs(p))
    return O.fir_function_order(x, fir)


FAMILIES = {"convolve": (random_convolve_case, 1), "chain": (random_chain_case, 2), "function": (random_function_case, 3)}


def more_cases(family: str):
    gen, salt = FAMILIES[family]
    rng = np.random.default_rng(SEED + salt)
    for i in range(COUNT_MORE):
        p, x = gen(rng)
        yield i, p, x
