"""Loaders for the fixtures in tests/golden (produced by tests/golden/make_golden.py)."""

from __future__ import annotations

import functools
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@functools.lru_cache(maxsize=None)
def cases():
    man = json.load(open(os.path.join(GOLDEN, "cases.json")))
    arr = np.load(os.path.join(GOLDEN, "cases.npz"))
    return man, arr


def case_list(kind: str):
    man, _ = cases()
    return [c for c in man if c["kind"] == kind]


def case_xy(c):
    _, arr = cases()
    return arr[f"x{c['id']}"], arr[f"y{c['id']}"]


@functools.lru_cache(maxsize=None)
def tables():
    return np.load(os.path.join(GOLDEN, "tables.npz"))


@functools.lru_cache(maxsize=None)
def hashes():
    return json.load(open(os.path.join(GOLDEN, "hashes.json")))


@functools.lru_cache(maxsize=None)
def objective():
    return json.load(open(os.path.join(GOLDEN, "objective.json")))


@functools.lru_cache(maxsize=None)
def excerpts():
    return np.load(os.path.join(GOLDEN, "full_excerpts.npz"))


@functools.lru_cache(maxsize=None)
def wav(name: str):
    import scipy.io.wavfile as wavfile

    fs, data = wavfile.read(os.path.join(GOLDEN, "audio", name + ".wav"))
    return fs, data


def same_bits(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit equality including dtype and shape (distinguishes -0.0 from +0.0)."""
    return a.dtype == b.dtype and a.shape == b.shape and np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


def dsp_inputs(kind: str, seed: int, n: int, dtype):
    """The seeded inputs of one case of dsp_extra.json (same draws as tests/golden/make_golden_r02.py::dsp_inputs)."""
    rng = np.random.default_rng(seed)
    if kind == "polar_coordinates":
        return (rng.standard_normal(n) * 0.3).astype(dtype), (rng.standard_normal(n) * 0.3).astype(dtype)
    return (rng.standard_normal((n, 2)) * 0.3).astype(dtype), (rng.standard_normal((n, 2)) * 0.1).astype(dtype)


@functools.lru_cache(maxsize=None)
def dsp_extra():
    man = json.load(open(os.path.join(GOLDEN, "dsp_extra.json")))
    arr = np.load(os.path.join(GOLDEN, "dsp_extra.npz"))
    return man, arr
