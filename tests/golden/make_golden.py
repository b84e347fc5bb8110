"""Regenerate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (it reads ``/root/reference``; the GPU box has no such path):

    python tests/golden/make_golden.py

Writes ``cases.npz`` + ``cases.json`` (small seeded inputs with the reference's outputs),
``tables.npz`` (tap tables), ``hashes.json`` (sha256 of full-size outputs on the wav fixtures)
and ``objective.json`` (objective known answers).  The wav inputs under ``audio/`` are verbatim
copies of the reference's ``audio/{viola,guitar,vocal}.wav`` (public-domain fixtures; the
configs in BASELINE.json name them).  Nothing here is imported by the product.
"""

from __future__ import annotations

import hashlib
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))

import scipy.io.wavfile as wavfile  # noqa: E402
from vndecorrelate.decorrelation import (  # noqa: E402
    HaasEffect,
    SignalChain,
    VelvetNoise,
    convolve_velvet_noise,
    generate_velvet_noise,
)
from vndecorrelate.optimization import (  # noqa: E402
    get_local_minima,
    grid_scan,
    optimize_haas_delay,
    optimize_velvet_noise,
    symmetry_aware_objective,
)
from vndecorrelate.utils.dsp import (  # noqa: E402
    apply_stereo_width,
    encode_signal_to_side_channel,
    polar_coordinates,
    rms_normalize,
)

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rows(vn: VelvetNoise) -> np.ndarray:
    out = []
    for ch, seq in enumerate(vn.velvet_noise):
        for si, seg in enumerate(seq):
            out += [(ch, si, int(i), -1) for i in seg[0]]
            out += [(ch, si, int(i), 1) for i in seg[1]]
    return np.array(out, dtype=np.int32).reshape(-1, 4)


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main() -> None:
    arrays: dict[str, np.ndarray] = {}
    manifest: list[dict] = []

    def add(kind: str, params: dict, x: np.ndarray, y: np.ndarray, extra: dict | None = None):
        k = len(manifest)
        arrays[f"x{k}"] = x
        arrays[f"y{k}"] = y
        manifest.append({"id": k, "kind": kind, "params": params, **(extra or {})})

    rng = np.random.default_rng(20261018)

    # ---- A/B: VelvetNoise.convolve (fp32 and fp64 input), LR layouts of 1..4 channels -------
    for L in (1, 7, 1000, 3001):
        for num_outs, filt in ((2, (0, 1)), (1, (0,)), (3, (0, 1, 2)), (4, (0, 1)), (2, (0,))):
            p = dict(sample_rate_hz=44100, duration_seconds=0.03, num_impulses=30, num_outs=num_outs,
                     filtered_channels=list(filt), seed=3, mode="LR")
            vn = VelvetNoise(**{**p, "filtered_channels": tuple(filt)})
            x = (rng.standard_normal((L, num_outs)) * 0.25).astype(np.float32)
            add("vn_convolve", p, x, vn.convolve(x))
    for L in (513, 2500):
        p = dict(sample_rate_hz=44100, duration_seconds=0.03, num_impulses=30, num_outs=2,
                 filtered_channels=[0, 1], seed=1, mode="LR")
        vn = VelvetNoise(**{**p, "filtered_channels": (0, 1)})
        x = rng.random((L, 2))  # float64, like tests/test_decorrelation.py:173
        add("vn_convolve", p, x, vn.convolve(x))

    # ---- C: VelvetNoise.decorrelate variants ------------------------------------------------
    base = dict(sample_rate_hz=44100, duration_seconds=0.03, num_impulses=30, seed=1)
    variants = [
        {},
        {"width": 0.5},
        {"width": 0.0},
        {"width": 1.0},
        {"mode": "LR"},
        {"mode": "LR", "width": 0.3},
        {"normalizer": None},
        {"mode": "LR", "normalizer": None, "filtered_channels": [0]},
        {"segment_envelope": [1.0]},
        {"segment_envelope": [1.0, 0.5, 0.25]},
        {"segment_envelope": [], "mode": "LR"},
        {"log_distribution_strength": 0.0},
        {"log_distribution_strength": 0.37, "seed": 9},
        {"duration_seconds": 0.5, "num_impulses": 15},  # FIR (22050) longer than the signal
        {"duration_seconds": 0.055, "num_impulses": 45, "sample_rate_hz": 48000},
        {"num_impulses": 300, "duration_seconds": 0.3, "sample_rate_hz": 96000, "mode": "LR", "normalizer": None},
    ]
    for v in variants:
        p = {**base, **v}
        kw = dict(p)
        if "normalizer" in kw:
            assert kw["normalizer"] is None
        if "filtered_channels" in kw:
            kw["filtered_channels"] = tuple(kw["filtered_channels"])
        if "segment_envelope" in kw:
            kw["segment_envelope"] = tuple(kw["segment_envelope"])
        for L, mono in ((2000, False), (1000, True), (4097, False)):
            x = (rng.standard_normal(L if mono else (L, 2)) * 0.3).astype(np.float32)
            y = VelvetNoise(**kw).decorrelate(x)
            add("vn_decorrelate", {k: (None if (k == "normalizer") else val) for k, val in p.items()}, x, y,
                {"normalizer_none": "normalizer" in p})
    # envelope given as a long list after construction (tests/test_decorrelation.py:135-158)
    vn = VelvetNoise(num_impulses=15, duration_seconds=0.5, sample_rate_hz=44100, segment_envelope=(1.0, 0.5, 0.25), seed=4)
    x = (rng.standard_normal((30000, 2)) * 0.3).astype(np.float32)
    vn.segment_envelope = [1.0] * 1000
    add("vn_decorrelate_env_swap", dict(num_impulses=15, duration_seconds=0.5, sample_rate_hz=44100,
                                         segment_envelope=[1.0, 0.5, 0.25], seed=4, new_envelope_len=1000), x, vn(x))
    # int16 input is not rescaled (utils/dsp.py:66-68)
    xi = (rng.standard_normal((1500, 2)) * 8000).astype(np.int16)
    add("vn_decorrelate", dict(base), xi, VelvetNoise(**base).decorrelate(xi), {"normalizer_none": False})

    # ---- D: HaasEffect ----------------------------------------------------------------------
    for mode in ("LR", "MS"):
        for ch in (0, 1):
            for mono in (False, True):
                for width in (None, 0.3):
                    p = dict(sample_rate_hz=44100, delay_time_seconds=0.0113, delayed_channel=ch, mode=mode, width=width)
                    x = (rng.standard_normal(700 if mono else (700, 2)) * 0.3).astype(np.float32)
                    add("haas", p, x, HaasEffect(**p).decorrelate(x))
    p = dict(sample_rate_hz=44100, delay_time_seconds=0.0, delayed_channel=0, mode="LR", width=None)
    x = (rng.standard_normal((64, 2)) * 0.3).astype(np.float32)
    add("haas", p, x, HaasEffect(**p).decorrelate(x))
    x = rng.standard_normal((300, 2))  # float64 input is cast to fp32 first (decorrelation.py:194)
    p = dict(sample_rate_hz=48000, delay_time_seconds=0.02, delayed_channel=1, mode="MS", width=None)
    add("haas", p, x, HaasEffect(**p).decorrelate(x))

    # ---- E: chains --------------------------------------------------------------------------
    x = (rng.standard_normal((5000, 2)) * 0.3).astype(np.float32)
    chain = (SignalChain(sample_rate_hz=44100)
             .velvet_noise(duration_seconds=0.03, num_impulses=30, log_distribution_strength=1.0, seed=1)
             .haas_effect(delay_time_seconds=0.02, mode="LR"))
    add("chain_cfg2", {}, x, chain(x))
    chain = (SignalChain(sample_rate_hz=44100)
             .velvet_noise(duration_seconds=0.02, num_impulses=30, seed=1, log_distribution_strength=1.0, mode="MS", filtered_channels=(0, 1))
             .haas_effect(delay_time_seconds=0.02, delayed_channel=1, mode="LR"))
    add("chain_example", {}, x, chain(x))
    chain = (SignalChain(sample_rate_hz=44100)
             .velvet_noise(duration_seconds=0.03, num_impulses=30, width=0.5, seed=5)
             .haas_effect(delay_time_seconds=0.0197, delayed_channel=1, mode="LR")
             .haas_effect(delay_time_seconds=0.0096, delayed_channel=1, mode="MS"))
    xm = (rng.standard_normal(1000) * 0.3).astype(np.float32)
    add("chain_hetero", {}, xm, chain(xm))

    # ---- F: convolve_velvet_noise -----------------------------------------------------------
    gkw = dict(duration_seconds=0.03, num_impulses=30, num_outs=2, sample_rate_hz=44100,
               segment_envelope=(0.85, 0.55, 0.35, 0.2), log_distribution_strength=1.0, seed=1)
    fir32 = generate_velvet_noise(**gkw)
    arrays["fir32_cfg1"] = fir32
    x32 = (rng.standard_normal((3000, 2)) * 0.3).astype(np.float32)
    x64 = rng.random((3000, 2))
    add("fn_convolve", {"fir": "fir32_cfg1"}, x32, convolve_velvet_noise(x32, fir32))
    add("fn_convolve", {"fir": "fir32_cfg1"}, x64, convolve_velvet_noise(x64, fir32))
    fir64 = VelvetNoise(sample_rate_hz=44100, duration_seconds=0.03, num_impulses=30, seed=1).FIR
    arrays["fir64_cfg1"] = fir64
    add("fn_convolve", {"fir": "fir64_cfg1"}, x32, convolve_velvet_noise(x32, fir64))
    xs = (rng.standard_normal((200, 2)) * 0.3).astype(np.float32)  # shorter than the FIR
    add("fn_convolve", {"fir": "fir32_cfg1"}, xs, convolve_velvet_noise(xs, fir32))
    g3 = dict(duration_seconds=0.05, num_impulses=40, num_outs=3, sample_rate_hz=48000,
              segment_envelope=(1.0, 0.5), log_distribution_strength=0.4, seed=11)
    arrays["fir32_g3"] = generate_velvet_noise(**g3)
    x3 = (rng.standard_normal((2600, 3)) * 0.3).astype(np.float32)
    add("fn_convolve", {"fir": "fir32_g3"}, x3, convolve_velvet_noise(x3, arrays["fir32_g3"]))

    # ---- helpers: encode / width / rms on their own -----------------------------------------
    a = (rng.standard_normal((999, 2)) * 0.3).astype(np.float32)
    b = (rng.standard_normal((999, 2)) * 0.3).astype(np.float32)
    t = b.copy(); encode_signal_to_side_channel(a, t); add("encode", {}, np.stack((a, b)), t)
    t = b.copy(); apply_stereo_width(t, 0.35); add("width", {"width": 0.35}, b, t)
    t = b.copy(); rms_normalize(a, t); add("rms", {}, np.stack((a, b)), t)
    big = (rng.standard_normal((70001, 2)) * 0.3).astype(np.float32)
    t = big[::-1].copy(); rms_normalize(big, t); add("rms", {}, np.stack((big, big[::-1])), t)

    np.savez_compressed(os.path.join(HERE, "cases.npz"), **arrays)
    json.dump(manifest, open(os.path.join(HERE, "cases.json"), "w"), indent=1)

    # ---- tap tables -------------------------------------------------------------------------
    tabs: dict[str, np.ndarray] = {}
    meta: dict[str, dict] = {}
    vn1 = VelvetNoise(sample_rate_hz=44100, duration_seconds=0.03, num_impulses=30, seed=1)
    tabs["cfg1"] = rows(vn1)
    meta["cfg1"] = {"sha256": sha(tabs["cfg1"])}
    vn3 = VelvetNoise(sample_rate_hz=48000, duration_seconds=0.03, num_impulses=30, num_outs=4096,
                      filtered_channels=tuple(range(4096)), mode="LR", normalizer=None, seed=1)
    r3 = rows(vn3)
    meta["cfg3"] = {"sha256": sha(r3), "shape": list(r3.shape)}
    tabs["cfg3_head"] = r3[r3[:, 0] < 4]
    tabs["cfg3_tail"] = r3[r3[:, 0] >= 4092]
    vn4 = VelvetNoise(sample_rate_hz=96000, duration_seconds=0.3, num_impulses=300, num_outs=4,
                      filtered_channels=(0, 1, 2, 3), mode="LR", normalizer=None, seed=1)
    tabs["cfg4_4ch"] = rows(vn4)
    meta["cfg4_4ch"] = {"sha256": sha(tabs["cfg4_4ch"])}
    vn4w = VelvetNoise(sample_rate_hz=96000, duration_seconds=0.3, num_impulses=300, num_outs=1024,
                       filtered_channels=tuple(range(1024)), mode="LR", normalizer=None, seed=1)
    r4 = rows(vn4w)
    meta["cfg4"] = {"sha256": sha(r4), "shape": list(r4.shape), "max_index": int(r4[:, 2].max())}
    for k in (0.0, 0.123, 0.5, 1.0):
        vk = VelvetNoise(sample_rate_hz=48000, duration_seconds=0.03, num_impulses=30, log_distribution_strength=k,
                         normalizer=None, filtered_channels=(0,), mode="LR", seed=1)
        tabs[f"cand_{k}"] = rows(vk)
    kap = np.linspace(0.0, 1.0, 1024)
    h = hashlib.sha256()
    for k in kap:
        vk = VelvetNoise(sample_rate_hz=48000, duration_seconds=0.03, num_impulses=30, log_distribution_strength=k,
                         normalizer=None, filtered_channels=(0,), mode="LR", seed=1)
        h.update(rows(vk).tobytes())
    meta["cfg5_1024_candidates"] = {"sha256_concat": h.hexdigest()}
    # generate_velvet_noise non-zeros
    for name, kw in (("gen_cfg1", gkw), ("gen_g3", g3),
                     ("gen_trunc", dict(duration_seconds=0.0301, num_impulses=20, num_outs=2, sample_rate_hz=44100, seed=2))):
        f = generate_velvet_noise(**kw)
        nz = np.argwhere(f != 0)
        tabs[name + "_nz"] = nz.astype(np.int32)
        tabs[name + "_val"] = f[f != 0]
        meta[name] = {"shape": list(f.shape), "kwargs": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}}
    fir_prop = VelvetNoise(sample_rate_hz=44100, duration_seconds=0.055, num_impulses=45, seed=6).FIR
    tabs["fir_prop_055_45_seed6"] = fir_prop
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **tabs)

    # ---- full-size hashes on the wav fixtures -----------------------------------------------
    hashes: dict[str, object] = {"tables": meta}
    wavs = {n: wavfile.read(os.path.join(REF, "audio", n + ".wav"))[1] for n in ("viola", "guitar", "vocal", "drums")}
    hashes["wav_sha256"] = {n: sha(a) for n, a in wavs.items()}
    hashes["wav_shape"] = {n: list(a.shape) for n, a in wavs.items()}
    hashes["cfg1_convolve_viola"] = sha(vn1.convolve(wavs["viola"]))
    y1 = vn1.decorrelate(wavs["viola"])
    hashes["cfg1_decorrelate_viola"] = sha(y1)
    arrays_full = {"cfg1_decorrelate_viola_head": y1[:4096], "cfg1_decorrelate_viola_tail": y1[-4096:]}
    cfg2 = (SignalChain(sample_rate_hz=44100)
            .velvet_noise(duration_seconds=0.03, num_impulses=30, log_distribution_strength=1.0, seed=1)
            .haas_effect(delay_time_seconds=0.02, mode="LR"))
    y2 = cfg2(wavs["guitar"])
    hashes["cfg2_chain_guitar"] = {"sha256": sha(y2), "shape": list(y2.shape), "dtype": str(y2.dtype)}
    arrays_full["cfg2_chain_guitar_head"] = y2[:2048]
    arrays_full["cfg2_chain_guitar_tail"] = y2[-2048:]
    ex = (SignalChain(sample_rate_hz=44100)
          .velvet_noise(duration_seconds=0.02, num_impulses=30, seed=1, log_distribution_strength=1.0, mode="MS", filtered_channels=(0, 1))
          .haas_effect(delay_time_seconds=0.02, delayed_channel=1, mode="LR"))
    hashes["example_chain"] = {}
    for n in ("viola", "vocal", "guitar", "drums"):
        y = ex(wavs[n])
        hashes["example_chain"][n] = {"sha256": sha(y), "shape": list(y.shape)}
    # the two outputs the reference commits must be what its own code produces here
    for n in ("viola", "vocal"):
        committed = wavfile.read(os.path.join(REF, "audio", n + "_decorrelated.wav"))[1]
        assert sha(committed) == hashes["example_chain"][n]["sha256"], n
        hashes["example_chain"][n]["equals_reference_committed_wav"] = True
    gains = np.sqrt(np.mean(np.square(wavs["viola"]), axis=0))
    hashes["rms_viola_input"] = [float(g) for g in gains]
    np.savez_compressed(os.path.join(HERE, "full_excerpts.npz"), **arrays_full)
    json.dump(hashes, open(os.path.join(HERE, "hashes.json"), "w"), indent=1)

    # ---- objective known answers ------------------------------------------------------------
    obj: dict[str, object] = {}
    viola = wavs["viola"]
    okw = dict(angle_limit=np.pi / 4, lambda_mean=5.0, lambda_skew=2.0, lambda_correlation=15.0, lambda_penalty=1e3)
    rowsv = []
    for k in (0.0, 0.25, 0.5, 0.75, 1.0):
        d = VelvetNoise(sample_rate_hz=44100, duration_seconds=0.03, num_impulses=30, log_distribution_strength=k,
                        normalizer=None, filtered_channels=(0,), mode="LR", seed=1)
        y = d.decorrelate(viola)
        r, th, w = polar_coordinates(y[:, 0], y[:, 1], normalize=False)
        rowsv.append({"kappa": k, "objective": float(symmetry_aware_objective(viola, d, **okw)),
                      "spread": float(np.sum(w * th**2)), "centroid": float(np.sum(w * th)),
                      "m3": float(np.sum(w * th**3)), "max_abs_theta": float(np.max(np.abs(th))),
                      "sum_r": float(r.sum()), "dot_lr": float(np.dot(y[:, 0], y[:, 1])),
                      "norm_l": float(np.linalg.norm(y[:, 0]))})
    obj["viola_vn"] = rowsv
    obj["viola_haas"] = [{"tau": t, "objective": float(symmetry_aware_objective(
        viola, HaasEffect(sample_rate_hz=44100, delay_time_seconds=t, mode="LR"), **okw))} for t in (0.0, 0.005, 0.01, 0.02)]
    # small sweeps (grid stage) on an excerpt and on a coloured clip
    from scipy.signal import lfilter

    def clip(i, n):
        g = np.random.default_rng(1000 + i)
        m = lfilter([0.02], [1, -0.98], g.standard_normal(n))
        s = 0.3 * lfilter([0.02], [1, -0.98], g.standard_normal(n))
        x = np.column_stack((m + s, m - s))
        return (x / np.max(np.abs(x)) * 0.5).astype(np.float32)

    sweeps = {}
    for name, sig, fs in (("viola_60k", viola[40000:100000], 44100), ("clip0_48k", clip(0, 48000), 48000), ("clip1_96k", clip(1, 96000), 48000)):
        kappas = np.linspace(0.0, 1.0, 32)
        ds = [VelvetNoise(sample_rate_hz=fs, duration_seconds=0.03, num_impulses=30, log_distribution_strength=k,
                          normalizer=None, filtered_channels=(0,), mode="LR", seed=1) for k in kappas]
        sc = quiet(grid_scan, sig, ds, **okw)
        taus = np.linspace(0.0, 0.03, 32)
        hs = [HaasEffect(sample_rate_hz=fs, delay_time_seconds=t, mode="LR") for t in taus]
        sh = quiet(grid_scan, sig, hs, **okw)
        sweeps[name] = {"fs": fs, "vn_scores": [float(s) for s in sc], "vn_dtype": str(sc.dtype),
                        "vn_argmin": int(np.argmin(sc)), "vn_minima": get_local_minima(sc, 32),
                        "haas_scores": [float(s) for s in sh], "haas_argmin": int(np.argmin(sh)),
                        "haas_minima": get_local_minima(sh, 32), "input_sha256": sha(sig)}
    obj["sweeps"] = sweeps
    # full optimisers on a short excerpt (keeps the generator run short)
    ex_sig = viola[40000:80000]
    obj["optimize_vn_viola_40k"] = {
        "kappa": float(quiet(optimize_velvet_noise, input_signal=ex_sig, sample_rate_hz=44100, duration_seconds=0.03,
                             num_impulses=15, seed=1, grid_size=48)), "grid_size": 48, "num_impulses": 15}
    obj["optimize_haas_viola_40k"] = {
        "tau": float(quiet(optimize_haas_delay, input_signal=ex_sig, sample_rate_hz=44100, max_delay_seconds=0.03, grid_size=48)),
        "grid_size": 48}
    json.dump(obj, open(os.path.join(HERE, "objective.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
