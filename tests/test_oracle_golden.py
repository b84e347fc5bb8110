"""Pins the CPU oracle (oracle/vnd_oracle.py) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  Everything here is bit-exact unless a tolerance is
written next to the assertion.  CPU only."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import vnd_oracle as O
from tests import _golden as G


def _taps_from_params(p):
    env = p.get("segment_envelope", list(O.DEFAULT_ENVELOPE))
    n_seg = len(env) if len(env) else 1
    return O.class_taps(
        sample_rate_hz=p["sample_rate_hz"],
        duration_seconds=p.get("duration_seconds", 0.03),
        num_impulses=p.get("num_impulses", 30),
        num_outs=p.get("num_outs", 2),
        num_segments=n_seg,
        log_distribution_strength=p.get("log_distribution_strength", 1.0),
        filtered_channels=tuple(p.get("filtered_channels", (0, 1))),
        seed=p.get("seed"),
    ), (tuple(env) if len(env) else (1.0,))


# ---------------------------------------------------------------- tap tables


def test_table_cfg1_rows_and_hash():
    t = O.class_taps(sample_rate_hz=44100, seed=1)
    rows = O.table_rows(t)
    assert np.array_equal(rows, G.tables()["cfg1"])
    assert G.sha(rows) == G.hashes()["tables"]["cfg1"]["sha256"]
    # SURVEY.md Appendix C, channel 0, first segment: (4,-) (7,-) (22,-) (28,-) (2,+) (10,+) (13,+) (17,+)
    assert [int(i) for i in t[0][0][0]] == [4, 7, 22, 28]
    assert [int(i) for i in t[0][0][1]] == [2, 10, 13, 17]


def test_table_cfg3_hash():
    t = O.class_taps(sample_rate_hz=48000, num_outs=4096, filtered_channels=tuple(range(4096)), seed=1)
    rows = O.table_rows(t)
    assert list(rows.shape) == G.hashes()["tables"]["cfg3"]["shape"]
    assert G.sha(rows) == G.hashes()["tables"]["cfg3"]["sha256"]


def test_table_cfg4_hashes():
    t = O.class_taps(sample_rate_hz=96000, duration_seconds=0.3, num_impulses=300, num_outs=4, filtered_channels=(0, 1, 2, 3), seed=1)
    assert G.sha(O.table_rows(t)) == G.hashes()["tables"]["cfg4_4ch"]["sha256"]


@pytest.mark.parametrize("kappa", [0.0, 0.123, 0.5, 1.0])
def test_candidate_tables(kappa):
    t = O.vn_candidate_taps(kappa, sample_rate_hz=48000, duration_seconds=0.03, num_impulses=30, seed=1)
    assert np.array_equal(O.table_rows(t), G.tables()[f"cand_{kappa}"])
    assert t[1] == []  # channel 1 is not filtered


def test_filtered_channel_quirk_raises_like_reference():
    with pytest.raises(IndexError):  # decorrelation.py:531 indexes the draw by channel number
        O.class_taps(sample_rate_hz=44100, filtered_channels=(1,), seed=1)


@pytest.mark.parametrize("name", ["gen_cfg1", "gen_g3", "gen_trunc"])
def test_dense_fir(name):
    meta = G.hashes()["tables"][name]
    kw = dict(meta["kwargs"])
    if "segment_envelope" in kw:
        kw["segment_envelope"] = tuple(kw["segment_envelope"])
    f = O.dense_fir(**kw)
    assert list(f.shape) == meta["shape"] and f.dtype == np.float32
    assert np.array_equal(np.argwhere(f != 0).astype(np.int32), G.tables()[name + "_nz"])
    assert np.array_equal(f[f != 0], G.tables()[name + "_val"])


def test_class_dense_fir_property():
    t = O.class_taps(sample_rate_hz=44100, duration_seconds=0.055, num_impulses=45, seed=6)
    f = O.class_dense_fir(t, O.DEFAULT_ENVELOPE, O.class_fir_length(44100, 0.055))
    assert G.same_bits(f, G.tables()["fir_prop_055_45_seed6"])
    assert f.shape == (2426, 2)  # tests/test_decorrelation.py:108


# ---------------------------------------------------------------- FIR, decorrelate, Haas, chains


@pytest.mark.parametrize("c", G.case_list("vn_convolve"), ids=lambda c: f"case{c['id']}")
def test_vn_convolve_cases(c):
    x, y = G.case_xy(c)
    taps, env = _taps_from_params(c["params"])
    got = O.fir_class_order(x, taps, env, c["params"]["num_outs"])
    assert G.same_bits(got, y)


@pytest.mark.parametrize("c", G.case_list("vn_decorrelate"), ids=lambda c: f"case{c['id']}")
def test_vn_decorrelate_cases(c):
    x, y = G.case_xy(c)
    p = c["params"]
    taps, env = _taps_from_params(p)
    got = O.vn_decorrelate(
        x, taps, envelope=env, num_outs=p.get("num_outs", 2), ms_mode=p.get("mode", "MS") == "MS",
        width=p.get("width"), normalizer=None if c["normalizer_none"] else "rms",
    )
    assert G.same_bits(got, y)


def test_vn_envelope_swapped_after_construction():
    (c,) = G.case_list("vn_decorrelate_env_swap")
    x, y = G.case_xy(c)
    p = c["params"]
    # taps were generated for 3 segments; the 1000-entry envelope is only looked up by segment index
    taps = O.class_taps(sample_rate_hz=44100, duration_seconds=0.5, num_impulses=15, num_segments=3, seed=4)
    got = O.vn_decorrelate(x, taps, envelope=[1.0] * p["new_envelope_len"])
    assert G.same_bits(got, y)


@pytest.mark.parametrize("c", G.case_list("haas"), ids=lambda c: f"case{c['id']}")
def test_haas_cases(c):
    x, y = G.case_xy(c)
    p = c["params"]
    got = O.haas(x, sample_rate_hz=p["sample_rate_hz"], delay_time_seconds=p["delay_time_seconds"],
                 delayed_channel=p["delayed_channel"], ms_mode=p["mode"] == "MS", width=p["width"])
    assert got.dtype == np.float64
    assert G.same_bits(got, y)


def _chain_cfg2(x, fs=44100):
    t = O.class_taps(sample_rate_hz=fs, seed=1)
    return O.haas(O.vn_decorrelate(x, t), sample_rate_hz=fs, delay_time_seconds=0.02)


def _chain_example(x, fs=44100):
    t = O.class_taps(sample_rate_hz=fs, duration_seconds=0.02, seed=1)
    return O.haas(O.vn_decorrelate(x, t), sample_rate_hz=fs, delay_time_seconds=0.02, delayed_channel=1)


def test_chain_cases():
    (c,) = G.case_list("chain_cfg2")
    x, y = G.case_xy(c)
    assert G.same_bits(_chain_cfg2(x), y)
    (c,) = G.case_list("chain_example")
    x, y = G.case_xy(c)
    assert G.same_bits(_chain_example(x), y)
    (c,) = G.case_list("chain_hetero")
    x, y = G.case_xy(c)
    t = O.class_taps(sample_rate_hz=44100, seed=5)
    z = O.vn_decorrelate(x, t, width=0.5)
    z = O.haas(z, sample_rate_hz=44100, delay_time_seconds=0.0197, delayed_channel=1)
    z = O.haas(z, sample_rate_hz=44100, delay_time_seconds=0.0096, delayed_channel=1, ms_mode=True)
    assert G.same_bits(z, y)


@pytest.mark.parametrize("c", G.case_list("fn_convolve"), ids=lambda c: f"case{c['id']}")
def test_fn_convolve_cases(c):
    x, y = G.case_xy(c)
    fir = G.cases()[1][c["params"]["fir"]]
    assert G.same_bits(O.fir_function_order(x, fir), y)


def test_helpers():
    (c,) = G.case_list("encode")
    xy, y = G.case_xy(c)
    t = xy[1].copy()
    O.side_encode(xy[0], t)
    assert G.same_bits(t, y)
    (c,) = G.case_list("width")
    x, y = G.case_xy(c)
    t = x.copy()
    O.stereo_width(t, c["params"]["width"])
    assert G.same_bits(t, y)
    for c in G.case_list("rms"):
        xy, y = G.case_xy(c)
        t = xy[1].copy()
        O.rms_match(xy[0], t)
        assert G.same_bits(t, y)
        # the sequential-sum restatement (what the CUDA gain kernel implements) gives the same gain
        for ch in range(2):
            mx = O.seq_sumsq_f32(xy[0][:, ch]) / np.float32(len(t))
            my = O.seq_sumsq_f32(xy[1][:, ch]) / np.float32(len(t))
            gain = np.sqrt(mx) / np.sqrt(my + np.float32(1e-10))
            assert G.same_bits((xy[1][:, ch] * gain).astype(np.float32), np.ascontiguousarray(y[:, ch]))


def test_reference_dsp_known_values():
    # tests/test_dsp.py:119-142 (STEREO-mode known value) and :22-46 (width extremes)
    x = np.array([[0.707, 0.3535], [0.707, 0.3535]])
    y = np.array([[1.0, 1.0], [1.0, 1.0]])
    O.rms_match(x, y, stereo_mode=True)
    assert np.allclose(np.array([0.55893258, 0.55893258]), y)
    a = np.column_stack((np.ones(100), np.zeros(100))).astype(np.float32)
    O.stereo_width(a, 1.0)
    O.lr_to_ms(a)
    assert np.sum(a[:, 0]) == 0.0 and np.sum(a[:, 1]) != 0.0
    with pytest.raises(ValueError):
        O.side_encode(np.zeros((100, 2)), np.zeros((120, 3)))


# ---------------------------------------------------------------- full-size wav goldens


def test_cfg1_viola_full():
    fs, x = G.wav("viola")
    assert G.sha(x) == G.hashes()["wav_sha256"]["viola"]
    t = O.class_taps(sample_rate_hz=fs, seed=1)
    assert G.sha(O.fir_class_order(x, t, O.DEFAULT_ENVELOPE, 2)) == G.hashes()["cfg1_convolve_viola"]
    y = O.vn_decorrelate(x, t)
    assert G.sha(y) == G.hashes()["cfg1_decorrelate_viola"]
    assert G.same_bits(y[:4096], G.excerpts()["cfg1_decorrelate_viola_head"])


def test_cfg2_guitar_full():
    fs, x = G.wav("guitar")
    y = _chain_cfg2(x, fs)
    h = G.hashes()["cfg2_chain_guitar"]
    assert list(y.shape) == h["shape"] and str(y.dtype) == h["dtype"]
    assert G.sha(y) == h["sha256"]


@pytest.mark.parametrize("name", ["viola", "vocal"])
def test_reference_committed_goldens(name):
    """audio/{viola,vocal}_decorrelated.wav of the reference == tests/test_example.py chain."""
    fs, x = G.wav(name)
    h = G.hashes()["example_chain"][name]
    assert h["equals_reference_committed_wav"]
    y = _chain_example(x, fs)
    assert list(y.shape) == h["shape"]
    assert G.sha(y) == h["sha256"]


# ---------------------------------------------------------------- objective and sweep


def test_objective_known_answers_viola():
    fs, x = G.wav("viola")
    for row in G.objective()["viola_vn"]:
        t = O.vn_candidate_taps(row["kappa"], sample_rate_hz=fs, duration_seconds=0.03, num_impulses=30, seed=1)
        y = O.vn_candidate_signal(x, t)
        got = O.objective(y)
        assert isinstance(got, np.float32)
        # same numpy on the same machine reproduces the reference bit-for-bit; across machines the
        # fp32 arctan2 may differ in the last ulp, worth up to ~2e-4 here
        assert abs(float(got) - row["objective"]) <= 5e-4
    for row in G.objective()["viola_haas"]:
        y = O.haas(x, sample_rate_hz=fs, delay_time_seconds=row["tau"])
        got = O.objective(y)
        assert isinstance(got, np.float64)
        assert abs(float(got) - row["objective"]) <= 1e-9 * max(1.0, abs(row["objective"]))


def test_small_sweeps():
    fs, viola = G.wav("viola")
    sw = G.objective()["sweeps"]
    sigs = {"viola_60k": viola[40000:100000], "clip0_48k": O.coloured_clip(0, 48000), "clip1_96k": O.coloured_clip(1, 96000)}
    for name, sig in sigs.items():
        s = sw[name]
        assert G.sha(sig) == s["input_sha256"]
        sc = O.vn_grid_scores(sig, np.linspace(0.0, 1.0, 32), sample_rate_hz=s["fs"], duration_seconds=0.03, num_impulses=30, seed=1)
        assert str(sc.dtype) == s["vn_dtype"]
        assert np.max(np.abs(sc.astype(np.float64) - np.array(s["vn_scores"]))) <= 5e-4
        assert int(np.argmin(sc)) == s["vn_argmin"]
        sh = O.haas_grid_scores(sig, np.linspace(0.0, 0.03, 32), sample_rate_hz=s["fs"])
        assert np.allclose(sh, np.array(s["haas_scores"]), rtol=1e-9, atol=1e-9)
        assert int(np.argmin(sh)) == s["haas_argmin"]
        assert O.local_minima(sh) == s["haas_minima"]


def test_optimisers_short_excerpt():
    fs, viola = G.wav("viola")
    sig = viola[40000:80000]
    ref = G.objective()["optimize_vn_viola_40k"]
    k, _, _ = O.optimize_vn(sig, sample_rate_hz=fs, duration_seconds=0.03, num_impulses=ref["num_impulses"], seed=1, grid_size=ref["grid_size"])
    assert abs(k - ref["kappa"]) <= 1e-4  # Brent's xatol (optimization.py:148)
    ref = G.objective()["optimize_haas_viola_40k"]
    t, _, _ = O.optimize_haas(sig, sample_rate_hz=fs, max_delay_seconds=0.03, grid_size=ref["grid_size"])
    assert abs(t - ref["tau"]) <= 1e-9


def test_oracle_random_tap_tables_match_reference():
    """The oracle's tap generation against the 240 random tuples hashed from the reference
    (tests/golden/make_golden_r02.py::random_tables)."""
    import json
    import os

    recs = json.load(open(os.path.join(G.GOLDEN, "tables_random.json")))
    for rec in recs[::3]:  # the pure-Python loops of the oracle are slow: every third tuple
        p = rec["params"]
        taps = O.class_taps(sample_rate_hz=p["sample_rate_hz"], duration_seconds=p["duration_seconds"], num_impulses=p["num_impulses"],
                            num_outs=p["num_outs"], num_segments=len(p["segment_envelope"]), log_distribution_strength=p["log_distribution_strength"],
                            filtered_channels=tuple(p["filtered_channels"]), seed=p["seed"])
        assert G.sha(O.table_rows(taps)) == rec["class_rows_sha256"], p
        fir = O.dense_fir(duration_seconds=p["duration_seconds"], num_impulses=p["num_impulses"], num_outs=p["num_outs"],
                          sample_rate_hz=p["sample_rate_hz"], segment_envelope=tuple(p["segment_envelope"]),
                          log_distribution_strength=p["log_distribution_strength"], seed=p["seed"])
        assert G.sha(fir) == rec["dense_sha256"], p


def test_oracle_matches_reference_on_random_decorrelate_cases():
    """The 120 seeded random ``VelvetNoise.decorrelate`` cases of tests/_random_cases.py: the oracle's output hashes (dtype,
    shape, sha256 of the bytes) equal the ones tests/golden/make_golden_r02b.py took from the unmodified reference; the
    cases the reference rejects are listed with its exception type and are checked against the library on the GPU."""
    import hashlib
    import json
    import os

    from tests import _random_cases as RC

    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "random_decorrelate.json")))
    assert golden["seed"] == RC.SEED and golden["count"] == RC.COUNT == len(golden["cases"])
    checked = 0
    for (i, p, x), ref in zip(RC.cases(), golden["cases"]):
        assert ref["id"] == i
        if "error" in ref:
            continue
        y = np.ascontiguousarray(RC.oracle_output(O, p, x))
        assert str(y.dtype) == ref["dtype"] and list(y.shape) == ref["shape"], (i, p)
        assert hashlib.sha256(y.tobytes()).hexdigest() == ref["sha256"], (i, p)
        checked += 1
    assert checked >= 100


@pytest.mark.parametrize("family", ["convolve", "chain", "function"])
def test_oracle_matches_reference_on_more_random_families(family):
    """60 seeded random cases each of multichannel ``convolve`` (C and Fortran order, float32 / float64, an unused input
    channel), ``SignalChain`` velvet noise + Haas (all modes, widths, mono, int16) and the function path
    ``convolve_velvet_noise(generate_velvet_noise(...))``: the oracle reproduces the reference's hashes."""
    import hashlib
    import json
    import os

    from tests import _random_cases as RC

    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "random_decorrelate.json")))
    fn = {"convolve": RC.oracle_convolve, "chain": RC.oracle_chain, "function": RC.oracle_function}[family]
    rows = golden["more"][family]
    assert golden["count_more"] == RC.COUNT_MORE == len(rows)
    checked = 0
    for (i, p, x), ref in zip(RC.more_cases(family), rows):
        if "error" in ref:
            continue
        y = np.ascontiguousarray(fn(O, p, x))
        assert str(y.dtype) == ref["dtype"] and list(y.shape) == ref["shape"], (i, p)
        assert hashlib.sha256(y.tobytes()).hexdigest() == ref["sha256"], (i, p)
        checked += 1
    assert checked >= 55
