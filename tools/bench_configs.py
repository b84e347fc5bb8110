#!/usr/bin/env python
"""Timings of the BASELINE.json configurations that are not the bench.py line (run on a GPU box).

cfg1  VelvetNoise(0.03 s, 30 impulses, seed 1).decorrelate(viola)            numpy in/out and CUDA-tensor in/out
cfg2  SignalChain velvet_noise + haas_effect on guitar (one fused launch)   numpy in/out and CUDA-tensor in/out
cfg4  300 impulses over 0.3 s at 96 kHz (28 800-sample halo), planar slab    device resident, Gsamples/s
cfg5  optimize_velvet_noise grid stage: clips x 1024 candidates             device resident, evaluations/s

One JSON object per configuration on stdout (and in gpurun_out/configs.json).  CPU columns time the
oracle (the numpy port of the reference loops) on the same inputs or on a stated sub-sample.
"""

from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import vnd_oracle as O  # noqa: E402
from tests import _golden as G  # noqa: E402
from vndecorrelate_b200 import decorrelation as D  # noqa: E402
from vndecorrelate_b200 import optimization as OPT  # noqa: E402
from vndecorrelate_b200 import taps as T  # noqa: E402


def cuda_ms(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))


def wall_ms(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def main():
    out = []
    dev = torch.device("cuda", 0)

    # ---- cfg1
    fs, x = G.wav("viola")
    vn = D.VelvetNoise(sample_rate_hz=fs, duration_seconds=0.03, num_impulses=30, seed=1)
    xt = torch.from_numpy(x).to(dev)
    taps = O.class_taps(sample_rate_hz=fs, seed=1)
    t_cpu = wall_ms(lambda: O.vn_decorrelate(x, taps), reps=5, warm=1)
    rec = {"config": "cfg1 viola VelvetNoise.decorrelate", "frames": int(x.shape[0]), "numpy_in_out_ms": wall_ms(lambda: vn.decorrelate(x)),
           "cuda_tensor_ms": cuda_ms(lambda: vn.decorrelate(xt)), "cpu_oracle_ms": t_cpu, "samples_out": int(x.size)}
    rec["Msamples_per_s_numpy_path"] = rec["samples_out"] / rec["numpy_in_out_ms"] / 1e3
    out.append(rec)

    # ---- cfg2
    fs, x = G.wav("guitar")
    chain = (D.SignalChain(sample_rate_hz=fs).velvet_noise(duration_seconds=0.03, num_impulses=30, log_distribution_strength=1.0, seed=1)
             .haas_effect(delay_time_seconds=0.02, mode="LR"))
    xt = torch.from_numpy(x).to(dev)
    taps = O.class_taps(sample_rate_hz=fs, seed=1)

    def cpu_chain():
        return O.haas(O.vn_decorrelate(x, taps), sample_rate_hz=fs, delay_time_seconds=0.02)

    t_cpu = wall_ms(cpu_chain, reps=3, warm=1)
    rec = {"config": "cfg2 guitar SignalChain velvet_noise + haas_effect (fused)", "frames": int(x.shape[0]),
           "numpy_in_out_ms": wall_ms(lambda: chain(x)), "cuda_tensor_ms": cuda_ms(lambda: chain(xt)), "cpu_oracle_ms": t_cpu}
    out.append(rec)

    # ---- cfg4 (channel sub-slab of the 1024-channel config; the kernel is per channel)
    C4, L4 = 64, 57_600_000
    vn4 = D.VelvetNoise(sample_rate_hz=96000, duration_seconds=0.3, num_impulses=300, num_outs=C4, filtered_channels=tuple(range(C4)),
                        mode="LR", normalizer=None, seed=1)
    slab = torch.empty((C4, L4), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(4321)
    for c0 in range(0, C4, 8):
        slab[c0:c0 + 8].normal_(0.0, 1.0, generator=g).mul_(0.1)
    ms = cuda_ms(lambda: vn4.convolve(slab.t()), reps=5, warm=2)
    n_cpu = 200_000
    xs = np.ascontiguousarray(slab[:2, :n_cpu].cpu().numpy().T)
    t4 = O.class_taps(sample_rate_hz=96000, duration_seconds=0.3, num_impulses=300, num_outs=2, filtered_channels=(0, 1), seed=1)
    t0 = time.perf_counter()
    O.fir_class_order(xs, t4, O.DEFAULT_ENVELOPE, 2)
    cpu_s = time.perf_counter() - t0
    rec = {"config": "cfg4 300 impulses / 0.3 s @ 96 kHz, planar slab", "channels": C4, "frames": L4, "ms": ms,
           "Gsamples_per_s": C4 * L4 / ms / 1e6, "algorithmic_GBps": 8 * C4 * L4 / ms / 1e6,
           "cpu_oracle_Gsamples_per_s": 2 * n_cpu / cpu_s / 1e9, "cpu_sample": f"2 channels x {n_cpu} frames, 1 core"}
    out.append(rec)
    del slab
    torch.cuda.empty_cache()

    # ---- cfg5 grid stage: clips x 1024 candidates
    n_clips, frames, grid = 8, 1_440_000, 1024
    clips = np.stack([O.coloured_clip(i, frames).T for i in range(n_clips)]).astype(np.float32)
    kappas = np.linspace(0.0, 1.0, grid)
    t0 = time.perf_counter()
    tables = [T.generate_tap_table(sample_rate_hz=48000, duration_seconds=0.03, num_impulses=30, num_outs=2, num_segments=4,
                                   log_distribution_strength=float(k), filtered_channels=(0,), seed=1) for k in kappas]
    prog = T.candidate_program(tables, O.DEFAULT_ENVELOPE, frames)
    table_s = time.perf_counter() - t0
    ct = torch.from_numpy(clips).to(dev)
    ms = cuda_ms(lambda: OPT.vn_objective_partials(ct, prog), reps=3, warm=1)
    partials = OPT.vn_objective_partials(ct, prog).cpu().numpy()
    scores = OPT.vn_scores_from_partials(partials, angle_limit=np.pi / 4, lambda_mean=5.0, lambda_skew=2.0, lambda_correlation=15.0,
                                         lambda_penalty=1e3)
    # CPU oracle on a sub-sample: 1 clip x 4 candidates
    t0 = time.perf_counter()
    ref = O.vn_grid_scores(np.ascontiguousarray(clips[0].T), kappas[:4], sample_rate_hz=48000, duration_seconds=0.03, num_impulses=30, seed=1)
    cpu_s = time.perf_counter() - t0
    rec = {"config": "cfg5 optimize_velvet_noise grid stage", "clips": n_clips, "frames": frames, "candidates": grid, "ms": ms,
           "evaluations_per_s": n_clips * grid / ms * 1e3, "frame_evaluations_per_s": n_clips * grid * frames / ms * 1e3,
           "tap_tables_host_s": table_s, "cpu_oracle_evaluations_per_s": 4 / cpu_s, "cpu_sample": "1 clip x 4 candidates, 1 core",
           "max_abs_score_diff_vs_oracle_on_sample": float(np.max(np.abs(scores[0, :4].astype(np.float64) - ref.astype(np.float64)))),
           "argmin_clip0": int(np.argmin(scores[0]))}
    out.append(rec)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        for r in out:
            line = json.dumps(r)
            print(line)
            f.write(line + "\n")


if __name__ == "__main__":
    main()
