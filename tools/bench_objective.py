#!/usr/bin/env python
"""cfg5 grid stage alone (run on a GPU box): clips x candidates through vn_objective_kernel.

    python tools/bench_objective.py [--clips 2] [--frames 1440000] [--candidates 1024] [--reps 3] [--check]

Prints one JSON line: ms per sweep, (clip, candidate) evaluations/s, frame-evaluations/s; with --check
also the largest |score difference| against the CPU oracle on clip 0 x 4 candidates and the argmin."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import vnd_oracle as O  # noqa: E402
from vndecorrelate_b200 import optimization as OPT  # noqa: E402
from vndecorrelate_b200 import taps as T  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=2)
ap.add_argument("--frames", type=int, default=1_440_000)
ap.add_argument("--candidates", type=int, default=1024)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()

clips = np.stack([O.coloured_clip(i, a.frames).T for i in range(a.clips)]).astype(np.float32)
kappas = np.linspace(0.0, 1.0, a.candidates)
tables = [T.generate_tap_table(sample_rate_hz=48000, duration_seconds=0.03, num_impulses=30, num_outs=2, num_segments=4,
                               log_distribution_strength=float(k), filtered_channels=(0,), seed=1) for k in kappas]
prog = T.candidate_program(tables, O.DEFAULT_ENVELOPE, a.frames)
ct = torch.from_numpy(clips).cuda()
OPT.vn_objective_partials(ct, prog)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
ev[0].record()
for i in range(a.reps):
    OPT.vn_objective_partials(ct, prog)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps)]))
rec = {"clips": a.clips, "frames": a.frames, "candidates": a.candidates, "ms": ms, "evaluations_per_s": a.clips * a.candidates / ms * 1e3,
       "frame_evaluations_per_s": a.clips * a.candidates * a.frames / ms * 1e3}
if a.check:
    partials = OPT.vn_objective_partials(ct, prog).cpu().numpy()
    scores = OPT.vn_scores_from_partials(partials, angle_limit=np.pi / 4, lambda_mean=5.0, lambda_skew=2.0, lambda_correlation=15.0, lambda_penalty=1e3)
    n = min(4, a.candidates)
    ref = O.vn_grid_scores(np.ascontiguousarray(clips[0].T), kappas[:n], sample_rate_hz=48000, duration_seconds=0.03, num_impulses=30, seed=1)
    rec["max_abs_score_diff_vs_oracle"] = float(np.max(np.abs(scores[0, :n].astype(np.float64) - ref.astype(np.float64))))
    rec["argmin_clip0"] = int(np.argmin(scores[0]))
print(json.dumps(rec))
