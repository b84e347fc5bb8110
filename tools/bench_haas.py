#!/usr/bin/env python
"""optimize_haas_delay's grid stage alone (run on a GPU box): clips x delays through haas_objective_kernel.

    python tools/bench_haas.py [--clips 8] [--frames 1440000] [--delays 1024]

Prints one JSON line: ms per sweep, (clip, delay) evaluations/s, frame-evaluations/s."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vndecorrelate_b200 import optimization as OPT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--frames", type=int, default=1_440_000)
ap.add_argument("--delays", type=int, default=1024)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
rng = np.random.default_rng(0)
clips = (rng.standard_normal((a.clips, 2, a.frames)) * 0.1).astype(np.float32)
delays = np.round(np.linspace(0.0, 0.03, a.delays) * 48000).astype(np.int32)
ct = torch.from_numpy(clips).cuda()
OPT.haas_objective_partials(ct, delays)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
ev[0].record()
for i in range(a.reps):
    OPT.haas_objective_partials(ct, delays)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps)]))
print(json.dumps({"clips": a.clips, "frames": a.frames, "delays": a.delays, "ms": ms, "evaluations_per_s": a.clips * a.delays / ms * 1e3,
                  "frame_evaluations_per_s": a.clips * a.delays * a.frames / ms * 1e3}))
