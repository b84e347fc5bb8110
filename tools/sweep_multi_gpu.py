#!/usr/bin/env python
"""BASELINE config 5 on N GPUs: the optimize_velvet_noise grid stage for many clips, sharded by clips
(or candidates) with one NCCL all-gather of the score matrix.  Launch with

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_multi_gpu.py

Every rank ends up with the full (clips x candidates) score matrix and so with the same argmin and
local-minima set per clip (optimization.py:120-128 needs both grid neighbours of every point).  Rank 0
prints one JSON line; with N == 1 the line is the single-GPU reference the others are compared with
through the checksum.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import vnd_oracle as O  # noqa: E402  (clip generator only: the recipe of SURVEY.md section 8d)
from vndecorrelate_b200 import optimization as OPT  # noqa: E402
from vndecorrelate_b200 import sharding as S  # noqa: E402
from vndecorrelate_b200 import taps as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=16)
    ap.add_argument("--frames", type=int, default=1_440_000)
    ap.add_argument("--grid", type=int, default=1024)
    ap.add_argument("--by", choices=["clips", "candidates"], default="clips")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kw = dict(angle_limit=np.pi / 4, lambda_mean=5.0, lambda_skew=2.0, lambda_correlation=15.0, lambda_penalty=1e3)
    clips = np.stack([O.coloured_clip(i, args.frames).T for i in range(args.clips)]).astype(np.float32)  # identical on every rank
    kappas = np.linspace(0.0, 1.0, args.grid)
    tables = [T.generate_tap_table(sample_rate_hz=48000, duration_seconds=0.03, num_impulses=30, num_outs=2, num_segments=4,
                                   log_distribution_strength=float(k), filtered_channels=(0,), seed=1) for k in kappas]
    prog = T.candidate_program(tables, O.DEFAULT_ENVELOPE, args.frames)
    clips_dev = torch.from_numpy(clips).to(dev)

    def score_fn(sub_clips, sub_prog):
        p = OPT.vn_objective_partials(sub_clips, sub_prog)
        return OPT.vn_scores_from_partials(p.cpu().numpy(), **kw)

    # warm-up, then one timed sweep (device time of the local shard + the gather, max over ranks)
    S.sweep_scores_sharded(clips_dev, prog, score_fn, by=args.by, device=dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    scores = S.sweep_scores_sharded(clips_dev, prog, score_fn, by=args.by, device=dev)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    argmins, minima = S.select_from_scores(scores)
    digest = hashlib.sha256(np.ascontiguousarray(scores).tobytes()).hexdigest()
    same = torch.tensor([int(digest[:12], 16)], dtype=torch.int64, device=dev)
    if world > 1:
        lo, hi = same.clone(), same.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = bool(lo.item() == hi.item())
    else:
        identical = True
    if rank == 0:
        print(json.dumps({
            "config": "cfg5 grid stage", "n_gpus": world, "by": args.by, "clips": args.clips, "frames": args.frames, "candidates": args.grid,
            "ms": float(ms.item()), "evaluations_per_s": args.clips * args.grid / float(ms.item()) * 1e3,
            "scores_sha256": digest, "all_ranks_hold_identical_scores": identical, "argmin_per_clip": argmins,
            "local_minima_per_clip": [len(m) for m in minima],
            "collective": "one all_gather of the float32 score rows (NCCL)" if world > 1 else "none"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
