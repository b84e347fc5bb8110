"""Multi-GPU partitioning of the hot path (one process per GPU, ``torch.distributed``).

Channels, clips and optimisation candidates are independent, so the work is partitioned with no
data-path collective:

* multichannel slabs — contiguous channel blocks per rank, each rank using the matching SLICE of
  the single host-generated tap table (never regenerated per shard: the RNG draw layout depends on
  the total channel count, reference ``decorrelation.py:510-520``);
* the ``optimize_velvet_noise`` sweep over many clips — clips (or candidates) are split over
  ranks; the one exchange is an all-gather of the float32 score matrix (n_clips x grid, 256 KB for
  64 x 1024) so that every rank can run ``get_local_minima`` / argmin on complete rows
  (``optimization.py:120-128`` needs both grid neighbours of every point).

The collective goes through ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests,
where the scoring function is injected).
"""

from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from .taps import TapProgram


def block_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block ``[start, stop)`` of ``total`` items for ``rank``; the first ``total % world``
    ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_program(program: TapProgram, rank: int, world: int) -> tuple[TapProgram, int, int]:
    """The rank's slice of a full tap program and its channel range."""
    start, stop = block_range(program.channels, rank, world)
    return program.slice_channels(start, stop), start, stop


def _dist():
    import torch.distributed as dist

    return dist


def _world(group=None) -> tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def collective_device(group=None):
    """Where a collective of ``group`` must run: the current CUDA device for NCCL, ``None`` (CPU) for gloo."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl":
        import torch

        return torch.device("cuda", torch.cuda.current_device())
    return None


def all_gather_rows(local_rows: np.ndarray, counts: Sequence[int], *, device=None, group=None) -> np.ndarray:
    """Concatenate per-rank row blocks (rank r contributes ``counts[r]`` rows) on every rank.

    ONE collective on ONE tensor: the blocks are padded to the longest and gathered with
    ``all_gather_into_tensor`` into a single ``(world * longest, ...)`` buffer (no list of per-rank tensors, one
    device-to-host copy); ``device`` selects where it runs (a CUDA device for NCCL, ``None``/cpu for gloo)."""
    import torch

    dist = _dist()
    rank, world = _world(group)
    if world == 1:
        return np.ascontiguousarray(local_rows)
    width = tuple(local_rows.shape[1:])
    longest = max(counts)
    pad = np.zeros((longest, *width), dtype=local_rows.dtype)
    pad[: local_rows.shape[0]] = local_rows
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device, non_blocking=True)
    out = torch.empty((world * longest, *width), dtype=t.dtype, device=t.device)  # rank blocks concatenated along dim 0
    dist.all_gather_into_tensor(out, t, group=group)
    full = out.cpu().numpy()
    if all(c == longest for c in counts):
        return full
    return np.concatenate([full[r * longest: r * longest + counts[r]] for r in range(world)], axis=0)


def sweep_scores_sharded(
    clips,
    program: TapProgram,
    score_fn: Callable[[object, TapProgram], np.ndarray],
    *,
    by: str = "clips",
    device=None,
    group=None,
) -> np.ndarray:
    """Scores ``(n_clips, n_candidates)`` of a sweep, computed shard-wise and all-gathered.

    ``clips``: ``(n_clips, 2, frames)`` (numpy or CUDA tensor, identical on every rank);
    ``score_fn(clips_subset, program_subset)`` returns the scores of a sub-problem — on GPUs this is
    ``vn_scores_from_partials(vn_objective_partials(...))``.  ``by`` picks the partitioned axis."""
    rank, world = _world(group)
    n_clips = clips.shape[0]
    n_cand = program.channels
    if by == "clips":
        counts = [block_range(n_clips, r, world)[1] - block_range(n_clips, r, world)[0] for r in range(world)]
        lo, hi = block_range(n_clips, rank, world)
        local = score_fn(clips[lo:hi], program) if hi > lo else np.zeros((0, n_cand), dtype=np.float32)
        return all_gather_rows(np.ascontiguousarray(local), counts, device=device, group=group)
    if by == "candidates":
        counts = [block_range(n_cand, r, world)[1] - block_range(n_cand, r, world)[0] for r in range(world)]
        lo, hi = block_range(n_cand, rank, world)
        local = score_fn(clips, program.slice_channels(lo, hi)) if hi > lo else np.zeros((n_clips, 0), dtype=np.float32)
        gathered = all_gather_rows(np.ascontiguousarray(local.T), counts, device=device, group=group)  # rows = candidates
        return np.ascontiguousarray(gathered.T)
    raise ValueError("by must be 'clips' or 'candidates'")


def select_from_scores(scores: np.ndarray):
    """Per clip: argmin index and the local-minima set of its score row (``optimization.py:120-128``)."""
    from .optimization import get_local_minima

    grid = scores.shape[1]
    return [int(np.argmin(row)) for row in scores], [get_local_minima(row, grid) for row in scores]
