#!/usr/bin/env python
"""Benchmark of the velvet-noise sparse-FIR hot path (BASELINE.json metric:
"Gsamples/s out (ch x samples) ... + % HBM roofline vs CPU numpy").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config 3 of BASELINE.json, the one the metric is quoted on): 4096 channels x 10 min @
48 kHz fp32 with one 30 ms / 30-impulse velvet filter per channel, partitioned by channel over the
8 GPUs of a box = 512 channels x 28.8 M samples per GPU (59 GB in + 59 GB out, resident in HBM).
Scaling is WEAK: every rank always holds one such 512-channel shard (N = 8 is the full config), the
tap table is generated once for all 4096 channels and sliced per rank, and there is no data-path
collective.  A step is one pass of the FIR kernel over the rank's whole shard.

`value` times the kernel with the shard resident in HBM (CUDA events on the launching stream);
`e2e` times the C-ABI call a user of the reference API makes with HOST buffers
(vnd_sparse_fir_stream_host: pinned host slab -> device -> pinned host slab, copies inside the timed
region); `cpu_baseline` / `--impl reference` time the numpy port of the reference's own loop
(oracle/vnd_oracle.py, pinned to the reference's golden vectors) on the host cores.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FS = 48000
FRAMES = 28_800_000  # 10 min @ 48 kHz
TOTAL_CHANNELS = 4096
CHANNELS_PER_GPU = 512
METRIC = "Gsamples/s out (ch x samples), velvet-noise sparse FIR, cfg3 shard per GPU"
UNIT = "Gsamples/s"
BYTES_PER_SAMPLE = 8  # algorithmic: 4 B read + 4 B written per output sample


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)), "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU arm: the numpy port of the reference loop (oracle), fanned out over host cores
# --------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, channels, frames = args
    from oracle import vnd_oracle as O

    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((frames, channels)) * 0.1).astype(np.float32)  # C-order (frames, channels): the reference's layout
    taps = O.class_taps(sample_rate_hz=FS, num_outs=channels, filtered_channels=tuple(range(channels)), seed=1)
    t0 = time.perf_counter()
    y = O.fir_class_order(x, taps, O.DEFAULT_ENVELOPE, channels)
    dt = time.perf_counter() - t0
    return dt, float(y[0, 0])


def cpu_rate(workers: int, channels_per_worker: int, frames: int, repeats: int = 1):
    """Gsamples/s of the numpy port with `workers` processes, best of `repeats`."""
    import multiprocessing as mp

    best = None
    if workers == 1:
        for r in range(repeats):
            dt, _ = _cpu_worker((100 + r, channels_per_worker, frames))  # the FIR alone, input generation excluded
            best = dt if best is None else min(best, dt)
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            for r in range(repeats):
                res = pool.map(_cpu_worker, [(100 + r * workers + w, channels_per_worker, frames) for w in range(workers)])
                dt = max(d for d, _ in res)  # slowest worker's FIR time (workers run concurrently)
                best = dt if best is None else min(best, dt)
    samples = workers * channels_per_worker * frames
    return samples / best / 1e9, best


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, cores)
    ch, frames = 4, 960_000  # 20 s @ 48 kHz per worker and step
    sample = f"{workers} processes x {ch} channels x {frames} frames (20 s @ 48 kHz), C-order (frames, channels) fp32, per step"
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_rate(workers, ch, frames)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_rate(workers, ch, frames)
        t_total += dt
    value = workers * ch * frames * args.steps / t_total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, note="CPU arm: bounded sample of the same workload"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, note: str = "") -> dict:
    cfg = {
        "workload": f"BASELINE configs[2]: {TOTAL_CHANNELS} ch x 10 min @ 48 kHz fp32, 30 ms / 30-impulse velvet FIR per channel; "
                    f"{args.channels_per_gpu} channels x {args.frames} frames per GPU (channel-sharded, planar)",
        "channels_per_gpu": args.channels_per_gpu, "frames": args.frames, "sample_rate_hz": FS, "num_impulses": 30, "duration_seconds": 0.03,
        "partitioning": "channels; one 4096-channel tap table generated once (seed 1) and sliced per rank; no collective",
        "l2": "inputs (tens of GB per GPU) are far larger than the 126 MB L2; no explicit flush",
    }
    if note:
        cfg["note"] = note
    return cfg


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import ctypes as C

    import torch
    import torch.distributed as dist

    from vndecorrelate_b200 import _native as N
    from vndecorrelate_b200 import runtime as R
    from vndecorrelate_b200.decorrelation import VelvetNoise

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = N.lib()

    # one table for the whole 4096-channel job, sliced per rank
    vn = VelvetNoise(sample_rate_hz=FS, duration_seconds=0.03, num_impulses=30, num_outs=TOTAL_CHANNELS,
                     filtered_channels=tuple(range(TOTAL_CHANNELS)), mode="LR", normalizer=None, seed=1)
    Cg, L = args.channels_per_gpu, args.frames
    c0 = (rank * Cg) % TOTAL_CHANNELS
    if c0 + Cg > TOTAL_CHANNELS:
        c0 = TOTAL_CHANNELS - Cg
    prog = vn.tap_program(L).slice_channels(c0, c0 + Cg)

    # device-resident synthetic shard: randn(seed 1234 + rank) * 0.1, generated in channel groups
    x = torch.empty((Cg, L), dtype=torch.float32, device=dev)
    y = torch.empty((Cg, L), dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    grp = max(1, min(Cg, (1 << 28) // max(L, 1)))
    for g0 in range(0, Cg, grp):
        g1 = min(Cg, g0 + grp)
        x[g0:g1].normal_(0.0, 1.0, generator=gen).mul_(0.1)
    torch.cuda.synchronize()

    sx = R.torch_signal(x.t())
    sy = R.torch_signal(y.t())
    ps = R.device_program(prog, dev)
    stream = torch.cuda.current_stream(dev)

    def step():
        N.check(lib.vnd_sparse_fir_dev(C.byref(sx), C.byref(sy), C.byref(ps), stream.cuda_stream), "vnd_sparse_fir_dev")

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # parity spot check (outside the timed region): 2 channels x 2^16 frames against the oracle
    parity = "skipped"
    if rank == 0:
        from oracle import vnd_oracle as O

        taps = O.class_taps(sample_rate_hz=FS, num_outs=TOTAL_CHANNELS, filtered_channels=tuple(range(TOTAL_CHANNELS)), seed=1)
        n = min(L, 1 << 16)
        halo = prog.halo
        ok = True
        for ch in (0, Cg - 1):
            m = min(L, n + halo)
            col = x[ch, :m].cpu().numpy()
            want = O.fir_class_order(np.stack((col, col), axis=1), [taps[c0 + ch], []], O.DEFAULT_ENVELOPE, 2)[:, 0]
            valid = m if m == L else n  # outputs whose taps all lie inside the excerpt
            got = y[ch, :valid].cpu().numpy()
            ok = ok and want[:valid].tobytes() == got.tobytes()
        parity = "bit-exact vs oracle on 2 channels x %d frames" % n if ok else "MISMATCH"
        if not ok:
            raise SystemExit("bench: GPU output differs from the oracle; refusing to report a number")

    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = N.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    launches = N.launch_count() - launches0
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    t = torch.tensor([total_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms = float(tmax[0])
        launches = int(t[1])
    samples_per_step = world * Cg * L
    value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e9

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / mean launch duration
    kernel_ms = float(np.mean(per_step))
    achieved = BYTES_PER_SAMPLE * Cg * L / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "fir_tile_traffic.json" if os.environ.get("VND_DISABLE_TMEM") == "1" else "fir_tmem_traffic.json")
    if os.path.exists(tpath):
        try:
            tr = json.load(open(tpath))
            traffic = tr["dram_bytes_per_sample"] * Cg * L
        except Exception:
            traffic = None
    if os.environ.get("VND_DISABLE_TMEM") != "1":
        kernel_name = ("fir_tmem_kernel (tensor-memory Hankel rows + tcgen05.ld gathers, FADD2, warp-specialized, TMA bulk copies); "
                       "the last ~1.5 tiles of every channel run on fir_tile_kernel in a second launch inside the same step")
    elif os.environ.get("VND_DISABLE_WINDOW") != "1":
        kernel_name = "fir_window_kernel<R=29,W=32> (register-window, persistent, TMA double-buffered)"
    else:
        kernel_name = "fir_tile_kernel<float,SEGMENTED>"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kernel_name, "kernel_ms": kernel_ms, "peak_source": peak_src,
                "note": "HBM is not the binding resource: every (output, tap) pair moves one word from on-chip memory to a register and costs one add; the kernel is bound by the shared-memory pipe and on-chip latency (ncu: shared-memory pipe ~67 % busy, issue slots ~57 %, DRAM ~35 %). See DESIGN.md section 4 and profiles/r01_summary.md"}

    # end to end through the C ABI with HOST buffers (copies inside the timed region)
    e2e = None
    try:
        Ce = min(args.e2e_channels, Cg)
        hx, hy = R.PinnedArray((Ce, L)), R.PinnedArray((Ce, L))
        hx.array[...] = x[:Ce].cpu().numpy()
        sub = prog.slice_channels(0, Ce)
        hs = sub.host_struct()
        ctx = R.HostContext.get(local_rank)
        chunk = max(1, Ce // max(1, args.e2e_chunks))  # channels per pipeline stage: more stages = shorter fill and drain

        def e2e_step():
            N.check(lib.vnd_sparse_fir_stream_host(ctx.handle, hx.array.ctypes.data, hy.array.ctypes.data, L, Ce, C.byref(hs), chunk), "vnd_sparse_fir_stream_host")

        e2e_step()
        if world > 1:
            dist.barrier()
        e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        same = bool(np.array_equal(hy.array[0, : 1 << 16], y[0, : 1 << 16].cpu().numpy()))
        e2e = {"value": world * Ce * L * e_steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": Ce * L * 4, "d2h_bytes_per_step": Ce * L * 4,
               "steps": e_steps, "channels": Ce, "frames": L, "api": "vnd_sparse_fir_stream_host (pinned host slabs, 3-stream pipeline)",
               "matches_device_path": same}
    except Exception as exc:  # report, do not hide
        e2e = {"value": None, "unit": UNIT, "error": str(exc)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ch, fr = 64, 960_000  # ~20 s of CPU work on one core; the strided column access of the C-order layout gets slower with more channels
        rate, secs = cpu_rate(1, ch, fr, repeats=1)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"{ch} channels x {fr} frames (20 s @ 48 kHz) C-order (frames, channels) fp32, numpy port of VelvetNoise.convolve, one pass ({secs:.1f} s)"}

    if rank == 0:
        cfg = workload_config(args)
        cfg["parity_spot_check"] = parity
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--channels-per-gpu", type=int, default=CHANNELS_PER_GPU)
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--e2e-channels", type=int, default=32)
    ap.add_argument("--e2e-chunks", type=int, default=32, help="pipeline stages of the end-to-end call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
