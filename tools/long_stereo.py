#!/usr/bin/env python
"""VelvetNoise.decorrelate (default RMS normaliser) on a 10-minute 44.1 kHz stereo signal resident on the device: run under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` for the per-kernel
time and DRAM traffic of the RMS path (launch list), or plain for the wall time of one call."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vndecorrelate_b200.decorrelation import VelvetNoise  # noqa: E402

frames = 600 * 44100
x = torch.randn((frames, 2), device="cuda") * 0.1
vn = VelvetNoise(sample_rate_hz=44100, duration_seconds=0.03, num_impulses=30, seed=1)
vn.decorrelate(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
y = vn.decorrelate(x)
torch.cuda.synchronize()
print("frames", frames, "ms", (time.perf_counter() - t0) * 1e3, "algorithmic MB", frames * 2 * 8 / 1e6)
