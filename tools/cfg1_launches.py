#!/usr/bin/env python
"""cfg1 / cfg2 on CUDA tensors, three calls each (run under `ncu --metrics gpu__time_duration.sum` for the launch timeline)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import _golden as G  # noqa: E402
from vndecorrelate_b200 import decorrelation as D  # noqa: E402

fs, x = G.wav("viola")
vn = D.VelvetNoise(sample_rate_hz=fs, duration_seconds=0.03, num_impulses=30, seed=1)
xt = torch.from_numpy(x).cuda()
for _ in range(3):
    y = vn.decorrelate(xt)
fs, g = G.wav("guitar")
chain = D.SignalChain(sample_rate_hz=fs).velvet_noise(duration_seconds=0.03, num_impulses=30, log_distribution_strength=1.0, seed=1).haas_effect(
    delay_time_seconds=0.02, mode="LR")
gt = torch.from_numpy(g).cuda()
for _ in range(3):
    z = chain(gt)
torch.cuda.synchronize()
print("ok", tuple(y.shape), tuple(z.shape))
