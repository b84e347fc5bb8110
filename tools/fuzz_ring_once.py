"""One-off stress of the long-filter paths (ring-buffer kernel + tile tails) against the oracle on the GPU box."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import vnd_oracle as O
import vndecorrelate_b200.decorrelation as api
rng = np.random.default_rng(424242)
bad = 0
for i in range(40):
    fs = int(rng.choice([48000, 96000]))
    dur = float(rng.uniform(0.1, 0.36))
    n_imp = int(rng.integers(8, 301))
    env = tuple(float(v) for v in rng.choice([1.0, 0.85, 0.55, 0.35, 0.2, -0.5], size=int(rng.integers(1, 6))))
    Cn = int(rng.integers(1, 7))
    k = int(rng.integers(1, Cn + 1))
    frames = int(rng.integers(110000, 900000))
    kappa = float(rng.choice([0.0, 1.0, rng.uniform(0, 1)]))
    seed = int(rng.integers(0, 1000))
    vn = api.VelvetNoise(sample_rate_hz=fs, duration_seconds=dur, num_impulses=n_imp, num_outs=Cn, filtered_channels=tuple(range(k)), mode="LR",
                         normalizer=None, segment_envelope=env, log_distribution_strength=kappa, seed=seed)
    g = torch.Generator(device="cuda").manual_seed(i)
    slab = torch.randn((Cn, frames), generator=g, device="cuda") * 0.1
    y = vn.convolve(slab.t())
    taps = O.class_taps(sample_rate_hz=fs, duration_seconds=dur, num_impulses=n_imp, num_outs=Cn, num_segments=len(env), filtered_channels=tuple(range(k)),
                        log_distribution_strength=kappa, seed=seed)
    want = O.fir_class_order(np.ascontiguousarray(slab.cpu().numpy().T), taps, env, Cn)
    got = np.ascontiguousarray(y.cpu().numpy())
    ok = got.tobytes() == want.tobytes()
    if not ok:
        bad += 1
        b = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
        print("MISMATCH", i, fs, dur, n_imp, env, Cn, k, frames, kappa, seed, len(b), b[0], b[-1])
print("cases 40 mismatches", bad)
