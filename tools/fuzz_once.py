import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import vnd_oracle as O
from tests import _random_cases as RC
import vndecorrelate_b200.decorrelation as api
bad = 0; n = 0; errs = 0
for seed in range(1, 9):
    rng = np.random.default_rng(7000 + seed)
    for i in range(150):
        p, x = RC.random_vn_case(rng)
        try:
            vn = api.VelvetNoise(**RC.vn_kwargs(p)); got = vn.decorrelate(x)
        except ValueError:
            errs += 1; continue
        want = RC.oracle_output(O, p, x); n += 1
        if got.dtype != want.dtype or got.shape != want.shape or got.tobytes() != np.ascontiguousarray(want).tobytes():
            bad += 1; print("MISMATCH decorrelate", seed, i, p, x.shape, x.dtype)
    for fam, fn, orc in (("convolve", lambda p, x: api.VelvetNoise(**RC.convolve_kwargs(p)).convolve(x), RC.oracle_convolve),
                         ("chain", lambda p, x: RC.run_chain(api.SignalChain, p, x), RC.oracle_chain),
                         ("function", lambda p, x: api.convolve_velvet_noise(x, api.generate_velvet_noise(**RC.function_kwargs(p))), RC.oracle_function)):
        gen = RC.FAMILIES[fam][0]
        for i in range(60):
            p, x = gen(rng)
            try:
                got = fn(p, x)
            except ValueError:
                errs += 1; continue
            want = orc(O, p, x); n += 1
            if got.dtype != want.dtype or got.shape != want.shape or np.ascontiguousarray(got).tobytes() != np.ascontiguousarray(want).tobytes():
                bad += 1; print("MISMATCH", fam, seed, i, p, x.shape, x.dtype)
print("cases", n, "rejected (ValueError)", errs, "mismatches", bad)
