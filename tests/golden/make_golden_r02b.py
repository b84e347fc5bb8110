"""Round 2, third session: hashes of the UNMODIFIED reference's ``VelvetNoise.decorrelate`` on the 120 seeded random cases
of ``tests/_random_cases.py`` and on its three further families of 60 cases each - multichannel ``convolve`` in C and
Fortran order, ``SignalChain`` velvet noise + Haas, ``convolve_velvet_noise(generate_velvet_noise(...))`` - (build container
only: reads /root/reference):

    python tests/golden/make_golden_r02b.py        ->  tests/golden/random_decorrelate.json

Nothing here is imported by the product."""

from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from vndecorrelate.decorrelation import SignalChain, VelvetNoise, convolve_velvet_noise, generate_velvet_noise  # noqa: E402

from tests import _random_cases as RC  # noqa: E402

out = []
for i, p, x in RC.cases():
    kw = RC.vn_kwargs(p)
    try:
        y = VelvetNoise(**kw).decorrelate(x.copy())
        y = np.ascontiguousarray(y)
        out.append(dict(id=i, dtype=str(y.dtype), shape=list(y.shape), sha256=hashlib.sha256(y.tobytes()).hexdigest()))
    except Exception as e:  # the reference's own error is the contract
        out.append(dict(id=i, error=type(e).__name__))
print(len(out), "decorrelate cases;", sum("error" in c for c in out), "raise in the reference")


def record(fn):
    import contextlib
    import io

    try:
        with contextlib.redirect_stdout(io.StringIO()):
            y = np.ascontiguousarray(fn())
        return dict(dtype=str(y.dtype), shape=list(y.shape), sha256=hashlib.sha256(y.tobytes()).hexdigest())
    except Exception as e:  # the reference's own error is the contract
        return dict(error=type(e).__name__)


more = {}
for fam in RC.FAMILIES:
    rows = []
    for i, p, x in RC.more_cases(fam):
        if fam == "convolve":
            r = record(lambda: VelvetNoise(**RC.convolve_kwargs(p)).convolve(x.copy(order="K")))
        elif fam == "chain":
            r = record(lambda: RC.run_chain(SignalChain, p, x.copy()))
        else:
            r = record(lambda: convolve_velvet_noise(x.copy(), generate_velvet_noise(**RC.function_kwargs(p))))
        r["id"] = i
        rows.append(r)
    more[fam] = rows
    print(fam, len(rows), "cases;", sum("error" in c for c in rows), "raise in the reference:", sorted({c["error"] for c in rows if "error" in c}))
json.dump(dict(seed=RC.SEED, count=RC.COUNT, cases=out, count_more=RC.COUNT_MORE, more=more), open(os.path.join(HERE, "random_decorrelate.json"), "w"), indent=0)
