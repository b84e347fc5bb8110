#!/usr/bin/env bash
# Round 2: the driver's sequence on one GPU - gpu tests, smoke, bench (both arms) - outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_r02.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02.json").read().strip().splitlines()[-1])
    print("cfg3 %.1f Gs/s frac %.3f  e2e %s  parity %s" % (d["value"], d["roofline"]["frac"], json.dumps(d["e2e"])[:400], d["parity"][:60]))
    for k, v in d["configs"].items():
        print(k, json.dumps(v)[:700])
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
PY
if [ -n "$REFARM" ]; then echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json; tail -c 1500 gpurun_out/bench_r02_reference.json; fi
