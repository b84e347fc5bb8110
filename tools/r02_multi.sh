#!/usr/bin/env bash
# Round 2: the driver's multi-GPU launch of bench.py on N GPUs of one box (N from $1), plus the reference arm the same way.
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_r02_n$N.err | tail -12
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_r02_n{n}.json").read().strip().splitlines()[-1])
    print("N=%s cfg3 %.1f Gs/s (%.1f per GPU) frac %.3f  e2e %.2f (ceiling %.2f)" % (n, d["value"], d["value"] / d["n_gpus"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["copy_ceiling"]["value"]))
    for k, v in d["configs"].items():
        print(k, v.get("value"), v.get("unit"), "ms", v.get("ms"), "e2e", (v.get("e2e") or {}).get("value"), v.get("collective", ""), v.get("scores_sha256", "")[:16], v.get("error", ""))
except Exception as e:
    print("bench parse failed", e)
PY
