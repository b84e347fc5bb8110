#!/usr/bin/env bash
# Sweeps the lane-quarter start-of-run delay of the tensor-memory kernel (run on the GPU box).
cd "$(dirname "$0")/.."
for ns in "$@"; do
  v=$(VND_TM_STAGGER_NS=$ns timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --channels-per-gpu ${CH:-148} --e2e-channels 2 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('%.1f Gs/s  frac %.3f  parity %s' % (d['value'], d['roofline']['frac'], d['config']['parity_spot_check'][:9]))")
  echo "stagger $ns ns: $v" | tee -a gpurun_out/stagger_sweep.txt
done
