#!/usr/bin/env bash
# Rebuilds with different extra defines and prints the short cfg3 bench value for each (GPU box).
cd "$(dirname "$0")/.."
for d in "$@"; do
  VND_EXTRA_DEFS="$d" bash vndecorrelate_b200/csrc/build.sh > /dev/null 2>&1 || { echo "$d: build failed"; continue; }
  v=$(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --channels-per-gpu ${CH:-64} --e2e-channels 2 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('%.1f Gs/s %s' % (d['value'], d['config']['parity_spot_check'][:9]))" 2>&1 | tail -1)
  echo "[$d] CH=${CH:-64}: $v" | tee -a gpurun_out/run_sweep.txt
done
bash vndecorrelate_b200/csrc/build.sh > /dev/null 2>&1
