#!/usr/bin/env bash
# Round 2, third session: parity of the ring-buffer kernel for long filters and the config-4 bench with and without it.
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring_kernel or cfg4 or direct_kernel" 2>&1 | tail -4
for dis in 0 1; do
  echo "== VND_DISABLE_RING=$dis"
  VND_DISABLE_RING=$dis timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --channels-per-gpu 16 --frames 2000000 --e2e-channels 2 --configs 4 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['configs']['cfg4']; print(json.dumps({k: c.get(k) for k in ('value', 'ms', 'lsu_pipe', 'parity')})[:400])"
done
