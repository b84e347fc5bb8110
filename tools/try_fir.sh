#!/usr/bin/env bash
# Parity of the tensor-memory kernel + a short cfg3 bench with TIGHT timeouts (experimental builds may
# deadlock; a hung kernel must not eat the GPU budget).  Usage: [CH=64] tools/try_fir.sh
cd "$(dirname "$0")/.."
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tmem or size_independent" 2>&1 | tail -3
timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --channels-per-gpu ${CH:-64} --e2e-channels 2 --configs "" 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('%.1f Gs/s  frac %.3f  clk %s  parity %s' % (d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['parity'][:9]))
except Exception as e:
    print('bench failed', e)"
