#!/usr/bin/env bash
# Rebuilds the library with different register-window kernel shapes and prints the bench value for
# each (run on the GPU box).  Usage: tools/sweep_window.sh "R W NW MINB [RUN]" ...
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  VND_EXTRA_DEFS="-DVND_WIN_R=$1 -DVND_WIN_W=$2 -DVND_WIN_NW=$3 -DVND_WIN_MINB=$4 -DVND_WIN_RUN=${5:-32}" \
    bash vndecorrelate_b200/csrc/build.sh > gpurun_out/sweep_build.log 2>&1 || { echo "$cfg: BUILD FAILED"; tail -3 gpurun_out/sweep_build.log; continue; }
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --channels-per-gpu 64 --e2e-channels 2 > gpurun_out/sweep_run.log 2>&1
  v=$(tail -1 gpurun_out/sweep_run.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f Gs/s frac %.3f clk %s parity %s' % (d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['config']['parity_spot_check'][:9]))" 2>/dev/null || tail -2 gpurun_out/sweep_run.log)
  echo "R W NW MINB RUN = $cfg : $v" | tee -a gpurun_out/sweep_results.txt
done
# leave the default build in place
bash vndecorrelate_b200/csrc/build.sh > /dev/null 2>&1
