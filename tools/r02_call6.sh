#!/usr/bin/env bash
# Round 2, call 6: objective kernel after the moment diet; cfg3 A/B with the round-1 kernel FILE inside this tree; Haas
# objective at scale; the RMS path on a 10-minute stereo file; full bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== objective + dsp tests"; timeout 900 python -m pytest tests -q -m gpu -k "objective or sweeps or optimis or cfg5 or batch or polar or normalis" 2>&1 | tail -4
run() { echo -n "$1: "; shift; env "$@" timeout 300 python tools/bench_objective.py --clips 16 --reps 3 --check 2>&1 | tail -1; }
run "smem objective kernel" A=1
for w in 608 704; do :; done
b() { timeout 300 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-channels 2 --configs "" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('%.1f Gs/s  frac %.3f' % (d['value'], d['roofline']['frac']))"; }
echo -n "cfg3 this tree:            "; b A=1
echo -n "cfg3 round-1 kernel file:  "; b VND_B200_LIB=$PWD/vndecorrelate_b200/_lib/libvnd_b200_r1tm.so
echo -n "cfg3 this tree:            "; b A=1
echo -n "cfg3 round-1 kernel file:  "; b VND_B200_LIB=$PWD/vndecorrelate_b200/_lib/libvnd_b200_r1tm.so
echo "== haas objective, 8 clips x 1024 delays"; timeout 300 python tools/bench_haas.py --clips 8 2>&1 | tail -1
echo "== 10-minute stereo file, RMS path"; timeout 300 python tools/long_stereo.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_long_stereo_launches.csv python tools/long_stereo.py > gpurun_out/r02_long_stereo_ncu.log 2>&1; echo "ncu rc=$?"
echo "== full bench"; STEPS=20 bash tools/r02_full.sh 2>&1 | tail -14
