#!/usr/bin/env bash
# Round 2, third session: `--set full` capture of the ring-buffer kernel (config 4) on a short configuration, after the same
# command exited 0 without ncu.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
small="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --channels-per-gpu 16 --frames 2000000 --e2e-channels 2 --configs 4 --cfg4-channels-per-gpu 16 --cfg4-frames 8000000"
$small > gpurun_out/r02_prof_plain_ring.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fir_ring -s 3 -c 1 -f -o gpurun_out/r02_fir_ring $small > gpurun_out/r02_prof_ncu_ring.log 2>&1
echo "fir_ring capture rc=$?"
ls -la gpurun_out/r02_fir_ring.ncu-rep
