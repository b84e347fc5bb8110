#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== optimiser tests (two batches in flight)"; timeout 900 python -m pytest tests -q -m gpu -k "optimis or cfg5 or batch or sweeps" 2>&1 | tail -3
small="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --channels-per-gpu 16 --frames 2000000 --e2e-channels 2 --configs 4 --cfg4-channels-per-gpu 16 --cfg4-frames 8000000"
$small > gpurun_out/r02_prof_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fir_tile_kernel<float, 0, 1024" -s 3 -c 1 -f -o gpurun_out/r02_fir_tile_long $small > gpurun_out/r02_prof_ncu5.log 2>&1; echo "fir_tile long capture rc=$?"; tail -3 gpurun_out/r02_prof_ncu5.log | cut -c1-200
STEPS=20 bash tools/r02_full.sh 2>&1 | tail -12 | cut -c1-700
