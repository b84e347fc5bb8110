#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tensor-memory objective variants"; timeout 600 python -m pytest tests -q -m gpu -k "tensor_memory_variants or cfg5" 2>&1 | tail -3
run() { echo -n "$1: "; shift; env "$@" timeout 300 python tools/bench_objective.py --clips 16 --reps 3 --check 2>&1 | tail -1; }
run "shared-memory kernel        " A=1
run "tmem 16 frames/lane x 20 warps" VND_OBJ_TMEM=1
run "tmem 32 frames/lane x 12 warps" VND_OBJ_TMEM=2
small="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --channels-per-gpu 16 --frames 2000000 --e2e-channels 2 --configs 4 --cfg4-channels-per-gpu 16 --cfg4-frames 8000000"
$small > gpurun_out/r02_prof_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fir_tile_kernel.*1024" -s 3 -c 1 -f -o gpurun_out/r02_fir_tile_long $small > gpurun_out/r02_prof_ncu5.log 2>&1; echo "fir_tile long capture rc=$?"; tail -2 gpurun_out/r02_prof_ncu5.log | cut -c1-200
