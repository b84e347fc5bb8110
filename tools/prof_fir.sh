#!/usr/bin/env bash
# Profiles the dominant FIR kernel of a short bench run (run on the GPU box).
# Usage: tools/prof_fir.sh <kernel-regex> <output-name> [extra bench args]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
pat="$1"; name="$2"; shift 2
cmd="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --channels-per-gpu 16 --frames 4000000 --e2e-channels 2 $*"
$cmd > gpurun_out/${name}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -f -o gpurun_out/$name $cmd > gpurun_out/${name}_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/${name}_ncu.log
