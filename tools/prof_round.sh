#!/usr/bin/env bash
# Round-end evidence (run on the GPU box): launch list of the bench command and one full capture of the
# dominant kernel on a short configuration.  Usage: tools/prof_round.sh <tag>
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag="$1"
cmd="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-channels 2"
$cmd > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list rc=$?"
small="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --channels-per-gpu 148 --frames 6000000 --e2e-channels 2"
$small > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fir_tmem -s 1 -c 1 -f -o gpurun_out/${tag}_fir_tmem $small > gpurun_out/${tag}_ncu2.log 2>&1
echo "full capture rc=$?"
