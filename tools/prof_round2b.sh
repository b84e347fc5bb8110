#!/usr/bin/env bash
# Round 2, third session: refreshed launch list of the bench command and `--set full` captures of the dominant kernels,
# plus the Haas objective kernel (new this session).  Every ncu command runs only after the same command exited 0 without ncu.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/prof_round2.sh
haas="python tools/bench_haas.py --clips 2 --reps 1"
$haas > gpurun_out/r02_prof_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:haas_objective -s 1 -c 1 -f -o gpurun_out/r02_haas_objective $haas > gpurun_out/r02_prof_ncu5.log 2>&1
echo "haas_objective capture rc=$?"
ls -la gpurun_out/*.ncu-rep
