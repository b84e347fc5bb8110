#!/usr/bin/env bash
# Round-2 evidence (run on the GPU box, one GPU): launch list of the bench command, then one `--set full` capture of
# each dominant kernel (cfg3 fir_tmem_kernel, cfg4 fir_tile_kernel, cfg5 vn_objective_kernel) on short configurations.
# Every ncu command runs only after the same command has exited 0 without ncu.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cfg5-clips 4"
$cmd > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv $cmd > gpurun_out/r02_prof_ncu1.log 2>&1
echo "launch list rc=$?"
small="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --channels-per-gpu 148 --frames 6000000 --e2e-channels 2 --configs 4 --cfg4-channels-per-gpu 16 --cfg4-frames 8000000"
$small > gpurun_out/r02_prof_plain2.log 2>&1 && {
ncu --set full --clock-control none --import-source on -k regex:fir_tmem -s 3 -c 1 -f -o gpurun_out/r02_fir_tmem $small > gpurun_out/r02_prof_ncu2.log 2>&1; echo "fir_tmem capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fir_tile -s 3 -c 1 -f -o gpurun_out/r02_fir_tile_long $small > gpurun_out/r02_prof_ncu3.log 2>&1; echo "fir_tile capture rc=$?"; }
obj="python tools/bench_objective.py --clips 2 --reps 1"
$obj > gpurun_out/r02_prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vn_objective -s 1 -c 1 -f -o gpurun_out/r02_vn_objective $obj > gpurun_out/r02_prof_ncu4.log 2>&1
echo "vn_objective capture rc=$?"
ls -la gpurun_out/*.ncu-rep
