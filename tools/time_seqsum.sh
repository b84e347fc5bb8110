#!/usr/bin/env bash
# Kernel time of the numpy-order sum of squares on cfg1 / cfg2 (run on the GPU box): ncu launch list of tools/cfg1_launches.py.
cd "$(dirname "$0")/.."
timeout 100 python tools/cfg1_launches.py > /dev/null 2>&1 || { echo "cfg1_launches failed"; exit 1; }
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/seqsum_launches.csv python tools/cfg1_launches.py > /dev/null 2>&1
grep seq_sumsq gpurun_out/seqsum_launches.csv | awk -F'","' '{gsub(/"/,"",$NF); print $5, $NF " ns"}' | sort | uniq -c
