#!/usr/bin/env bash
# Timing probes of the tensor-memory kernel (results wrong by construction): -DVND_TM_PROBE=n builds with the timeline
# probes into vndecorrelate_b200/_lib_probe<n>/ (1: two x16 loads of the same columns per tap, i.e. half the distinct
# bytes but the same instruction count as two halves; 2: a quarter of the adds per tensor-memory tap; 3: one x16 load per tap; 4: tap columns computed instead of read from the operation words; 5: far taps without the block-crossing path; 6: every far tap as an aligned one (packed adds, eight loads); 7: both).
set -euo pipefail
n="$1"
cd "$(dirname "$0")/../vndecorrelate_b200/csrc"
mkdir -p ../_lib_probe$n
sed "s#out=\"\$here/../_lib\"#out=\"\$here/../_lib_probe$n\"#" build.sh > ./.build_probe_tmp.sh
VND_EXTRA_DEFS="-DVND_TM_TRACE -DVND_TM_PROBE=$n" bash ./.build_probe_tmp.sh
rm -f ./.build_probe_tmp.sh
