#!/usr/bin/env bash
# Debug build of the library with the tensor-memory kernel's timeline probes (-DVND_TM_TRACE) into
# vndecorrelate_b200/_lib_trace/; use it with VND_B200_LIB=$PWD/vndecorrelate_b200/_lib_trace/libvnd_b200.so python tools/trace_tmem.py
set -euo pipefail
cd "$(dirname "$0")/../vndecorrelate_b200/csrc"
mkdir -p ../_lib_trace
sed 's#out="$here/../_lib"#out="$here/../_lib_trace"#' build.sh > ./.build_trace_tmp.sh
VND_EXTRA_DEFS=-DVND_TM_TRACE bash ./.build_trace_tmp.sh
rm -f ./.build_trace_tmp.sh
