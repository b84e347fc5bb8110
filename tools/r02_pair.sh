#!/usr/bin/env bash
# Round 2: paired-loop variants of the tensor-memory FIR (VND_TM_SHAPE 5-8) against the default shape, tight timeouts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for s in ${SHAPES:-0 5 6 7 8}; do
  echo "== VND_TM_SHAPE=$s"
  VND_TM_SHAPE=$s CH=${CH:-148} bash tools/try_fir.sh 2>&1 | tail -4
done | tee gpurun_out/r02_pair.txt
