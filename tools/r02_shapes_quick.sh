#!/usr/bin/env bash
# Round 2: short cfg3 bench (+ tmem parity tests) for the TmShapes named in $SHAPES, tight timeouts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for s in ${SHAPES:-0}; do
  echo "== VND_TM_SHAPE=$s"
  VND_TM_SHAPE=$s CH=${CH:-148} bash tools/try_fir.sh 2>&1 | tail -4
done | tee -a gpurun_out/r02_shapes.txt
