#!/usr/bin/env bash
# Round 2: parity + short cfg3 bench for every TmShape of the tensor-memory FIR (VND_TM_SHAPE), tight timeouts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu suite (default shape)"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for s in ${SHAPES:-0 1 2 3 4 5}; do
  echo "== VND_TM_SHAPE=$s"
  VND_TM_SHAPE=$s CH=${CH:-148} bash tools/try_fir.sh 2>&1 | tail -4
done | tee gpurun_out/r02_shapes.txt
