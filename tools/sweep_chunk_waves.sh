#!/bin/bash
# psi chunk size of the general predictor (waves of point tiles per launch): does a chunk that fits the 126 MB L2 beat fewer, larger launches?
O=gpurun_out; mkdir -p $O
for w in 1 2 3 4 8; do
  echo "== KB200_CHUNK_WAVES=$w" | tee -a $O/sweep_chunk_waves.log
  KB200_CHUNK_WAVES=$w KB200_NO_FUSED_PREDICT=1 python tools/bench_extra.py c3 c3d10 c4var c5pred 2>&1 | grep '"case"' | cut -c1-260 | tee -a $O/sweep_chunk_waves.log
done
