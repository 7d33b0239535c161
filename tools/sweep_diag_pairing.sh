#!/bin/bash
# One diagonal CTA per SM (KB200_DIAG_SMEM_KB=116) against the default two, for 2 / 3 / 4 candidate groups and the dual-stream schedule.
O=gpurun_out; mkdir -p $O
for kb in 0 116; do for g in 2 3 4; do for dual in 0 1; do
  echo "== DIAG_SMEM_KB=$kb GROUPS=$g DUAL=$dual" | tee -a $O/sweep_diag_pairing.log
  KB200_DIAG_SMEM_KB=$kb KB200_GROUPS=$g KB200_DUAL=$dual KB200_TIME_C2_ONLY=1 python tools/time_c2.py 2>&1 | head -1 | tee -a $O/sweep_diag_pairing.log
done; done; done
