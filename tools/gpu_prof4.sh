#!/bin/bash
# ncu --set full with source correlation: diagonal kernel j = 0 (POTF2 only) and j = 7, panel kernel j = 1, C2 step.  usage: tools/gpu_prof4.sh <tag>
TAG=${1:-r2y}
O=gpurun_out
mkdir -p $O
python tools/prof_run.py lik 2 > $O/plain_lik.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:chol_diag_half -s 16 -c 1 -f -o $O/prof_diag0_$TAG python tools/prof_run.py lik 2 > $O/ncu_full_diag0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chol_diag_half -s 23 -c 1 -f -o $O/prof_diag7_$TAG python tools/prof_run.py lik 2 > $O/ncu_full_diag7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chol_panel_half -s 14 -c 1 -f -o $O/prof_panel1_$TAG python tools/prof_run.py lik 2 > $O/ncu_full_panel1.log 2>&1
ls -la $O/prof_diag0_$TAG.ncu-rep $O/prof_diag7_$TAG.ncu-rep $O/prof_panel1_$TAG.ncu-rep
