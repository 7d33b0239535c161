#!/bin/bash
# ncu --set full captures with source correlation of the two kernels furthest below their roof (round 2).
# usage: tools/gpu_prof2.sh <tag>
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
python tools/prof_run.py pred 2 > $O/plain_pred.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:predict_small -s 1 -c 1 -f -o $O/prof_pred_$TAG python tools/prof_run.py pred 2 > $O/ncu_full_pred.log 2>&1
python tools/prof_run.py lik 2 > $O/plain_lik.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chol_diag_half -s 18 -c 3 -f -o $O/prof_diag_$TAG python tools/prof_run.py lik 2 > $O/ncu_full_lik.log 2>&1
python tools/ncu_summary.py $O/prof_pred_$TAG.ncu-rep $O/${TAG}_pred_full.md > /dev/null 2>&1
python tools/ncu_summary.py $O/prof_diag_$TAG.ncu-rep $O/${TAG}_diag_full.md > /dev/null 2>&1
ls -la $O/*.ncu-rep
