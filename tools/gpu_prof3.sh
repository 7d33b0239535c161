#!/bin/bash
# ncu of the general predictor on the C3 model (n = 200: pure solve of sample tile 0, ragged sample tile 1) -- launch list and
# one --set full capture each of the two panel launches.   usage: tools/gpu_prof3.sh <tag>
TAG=${1:-r2y}
O=gpurun_out
mkdir -p $O
export KB200_NO_FUSED_PREDICT=1
python tools/prof_run.py pred 2 > $O/plain_predg.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_predg_launches.csv python tools/prof_run.py pred 2 > $O/ncu_predg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chol_panel_half -s 2 -c 2 -f -o $O/prof_predg_$TAG python tools/prof_run.py pred 2 > $O/ncu_full_predg.log 2>&1
python tools/ncu_summary.py $O/prof_predg_$TAG.ncu-rep $O/${TAG}_predg_full.md > /dev/null 2>&1
ls -la $O/*predg*
