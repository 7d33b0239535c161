#!/bin/bash
# One gpurun call: GPU tests, parity report, bench, then the ncu passes of B200_PROFILING.md.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/env_$TAG.txt; nproc >> $O/env_$TAG.txt; ls /root/reference >> $O/env_$TAG.txt 2>&1
python -m pytest tests -m gpu -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log; tail -15 $O/pytest_$TAG.log
python tools/parity_report.py 2>&1 | tee $O/parity_$TAG.log
python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cat $O/bench_$TAG.json; tail -5 $O/bench_$TAG.err
if [ "$2" != "nocu" ]; then
python tools/prof_run.py lik 2 > $O/plain_lik.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_lik_$TAG.csv python tools/prof_run.py lik 2 > $O/ncu_lik.log 2>&1
python tools/prof_run.py pred 2 > $O/plain_pred.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_pred_$TAG.csv python tools/prof_run.py pred 2 > $O/ncu_pred.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'chol_diag|chol_panel|corr_tile|lik_finalize' -s 17 -c 17 -f -o $O/prof_lik_$TAG python tools/prof_run.py lik 2 > $O/ncu_full_lik.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'predict_small|chol_panel|corr_tile|predict_epi' -s 2 -c 3 -f -o $O/prof_pred_$TAG python tools/prof_run.py pred 2 > $O/ncu_full_pred.log 2>&1
# summaries are made on the box; gpurun copies back at most 64 MiB, so the largest capture is dropped when the
# directory is over that (the .md keeps what profiles/ needs)
python tools/ncu_summary.py $O/prof_lik_$TAG.ncu-rep $O/${TAG}_lik_full.md > /dev/null 2>&1
python tools/ncu_summary.py $O/prof_pred_$TAG.ncu-rep $O/${TAG}_pred_full.md > /dev/null 2>&1
if [ $(du -sm $O | cut -f1) -gt 55 ]; then rm -f $O/prof_lik_$TAG.ncu-rep; fi
ls -la $O
fi
