#!/bin/bash
# Final round-2 evidence (after the ragged-tile, pure-solve and blocked-POTF2 changes): GPU tests, bench (both arms), launch lists, --set full captures of every hot kernel, traffic.json.
TAG=${1:-r2zz}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --durations=10 > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log; tail -4 $O/pytest_$TAG.log
cp $O/parity_c2_counts.json $O/${TAG}_parity_c2_counts.json 2>/dev/null
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -3 $O/bench_$TAG.err
python tools/stress_ragged.py 7 30 > $O/stress_ragged_$TAG.log 2>&1; tail -1 $O/stress_ragged_$TAG.log
python tools/time_pred_small.py > $O/time_pred_small_$TAG.log 2>&1; python tools/time_small.py > $O/time_small_$TAG.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log
python tools/prof_run.py lik 2 > $O/plain_lik.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_lik_$TAG.csv python tools/prof_run.py lik 2 > $O/ncu_lik.log 2>&1
python tools/prof_run.py pred 2 > $O/plain_pred.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_pred_$TAG.csv python tools/prof_run.py pred 2 > $O/ncu_pred.log 2>&1
ncu --set full --clock-control none -k regex:'chol_diag|chol_panel|solve_rows|corr_tile|lik_finalize' -s 17 -c 17 -f -o $O/prof_lik_$TAG python tools/prof_run.py lik 2 > $O/ncu_full_lik.log 2>&1
ncu --set full --clock-control none -k regex:'corr_tile|solve_rows|chol_panel|predict_epilogue' -s 20 -c 4 -f -o $O/prof_pred_$TAG python tools/prof_run.py pred 2 > $O/ncu_full_pred.log 2>&1
python tools/prof_run.py pred4 2 > $O/plain_pred4.log 2>&1 &&
ncu --set full --clock-control none -k regex:'corr_tile' -s 1 -c 2 -f -o $O/prof_pred4_$TAG python tools/prof_run.py pred4 2 > $O/ncu_full_pred4.log 2>&1
python tools/prof_run.py small 2 > $O/plain_small.log 2>&1 &&
ncu --set full --clock-control none -k regex:'lik_small' -s 1 -c 2 -f -o $O/prof_small_$TAG python tools/prof_run.py small 2 > $O/ncu_full_small.log 2>&1
for k in lik pred pred4 small; do python tools/ncu_summary.py $O/prof_${k}_$TAG.ncu-rep $O/${TAG}_${k}_full.md > /dev/null 2>&1; done
python tools/make_traffic.py $O/prof_lik_$TAG.ncu-rep $O/prof_pred_$TAG.ncu-rep $O/prof_small_$TAG.ncu-rep > $O/traffic_$TAG.log 2>&1; cp profiles/traffic.json $O/traffic_$TAG.json
rm -f $O/prof_lik_$TAG.ncu-rep
ls -la $O | tail -30
