O=gpurun_out; mkdir -p $O
python tools/worst_candidates.py 2>&1 | tee $O/worst_r1o.log
python tools/stress_half.py 2>&1 | tee -a $O/worst_r1o.log
python tools/stress_rl.py 2>&1 | tee -a $O/worst_r1o.log
python -m pytest tests -m gpu -q -x > $O/pytest_r1o.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_r1o.log; tail -8 $O/pytest_r1o.log
python bench.py --steps 10 --warmup 3 > $O/bench_r1o.json 2> $O/bench_r1o.err; echo "bench rc=$?"; cat $O/bench_r1o.json; tail -3 $O/bench_r1o.err
