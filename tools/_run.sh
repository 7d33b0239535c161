O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q > $O/pytest_r1q.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_r1q.log
python bench.py --steps 10 --warmup 3 > $O/bench_r1q.json 2> $O/bench_r1q.err; echo "bench rc=$?"; wc -l $O/bench_r1q.json; python -c "
import json; j=json.load(open('gpurun_out/bench_r1q.json')); print(j['value'], j['e2e']['value'], j['predict']['value'], j['other_configs']['c5_ehvi3d']['independent'], j['cpu_baseline']['value'])"
