O=gpurun_out; mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/bench_r1p.json 2> $O/bench_r1p.err; echo "bench rc=$?"; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_r1p.json'))
print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['clocks'])
print(json.dumps(j['other_configs']['c5_ehvi3d'], indent=1))
print(j['cpu_baseline'])
PY
tail -3 $O/bench_r1p.err
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2
