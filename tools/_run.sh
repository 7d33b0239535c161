python -m pytest tests/test_ehvi3d.py -m gpu -q -x 2>&1 | tail -3
python tools/prof_ehvi3d.py 100 1000 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ehvi3d.csv python tools/prof_ehvi3d.py 100 1000 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_ehvi3d.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]: print(r[4][:40], r[7], r[8], r[-1])
PY
