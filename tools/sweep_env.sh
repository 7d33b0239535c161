#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ...   -> likelihood evals/s of bench.py for every value of an env knob
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --predict-points 2e5 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); print('$VAR=$v', round(b['value']), 'evals/s', round(b['ms_per_step'],3), 'ms', b['kernel_ms'])"
done
