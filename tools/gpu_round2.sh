#!/bin/bash
# Round-2 gpurun call: GPU tests (parity counts kept), bench (ours + reference arm), optional extras.
# usage: tools/gpu_round2.sh <tag> [tests|bench|all]
TAG=${1:-r2a}
WHAT=${2:-all}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/env_$TAG.txt; nproc >> $O/env_$TAG.txt
if [ "$WHAT" = "all" ] || [ "$WHAT" = "tests" ]; then
  python -m pytest tests -m gpu -q --durations=15 $PYTEST_EXTRA > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log; tail -30 $O/pytest_$TAG.log
  [ -f $O/parity_c2_counts.json ] && cp $O/parity_c2_counts.json $O/${TAG}_parity_c2_counts.json
fi
if [ "$WHAT" = "all" ] || [ "$WHAT" = "bench" ]; then
  python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; head -c 3000 $O/bench_$TAG.json; echo; tail -5 $O/bench_$TAG.err
fi
