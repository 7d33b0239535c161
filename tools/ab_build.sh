#!/bin/bash
# A/B of compile-time kernel variants on the GPU box: rebuild libkb200.so with each set of -D flags, time the C2 generation.
# usage: tools/ab_build.sh "<flags A>" "<flags B>" ...   ("" = default build); the default build is restored at the end
O=gpurun_out; mkdir -p $O
for V in "$@"; do
  KB200_NVCC_EXTRA="$V" python -m kadal_b200._build --force > /dev/null 2>$O/ab_build.err || { echo "build failed: $V"; tail -5 $O/ab_build.err; continue; }
  KB200_NVCC_EXTRA="$V" python tools/time_c2.py 2>>$O/ab_build.err | tee -a $O/ab_build.log
  python tools/bench_extra.py c4 2>>$O/ab_build.err | grep '"case"' | cut -c1-160 | tee -a $O/ab_build.log
done
python -m kadal_b200._build --force > /dev/null 2>&1
