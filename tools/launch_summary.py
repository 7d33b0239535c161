"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list, split at lik_finalize_kernel
(one block per fit).   python tools/launch_summary.py gpurun_out/launches_append_4000.csv"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    seq.append((name, v))
cuts = [i + 1 for i, (n, _) in enumerate(seq) if "lik_finalize" in n] + [len(seq)]
lo = 0
for k, hi in enumerate(cuts):
    if hi == lo:
        continue
    c = collections.OrderedDict()
    for n, v in seq[lo:hi]:
        a = c.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    print("block %d (launches %d..%d)" % (k, lo, hi - 1))
    for n, (cnt, v) in c.items():
        print("   %-58s %5d %10.1f us" % (n[:58], cnt, v))
    print("   total %.1f us" % sum(v for _, v in seq[lo:hi]))
    lo = hi
