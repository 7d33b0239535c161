"""Per-kernel SASS evidence for profiles/: instruction tally (DMMA = mma.sync.m8n8k4.f64, UBLKCP = cp.async.bulk,
SYNCS = mbarrier, DFMA/DADD/DMUL, MUFU, LDS/STS, SHFL) and resource usage (registers, stack = spills, shared memory)
of every kernel in kadal_b200/_lib/*.o, from `cuobjdump -sass` and `cuobjdump -res-usage` (no GPU needed).

    python tools/sass_tally.py > profiles/r2_sass_tally.md
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["DMMA", "UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "SHFL", "LDG", "STG", "BAR", "LDL", "STL"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("kb200::", "").replace("void ", "")
    except Exception:
        return name


def main():
    print("# SASS tally per kernel (sm_100a, `cuobjdump -sass` / `-res-usage` of kadal_b200/_lib/*.o)\n")
    print("DMMA = `mma.sync.aligned.m8n8k4.f64` (FP64 tensor pipe); UBLKCP = `cp.async.bulk` (TMA 1-D); SYNCS = mbarrier; "
          "LDL / STL = local-memory (spill) accesses; STACK = bytes of stack frame.\n")
    print("| object | kernel | REG | STACK | " + " | ".join(OPS) + " | total |")
    print("|---|---|---|---|" + "---|" * (len(OPS) + 1))
    for obj in sorted(glob.glob(os.path.join(ROOT, "kadal_b200", "_lib", "*.o"))):
        res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
        usage = dict((m.group(1), (m.group(2), m.group(3))) for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", res))
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur, tally = None, collections.OrderedDict()
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                tally[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and cur:
                tally[cur][m.group(1)] += 1
                tally[cur]["__total"] += 1
        for fn, c in tally.items():
            reg, stack = usage.get(fn, ("?", "?"))
            print("| %s | `%s` | %s | %s | " % (os.path.basename(obj), demangle(fn), reg, stack) +
                  " | ".join(str(c.get(o, 0)) for o in OPS) + " | %d |" % c["__total"])


if __name__ == "__main__":
    main()
