"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_lik_r1a.ncu-rep profiles/r1a_lik_full.md
"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_pipe%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math_thr"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(m, s, hdr.index(m)) for m, s in METRICS if m in hdr]
    kn = hdr.index("Kernel Name")
    lines = ["# ncu --set full summary of `%s`" % rep.split("/")[-1], "",
             "(`ncu -i <rep> --page raw --csv`; durations are cold-cache / serialised replays: compare shares, not absolutes)", "",
             "| # | kernel | " + " | ".join("%s [%s]" % (s, units[i]) if units[i] else s for _, s, i in cols) + " |",
             "|---|---|" + "---|" * len(cols)]
    for k, r in enumerate(rows[2:]):
        name = r[kn].split("(")[0].replace("void ", "")
        vals = []
        for _, _, i in cols:
            v = r[i]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            vals.append(v)
        lines.append("| %d | %s | " % (k, name) + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
