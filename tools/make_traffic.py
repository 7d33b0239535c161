"""profiles/traffic.json from an `ncu --set full` capture of one likelihood pass: per kernel class the
average DRAM bytes (read + write) per launch, which bench.py reports as roofline.traffic.

    python tools/make_traffic.py gpurun_out/prof_lik_<tag>.ncu-rep [<pred rep>]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASS = {"corr_tile_kernel": "assemble", "chol_diag_kernel": "chol_diag", "chol_panel_kernel": "chol_panel",
         "chol_diag_half_kernel": "chol_diag", "chol_panel_half_kernel": "chol_panel", "solve_rows_half_kernel": "chol_panel",
         "lik_finalize_kernel": "finalize", "predict_small_kernel": "predict_small", "lik_small_kernel": "lik_small"}


def scale(unit):
    u = unit.lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)


def main():
    out = {}
    for rep in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        acc = {}
        # a capture of the tiled predictor (prof_pred_*): its variance-solve launches (solve_rows_half_kernel + chol_panel_half_kernel
        # in prediction mode) form the class bench.py's predict.roofline reports; psi and epilogue launches are listed beside them
        pred = os.path.basename(rep).startswith("prof_pred_")
        for r in rows[2:]:
            name = r[kn].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
            cls = CLASS.get(name)
            if pred:
                cls = {"solve_rows_half_kernel": "predict_tiled", "chol_panel_half_kernel": "predict_tiled",
                       "corr_tile_kernel": "predict_tiled_psi", "predict_epilogue_kernel": "predict_tiled_epilogue",
                       "predict_small_kernel": "predict_small"}.get(name)
            if cls is None:
                continue
            b = float(r[rd].replace(",", "")) * scale(units[rd]) + float(r[wr].replace(",", "")) * scale(units[wr])
            acc.setdefault(cls, []).append(b)
        for cls, v in acc.items():
            out[cls] = {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v), "source": os.path.basename(rep)}
    path = os.path.join(ROOT, "profiles", "traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
