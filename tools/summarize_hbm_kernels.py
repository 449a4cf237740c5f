"""HBM-bound kernels: achieved DRAM bandwidth per launch from an `ncu --set full` raw CSV, against the measured copy peak.

    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/summarize_hbm_kernels.py raw.csv [peak_GBs] > profiles/...md
"""
import csv
import json
import os
import re
import sys


def main(path, peak):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, k, scale_to=None):
        v, u = float(r[ix[k]]), units[ix[k]]
        if scale_to == "bytes":
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        if scale_to == "s":
            return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1e-9)
        return v
    print(f"# {os.path.basename(path)}: HBM-bound kernels, `ncu --set full --clock-control none` (cold caches), peak = {peak:.0f} GB/s "
          f"(MEASURED_PEAKS.json copy bandwidth)\n")
    print("| # | kernel | time us | DRAM read MB | DRAM write MB | achieved GB/s | of peak | issue % | occupancy % | top stall |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---|")
    for i, r in enumerate(data):
        name = re.sub(r"^void ", "", r[ix["Kernel Name"]]).split("(")[0].replace("md::", "")
        t = val(r, "gpu__time_duration.sum", "s")
        rd, wr = val(r, "dram__bytes_read.sum", "bytes"), val(r, "dram__bytes_write.sum", "bytes")
        gbs = (rd + wr) / t / 1e9
        stalls = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: float(r[j])
                  for h, j in ix.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
        top = max(stalls.items(), key=lambda kv: kv[1]) if stalls else ("", 0)
        print(f"| {i} | `{name}` | {t * 1e6:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {gbs / peak:.2f} | "
              f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {top[0]} {top[1]:.1f} |")


if __name__ == "__main__":
    pk = 6571.9
    try:
        pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else pk)
