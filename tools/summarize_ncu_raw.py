"""Key metrics per launch from `ncu -i X.ncu-rep --page raw --csv` (one row per launch) as a markdown table.

    python tools/summarize_ncu_raw.py gpurun_out/r01_ncu_gemm_full_raw.csv > profiles/r01_ncu_gemm_full.md
"""
import csv
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma inst %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
    ("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "L1->XBAR %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("smsp__inst_executed.sum", "warp inst"),
]


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    use = [(k, n) for k, n in COLS if k in idx]
    print(f"# {path}: {len(data)} launches (`ncu --set full --clock-control none`; cold caches, serialised)\n")
    print("| # | kernel | " + " | ".join(n for _, n in use) + " | top stalls (pc samples) |")
    print("|---|---|" + "---:|" * len(use) + "---|")
    for li, r in enumerate(data):
        name = re.sub(r"^void ", "", r[idx["Kernel Name"]]).split("(")[0]
        cells = []
        for k, _ in use:
            v, u = r[idx[k]], units[idx[k]]
            try:
                f = float(v.replace(",", ""))
                if k == "gpu__time_duration.sum":
                    f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                    cells.append(f"{f:.1f} us")
                elif u.endswith("byte"):
                    f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                    cells.append(f"{f:.1f} MB")
                else:
                    cells.append(f"{f:.1f} {u}".strip() if u not in ("", "%") else (f"{f:.1f}" if u == "%" else f"{f:.0f}"))
            except ValueError:
                cells.append(v)
        st = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        top = ", ".join(f"{n} {v / tot:.0%}" for v, n in sorted(st, reverse=True)[:4])
        print(f"| {li} | `{name}` | " + " | ".join(cells) + f" | {top} |")


def write_traffic(path, launch, out):
    """profiles/ncu_gemm_traffic.json for bench.py's roofline.traffic: DRAM bytes (read + write) of launch #`launch`."""
    import json
    rows = list(csv.reader(open(path, newline="")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    r = data[launch]
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[idx[k]].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[k]]]
    name = re.sub(r"^void ", "", r[idx["Kernel Name"]]).split("(")[0]
    json.dump({"dram_bytes": tot, "launch": f"#{launch} {name}", "source": path}, open(out, "w"), indent=1)
    print(f"wrote {out}: {tot / 1e6:.1f} MB for launch #{launch} {name}", file=sys.stderr)


if __name__ == "__main__":
    main(sys.argv[1])
    if len(sys.argv) >= 5 and sys.argv[2] == "--traffic":
        write_traffic(sys.argv[1], int(sys.argv[3]), sys.argv[4])
