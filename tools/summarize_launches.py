"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

    python tools/summarize_launches.py gpurun_out/r01_launches_c2.csv > profiles/r01_launches_c2_summary.md
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(?:md::)?([A-Za-z0-9_:]+)(<[^(]*>)?\(", name)
    if m and not name.startswith("at::"):
        return m.group(1) + (m.group(2) or "")
    m = re.match(r"at::native::(?:<unnamed>::)?([A-Za-z0-9_]+)", name)
    if m:
        return "torch:" + m.group(1)
    return name.split("(")[0][:70]


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        rows.append((short(r["Kernel Name"]), ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.2f} ms of kernel time (ncu: serialised, cold caches)\n")
    print("| kernel | launches | total ms | share | mean us |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns / 1e6:.3f} | {ns / tot:.4f} | {ns / n / 1e3:.1f} |")
    fam = defaultdict(float)
    for k, (n, ns) in agg.items():
        key = ("gemm_tcgen05_kernel" if k.startswith("gemm_tcgen05") else "attention" if k.startswith("attn_") else
               "torch (RNG, zero-fill, copies)" if k.startswith("torch:") or k.startswith("at::") else "row / element-wise kernels")
        fam[key] += ns
    print("\n| family | share |\n|---|---:|")
    for k, ns in sorted(fam.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {ns / tot:.4f} |")


if __name__ == "__main__":
    main(sys.argv[1])
