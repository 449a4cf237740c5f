"""Per-kernel SASS evidence from the built library: which instruction families each kernel of libmicrodit_b200.so uses.

    python tools/sass_summary.py > profiles/rNN_sass_summary.md

Runs `cuobjdump -sass` (no GPU needed) and counts, per kernel, the mnemonics that prove the Blackwell path
(/opt/skills/guides/B200_PROFILING.md): UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (tcgen05.ld / st),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus HMMA (mma.sync), LDGSTS (cp.async), MUFU, RED/ATOM, and local-memory
traffic (LDL / STL: spills or runtime-indexed arrays) and the 256-bit global accesses (STG.256 / LDG.256).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "micro_diffusion_b200", "libmicrodit_b200.so")
FAMILIES = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "MUFU", "RED", "ATOM", "LDL", "STL",
            "STG.256", "LDG.256"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            if ".256" in line and ("STG." in line or "LDG." in line):   # 256-bit global accesses (sm_100)
                counts[cur]["STG.256" if "STG." in line else "LDG.256"] += 1
            for f in FAMILIES:
                if op.startswith(f):
                    counts[cur][f] += 1
                    break
    names = demangle(order)
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (`cuobjdump -sass`, sm_100a)\n")
    print("| kernel | instr | " + " | ".join(FAMILIES) + " |")
    print("|---|---:|" + "---:|" * len(FAMILIES))
    for k in sorted(order, key=lambda k: names[k]):
        n = re.sub(r"^void ", "", names[k]).split("(")[0].replace("md::", "")
        c = counts[k]
        print(f"| `{n}` | {c['_total']} | " + " | ".join(str(c[f]) if c[f] else "" for f in FAMILIES) + " |")


if __name__ == "__main__":
    sys.exit(main())
