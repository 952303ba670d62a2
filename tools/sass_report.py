"""Per-kernel SASS mnemonic counts of libdstd_b200.so (what proves a Blackwell-native kernel: UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UBLKCP / UTMALDG = TMA, HMMA = legacy mma.sync, LDGSTS = cp.async).

    python tools/sass_report.py > profiles/r02_sass_mnemonics.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dstd_gcn_b200", "csrc", "libdstd_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "LDGSTS", "SYNCS", "R2UR", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    counts[cur][w] += 1
    print("# SASS mnemonics per kernel (`cuobjdump -sass dstd_gcn_b200/csrc/libdstd_b200.so`, sm_100a)\n")
    print("UTCHMMA = `tcgen05.mma`, LDTM = `tcgen05.ld`, UTCBAR = `tcgen05.commit`, UBLKCP = `cp.async.bulk` (TMA bulk copy), "
          "HMMA = `mma.sync` (legacy tensor path), LDGSTS = `cp.async`, SYNCS = mbarrier ops.\n")
    cols = [w for w in WATCH if any(c[w] for c in counts.values())]
    print("| kernel | instructions | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for k, c in counts.items():
        if not any(c[w] for w in cols if w not in ("MUFU", "R2UR")):
            continue
        print(f"| `{k}` | {c['total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print(f"\nWhole library: " + ", ".join(f"{w} {tot[w]}" for w in cols) + f"; {len(counts)} kernels.")


if __name__ == "__main__":
    main()
