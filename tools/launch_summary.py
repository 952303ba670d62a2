"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into a per-kernel markdown table.

    python tools/launch_summary.py gpurun_out/launches.csv "title line" > profiles/rNN_launches_xxx.md
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    lines = [ln for ln in open(path) if ln.startswith('"')]
    tot, cnt = defaultdict(float), defaultdict(int)
    for row in csv.DictReader(lines):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        if name.startswith("at::"):
            name = row["Kernel Name"].replace("void ", "")[:70]
        tot[name] += float(row["Metric Value"]) / 1e6
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {title}\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / total:.1f}% | {1e3 * tot[k] / cnt[k]:.1f} |")
    print(f"\nTotal {total:.2f} ms over {sum(cnt.values())} launches.")


if __name__ == "__main__":
    main()
