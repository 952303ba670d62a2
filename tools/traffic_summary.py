"""Sum ncu DRAM traffic over whole training steps of `bench.py --no-graph` (window = launches between the first and
the last Adam kernel in the capture, i.e. an integer number of identical steps) and write profiles/<name>.json with
bytes per step and per kernel; bench.py reports it as roofline.traffic.

    python tools/traffic_summary.py gpurun_out/traffic.csv profiles/r01_dram_traffic_h36m_b256.json
"""
import collections
import csv
import json
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    lines = [l for l in open(src) if not l.startswith("==")]
    order = {}
    for row in csv.DictReader(lines):
        order.setdefault(row["ID"], (len(order), row["Kernel Name"]))
    adam = [i for i, n in order.values() if "adam_kernel" in n]
    lo, hi, nsteps = adam[0], adam[-1], len(adam) - 1
    per_kernel = collections.defaultdict(lambda: {"launches": 0, "dram_bytes": 0.0, "time_ns": 0.0})
    seen = set()
    for row in csv.DictReader(lines):
        if not (lo < order[row["ID"]][0] <= hi):
            continue
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("dstd::", "")
        metric, unit = row["Metric Name"], row["Metric Unit"]
        k = per_kernel[name]
        if metric.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            k["dram_bytes"] += v * mult
        elif metric == "gpu__time_duration.sum":
            mult = {"ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
            k["time_ns"] += v * mult
            if row["ID"] not in seen:
                seen.add(row["ID"])
                k["launches"] += 1
    total = sum(k["dram_bytes"] for k in per_kernel.values())
    out = {"source": src, "steps_in_capture": nsteps, "dram_bytes_per_step": total / nsteps,
           "kernels": {n: {"launches_per_step": k["launches"] / nsteps, "dram_bytes_per_step": k["dram_bytes"] / nsteps,
                           "time_ms_per_step": k["time_ns"] / nsteps * 1e-6}
                       for n, k in sorted(per_kernel.items(), key=lambda x: -x[1]["dram_bytes"])[:30]}}
    json.dump(out, open(dst, "w"), indent=1)
    print(f"DRAM traffic per step: {total / nsteps / 1e9:.3f} GB")
    for n, k in list(out["kernels"].items())[:12]:
        print(f"  {k['dram_bytes_per_step'] / 1e9:7.3f} GB  {k['time_ms_per_step']:7.3f} ms  "
              f"{k['dram_bytes_per_step'] / max(k['time_ms_per_step'], 1e-9) / 1e6:7.1f} GB/s  {n[:60]}")


if __name__ == "__main__":
    main()
