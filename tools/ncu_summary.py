"""Summarise an `ncu --set full` report (read here, without a GPU) into a markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.md
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{rep}`\n")
    for r in rows[2:]:
        print(f"## `{r[idx['Kernel Name']]}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
        print("| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in idx:
                print(f"| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace(
                        "_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        top = ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:6])
        print(f"| top stall reasons (warps per issue) | {top} |\n")


if __name__ == "__main__":
    main()
