"""Summarise an `ncu --page source --csv` export of one kernel: warp-stall samples aggregated between block barriers
(BAR.SYNC), with the dominant stall reasons and opcodes of each segment.  Used to find which phase of the fused unit
kernels the time goes to (the CUDA-C correlation of the csv page is SASS only).

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:unit_fwd > src.csv ; python tools/sass_segments.py src.csv
"""
import csv
import sys


def analyze(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    data = [r for r in rows[1:] if r[0] != hdr[0]]          # one kernel per file (filter with --kernel-name)
    si = hdr.index('# Samples')
    ii = hdr.index('Instructions Executed')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[si]) for r in data)
    itot = sum(int(r[ii]) for r in data)
    print('total samples', tot, 'SASS instructions', len(data), 'executed warp instructions', itot)
    seg, cur, start = [], 0, 0
    for idx, r in enumerate(data):
        cur += int(r[si])
        if 'BAR.SYNC' in r[1] or 'BAR.RED' in r[1]:
            seg.append((start, idx, cur))
            start, cur = idx + 1, 0
    seg.append((start, len(data) - 1, cur))
    for s in seg:
        if s[2] <= tot * 0.01:
            continue
        agg, kinds = {}, {}
        for r in data[s[0]:s[1] + 1]:
            for c in stall_cols:
                v = int(r[c] or 0)
                if v:
                    agg[hdr[c]] = agg.get(hdr[c], 0) + v
            toks = r[1].split()
            op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
            kinds[op] = kinds.get(op, 0) + int(r[si])
        top = sorted(agg.items(), key=lambda x: -x[1])[:4]
        topk = sorted(kinds.items(), key=lambda x: -x[1])[:6]
        iex = sum(int(r[ii]) for r in data[s[0]:s[1] + 1])
        print(f"  instr {s[0]:5d}-{s[1]:5d}: {100 * s[2] / tot:5.1f}% of samples, {100 * iex / itot:5.1f}% of executed instructions  stalls {top}  ops {topk}")


if __name__ == '__main__':
    analyze(sys.argv[1])
