#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on).

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python scripts/ncu_lines.py src.csv [top]

Sums, per CUDA source line, the warp instructions executed, the stall samples and the shared-memory wavefronts.
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur = None
    hdr = None
    agg = defaultdict(lambda: defaultdict(float))
    src = {}
    for r in csv.reader(open(path, errors='replace')):
        if not r:
            continue
        if r[0] in ('File Name', 'File Path'):
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or not r[0].strip() or not r[0].strip().isdigit():
            continue
        key = (cur, int(r[0]))
        src[key] = r[1]
        for name in ('Instructions Executed', '# Samples', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_barrier',
                     'stall_math', 'stall_branch_resolving', 'stall_selected', 'stall_not_selected', 'L1 Wavefronts Shared',
                     'stall_no_inst', 'stall_mio'):
            try:
                agg[key][name] += float(r[hdr.index(name)] or 0)
            except ValueError:
                pass
    ti = sum(a['Instructions Executed'] for a in agg.values())
    ts = sum(a['# Samples'] for a in agg.values())
    tw = sum(a['L1 Wavefronts Shared'] for a in agg.values())
    print(f"total warp instructions {ti:.4e}  samples {ts:.0f}  shared wavefronts {tw:.4e}")
    print(f"{'file:line':>24} {'inst%':>6} {'smp%':>6} {'wf%':>6} {'long':>5} {'short':>5} {'wait':>5} {'bar':>5} {'math':>5} {'br':>5} {'sel':>5}  source")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]['# Samples'])[:top]:
        s = max(a['# Samples'], 1)
        print(f"{key[0][:16] + ':' + str(key[1]):>24} {100 * a['Instructions Executed'] / ti:6.2f} {100 * a['# Samples'] / ts:6.2f} "
              f"{100 * a['L1 Wavefronts Shared'] / max(tw, 1):6.2f} {100 * a['stall_long_sb'] / s:5.0f} {100 * a['stall_short_sb'] / s:5.0f} "
              f"{100 * a['stall_wait'] / s:5.0f} {100 * a['stall_barrier'] / s:5.0f} {100 * a['stall_math'] / s:5.0f} "
              f"{100 * a['stall_branch_resolving'] / s:5.0f} {100 * (a['stall_selected'] + a['stall_not_selected']) / s:5.0f}  {src[key].strip()[:100]}")


if __name__ == '__main__':
    main()
