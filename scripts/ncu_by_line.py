#!/usr/bin/env python
"""Aggregate an ncu SASS-level source page by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > sass.csv
    cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > dis.txt
    python scripts/ncu_by_line.py sass.csv dis.txt <mangled-kernel-substring> [top]

Pairs the i-th SASS instruction of the kernel in the nvdisasm listing (which carries
`//## File "...", line N` markers) with the i-th row of the ncu CSV, then sums instructions
executed and stall samples per source line.
"""
import csv
import re
import sys
from collections import defaultdict


def disasm_lines(path, kernel_sub):
    lines = []
    cur = None
    active = False
    for raw in open(path, errors='replace'):
        if raw.startswith('//--------------------- .text.'):
            active = kernel_sub in raw
            cur = None
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', raw)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith('.cu') else -int(m.group(2))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', raw):
            lines.append(cur)
    return lines


def main():
    sass_csv, dis, sub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[h]
    body = [r for r in rows[h + 1:] if r and r[0].strip()]
    lines = disasm_lines(dis, sub)
    if len(lines) != len(body):
        print(f"warning: {len(lines)} disassembled vs {len(body)} profiled instructions", file=sys.stderr)
    ci = {n: hdr.index(n) for n in ('Instructions Executed', '# Samples', 'stall_long_sb', 'stall_short_sb', 'stall_wait',
                                    'stall_barrier', 'stall_math', 'stall_mio', 'L1 Wavefronts Shared Excessive')}
    agg = defaultdict(lambda: defaultdict(float))
    for ln, r in zip(lines, body):
        for k, i in ci.items():
            try:
                agg[ln][k] += float(r[i] or 0)
            except ValueError:
                pass
    tot_i = sum(a['Instructions Executed'] for a in agg.values())
    tot_s = sum(a['# Samples'] for a in agg.values())
    src = {}
    try:
        for n, l in enumerate(open('mcmc_spec_b200/csrc/mcmcspec.cu'), 1):
            src[n] = l.rstrip()
    except OSError:
        pass
    print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
    print(f"{'line':>6} {'inst%':>6} {'smp%':>6} {'long':>6} {'short':>6} {'wait':>6} {'bar':>6} {'math':>6}  source")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]['# Samples'])[:top]:
        s = max(a['# Samples'], 1)
        print(f"{str(ln):>6} {100 * a['Instructions Executed'] / tot_i:6.2f} {100 * a['# Samples'] / tot_s:6.2f} "
              f"{100 * a['stall_long_sb'] / s:6.1f} {100 * a['stall_short_sb'] / s:6.1f} {100 * a['stall_wait'] / s:6.1f} "
              f"{100 * a['stall_barrier'] / s:6.1f} {100 * a['stall_math'] / s:6.1f}  {src.get(ln, '')[:90] if isinstance(ln, int) and ln > 0 else '(header/inlined lib)'}")


if __name__ == '__main__':
    main()
