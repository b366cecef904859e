#!/usr/bin/env python
"""Run every kernel of the library once or twice at BASELINE shapes (for an ncu per-kernel table).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum \\
        --clock-control none --csv --log-file gpurun_out/kernels.csv python scripts/kernel_table.py
    python scripts/kernel_table.py summarize gpurun_out/kernels.csv > profiles/r1_kernel_table.md
"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import numpy as np
    import torch
    from mcmc_spec_b200 import native, synthetic
    from mcmc_spec_b200.ensemble import CudaBackend, EnsembleSampler
    from mcmc_spec_b200.gridbuild import gaussian_taps
    from mcmc_spec_b200.posterior import LogPosterior
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    from bench_configs import central_third

    # ---- setup kernels at the koi2298 size: 104 nodes, 135 000-pixel rows, 9750-pixel window, 93 taps
    rng = np.random.default_rng(0)
    wl = np.arange(3000.0, 30000.0, 0.2)
    node_wl, node_fl = [], []
    for k in range(104):
        n = 150000 + int(rng.integers(0, 20000))
        w = np.sort(rng.uniform(2900.0, 30100.0, n)); w[0], w[-1] = 2900.0, 30100.0
        node_wl.append(w); node_fl.append(1e5 * (1 + 0.3 * np.sin(w / 50.0 + k)))
    inwin = np.where((wl >= 6450.0) & (wl <= 8400.0))[0]
    taps = gaussian_taps(wl[inwin], 1700)
    native.build_grid(node_wl, node_fl, wl, int(inwin[0]), int(inwin.size), taps)
    del node_wl, node_fl

    def evaluate(prob, nw, tag):
        lp = LogPosterior(prob)
        th = torch.tensor(central_third(prob, nw), device='cuda')
        out = torch.empty(nw, dtype=torch.float64, device='cuda')
        lp.engine.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(2):
            lp.engine.lnprob_device(th.data_ptr(), nw, out.data_ptr())
        torch.cuda.synchronize()
        print(f'# {tag}: bytes_per_eval {lp.tables.bytes_per_eval()} nw {nw}', flush=True)
        return lp

    lp = evaluate(synthetic.make_problem_c4(), 32768, 'C4')
    # the same batch in the bracket-only mode (ms_set_fast)
    lp.engine.set_fast(True)
    th = torch.tensor(central_third(lp.problem, 32768), device='cuda')
    out = torch.empty(32768, dtype=torch.float64, device='cuda')
    for _ in range(2):
        lp.engine.lnprob_device(th.data_ptr(), 32768, out.data_ptr())
    torch.cuda.synchronize()
    lp.close()
    p1 = synthetic.make_problem(truth=synthetic.consistent_truth(3850.0, 3425.0, 0.12, 2.0732e-3))
    lp = evaluate(p1, 8192, 'C1')
    # the reference-sized ensemble: one fused iteration (split sort, propose, prep, stream x cluster of 8, accept, scatter)
    s = EnsembleSampler(50, 6, CudaBackend(lp.engine), seed=1)
    p0 = p1.truth * (1 + 0.01 * rng.standard_normal((50, 6))); p0[:, 2] = np.abs(p0[:, 2])
    s.initialize(p0)
    s.run_mcmc(2)
    torch.cuda.synchronize()
    lp.close()
    lp = evaluate(synthetic.make_problem_c5(), 16384, 'C5')
    lp.close()


def summarize(path):
    peak = 6455.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    rows = list(csv.reader(open(path, errors='replace')))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    ki, mi, vi, ui, ii = (hdr.index(n) for n in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
    launches = OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9,
                 'sector': 32.0}.get(u, 1.0)
        d = launches.setdefault(r[ii], {'name': r[ki]})
        d[r[mi]] = v * scale
    print('| # | kernel | grid context | duration [us] | DRAM read+write [MB] | DRAM GB/s | of HBM peak | L2->SM [MB] | L2->SM GB/s |')
    print('|---|---|---|---|---|---|---|---|---|')
    for k, d in launches.items():
        name = d['name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        if name.startswith(('at::', 'at_cuda', 'cub::', 'native::')) or 'elementwise' in name or 'Memset' in name:
            continue
        us = d.get('gpu__time_duration.sum', 0.0)
        dram = d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
        l2 = d.get('lts__t_sectors_srcunit_tex_op_read.sum', 0.0)
        gbs = dram / us / 1e3 if us else 0.0
        print(f"| {k} | `{name[:70]}` | | {us:.1f} | {dram / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.2f} | {l2 / 1e6:.1f} | {l2 / us / 1e3 if us else 0:.0f} |")


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == 'summarize':
        summarize(sys.argv[2])
    else:
        run()
