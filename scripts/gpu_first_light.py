"""First-light GPU check: small-problem parity vs the oracle + a rough C4 timing."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
from mcmc_spec_b200 import synthetic as syn
from mcmc_spec_b200.posterior import LogPosterior
from oracle_adapter import to_oracle

p = syn.make_problem_small()
lp = LogPosterior(p)
o = to_oracle(p)
th = syn.sample_prior_box(p, 256, seed=3)
g = lp(th)
c = o.lnprob_batch(th)
fin = np.isfinite(c)
print('inf sets equal:', np.array_equal(np.isfinite(g), fin))
rel = np.abs(g[fin] - c[fin]) / np.abs(c[fin])
print('max rel err', rel.max(), 'n', fin.sum())
parts = lp.parts(th[:4])
for k in range(4):
    d = {}
    o.lnprob(th[k], d)
    print(k, 'chi2', parts['chi2_spec'][k], d.get('chi2_spec'), 'con', parts['contrast'][k], d.get('contrast'),
          'phot', parts['phot'][k][:2], None if 'phot' not in d else d['phot'][:2], 'lp', parts['lnprior'][k], d['lp'])

import torch
pc = syn.make_problem_c4()
t0 = time.time(); L = LogPosterior(pc); print('c4 create', time.time() - t0)
oc = to_oracle(pc)
thc = syn.sample_prior_box(pc, 4096, seed=5)
gc = L(thc)
cc = oc.lnprob_batch(thc[:64])
fin = np.isfinite(cc)
print('c4 max rel', (np.abs(gc[:64][fin] - cc[fin]) / np.abs(cc[fin])).max())
nw = 32768
thb = syn.sample_prior_box(pc, nw, seed=6, edge_cases=False)
d_th = torch.tensor(thb, device='cuda'); d_out = torch.empty(nw, dtype=torch.float64, device='cuda')
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    L.engine.lnprob_device(d_th.data_ptr(), nw, d_out.data_ptr())
    import ctypes
    torch.cuda.synchronize()
    # handle stream is non-blocking: sync it through the runtime
    ctypes.CDLL(None)
    dt = time.time() - t0
    print('c4 device eval', nw, 'walkers', dt, 's ->', nw / dt, 'evals/s', 'alg GB/s', nw / dt * L.tables.bytes_per_eval() / 1e9)
print(d_out[:4].cpu().numpy(), gc[:4])
