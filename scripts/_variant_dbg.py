import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from mcmc_spec_b200 import synthetic
from mcmc_spec_b200.posterior import LogPosterior
p = synthetic.make_problem_small()
lp = LogPosterior(p)
lp.engine.set_group(True)
rng = np.random.default_rng(3)
theta = np.asarray(p.truth) * (1 + 0.01 * rng.standard_normal((16384, 6)))
theta[:, 2] = np.abs(theta[:, 2])
big = lp(theta); big = lp(theta)
small = np.concatenate([lp(theta[k:k + 1024]) for k in range(0, 16384, 1024)])
small = np.concatenate([lp(theta[k:k + 1024]) for k in range(0, 16384, 1024)])
d = big != small
print('n diff', d.sum(), 'of', d.size, 'max rel', np.nanmax(np.abs(big-small)/np.abs(big)))
print(big[d][:5], small[d][:5])
parts_b = lp.parts(theta[:2048]) if hasattr(lp,'parts') else None
