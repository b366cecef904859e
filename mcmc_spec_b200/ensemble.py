"""Affine-invariant ensemble sampler around the device log-posterior, sharded by walker.

What the reference drives (commented block, /root/reference/mft6.py:1491-1527):
``emcee.EnsembleSampler(nwalkers, ndim, logposterior, ...)`` with the default move -- emcee >= 3
``RedBlueMove`` / ``StretchMove`` (a = 2).  emcee is not vendored by the reference and is not
installed here, so its published algorithm is restated:

    per iteration:  inds = arange(nw) % 2, shuffled           (randomize_split=True)
    for split in (0, 1):   S = walkers with inds == split, C = the rest (already updated)
        z = ((a - 1) u + 1)^2 / a ;  j ~ randint(|C|) ;  q = c_j - (c_j - s) z
        accept if  ln u' < (ndim - 1) ln z + lnp(q) - lnp(s)

B200 layout: the full ensemble state (positions + log-probs, nw x (ndim+1) doubles, 3.7 MB at
65 536 walkers) is replicated on every rank; each rank proposes / evaluates / accepts its
contiguous shard of the active half and the shards are exchanged by ONE all-gather per half step
(NCCL over NVLink; the only collective of the path).  Random numbers come from a counter-based
generator keyed by (seed, iteration, split, slot in the half), so the chain does not depend on the
number of ranks.  Positions never leave the device.

``backend`` supplies propose / lnprob / accept on torch tensors.  The product backend is
``CudaBackend`` (the C ABI kernels); tests plug in a numpy backend to exercise the sharding logic
under gloo on CPUs.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class CudaBackend:
    """propose / lnprob / accept through libmcmcspec.so, enqueued on torch's current CUDA stream
    (so they order with the index/gather plumbing and the NCCL all-gather around them)."""

    def __init__(self, engine):
        self.engine = engine
        self.ndim = engine.ndim
        self.device = torch.device('cuda', engine.device)

    def _bind(self):
        self.engine.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def lnprob(self, theta):
        self._bind()
        theta = theta.contiguous()
        out = torch.empty(theta.shape[0], dtype=torch.float64, device=theta.device)
        self.engine.lnprob_device(theta.data_ptr(), theta.shape[0], out.data_ptr())
        return out

    def propose(self, x, c, a, seed, iteration, split, gid0):
        self._bind()
        x = x.contiguous(); c = c.contiguous()
        q = torch.empty_like(x)
        logfac = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
        self.engine.stretch_propose(x.data_ptr(), x.shape[0], c.data_ptr(), c.shape[0], a, seed, iteration, split,
                                    gid0, q.data_ptr(), logfac.data_ptr())
        return q, logfac

    def accept(self, x, lnp_x, q, lnp_q, logfac, seed, iteration, split, gid0, accepted):
        self._bind()
        self.engine.stretch_accept(x.data_ptr(), lnp_x.data_ptr(), q.data_ptr(), lnp_q.data_ptr(), logfac.data_ptr(),
                                   x.shape[0], seed, iteration, split, gid0, accepted.data_ptr())


class EnsembleSampler:
    """Stretch-move ensemble sampler; attribute names follow emcee where they exist.

    nwalkers must be even and nwalkers/2 divisible by the world size.  ``group`` is the
    torch.distributed process group (None = single process, no collective).
    """

    def __init__(self, nwalkers, ndim, backend, a=2.0, seed=0, group=None, rank=0, world_size=1,
                 randomize_split=True, device=None):
        if nwalkers % 2 or (nwalkers // 2) % world_size:
            raise ValueError("nwalkers must be even and nwalkers/2 divisible by the number of ranks")
        self.nwalkers, self.ndim, self.backend = nwalkers, ndim, backend
        self.a, self.seed = float(a), int(seed)
        self.group, self.rank, self.world_size = group, rank, world_size
        self.randomize_split = randomize_split
        self.device = device if device is not None else getattr(backend, 'device', torch.device('cpu'))
        self.half = nwalkers // 2
        self.shard = self.half // world_size
        self.iteration = 0
        self.state = None                       # [nw, ndim+1]: positions | log-prob (replicated)
        self.naccepted = torch.zeros(nwalkers, dtype=torch.int32, device=self.device)
        self._perm_gen = torch.Generator(device='cpu')
        self.chain = []                         # optional thinned copies of (coords, log_prob)

    # ------------------------------------------------------------------ state
    def initialize(self, p0):
        """p0 [nw, ndim] (numpy or tensor): evaluates lnprob of the local shards and gathers."""
        p0 = torch.as_tensor(np.asarray(p0) if not torch.is_tensor(p0) else p0, dtype=torch.float64).to(self.device)
        if p0.shape != (self.nwalkers, self.ndim):
            raise ValueError("p0 must be [nwalkers, ndim]")
        per = self.nwalkers // self.world_size
        mine = p0[self.rank * per:(self.rank + 1) * per].contiguous()
        lnp = self.backend.lnprob(mine)
        packed = torch.cat([mine, lnp[:, None]], dim=1).contiguous()
        self.state = self._all_gather(packed)
        self.iteration = 0
        self.naccepted.zero_()

    @property
    def coords(self):
        return self.state[:, :self.ndim]

    @property
    def log_prob(self):
        return self.state[:, self.ndim]

    @property
    def acceptance_fraction(self):
        return (self.naccepted.to(torch.float64) / max(self.iteration, 1)).cpu().numpy()

    # ------------------------------------------------------------------ plumbing
    def _all_gather(self, packed):
        if self.world_size == 1:
            return packed
        out = torch.empty((self.world_size * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(out, packed, group=self.group)
        return out

    def _split_indices(self):
        nw = self.nwalkers
        if self.randomize_split:
            self._perm_gen.manual_seed((self.seed * 1000003 + self.iteration) & 0x7fffffffffffffff)
            perm = torch.randperm(nw, generator=self._perm_gen)
        else:
            perm = torch.arange(nw)
        inds = torch.arange(nw) % 2
        halves = [perm[inds == 0], perm[inds == 1]]
        return [h.to(self.device, non_blocking=True) for h in halves]

    # ------------------------------------------------------------------ one iteration
    def step(self):
        """One ensemble iteration = two half steps; every walker gets one proposal."""
        halves = self._split_indices()
        nd = self.ndim
        for split in (0, 1):
            s_idx, c_idx = halves[split], halves[1 - split]
            lo = self.rank * self.shard
            mine = s_idx[lo:lo + self.shard]
            cur = self.state.index_select(0, mine)                      # [shard, ndim+1]
            x = cur[:, :nd].contiguous()
            lnp_x = cur[:, nd].contiguous()
            comp = self.state.index_select(0, c_idx)[:, :nd].contiguous()
            q, logfac = self.backend.propose(x, comp, self.a, self.seed, self.iteration, split, lo)
            lnp_q = self.backend.lnprob(q)
            acc = torch.zeros(self.shard, dtype=torch.int32, device=self.device)
            self.backend.accept(x, lnp_x, q, lnp_q, logfac, self.seed, self.iteration, split, lo, acc)
            packed = torch.cat([x, lnp_x[:, None], acc.to(torch.float64)[:, None]], dim=1).contiguous()
            full = self._all_gather(packed)                              # [half, ndim+2] in slot order
            self.state.index_copy_(0, s_idx, full[:, :nd + 1])
            self.naccepted.index_add_(0, s_idx, full[:, nd + 1].to(torch.int32))
        self.iteration += 1

    def run_mcmc(self, nsteps, thin=0):
        for n in range(nsteps):
            self.step()
            if thin and n % thin == 0:
                self.chain.append(self.state.detach().cpu().numpy().copy())
        return self.state

    # ------------------------------------------------------------------ emcee-style file dumps
    def save_coords(self, path):
        """np.savetxt of the walker coordinates, one row per walker (mft6.py:1494-1510 format)."""
        with open(path, 'ab') as f:
            f.write(b"\n")
            np.savetxt(f, self.coords.detach().cpu().numpy())
