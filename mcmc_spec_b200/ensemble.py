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
contiguous shard of the active half, in place, and the shards are exchanged by ONE all-gather per
half step (NCCL over NVLink; the only collective of the path).  Random numbers -- the split
permutation, z, the partner and the acceptance draw -- come from a counter-based generator keyed
by (seed, iteration, split, slot in the half), so the chain does not depend on the number of
ranks.  Positions never leave the device; a half step is one C call (proposal folded into the preparation kernel,
evaluation, acceptance) plus, when sharded, the collective and the scatter of the gathered rows.

``backend`` supplies the five primitives on torch tensors.  The product backend is
``CudaBackend`` (the C ABI kernels); tests plug in a numpy mirror to exercise the sharding logic
under gloo on CPUs.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class CudaBackend:
    """The sampler primitives through libmcmcspec.so, enqueued on torch's current CUDA stream (so
    they order with the NCCL all-gather around them)."""

    def __init__(self, engine):
        self.engine = engine
        self.ndim = engine.ndim
        self.device = torch.device('cuda', engine.device)

    def _bind(self):
        self.engine.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def lnprob(self, theta, out=None):
        self._bind()
        theta = theta.contiguous()
        if out is None:
            out = torch.empty(theta.shape[0], dtype=torch.float64, device=theta.device)
        self.engine.lnprob_device(theta.data_ptr(), theta.shape[0], out.data_ptr())
        return out

    def split_keys(self, seed, iteration, n):
        self._bind()
        keys = torch.empty(n, dtype=torch.int64, device=self.device)
        self.engine.split_keys(seed, iteration, n, keys.data_ptr())
        return keys

    def propose(self, state, s_idx, c_idx, a, seed, iteration, split, gid0, q, logfac):
        self._bind()
        self.engine.stretch_propose(state.data_ptr(), s_idx.data_ptr(), s_idx.shape[0], c_idx.data_ptr(), c_idx.shape[0],
                                    a, seed, iteration, split, gid0, q.data_ptr(), logfac.data_ptr())

    def accept(self, state, s_idx, q, lnp_q, logfac, seed, iteration, split, gid0, packed):
        self._bind()
        self.engine.stretch_accept(state.data_ptr(), s_idx.data_ptr(), q.data_ptr(), lnp_q.data_ptr(), logfac.data_ptr(),
                                   s_idx.shape[0], seed, iteration, split, gid0, packed.data_ptr())

    def scatter(self, state, naccepted, idx, packed):
        self._bind()
        self.engine.stretch_scatter(state.data_ptr(), naccepted.data_ptr(), idx.data_ptr(), packed.data_ptr(),
                                    idx.shape[0])

    def halfstep(self, state, naccepted, s_idx, c_idx, a, seed, iteration, split, gid0, q, logfac, lnp_q, packed=None):
        """Propose, evaluate and accept one rank's share of a half step in one C call; ``packed`` None: in place."""
        self._bind()
        self.engine.stretch_halfstep(state.data_ptr(), naccepted.data_ptr(), s_idx.data_ptr(), s_idx.shape[0], c_idx.data_ptr(),
                                     c_idx.shape[0], a, seed, iteration, split, gid0, q.data_ptr(), logfac.data_ptr(),
                                     lnp_q.data_ptr(), None if packed is None else packed.data_ptr())

    def set_shard(self, rank, world):
        self.engine.set_group_shard(rank, world)

    def group_will_run(self, n):
        return self.engine.group_will_run(n)

    def evaluate(self, state, s_idx, c_idx, a, seed, iteration, split, gid0, q, logfac, lnp_q):
        """Propose for the whole half, evaluate this rank's groups (lnp_q is 0 for the other ranks' walkers)."""
        self._bind()
        self.engine.stretch_evaluate(state.data_ptr(), s_idx.data_ptr(), s_idx.shape[0], c_idx.data_ptr(), c_idx.shape[0], a, seed,
                                     iteration, split, gid0, q.data_ptr(), logfac.data_ptr(), lnp_q.data_ptr())

    def accept_inplace(self, state, naccepted, s_idx, q, lnp_q, logfac, seed, iteration, split, gid0):
        self._bind()
        self.engine.stretch_accept_inplace(state.data_ptr(), naccepted.data_ptr(), s_idx.data_ptr(), s_idx.shape[0], q.data_ptr(),
                                           lnp_q.data_ptr(), logfac.data_ptr(), seed, iteration, split, gid0)

    FUSED_MAX_WALKERS = 2048

    def iteration(self, state, naccepted, a, seed, iteration, randomize_split):
        """One whole iteration enqueued by a single C call (single rank, small ensembles)."""
        self._bind()
        self.engine.stretch_iteration(state.data_ptr(), naccepted.data_ptr(), state.shape[0], a, seed, iteration,
                                      randomize_split)


class EnsembleSampler:
    """Stretch-move ensemble sampler; attribute names follow emcee where they exist.

    nwalkers must be even and nwalkers/2 divisible by the world size.  ``group`` is the
    torch.distributed process group (None = default group; single process = no collective).
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
        self.chain = []                         # optional thinned copies of the state
        f64 = dict(dtype=torch.float64, device=self.device)
        self._q = torch.empty((self.shard, ndim), **f64)
        self._logfac = torch.empty(self.shard, **f64)
        self._lnp_q = torch.empty(self.shard, **f64)
        self._packed = torch.empty((self.shard, ndim + 2), **f64)
        self._full = self._packed if world_size == 1 else torch.empty((self.half, ndim + 2), **f64)
        self._fixed = [torch.arange(0, nwalkers, 2, device=self.device), torch.arange(1, nwalkers, 2, device=self.device)]
        self._side, self._next = None, None
        self._by_group = None                   # N > 1: shard the evaluation by groups of cell mates (decided at the first step)
        self.shard_by_group = True              # True: by the shard size (see step()); False: never; 'always': whenever possible
        self.fused = True                       # allow the one-call iteration where the backend offers it
        self._stored, self._iter0 = None, 0

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
        if self.world_size == 1:
            self.state = packed
        else:
            self.state = torch.empty((self.nwalkers, self.ndim + 1), dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(self.state, packed, group=self.group)
        self.iteration = 0
        self._iter0 = 0
        self.naccepted.zero_()

    @property
    def coords(self):
        return self.state[:, :self.ndim]

    @property
    def log_prob(self):
        return self.state[:, self.ndim]

    def checksum(self):
        """Order-independent exact fingerprint of the ensemble (positions and log-probs): the sum of the state's
        64-bit patterns modulo 2^64.  The random streams do not depend on the number of ranks and a walker's
        log-posterior does not depend on its batch, so the value is the same at 1, 2, 4 and 8 GPUs."""
        return int(self.state.contiguous().view(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF

    @property
    def acceptance_fraction(self):
        return (self.naccepted.to(torch.float64) / max(self.iteration - self._iter0, 1)).cpu().numpy()

    # ------------------------------------------------------------------ one iteration
    def _compute_halves(self, iteration):
        keys = self.backend.split_keys(self.seed, iteration, self.nwalkers)
        perm = torch.argsort(keys, stable=True)
        return [perm[0::2].contiguous(), perm[1::2].contiguous()]       # inds = arange(nw) % 2, shuffled

    def _halves(self):
        """Red/blue split of this iteration.  The permutation depends only on (seed, iteration), so on a
        GPU the one for the *next* iteration is computed on a side stream while this iteration's kernels
        run, which takes the 65 536-key sort off the critical path."""
        if not self.randomize_split:
            return self._fixed
        if self.device.type != 'cuda':
            return self._compute_halves(self.iteration)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        if self._next is None or self._next[0] != self.iteration:
            self._next = (self.iteration, self._compute_halves(self.iteration), None)
        it, halves, ev = self._next
        if ev is not None:
            main.wait_event(ev)
        for h in halves:
            h.record_stream(main)
        with torch.cuda.stream(self._side):
            nxt = self._compute_halves(self.iteration + 1)
            ev2 = torch.cuda.Event()
            ev2.record(self._side)
        self._next = (self.iteration + 1, nxt, ev2)
        return halves

    def step(self):
        """One ensemble iteration = two half steps; every walker gets one proposal."""
        be = self.backend
        if (self.world_size == 1 and hasattr(be, 'iteration') and self.nwalkers <= be.FUSED_MAX_WALKERS
                and self.fused and self.state.is_contiguous()):
            # small single-rank ensembles are launch bound: the whole iteration goes down in one call
            be.iteration(self.state, self.naccepted, self.a, self.seed, self.iteration, self.randomize_split)
            self.iteration += 1
            return
        halves = self._halves()
        if self._by_group is None:
            # Several ranks and a backend with the group kernel: every rank holds the whole ensemble, proposes for the whole
            # half and evaluates every world_size-th GROUP of cell mates (whole 32-walker groups per GPU instead of 1/N of
            # every cell's walkers, which leaves most groups part filled); the log-probs are summed over the ranks (each
            # walker's is non-zero on exactly one) and every rank applies the same acceptance test: no positions travel.
            # It pays once a rank's 1/N of the half no longer fills the GPU with groups of its own (one group per SM,
            # 32 walkers each): below that the part-filled groups of a walker shard cost a second round of CTAs, above
            # it the replicated preparation and the all-reduce cost more than the fuller groups save.
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count if self.device.type == 'cuda' else 0
            small = self.shard_by_group == 'always' or self.shard <= 32 * sms
            self._by_group = bool(self.world_size > 1 and self.shard_by_group and hasattr(be, 'evaluate')
                                  and small and be.group_will_run(self.half))
            if self._by_group:
                be.set_shard(self.rank, self.world_size)
                f64 = dict(dtype=torch.float64, device=self.device)
                self._qh = torch.empty((self.half, self.ndim), **f64)
                self._lfh = torch.empty(self.half, **f64)
                self._lnph = torch.empty(self.half, **f64)
        for split in (0, 1):
            s_idx, c_idx = halves[split], halves[1 - split]
            if self._by_group:
                be.evaluate(self.state, s_idx, c_idx, self.a, self.seed, self.iteration, split, 0, self._qh, self._lfh, self._lnph)
                dist.all_reduce(self._lnph, op=dist.ReduceOp.SUM, group=self.group)
                be.accept_inplace(self.state, self.naccepted, s_idx, self._qh, self._lnph, self._lfh, self.seed, self.iteration,
                                  split, 0)
                continue
            lo = self.rank * self.shard
            mine = s_idx[lo:lo + self.shard]
            if hasattr(be, 'halfstep'):
                # one C call: proposal folded into the preparation kernel, evaluation, acceptance -- in place when this
                # rank owns the whole half, else into the packed rows of the all-gather
                if self.world_size == 1:
                    be.halfstep(self.state, self.naccepted, mine, c_idx, self.a, self.seed, self.iteration, split, lo,
                                self._q, self._logfac, self._lnp_q)
                    continue
                be.halfstep(self.state, self.naccepted, mine, c_idx, self.a, self.seed, self.iteration, split, lo,
                            self._q, self._logfac, self._lnp_q, self._packed)
            else:
                be.propose(self.state, mine, c_idx, self.a, self.seed, self.iteration, split, lo, self._q, self._logfac)
                be.lnprob(self._q, out=self._lnp_q)
                be.accept(self.state, mine, self._q, self._lnp_q, self._logfac, self.seed, self.iteration, split, lo,
                          self._packed)
            if self.world_size > 1:
                dist.all_gather_into_tensor(self._full, self._packed, group=self.group)   # slot order
            be.scatter(self.state, self.naccepted, s_idx, self._full)
        self.iteration += 1

    def run_mcmc(self, nsteps, thin=0):
        for n in range(nsteps):
            self.step()
            if thin and n % thin == 0:
                self.chain.append(self.state.detach().cpu().numpy().copy())
        return self.state

    # ------------------------------------------------------------------ the emcee calls the reference makes
    # (mft6.py:1491-1527: sampler.sample(pos, iterations=n) yielding a state with .coords, get_last_sample(),
    #  reset(), get_autocorr_time(quiet=True), acceptance_fraction, chain)
    def sample(self, initial_state=None, iterations=1, store=True):
        """Generator over iterations, like ``emcee.EnsembleSampler.sample``: yields the sampler itself after every
        step (``.coords`` and ``.log_prob`` are the current ensemble).  ``initial_state``: positions [nw, ndim], or
        the object returned by ``get_last_sample()``.  With ``store`` the positions of every step are kept (on the
        device) for ``chain`` / ``get_autocorr_time``."""
        if initial_state is not None and initial_state is not self:
            pos = initial_state.coords if hasattr(initial_state, 'coords') else initial_state
            if torch.is_tensor(pos):
                pos = pos.detach().clone()
            it = self.iteration
            self.initialize(pos)
            self.iteration = self._iter0 = it        # the random streams go on, they never repeat an iteration index
        if store and self._stored is None:
            self._stored = []
        for _ in range(int(iterations)):
            self.step()
            if store:
                self._stored.append(self.coords.detach().clone())
            yield self

    def get_last_sample(self):
        return self

    def reset(self):
        """Forget the stored chain and the acceptance counts; positions stay (emcee ``reset``)."""
        self._stored = None
        self.naccepted.zero_()
        self._iter0 = self.iteration

    def get_chain(self, flat=False):
        """Stored positions [nsteps, nwalkers, ndim] (``flat``: [nsteps * nwalkers, ndim]) as a numpy array."""
        if not self._stored:
            return np.empty((0, self.nwalkers, self.ndim))
        x = torch.stack(self._stored).cpu().numpy()
        return x.reshape(-1, self.ndim) if flat else x

    @property
    def stored_chain(self):
        """emcee's legacy ``sampler.chain``: [nwalkers, nsteps, ndim] (what mft6.py:1527 flattens into samples.txt)."""
        return np.swapaxes(self.get_chain(), 0, 1)

    def get_autocorr_time(self, quiet=True, c=5):
        from .fit import integrated_autocorr_time
        if not self._stored:
            return np.full(self.ndim, np.nan)
        return integrated_autocorr_time(torch.stack(self._stored), c=c)

    # ------------------------------------------------------------------ emcee-style file dumps
    def save_coords(self, path):
        """np.savetxt of the walker coordinates, one row per walker (mft6.py:1494-1510 format)."""
        with open(path, 'ab') as f:
            f.write(b"\n")
            np.savetxt(f, self.coords.detach().cpu().numpy())
