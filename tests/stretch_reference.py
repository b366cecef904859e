"""numpy mirror of the device stretch-move kernels and their counter-based generator (TEST code).

Restates emcee's StretchMove arithmetic (emcee/moves/stretch.py, not vendored by the reference):
zz = ((a - 1) * u + 1)**2 / a;  factors = (ndim - 1) * log(zz);  q = c[j] - (c[j] - s) * zz;
accept where log(u') < factors + lnp(q) - lnp(s).
"""
import numpy as np
import torch

M = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over='ignore'):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def counter_uniform(seed, iteration, split, purpose, gid):
    k = mix64(np.uint64(seed))
    k = mix64(k ^ np.uint64(iteration))
    k = mix64(k ^ np.uint64(split * 4 + purpose))
    r = mix64(k ^ mix64(np.asarray(gid, dtype=np.uint64)))
    return ((r >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def split_keys(seed, iteration, n):
    k = mix64(np.uint64(seed) ^ np.uint64(0x5851F42D4C957F2D))
    k = mix64(k ^ np.uint64(iteration))
    return (mix64(k ^ mix64(np.arange(n, dtype=np.uint64))) >> np.uint64(1)).astype(np.int64)


def propose(x, c, a, seed, iteration, split, gid0):
    n_s, ndim = x.shape
    gid = gid0 + np.arange(n_s)
    u = counter_uniform(seed, iteration, split, 0, gid)
    root = (a - 1.0) * u + 1.0
    z = root * root / a
    j = np.minimum((counter_uniform(seed, iteration, split, 1, gid) * c.shape[0]).astype(np.int64), c.shape[0] - 1)
    cj = c[j]
    q = cj - (cj - x) * z[:, None]
    return q, (ndim - 1) * np.log(z)


def accept_mask(lnp_x, lnp_q, logfac, seed, iteration, split, gid0):
    gid = gid0 + np.arange(lnp_x.shape[0])
    u = counter_uniform(seed, iteration, split, 2, gid)
    with np.errstate(invalid='ignore'):
        return np.log(u) < logfac + lnp_q - lnp_x


class NumpyBackend:
    """EnsembleSampler backend on CPU tensors with a user lnprob(theta[n, ndim]) -> [n] (numpy)."""

    def __init__(self, lnprob_fn, ndim):
        self.fn, self.ndim = lnprob_fn, ndim
        self.device = torch.device('cpu')

    def lnprob(self, theta, out=None):
        res = torch.from_numpy(np.asarray(self.fn(theta.numpy().copy()), dtype=np.float64))
        if out is None:
            return res
        out.copy_(res)
        return out

    def split_keys(self, seed, iteration, n):
        return torch.from_numpy(split_keys(seed, iteration, n))

    def propose(self, state, s_idx, c_idx, a, seed, iteration, split, gid0, q, logfac):
        nd = self.ndim
        st = state.numpy()
        qq, lf = propose(st[s_idx.numpy(), :nd], st[c_idx.numpy(), :nd], a, seed, iteration, split, gid0)
        q.copy_(torch.from_numpy(qq)); logfac.copy_(torch.from_numpy(lf))

    def accept(self, state, s_idx, q, lnp_q, logfac, seed, iteration, split, gid0, packed):
        nd = self.ndim
        cur = state.numpy()[s_idx.numpy()]
        m = accept_mask(cur[:, nd], lnp_q.numpy(), logfac.numpy(), seed, iteration, split, gid0)
        out = np.concatenate([cur, np.zeros((cur.shape[0], 1))], axis=1)
        out[m, :nd] = q.numpy()[m]; out[m, nd] = lnp_q.numpy()[m]; out[m, nd + 1] = 1.0
        packed.copy_(torch.from_numpy(out))

    def scatter(self, state, naccepted, idx, packed):
        nd = self.ndim
        state[idx] = packed[:, :nd + 1]
        naccepted[idx] += packed[:, nd + 1].to(torch.int32)


class GroupShardedNumpyBackend(NumpyBackend):
    """The numpy backend with the entry points of the evaluation sharded by groups (CudaBackend.evaluate /
    accept_inplace / set_shard / group_will_run): every rank proposes for the whole half, evaluates the walkers of
    every world-th block of 4 ("groups"), leaves the others at 0; the sampler sums over the ranks."""

    GROUP = 4

    def __init__(self, lnprob_fn, ndim):
        super().__init__(lnprob_fn, ndim)
        self.rank, self.world = 0, 1

    def set_shard(self, rank, world):
        self.rank, self.world = rank, world

    def group_will_run(self, n):
        return True

    def evaluate(self, state, s_idx, c_idx, a, seed, iteration, split, gid0, q, logfac, lnp_q):
        self.propose(state, s_idx, c_idx, a, seed, iteration, split, gid0, q, logfac)
        n = q.shape[0]
        mine = (np.arange(n) // self.GROUP) % self.world == self.rank
        out = np.zeros(n)
        if mine.any():
            out[mine] = np.asarray(self.fn(q.numpy()[mine].copy()), dtype=np.float64)
        lnp_q.copy_(torch.from_numpy(out))

    def accept_inplace(self, state, naccepted, s_idx, q, lnp_q, logfac, seed, iteration, split, gid0):
        nd = self.ndim
        idx = s_idx.numpy()
        cur = state.numpy()[idx]
        m = accept_mask(cur[:, nd], lnp_q.numpy(), logfac.numpy(), seed, iteration, split, gid0)
        st = state.numpy()
        rows = idx[m]
        st[rows, :nd] = q.numpy()[m]; st[rows, nd] = lnp_q.numpy()[m]
        naccepted.numpy()[rows] += 1

