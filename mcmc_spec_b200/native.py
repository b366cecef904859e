"""ctypes binding of libmcmcspec.so (C ABI in include/mcmcspec.h).

This is the stub a maintainer of the reference would add next to mft6.py: it mirrors
``ms_config`` field for field, loads the in-tree library and fails loudly when the library (or a
CUDA device) is missing -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .tables import DeviceTables, TILE

MS_ABI_VERSION = 2
MS_MAX_CONTRAST, MS_MAX_PHOT, MS_MAX_NDIM = 4, 12, 8
MS_HOST, MS_DEVICE = 0, 1
MS_W_PRIOR_BOUNDS, MS_W_OFF_ISOCHRONE, MS_W_OFF_GRID, MS_W_NONFINITE = 1, 2, 4, 8

# MCMCSPEC_LIB: another build of the same library (kernel tuning variants, scripts/tune_variants.py)
_LIB_PATH = os.environ.get('MCMCSPEC_LIB') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libmcmcspec.so')

EXPORTS = ['ms_config_size', 'ms_abi_version', 'ms_create', 'ms_create_shared', 'ms_destroy', 'ms_last_error', 'ms_stream',
           'ms_device', 'ms_lnprob', 'ms_lnprob_parts', 'ms_model_spectrum', 'ms_component_mags', 'ms_set_optimizer_data', 'ms_optimizer_chisq', 'ms_broaden_rows', 'ms_build_grid',
           'ms_split_keys', 'ms_stretch_propose', 'ms_stretch_accept', 'ms_stretch_scatter', 'ms_stretch_iteration', 'ms_stretch_halfstep', 'ms_stretch_evaluate', 'ms_stretch_accept_inplace', 'ms_set_group_shard', 'ms_group_will_run', 'ms_launch_count', 'ms_set_stream', 'ms_use_own_stream', 'ms_set_cluster', 'ms_set_pairing', 'ms_set_fast', 'ms_set_cell_order', 'ms_set_group', 'ms_group_supported', 'ms_short_circuited', 'ms_set_timing',
           'ms_kernel_time']

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class MsConfig(C.Structure):
    _fields_ = [
        ('abi_version', C.c_int32), ('device', C.c_int32),
        ('n_teff', C.c_int32), ('n_logg', C.c_int32), ('n_feh', C.c_int32), ('n_reg', C.c_int32),
        ('row_len', C.c_int64), ('reg_lo', C.c_int64),
        ('teff_nodes', _dp), ('logg_nodes', _dp), ('feh_nodes', _dp), ('grid', _dp), ('feh', C.c_double),
        ('n_data', C.c_int32), ('n_tiles', C.c_int32),
        ('rebin_hi', _ip), ('rebin_t', _dp), ('ext_lo', _dp), ('ext_hi', _dp),
        ('data_flux', _dp), ('data_err', _dp), ('data_x', _dp),
        ('poly_basis', C.c_double * 9), ('data_median', C.c_double),
        ('n_contrast', C.c_int32), ('n_phot', C.c_int32),
        ('c_tile_lo', _ip), ('c_tile_hi', _ip), ('c_offset', _lp), ('c_weights', _dp),
        ('c_weights_len', C.c_int64), ('c_mag', _dp), ('c_err', _dp),
        ('p_tile_lo', _ip), ('p_tile_hi', _ip), ('p_offset', _lp), ('p_weights', _dp),
        ('p_weights_len', C.c_int64), ('p_norm', _dp), ('p_ext', _dp), ('p_mag', _dp), ('p_err', _dp),
        ('n_iso', C.c_int32), ('n_dm', C.c_int32), ('n_dust_samples', C.c_int32), ('nspec', C.c_int32),
        ('iso_teff', _dp), ('iso_logg', _dp), ('iso_lum', _dp), ('dust_dm_edges', _dp), ('dust_samples', _dp),
        ('tmin', C.c_double), ('tmax', C.c_double),
        ('prior_mu', C.c_double * MS_MAX_NDIM), ('prior_sig', C.c_double * MS_MAX_NDIM),
        ('ext_flag', C.c_int32), ('dist_fit', C.c_int32), ('rad_prior', C.c_int32), ('flavour', C.c_int32),
        ('feh_free', C.c_int32), ('reserved0', C.c_int32),
    ]


class NativeError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """Load libmcmcspec.so and declare prototypes.  Raises NativeError if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or _LIB_PATH
    if not os.path.exists(path):
        raise NativeError(f"{path} is missing: build it with `python -m mcmc_spec_b200.build` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.ms_config_size.restype = C.c_int64
    lib.ms_abi_version.restype = C.c_int32
    lib.ms_create.argtypes = [C.POINTER(MsConfig), C.POINTER(vp)]
    lib.ms_create_shared.argtypes = [C.POINTER(MsConfig), vp, C.POINTER(vp)]
    lib.ms_destroy.argtypes = [vp]
    lib.ms_last_error.argtypes = [vp]
    lib.ms_last_error.restype = C.c_char_p
    lib.ms_stream.argtypes = [vp]
    lib.ms_stream.restype = vp
    lib.ms_device.argtypes = [vp]
    lib.ms_lnprob.argtypes = [vp, vp, C.c_int64, vp, C.c_int]
    lib.ms_lnprob_parts.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    lib.ms_model_spectrum.argtypes = [vp, vp, C.c_int64, vp, C.c_int]
    lib.ms_component_mags.argtypes = [vp, vp, C.c_int64, vp, C.c_int]
    lib.ms_set_optimizer_data.argtypes = [vp, vp, C.c_int32]
    lib.ms_optimizer_chisq.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64, vp]
    lib.ms_broaden_rows.argtypes = [C.c_int, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp, C.c_int32]
    lib.ms_build_grid.argtypes = [C.c_int, vp, vp, vp, C.c_int64, vp, C.c_int64, C.c_int64, C.c_int64, vp, C.c_int32, vp]
    lib.ms_split_keys.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int64, vp]
    lib.ms_stretch_propose.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int64, C.c_int32, C.c_double, C.c_uint64,
                                       C.c_uint64, C.c_int32, C.c_int64, vp, vp]
    lib.ms_stretch_accept.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64,
                                      C.c_int32, C.c_int64, vp]
    lib.ms_stretch_scatter.argtypes = [vp, vp, vp, vp, vp, C.c_int64, C.c_int32]
    lib.ms_stretch_iteration.argtypes = [vp, vp, vp, C.c_int64, C.c_double, C.c_uint64, C.c_uint64, C.c_int32]
    lib.ms_stretch_halfstep.argtypes = [vp, vp, vp, vp, C.c_int64, vp, C.c_int64, C.c_double, C.c_uint64, C.c_uint64, C.c_int32,
                                        C.c_int64, vp, vp, vp, vp]
    lib.ms_launch_count.argtypes = [vp]
    lib.ms_launch_count.restype = C.c_int64
    lib.ms_set_stream.argtypes = [vp, vp]
    lib.ms_use_own_stream.argtypes = [vp]
    lib.ms_set_cluster.argtypes = [vp, C.c_int]
    lib.ms_set_pairing.argtypes = [vp, C.c_int]
    lib.ms_set_fast.argtypes = [vp, C.c_int]
    lib.ms_set_cell_order.argtypes = [vp, C.c_int]
    lib.ms_set_group.argtypes = [vp, C.c_int, C.c_int64]
    lib.ms_set_group_shard.argtypes = [vp, C.c_int, C.c_int]
    lib.ms_group_will_run.argtypes = [vp, C.c_int64]
    lib.ms_stretch_evaluate.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int64, C.c_double, C.c_uint64, C.c_uint64, C.c_int32,
                                        C.c_int64, vp, vp, vp]
    lib.ms_stretch_accept_inplace.argtypes = [vp, vp, vp, vp, C.c_int64, vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int32, C.c_int64]
    lib.ms_group_supported.argtypes = [vp]
    lib.ms_short_circuited.argtypes = [vp, C.POINTER(C.c_int64), C.c_int]
    lib.ms_set_timing.argtypes = [vp, C.c_int]
    lib.ms_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    for name in EXPORTS:
        getattr(lib, name)
    if lib.ms_config_size() != C.sizeof(MsConfig):
        raise NativeError(f"ms_config layout mismatch: library {lib.ms_config_size()} bytes, "
                          f"binding {C.sizeof(MsConfig)} bytes")
    if lib.ms_abi_version() != MS_ABI_VERSION:
        raise NativeError("libmcmcspec.so ABI version mismatch")
    if path == _LIB_PATH:
        _lib = lib
    return lib


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def make_config(t: DeviceTables, device=0):
    """DeviceTables -> (MsConfig, keepalive list of the numpy arrays it points into)."""
    keep = []

    def dbl(a):
        a = np.ascontiguousarray(a, dtype=np.float64); keep.append(a); return _p(a, C.c_double)

    def i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32); keep.append(a); return _p(a, C.c_int32)

    def i64(a):
        a = np.ascontiguousarray(a, dtype=np.int64); keep.append(a); return _p(a, C.c_int64)

    cfg = MsConfig()
    cfg.abi_version, cfg.device = MS_ABI_VERSION, int(device)
    cfg.n_teff, cfg.n_logg, cfg.n_feh = len(t.teff_nodes), len(t.logg_nodes), len(t.feh_nodes)
    cfg.n_reg, cfg.row_len, cfg.reg_lo = t.n_reg, t.row_len, t.reg_lo
    cfg.teff_nodes, cfg.logg_nodes, cfg.feh_nodes = dbl(t.teff_nodes), dbl(t.logg_nodes), dbl(t.feh_nodes)
    grid = np.ascontiguousarray(t.grid, dtype=np.float64)
    if grid.shape[-1] != t.row_len or grid.size != cfg.n_teff * cfg.n_logg * cfg.n_feh * t.row_len:
        raise ValueError("grid shape does not match the node lists")
    keep.append(grid)
    cfg.grid = _p(grid, C.c_double)
    cfg.feh = t.feh
    cfg.n_data, cfg.n_tiles = t.n_data, t.n_tiles
    cfg.rebin_hi, cfg.rebin_t = i32(t.rebin_hi), dbl(t.rebin_t)
    cfg.ext_lo, cfg.ext_hi = dbl(t.ext_lo), dbl(t.ext_hi)
    cfg.data_flux, cfg.data_err, cfg.data_x = dbl(t.data_flux), dbl(t.data_err), dbl(t.data_x)
    cfg.poly_basis = (C.c_double * 9)(*np.asarray(t.poly_basis, dtype=np.float64).ravel())
    cfg.data_median = t.data_median
    cfg.n_contrast, cfg.n_phot = t.contrast.n, t.phot.n
    if cfg.n_contrast > MS_MAX_CONTRAST or cfg.n_phot > MS_MAX_PHOT:
        raise ValueError("too many filters for the compiled kernels")
    cfg.c_tile_lo, cfg.c_tile_hi, cfg.c_offset = i32(t.contrast.tile_lo), i32(t.contrast.tile_hi), i64(t.contrast.offset)
    cfg.c_weights, cfg.c_weights_len = dbl(t.contrast.weights), int(t.contrast.weights.size)
    cfg.c_mag, cfg.c_err = dbl(t.c_mag), dbl(t.c_err)
    cfg.p_tile_lo, cfg.p_tile_hi, cfg.p_offset = i32(t.phot.tile_lo), i32(t.phot.tile_hi), i64(t.phot.offset)
    cfg.p_weights, cfg.p_weights_len = dbl(t.phot.weights), int(t.phot.weights.size)
    cfg.p_norm, cfg.p_ext, cfg.p_mag, cfg.p_err = dbl(t.p_norm), dbl(t.p_ext), dbl(t.p_mag), dbl(t.p_err)
    cfg.n_iso, cfg.n_dm, cfg.n_dust_samples, cfg.nspec = len(t.iso_teff), len(t.dust_dm_edges), t.dust_samples.shape[0], t.nspec
    cfg.iso_teff, cfg.iso_logg, cfg.iso_lum = dbl(t.iso_teff), dbl(t.iso_logg), dbl(t.iso_lum)
    cfg.dust_dm_edges, cfg.dust_samples = dbl(t.dust_dm_edges), dbl(t.dust_samples)
    cfg.tmin, cfg.tmax = t.tmin, t.tmax
    mu = np.zeros(MS_MAX_NDIM); sig = np.ones(MS_MAX_NDIM)
    mu[:len(t.prior_mu)] = t.prior_mu; sig[:len(t.prior_sig)] = t.prior_sig
    cfg.prior_mu = (C.c_double * MS_MAX_NDIM)(*mu)
    cfg.prior_sig = (C.c_double * MS_MAX_NDIM)(*sig)
    cfg.ext_flag, cfg.dist_fit, cfg.rad_prior = t.ext_flag, t.dist_fit, t.rad_prior
    cfg.flavour = int(getattr(t, 'flavour', 0))
    cfg.feh_free = int(getattr(t, 'feh_free', 0))
    assert TILE == 1024
    return cfg, keep


class Engine:
    """Owns one ``ms_handle`` (one GPU, one stream).  Thin, typed wrapper over the C ABI."""

    def __init__(self, tables: DeviceTables, device=0, lib=None, grid_owner=None):
        """``grid_owner``: another Engine on the same device whose device copy of the model grid is
        reused (ms_create_shared); it is kept alive for as long as this engine lives."""
        self.lib = lib or load_library()
        self._grid_owner = grid_owner
        self.tables = tables
        self.ndim = 2 * tables.nspec + 2 + int(getattr(tables, 'feh_free', 0))      # columns of theta
        self.n_data, self.n_contrast, self.n_phot, self.nspec = tables.n_data, tables.contrast.n, tables.phot.n, tables.nspec
        cfg, keep = make_config(tables, device)
        h = C.c_void_p()
        if grid_owner is None:
            rc = self.lib.ms_create(C.byref(cfg), C.byref(h))
        else:
            rc = self.lib.ms_create_shared(C.byref(cfg), grid_owner._h, C.byref(h))
        del keep
        if rc != 0:
            raise NativeError(f"ms_create failed ({rc}): {self.lib.ms_last_error(None).decode()}")
        self._h = h

    # -- lifetime
    def close(self):
        if getattr(self, '_h', None):
            self.lib.ms_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise NativeError(f"{what} failed ({rc}): {self.lib.ms_last_error(self._h).decode()}")

    @property
    def stream(self):
        return self.lib.ms_stream(self._h)

    @property
    def device(self):
        return self.lib.ms_device(self._h)

    @property
    def launch_count(self):
        return int(self.lib.ms_launch_count(self._h))

    def set_stream(self, stream_ptr):
        """Enqueue on a caller-owned cudaStream_t (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        self._check(self.lib.ms_set_stream(self._h, stream_ptr), 'ms_set_stream')

    def use_own_stream(self):
        self._check(self.lib.ms_use_own_stream(self._h), 'ms_use_own_stream')

    def set_cluster(self, ctas_per_walker):
        """CTAs per walker: 0 automatic, 1 never, 2/4/8 forced (thread-block clusters for small ensembles)."""
        self._check(self.lib.ms_set_cluster(self._h, int(ctas_per_walker)), 'ms_set_cluster')

    def set_pairing(self, enabled):
        """Two cell-mates per CTA for large batches: True / False / None (automatic, the default); values are
        bit-identical either way."""
        self._check(self.lib.ms_set_pairing(self._h, -1 if enabled is None else int(bool(enabled))), 'ms_set_pairing')

    def set_fast(self, enabled):
        """Bracket-only evaluation (reads 2 model pixels per observed pixel and per-row band fluxes instead of
        streaming whole rows); off by default."""
        self._check(self.lib.ms_set_fast(self._h, int(bool(enabled))), 'ms_set_fast')

    def set_cell_order(self, mode):
        """Launch order of large batches: True grouped by grid cell, False natural, None automatic (default)."""
        self._check(self.lib.ms_set_cell_order(self._h, -1 if mode is None else int(bool(mode))), 'ms_set_cell_order')

    def set_group(self, mode, min_walkers=0):
        """Group evaluation of large posterior batches (cell-mates share one TMA-staged stream of their grid rows):
        True whenever the tables fit, False never, None automatic (batches >= ``min_walkers``, default 1024)."""
        self._check(self.lib.ms_set_group(self._h, -1 if mode is None else int(bool(mode)), int(min_walkers)), 'ms_set_group')

    @property
    def group_supported(self):
        return bool(self.lib.ms_group_supported(self._h))

    def short_circuited(self, reset=False):
        """Evaluations cut short by the prior box / tables (no grid row read) since creation or the last reset."""
        c = C.c_int64(0)
        self._check(self.lib.ms_short_circuited(self._h, C.byref(c), int(bool(reset))), 'ms_short_circuited')
        return int(c.value)

    def set_timing(self, enabled):
        self._check(self.lib.ms_set_timing(self._h, int(bool(enabled))), 'ms_set_timing')

    def kernel_time(self):
        """(total ms, launches, walkers) of the streaming kernel since the last call; resets."""
        ms, nl, nw = C.c_double(), C.c_int64(), C.c_int64()
        self._check(self.lib.ms_kernel_time(self._h, C.byref(ms), C.byref(nl), C.byref(nw)), 'ms_kernel_time')
        return ms.value, nl.value, nw.value

    # -- host-buffer calls (numpy in, numpy out)
    def lnprob_host(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if theta.ndim != 2 or theta.shape[1] != self.ndim:
            raise ValueError(f"theta must be [nw, {self.ndim}]")
        out = np.empty(theta.shape[0], dtype=np.float64)
        self._check(self.lib.ms_lnprob(self._h, theta.ctypes.data, theta.shape[0], out.ctypes.data, MS_HOST), 'ms_lnprob')
        return out

    def lnprob_parts_host(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        nw = theta.shape[0]
        out = dict(lnprob=np.empty(nw), chi2_spec=np.empty(nw), contrast=np.empty((nw, self.n_contrast)),
                   phot=np.empty((nw, self.n_phot)), lnprior=np.empty(nw), logg=np.empty((nw, self.nspec)),
                   status=np.empty(nw, dtype=np.int32))
        self._check(self.lib.ms_lnprob_parts(
            self._h, theta.ctypes.data, nw, out['lnprob'].ctypes.data, out['chi2_spec'].ctypes.data,
            out['contrast'].ctypes.data, out['phot'].ctypes.data, out['lnprior'].ctypes.data,
            out['logg'].ctypes.data, out['status'].ctypes.data, MS_HOST), 'ms_lnprob_parts')
        return out

    def model_spectrum_host(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        out = np.empty((theta.shape[0], self.n_data), dtype=np.float64)
        self._check(self.lib.ms_model_spectrum(self._h, theta.ctypes.data, theta.shape[0], out.ctypes.data, MS_HOST),
                    'ms_model_spectrum')
        return out

    def component_mags_host(self, theta):
        """[nw, n_contrast, nspec] band magnitudes of the individual stars (make_composite's ``mag``)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        out = np.empty((theta.shape[0], self.n_contrast, self.nspec), dtype=np.float64)
        self._check(self.lib.ms_component_mags(self._h, theta.ctypes.data, theta.shape[0], out.ctypes.data, MS_HOST),
                    'ms_component_mags')
        return out

    # -- optimiser stage
    def set_optimizer_data(self, flux_sets):
        flux_sets = np.ascontiguousarray(np.atleast_2d(flux_sets), dtype=np.float64)
        if flux_sets.shape[1] != self.n_data:
            raise ValueError("flux_sets must be [n_sets, n_data]")
        self._check(self.lib.ms_set_optimizer_data(self._h, flux_sets.ctypes.data, flux_sets.shape[0]), 'ms_set_optimizer_data')
        self._n_opt_sets = flux_sets.shape[0]

    def optimizer_chisq(self, theta, set_index, rad_sigma, av_prior=None, ref_median=None):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        nw = theta.shape[0]
        set_index = np.ascontiguousarray(set_index, dtype=np.int32)
        rad_sigma = np.ascontiguousarray(rad_sigma, dtype=np.float64)
        if theta.shape != (nw, self.ndim) or set_index.shape != (nw,) or rad_sigma.shape != (nw, self.nspec):
            raise ValueError("optimizer_chisq: shape mismatch")
        avp = None if av_prior is None else np.ascontiguousarray(av_prior, dtype=np.float64)
        med = None if ref_median is None else np.ascontiguousarray(np.broadcast_to(ref_median, (nw,)), dtype=np.float64)
        out = np.empty(nw, dtype=np.float64)
        self._check(self.lib.ms_optimizer_chisq(self._h, theta.ctypes.data, set_index.ctypes.data, rad_sigma.ctypes.data,
                                                None if avp is None else avp.ctypes.data,
                                                None if med is None else med.ctypes.data, nw, out.ctypes.data),
                    'ms_optimizer_chisq')
        return out

    # -- device-pointer calls (asynchronous on the handle's stream)
    def lnprob_device(self, theta_ptr, nw, lnprob_ptr):
        self._check(self.lib.ms_lnprob(self._h, theta_ptr, nw, lnprob_ptr, MS_DEVICE), 'ms_lnprob')

    def split_keys(self, seed, iteration, n, keys_ptr):
        self._check(self.lib.ms_split_keys(self._h, seed, iteration, n, keys_ptr), 'ms_split_keys')

    def stretch_propose(self, state_ptr, s_idx_ptr, n_s, c_idx_ptr, n_c, a, seed, iteration, split, gid0, q_ptr,
                        logfac_ptr):
        self._check(self.lib.ms_stretch_propose(self._h, state_ptr, s_idx_ptr, n_s, c_idx_ptr, n_c, self.ndim, a, seed,
                                                iteration, split, gid0, q_ptr, logfac_ptr), 'ms_stretch_propose')

    def stretch_accept(self, state_ptr, s_idx_ptr, q_ptr, lnp_q_ptr, logfac_ptr, n_s, seed, iteration, split, gid0,
                       packed_ptr):
        self._check(self.lib.ms_stretch_accept(self._h, state_ptr, s_idx_ptr, q_ptr, lnp_q_ptr, logfac_ptr, n_s,
                                               self.ndim, seed, iteration, split, gid0, packed_ptr), 'ms_stretch_accept')

    def stretch_iteration(self, state_ptr, naccepted_ptr, nw, a, seed, iteration, randomize_split=True):
        self._check(self.lib.ms_stretch_iteration(self._h, state_ptr, naccepted_ptr, nw, a, seed, iteration,
                                                  int(bool(randomize_split))), 'ms_stretch_iteration')

    def stretch_halfstep(self, state_ptr, naccepted_ptr, s_idx_ptr, n_s, c_idx_ptr, n_c, a, seed, iteration, split, gid0,
                         q_ptr, logfac_ptr, lnp_ptr, packed_ptr=None):
        """Propose + evaluate + accept for one rank's share of a half step (ms_stretch_halfstep); ``packed_ptr`` None:
        the accepted rows are written in place."""
        self._check(self.lib.ms_stretch_halfstep(self._h, state_ptr, naccepted_ptr, s_idx_ptr, n_s, c_idx_ptr, n_c, a, seed,
                                                 iteration, split, gid0, q_ptr, logfac_ptr, lnp_ptr, packed_ptr),
                    'ms_stretch_halfstep')

    def set_group_shard(self, rank, world):
        """This engine evaluates the groups rank, rank + world, ... of every ms_stretch_evaluate batch (ms_set_group_shard)."""
        self._check(self.lib.ms_set_group_shard(self._h, int(rank), int(world)), 'ms_set_group_shard')

    def group_will_run(self, nw):
        return bool(self.lib.ms_group_will_run(self._h, int(nw)))

    def stretch_evaluate(self, state_ptr, s_idx_ptr, n_s, c_idx_ptr, n_c, a, seed, iteration, split, gid0, q_ptr, logfac_ptr, lnp_ptr):
        """Proposal for all n_s walkers + evaluation of this rank's groups (ms_stretch_evaluate)."""
        self._check(self.lib.ms_stretch_evaluate(self._h, state_ptr, s_idx_ptr, n_s, c_idx_ptr, n_c, a, seed, iteration, split, gid0,
                                                 q_ptr, logfac_ptr, lnp_ptr), 'ms_stretch_evaluate')

    def stretch_accept_inplace(self, state_ptr, naccepted_ptr, s_idx_ptr, n_s, q_ptr, lnp_ptr, logfac_ptr, seed, iteration, split, gid0):
        self._check(self.lib.ms_stretch_accept_inplace(self._h, state_ptr, naccepted_ptr, s_idx_ptr, n_s, q_ptr, lnp_ptr, logfac_ptr,
                                                       seed, iteration, split, gid0), 'ms_stretch_accept_inplace')

    def stretch_scatter(self, state_ptr, naccepted_ptr, idx_ptr, packed_ptr, n):
        self._check(self.lib.ms_stretch_scatter(self._h, state_ptr, naccepted_ptr, idx_ptr, packed_ptr, n, self.ndim),
                    'ms_stretch_scatter')


def broaden_rows(rows, win_lo, win_len, taps, device=0, lib=None):
    """In-place window broadening of host rows [n_rows, row_len] on the GPU (ms_broaden_rows)."""
    lib = lib or load_library()
    if rows.dtype != np.float64 or not rows.flags.c_contiguous or rows.ndim != 2:
        raise ValueError("rows must be a C-contiguous float64 [n_rows, row_len] array")
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    rc = lib.ms_broaden_rows(int(device), rows.ctypes.data, rows.shape[0], rows.shape[1], int(win_lo),
                             int(win_len), taps.ctypes.data, int(taps.size))
    if rc != 0:
        raise NativeError(f"ms_broaden_rows failed ({rc}): {lib.ms_last_error(None).decode()}")
    return rows


def build_grid(node_wl, node_flux, wl, win_lo=0, win_len=0, taps=None, device=0, lib=None):
    """Resample ragged node spectra onto ``wl`` and broaden the window, on the GPU (ms_build_grid).
    ``node_wl`` / ``node_flux``: lists of ascending wavelength / flux arrays, one pair per grid row.
    Returns rows [n_rows, len(wl)].  Raises ValueError for an out-of-range abscissa (interp1d)."""
    lib = lib or load_library()
    lens = np.array([len(w) for w in node_wl], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    sx = np.ascontiguousarray(np.concatenate([np.asarray(w, dtype=np.float64) for w in node_wl]))
    sy = np.ascontiguousarray(np.concatenate([np.asarray(f, dtype=np.float64) for f in node_flux]))
    if sx.shape != sy.shape:
        raise ValueError("node wavelength and flux arrays differ in length")
    wl = np.ascontiguousarray(wl, dtype=np.float64)
    out = np.empty((len(lens), wl.size), dtype=np.float64)
    tp = np.ascontiguousarray(taps if taps is not None else [1.0], dtype=np.float64)
    rc = lib.ms_build_grid(int(device), sx.ctypes.data, sy.ctypes.data, off.ctypes.data, len(lens), wl.ctypes.data,
                           wl.size, int(win_lo), int(win_len), tp.ctypes.data, int(tp.size), out.ctypes.data)
    if rc != 0:
        msg = lib.ms_last_error(None).decode()
        if 'interpolation range' in msg:
            raise ValueError(msg)
        raise NativeError(f"ms_build_grid failed ({rc}): {msg}")
    return out
