/*
 * mcmcspec.h -- C ABI of libmcmcspec.so: the B200 (sm_100a) log-posterior engine for
 * unresolved binary stars.  Drop-in boundary for the per-walker `logposterior` path of
 * kendallsullivan/mcmc_spec (citations are file:line into that repository's mft6.py).
 *
 * The reference has no FFI: its boundary is the Python callable handed to the ensemble sampler,
 *     emcee.EnsembleSampler(nwalkers, ndim, logposterior, args=[fr, nspec, ndust, data, err,
 *         broadening, r, specs, ctm, ptm, tmi, tma, vs, tmin, tmax, matrix, ra, dec], kwargs=...)
 * (mft6.py:1491).  Everything `args` carries is folded once into an `ms_config` and uploaded by
 * ms_create(); ms_lnprob() then is `logposterior` for a whole batch of walkers.  The binding a
 * maintainer adds on the reference side is the ctypes stub in INTEGRATION.md (shipped as
 * mcmc_spec_b200/native.py).
 *
 * Conventions: plain pointers and sizes only; all floating point is IEEE binary64; every entry
 * point returns an int status (MS_OK == 0) and never throws; per-walker domain failures (outside
 * the prior box, off the isochrone or the grid, NaN chi-square) give lnprob = -inf, never NaN
 * (mft6.py:1227-1230, :1202-1205, :1465-1466).  Walker vectors theta are
 * [T1, T2, Av, R1, R2/R1, parallax] (mft6.py:1146), row-major [nw][ndim].
 * A handle owns one CUDA device and one stream; calls on a handle must be serialised.
 */
#ifndef MCMCSPEC_H
#define MCMCSPEC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MS_ABI_VERSION 2
#define MS_TILE 1024            /* model pixels per streaming tile; filter weights are tile padded */
#define MS_MAX_CONTRAST 4
#define MS_MAX_PHOT 12
#define MS_MAX_NDIM 8

enum {
  MS_OK = 0,
  MS_ERR_INVALID = 1,           /* bad argument / inconsistent config */
  MS_ERR_CUDA = 2,              /* CUDA runtime error; see ms_last_error */
  MS_ERR_UNSUPPORTED = 3,       /* shape outside what the kernels are built for */
  MS_ERR_NO_DEVICE = 4
};

/* Where a pointer argument lives. */
enum { MS_HOST = 0, MS_DEVICE = 1 };

/* Per-walker status bits returned by ms_lnprob_parts (diagnostics; lnprob is -inf if any set). */
enum {
  MS_W_PRIOR_BOUNDS = 1,        /* logprior bounds, mft6.py:1227-1230 */
  MS_W_OFF_ISOCHRONE = 2,       /* Teff outside the isochrone table: interp1d ValueError, mft6.py:76/:95 */
  MS_W_OFF_GRID = 4,            /* log g outside the grid nodes: get_spec index error, mft6.py:467-477 */
  MS_W_NONFINITE = 8            /* non-finite theta or NaN chi-square, mft6.py:1202 */
};

typedef struct ms_handle ms_handle;

/*
 * All pointers are HOST pointers, read during ms_create() only.
 *
 * Replaces (read-only inputs of logposterior):
 *   grid, *_nodes ........ `specs` dict of pre-broadened rows + 'wl'           mft6.py:342-385, :497-500
 *   reg_lo, n_reg ........ region cut of get_spec via make_composite           mft6.py:687, :542
 *   rebin_*, ext_* ....... interp1d(wave1, spec)(wl*1e4) after extinct()       mft6.py:1162, :1169-1170
 *   data_*, poly_basis ... `data`, `err`, norm_spec's Polynomial.fit           mft6.py:1173-1179, :193-196
 *   c_* .................. `ctm`, fr[0:2] (contrast filters, mags, errors)     mft6.py:717-741, :1182
 *   p_* .................. pyphot library filters, fr[3:5], phot_cwl           mft6.py:771-783, :1163, :1188
 *   iso_* ................ `matrix` rows of log-age 9.0, first 220             mft6.py:70-76, :90-95
 *   dust_* ............... bayestar(SkyCoord(ra, dec, 1/plx), 'samples')       mft6.py:1233-1239
 *   prior_*, tmin, tmax .. `prior`, grid extrema                               mft6.py:1241-1260, :1479-1487
 */
typedef struct ms_config {
  int32_t abi_version;          /* MS_ABI_VERSION */
  int32_t device;               /* CUDA device ordinal */

  /* model grid: rows [n_feh][n_logg][n_teff][row_len], pixels [reg_lo, reg_lo+n_reg) consumed */
  int32_t n_teff, n_logg, n_feh;
  int32_t n_reg;
  int64_t row_len;
  int64_t reg_lo;
  const double* teff_nodes;
  const double* logg_nodes;
  const double* feh_nodes;
  const double* grid;
  double feh;                   /* fixed metallicity of the fit (3-D grids only) */

  /* observed spectrum, ascending wavelength */
  int32_t n_data;
  int32_t n_tiles;              /* ceil(n_reg / MS_TILE) */
  const int32_t* rebin_hi;      /* [n_data] region-relative upper bracketing model pixel, >= 1 */
  const double* rebin_t;        /* [n_data] (x - wl[hi-1]) / (wl[hi] - wl[hi-1]) */
  const double* ext_lo;         /* [n_data] -0.4*log2(10)*k(wl[hi-1]);  E = exp2(Av*ext) */
  const double* ext_hi;         /* [n_data] same at wl[hi] */
  const double* data_flux;      /* [n_data] */
  const double* data_err;       /* [n_data] */
  const double* data_x;         /* [n_data] wavelengths mapped onto [-1, 1] */
  double poly_basis[9];         /* orthonormal quadratic basis, q_k(x) = sum_j B[3k+j] x^j */
  double data_median;           /* np.median(data_flux) */

  /* contrast filters: per-pixel quadrature weights, whole tiles, zero padded */
  int32_t n_contrast;
  int32_t n_phot;
  const int32_t* c_tile_lo;     /* [n_contrast] */
  const int32_t* c_tile_hi;
  const int64_t* c_offset;      /* element offset into c_weights */
  const double* c_weights;
  int64_t c_weights_len;
  const double* c_mag;          /* observed contrasts and errors */
  const double* c_err;

  /* photometric filters */
  const int32_t* p_tile_lo;     /* [n_phot] */
  const int32_t* p_tile_hi;
  const int64_t* p_offset;
  const double* p_weights;
  int64_t p_weights_len;
  const double* p_norm;         /* 1 / (normalising integral * zero-point flux) */
  const double* p_ext;          /* CCM89 a+b/Rv at the filter's mean wavelength */
  const double* p_mag;
  const double* p_err;

  /* isochrone (sorted by Teff), dust sightline, prior */
  int32_t n_iso;
  int32_t n_dm;
  int32_t n_dust_samples;
  int32_t nspec;
  const double* iso_teff;
  const double* iso_logg;
  const double* iso_lum;
  const double* dust_dm_edges;  /* [n_dm] */
  const double* dust_samples;   /* [n_dust_samples][n_dm] */
  double tmin, tmax;
  double prior_mu[MS_MAX_NDIM]; /* 0 -> no Gaussian prior on that parameter */
  double prior_sig[MS_MAX_NDIM];
  int32_t ext_flag;             /* `a` / `ext` keyword */
  int32_t dist_fit;
  int32_t rad_prior;
  int32_t flavour;              /* 0: mft6.py; 1: mft6_nospec.py (no spectral term, parallax <= 1/100) */
  int32_t feh_free;             /* extension for 3-D grids (binaries): [Fe/H] is a walker parameter, the trailing
                                   column of theta ([nw][ndim + 1]); outside the nodes -> -inf.  0: fixed `feh` */
  int32_t reserved0;
} ms_config;

/* sizeof(ms_config) as compiled, so a binding can verify its mirror of the struct. */
int64_t ms_config_size(void);
int32_t ms_abi_version(void);

/* Upload all tables; replaces the per-call hand-over of `args` (mft6.py:1491). */
int ms_create(const ms_config* cfg, ms_handle** out);
/*
 * Same, but the model grid is not uploaded again: the new handle reads the device rows of
 * `grid_owner` (same device, same grid shape and consumed region; cfg->grid is ignored).  For
 * several fits of one grid to different observed spectra (independent ensembles).  The owner must
 * be destroyed after every handle that shares its grid.
 */
int ms_create_shared(const ms_config* cfg, const ms_handle* grid_owner, ms_handle** out);
int ms_destroy(ms_handle* h);
/* Message of the last failing call on this handle (or of the last failing ms_create if h==NULL). */
const char* ms_last_error(const ms_handle* h);
/* The handle's CUDA stream (cudaStream_t) and device ordinal, for callers that enqueue around it. */
void* ms_stream(const ms_handle* h);
int ms_device(const ms_handle* h);

/*
 * logposterior for nw walkers (mft6.py:1459-1470).  theta [nw][ndim], lnprob [nw]; `where` says
 * whether both are MS_HOST (copied through pinned staging, call returns after the result landed)
 * or MS_DEVICE (enqueued on the handle's stream, asynchronous).
 */
int ms_lnprob(ms_handle* h, const double* theta, int64_t nw, double* lnprob, int where);

/*
 * Same evaluation with the intermediate quantities of loglikelihood exposed (any pointer may be
 * NULL): chi2_spec [nw] = sum(chisq)/N (mft6.py:1179), contrast [nw][n_contrast] (mft6.py:741),
 * phot [nw][n_phot] extincted magnitudes (mft6.py:1163), lnprior [nw] (mft6.py:1462),
 * logg [nw][nspec] (mft6.py:1149), status [nw] MS_W_* bits.
 */
int ms_lnprob_parts(ms_handle* h, const double* theta, int64_t nw, double* lnprob, double* chi2_spec,
                    double* contrast, double* phot, double* lnprior, double* logg, int32_t* status,
                    int where);

/*
 * Optimiser stage (optimize_fit / fit_spec, mft6.py:856-1137, :1686-1765).
 * ms_set_optimizer_data: the observed spectra after fit_spec's one-off norm_spec against each
 *   starting guess (mft6.py:889), flux_sets [n_sets][n_data], host pointer.
 * ms_optimizer_chisq: fit_spec's `test_cs` for nw proposals (mft6.py:995-1046) -- 3 x the mean spectral
 *   chi-square of the median-scaled model against flux_sets[set_index[w]], times (n_contrast + n_phot),
 *   plus the contrast and photometric chi-squares, plus opt_prior terms in chi-square units: Av against
 *   the Bayestar moments at the proposal's distance (av_prior[w] replaces theta's Av in that term when
 *   av_prior != NULL, as the initial evaluation at mft6.py:927 needs), the parallax prior when
 *   dist_fit, and the radius prior with sigmas rad_sigma [nw][nspec] (mft6.py:1038).  The model is
 *   scaled to the median of its flux set (mft6.py:1011) unless ref_median [nw] is given (the initial
 *   evaluation scales to the median of the *un-renormalised* spectrum, mft6.py:888).  The box of
 *   mft6.py:984 is the caller's business.  +inf where the forward model is undefined.  Host pointers.
 */
int ms_set_optimizer_data(ms_handle* h, const double* flux_sets, int32_t n_sets);
int ms_optimizer_chisq(ms_handle* h, const double* theta, const int32_t* set_index, const double* rad_sigma,
                       const double* av_prior, const double* ref_median, int64_t nw, double* chisq);

/* Median-scaled model on the observed pixels, out [nw][n_data] (init_cspec after mft6.py:1173). */
int ms_model_spectrum(ms_handle* h, const double* theta, int64_t nw, double* out, int where);

/*
 * Band magnitudes of each component, out [nw][n_contrast][nspec] = -2.5 log10(trapz(F_k * T, lambda))
 * (mft6.py:729-733): what the posterior-predictive sweep of plot_results derives per sample through
 * make_composite(plot=True) (mft6.py:785-828, :2486-2509) -- e.g. the Kepler magnitudes of primary
 * and secondary, hence Delta Kp and the planet-radius correction factors -- when the band of interest
 * is passed as one of the contrast filters.  NaN for walkers outside the prior box.
 */
int ms_component_mags(ms_handle* h, const double* theta, int64_t nw, double* out, int where);

/*
 * Instrumental broadening of grid rows in place (host rows [n_rows][row_len]): the operator of
 * spec_interpolator/broaden, mft6.py:366-383 and :124-130 -- pixels [win_lo, win_lo+win_len) of
 * every row are convolved with the sum-normalised `taps` (odd n_taps, zero padded 'same'), then
 * the first 5 and last 10 window samples are patched.  Runs on `device`.
 */
int ms_broaden_rows(int device, double* rows, int64_t n_rows, int64_t row_len, int64_t win_lo,
                    int64_t win_len, const double* taps, int32_t n_taps);

/*
 * Model-grid assembly on the GPU: the numeric part of spec_interpolator (btsettl branch), mft6.py:343-383.
 * Node r's native spectrum is (src_x, src_y)[src_off[r] .. src_off[r+1]) with ascending src_x; every node
 * is linearly interpolated onto the common grid x_dst[row_len] (interp1d semantics, bit-for-bit np.interp;
 * an abscissa outside a node's range is an error, as interp1d's bounds_error raises), then -- when
 * win_len > 0 -- pixels [win_lo, win_lo+win_len) are broadened as by ms_broaden_rows and spliced between
 * the untouched blue and red parts.  rows_out: host [n_rows][row_len].
 */
int ms_build_grid(int device, const double* src_x, const double* src_y, const int64_t* src_off, int64_t n_rows,
                  const double* x_dst, int64_t row_len, int64_t win_lo, int64_t win_len, const double* taps,
                  int32_t n_taps, double* rows_out);

/*
 * Stretch-move half step on a device-resident ensemble (emcee RedBlueMove/StretchMove, a = 2; the
 * sampler the reference drives at mft6.py:1491-1510).  `state` [nw][ndim+1] holds positions and
 * log-probabilities; s_idx [n_s] are the state rows this caller updates (its shard of the active
 * half), c_idx [n_c] the complementary half.  All pointers are device pointers.
 *   ms_split_keys:      keys[i] (63-bit) whose argsort is the iteration's random red/blue split.
 *   ms_stretch_propose: z = ((a-1)u+1)^2/a, j ~ randint(n_c),
 *                       q[s] = c_j - (c_j - x_s) z, logfac[s] = (ndim-1) ln z.
 *   ms_stretch_accept:  accept where ln u' < logfac + lnp_q - lnp_x; packed[s] = new position |
 *                       new log-prob | accepted (0/1)   [n_s][ndim+2] -- the all-gather payload.
 *   ms_stretch_scatter: state[idx[k]] = packed[k][0..ndim]; naccepted[idx[k]] += packed[k][ndim+1].
 * Random numbers are counter based: stream (seed, iteration, split, gid0 + s), so a sharded ensemble
 * draws the same numbers as an unsharded one.
 */
int ms_split_keys(ms_handle* h, uint64_t seed, uint64_t iteration, int64_t n, int64_t* keys);
int ms_stretch_propose(ms_handle* h, const double* state, const int64_t* s_idx, int64_t n_s, const int64_t* c_idx,
                       int64_t n_c, int32_t ndim, double a, uint64_t seed, uint64_t iteration, int32_t split,
                       int64_t gid0, double* q, double* logfac);
int ms_stretch_accept(ms_handle* h, const double* state, const int64_t* s_idx, const double* q, const double* lnp_q,
                      const double* logfac, int64_t n_s, int32_t ndim, uint64_t seed, uint64_t iteration,
                      int32_t split, int64_t gid0, double* packed);
int ms_stretch_scatter(ms_handle* h, double* state, int32_t* naccepted, const int64_t* idx, const double* packed,
                       int64_t n, int32_t ndim);

/*
 * One half step of the stretch move for ensembles of any size, sharded or not, in one call: the proposal for the
 * n_s walkers state[s_idx[k]] (partners drawn from state[c_idx[0..n_c)]; random streams keyed by (seed, iteration,
 * split, gid0 + k), as ms_stretch_propose) is folded into the preparation kernel, the proposals are evaluated, and the
 * acceptance test is applied -- in place when `packed` is NULL (this rank owns the whole active half), otherwise into
 * packed[n_s][ncol + 2] = new position | new log-prob | accepted, the rows the ranks exchange (one all-gather) and
 * ms_stretch_scatter applies.  q [n_s][ncol], logfac [n_s], lnp_q [n_s] are device scratch owned by the caller.
 * All pointers are device pointers; work is enqueued on the handle's stream.
 */
int ms_stretch_halfstep(ms_handle* h, double* state, int32_t* naccepted, const int64_t* s_idx, int64_t n_s,
                        const int64_t* c_idx, int64_t n_c, double a, uint64_t seed, uint64_t iteration, int32_t split,
                        int64_t gid0, double* q, double* logfac, double* lnp_q, double* packed);

/*
 * The same half step in two calls, for N processes that each hold the whole ensemble and share the evaluation by
 * groups (ms_set_group_shard): ms_stretch_evaluate proposes for all n_s walkers of the half and evaluates this
 * rank's groups (lnp_q is 0 for the other ranks' walkers: the caller sums lnp_q over the ranks, one all-reduce);
 * ms_stretch_accept_inplace then applies the acceptance test to all n_s walkers in place on every rank -- the ranks'
 * ensembles stay identical, no positions are exchanged.  MS_ERR_UNSUPPORTED when a shard is set and the batch does
 * not take the group kernel (ms_group_will_run).
 */
int ms_stretch_evaluate(ms_handle* h, const double* state, const int64_t* s_idx, int64_t n_s, const int64_t* c_idx, int64_t n_c,
                        double a, uint64_t seed, uint64_t iteration, int32_t split, int64_t gid0, double* q, double* logfac,
                        double* lnp_q);
int ms_stretch_accept_inplace(ms_handle* h, double* state, int32_t* naccepted, const int64_t* s_idx, int64_t n_s, const double* q,
                              const double* lnp_q, const double* logfac, uint64_t seed, uint64_t iteration, int32_t split,
                              int64_t gid0);

/*
 * One whole ensemble iteration (both half steps) for a single-rank ensemble of at most 2048 walkers, enqueued
 * from one call: the iteration's red/blue split (hashed keys sorted in one CTA; randomize_split = 0 gives
 * evens/odds), then per half: propose, evaluate (prep + streaming kernel), accept, scatter -- in place on
 * `state` [nw][ndim+1] (device).  Same random streams and arithmetic as the per-kernel entry points above,
 * i.e. the same chain; what it removes is the host round trip between launches, which dominates ensembles as
 * small as the reference's 50 walkers (emcee loop at mft6.py:1491-1510).
 */
int ms_stretch_iteration(ms_handle* h, double* state, int32_t* naccepted, int64_t nw, double a, uint64_t seed,
                         uint64_t iteration, int32_t randomize_split);

/*
 * Enqueue on the caller's stream (a cudaStream_t, taken literally: 0 is the CUDA legacy default
 * stream) instead of the handle's own, e.g. the stream a collective library runs on.  The caller
 * keeps ownership.  ms_use_own_stream() goes back to the stream the handle created.
 */
int ms_set_stream(ms_handle* h, void* stream);
int ms_use_own_stream(ms_handle* h);

/*
 * CTAs per walker in the streaming kernel.  One CTA per walker cannot fill the GPU for ensembles of a
 * few dozen walkers (the reference runs 50): a thread-block cluster of 2/4/8 CTAs then shares a walker
 * (partial sums and the rebinned model meet in the leader CTA through distributed shared memory).
 * 0 = automatic (clusters only while nw * ctas <= the number of SMs), 1 = never, 2/4/8 = forced.  Values
 * agree with the one-CTA path to rounding (the partial sums are combined in a different order).
 */
int ms_set_cluster(ms_handle* h, int ctas_per_walker);

/*
 * Walkers per CTA for large batches.  enabled = 1: batches of at least 8 walkers per SM are launched in
 * grid-cell order with two cell-mates per CTA; the two thread groups request the same bracketing rows,
 * filter weights and per-pixel records back to back, so every second request is served by the SM's L1
 * instead of L2 (L2->SM traffic halves).  enabled = 0: always one walker per CTA.  enabled = -1
 * (default): on for stream-dominated shapes (model pixels >= 16 x observed pixels), where it pays.
 * A walker's value does not depend on the setting or on its partner (bit-identical).
 */
int ms_set_pairing(ms_handle* h, int enabled);

/*
 * Bracket-only evaluation (default off; the streaming kernels stay the reference path).  Interpolation in the grid,
 * dilution, np.trapz and the linear rebinning of mft6.py:1169 are linear in the grid rows, and the interpolation
 * weights do not depend on wavelength.  With enabled = 1 an evaluation therefore reads (i) only the two model
 * pixels that bracket each observed pixel, from a compact [row][n_data][2] copy of the grid built at ms_create
 * (extinction is still applied at those model pixels before the rebinning, mft6.py:1162 -> :1169), and (ii) the
 * band fluxes of its 2^D rows per star from a [row][band] table, instead of streaming n_star * 2^D whole rows:
 * n_star * 2^D * (2 n_data + n_bands) * 8 bytes per evaluation.  The model on the observed pixels is
 * bit-identical to the streaming path; chi-square, contrasts and magnitudes agree to rounding (summation order).  Posterior flavours
 * only (the optimiser stage keeps streaming).
 */
int ms_set_fast(ms_handle* h, int enabled);

/*
 * Launch order of batches of >= 1024 walkers.  1: grouped by grid cell (counting sort on the device), so that
 * CTAs resident together stream the same rows and a grid larger than L2 is fetched from HBM about once per
 * cell group instead of once per walker; 0: the caller's order; -1 (default): grouped when the grid exceeds
 * the device's L2.  Values do not depend on the order.
 */
int ms_set_cell_order(ms_handle* h, int mode);

/*
 * Group evaluation of large posterior batches (binaries on a (Teff, log g) grid).  Walkers are sorted by grid
 * cell on the device and handed out in groups of up to 8 cell-mates; one persistent CTA streams the group's
 * 2 * 2^D bracketing rows ONCE through a 4-stage shared-memory ring filled by bulk tensor copies (TMA, mbarrier
 * completion), integrates the staged rows through the passbands, gathers the bracketing model pixels of the
 * observed pixels, and one warp per walker does that walker's extinction, rebinning, quadratic renormalisation,
 * exact median and chi-square.  The reference's per-walker contract is unchanged (get_spec / make_composite /
 * loglikelihood of mft6.py are linear in the grid rows up to the extinction, which stays per walker and per
 * model pixel); values agree with the one-CTA-per-walker kernel to rounding (summation order of the band
 * integrals), and a walker's value does not depend on which other walkers share its group.
 * mode: -1 (default) automatic -- batches of at least `min_walkers` (default 1024; <= 0 keeps the current
 * threshold), as long as the batch does not spread so thinly over the grid cells that a group holds fewer than ~6
 * walkers (a group costs the same with 3 busy lanes as with 32; the count of the previous batch of the same size
 * decides, with hysteresis; the first batch of a size waits once for its own count); 0 never; 1 whenever the handle
 * can take it.
 * ms_group_supported: 1 when this handle can take it (nspec = 2, a (Teff, log g) or (Teff, log g, [Fe/H]) grid with at
 * least two nodes per axis, more than one grid cell, at most 65 535 observed pixels).
 * ms_set_group_shard: this handle evaluates the groups rank, rank + world, ... of every batch and leaves the other
 * walkers' lnprob at 0 -- N processes that hold the same ensemble each evaluate 1/N of the groups of the WHOLE half
 * (whole groups per GPU instead of 1/N of every cell's walkers) and sum the results (one all-reduce).  Only
 * ms_stretch_evaluate honours it.  ms_group_will_run: whether a posterior batch of nw walkers takes the group kernel.
 */
int ms_set_group(ms_handle* h, int mode, int64_t min_walkers);
int ms_group_supported(const ms_handle* h);
int ms_set_group_shard(ms_handle* h, int rank, int world);
int ms_group_will_run(const ms_handle* h, int64_t nw);

/*
 * Evaluations that were cut short before any grid row was read (outside the prior box, off the isochrone or the
 * grid, non-finite input: -inf per mft6.py:1227-1230, :1465-1466) since the handle was created or the counter was
 * last reset.  Synchronises the handle's stream.  bench.py reports it so that a throughput figure cannot hide
 * skipped work.
 */
int ms_short_circuited(ms_handle* h, int64_t* count, int reset);

/* Number of kernel launches this handle has issued (bench.py's gpu_launches). */
int64_t ms_launch_count(const ms_handle* h);

/*
 * Device-side timing of the streaming kernel: after ms_set_timing(h, 1) every launch of
 * lnprob_stream_kernel is bracketed by CUDA events on its stream.  ms_kernel_time() synchronises
 * those events, returns their summed duration in milliseconds and the launch / walker counts they
 * cover, and resets the accumulation.
 */
int ms_set_timing(ms_handle* h, int enabled);
int ms_kernel_time(ms_handle* h, double* total_ms, int64_t* n_launches, int64_t* n_walkers);

#ifdef __cplusplus
}
#endif
#endif /* MCMCSPEC_H */
