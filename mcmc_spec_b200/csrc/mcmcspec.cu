// mcmcspec.cu -- sm_100a kernels and C ABI of libmcmcspec.so (see include/mcmcspec.h).
//
// One evaluation of mft6.py's logposterior (mft6.py:1459-1470) is split into
//   walker_prep_kernel   : one thread per walker -- prior (mft6.py:1207-1327), isochrone log g and
//                          radius (mft6.py:66-98), grid bracket + bilinear/trilinear coefficients
//                          folded with the (R/d)^2 dilution (mft6.py:439-519, :689-700).
//   lnprob_stream_kernel : one CTA per walker -- streams the 2*2^D bracketing grid rows once
//                          (vectorised, coalesced 16-byte loads), forms both component spectra in
//                          registers, accumulates the contrast and photometric band integrals
//                          (mft6.py:717-783), extincts and rebins onto the observed pixels through
//                          a shared-memory tile (mft6.py:1162, :1169-1170), then finishes in shared
//                          memory: exact median (mft6.py:1173), quadratic renormalisation
//                          (mft6.py:193-196), chi-square and the final sum (mft6.py:1178-1205).
// No intermediate spectrum ever goes to HBM.  All arithmetic is binary64.
//
// Setup-time operator: broaden_rows_kernel = spec_interpolator's window broadening
// (mft6.py:124-130, :366-383) as a shared-memory banded stencil.
// Sampler support: split_keys / stretch_propose / stretch_accept / stretch_scatter kernels (emcee
// stretch move, a = 2) working in place on the replicated ensemble state.

#include "mcmcspec.h"

#include <cooperative_groups.h>
#include <cuda.h>
#if !defined(MS_TU) || MS_TU == 0 || MS_TU == 1
#include <cub/device/device_radix_sort.cuh>
#endif
#include <cuda_runtime.h>
#include <math_constants.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

namespace cg = cooperative_groups;

// Translation units.  The streaming kernel has ~60 instantiations; `python -m mcmc_spec_b200.build` compiles this file
// three times in parallel with -DMS_TU=1/2/3 and links the objects (1: the C ABI, all small kernels and the
// (Teff, log g) grids with <= 2 + 6 bands; 2: (Teff, log g) grids with the wide band set; 3: grids with [Fe/H]).
// MS_TU = 0 (default, e.g. a plain `nvcc -c mcmcspec.cu`) keeps everything in one unit.
#ifndef MS_TU
#define MS_TU 0
#endif
#define MS_HAS_MAIN (MS_TU == 0 || MS_TU == 1)
#define MS_HAS_NR4_WIDE (MS_TU == 0 || MS_TU == 2)
#define MS_HAS_NR8 (MS_TU == 0 || MS_TU == 3)
#define MS_HAS_GROUP (MS_TU == 0 || MS_TU == 4)      // lnprob_group_kernel (group_kernel.cuh)
#ifndef MS_GROUP_FILL_LO
#define MS_GROUP_FILL_LO 6.0                         // walkers per group below which the per-walker kernel takes over ...
#define MS_GROUP_FILL_HI 9.0                         // ... and above which the group kernel comes back
#endif
#ifndef MS_GROUP_VW
#define MS_GROUP_VW 16                               // partial sums per walker of the group kernel = walker warps with one CTA per SM
#endif
#define MS_GROUP_NWW_WIDE MS_GROUP_VW
#define MS_GROUP_NWW (MS_GROUP_VW / 2)               // walker warps with two CTAs per SM

namespace msd {   // types shared by the translation units (external linkage)

constexpr int MAXWARP = 16;              // the streaming kernel runs with 256 or 512 threads per CTA
constexpr int TILE = MS_TILE;           // model pixels per tile (4 per thread)
constexpr int MAX_STAR = 3;              // binaries and triples (mft6.py nspec = 2 / 3)
constexpr int MAX_CORNER = 8;
constexpr int BAND_COLS = MS_MAX_CONTRAST + MS_MAX_PHOT;

constexpr double RSUN_CM = 6.957e10;    // mft6.py:81, :691
constexpr double PC_CM = 3.086e18;      // mft6.py:691
constexpr double SIGMA_SB = 5.670374e-5;  // mft6.py:79
constexpr double LSUN = 3.839e33;       // mft6.py:80

constexpr int CHUNK = 128;               // model pixels a warp owns per step: 32 lanes x 4 pixels (one 256-bit load per row)
constexpr int CHUNKS_PER_TILE = TILE / CHUNK;
constexpr int SEG = CHUNK + CHUNK / 16 + 4;   // shared-memory doubles per warp: previous pixel + CHUNK composites, skewed

// Slot of local pixel q (q = 0: the pixel before the chunk, q = j + 1: chunk pixel j).  One slot is
// skipped every 16 so that the stride-4 reads of the rebinning (observed pixels are ~4 model pixels
// apart) fall into distinct banks.
__device__ __forceinline__ int seg_slot(int q) { return q + (q >> 4); }

static_assert(TILE % CHUNK == 0, "filter masks are per tile, work is per chunk");

// ------------------------------------------------------------------------------------------
// Device-side view of the tables (passed by value as a kernel parameter).
// ------------------------------------------------------------------------------------------
struct DevTables {
  // grid
  const double* grid;          // [n_rows][stride], region cut, zero padded to whole tiles
  long long stride;
  const double* teff_nodes; const double* logg_nodes; const double* feh_nodes;
  int n_teff, n_logg, n_feh;
  int n_reg, n_tiles;
  double feh;
  // observed spectrum
  int n_data;
  // 32-byte records, one 256-bit load per observed pixel and lane (fully coalesced):
  const double* rebin_rec;      // [N][4] {t, ext_lo, (ext_hi - ext_lo) ln 2, upper bracketing pixel (as a double)}
  const double* data_rec;       // [N][4] {x, flux, 1/err^2, 0}
  const double* data_inv_e2;    // [N]
  const int4* chunk_meta;       // [n_tiles*CHUNKS_PER_TILE] {contrast filter mask, photometric filter mask, first observed
                                // pixel owned by the chunk (bit 31: its lower bracket is the last pixel of the previous
                                // chunk), one past the last observed pixel}: one 16-byte load per chunk
  double dext_max;              // max |(ext_hi - ext_lo) ln 2|
  const double* opt_median;     // [n_sets] np.median of each renormalised optimiser flux set
  double basis[9];
  double data_median;
  // filters: pointer such that weight of region pixel p is wbase[p]
  int n_contrast, n_phot;
  const double* c_wbase[MS_MAX_CONTRAST];
  const double* p_wbase[MS_MAX_PHOT];
  double c_mag[MS_MAX_CONTRAST], c_err[MS_MAX_CONTRAST];
  double p_norm[MS_MAX_PHOT], p_ext[MS_MAX_PHOT], p_mag[MS_MAX_PHOT], p_err[MS_MAX_PHOT];
  // isochrone, dust, prior
  const double* iso_teff; const double* iso_logg; const double* iso_lum; int n_iso;
  const double* dust_dm; const double* dust_samples; int n_dm, n_dust_samples;
  double tmin, tmax;
  double prior_mu[MS_MAX_NDIM], prior_sig[MS_MAX_NDIM];
  int nspec, ndim, ext_flag, dist_fit, rad_prior;
  int nospec;                   // mft6_nospec.py flavour (no spectral term), also for the optimiser's chi-square
  int feh_free, ncol;           // per-walker [Fe/H] as an extra trailing column of theta; ncol = ndim + feh_free
  int n_cells;                 // number of cell keys (bins of the launch-order counting sort)
  // bracket-only mode (ms_set_fast): what an evaluation needs of the grid when linearity is used
  const double2* bracket;      // [n_rows][n_data] {row[hi-1], row[hi]}: the two model pixels bracketing each observed pixel
  const double* bandsum;       // [n_rows][BAND_COLS] sum_p w_f[p] row[p]: contrast f -> column f, photometric f -> MS_MAX_CONTRAST + f
};

// What walker_prep_kernel hands to the streaming kernel.
struct __align__(16) WalkerPlan {
  double coef[MAX_STAR][MAX_CORNER];   // interpolation weight x (R/d)^2, per grid row
  int row[MAX_STAR][MAX_CORNER];       // grid row index
  double av;
  double lp;                           // log prior (finite when status == 0)
  double logg[MAX_STAR];
  int status;                          // MS_W_* bits
  int cell;                            // grid-cell key of the two stars (walkers of one cell read the same rows)
};

// What is evaluated for a walker.
//   MODE_POSTERIOR : logposterior of mft6.py (:1459-1470)
//   MODE_NOSPEC    : logposterior of mft6_nospec.py -- the spectral term is dropped (its :1163-1196) and
//                    the parallax box is [1/3000, 1/100]
//   MODE_OPTIMIZER : fit_spec's test chi-square (mft6.py:995-1046): 3 x mean spectral chi-square against
//                    a fixed, once-renormalised observed spectrum, chi-square priors, no prior box
enum { MODE_POSTERIOR = 0, MODE_NOSPEC = 1, MODE_OPTIMIZER = 2 };

struct OptIn {
  const double* flux_sets;      // [n_sets][n_data] observed flux after fit_spec's one-off norm_spec (mft6.py:889)
  const int* set_index;         // [nw] which set a walker is compared with
  const double* rad_sigma;      // [nw][nspec] sigma of the radius prior (mft6.py:1038: si[2])
  const double* av_prior;       // [nw] Av entering the prior term, or NULL (= theta's Av)
  const double* ref_median;     // [nw] median the model is scaled to, or NULL (= median of the walker's set)
};

// Stretch-move proposal folded into the prep kernel (ms_stretch_iteration): thread w draws z and its partner,
// writes the proposal q[w] = theta[w] and the log of the Jacobian factor, then goes on preparing that walker.
struct ProposeIn {
  const double* state;          // [nw_total][ncol + 1], or NULL: theta is given
  const long long* s_idx; const long long* c_idx; long long n_c;
  double a; unsigned long long seed, iteration; int split; long long gid0;
  double* q; double* logfac;
};

struct PartsOut {
  double* chi2_spec; double* contrast; double* phot; double* lnprior; double* logg; int* status;
  double* model;                        // [nw][n_data] median-scaled model
  double* comp_mag;                     // [nw][n_contrast][nspec] -2.5 log10 of each component's band flux (mft6.py:733)
};

}  // namespace msd
using namespace msd;

// launchers of the streaming-kernel families that live in other translation units (see MS_TU)
struct ms_handle;
int ms_tu_launch_nr4_wide(ms_handle* h, long long nw, double* d_lnp, const msd::PartsOut& parts, int mode, const msd::OptIn& opt);
int ms_tu_launch_nr8(ms_handle* h, int wide, long long nw, double* d_lnp, const msd::PartsOut& parts, int mode, const msd::OptIn& opt);
int ms_tu_launch_group(ms_handle* h, long long nw, double* d_lnp, const msd::PartsOut& parts);

namespace {

// ------------------------------------------------------------------------------------------
// Small device helpers
// ------------------------------------------------------------------------------------------

// 1/x to ~1 ulp for finite non-zero x: one third-order step on the hardware seed (rcp.approx.ftz.f64 works from the
// high word only, relative error e ~ 2^-20: r0 (1 + e + e^2) leaves e^3 ~ 2^-60); 0 -> inf, NaN -> NaN.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}

// Blackwell 256-bit global loads (LDG.E.256): four consecutive pixels per lane and instruction.
struct __align__(32) dbl4 { double x, y, z, w; };
__device__ __forceinline__ dbl4 ldg_stream4(const double* p) {
  dbl4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ dbl4 ldg_cached4(const double* p) {
  dbl4 v;
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

#ifndef MS_NT_SMALL
#define MS_NT_SMALL 256                  // CTA width when two CTAs share an SM
#endif
#ifndef MS_NT_LARGE
#define MS_NT_LARGE 512                  // CTA width when one CTA owns an SM
#endif
#ifndef MS_PAIRING_DEFAULT
#define MS_PAIRING_DEFAULT -1
#endif
#ifndef MS_FORCE_LARGE
#define MS_FORCE_LARGE 0
#endif
#ifndef MS_CHI_UNROLL
#define MS_CHI_UNROLL 4
#endif

// Barrier over the NT threads of one walker's thread group.  A CTA holds one group (bar 0 == __syncthreads)
// or two groups working on two walkers (bars 1 and 2), see lnprob_stream_kernel.
template <int NT>
__device__ __forceinline__ void gsync(int bar) { asm volatile("bar.sync %0, %1;" :: "r"(bar), "n"(NT) : "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of K per-thread values over the CTA, in a fixed order; result valid in every thread.
template <int K, int NWARP>
__device__ __forceinline__ void block_sum(double (&v)[K], double* scratch /* [NWARP*K] */, int bar) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x % (NWARP * 32)) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) scratch[warp * K + k] = s;
  }
  gsync<NWARP * 32>(bar);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) s += scratch[w * K + k];
    v[k] = s;
  }
  gsync<NWARP * 32>(bar);
}

// np.interp / interp1d(kind='linear', bounds_error=True) on an ascending table.
// Two ordinates over the same abscissae (isochrone log g and luminosity at one Teff): one bisection.
__device__ bool interp_table2(const double* __restrict__ x, const double* __restrict__ y1, const double* __restrict__ y2, int n,
                              double v, double* out1, double* out2) {
  if (!(v >= x[0] && v <= x[n - 1])) return false;
  if (v == x[n - 1]) { *out1 = y1[n - 1]; *out2 = y2[n - 1]; return true; }
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (x[mid] <= v) lo = mid; else hi = mid;
  }
  if (x[lo] == v) { *out1 = y1[lo]; *out2 = y2[lo]; return true; }
  const double dx = x[lo + 1] - x[lo], dv = v - x[lo];
  *out1 = ((y1[lo + 1] - y1[lo]) / dx) * dv + y1[lo];
  *out2 = ((y2[lo + 1] - y2[lo]) / dx) * dv + y2[lo];
  return true;
}

// find_nearest + partner on the far side (mft6.py:100-113, :439-453).  w2 is the weight of node i2.
__device__ bool bracket_nodes(const double* __restrict__ nodes, int n, double v, int* i1, int* i2, double* w2) {
  int best = 0;
  double bd = fabs(nodes[0] - v);
  for (int i = 1; i < n; ++i) {
    const double d = fabs(nodes[i] - v);
    if (d < bd) { bd = d; best = i; }
  }
  *i1 = best;
  if (nodes[best] == v) { *i2 = best; *w2 = 0.0; return true; }
  const int other = (nodes[best] > v) ? best - 1 : best + 1;
  if (other < 0 || other >= n) return false;     // reference: wrap-around / IndexError
  *i2 = other;
  *w2 = (v - nodes[best]) / (nodes[other] - nodes[best]);
  return true;
}

__device__ double stefan_boltzmann_radius(double lum, double teff) {
  // mft6.py:83
  return sqrt(lum * LSUN / (4.0 * CUDART_PI * SIGMA_SB * (teff * teff * teff * teff))) / RSUN_CM;
}

// bayestar(SkyCoord(distance), mode='samples') * 3.1 * 0.884 -> mean, std   (mft6.py:1233-1238)
__device__ void dust_prior_moments(const double* __restrict__ dust_dm, const double* __restrict__ dust_samples, int n, int ns,
                                   double plx, double* mu, double* sig) {
  const double d_kpc = (1.0 / plx) / 1.0e3;
  const double dm = 5.0 * (log10(d_kpc) + 2.0);
  int c = 0;                                           // np.searchsorted(edges, dm) (side='left')
  { int lo = 0, hi = n; while (lo < hi) { int mid = (lo + hi) >> 1; if (dust_dm[mid] < dm) lo = mid + 1; else hi = mid; } c = lo; }
  double s[16];
  double a = 0.0;
  if (c == 0) a = pow(10.0, 0.2 * (dm - dust_dm[0]));
  else if (c < n) a = (dust_dm[c] - dm) / (dust_dm[c] - dust_dm[c - 1]);
  double sum = 0.0;
  for (int k = 0; k < ns; ++k) {
    double e;
    if (c == 0) e = a * dust_samples[k * n];
    else if (c == n) e = dust_samples[k * n + n - 1];
    else e = (1.0 - a) * dust_samples[k * n + c] + a * dust_samples[k * n + c - 1];
    e = e * 3.1 * 0.884;
    s[k] = e;
    sum += e;
  }
  const double mean = sum / ns;
  double var = 0.0;
  for (int k = 0; k < ns; ++k) { const double d = s[k] - mean; var += d * d; }
  *mu = mean;
  *sig = sqrt(var / ns);
}

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x);
__host__ __device__ __forceinline__ double counter_uniform(unsigned long long seed, unsigned long long iteration,
                                                           int split, int purpose, unsigned long long gid);

// ------------------------------------------------------------------------------------------
// walker_prep_kernel
// ------------------------------------------------------------------------------------------
__global__ void walker_prep_kernel(const DevTables T, const double* __restrict__ theta, long long nw,
                                   WalkerPlan* __restrict__ plans_out, unsigned* __restrict__ cell_hist,
                                   const int mode, const OptIn opt, const ProposeIn prop,
                                   unsigned long long* __restrict__ short_circuited) {
  // The isochrone, node and dust tables (a few KB) are staged in shared memory once per CTA: the searches below
  // are chains of dependent loads (two bisections of the isochrone per star, the dust slice, three node scans),
  // ~60 L2 round trips per walker when read from global memory -- which made this kernel 28 us, more than a
  // third of a 50-walker sampler iteration.
  extern __shared__ __align__(16) unsigned char prep_smem[];
  double* s_iso_teff = reinterpret_cast<double*>(prep_smem);
  double* s_iso_logg = s_iso_teff + T.n_iso;
  double* s_iso_lum = s_iso_logg + T.n_iso;
  double* s_teff = s_iso_lum + T.n_iso;
  double* s_logg = s_teff + T.n_teff;
  double* s_feh = s_logg + T.n_logg;
  double* s_dm = s_feh + T.n_feh;
  double* s_dust = s_dm + T.n_dm;
  for (int k = threadIdx.x; k < T.n_iso; k += blockDim.x) {
    s_iso_teff[k] = T.iso_teff[k]; s_iso_logg[k] = T.iso_logg[k]; s_iso_lum[k] = T.iso_lum[k];
  }
  for (int k = threadIdx.x; k < T.n_teff; k += blockDim.x) s_teff[k] = T.teff_nodes[k];
  for (int k = threadIdx.x; k < T.n_logg; k += blockDim.x) s_logg[k] = T.logg_nodes[k];
  for (int k = threadIdx.x; k < T.n_feh; k += blockDim.x) s_feh[k] = T.feh_nodes[k];
  for (int k = threadIdx.x; k < T.n_dm; k += blockDim.x) s_dm[k] = T.dust_dm[k];
  for (int k = threadIdx.x; k < T.n_dm * T.n_dust_samples; k += blockDim.x) s_dust[k] = T.dust_samples[k];
  __syncthreads();
  const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (w >= nw) return;
  // every exit stores the plan and counts the walker in its cell bin
  auto store = [&](const WalkerPlan& p) {
    plans_out[w] = p;
    if (cell_hist) {
      // warp-aggregated: the walkers of a chain near its mode share a handful of cells
      const unsigned peers = __match_any_sync(__activemask(), p.cell);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(cell_hist + p.cell, (unsigned)__popc(peers));
    }
    if (p.status != 0) atomicAdd(short_circuited, 1ull);      // never reaches the grid: -inf without streaming
  };
  const int ndim = T.ndim, ns = T.nspec;
  double p[MS_MAX_NDIM];
  bool finite = true;
  double fehv = T.feh;
  if (prop.state) {
    // emcee StretchMove (a = 2), same arithmetic and random streams as stretch_propose_kernel
    const unsigned long long gid = (unsigned long long)(prop.gid0 + w);
    const double u = counter_uniform(prop.seed, prop.iteration, prop.split, 0, gid);
    const double root = (prop.a - 1.0) * u + 1.0;
    const double z = root * root / prop.a;
    long long j = (long long)(counter_uniform(prop.seed, prop.iteration, prop.split, 1, gid) * (double)prop.n_c);
    if (j >= prop.n_c) j = prop.n_c - 1;
    const double* x = prop.state + prop.s_idx[w] * (T.ncol + 1);
    const double* c = prop.state + prop.c_idx[j] * (T.ncol + 1);
    for (int k = 0; k < T.ncol; ++k) {
      const double qk = c[k] - (c[k] - x[k]) * z;
      prop.q[w * T.ncol + k] = qk;
      if (k < ndim) p[k] = qk; else fehv = qk;
    }
    prop.logfac[w] = (double)(T.ncol - 1) * log(z);
    for (int k = 0; k < ndim; ++k) finite = finite && isfinite(p[k]);
  } else {
    for (int k = 0; k < ndim; ++k) { p[k] = theta[w * T.ncol + k]; finite = finite && isfinite(p[k]); }
    // 3-D grids: the metallicity is fixed for the fit, or (an extension, for grids beyond L2) a walker parameter
    if (T.feh_free) fehv = theta[w * T.ncol + ndim];
  }
  finite = finite && isfinite(fehv);

  WalkerPlan plan;
  for (int s = 0; s < MAX_STAR; ++s) {
    plan.logg[s] = CUDART_NAN;
    for (int c = 0; c < MAX_CORNER; ++c) { plan.coef[s][c] = 0.0; plan.row[s][c] = 0; }
  }
  plan.av = 0.0; plan.lp = -CUDART_INF; plan.status = 0; plan.cell = T.n_cells - 1;
  if (!finite) { plan.status = MS_W_NONFINITE; store(plan); return; }

  // layout mft6.py:1146: T[0..ns), Av, R1, ratios[1..ns), parallax
  const double* temps = p;
  const double av = p[ns];
  const double* rad = p + ns + 1;          // rad[0] = R1, rad[1] = R2/R1
  const double plx = p[2 * ns + 1];
  // What logprior calls `rad`.  The 6-parameter dist_fit branch with ext == False slices p0[nspec:2*nspec]
  // (mft6.py:1219-1221), i.e. the Av slot and R1, for its bounds and its radius prior; every other branch reads the
  // radii where the likelihood does (mft6.py:1216, :1278-1279, :1340-1342, :1400-1402).
  const double* prad = (ns == 2 && T.dist_fit && !T.ext_flag && mode != MODE_OPTIMIZER) ? p + ns : rad;

  // ---- logprior bounds (mft6.py:1227-1230 / :1287-1290).  fit_spec keeps its proposals inside its own
  // box on the host (mft6.py:984-985, :1069-1103), so the optimiser mode has none here.
  if (mode != MODE_OPTIMIZER) {
    bool out = false;
    for (int s = 0; s < ns; ++s) out = out || (temps[s] > T.tmax) || (temps[s] < T.tmin);
    if (ns == 2) {
      if (T.dist_fit) {
        const double plx_hi = (mode == MODE_NOSPEC) ? 1.0 / 100.0 : 1.0 / 4.0;     // mft6_nospec.py:1228
        for (int s = 0; s < ns; ++s) out = out || (prad[s] < 0.05);
        out = out || (prad[0] > 1.5) || (plx < 1.0 / 3000.0) || (plx > plx_hi);
      } else {
        out = out || (rad[1] < 0.05) || (rad[0] < 0.05);
      }
    } else {
      // triples, mft6.py:1346 / :1408: no upper bound on R1, parallax from 1/1000; without a distance
      // only the two ratios and the sign of the parallax are checked
      if (T.dist_fit) {
        for (int s = 0; s < ns; ++s) out = out || (rad[s] < 0.05);
        out = out || (plx < 1.0 / 1000.0) || (plx > 1.0 / 4.0);
      } else {
        out = out || (rad[1] < 0.05) || (rad[2] < 0.05) || (plx < 0.0);
      }
    }
    if (T.ext_flag && av < 0.0) out = true;
    if (out) { plan.status = MS_W_PRIOR_BOUNDS; store(plan); return; }
  } else if (!(plx > 0.0)) {
    plan.status = MS_W_PRIOR_BOUNDS; store(plan); return;
  }

  // ---- isochrone: log g and luminosity per star (mft6.py:66-98)
  double lum[MAX_STAR];
  for (int s = 0; s < ns; ++s) {
    double g, l;
    if (!interp_table2(s_iso_teff, s_iso_logg, s_iso_lum, T.n_iso, temps[s], &g, &l)) {
      plan.status = MS_W_OFF_ISOCHRONE; store(plan); return;
    }
    plan.logg[s] = g; lum[s] = l;
  }

  // ---- prior terms, summed in the reference's order
  double lp = 0.0;
  if (mode == MODE_OPTIMIZER) {
    // opt_prior terms in chi-square units (mft6.py:833-854, :1030-1046)
    double mu, sig;
    dust_prior_moments(s_dm, s_dust, T.n_dm, T.n_dust_samples, plx, &mu, &sig);
    if (sig == 0.0) sig = 0.05;
    const double avp = opt.av_prior ? opt.av_prior[w] : av;
    { const double z = (avp - mu) / sig; lp += z * z; }
    if (T.dist_fit) { const double z = (plx - T.prior_mu[ndim - 1]) / T.prior_sig[ndim - 1]; lp += z * z; }
    if (T.rad_prior) {
      // opt_prior(var_par[2], [R1m, R2m/R1m (, R3m/R1m)], si[2])   (mft6.py:1037-1050)
      const double r1m = stefan_boltzmann_radius(lum[0], temps[0]);
      for (int s = 0; s < ns; ++s) {
        const double m = (s == 0) ? r1m : stefan_boltzmann_radius(lum[s], temps[s]) / r1m;
        const double z = (rad[s] - m) / opt.rad_sigma[w * ns + s];
        lp += z * z;
      }
    }
  } else {
    if (T.ext_flag) {
      double mu, sig;
      dust_prior_moments(s_dm, s_dust, T.n_dm, T.n_dust_samples, plx, &mu, &sig);
      if (sig == 0.0) sig = 0.05;
      const double z = (av - mu) / sig;
      lp += -0.5 * (z * z);
    }
    const int nprior = T.dist_fit ? ndim : ndim - 2;       // ps has 2*ns+2 or 2*ns entries
    for (int k = 0; k < nprior; ++k) {
      if (T.prior_mu[k] != 0.0) {
        const double z = (p[k] - T.prior_mu[k]) / T.prior_sig[k];
        lp += -0.5 * (z * z);
      }
    }
    if (T.rad_prior) {
      // mft6.py:1262-1269 (binaries), :1381-1389 (triples): 2 % Gaussians on R1 and on the ratios
      const double r1m = stefan_boltzmann_radius(lum[0], temps[0]);
      for (int s = 0; s < ns; ++s) {
        const double m = (s == 0) ? r1m : stefan_boltzmann_radius(lum[s], temps[s]) / r1m;
        const double z = (prad[s] - m) / (0.02 * m);
        lp += -0.5 * (z * z);
      }
    }
  }
  if (!isfinite(lp)) { plan.status = MS_W_NONFINITE; store(plan); return; }   // mft6.py:1465

  // ---- grid brackets and coefficients.  The corners are stored in a canonical order -- lower node first on every
  // axis -- so that all walkers of one grid cell carry the same row list (the group kernel streams those rows once
  // for the whole group); find_nearest's "nearest node first" order (mft6.py:439-453) only decides which weight
  // goes with which node.
  auto canon = [](int i1, int i2, double w2, int n, int* lo, int* hi, double* wlo, double* whi) {
    if (i1 == i2) {
      // on a node: the node above is the (weight 0) partner; on the last node the one below is, so that the
      // corners are always the 2 x 2 (x 2) block starting at the lower node (one tensor-copy box)
      if (i1 < n - 1 || n == 1) { *lo = i1; *hi = min(i1 + 1, n - 1); *wlo = 1.0; *whi = 0.0; }
      else { *lo = n - 2; *hi = n - 1; *wlo = 0.0; *whi = 1.0; }
    }
    else if (i1 < i2) { *lo = i1; *hi = i2; *wlo = 1.0 - w2; *whi = w2; }
    else { *lo = i2; *hi = i1; *wlo = w2; *whi = 1.0 - w2; }
  };
  int z1 = 0, z2 = 0; double wz = 0.0;
  if (T.n_feh > 1 && !bracket_nodes(s_feh, T.n_feh, fehv, &z1, &z2, &wz)) {
    plan.status = MS_W_OFF_GRID; store(plan); return;
  }
  int zlo = 0, zhi = 0; double czl = 1.0, czh = 0.0;
  if (T.n_feh > 1) canon(z1, z2, wz, T.n_feh, &zlo, &zhi, &czl, &czh);
  const double di = 1.0 / plx;                            // mft6.py:690
  for (int s = 0; s < ns; ++s) {
    int t1, t2, g1, g2; double wt, wg;
    if (!bracket_nodes(s_teff, T.n_teff, temps[s], &t1, &t2, &wt) ||
        !bracket_nodes(s_logg, T.n_logg, plan.logg[s], &g1, &g2, &wg)) {
      plan.status = MS_W_OFF_GRID; store(plan); return;
    }
    const double base = (s == 0) ? rad[0] * RSUN_CM / (di * PC_CM) : rad[0] * rad[s] * RSUN_CM / (di * PC_CM);
    const double scale = base * base;                     // mft6.py:691, :700
    int it[2], ig[2]; double ct[2], cg[2];
    canon(t1, t2, wt, T.n_teff, &it[0], &it[1], &ct[0], &ct[1]);
    canon(g1, g2, wg, T.n_logg, &ig[0], &ig[1], &cg[0], &cg[1]);
    const double cz[2] = {czl, czh};
    const int iz[2] = {zlo, zhi};
    // corners in the order a (pixel, Teff, log g[, Fe/H]) tensor box lands in shared memory, Teff fastest:
    // (t_lo,g_lo) (t_hi,g_lo) (t_lo,g_hi) (t_hi,g_hi) [x z_lo, z_hi]
    const int nz = (T.n_feh > 1) ? 2 : 1;
    int c = 0;
    for (int a = 0; a < nz; ++a)
      for (int d = 0; d < 2; ++d)
        for (int b = 0; b < 2; ++b) {
          plan.row[s][c] = (iz[a] * T.n_logg + ig[d]) * T.n_teff + it[b];
          plan.coef[s][c] = ((nz == 2) ? cz[a] : 1.0) * ct[b] * cg[d] * scale;
          ++c;
        }
  }
  {
    // cell key: lower Teff / log g node of each star's bracket
    int key = 0;
    const int per_star = T.n_teff * T.n_logg;
    for (int s = 0; s < 2; ++s) {                                       // the first two stars name the cell
      const int r_a = plan.row[s][0] % per_star, r_b = plan.row[s][3] % per_star;   // (t1,g1) and (t2,g2)
      const int tl = min(r_a % T.n_teff, r_b % T.n_teff), gl = min(r_a / T.n_teff, r_b / T.n_teff);
      key = key * per_star + gl * T.n_teff + tl;
    }
    if (T.n_feh > 1) key = key * T.n_feh + zlo;
    plan.cell = (T.n_cells > 1) ? min(key, T.n_cells - 2) : 0;
  }
  plan.av = ((T.ext_flag || mode == MODE_OPTIMIZER) && av > 0.0) ? av : 0.0;   // mft6.py:1161, :1001
  plan.lp = lp;
  store(plan);
}

// ------------------------------------------------------------------------------------------
// Launch order: counting sort of the walkers by grid cell, so that CTAs resident at the same time
// stream the same rows and find them in L2 (matters once the grid exceeds L2: config C5).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) cell_scan_kernel(unsigned* __restrict__ hist, int n_cells) {
  // exclusive prefix sum in place, single CTA
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_cells; base += 1024) {
    const int i = base + tid;
    const unsigned v = (i < n_cells) ? hist[i] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const unsigned before = carry + (warp ? warp_tot[warp - 1] : 0u) + incl - v;
    if (i < n_cells) hist[i] = before;
    __syncthreads();
    if (tid == 1023) carry = before + v;
    __syncthreads();
  }
}

__global__ void cell_keys_kernel(const WalkerPlan* __restrict__ plans, long long nw, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (w >= nw) return;
  keys[w] = (unsigned)plans[w].cell; vals[w] = (int)w;
}

__global__ void cell_scatter_kernel(const WalkerPlan* __restrict__ plans, long long nw, unsigned* __restrict__ offsets,
                                    int* __restrict__ order) {
  const long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (w >= nw) return;
  // warp-aggregated: one atomic per distinct cell and warp, the lanes of a cell take consecutive slots
  const int cell = plans[w].cell;
  const int lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(__activemask(), cell);
  const int leader = __ffs(peers) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(offsets + cell, (unsigned)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  order[base + __popc(peers & ((1u << lane) - 1u))] = (int)w;
}

// ------------------------------------------------------------------------------------------
// Exact median of sM[0..n) (np.median: mean of the two middle order statistics for even n) by an
// MSD radix select, fused with the chi-square pass.
//
// Because norm_spec's fit is linear in data/model and the median scaling is a common factor
// (mft6.py:1173-1179),  chi2 = s^2 * sum_i ((m_i - d_i / p0(x_i))^2 / e_i^2)  with m the
// *unscaled* model, p0 the quadratic fitted to d/m and s = median(d)/median(m): the chi-square
// sum does not need the median and shares the first histogram pass over sM.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ordered_key(double v) {
  const long long b = __double_as_longlong(v);
  return (b < 0) ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

constexpr int LIST_CAP = 256;            // finish by counting once the bin holds at most this many keys
static_assert(LIST_CAP <= 8 * SEG, "the candidate list lives in the (idle) composite tile buffer");

struct SelectScratch {
  unsigned hist[256];
  unsigned long long kmin[MAXWARP], kmax[MAXWARP];
  unsigned long long bcast[4];
  unsigned count;
  int nanflag;
};

// Locate the bin holding `rank` in S->hist (256 bins).  Warp 0 only; writes
// bcast[0] = bin, bcast[1] = elements below the bin, bcast[2] = elements in the bin.
__device__ __forceinline__ void locate_rank_bin(SelectScratch* S, unsigned rank, int lane) {
  unsigned c[8]; unsigned tot = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = S->hist[lane * 8 + j]; tot += c[j]; }
  unsigned incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  const unsigned excl = incl - tot;
  if (rank >= excl && rank < incl) {                   // exactly one lane
    unsigned run = excl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (rank >= run && rank < run + c[j]) {
        S->bcast[0] = (unsigned long long)(lane * 8 + j); S->bcast[1] = run; S->bcast[2] = c[j];
      }
      run += c[j];
    }
  }
}

// One histogram pass over the elements whose key matches `prefix` on its top `bits_done` bits.
// Warp-aggregated shared-memory atomics (neighbouring pixels of a smooth spectrum share a bin).
template <int NT>
__device__ __forceinline__ void histogram_pass(const double* sM, int n, SelectScratch* S, unsigned long long prefix,
                                               int bits_done, int shift, int width) {
  const int tid = threadIdx.x % NT, lane = tid & 31;
  const unsigned long long himask = (bits_done == 0) ? 0ull : (~0ull << (64 - bits_done));
  const int rounds = (n + NT - 1) / NT;
  for (int r = 0; r < rounds; ++r) {
    const int i = r * NT + tid;
    unsigned d = 0xffffffffu;
    if (i < n) {
      const unsigned long long k = ordered_key(sM[i]);
      if ((k & himask) == prefix) d = (unsigned)((k >> shift) & ((1u << width) - 1u));
    }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (d != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&S->hist[d], (unsigned)__popc(peers));
  }
}

// Median of sM[0..n) given the block-wide key range [kmin, kmax]; also returns, through *chi_out,
// sum_i (m_i - d_i/p0(x_i))^2 * inv_e2_i with p0 = a[0] q0 + a[1] q1 + a[2] q2.
// `list` is scratch for LIST_CAP keys.
template <int NT, bool DO_CHI>
__device__ double block_median_chi(const double* sM, int n, SelectScratch* S, unsigned long long* list,
                                   unsigned long long kmin, unsigned long long kmax, const DevTables& T,
                                   const double (&a)[3], double* chi_out, double* red_scratch, int bar) {
  constexpr int NWARP = NT / 32;
  const int tid = threadIdx.x % NT, lane = tid & 31, warp = tid >> 5;
  const int rounds = (n + NT - 1) / NT;
  unsigned rank = (unsigned)((n - 1) >> 1);
  const bool flat = (kmin == kmax);
  int bits_done = flat ? 64 : __clzll((long long)(kmin ^ kmax));
  unsigned long long prefix = (bits_done == 0) ? 0ull : (bits_done == 64 ? kmin : (kmin & (~0ull << (64 - bits_done))));
  const int width0 = flat ? 0 : min(8, 64 - bits_done);
  const int shift0 = 64 - bits_done - width0;

  // ---- pass 1: chi-square terms + histogram of the first 8 differing key bits
  for (int b = tid; b < 256; b += NT) S->hist[b] = 0;
  if (tid == 0) S->count = 0;
  gsync<NT>(bar);
  double chi = 0.0;
  const double b0 = T.basis[0], b3 = T.basis[3], b4 = T.basis[4], b6 = T.basis[6], b7 = T.basis[7], b8 = T.basis[8];
  const double a0b0 = a[0] * b0;
  constexpr int U = MS_CHI_UNROLL;                           // elements per thread and iteration
  for (int r = 0; r < rounds; r += U) {
    // all table loads of the iteration are issued before any is consumed
    double m[U], x[U], d[U], wgt[U];
    bool v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = (r + u) * NT + tid;
      v[u] = (r + u < rounds) && (i < n);
      m[u] = 0.0; x[u] = 0.0; d[u] = 0.0; wgt[u] = 0.0;
      if (v[u]) {
        m[u] = sM[i];
        if constexpr (DO_CHI) { const dbl4 rec = ldg_cached4(T.data_rec + 4 * (long long)i); x[u] = rec.x; d[u] = rec.y; wgt[u] = rec.z; }
      }
    }
    unsigned dg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      dg[u] = 0xffffffffu;
      if (v[u]) {
        if constexpr (DO_CHI) {
          const double pfit = a0b0 + a[1] * fma(b4, x[u], b3) + a[2] * fma(fma(b8, x[u], b7), x[u], b6);
          const double res = m[u] - d[u] * fast_rcp(pfit);                // mft6.py:196, :120 (unscaled)
          chi = fma(res * res, wgt[u], chi);
        }
        if (!flat) dg[u] = (unsigned)((ordered_key(m[u]) >> shift0) & ((1u << width0) - 1u));
      }
    }
    if (!flat) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned peers = __match_any_sync(0xffffffffu, dg[u]);
        if (dg[u] != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&S->hist[dg[u]], (unsigned)__popc(peers));
      }
    }
  }
  {
    double c1[1] = {chi};
    block_sum<1, NWARP>(c1, red_scratch, bar);                 // includes the barriers that publish hist
    *chi_out = c1[0];
  }
  if (flat) return key_value(kmin);

  // ---- narrow down to a bin small enough to finish by counting
  int width = width0, shift = shift0;
  unsigned in_bin = 0;
  while (true) {
    if (warp == 0) locate_rank_bin(S, rank, lane);
    gsync<NT>(bar);
    const unsigned digit = (unsigned)S->bcast[0];
    rank -= (unsigned)S->bcast[1];
    in_bin = (unsigned)S->bcast[2];
    prefix |= ((unsigned long long)digit) << shift;
    bits_done += width;
    gsync<NT>(bar);
    if (in_bin <= (unsigned)LIST_CAP || bits_done >= 64) break;
    width = min(8, 64 - bits_done);
    shift = 64 - bits_done - width;
    for (int b = tid; b < 256; b += NT) S->hist[b] = 0;
    gsync<NT>(bar);
    histogram_pass<NT>(sM, n, S, prefix, bits_done, shift, width);
    gsync<NT>(bar);
  }

  double v1, v2;
  bool need_next = false;                                 // upper middle element lies beyond this bin
  if (bits_done >= 64) {
    // all 64 bits fixed: `in_bin` copies of one value
    v1 = key_value(prefix);
    v2 = v1;
    need_next = (rank + 1 >= in_bin);
  } else {
    // ---- compact the bin's keys, then rank them by counting
    const unsigned long long himask = ~0ull << (64 - bits_done);
    for (int i = tid; i < n; i += NT) {
      const unsigned long long k = ordered_key(sM[i]);
      if ((k & himask) == prefix) list[atomicAdd(&S->count, 1u)] = k;
    }
    gsync<NT>(bar);
    const unsigned cnt = S->count;                         // == in_bin
    for (unsigned t = tid; t < cnt; t += NT) {
      const unsigned long long mine = list[t];
      unsigned below = 0;
      for (unsigned j = 0; j < cnt; ++j) {
        const unsigned long long o = list[j];
        below += (o < mine) || (o == mine && j < t);
      }
      if (below == rank) S->bcast[0] = mine;
      if (below == rank + 1) S->bcast[1] = mine;
    }
    gsync<NT>(bar);
    v1 = key_value(S->bcast[0]);
    v2 = key_value(S->bcast[1]);
    need_next = (rank + 1 >= cnt);
    gsync<NT>(bar);
  }
  if (n & 1) return v1;
  if (need_next) {
    // rare: the lower median is the largest element of its bin -> upper median = min{key > v1}
    const unsigned long long k1 = ordered_key(v1);
    unsigned long long kgt = ~0ull;
    for (int i = tid; i < n; i += NT) {
      const unsigned long long k = ordered_key(sM[i]);
      if (k > k1) kgt = min(kgt, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kgt = min(kgt, __shfl_xor_sync(0xffffffffu, kgt, o));
    if (lane == 0) S->kmin[warp] = kgt;
    gsync<NT>(bar);
    kgt = S->kmin[0];
#pragma unroll
    for (int w = 1; w < NWARP; ++w) kgt = min(kgt, S->kmin[w]);
    gsync<NT>(bar);
    v2 = key_value(kgt);
  }
  return (v1 + v2) / 2.0;
}

// One observed pixel: extinct() at the two bracketing model pixels, interp1d onto the observed wavelength
// (mft6.py:1162, :1169), then the bookkeeping of phase B (norm_spec moments in the orthonormal basis on the
// unscaled model, mft6.py:194-195; key range and NaN flag for the median).  Returns the model value.
template <int MODE>
__device__ __forceinline__ double observed_pixel(const DevTables& T, const dbl4& rec, const dbl4& drec, double flo, double fhi,
                                                 double av, bool small_dext, double (&mom)[3], unsigned long long& kmin,
                                                 unsigned long long& kmax, int& has_nan) {
  const double elo = exp2(av * rec.y);
  double ehi;
  if (small_dext) {
    const double xx = av * rec.z;                                // E_hi = E_lo * exp(xx), |xx| < 4e-3
    double pz = fma(xx, 1.0 / 120.0, 1.0 / 24.0);
    pz = fma(pz, xx, 1.0 / 6.0); pz = fma(pz, xx, 0.5); pz = fma(pz, xx, 1.0); pz = fma(pz, xx, 1.0);
    ehi = elo * pz;
  } else {
    ehi = exp2(av * fma(rec.z, 1.4426950408889634, rec.y));   // ext_hi = ext_lo + dext / ln 2
  }
  const double ylo = flo * elo, yhi = fhi * ehi;
  const double m = ylo + (yhi - ylo) * rec.x;
  if constexpr (MODE == MODE_POSTERIOR) {
    const double frac = drec.y * fast_rcp(m);
    const double x = drec.x;
    const double q1 = fma(T.basis[4], x, T.basis[3]);
    const double q2 = fma(fma(T.basis[8], x, T.basis[7]), x, T.basis[6]);
    mom[0] = fma(T.basis[0], frac, mom[0]); mom[1] = fma(q1, frac, mom[1]); mom[2] = fma(q2, frac, mom[2]);
  }
  has_nan |= (m != m);
  const unsigned long long k = ordered_key(m);
  kmin = min(kmin, k); kmax = max(kmax, k);
  return m;
}

#include "group_kernel.cuh"

// ------------------------------------------------------------------------------------------
// lnprob_stream_kernel: one thread group of NT threads per walker, G groups per CTA.
//   NR  grid rows per star (4: Teff x log g, 8: with [Fe/H]);  NC / NP compile-time filter caps;
//   NS   number of stars (2, or 3 for triples: mft6.py nspec);
//   MODE what is evaluated (MODE_POSTERIOR / MODE_NOSPEC / MODE_OPTIMIZER);
//   NT  threads per walker: 256, or 512 when the observed spectrum is so long that only one walker
//       fits an SM;
//   G   walkers per CTA.  G = 1: one CTA per walker, two CTAs per SM.  G = 2: the CTA evaluates two
//       walkers that the host made neighbours in the launch order because they sit in the same grid
//       cell (same bracketing rows).  Warp k of either group owns the same chunks, the two warps meet
//       at a two-warp named barrier before every chunk and then request the same row sectors, filter
//       weights and per-pixel records back to back, so the second request is served by the SM's L1
//       (hit or pending-miss merge) instead of L2.  Each walker's arithmetic is exactly that of G = 1;
//       the pairing only changes where its operands come from, never its value.
//   FAST 1: bracket-only evaluation (ms_set_fast; SURVEY.md section 8(d), "algebraic note").  Interpolation, dilution,
//       trapezoid integration and linear rebinning are all linear in the grid rows and the interpolation
//       weights do not depend on wavelength, so (i) the spectral term needs only the two model pixels that
//       bracket each observed pixel -- a compact [row][N][2] copy of the grid, read fully coalesced -- and (ii)
//       every band flux of a component is the same weighted sum of per-row band fluxes tabulated once.  The model
//       vector is bit-identical to the streaming path, moments and band sums agree to rounding.  The streaming path stays
//       the default.
// ------------------------------------------------------------------------------------------
template <int NR, int NC, int NP, int NT, int MODE, int NS, int G, int FAST>
__global__ void __launch_bounds__(NT * G, (NT * G <= 256) ? 2 : 1)
lnprob_stream_kernel(const __grid_constant__ DevTables T, const WalkerPlan* __restrict__ plans,
                     const int* __restrict__ order, double* __restrict__ lnprob, const PartsOut parts,
                     const OptIn opt, const int csize, const long long nw, const long long pair_stride,
                     const unsigned group_bytes, double* __restrict__ model_scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NWARP = NT / 32;
  constexpr int KRED = NS * NC + NP;
  const int grp = (G > 1) ? (int)(threadIdx.x / NT) : 0;             // which of the CTA's walkers this thread serves
  const int bar = (G > 1) ? 1 + grp : 0;                             // the group's named barrier (0 == __syncthreads)
  unsigned char* smem_grp = smem_raw + (size_t)grp * group_bytes;
  double* sF = reinterpret_cast<double*>(smem_grp);                  // [NWARP*SEG] per-warp composite chunks
  double* sRed = sF + NWARP * SEG;                                          // [NWARP*KRED]
  SelectScratch* sSel = reinterpret_cast<SelectScratch*>(sRed + NWARP * (KRED > 4 ? KRED : 4));
  // [n_data] model on the observed pixels: in shared memory, or -- observed spectra too long for it (more than
  // ~27 000 pixels) -- in a global scratch row per launch slot, which then lives in L2 between the two phases
  double* sM = model_scratch ? model_scratch + (size_t)(blockIdx.x / csize) * (size_t)T.n_data
                             : reinterpret_cast<double*>(sSel + 1);
  __shared__ WalkerPlan sPlanG[G];
  WalkerPlan& sPlan = sPlanG[grp];

  // Small ensembles (the reference runs 50 walkers) cannot fill 148 SMs with one CTA per walker: a
  // thread-block cluster of `csize` CTAs then shares a walker.  Each CTA streams every csize-th group of
  // chunks, rebinned pixels are stored straight into the leader CTA's model vector through distributed
  // shared memory, the band sums are combined over the cluster, and the leader alone runs phase B.
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = (csize > 1) ? cluster.block_rank() : 0u;
  __shared__ double sExport[MAX_STAR * MS_MAX_CONTRAST + MS_MAX_PHOT + 3];
  __shared__ unsigned long long sExportKey[3];

  const int tid = (G > 1) ? (int)(threadIdx.x % NT) : (int)threadIdx.x;
  long long slot;
  if constexpr (G > 1) {
    // CTA b takes pair (b * pair_stride) mod n_pairs of the cell-sorted order: neighbours in time come from
    // different cells (spreads the L2 slices in use), the two walkers of a CTA from the same one
    const long long n_pairs = (nw + G - 1) / G;
    slot = (((long long)blockIdx.x * pair_stride) % n_pairs) * G + grp;
  } else {
    slot = blockIdx.x / csize;
  }
  const bool present = slot < nw;
  const long long w = present ? (order ? order[slot] : slot) : 0;   // launch order: grouped by grid cell
  if (present) {
    // cooperative copy of the plan (16-byte chunks)
    const int4* src = reinterpret_cast<const int4*>(plans + w);
    int4* dst = reinterpret_cast<int4*>(&sPlan);
    if (tid < (int)(sizeof(WalkerPlan) / sizeof(int4))) dst[tid] = src[tid];
  } else if (tid == 0) {
    sPlan.status = MS_W_NONFINITE;
  }
  __syncthreads();
  // the two warps of a pair synchronise only while both walkers stream (uniform over the CTA)
  bool paired = false;
  if constexpr (G > 1) paired = (sPlanG[0].status == 0) && (sPlanG[G - 1].status == 0);
  if (!present) return;
  const int status = sPlan.status;
  if (status != 0) {                                                  // mft6.py:1465-1466
    if (crank != 0) return;
    if (tid == 0) {
      lnprob[w] = (MODE == MODE_OPTIMIZER) ? CUDART_INF : -CUDART_INF;
      if (parts.status) parts.status[w] = status;
      if (parts.lnprior) parts.lnprior[w] = -CUDART_INF;
      if (parts.chi2_spec) parts.chi2_spec[w] = CUDART_NAN;
      if (parts.logg) for (int s = 0; s < T.nspec; ++s) parts.logg[w * T.nspec + s] = sPlan.logg[s];
      if (parts.contrast) for (int c = 0; c < T.n_contrast; ++c) parts.contrast[w * T.n_contrast + c] = CUDART_NAN;
      if (parts.phot) for (int f = 0; f < T.n_phot; ++f) parts.phot[w * T.n_phot + f] = CUDART_NAN;
      if (parts.comp_mag) for (int f = 0; f < T.n_contrast * NS; ++f) parts.comp_mag[w * T.n_contrast * NS + f] = CUDART_NAN;
    }
    if (parts.model) for (int i = tid; i < T.n_data; i += NT) parts.model[w * T.n_data + i] = CUDART_NAN;
    return;
  }

  const int lane = tid & 31, warp = tid >> 5;
  double cf[NS][NR];
  const double* pr[NS][NR];                                          // this lane's pixel pair in each bracketing row,
#pragma unroll                                                       // advanced past the other warps' chunks per step
  for (int s = 0; s < NS; ++s) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      cf[s][r] = sPlan.coef[s][r];
      pr[s][r] = T.grid + (long long)sPlan.row[s][r] * T.stride + (crank * NWARP + warp) * CHUNK + 4 * lane;
    }
  }
  const double av = sPlan.av;
  const bool small_dext = av * T.dext_max < 4e-3;                     // exp(x) by a 5th-order polynomial: x^6/720 < 6e-18
  double* seg = sF + warp * SEG;                                      // see seg_slot()

  double accC[NC > 0 ? NC : 1][NS];
  double accP[NP > 0 ? NP : 1];
#pragma unroll
  for (int f = 0; f < NC; ++f) {
#pragma unroll
    for (int st = 0; st < NS; ++st) accC[f][st] = 0.0;
  }
#pragma unroll
  for (int f = 0; f < NP; ++f) accP[f] = 0.0;

  double mom[3] = {0.0, 0.0, 0.0};                                   // sum_i q_k(x_i) d_i / m_i
  unsigned long long kmin = ~0ull, kmax = 0ull;                      // key range of the model, for the median
  int has_nan = 0;

  if constexpr (FAST) {
    // ---------------------------------------------------------------- phase A, bracket-only
    if constexpr (MODE != MODE_NOSPEC) {
      const int n_obs = T.n_data;
      const double2* br[NS][NR];
#pragma unroll
      for (int st = 0; st < NS; ++st) {
#pragma unroll
        for (int r = 0; r < NR; ++r) br[st][r] = T.bracket + (long long)sPlan.row[st][r] * n_obs;
      }
      for (int i = tid; i < n_obs; i += NT) {
        const dbl4 rec = ldg_cached4(T.rebin_rec + 4 * (long long)i);
        dbl4 drec; drec.x = drec.y = drec.z = drec.w = 0.0;
        if constexpr (MODE == MODE_POSTERIOR) drec = ldg_cached4(T.data_rec + 4 * (long long)i);
        double2 b[NS][NR];
#pragma unroll
        for (int st = 0; st < NS; ++st) {
#pragma unroll
          for (int r = 0; r < NR; ++r) b[st][r] = __ldg(br[st][r] + i);
        }
        // same operation order as the streaming path: per star an FMA chain over its rows, then the sum over stars
        double flo = 0.0, fhi = 0.0;
#pragma unroll
        for (int st = 0; st < NS; ++st) {
          double l = cf[st][0] * b[st][0].x, u = cf[st][0] * b[st][0].y;
#pragma unroll
          for (int r = 1; r < NR; ++r) { l = fma(cf[st][r], b[st][r].x, l); u = fma(cf[st][r], b[st][r].y, u); }
          flo = (st == 0) ? l : flo + l; fhi = (st == 0) ? u : fhi + u;
        }
        sM[i] = observed_pixel<MODE>(T, rec, drec, flo, fhi, av, small_dext, mom, kmin, kmax, has_nan);
      }
    }
  } else {
  // ------------------------------------------------------------------ phase A: stream the rows
  // Each warp owns every NWARP-th contiguous 128-pixel chunk of the region and the observed pixels
  // whose upper bracket falls into it, so the rebinning needs no block-wide barrier and the warps
  // drift apart: while some wait on their row loads others compute.
  const int n_chunks = T.n_tiles * CHUNKS_PER_TILE;
  const int cstep = csize * NWARP;
  double* sM_dst = (csize > 1) ? cluster.map_shared_rank(sM, 0) : sM;   // the leader's model vector
  for (int c = crank * NWARP + warp; c < n_chunks; c += cstep) {
    const int cbase = c * CHUNK;                                      // first pixel of this warp's chunk
    const int4 meta = __ldg(T.chunk_meta + c);
    const unsigned cmask = (unsigned)meta.x, pmask = (unsigned)meta.y;
    const int ds_raw = meta.z;
    const int d0 = ds_raw & 0x7fffffff;
    const int d1 = meta.w;
    const bool stage = (MODE != MODE_NOSPEC) && d1 > d0;
    const bool need_prev = stage && ds_raw < 0;                        // warp-uniform
    if constexpr (G > 1) {
      if (paired) asm volatile("bar.sync %0, 64;" :: "r"(1 + G + warp) : "memory");   // meet the partner's warp
    }
    if (need_prev && lane == 0) {
      // composite of the pixel just before the chunk (owned by another warp / the previous tile)
      double tot = 0.0;
#pragma unroll
      for (int st = 0; st < NS; ++st) {
        double f = 0.0;
#pragma unroll
        for (int r = 0; r < NR; ++r) f = fma(cf[st][r], __ldg(pr[st][r] - 1), f);
        tot += f;
      }
      seg[seg_slot(0)] = tot;
    }
    {
      const int px = cbase + 4 * lane;                                // this lane's four pixels of the chunk
      dbl4 fs[NS];                                                    // their flux, per star
      if constexpr (NS * NR <= 8) {
        // all rows of all stars in flight at once (binaries on a Teff x log g grid): 8 x 32 bytes per lane
        dbl4 a[NS][NR];
#pragma unroll
        for (int st = 0; st < NS; ++st) {
#pragma unroll
          for (int r = 0; r < NR; ++r) a[st][r] = (G > 1) ? ldg_cached4(pr[st][r]) : ldg_stream4(pr[st][r]);
        }
#pragma unroll
        for (int st = 0; st < NS; ++st) {
          const double c = cf[st][0];
          fs[st].x = c * a[st][0].x; fs[st].y = c * a[st][0].y; fs[st].z = c * a[st][0].z; fs[st].w = c * a[st][0].w;
#pragma unroll
          for (int r = 1; r < NR; ++r) {
            const double cr = cf[st][r];
            fs[st].x = fma(cr, a[st][r].x, fs[st].x); fs[st].y = fma(cr, a[st][r].y, fs[st].y);
            fs[st].z = fma(cr, a[st][r].z, fs[st].z); fs[st].w = fma(cr, a[st][r].w, fs[st].w);
          }
        }
      } else {
        // 8 corners per star or three stars: one star at a time
#pragma unroll
        for (int st = 0; st < NS; ++st) {
          dbl4 a[NR];
#pragma unroll
          for (int r = 0; r < NR; ++r) a[r] = (G > 1) ? ldg_cached4(pr[st][r]) : ldg_stream4(pr[st][r]);
          const double c = cf[st][0];
          fs[st].x = c * a[0].x; fs[st].y = c * a[0].y; fs[st].z = c * a[0].z; fs[st].w = c * a[0].w;
#pragma unroll
          for (int r = 1; r < NR; ++r) {
            const double cr = cf[st][r];
            fs[st].x = fma(cr, a[r].x, fs[st].x); fs[st].y = fma(cr, a[r].y, fs[st].y);
            fs[st].z = fma(cr, a[r].z, fs[st].z); fs[st].w = fma(cr, a[r].w, fs[st].w);
          }
        }
      }
      dbl4 comp = fs[0];                                              // mft6.py:744 / :753
#pragma unroll
      for (int st = 1; st < NS; ++st) { comp.x += fs[st].x; comp.y += fs[st].y; comp.z += fs[st].z; comp.w += fs[st].w; }
#pragma unroll
      for (int f = 0; f < NC; ++f) {
        if ((cmask >> f) & 1u) {                                      // warp-uniform
          const dbl4 wv = ldg_cached4(T.c_wbase[f] + px);
#pragma unroll
          for (int st = 0; st < NS; ++st) {
            double acc = accC[f][st];
            acc = fma(wv.x, fs[st].x, acc); acc = fma(wv.y, fs[st].y, acc);
            acc = fma(wv.z, fs[st].z, acc); acc = fma(wv.w, fs[st].w, acc);
            accC[f][st] = acc;
          }
        }
      }
#pragma unroll
      for (int f = 0; f < NP; ++f) {
        if ((pmask >> f) & 1u) {
          const dbl4 wv = ldg_cached4(T.p_wbase[f] + px);
          double acc = accP[f];
          acc = fma(wv.x, comp.x, acc); acc = fma(wv.y, comp.y, acc); acc = fma(wv.z, comp.z, acc); acc = fma(wv.w, comp.w, acc);
          accP[f] = acc;
        }
      }
      if (stage) {
        seg[seg_slot(4 * lane + 1)] = comp.x; seg[seg_slot(4 * lane + 2)] = comp.y;
        seg[seg_slot(4 * lane + 3)] = comp.z; seg[seg_slot(4 * lane + 4)] = comp.w;
      }
    }
#pragma unroll
    for (int st = 0; st < NS; ++st) {
#pragma unroll
      for (int r = 0; r < NR; ++r) pr[st][r] += cstep * CHUNK;
    }
    if (stage) {
      __syncwarp();
      for (int i = d0 + lane; i < d1; i += 32) {
        const dbl4 rec = ldg_cached4(T.rebin_rec + 4 * (long long)i);
        dbl4 drec; drec.x = drec.y = drec.z = drec.w = 0.0;            // issued with the other table load
        if constexpr (MODE == MODE_POSTERIOR) drec = ldg_cached4(T.data_rec + 4 * (long long)i);
        const int hl = (int)rec.w - cbase;                            // 0 .. CHUNK-1
        const double flo = seg[seg_slot(hl)];
        const double fhi = seg[seg_slot(hl + 1)];
        sM_dst[i] = observed_pixel<MODE>(T, rec, drec, flo, fhi, av, small_dext, mom, kmin, kmax, has_nan);
      }
      __syncwarp();
    }
  }
  }  // !FAST
  gsync<NT>(bar);

  // ------------------------------------------------------------------ band integrals
  double red[KRED > 0 ? KRED : 1];
  if constexpr (FAST) {
    // component band fluxes from the per-row table (only thread 0 needs them, for the final sum)
#pragma unroll
    for (int k = 0; k < KRED; ++k) red[k] = 0.0;
    if (tid == 0) {
#pragma unroll
      for (int f = 0; f < NC; ++f) {
#pragma unroll
        for (int st = 0; st < NS; ++st) {
          double acc = 0.0;
#pragma unroll
          for (int r = 0; r < NR; ++r) acc = fma(cf[st][r], __ldg(T.bandsum + (long long)sPlan.row[st][r] * BAND_COLS + f), acc);
          red[NS * f + st] = acc;
        }
      }
#pragma unroll
      for (int f = 0; f < NP; ++f) {
        double tot = 0.0;
#pragma unroll
        for (int st = 0; st < NS; ++st) {
          double acc = 0.0;
#pragma unroll
          for (int r = 0; r < NR; ++r)
            acc = fma(cf[st][r], __ldg(T.bandsum + (long long)sPlan.row[st][r] * BAND_COLS + MS_MAX_CONTRAST + f), acc);
          tot += acc;
        }
        red[NS * NC + f] = tot;
      }
    }
  } else {
#pragma unroll
  for (int f = 0; f < NC; ++f) {
#pragma unroll
    for (int st = 0; st < NS; ++st) red[NS * f + st] = accC[f][st];
  }
#pragma unroll
  for (int f = 0; f < NP; ++f) red[NS * NC + f] = accP[f];
  block_sum<KRED, NWARP>(red, sRed, bar);
  }

  // ------------------------------------------------------------------ phase B: spectral chi-square
  const int n = T.n_data;
  double chi_un = CUDART_NAN, med = CUDART_NAN, scale = 1.0;
  if constexpr (MODE != MODE_NOSPEC) {
    block_sum<3, NWARP>(mom, sRed, bar);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
      kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
      has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o);
    }
    if (tid == 0) sSel->nanflag = 0;
    gsync<NT>(bar);
    if (lane == 0) { sSel->kmin[warp] = kmin; sSel->kmax[warp] = kmax; if (has_nan) sSel->nanflag = 1; }
    gsync<NT>(bar);
    kmin = sSel->kmin[0]; kmax = sSel->kmax[0];
#pragma unroll
    for (int w2 = 1; w2 < NWARP; ++w2) { kmin = min(kmin, sSel->kmin[w2]); kmax = max(kmax, sSel->kmax[w2]); }
    has_nan = sSel->nanflag;
    gsync<NT>(bar);
  }
  if (csize > 1) {
    // combine the cluster's partial sums in rank order (deterministic), then let the helpers go
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < KRED; ++k) sExport[k] = red[k];
      sExport[KRED] = mom[0]; sExport[KRED + 1] = mom[1]; sExport[KRED + 2] = mom[2];
      sExportKey[0] = kmin; sExportKey[1] = kmax; sExportKey[2] = (unsigned long long)has_nan;
    }
    cluster.sync();                                                   // also publishes the remote sM stores
    if (crank == 0) {
      for (int r = 1; r < csize; ++r) {
        const double* e = cluster.map_shared_rank(sExport, r);
        const unsigned long long* ek = cluster.map_shared_rank(sExportKey, r);
#pragma unroll
        for (int k = 0; k < KRED; ++k) red[k] += e[k];
        mom[0] += e[KRED]; mom[1] += e[KRED + 1]; mom[2] += e[KRED + 2];
        kmin = min(kmin, ek[0]); kmax = max(kmax, ek[1]); has_nan |= (int)ek[2];
      }
    }
    cluster.sync();                                                   // helpers' shared memory was read
    if (crank != 0) return;
  }
  if constexpr (MODE != MODE_NOSPEC) {
    unsigned long long* list = reinterpret_cast<unsigned long long*>(sF);
    if constexpr (MODE == MODE_POSTERIOR) {
      if (!has_nan) med = block_median_chi<NT, true>(sM, n, sSel, list, kmin, kmax, T, mom, &chi_un, sRed, bar);  // mft6.py:1173
      scale = T.data_median / med;
    } else {
      // fit_spec: the observed spectrum was renormalised once (mft6.py:889); every proposal is only
      // median-scaled against it (mft6.py:1010-1014), so the scale does not factor out of the sum
      double dummy;
      if (!has_nan) med = block_median_chi<NT, false>(sM, n, sSel, list, kmin, kmax, T, mom, &dummy, sRed, bar);
      const double* dset = opt.flux_sets + (long long)opt.set_index[w] * n;
      scale = (opt.ref_median ? opt.ref_median[w] : T.opt_median[opt.set_index[w]]) / med;   // np.median(flux), mft6.py:1011 (:888 at the start)
      double chi[1] = {0.0};
      for (int i = tid; i < n; i += NT) {
        const double r = scale * sM[i] - dset[i];
        chi[0] = fma(r * r, T.data_inv_e2[i], chi[0]);
      }
      block_sum<1, NWARP>(chi, sRed, bar);
      chi_un = chi[0] / (scale * scale);                              // so that iic below is uniform
    }
    if (parts.model) for (int i = tid; i < n; i += NT) parts.model[w * n + i] = sM[i] * scale;
  }

  // ------------------------------------------------------------------ final sum (mft6.py:1178-1205)
  if (tid == 0) {
    const double iic = (MODE == MODE_NOSPEC) ? 0.0 : (scale * scale) * chi_un / (double)n;   // mft6.py:1173-1179
    double icontrast = 0.0, iphot = 0.0;
    for (int f = 0; f < NC; ++f) {
      if (f < T.n_contrast) {
        // binaries: secondary - primary (mft6.py:741); triples: the first half of the filters measures the
        // secondary, the second half the tertiary, both against the primary (mft6.py:747-749)
        const int other = (NS == 2 || f < T.n_contrast / 2) ? 1 : 2;
        const double m0 = -2.5 * log10(red[NS * f]);
        const double m1 = -2.5 * log10(red[NS * f + other]);
        const double con = m1 - m0;
        const double d = con - T.c_mag[f];
        icontrast += (d * d) / (T.c_err[f] * T.c_err[f]);
        if (parts.contrast) parts.contrast[w * T.n_contrast + f] = con;
        if (parts.comp_mag)
          for (int st = 0; st < NS; ++st) parts.comp_mag[(w * T.n_contrast + f) * NS + st] = -2.5 * log10(red[NS * f + st]);
      }
    }
    for (int f = 0; f < NP; ++f) {
      if (f < T.n_phot) {
        double mag = -2.5 * log10(red[NS * NC + f] * T.p_norm[f]);    // mft6.py:780-782
        mag += av * T.p_ext[f];                                       // mft6.py:1163 (av == 0 when skipped)
        const double d = mag - T.p_mag[f];
        iphot += (d * d) / (T.p_err[f] * T.p_err[f]);
        if (parts.phot) parts.phot[w * T.n_phot + f] = mag;
      }
    }
    double out;
    int st = 0;
    if constexpr (MODE == MODE_OPTIMIZER) {
      // test_cs of fit_spec (mft6.py:1016, :1030-1046): x3 spectral weight, priors in chi-square units
      // mft6_nospec.py's fit_spec comments the spectral term out (its :1016-1030): contrast + photometry + priors
      out = T.nospec ? (icontrast + iphot + sPlan.lp)
                     : ((iic * 3.0) * (double)(T.n_contrast + T.n_phot) + icontrast + iphot + sPlan.lp);
      if (out != out) { out = CUDART_INF; st |= MS_W_NONFINITE; }     // a NaN never wins `test_cs < chi`
    } else {
      const double init_cs = (MODE == MODE_NOSPEC) ? icontrast + iphot                             // mft6_nospec.py:1192
                                                   : iic * (double)(T.n_contrast + T.n_phot) + icontrast + iphot;   // mft6.py:1191
      const double lnl = (init_cs != init_cs) ? -CUDART_INF : -0.5 * init_cs;                     // mft6.py:1202-1205
      out = sPlan.lp + lnl;
      if (out != out) { out = -CUDART_INF; }
      if (init_cs != init_cs) st |= MS_W_NONFINITE;
    }
    lnprob[w] = out;
    if (parts.status) parts.status[w] = st;
    if (parts.lnprior) parts.lnprior[w] = sPlan.lp;
    if (parts.chi2_spec) parts.chi2_spec[w] = iic;
    if (parts.logg) for (int s = 0; s < T.nspec; ++s) parts.logg[w * T.nspec + s] = sPlan.logg[s];
  }
}

// ------------------------------------------------------------------------------------------
// Setup of the bracket-only mode: the two bracketing model pixels of every observed pixel, per grid row, and the
// band flux of every grid row through every filter.
// ------------------------------------------------------------------------------------------
__global__ void bracket_gather_kernel(const double* __restrict__ grid, long long stride, const double* __restrict__ rebin_rec,
                                      int n_data, double2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = blockIdx.y;
  if (i >= n_data) return;
  const int hi = (int)rebin_rec[4 * (long long)i + 3];
  const double* g = grid + row * stride;
  out[row * n_data + i] = make_double2(g[hi - 1], g[hi]);
}

// one CTA per grid row: out[row * BAND_COLS + col] = sum over the band's pixels of w[p] * row[p] (fixed order)
__global__ void __launch_bounds__(256)
bandsum_kernel(const double* __restrict__ grid, long long stride, const double* __restrict__ wbase, int p_lo, int p_hi,
               int col, double* __restrict__ out) {
  __shared__ double s_part[8];
  const long long row = blockIdx.x;
  const double* g = grid + row * stride;
  double acc = 0.0;
  for (int p = p_lo + threadIdx.x; p < p_hi; p += 256) acc = fma(wbase[p], g[p], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    out[row * BAND_COLS + col] = t;
  }
}

// ------------------------------------------------------------------------------------------
// broaden_rows_kernel: 'same' zero-padded convolution of the window of each row with `taps`
// (mft6.py:128; np.convolve semantics), then the 5/10-sample edge patch (mft6.py:129-130).
// grid = (ceil(win_len / BT), n_rows); dynamic smem = (BT + n_taps - 1) doubles + n_taps doubles.
// ------------------------------------------------------------------------------------------
constexpr int BT = 1024;   // output pixels per CTA
constexpr int BR = 8;      // adjacent output pixels per thread (register tile)
constexpr int BTHREADS = BT / BR;

// Shared-memory slot of staged input i: one pad every 8 doubles, so that the threads' stride of BR = 8 inputs
// becomes 9 doubles (odd): the 64-bit loads of a half warp fall into distinct bank pairs.
__device__ __forceinline__ int bpad(int i) { return i + (i >> 3); }

__global__ void __launch_bounds__(BTHREADS)
broaden_rows_kernel(const double* __restrict__ in, double* __restrict__ out, long long row_len,
                    long long win_lo, int win_len, const double* __restrict__ taps, int n_taps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sIn = reinterpret_cast<double*>(smem_raw);        // [bpad(BT + n_taps - 1) + 1]
  double* sTap = sIn + bpad(BT + n_taps - 1) + 1;           // [n_taps]
  const int half = n_taps / 2;                              // odd n_taps: symmetric support
  const long long row = blockIdx.y;
  const int o0 = blockIdx.x * BT;
  const double* src = in + row * row_len + win_lo;
  for (int i = threadIdx.x; i < BT + n_taps - 1; i += BTHREADS) {
    const int j = o0 - half + i;                            // window-relative input pixel
    sIn[bpad(i)] = (j >= 0 && j < win_len) ? src[j] : 0.0;  // zero padding outside the window
  }
  for (int i = threadIdx.x; i < n_taps; i += BTHREADS) sTap[i] = taps[i];
  __syncthreads();
  // np.convolve(y, e, 'same')[o] = sum_k e[k] * y[o + half - k], k ascending (the summation order is kept).
  // Thread t owns outputs o0 + 8t .. 8t+7.  Step k needs the eight inputs 8t + q + m, m = n_taps - 1 - k; step
  // k + 1 needs the same window shifted down by one, so one new input and one (broadcast) tap are loaded per
  // eight FMAs.  The window lives in registers; input q of step k sits in slot (q - k) mod 8, which makes the
  // shift a renaming once the loop is unrolled by eight.
  const int base = BR * threadIdx.x;
  double acc[BR], win[BR];
#pragma unroll
  for (int q = 0; q < BR; ++q) { acc[q] = 0.0; win[q] = sIn[bpad(base + q + n_taps - 1)]; }
  for (int k0 = 0; k0 < n_taps; k0 += BR) {
#pragma unroll
    for (int j = 0; j < BR; ++j) {
      const int k = k0 + j;
      if (k < n_taps) {
        const double tap = sTap[k];
#pragma unroll
        for (int q = 0; q < BR; ++q) acc[q] = fma(tap, win[(q - j + BR) & (BR - 1)], acc[q]);
        const int m_next = n_taps - 2 - k;                  // window offset of step k + 1
        if (m_next >= 0) win[(BR - 1 - j) & (BR - 1)] = sIn[bpad(base + m_next)];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < BR; ++q) {
    const int o = o0 + base + q;
    if (o < win_len) out[row * row_len + win_lo + o] = acc[q];
  }
}

__global__ void broaden_patch_kernel(double* __restrict__ rows, long long row_len, long long win_lo, int win_len) {
  // broad[0:5] = broad[5]; broad[n-10:n] = broad[n-11]        (mft6.py:129-130)
  const long long row = blockIdx.x;
  double* b = rows + row * row_len + win_lo;
  const int t = threadIdx.x;
  if (win_len < 16) return;
  const double head = b[5], tail = b[win_len - 11];
  __syncthreads();
  if (t < 5) b[t] = head;
  else if (t < 15) b[win_len - 10 + (t - 5)] = tail;
}

// ------------------------------------------------------------------------------------------
// resample_rows_kernel: spec_interpolator's ``interp1d(spold, spec1)(wl)`` (mft6.py:369-371) for all
// nodes at once: row r of the output is the node's native spectrum (src_x, src_y)[off[r] .. off[r+1])
// linearly interpolated onto the common wavelength grid x_dst.  Same arithmetic as scipy's linear
// interp1d on 1-D data (np.interp): bracket dx[j] <= x < dx[j+1], slope * (x - dx[j]) + dy[j] with
// separately rounded multiply and add (no FMA contraction), exact node hits return the node value.
// Out-of-range abscissae raise `*oob` (interp1d's bounds_error=True).
// grid = (ceil(row_len / 256), n_rows)
// ------------------------------------------------------------------------------------------
inline int bpad_host(int i) { return i + (i >> 3); }

// largest j with x[j] <= v, for x[0] <= v (ascending x[lo..hi])
__device__ __forceinline__ long long bracket_below(const double* __restrict__ x, long long lo, long long hi, double v) {
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if (x[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
resample_rows_kernel(const double* __restrict__ src_x, const double* __restrict__ src_y,
                     const long long* __restrict__ off, const double* __restrict__ x_dst, long long row_len,
                     double* __restrict__ out, int* __restrict__ oob) {
  constexpr int WIN = 64;                                   // node samples staged per warp
  __shared__ double s_win[8][WIN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long r = blockIdx.y;
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const double* x = src_x + off[r];
  const double* y = src_y + off[r];
  const long long n = off[r + 1] - off[r];
  const bool live = j < row_len;
  const double v = live ? x_dst[j] : 0.0;
  const bool inside = live && (v >= x[0] && v <= x[n - 1]);
  if (live && !inside) { atomicOr(oob, (v < x[0]) ? 1 : 2); out[r * row_len + j] = CUDART_NAN; }
  // A warp's 32 ascending abscissae bracket into one short, monotone stretch of the node spectrum.  Its start
  // is found once per warp by a 32-ary search (every lane probes one splitter per round: 4 rounds for 10^5-10^6
  // samples instead of 17-20 dependent loads per lane), the stretch is staged in shared memory with two coalesced
  // loads, and each lane finishes with a 6-step search there.
  const unsigned in_mask = __ballot_sync(0xffffffffu, inside);
  if (in_mask == 0u) return;
  const int first = __ffs(in_mask) - 1;
  const double v0 = __shfl_sync(0xffffffffu, v, first);     // smallest in-range abscissa of the warp
  long long lo = 0, hi = n - 1;                             // invariant: x[lo] <= v0, and (hi == n - 1 or x[hi] > v0)
  while (hi - lo > 32) {
    const long long step = (hi - lo) / 32;
    const long long idx = lo + (long long)(lane + 1) * step;
    const bool le = (idx < hi) && (x[idx] <= v0);
    const int c = __popc(__ballot_sync(0xffffffffu, le));   // probes are ascending: c leading ones
    const long long lo_new = lo + (long long)c * step;
    if (c < 32 && lo + (long long)(c + 1) * step < hi) hi = lo + (long long)(c + 1) * step;
    lo = lo_new;
  }
  {
    const long long idx = lo + lane + 1;
    const bool le = (idx < hi) && (x[idx] <= v0);
    lo += __popc(__ballot_sync(0xffffffffu, le));
  }
  // stage x[lo .. lo + WIN)
  for (int k = lane; k < WIN; k += 32) s_win[warp][k] = (lo + k < n) ? x[lo + k] : CUDART_INF;
  __syncwarp();
  if (!inside) return;
  double res;
  if (v == x[n - 1]) {
    res = y[n - 1];
  } else {
    long long b;                                            // largest index with x[b] <= v
    if (s_win[warp][WIN - 1] > v) {
      int l = 0, h = WIN - 1;                               // s_win[l] <= v < s_win[h]
      while (h - l > 1) { const int mid = (l + h) >> 1; if (s_win[warp][mid] <= v) l = mid; else h = mid; }
      b = lo + l;
    } else {
      b = bracket_below(x, lo + WIN - 1, n - 1, v);         // node sampled much finer than the output grid
    }
    if (x[b] == v) {
      res = y[b];
    } else {
      const double slope = __ddiv_rn(__dsub_rn(y[b + 1], y[b]), __dsub_rn(x[b + 1], x[b]));
      res = __dadd_rn(__dmul_rn(slope, __dsub_rn(v, x[b])), y[b]);
      if (res != res) {                                     // np.interp: "if we get nan in one direction, try the other"
        res = __dadd_rn(__dmul_rn(slope, __dsub_rn(v, x[b + 1])), y[b + 1]);
        if (res != res && y[b] == y[b + 1]) res = y[b];
      }
    }
  }
  out[r * row_len + j] = res;
}

// ------------------------------------------------------------------------------------------
// Stretch move (emcee RedBlueMove + StretchMove, a = 2) with a counter-based generator.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// uniform in (0,1): stream (seed, iteration, split, purpose), counter = global walker id
__host__ __device__ __forceinline__ double counter_uniform(unsigned long long seed, unsigned long long iteration,
                                                           int split, int purpose, unsigned long long gid) {
  unsigned long long k = mix64(seed);
  k = mix64(k ^ iteration);
  k = mix64(k ^ (unsigned long long)(split * 4 + purpose));
  const unsigned long long r = mix64(k ^ mix64(gid));
  return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// 63-bit keys whose argsort is the iteration's random split of the ensemble (emcee shuffles the
// red/blue assignment every iteration); identical on every rank and in the numpy mirror.
__global__ void split_keys_kernel(unsigned long long seed, unsigned long long iteration, long long n,
                                  long long* __restrict__ keys) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long k = mix64(seed ^ 0x5851F42D4C957F2Dull);
  k = mix64(k ^ iteration);
  keys[i] = (long long)(mix64(k ^ mix64((unsigned long long)i)) >> 1);
}

// The iteration's red/blue split for small ensembles, in one CTA: hashed keys (as split_keys_kernel),
// bitonic sort of (key, index) pairs in shared memory -- the order torch.argsort(keys, stable=True) gives --
// and the two halves perm[0::2], perm[1::2] written out.  randomize == 0: the fixed split (evens, odds).
constexpr int SPLIT_SMALL_MAX = 2048;
__global__ void __launch_bounds__(1024)
split_perm_small_kernel(unsigned long long seed, unsigned long long iteration, int n, int npow, int randomize,
                        long long* __restrict__ h0, long long* __restrict__ h1) {
  __shared__ unsigned long long sk[SPLIT_SMALL_MAX];
  __shared__ int si[SPLIT_SMALL_MAX];
  const int tid = threadIdx.x;
  unsigned long long kbase = mix64(seed ^ 0x5851F42D4C957F2Dull);
  kbase = mix64(kbase ^ iteration);
  for (int i = tid; i < npow; i += 1024) {
    unsigned long long key = ~0ull;
    if (i < n) key = randomize ? (mix64(kbase ^ mix64((unsigned long long)i)) >> 1) : (unsigned long long)i;
    sk[i] = key; si[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= npow; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow; i += 1024) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const unsigned long long ka = sk[i], kb = sk[ixj];
          const int ia = si[i], ib = si[ixj];
          const bool a_gt_b = (ka > kb) || (ka == kb && ia > ib);
          if (a_gt_b == up) { sk[i] = kb; sk[ixj] = ka; si[i] = ib; si[ixj] = ia; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += 1024) ((i & 1) ? h1 : h0)[i >> 1] = (long long)si[i];
}

// state [nw][ndim+1] = positions | log-prob (replicated on every rank).  s_idx [n_s] are the state
// rows this rank updates (its shard of the active half), c_idx [n_c] the complementary half.
__global__ void stretch_propose_kernel(const double* __restrict__ state, const long long* __restrict__ s_idx,
                                       long long n_s, const long long* __restrict__ c_idx, long long n_c, int ndim,
                                       double a, unsigned long long seed, unsigned long long iteration, int split,
                                       long long gid0, double* __restrict__ q, double* __restrict__ logfac) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= n_s) return;
  const unsigned long long gid = (unsigned long long)(gid0 + s);
  const double u = counter_uniform(seed, iteration, split, 0, gid);
  const double root = (a - 1.0) * u + 1.0;
  const double z = root * root / a;
  long long j = (long long)(counter_uniform(seed, iteration, split, 1, gid) * (double)n_c);
  if (j >= n_c) j = n_c - 1;
  const double* x = state + s_idx[s] * (ndim + 1);
  const double* c = state + c_idx[j] * (ndim + 1);
  for (int k = 0; k < ndim; ++k) q[s * ndim + k] = c[k] - (c[k] - x[k]) * z;
  logfac[s] = (double)(ndim - 1) * log(z);
}

// packed [n_s][ndim+2] = new position | new log-prob | accepted (0/1): what the all-gather carries.
__global__ void stretch_accept_kernel(const double* __restrict__ state, const long long* __restrict__ s_idx,
                                      const double* __restrict__ q, const double* __restrict__ lnp_q,
                                      const double* __restrict__ logfac, long long n_s, int ndim,
                                      unsigned long long seed, unsigned long long iteration, int split, long long gid0,
                                      double* __restrict__ packed) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= n_s) return;
  const double* x = state + s_idx[s] * (ndim + 1);
  const double u = counter_uniform(seed, iteration, split, 2, (unsigned long long)(gid0 + s));
  const double lnpdiff = logfac[s] + lnp_q[s] - x[ndim];
  const bool acc = log(u) < lnpdiff;                                 // -inf proposals never pass; NaN never passes
  double* out = packed + s * (ndim + 2);
  for (int k = 0; k < ndim; ++k) out[k] = acc ? q[s * ndim + k] : x[k];
  out[ndim] = acc ? lnp_q[s] : x[ndim];
  out[ndim + 1] = acc ? 1.0 : 0.0;
}

// Single-rank form of accept + scatter: the active half S and its complement C are disjoint, every thread touches
// only its own row of the state, so the accepted proposal is written in place (no packed buffer, no collective).
__global__ void stretch_accept_inplace_kernel(double* __restrict__ state, int* __restrict__ naccepted,
                                              const long long* __restrict__ s_idx, const double* __restrict__ q,
                                              const double* __restrict__ lnp_q, const double* __restrict__ logfac,
                                              long long n_s, int ndim, unsigned long long seed, unsigned long long iteration,
                                              int split, long long gid0) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= n_s) return;
  double* x = state + s_idx[s] * (ndim + 1);
  const double u = counter_uniform(seed, iteration, split, 2, (unsigned long long)(gid0 + s));
  const double lnpdiff = logfac[s] + lnp_q[s] - x[ndim];
  if (log(u) < lnpdiff) {                                            // -inf proposals never pass; NaN never passes
    for (int k = 0; k < ndim; ++k) x[k] = q[s * ndim + k];
    x[ndim] = lnp_q[s];
    if (naccepted) naccepted[s_idx[s]] += 1;
  }
}

// state[idx[k]] = packed[k][0..ndim], naccepted[idx[k]] += packed[k][ndim+1]  (idx: the whole active half)
__global__ void stretch_scatter_kernel(double* __restrict__ state, int* __restrict__ naccepted,
                                       const long long* __restrict__ idx, const double* __restrict__ packed,
                                       long long n, int ndim) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= n) return;
  double* dst = state + idx[k] * (ndim + 1);
  const double* src = packed + k * (ndim + 2);
  for (int d = 0; d <= ndim; ++d) dst[d] = src[d];
  if (naccepted && src[ndim + 1] != 0.0) naccepted[idx[k]] += 1;
}

}  // namespace

// ==========================================================================================
// Host side: handle, uploads, launches
// ==========================================================================================
struct ms_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  DevTables T{};
  std::vector<void*> allocs;
  int corners = 4;               // rows per star
  int variant = 0;               // 0: (NC,NP)=(2,6), 1: (4,12)
  int threads = 256;             // CTA width of the streaming kernel
  int cluster = 0;               // CTAs per walker: 0 = automatic (small batches), 1 = never, 2/4/8 = forced
  int num_sms = 148;
  size_t smem_bytes = 0;
  // per-call buffers
  WalkerPlan* plans = nullptr; long long plans_cap = 0;
  // canonical cell order (ms_stretch_evaluate with a group shard: every rank must form the same groups)
  unsigned* sort_keys = nullptr; unsigned* sort_keys_out = nullptr; int* sort_vals = nullptr; void* sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0; long long sort_cap = 0;
  int* order = nullptr; unsigned* cell_hist = nullptr; bool use_order = false;
  int pairing = MS_PAIRING_DEFAULT;  // 0: one walker per CTA, 1: two cell-mates per CTA for large batches, -1: automatic
  bool pair_now = false;         // decided per call (evaluate_device)
  bool fast = false;             // bracket-only evaluation (ms_set_fast)
  int cell_order = -1;           // launch order: -1 automatic (cell-sorted when the grid exceeds L2), 0 natural, 1 cell-sorted
  int max_optin_smem = 0;
  bool nospec = false;           // mft6_nospec.py flavour of the posterior
  double* d_opt_flux = nullptr; double* d_opt_median = nullptr; int opt_sets = 0;   // fit_spec's renormalised spectra
  size_t grid_bytes = 0, l2_bytes = 0;
  long long grid_sig[4] = {0, 0, 0, 0};   // rows, row length, region start, region length
  double* d_theta = nullptr; double* d_lnp = nullptr; long long io_cap = 0;
  double* h_theta = nullptr; double* h_lnp = nullptr; long long pin_cap = 0;
  long long launches = 0;
  // ms_stretch_iteration scratch (single-rank fused iteration)
  long long* it_half[2] = {nullptr, nullptr}; double* it_q = nullptr; double* it_logfac = nullptr; double* it_lnp = nullptr;
  double* it_packed = nullptr; long long it_cap = 0;
  bool model_global = false; double* model_scratch = nullptr; size_t model_scratch_cap = 0;   // observed spectra beyond shared memory
  // group kernel (group_kernel.cuh): tables, per-call group list, model-vector scratch slots
  grp::GroupTables GT{};
  CUtensorMap grid_map{};         // (pixel, Teff, log g) view of the device grid for the group kernel's row boxes
  bool group_ok = false;         // the tables fit the kernel's shared-memory ring (weight slots, records per stage)
  int group_mode = -1;           // -1 automatic (large posterior batches), 0 never, 1 whenever possible
  int shard_rank = 0, shard_world = 1;   // ms_set_group_shard: this handle evaluates every shard_world-th group ...
  bool shard_active = false;             // ... of the batches of ms_stretch_evaluate (set for the duration of that call)
  int group_variant = 0;         // 0 by the number of groups, 1 always one CTA per SM, 2 always two (MCMCSPEC_GROUP_VARIANT, tests)
  bool group_now = false;        // decided per call
  bool group_probe = false;      // group kernel possible but the groups are too thin: per-walker kernel, group count still refreshed
  bool group_thin = false;       // sticky: the last counts said fewer than MS_GROUP_FILL_LO walkers per group
  long long group_min_walkers = 1024;
  grp::GroupDesc* groups = nullptr; long long groups_cap = 0;
  unsigned* cell_start = nullptr; unsigned* group_off = nullptr; int* n_groups = nullptr;
  double* group_scratch = nullptr; unsigned long long* group_list = nullptr; int* group_counter = nullptr; int group_ctas = 0;
  unsigned long long* short_circuited = nullptr;   // device counter: evaluations cut short by the prior box / tables
  int* h_n_groups = nullptr;     // pinned: group count of a recent call (picks the kernel variant of the next one)
  unsigned char* scratch = nullptr; size_t scratch_cap = 0;   // per-call inputs of ms_optimizer_chisq (grown on demand)
  cudaStream_t own_stream = nullptr;
  // optional device-side timing of the streaming kernel
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;      // start/stop pairs, reused
  size_t ev_used = 0;
  long long timed_walkers = 0;
  std::string err;
};

namespace {

std::string g_create_error;

// Device temporaries of one call: freed on every exit path.
struct DevMem {
  void* p = nullptr;
  DevMem() = default;
  DevMem(const DevMem&) = delete;
  DevMem& operator=(const DevMem&) = delete;
  ~DevMem() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename Tp> Tp* as() const { return static_cast<Tp*>(p); }
};

// NVTX ranges around the entry points (header-only NVTX3; a no-op unless a profiler is attached), replacing
// the reference's dead time.time() markers (mft6.py:397, :548, :1148, :1151, :3511-3514).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      set_err(h, std::string(#call) + ": " + cudaGetErrorString(e__));                \
      return MS_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

void set_err(ms_handle* h, const std::string& m) {
  if (h) h->err = m; else g_create_error = m;
}

template <typename Tp>
int upload(ms_handle* h, const Tp* host, size_t n, const Tp** dev) {
  void* p = nullptr;
  const size_t bytes = (n ? n : 1) * sizeof(Tp);
  CK(cudaMalloc(&p, bytes));
  h->allocs.push_back(p);
  if (n) CK(cudaMemcpy(p, host, n * sizeof(Tp), cudaMemcpyHostToDevice));
  *dev = static_cast<const Tp*>(p);
  return MS_OK;
}

long long gcd_ll(long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; }

template <int NR, int NC, int NP, int NT, int MODE, int NS>
int launch_stream_mode(ms_handle* h, long long nw, double* d_lnp, const PartsOut& parts, const OptIn& opt) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timing) {
    if (h->ev_used + 2 > h->ev_pool.size()) {
      for (int k = 0; k < 2; ++k) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->ev_pool.push_back(e); }
    }
    e0 = h->ev_pool[h->ev_used]; e1 = h->ev_pool[h->ev_used + 1];
    h->ev_used += 2; h->timed_walkers += nw;
    CK(cudaEventRecord(e0, h->stream));
  }
  const int* order = h->use_order ? h->order : nullptr;
  constexpr bool CAN_PAIR = (NT == MS_NT_SMALL) && (MODE == MODE_POSTERIOR) && (NS == 2);
  bool launched = false;
  if constexpr (CAN_PAIR) {
    if (h->pair_now) {
      // two cell-mates per CTA (see the kernel's header comment); `order` is the cell-sorted launch order
      auto kern2 = lnprob_stream_kernel<NR, NC, NP, NT, MODE, NS, 2, 0>;
      CK(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * h->smem_bytes)));
      const long long n_pairs = (nw + 1) / 2;
      long long stride = (long long)((double)n_pairs * 0.381966) | 1ll;
      while (gcd_ll(stride, n_pairs) != 1) stride += 2;
      kern2<<<(unsigned)n_pairs, 2 * NT, 2 * h->smem_bytes, h->stream>>>(h->T, h->plans, order, d_lnp, parts, opt, 1, nw,
                                                                          stride % n_pairs, (unsigned)h->smem_bytes, nullptr);
      launched = true;
    }
  }
  if constexpr (MODE != MODE_OPTIMIZER) {
    if (!launched && h->fast) {
      // bracket-only evaluation: one CTA per walker, no clusters (the work per walker is a few microseconds)
      auto kernf = lnprob_stream_kernel<NR, NC, NP, NT, MODE, NS, 1, 1>;
      CK(cudaFuncSetAttribute(kernf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
      if (h->model_global) {
        const size_t need = (size_t)nw * (size_t)h->T.n_data;
        if (need > h->model_scratch_cap) {
          if (need * 8 > (size_t)64 << 30) { set_err(h, "ms_lnprob: batch x observed pixels exceeds the 64 GiB model scratch; evaluate in smaller batches"); return MS_ERR_UNSUPPORTED; }
          CK(cudaStreamSynchronize(h->stream));
          if (h->model_scratch) cudaFree(h->model_scratch);
          h->model_scratch = nullptr; h->model_scratch_cap = 0;
          CK(cudaMalloc((void**)&h->model_scratch, need * 8));
          h->model_scratch_cap = need;
        }
      }
      kernf<<<(unsigned)nw, NT, h->smem_bytes, h->stream>>>(h->T, h->plans, order, d_lnp, parts, opt, 1, nw, 1ll,
                                                            (unsigned)h->smem_bytes, h->model_global ? h->model_scratch : nullptr);
      launched = true;
    }
  }
  if (!launched) {
    auto kern = lnprob_stream_kernel<NR, NC, NP, NT, MODE, NS, 1, 0>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    // cluster size: enough CTAs for two per SM, at most 8 per walker (portable limit)
    int csize = h->cluster;
    if (csize <= 0) {
      csize = 1;
      while (csize < 8 && (long long)nw * csize * 2 <= 2ll * h->num_sms) csize *= 2;
    }
    if (h->model_global) {
      // the model vector of every launch slot lives in a global scratch row (grown on demand)
      csize = 1;
      const size_t need = (size_t)nw * (size_t)h->T.n_data;
      if (need > h->model_scratch_cap) {
        if (need * 8 > (size_t)64 << 30) { set_err(h, "ms_lnprob: batch x observed pixels exceeds the 64 GiB model scratch; evaluate in smaller batches"); return MS_ERR_UNSUPPORTED; }
        CK(cudaStreamSynchronize(h->stream));
        if (h->model_scratch) cudaFree(h->model_scratch);
        h->model_scratch = nullptr; h->model_scratch_cap = 0;
        CK(cudaMalloc((void**)&h->model_scratch, need * 8));
        h->model_scratch_cap = need;
      }
    }
    if (csize == 1) {
      kern<<<(unsigned)nw, NT, h->smem_bytes, h->stream>>>(h->T, h->plans, order, d_lnp, parts, opt, 1, nw, 1ll,
                                                           (unsigned)h->smem_bytes, h->model_global ? h->model_scratch : nullptr);
    } else {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(nw * csize)); cfg.blockDim = dim3(NT);
      cfg.dynamicSmemBytes = h->smem_bytes; cfg.stream = h->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, kern, h->T, (const WalkerPlan*)h->plans, order, d_lnp, parts, opt, csize, (long long)nw,
                            1ll, (unsigned)h->smem_bytes, (double*)nullptr));
    }
  }
  CK(cudaGetLastError());
  if (h->timing) CK(cudaEventRecord(e1, h->stream));
  h->launches += 1;
  return MS_OK;
}

template <int NR, int NC, int NP, int NT>
int launch_stream_threads(ms_handle* h, long long nw, double* d_lnp, const PartsOut& parts, int mode, const OptIn& opt) {
  if (h->T.nspec == 3) {
    // triples: Teff x log g grids, posterior flavours only
    if constexpr (NR == 4) {
      if (mode == MODE_NOSPEC) return launch_stream_mode<NR, NC, NP, NT, MODE_NOSPEC, 3>(h, nw, d_lnp, parts, opt);
      if (mode == MODE_POSTERIOR) return launch_stream_mode<NR, NC, NP, NT, MODE_POSTERIOR, 3>(h, nw, d_lnp, parts, opt);
      if (mode == MODE_OPTIMIZER) return launch_stream_mode<NR, NC, NP, NT, MODE_OPTIMIZER, 3>(h, nw, d_lnp, parts, opt);
    }
    set_err(h, "triples are built for (Teff, log g) grids only");
    return MS_ERR_UNSUPPORTED;
  }
  if (mode == MODE_OPTIMIZER) return launch_stream_mode<NR, NC, NP, NT, MODE_OPTIMIZER, 2>(h, nw, d_lnp, parts, opt);
  if (mode == MODE_NOSPEC) return launch_stream_mode<NR, NC, NP, NT, MODE_NOSPEC, 2>(h, nw, d_lnp, parts, opt);
  return launch_stream_mode<NR, NC, NP, NT, MODE_POSTERIOR, 2>(h, nw, d_lnp, parts, opt);
}

template <int NR, int NC, int NP>
int launch_stream_variant(ms_handle* h, long long nw, double* d_lnp, const PartsOut& parts, int mode, const OptIn& opt) {
  if (h->threads == MS_NT_LARGE) return launch_stream_threads<NR, NC, NP, MS_NT_LARGE>(h, nw, d_lnp, parts, mode, opt);
  return launch_stream_threads<NR, NC, NP, MS_NT_SMALL>(h, nw, d_lnp, parts, mode, opt);
}

size_t stream_smem_bytes(int n_data, int nc, int np, int nwarp, int nstar) {
  const int kred = nstar * nc + np;
  size_t b = (size_t)nwarp * SEG * 8 + (size_t)nwarp * (kred > 4 ? kred : 4) * 8 + sizeof(SelectScratch);
  b += (size_t)n_data * 8;
  return (b + 15) & ~(size_t)15;
}

}  // namespace

#if MS_HAS_NR4_WIDE
int ms_tu_launch_nr4_wide(ms_handle* h, long long nw, double* d_lnp, const msd::PartsOut& parts, int mode, const msd::OptIn& opt) {
  return launch_stream_variant<4, MS_MAX_CONTRAST, MS_MAX_PHOT>(h, nw, d_lnp, parts, mode, opt);
}
#endif
#if MS_HAS_NR8
int ms_tu_launch_nr8(ms_handle* h, int wide, long long nw, double* d_lnp, const msd::PartsOut& parts, int mode, const msd::OptIn& opt) {
  if (wide) return launch_stream_variant<8, MS_MAX_CONTRAST, MS_MAX_PHOT>(h, nw, d_lnp, parts, mode, opt);
  return launch_stream_variant<8, 2, 6>(h, nw, d_lnp, parts, mode, opt);
}
#endif

#if MS_HAS_GROUP
int ms_tu_launch_group(ms_handle* h, long long nw, double* d_lnp, const msd::PartsOut& parts) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timing) {
    if (h->ev_used + 2 > h->ev_pool.size()) {
      for (int k = 0; k < 2; ++k) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->ev_pool.push_back(e); }
    }
    e0 = h->ev_pool[h->ev_used]; e1 = h->ev_pool[h->ev_used + 1];
    h->ev_used += 2; h->timed_walkers += h->shard_active ? nw / h->shard_world : nw;      // walkers this launch evaluates
    CK(cudaEventRecord(e0, h->stream));
  }
  // Few groups (at most one per SM): one CTA per SM with sixteen walker warps per group shortens the time a group takes;
  // otherwise two CTAs per SM with eight walker warps each.  The number of groups depends on how the walkers spread over
  // the grid cells and is only known on the device: the count of a recent call of the same batch size (copied back
  // asynchronously, never waited for) stands in for it -- successive half steps of a chain look alike.
  const long long min_groups = (nw + grp::GMAX - 1) / grp::GMAX;
  const long long max_groups = min_groups + std::min<long long>(h->T.n_cells, nw);
  const int seen = *reinterpret_cast<volatile int*>(h->h_n_groups);
  const long long seen_nw = *reinterpret_cast<volatile int*>(h->h_n_groups + 1);
  const long long expect_all = (seen > 0 && seen_nw == nw) ? (long long)seen : min_groups + min_groups / 3;
  const int sworld = h->shard_active ? h->shard_world : 1, srank = h->shard_active ? h->shard_rank : 0;
  const long long expect_groups = (expect_all + sworld - 1) / sworld;      // this rank's share
  grp::GroupTables GTs = h->GT;
  GTs.shard_rank = srank; GTs.shard_world = sworld;
  const bool wide = h->group_variant == 1 || (h->group_variant == 0 && expect_groups <= h->num_sms);
  auto launch = [&](auto kern, int nww, size_t ring, long long slots) -> int {
    const int smem = (int)(ring + (size_t)nww * grp::HBINS * 32 * 2);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const unsigned ctas = (unsigned)std::min<long long>(slots, max_groups);
    kern<<<ctas, nww * 32, smem, h->stream>>>(h->T, GTs, h->grid_map, h->plans, h->order, h->groups, h->n_groups, h->group_counter,
                                             d_lnp, parts, h->group_scratch, h->group_list);
    return MS_OK;
  };
  int rc;
  if (h->corners == 8) {
    const size_t ring = (size_t)grp::StageLayout<8>::ring;
    rc = wide ? launch(grp::lnprob_group_kernel<8, MS_GROUP_NWW_WIDE, 1>, MS_GROUP_NWW_WIDE, ring, h->num_sms)
              : launch(grp::lnprob_group_kernel<8, MS_GROUP_NWW, 2>, MS_GROUP_NWW, ring, h->group_ctas);
  } else {
    const size_t ring = (size_t)grp::StageLayout<4>::ring;
    rc = wide ? launch(grp::lnprob_group_kernel<4, MS_GROUP_NWW_WIDE, 1>, MS_GROUP_NWW_WIDE, ring, h->num_sms)
              : launch(grp::lnprob_group_kernel<4, MS_GROUP_NWW, 2>, MS_GROUP_NWW, ring, h->group_ctas);
  }
  if (rc) return rc;
  CK(cudaGetLastError());
  if (h->timing) CK(cudaEventRecord(e1, h->stream));
  h->h_n_groups[1] = (int)std::min<long long>(nw, 0x7fffffff);
  CK(cudaMemcpyAsync(h->h_n_groups, h->n_groups, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  h->launches += 1;
  return MS_OK;
}
#endif

#if MS_HAS_MAIN
namespace {

int ensure_plans(ms_handle* h, long long nw) {
  if (nw <= h->plans_cap) return MS_OK;
  if (h->plans) cudaFree(h->plans);
  if (h->order) cudaFree(h->order);
  if (h->sort_keys) cudaFree(h->sort_keys);
  if (h->sort_keys_out) cudaFree(h->sort_keys_out);
  if (h->sort_vals) cudaFree(h->sort_vals);
  if (h->sort_tmp) cudaFree(h->sort_tmp);
  h->plans = nullptr; h->order = nullptr; h->plans_cap = 0;
  long long cap = nw + nw / 4 + 64;
  CK(cudaMalloc((void**)&h->plans, (size_t)cap * sizeof(WalkerPlan)));
  CK(cudaMalloc((void**)&h->order, (size_t)cap * sizeof(int)));
  h->plans_cap = cap;
  return MS_OK;
}

int evaluate_device(ms_handle* h, const double* d_theta, long long nw, double* d_lnp, const PartsOut& parts,
                    int mode = -1, const OptIn& opt = OptIn{nullptr, nullptr, nullptr, nullptr, nullptr},
                    const ProposeIn& prop = ProposeIn{nullptr, nullptr, nullptr, 0, 0.0, 0ull, 0ull, 0, 0, nullptr, nullptr}) {
  if (mode < 0) mode = h->nospec ? MODE_NOSPEC : MODE_POSTERIOR;
  if (nw == 0) return MS_OK;
  if (mode != MODE_OPTIMIZER && h->T.nspec == 3 && h->T.dist_fit && !h->T.rad_prior) {
    // mft6.py:1381-1393: the 8-parameter dist_fit branch only returns inside `if rad_prior == True`; without it
    // logprior returns None and logposterior's np.isfinite(None) raises.  There is no value to reproduce.
    set_err(h, "logprior is undefined for a triple with dist_fit and without rad_prior (mft6.py:1381-1393 returns None)");
    return MS_ERR_UNSUPPORTED;
  }
  int rc = ensure_plans(h, nw);
  if (rc) return rc;
  const int pb = 128;
  const unsigned pg = (unsigned)((nw + pb - 1) / pb);
  // Grouping walkers by grid cell pays when the grid does not fit L2 (rows are then fetched from HBM once
  // per group instead of once per walker).  With an L2-resident grid it only concentrates the concurrent
  // CTAs on the same L2 slices (measured: C4 -4 %, C1 -8 %), so the natural order is kept there.
  h->use_order = nw >= 1024 && h->T.n_cells > 1 && (h->cell_order < 0 ? h->grid_bytes > h->l2_bytes : h->cell_order > 0);
  // Two walkers per CTA pay when (a) the batch still fills the machine with half as many CTAs, (b) both
  // walkers' shared memory fits one SM and (c) cell-mates exist to be paired: the launch order is then
  // always the cell-sorted one.
  // Measured (profiles/): pairing halves the L2->SM traffic (L1 hit rate 53 %), which pays on stream-dominated
  // shapes (C1/C2: +5 %) and costs the pair barriers where the observed-pixel work dominates (C4: -5 %).
  const bool want_pair = h->pairing > 0 || (h->pairing < 0 && (long long)h->T.n_reg >= 16ll * h->T.n_data);
  h->pair_now = want_pair && !h->fast && !h->model_global && h->cluster <= 1 && mode == MODE_POSTERIOR && h->T.nspec == 2 &&
                h->threads == MS_NT_SMALL && h->T.n_cells > 1 && nw >= 8ll * h->num_sms &&
                2 * h->smem_bytes + 2048 <= (size_t)h->max_optin_smem;
  if (h->pair_now) h->use_order = true;
  // Large posterior batches of binaries on a (Teff, log g) grid: groups of cell-mates share one stream of their
  // rows (group_kernel.cuh).  Needs the cell-sorted order and the group table.
  h->group_now = h->group_ok && h->group_mode != 0 && mode == MODE_POSTERIOR && !h->fast && h->cluster <= 1 &&
                 (h->group_mode > 0 || nw >= h->group_min_walkers);
  h->group_probe = false;
  if (h->group_now && h->group_mode < 0 && !h->shard_active) {
    // A group costs the same whether 32 lanes or 3 are busy: when the batch spreads so thinly over the grid cells that a
    // group holds only a few walkers (a chain far from convergence on a fine grid, C5's proposals), the per-walker kernel
    // is faster.  The group count of the last call of this batch size decides (copied back asynchronously, never waited
    // for), with hysteresis; the count keeps being refreshed on either path.  ms_set_group(1) forces the group kernel.
    const int seen = *reinterpret_cast<volatile int*>(h->h_n_groups);
    const long long seen_nw = *reinterpret_cast<volatile int*>(h->h_n_groups + 1);
    if (seen > 0 && seen_nw == nw) {
      const double fill = (double)nw / (double)seen;
      if (fill < MS_GROUP_FILL_LO) h->group_thin = true;
      else if (fill > MS_GROUP_FILL_HI) h->group_thin = false;
    }
    if (h->group_thin) { h->group_now = false; h->group_probe = true; }
  }
  if (h->group_now || h->group_probe) { h->use_order = true; h->pair_now = false; }
  if (h->use_order) CK(cudaMemsetAsync(h->cell_hist, 0, (size_t)h->T.n_cells * sizeof(unsigned), h->stream));
  const size_t prep_smem = (size_t)(3 * h->T.n_iso + h->T.n_teff + h->T.n_logg + h->T.n_feh + h->T.n_dm +
                                    h->T.n_dm * h->T.n_dust_samples) * sizeof(double);
  if (prep_smem > 48 * 1024) CK(cudaFuncSetAttribute(walker_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
  walker_prep_kernel<<<pg, pb, prep_smem, h->stream>>>(h->T, d_theta, nw, h->plans, h->use_order ? h->cell_hist : nullptr, mode, opt, prop,
                                                       h->short_circuited);
  CK(cudaGetLastError());
  h->launches += 1;
  if (h->group_now || h->group_probe) {
    const long long max_groups = (nw + grp::GMAX - 1) / grp::GMAX + std::min<long long>(h->T.n_cells, nw);
    if (max_groups > h->groups_cap) {
      CK(cudaStreamSynchronize(h->stream));
      if (h->groups) cudaFree(h->groups);
      h->groups = nullptr; h->groups_cap = 0;
      CK(cudaMalloc((void**)&h->groups, (size_t)(max_groups + max_groups / 4) * sizeof(grp::GroupDesc)));
      h->groups_cap = max_groups + max_groups / 4;
    }
    grp::group_scan_kernel<<<1, 1024, 0, h->stream>>>(h->cell_hist, h->T.n_cells, h->cell_start, h->group_off, h->n_groups,
                                                      h->group_counter);
    h->launches += 1;
    if (h->group_mode < 0 && !h->shard_active &&
        !(*reinterpret_cast<volatile int*>(h->h_n_groups) > 0 && *reinterpret_cast<volatile int*>(h->h_n_groups + 1) == (int)std::min<long long>(nw, 0x7fffffff))) {
      // first batch of this size: no earlier group count to go by -- wait for this one (once) before choosing the kernel
      int ng = 0;
      CK(cudaMemcpyAsync(&ng, h->n_groups, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      h->group_thin = ng > 0 && (double)nw / (double)ng < MS_GROUP_FILL_LO;
      h->group_now = !h->group_thin; h->group_probe = h->group_thin;
    }
  }
  if (h->group_now) {
    if (h->shard_active) {
      // the ranks that share this half must form the same groups: walkers in cell order, by index within a cell (a
      // stable radix sort on the cell number) instead of the order the scatter's atomics happen to give
      if (nw > h->sort_cap) {
        if (h->sort_keys) cudaFree(h->sort_keys);
        if (h->sort_keys_out) cudaFree(h->sort_keys_out);
        if (h->sort_vals) cudaFree(h->sort_vals);
        if (h->sort_tmp) cudaFree(h->sort_tmp);
        h->sort_keys = h->sort_keys_out = nullptr; h->sort_vals = nullptr; h->sort_tmp = nullptr; h->sort_cap = 0;
        const long long cap = nw + nw / 4 + 64;
        CK(cudaMalloc((void**)&h->sort_keys, (size_t)cap * sizeof(unsigned)));
        CK(cudaMalloc((void**)&h->sort_keys_out, (size_t)cap * sizeof(unsigned)));
        CK(cudaMalloc((void**)&h->sort_vals, (size_t)cap * sizeof(int)));
        size_t bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, h->sort_keys, h->sort_keys_out, h->sort_vals, h->order, (int)cap, 0, 32, h->stream));
        CK(cudaMalloc(&h->sort_tmp, bytes));
        h->sort_tmp_bytes = bytes; h->sort_cap = cap;
      }
      int bits = 1;
      while ((1ll << bits) < (long long)h->T.n_cells + 1 && bits < 32) ++bits;
      cell_keys_kernel<<<pg, pb, 0, h->stream>>>(h->plans, nw, h->sort_keys, h->sort_vals);
      size_t bytes = h->sort_tmp_bytes;
      CK(cub::DeviceRadixSort::SortPairs(h->sort_tmp, bytes, h->sort_keys, h->sort_keys_out, h->sort_vals, h->order, (int)nw, 0, bits, h->stream));
    } else {
      cell_scatter_kernel<<<pg, pb, 0, h->stream>>>(h->plans, nw, h->cell_hist, h->order);
    }
    grp::group_table_kernel<<<pg, pb, 0, h->stream>>>(h->plans, h->order, nw, h->cell_start, h->group_off, h->groups);
    CK(cudaGetLastError());
    h->launches += 2;
    return ms_tu_launch_group(h, nw, d_lnp, parts);
  }
  if (h->group_probe) {
    // per-walker kernel in cell order; the scan above also counted the groups this batch would have made
    cell_scatter_kernel<<<pg, pb, 0, h->stream>>>(h->plans, nw, h->cell_hist, h->order);
    CK(cudaGetLastError());
    h->h_n_groups[1] = (int)std::min<long long>(nw, 0x7fffffff);
    CK(cudaMemcpyAsync(h->h_n_groups, h->n_groups, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    h->launches += 1;
  } else if (h->use_order) {
    cell_scan_kernel<<<1, 1024, 0, h->stream>>>(h->cell_hist, h->T.n_cells);
    cell_scatter_kernel<<<pg, pb, 0, h->stream>>>(h->plans, nw, h->cell_hist, h->order);
    CK(cudaGetLastError());
    h->launches += 2;
  }
  if (h->corners == 4) {
    if (h->variant == 0) return launch_stream_variant<4, 2, 6>(h, nw, d_lnp, parts, mode, opt);
    return ms_tu_launch_nr4_wide(h, nw, d_lnp, parts, mode, opt);
  }
  return ms_tu_launch_nr8(h, h->variant != 0, nw, d_lnp, parts, mode, opt);
}

int ensure_io(ms_handle* h, long long nw) {
  if (nw > h->io_cap) {
    if (h->d_theta) cudaFree(h->d_theta);
    if (h->d_lnp) cudaFree(h->d_lnp);
    h->d_theta = h->d_lnp = nullptr; h->io_cap = 0;
    long long cap = nw + nw / 4 + 64;
    CK(cudaMalloc((void**)&h->d_theta, (size_t)cap * MS_MAX_NDIM * 8));
    CK(cudaMalloc((void**)&h->d_lnp, (size_t)cap * 8));
    h->io_cap = cap;
  }
  if (nw > h->pin_cap) {
    if (h->h_theta) cudaFreeHost(h->h_theta);
    if (h->h_lnp) cudaFreeHost(h->h_lnp);
    h->h_theta = h->h_lnp = nullptr; h->pin_cap = 0;
    long long cap = nw + nw / 4 + 64;
    CK(cudaMallocHost((void**)&h->h_theta, (size_t)cap * MS_MAX_NDIM * 8));
    CK(cudaMallocHost((void**)&h->h_lnp, (size_t)cap * 8));
    h->pin_cap = cap;
  }
  return MS_OK;
}

// scratch device buffer for a host-side optional output
struct HostOut {
  void* dev = nullptr; void* host = nullptr; size_t bytes = 0;
};

}  // namespace

extern "C" {

int64_t ms_config_size(void) { return (int64_t)sizeof(ms_config); }
int32_t ms_abi_version(void) { return MS_ABI_VERSION; }

const char* ms_last_error(const ms_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
void* ms_stream(const ms_handle* h) { return h ? (void*)h->stream : nullptr; }
int ms_device(const ms_handle* h) { return h ? h->device : -1; }
int64_t ms_launch_count(const ms_handle* h) { return h ? h->launches : 0; }

int ms_destroy(ms_handle* h) {
  if (!h) return MS_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (void* p : h->allocs) cudaFree(p);
  if (h->plans) cudaFree(h->plans);
  if (h->order) cudaFree(h->order);
  if (h->cell_hist) cudaFree(h->cell_hist);
  if (h->d_opt_flux) cudaFree(h->d_opt_flux);
  if (h->d_opt_median) cudaFree(h->d_opt_median);
  if (h->d_theta) cudaFree(h->d_theta);
  if (h->d_lnp) cudaFree(h->d_lnp);
  if (h->h_theta) cudaFreeHost(h->h_theta);
  if (h->h_lnp) cudaFreeHost(h->h_lnp);
  if (h->scratch) cudaFree(h->scratch);
  if (h->model_scratch) cudaFree(h->model_scratch);
  if (h->groups) cudaFree(h->groups);
  if (h->cell_start) cudaFree(h->cell_start);
  if (h->group_off) cudaFree(h->group_off);
  if (h->n_groups) cudaFree(h->n_groups);
  if (h->group_scratch) cudaFree(h->group_scratch);
  if (h->group_counter) cudaFree(h->group_counter);
  if (h->group_list) cudaFree(h->group_list);
  if (h->h_n_groups) cudaFreeHost(h->h_n_groups);
  if (h->short_circuited) cudaFree(h->short_circuited);
  if (h->it_half[0]) cudaFree(h->it_half[0]);
  if (h->it_half[1]) cudaFree(h->it_half[1]);
  if (h->it_q) cudaFree(h->it_q);
  if (h->it_logfac) cudaFree(h->it_logfac);
  if (h->it_lnp) cudaFree(h->it_lnp);
  if (h->it_packed) cudaFree(h->it_packed);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return MS_OK;
}

int ms_create(const ms_config* cfg, ms_handle** out) { return ms_create_shared(cfg, nullptr, out); }

int ms_create_shared(const ms_config* cfg, const ms_handle* grid_owner, ms_handle** out) {
  NvtxRange nvtx_range("ms_create_shared");
  if (!cfg || !out) { set_err(nullptr, "ms_create: null argument"); return MS_ERR_INVALID; }
  *out = nullptr;
  if (cfg->abi_version != MS_ABI_VERSION) { set_err(nullptr, "ms_create: ABI version mismatch"); return MS_ERR_INVALID; }
  if (cfg->nspec != 2 && cfg->nspec != 3) { set_err(nullptr, "ms_create: nspec must be 2 (binary) or 3 (triple)"); return MS_ERR_UNSUPPORTED; }
  if (cfg->feh_free && (cfg->n_feh < 2 || cfg->nspec != 2)) {
    set_err(nullptr, "ms_create: a free [Fe/H] needs a 3-D grid and is built for binaries"); return MS_ERR_UNSUPPORTED;
  }
  if (cfg->nspec == 3 && cfg->n_feh > 1) { set_err(nullptr, "ms_create: triples are built for (Teff, log g) grids only"); return MS_ERR_UNSUPPORTED; }
  if (cfg->n_contrast < 0 || cfg->n_contrast > MS_MAX_CONTRAST || cfg->n_phot < 0 || cfg->n_phot > MS_MAX_PHOT) {
    set_err(nullptr, "ms_create: too many filters"); return MS_ERR_UNSUPPORTED;
  }
  if (cfg->n_teff < 1 || cfg->n_logg < 1 || cfg->n_feh < 1 || cfg->n_reg < 2 || cfg->n_data < 1 ||
      cfg->n_iso < 2 || cfg->n_dm < 1 || cfg->n_dust_samples < 1 || cfg->n_dust_samples > 16 ||
      cfg->n_tiles != (cfg->n_reg + TILE - 1) / TILE || cfg->reg_lo < 0 || cfg->reg_lo + cfg->n_reg > cfg->row_len) {
    set_err(nullptr, "ms_create: inconsistent sizes"); return MS_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err(nullptr, "ms_create: no CUDA device"); return MS_ERR_NO_DEVICE; }
  if (cfg->device < 0 || cfg->device >= ndev) { set_err(nullptr, "ms_create: bad device ordinal"); return MS_ERR_INVALID; }

  ms_handle* h = new (std::nothrow) ms_handle();
  if (!h) { set_err(nullptr, "ms_create: out of host memory"); return MS_ERR_INVALID; }
  h->device = cfg->device;
  auto fail = [&](int rc) { g_create_error = h->err; ms_destroy(h); return rc; };
#define CKC(expr) do { int rc__ = (expr); if (rc__) return fail(rc__); } while (0)
#define CKH(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e__); return fail(MS_ERR_CUDA); } } while (0)
  CKH(cudaSetDevice(h->device));
  CKH(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;

  DevTables& T = h->T;
  T.n_teff = cfg->n_teff; T.n_logg = cfg->n_logg; T.n_feh = cfg->n_feh;
  T.n_reg = cfg->n_reg; T.n_tiles = cfg->n_tiles; T.feh = cfg->feh;
  T.n_data = cfg->n_data;
  T.n_contrast = cfg->n_contrast; T.n_phot = cfg->n_phot;
  T.nspec = cfg->nspec; T.ndim = 2 * cfg->nspec + 2;
  T.feh_free = (cfg->feh_free != 0); T.ncol = T.ndim + T.feh_free;
  T.ext_flag = cfg->ext_flag; T.dist_fit = cfg->dist_fit; T.rad_prior = cfg->rad_prior;
  T.opt_median = nullptr;
  h->nospec = (cfg->flavour == 1);
  T.nospec = h->nospec ? 1 : 0;
  T.tmin = cfg->tmin; T.tmax = cfg->tmax;
  T.data_median = cfg->data_median;
  for (int k = 0; k < 9; ++k) T.basis[k] = cfg->poly_basis[k];
  for (int k = 0; k < MS_MAX_NDIM; ++k) { T.prior_mu[k] = cfg->prior_mu[k]; T.prior_sig[k] = cfg->prior_sig[k]; }
  T.n_iso = cfg->n_iso; T.n_dm = cfg->n_dm; T.n_dust_samples = cfg->n_dust_samples;
  {
    const long long per_star = (long long)cfg->n_teff * cfg->n_logg;
    const long long cells = per_star * per_star * (cfg->n_feh > 1 ? cfg->n_feh : 1) + 1;   // + 1: walkers that never reach the grid
    T.n_cells = (cells <= (1 << 20)) ? (int)cells : 1;
    CKH(cudaMalloc((void**)&h->cell_hist, (size_t)T.n_cells * sizeof(unsigned)));
  }
  h->corners = (cfg->n_feh > 1) ? 8 : 4;
  h->variant = (cfg->n_contrast <= 2 && cfg->n_phot <= 6) ? 0 : 1;

  // ---- grid: region cut of every row, padded with zeros to whole tiles
  const long long n_rows = (long long)cfg->n_feh * cfg->n_logg * cfg->n_teff;
  T.stride = (long long)cfg->n_tiles * TILE;
  h->grid_sig[0] = n_rows; h->grid_sig[1] = cfg->row_len; h->grid_sig[2] = cfg->reg_lo; h->grid_sig[3] = cfg->n_reg;
  if (grid_owner) {
    // several fits of the same model grid (config C3: independent ensembles on different observed
    // spectra): reuse the owner's device copy of the rows.  The owner must outlive this handle.
    if (grid_owner->device != h->device || std::memcmp(grid_owner->grid_sig, h->grid_sig, sizeof(h->grid_sig)) != 0) {
      h->err = "ms_create_shared: grid shape / consumed region / device differ from the owner's"; return fail(MS_ERR_INVALID);
    }
    T.grid = grid_owner->T.grid;
    h->grid_bytes = grid_owner->grid_bytes; h->l2_bytes = grid_owner->l2_bytes;
  } else {
    // 256 bytes of zeros in front of row 0: the group kernel's bulk copies start two pixels before a stage
    void* p = nullptr;
    CKH(cudaMalloc(&p, (size_t)n_rows * T.stride * 8 + 256));
    h->allocs.push_back(p);
    CKH(cudaMemset(p, 0, (size_t)n_rows * T.stride * 8 + 256));
    double* g0 = static_cast<double*>(p) + 32;
    CKH(cudaMemcpy2D(g0, (size_t)T.stride * 8, cfg->grid + cfg->reg_lo, (size_t)cfg->row_len * 8,
                     (size_t)cfg->n_reg * 8, (size_t)n_rows, cudaMemcpyHostToDevice));
    T.grid = g0;
    h->grid_bytes = (size_t)n_rows * T.stride * 8;
    int l2 = 0;
    CKH(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device));
    h->l2_bytes = (size_t)l2;
  }
  CKH(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
  CKC(upload(h, cfg->teff_nodes, cfg->n_teff, &T.teff_nodes));
  CKC(upload(h, cfg->logg_nodes, cfg->n_logg, &T.logg_nodes));
  CKC(upload(h, cfg->feh_nodes, cfg->n_feh, &T.feh_nodes));

  // ---- observed spectrum tables
  std::vector<int> chunk_start;
  for (int i = 0; i < cfg->n_data; ++i) {
    if (cfg->rebin_hi[i] < 1 || cfg->rebin_hi[i] >= cfg->n_reg || (i && cfg->rebin_hi[i] < cfg->rebin_hi[i - 1])) {
      h->err = "ms_create: rebin_hi must be ascending and inside the region"; return fail(MS_ERR_INVALID);
    }
  }
  {
    const int n = cfg->n_data;
    const double LN2 = 0.69314718055994530942;
    std::vector<double> inv_e2(n), rrec(4 * (size_t)n), drec(4 * (size_t)n);
    double dmax = 0.0;
    for (int i = 0; i < n; ++i) {
      inv_e2[i] = 1.0 / (cfg->data_err[i] * cfg->data_err[i]);                                       // mft6.py:120
      const double dext = (cfg->ext_hi[i] - cfg->ext_lo[i]) * LN2;
      dmax = std::max(dmax, std::fabs(dext));
      rrec[4 * i] = cfg->rebin_t[i]; rrec[4 * i + 1] = cfg->ext_lo[i]; rrec[4 * i + 2] = dext; rrec[4 * i + 3] = (double)cfg->rebin_hi[i];
      drec[4 * i] = cfg->data_x[i]; drec[4 * i + 1] = cfg->data_flux[i]; drec[4 * i + 2] = inv_e2[i]; drec[4 * i + 3] = 0.0;
    }
    T.dext_max = dmax;
    CKC(upload(h, inv_e2.data(), inv_e2.size(), &T.data_inv_e2));
    CKC(upload(h, rrec.data(), rrec.size(), &T.rebin_rec));
    CKC(upload(h, drec.data(), drec.size(), &T.data_rec));
    // observed pixels owned by each 128-pixel chunk (rebin_hi is ascending)
    const int n_chunks = cfg->n_tiles * CHUNKS_PER_TILE;
    std::vector<int> cds((size_t)n_chunks + 1);
    int i = 0;
    for (int c = 0; c <= n_chunks; ++c) {
      while (i < n && cfg->rebin_hi[i] < c * CHUNK) ++i;
      cds[c] = i;
      if (c < n_chunks && i < n && cfg->rebin_hi[i] == c * CHUNK) cds[c] |= (int)0x80000000;
    }
    chunk_start = cds;
  }

  // ---- filters
  std::vector<unsigned> cmask(cfg->n_tiles, 0u), pmask(cfg->n_tiles, 0u);
  const double* d_cw = nullptr; const double* d_pw = nullptr;
  CKC(upload(h, cfg->c_weights, (size_t)cfg->c_weights_len, &d_cw));
  CKC(upload(h, cfg->p_weights, (size_t)cfg->p_weights_len, &d_pw));
  for (int f = 0; f < MS_MAX_CONTRAST; ++f) { T.c_wbase[f] = d_cw; T.c_mag[f] = 0; T.c_err[f] = 1; }
  for (int f = 0; f < MS_MAX_PHOT; ++f) { T.p_wbase[f] = d_pw; T.p_norm[f] = 1; T.p_ext[f] = 0; T.p_mag[f] = 0; T.p_err[f] = 1; }
  for (int f = 0; f < cfg->n_contrast; ++f) {
    const int lo = cfg->c_tile_lo[f], hi = cfg->c_tile_hi[f];
    if (lo < 0 || hi > cfg->n_tiles || hi < lo || cfg->c_offset[f] + (long long)(hi - lo) * TILE > cfg->c_weights_len) {
      h->err = "ms_create: contrast filter tiles out of range"; return fail(MS_ERR_INVALID);
    }
    for (int t = lo; t < hi; ++t) cmask[t] |= 1u << f;
    T.c_wbase[f] = d_cw + cfg->c_offset[f] - (long long)lo * TILE;
    T.c_mag[f] = cfg->c_mag[f]; T.c_err[f] = cfg->c_err[f];
  }
  for (int f = 0; f < cfg->n_phot; ++f) {
    const int lo = cfg->p_tile_lo[f], hi = cfg->p_tile_hi[f];
    if (lo < 0 || hi > cfg->n_tiles || hi < lo || cfg->p_offset[f] + (long long)(hi - lo) * TILE > cfg->p_weights_len) {
      h->err = "ms_create: photometric filter tiles out of range"; return fail(MS_ERR_INVALID);
    }
    for (int t = lo; t < hi; ++t) pmask[t] |= 1u << f;
    T.p_wbase[f] = d_pw + cfg->p_offset[f] - (long long)lo * TILE;
    T.p_norm[f] = cfg->p_norm[f]; T.p_ext[f] = cfg->p_ext[f];
    T.p_mag[f] = cfg->p_mag[f]; T.p_err[f] = cfg->p_err[f];
  }
  {
    const int n_chunks = cfg->n_tiles * CHUNKS_PER_TILE;
    std::vector<int4> meta(n_chunks);
    for (int c = 0; c < n_chunks; ++c)
      meta[c] = make_int4((int)cmask[c / CHUNKS_PER_TILE], (int)pmask[c / CHUNKS_PER_TILE], chunk_start[c],
                          chunk_start[c + 1] & 0x7fffffff);
    CKC(upload(h, meta.data(), meta.size(), &T.chunk_meta));
  }

  // ---- bracket-only mode tables, built on the device from the (possibly shared) device grid
  {
    double2* d_br = nullptr; double* d_bs = nullptr;
    CKH(cudaMalloc((void**)&d_br, (size_t)n_rows * cfg->n_data * sizeof(double2)));
    h->allocs.push_back(d_br);
    CKH(cudaMalloc((void**)&d_bs, (size_t)n_rows * BAND_COLS * sizeof(double)));
    h->allocs.push_back(d_bs);
    CKH(cudaMemset(d_bs, 0, (size_t)n_rows * BAND_COLS * sizeof(double)));
    dim3 gb((unsigned)((cfg->n_data + 255) / 256), (unsigned)n_rows);
    bracket_gather_kernel<<<gb, 256, 0, h->stream>>>(T.grid, T.stride, T.rebin_rec, cfg->n_data, d_br);
    for (int f = 0; f < cfg->n_contrast; ++f)
      if (cfg->c_tile_hi[f] > cfg->c_tile_lo[f])
        bandsum_kernel<<<(unsigned)n_rows, 256, 0, h->stream>>>(T.grid, T.stride, T.c_wbase[f], cfg->c_tile_lo[f] * TILE,
                                                                cfg->c_tile_hi[f] * TILE, f, d_bs);
    for (int f = 0; f < cfg->n_phot; ++f)
      if (cfg->p_tile_hi[f] > cfg->p_tile_lo[f])
        bandsum_kernel<<<(unsigned)n_rows, 256, 0, h->stream>>>(T.grid, T.stride, T.p_wbase[f], cfg->p_tile_lo[f] * TILE,
                                                                cfg->p_tile_hi[f] * TILE, MS_MAX_CONTRAST + f, d_bs);
    CKH(cudaGetLastError());
    T.bracket = d_br; T.bandsum = d_bs;
  }

  // ---- group kernel tables (group_kernel.cuh): stages of KP observed pixels over the bracket table
  {
    using namespace grp;
    const int n = cfg->n_data;
    const int KP = kp_of(h->corners);                     // 64, or 32 with an [Fe/H] axis
    const int n_stages = (n + KP - 1) / KP;
    const double LN2 = 0.69314718055994530942;
    bool ok = cfg->nspec == 2 && T.n_cells > 1 && n <= 65535 && n >= 1;
    // records, one block per stage: x = (ext - anchor of the stage) ln 2; d q_k(x) of norm_spec's orthonormal
    // quadratic basis
    std::vector<double2> anchor2((size_t)n_stages + 32, make_double2(0.0, std::numeric_limits<double>::infinity()));
    std::vector<double2> blk((size_t)n_stages * 3 * KP + 8, make_double2(0.0, 0.0));
    std::vector<double> dr2(4 * (size_t)n);
    double xmax = 0.0, dxmax = 0.0, pxmax = 0.0;
    for (int s = 0; s < n_stages; ++s) {
      const int d0 = s * KP, ns_ = std::min(KP, n - d0);
      const double anc = 0.5 * (cfg->ext_lo[d0] + cfg->ext_hi[d0 + ns_ - 1]);
      // the stage's common factor exp2(av * anchor) follows from the previous stage's by exp(av * step); every
      // ANCHOR_REFRESH-th one is taken exactly (step = +inf)
      anchor2[s].x = anc;
      if (s % ANCHOR_REFRESH != 0) {
        anchor2[s].y = (anc - anchor2[s - 1].x) * LN2;
        dxmax = std::max(dxmax, std::fabs(anchor2[s].y));
      }
      double2* pa = blk.data() + (size_t)s * 3 * KP;
      double2* pb = pa + KP; double2* pc = pb + KP;
      for (int k = 0; k < ns_; ++k) {
        const int i = d0 + k;
        const double xlo = (cfg->ext_lo[i] - anc) * LN2, xhi = (cfg->ext_hi[i] - anc) * LN2;
        xmax = std::max(xmax, std::max(std::fabs(xlo), std::fabs(xhi)));
        const double x = cfg->data_x[i], d = cfg->data_flux[i];
        const double q0 = T.basis[0], q1 = std::fma(T.basis[4], x, T.basis[3]);
        const double q2 = std::fma(std::fma(T.basis[8], x, T.basis[7]), x, T.basis[6]);
        pa[k] = make_double2(cfg->rebin_t[i], xlo);
        pb[k] = make_double2((cfg->ext_hi[i] - cfg->ext_lo[i]) * LN2, d * q0);
        pxmax = std::max(pxmax, std::fabs(pb[k].x));
        pc[k] = make_double2(d * q1, d * q2);
        dr2[4 * (size_t)i] = q1; dr2[4 * (size_t)i + 1] = q2; dr2[4 * (size_t)i + 2] = d;
        dr2[4 * (size_t)i + 3] = 1.0 / (cfg->data_err[i] * cfg->data_err[i]);
      }
    }
    grp::GroupTables& GT = h->GT;
    GT.n_stages = n_stages; GT.xmax = xmax; GT.dxmax = dxmax; GT.pxmax = pxmax; GT.n_teff = cfg->n_teff; GT.n_logg = cfg->n_logg;
    CKC(upload(h, blk.data(), blk.size(), &GT.recblk));
    CKC(upload(h, anchor2.data(), anchor2.size(), &GT.anchor));
    CKC(upload(h, dr2.data(), dr2.size(), &GT.drec2));
    // (bracket pair, Teff, log g[, Fe/H]) tensor map of the bracket table: one box = KP pixels of the 2 x 2 (x 2) bracketing
    // rows of a star
    if (ok && cfg->n_teff >= 2 && cfg->n_logg >= 2 && (cfg->n_feh == 1 || cfg->n_feh >= 2)) {
      typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
          qres != cudaDriverEntryPointSuccess) {
        ok = false;
      } else {
        const int rank = (h->corners == 8) ? 4 : 3;
        const cuuint64_t gdim[4] = {(cuuint64_t)2 * (cuuint64_t)n, (cuuint64_t)cfg->n_teff, (cuuint64_t)cfg->n_logg, (cuuint64_t)cfg->n_feh};
        const cuuint64_t gstr[3] = {(cuuint64_t)n * 16, (cuuint64_t)n * 16 * (cuuint64_t)cfg->n_teff,
                                    (cuuint64_t)n * 16 * (cuuint64_t)cfg->n_teff * (cuuint64_t)cfg->n_logg};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * KP), 2, 2, 2};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult cr = reinterpret_cast<EncodeFn>(fn)(&h->grid_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank,
                                                          const_cast<double2*>(T.bracket), gdim, gstr, box, estr,
                                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) ok = false;
      }
    } else {
      ok = false;
    }
    h->group_ok = ok;
    if (const char* v = std::getenv("MCMCSPEC_GROUP_VARIANT")) h->group_variant = (v[0] == 'w') ? 1 : (v[0] == 'n') ? 2 : 0;
    // persistent CTAs, two per SM
    h->group_ctas = 2 * h->num_sms;
    CKH(cudaMalloc((void**)&h->cell_start, ((size_t)T.n_cells + 1) * sizeof(unsigned)));
    CKH(cudaMalloc((void**)&h->group_off, (size_t)T.n_cells * sizeof(unsigned)));
    CKH(cudaMalloc((void**)&h->n_groups, sizeof(int)));
    CKH(cudaMalloc((void**)&h->group_counter, sizeof(int)));
    CKH(cudaMalloc((void**)&h->short_circuited, sizeof(unsigned long long)));
    CKH(cudaMemset(h->short_circuited, 0, sizeof(unsigned long long)));
    CKH(cudaMallocHost((void**)&h->h_n_groups, 2 * sizeof(int)));
    h->h_n_groups[0] = -1; h->h_n_groups[1] = 0;
    if (ok) {
      // per CTA: the transposed model vectors [n][32] and the walker warps' key lists [NWW_MAX + 1][LISTCAP][32]
      CKH(cudaMalloc((void**)&h->group_scratch, (size_t)h->group_ctas * (size_t)n * 32 * sizeof(double)));
      const size_t n_lists = std::max<size_t>((size_t)h->group_ctas * (MS_GROUP_NWW + 1), (size_t)h->num_sms * (MS_GROUP_NWW_WIDE + 1));
      CKH(cudaMalloc((void**)&h->group_list, n_lists * (size_t)LISTCAP * 32 * sizeof(unsigned long long)));
    }
  }

  // ---- isochrone, dust
  CKC(upload(h, cfg->iso_teff, cfg->n_iso, &T.iso_teff));
  CKC(upload(h, cfg->iso_logg, cfg->n_iso, &T.iso_logg));
  CKC(upload(h, cfg->iso_lum, cfg->n_iso, &T.iso_lum));
  CKC(upload(h, cfg->dust_dm_edges, cfg->n_dm, &T.dust_dm));
  CKC(upload(h, cfg->dust_samples, (size_t)cfg->n_dm * cfg->n_dust_samples, &T.dust_samples));

  // ---- shared memory budget of the streaming kernel
  const int nc = h->variant == 0 ? 2 : MS_MAX_CONTRAST, np = h->variant == 0 ? 6 : MS_MAX_PHOT;
  int max_optin = 0;
  CKH(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  h->max_optin_smem = max_optin;
  h->threads = MS_NT_SMALL;
  h->smem_bytes = stream_smem_bytes(cfg->n_data, nc, np, MS_NT_SMALL / 32, cfg->nspec);
  if (MS_FORCE_LARGE || 2 * (h->smem_bytes + 1024) > (size_t)max_optin + 1024) {
    // two CTAs no longer fit an SM: run one wide CTA instead (same warps in flight)
    h->threads = MS_NT_LARGE;
    h->smem_bytes = stream_smem_bytes(cfg->n_data, nc, np, MS_NT_LARGE / 32, cfg->nspec);
  }
  if (h->smem_bytes + 2048 > (size_t)max_optin) {
    // the observed spectrum does not fit shared memory: keep the per-walker model vector in a global scratch row
    // (L2 resident between the streaming and the chi-square phase); two 256-thread CTAs per SM again
    h->model_global = true;
    h->threads = MS_NT_SMALL;
    h->smem_bytes = stream_smem_bytes(0, nc, np, MS_NT_SMALL / 32, cfg->nspec);
  }
  CKH(cudaStreamSynchronize(h->stream));
#undef CKC
#undef CKH
  *out = h;
  return MS_OK;
}

int ms_set_stream(ms_handle* h, void* stream) {
  if (!h) return MS_ERR_INVALID;
  h->stream = (cudaStream_t)stream;
  return MS_OK;
}

int ms_use_own_stream(ms_handle* h) {
  if (!h) return MS_ERR_INVALID;
  h->stream = h->own_stream;
  return MS_OK;
}

int ms_set_cluster(ms_handle* h, int ctas_per_walker) {
  if (!h) return MS_ERR_INVALID;
  if (ctas_per_walker != 0 && ctas_per_walker != 1 && ctas_per_walker != 2 && ctas_per_walker != 4 && ctas_per_walker != 8) {
    set_err(h, "ms_set_cluster: 0 (automatic), 1, 2, 4 or 8"); return MS_ERR_INVALID;
  }
  h->cluster = ctas_per_walker;
  return MS_OK;
}

int ms_set_pairing(ms_handle* h, int enabled) {
  if (!h) return MS_ERR_INVALID;
  h->pairing = enabled < 0 ? -1 : (enabled != 0);
  return MS_OK;
}

int ms_set_fast(ms_handle* h, int enabled) {
  if (!h) return MS_ERR_INVALID;
  h->fast = enabled != 0;
  return MS_OK;
}

int ms_short_circuited(ms_handle* h, int64_t* count, int reset) {
  if (!h || !count) return MS_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  unsigned long long v = 0;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(&v, h->short_circuited, sizeof(v), cudaMemcpyDeviceToHost));
  if (reset) CK(cudaMemset(h->short_circuited, 0, sizeof(v)));
  *count = (int64_t)v;
  return MS_OK;
}

int ms_set_group(ms_handle* h, int mode, int64_t min_walkers) {
  if (!h) return MS_ERR_INVALID;
  h->group_mode = mode < 0 ? -1 : (mode != 0);
  if (min_walkers > 0) h->group_min_walkers = min_walkers;
  return MS_OK;
}

int ms_group_supported(const ms_handle* h) { return (h && h->group_ok) ? 1 : 0; }

int ms_set_group_shard(ms_handle* h, int rank, int world) {
  if (!h) return MS_ERR_INVALID;
  if (world < 1 || rank < 0 || rank >= world) { set_err(h, "ms_set_group_shard: need 0 <= rank < world"); return MS_ERR_INVALID; }
  h->shard_rank = rank; h->shard_world = world;
  return MS_OK;
}

int ms_group_will_run(const ms_handle* h, int64_t nw) {
  if (!h) return 0;
  return (h->group_ok && h->group_mode != 0 && !h->fast && h->cluster <= 1 &&
          (h->group_mode > 0 || nw >= h->group_min_walkers)) ? 1 : 0;
}

int ms_set_cell_order(ms_handle* h, int mode) {
  if (!h) return MS_ERR_INVALID;
  h->cell_order = mode < 0 ? -1 : (mode != 0);
  return MS_OK;
}

int ms_set_timing(ms_handle* h, int enabled) {
  if (!h) return MS_ERR_INVALID;
  h->timing = enabled != 0;
  return MS_OK;
}

int ms_kernel_time(ms_handle* h, double* total_ms, int64_t* n_launches, int64_t* n_walkers) {
  if (!h) return MS_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  double tot = 0.0;
  for (size_t k = 0; k + 1 < h->ev_used; k += 2) {
    CK(cudaEventSynchronize(h->ev_pool[k + 1]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev_pool[k], h->ev_pool[k + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = (int64_t)(h->ev_used / 2);
  if (n_walkers) *n_walkers = h->timed_walkers;
  h->ev_used = 0; h->timed_walkers = 0;
  return MS_OK;
}

int ms_set_optimizer_data(ms_handle* h, const double* flux_sets, int32_t n_sets) {
  if (!h) return MS_ERR_INVALID;
  if (!flux_sets || n_sets < 1) { set_err(h, "ms_set_optimizer_data: invalid argument"); return MS_ERR_INVALID; }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->d_opt_flux) cudaFree(h->d_opt_flux);
  if (h->d_opt_median) cudaFree(h->d_opt_median);
  h->d_opt_flux = h->d_opt_median = nullptr; h->opt_sets = 0; h->T.opt_median = nullptr;
  const int n = h->T.n_data;
  std::vector<double> med(n_sets), tmp(n);
  for (int k = 0; k < n_sets; ++k) {                      // np.median of each set (mft6.py:1011)
    std::copy(flux_sets + (size_t)k * n, flux_sets + (size_t)(k + 1) * n, tmp.begin());
    std::sort(tmp.begin(), tmp.end());
    med[k] = (n & 1) ? tmp[n / 2] : (tmp[n / 2 - 1] + tmp[n / 2]) / 2.0;
  }
  CK(cudaMalloc((void**)&h->d_opt_flux, (size_t)n_sets * n * 8));
  CK(cudaMalloc((void**)&h->d_opt_median, (size_t)n_sets * 8));
  CK(cudaMemcpy(h->d_opt_flux, flux_sets, (size_t)n_sets * n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_opt_median, med.data(), (size_t)n_sets * 8, cudaMemcpyHostToDevice));
  h->opt_sets = n_sets; h->T.opt_median = h->d_opt_median;
  return MS_OK;
}

int ms_optimizer_chisq(ms_handle* h, const double* theta, const int32_t* set_index, const double* rad_sigma,
                       const double* av_prior, const double* ref_median, int64_t nw, double* chisq) {
  NvtxRange nvtx_range("ms_optimizer_chisq");
  if (!h) return MS_ERR_INVALID;
  if (nw < 0 || (nw > 0 && (!theta || !set_index || !rad_sigma || !chisq))) { set_err(h, "ms_optimizer_chisq: null argument"); return MS_ERR_INVALID; }
  if (h->opt_sets < 1) { set_err(h, "ms_optimizer_chisq: call ms_set_optimizer_data first"); return MS_ERR_INVALID; }
  if (h->T.feh_free) { set_err(h, "ms_optimizer_chisq: the optimiser stage is built for a fixed [Fe/H]"); return MS_ERR_UNSUPPORTED; }
  if (nw == 0) return MS_OK;
  for (int64_t w = 0; w < nw; ++w)
    if (set_index[w] < 0 || set_index[w] >= h->opt_sets) { set_err(h, "ms_optimizer_chisq: set_index out of range"); return MS_ERR_INVALID; }
  CK(cudaSetDevice(h->device));
  const int ndim = h->T.ndim, ns = h->T.nspec;
  int rc = ensure_io(h, nw);
  if (rc) return rc;
  // host buffers in, host buffer out (the optimiser's random walk runs on the host); the per-call inputs live in
  // one handle-owned scratch buffer (this entry point is called thousands of times per fit)
  const size_t a8 = ((size_t)nw * 8 + 255) & ~(size_t)255;
  const size_t need = ns * a8 + 3 * a8;
  if (need > h->scratch_cap) {
    CK(cudaStreamSynchronize(h->stream));
    if (h->scratch) cudaFree(h->scratch);
    h->scratch = nullptr; h->scratch_cap = 0;
    CK(cudaMalloc((void**)&h->scratch, need + need / 2));
    h->scratch_cap = need + need / 2;
  }
  double* d_rs = reinterpret_cast<double*>(h->scratch);
  int* d_idx = reinterpret_cast<int*>(h->scratch + ns * a8);
  double* d_avp = av_prior ? reinterpret_cast<double*>(h->scratch + ns * a8 + a8) : nullptr;
  double* d_med = ref_median ? reinterpret_cast<double*>(h->scratch + ns * a8 + 2 * a8) : nullptr;
  cudaError_t e = cudaMemcpyAsync(h->d_theta, theta, (size_t)nw * ndim * 8, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rs, rad_sigma, (size_t)nw * ns * 8, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, set_index, (size_t)nw * 4, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && av_prior) e = cudaMemcpyAsync(d_avp, av_prior, (size_t)nw * 8, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && ref_median) e = cudaMemcpyAsync(d_med, ref_median, (size_t)nw * 8, cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) { set_err(h, std::string("ms_optimizer_chisq: ") + cudaGetErrorString(e)); rc = MS_ERR_CUDA; }
  if (rc == MS_OK) {
    PartsOut parts{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    OptIn opt{h->d_opt_flux, d_idx, d_rs, d_avp, d_med};
    rc = evaluate_device(h, h->d_theta, nw, h->d_lnp, parts, MODE_OPTIMIZER, opt);
  }
  if (rc == MS_OK) {
    e = cudaMemcpyAsync(chisq, h->d_lnp, (size_t)nw * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { set_err(h, std::string("ms_optimizer_chisq: ") + cudaGetErrorString(e)); rc = MS_ERR_CUDA; }
  } else {
    cudaStreamSynchronize(h->stream);
  }
  return rc;
}

int ms_lnprob_parts(ms_handle* h, const double* theta, int64_t nw, double* lnprob, double* chi2_spec,
                    double* contrast, double* phot, double* lnprior, double* logg, int32_t* status, int where) {
  NvtxRange nvtx_range("ms_lnprob_parts");
  if (!h) return MS_ERR_INVALID;
  if (nw < 0 || (nw > 0 && (!theta || !lnprob))) { set_err(h, "ms_lnprob: null argument"); return MS_ERR_INVALID; }
  if (nw == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  const int ndim = h->T.ncol;                          // columns of theta
  if (where == MS_DEVICE) {
    PartsOut parts{chi2_spec, contrast, phot, lnprior, logg, status, nullptr, nullptr};
    return evaluate_device(h, theta, nw, lnprob, parts);
  }
  if (where != MS_HOST) { set_err(h, "ms_lnprob: bad `where`"); return MS_ERR_INVALID; }
  int rc = ensure_io(h, nw);
  if (rc) return rc;
  std::memcpy(h->h_theta, theta, (size_t)nw * ndim * 8);
  CK(cudaMemcpyAsync(h->d_theta, h->h_theta, (size_t)nw * ndim * 8, cudaMemcpyHostToDevice, h->stream));
  // optional outputs: device scratch, copied back after the kernel
  struct Opt { void* host; size_t bytes; void* dev; };
  Opt opts[6] = {{chi2_spec, (size_t)nw * 8, nullptr}, {contrast, (size_t)nw * h->T.n_contrast * 8, nullptr},
                 {phot, (size_t)nw * h->T.n_phot * 8, nullptr}, {lnprior, (size_t)nw * 8, nullptr},
                 {logg, (size_t)nw * h->T.nspec * 8, nullptr}, {status, (size_t)nw * 4, nullptr}};
  DevMem opt_mem[6];
  for (int k = 0; k < 6; ++k)
    if (opts[k].host && opts[k].bytes) { CK(opt_mem[k].alloc(opts[k].bytes)); opts[k].dev = opt_mem[k].p; }
  PartsOut parts{(double*)opts[0].dev, (double*)opts[1].dev, (double*)opts[2].dev, (double*)opts[3].dev,
                 (double*)opts[4].dev, (int*)opts[5].dev, nullptr, nullptr};
  rc = evaluate_device(h, h->d_theta, nw, h->d_lnp, parts);
  if (rc == MS_OK) {
    cudaError_t e = cudaMemcpyAsync(h->h_lnp, h->d_lnp, (size_t)nw * 8, cudaMemcpyDeviceToHost, h->stream);
    for (auto& o : opts) if (o.dev && e == cudaSuccess) e = cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { set_err(h, std::string("ms_lnprob: ") + cudaGetErrorString(e)); rc = MS_ERR_CUDA; }
    else std::memcpy(lnprob, h->h_lnp, (size_t)nw * 8);
  }
  return rc;
}

int ms_lnprob(ms_handle* h, const double* theta, int64_t nw, double* lnprob, int where) {
  return ms_lnprob_parts(h, theta, nw, lnprob, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, where);
}

int ms_model_spectrum(ms_handle* h, const double* theta, int64_t nw, double* out, int where) {
  NvtxRange nvtx_range("ms_model_spectrum");
  if (!h) return MS_ERR_INVALID;
  if (nw < 0 || (nw > 0 && (!theta || !out))) { set_err(h, "ms_model_spectrum: null argument"); return MS_ERR_INVALID; }
  if (nw == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  const int ndim = h->T.ncol, n = h->T.n_data;
  int rc = ensure_io(h, nw);
  if (rc) return rc;
  const double* d_theta = theta;
  double* d_out = out;
  DevMem tmp_out;
  if (where == MS_HOST) {
    std::memcpy(h->h_theta, theta, (size_t)nw * ndim * 8);
    CK(cudaMemcpyAsync(h->d_theta, h->h_theta, (size_t)nw * ndim * 8, cudaMemcpyHostToDevice, h->stream));
    d_theta = h->d_theta;
    CK(tmp_out.alloc((size_t)nw * n * 8));
    d_out = tmp_out.as<double>();
  }
  PartsOut parts{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_out, nullptr};
  rc = evaluate_device(h, d_theta, nw, h->d_lnp, parts);
  if (where == MS_HOST) {
    if (rc == MS_OK) {
      cudaError_t e = cudaMemcpyAsync(out, d_out, (size_t)nw * n * 8, cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) { set_err(h, std::string("ms_model_spectrum: ") + cudaGetErrorString(e)); rc = MS_ERR_CUDA; }
    } else {
      cudaStreamSynchronize(h->stream);
    }
  }
  return rc;
}

int ms_component_mags(ms_handle* h, const double* theta, int64_t nw, double* out, int where) {
  NvtxRange nvtx_range("ms_component_mags");
  if (!h) return MS_ERR_INVALID;
  if (nw < 0 || (nw > 0 && (!theta || !out))) { set_err(h, "ms_component_mags: null argument"); return MS_ERR_INVALID; }
  if (nw == 0 || h->T.n_contrast == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  const int ndim = h->T.ncol;
  const size_t per = (size_t)h->T.n_contrast * h->T.nspec;
  int rc = ensure_io(h, nw);
  if (rc) return rc;
  const double* d_theta = theta;
  double* d_out = out;
  DevMem tmp_out;
  if (where == MS_HOST) {
    std::memcpy(h->h_theta, theta, (size_t)nw * ndim * 8);
    CK(cudaMemcpyAsync(h->d_theta, h->h_theta, (size_t)nw * ndim * 8, cudaMemcpyHostToDevice, h->stream));
    d_theta = h->d_theta;
    CK(tmp_out.alloc((size_t)nw * per * 8));
    d_out = tmp_out.as<double>();
  }
  PartsOut parts{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_out};
  rc = evaluate_device(h, d_theta, nw, h->d_lnp, parts);
  if (where == MS_HOST) {
    if (rc == MS_OK) {
      cudaError_t e = cudaMemcpyAsync(out, d_out, (size_t)nw * per * 8, cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) { set_err(h, std::string("ms_component_mags: ") + cudaGetErrorString(e)); rc = MS_ERR_CUDA; }
    } else {
      cudaStreamSynchronize(h->stream);
    }
  }
  return rc;
}

int ms_broaden_rows(int device, double* rows, int64_t n_rows, int64_t row_len, int64_t win_lo, int64_t win_len,
                    const double* taps, int32_t n_taps) {
  NvtxRange nvtx_range("ms_broaden_rows");
  ms_handle* h = nullptr;   // errors land in the create-error slot
  if (!rows || !taps || n_rows < 0 || row_len <= 0 || win_lo < 0 || win_len < 0 || win_lo + win_len > row_len ||
      n_taps < 1 || (n_taps % 2) == 0 || n_taps > 4097 || (win_len > 0 && win_len < n_taps) || win_len > 0x7fffffff) {
    set_err(h, "ms_broaden_rows: invalid argument (n_taps must be odd and not exceed the window)"); return MS_ERR_INVALID;
  }
  if (n_rows == 0 || win_len == 0) return MS_OK;
  CK(cudaSetDevice(device));
  const long long chunk = std::max<long long>(1, std::min<long long>(n_rows, (1ll << 28) / row_len));   // <= 2 GiB per buffer
  DevMem m_in, m_out, m_taps;
  CK(m_in.alloc((size_t)chunk * row_len * 8));
  CK(m_out.alloc((size_t)chunk * row_len * 8));
  CK(m_taps.alloc((size_t)n_taps * 8));
  double *d_in = m_in.as<double>(), *d_out = m_out.as<double>(), *d_taps = m_taps.as<double>();
  int rc = MS_OK;
  cudaError_t e = cudaMemcpy(d_taps, taps, (size_t)n_taps * 8, cudaMemcpyHostToDevice);
  const size_t smem = ((size_t)bpad_host(BT + n_taps - 1) + 1 + (size_t)n_taps) * 8;
  if (e == cudaSuccess && smem > 48 * 1024)
    e = cudaFuncSetAttribute(broaden_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (long long r0 = 0; r0 < n_rows && e == cudaSuccess; r0 += chunk) {
    const long long nr = std::min<long long>(chunk, n_rows - r0);
    e = cudaMemcpy(d_in, rows + r0 * row_len, (size_t)nr * row_len * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_out, d_in, (size_t)nr * row_len * 8, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) break;
    dim3 grid((unsigned)((win_len + BT - 1) / BT), (unsigned)nr);
    broaden_rows_kernel<<<grid, BTHREADS, smem>>>(d_in, d_out, row_len, win_lo, (int)win_len, d_taps, n_taps);
    broaden_patch_kernel<<<(unsigned)nr, 32>>>(d_out, row_len, win_lo, (int)win_len);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(rows + r0 * row_len, d_out, (size_t)nr * row_len * 8, cudaMemcpyDeviceToHost);
  }
  if (e != cudaSuccess) { set_err(h, std::string("ms_broaden_rows: ") + cudaGetErrorString(e)); rc = MS_ERR_CUDA; }
  return rc;
}

int ms_build_grid(int device, const double* src_x, const double* src_y, const int64_t* src_off, int64_t n_rows,
                  const double* x_dst, int64_t row_len, int64_t win_lo, int64_t win_len, const double* taps,
                  int32_t n_taps, double* rows_out) {
  NvtxRange nvtx_range("ms_build_grid");
  ms_handle* h = nullptr;   // errors land in the create-error slot
  if (!src_x || !src_y || !src_off || !x_dst || !rows_out || n_rows < 0 || row_len <= 0 || win_lo < 0 || win_len < 0 ||
      win_lo + win_len > row_len || win_len > 0x7fffffff) {
    set_err(h, "ms_build_grid: invalid argument"); return MS_ERR_INVALID;
  }
  if (win_len > 0 && (!taps || n_taps < 1 || (n_taps % 2) == 0 || n_taps > 4097 || win_len < n_taps)) {
    set_err(h, "ms_build_grid: n_taps must be odd and not exceed the window"); return MS_ERR_INVALID;
  }
  if (n_rows == 0) return MS_OK;
  for (int64_t r = 0; r < n_rows; ++r)
    if (src_off[r + 1] - src_off[r] < 2) { set_err(h, "ms_build_grid: a node spectrum needs at least two samples"); return MS_ERR_INVALID; }
  CK(cudaSetDevice(device));
  const size_t n_src = (size_t)src_off[n_rows];
  double *d_sx = nullptr, *d_sy = nullptr, *d_x = nullptr, *d_rows = nullptr, *d_rows2 = nullptr, *d_taps = nullptr;
  long long* d_off = nullptr; int* d_oob = nullptr;
  int rc = MS_OK;
  cudaError_t e = cudaMalloc((void**)&d_sx, n_src * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_sy, n_src * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_off, (size_t)(n_rows + 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_x, (size_t)row_len * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_rows, (size_t)n_rows * row_len * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_oob, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(d_oob, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(d_sx, src_x, n_src * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_sy, src_y, n_src * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_off, src_off, (size_t)(n_rows + 1) * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_x, x_dst, (size_t)row_len * 8, cudaMemcpyHostToDevice);
  int oob = 0;
  if (e == cudaSuccess) {
    dim3 grid((unsigned)((row_len + 255) / 256), (unsigned)n_rows);
    resample_rows_kernel<<<grid, 256>>>(d_sx, d_sy, d_off, d_x, row_len, d_rows, d_oob);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(&oob, d_oob, sizeof(int), cudaMemcpyDeviceToHost);
  }
  double* d_final = d_rows;
  if (e == cudaSuccess && oob == 0 && win_len > 0) {
    // window broadening + splice (mft6.py:373-376): the stencil reads the resampled rows and writes a copy
    e = cudaMalloc((void**)&d_rows2, (size_t)n_rows * row_len * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_taps, (size_t)n_taps * 8);
    if (e == cudaSuccess) e = cudaMemcpy(d_taps, taps, (size_t)n_taps * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_rows2, d_rows, (size_t)n_rows * row_len * 8, cudaMemcpyDeviceToDevice);
    const size_t smem = ((size_t)bpad_host(BT + n_taps - 1) + 1 + (size_t)n_taps) * 8;
    if (e == cudaSuccess && smem > 48 * 1024)
      e = cudaFuncSetAttribute(broaden_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      dim3 grid((unsigned)((win_len + BT - 1) / BT), (unsigned)n_rows);
      broaden_rows_kernel<<<grid, BTHREADS, smem>>>(d_rows, d_rows2, row_len, win_lo, (int)win_len, d_taps, n_taps);
      broaden_patch_kernel<<<(unsigned)n_rows, 32>>>(d_rows2, row_len, win_lo, (int)win_len);
      e = cudaGetLastError();
    }
    d_final = d_rows2;
  }
  if (e == cudaSuccess && oob == 0) e = cudaMemcpy(rows_out, d_final, (size_t)n_rows * row_len * 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { set_err(h, std::string("ms_build_grid: ") + cudaGetErrorString(e)); rc = MS_ERR_CUDA; }
  else if (oob) {
    set_err(h, (oob & 1) ? "A value in x_new is below the interpolation range." : "A value in x_new is above the interpolation range.");
    rc = MS_ERR_INVALID;
  }
  cudaFree(d_sx); cudaFree(d_sy); cudaFree(d_off); cudaFree(d_x); cudaFree(d_rows); cudaFree(d_oob);
  if (d_rows2) cudaFree(d_rows2);
  if (d_taps) cudaFree(d_taps);
  return rc;
}

int ms_stretch_iteration(ms_handle* h, double* state, int32_t* naccepted, int64_t nw, double a, uint64_t seed,
                         uint64_t iteration, int32_t randomize_split) {
  NvtxRange nvtx_range("ms_stretch_iteration");
  if (!h) return MS_ERR_INVALID;
  if (!state || nw < 2 || (nw % 2) || nw > SPLIT_SMALL_MAX) {
    set_err(h, "ms_stretch_iteration: nw must be even and at most 2048 (larger ensembles: the per-kernel entry points)");
    return MS_ERR_INVALID;
  }
  CK(cudaSetDevice(h->device));
  const int ndim = h->T.ncol;
  const long long half = nw / 2;
  if (half > h->it_cap) {
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 2; ++k) { if (h->it_half[k]) cudaFree(h->it_half[k]); h->it_half[k] = nullptr; }
    if (h->it_q) cudaFree(h->it_q); if (h->it_logfac) cudaFree(h->it_logfac);
    if (h->it_lnp) cudaFree(h->it_lnp); if (h->it_packed) cudaFree(h->it_packed);
    h->it_q = h->it_logfac = h->it_lnp = h->it_packed = nullptr; h->it_cap = 0;
    for (int k = 0; k < 2; ++k) CK(cudaMalloc((void**)&h->it_half[k], (size_t)half * 8));
    CK(cudaMalloc((void**)&h->it_q, (size_t)half * ndim * 8));
    CK(cudaMalloc((void**)&h->it_logfac, (size_t)half * 8));
    CK(cudaMalloc((void**)&h->it_lnp, (size_t)half * 8));
    CK(cudaMalloc((void**)&h->it_packed, (size_t)half * (ndim + 2) * 8));
    h->it_cap = half;
  }
  int npow = 2;
  while (npow < nw) npow <<= 1;
  split_perm_small_kernel<<<1, 1024, 0, h->stream>>>(seed, iteration, (int)nw, npow, randomize_split, h->it_half[0], h->it_half[1]);
  CK(cudaGetLastError());
  h->launches += 1;
  const unsigned g = (unsigned)((half + 127) / 128);
  const PartsOut none{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int split = 0; split < 2; ++split) {
    const long long* s_idx = h->it_half[split];
    const long long* c_idx = h->it_half[1 - split];
    // propose + prepare in one kernel, evaluate, accept in place: three launches per half step
    const ProposeIn prop{state, s_idx, c_idx, half, a, seed, iteration, split, 0ll, h->it_q, h->it_logfac};
    const int rc = evaluate_device(h, h->it_q, half, h->it_lnp, none, -1, OptIn{nullptr, nullptr, nullptr, nullptr, nullptr}, prop);
    if (rc) return rc;
    stretch_accept_inplace_kernel<<<g, 128, 0, h->stream>>>(state, naccepted, s_idx, h->it_q, h->it_lnp, h->it_logfac, half, ndim,
                                                            seed, iteration, split, 0ll);
    CK(cudaGetLastError());
    h->launches += 1;
  }
  return MS_OK;
}

int ms_stretch_halfstep(ms_handle* h, double* state, int32_t* naccepted, const int64_t* s_idx, int64_t n_s,
                        const int64_t* c_idx, int64_t n_c, double a, uint64_t seed, uint64_t iteration, int32_t split,
                        int64_t gid0, double* q, double* logfac, double* lnp_q, double* packed) {
  NvtxRange nvtx_range("ms_stretch_halfstep");
  if (!h) return MS_ERR_INVALID;
  if (n_s < 0 || n_c < 1 || !state || !s_idx || !c_idx || !q || !logfac || !lnp_q) {
    set_err(h, "ms_stretch_halfstep: invalid argument"); return MS_ERR_INVALID;
  }
  if (n_s == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  const int ndim = h->T.ncol;
  // proposal folded into the prep kernel, evaluation, then the acceptance test: in place (one rank owns the whole
  // half) or into the packed rows the ranks exchange
  const ProposeIn prop{state, (const long long*)s_idx, (const long long*)c_idx, n_c, a, seed, iteration, split, (long long)gid0, q, logfac};
  const PartsOut none{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const int rc = evaluate_device(h, q, n_s, lnp_q, none, -1, OptIn{nullptr, nullptr, nullptr, nullptr, nullptr}, prop);
  if (rc) return rc;
  const unsigned g = (unsigned)((n_s + 127) / 128);
  if (packed) {
    stretch_accept_kernel<<<g, 128, 0, h->stream>>>(state, (const long long*)s_idx, q, lnp_q, logfac, n_s, ndim, seed, iteration,
                                                    split, gid0, packed);
  } else {
    stretch_accept_inplace_kernel<<<g, 128, 0, h->stream>>>(state, naccepted, (const long long*)s_idx, q, lnp_q, logfac, n_s, ndim,
                                                            seed, iteration, split, gid0);
  }
  CK(cudaGetLastError());
  h->launches += 1;
  return MS_OK;
}

int ms_stretch_evaluate(ms_handle* h, const double* state, const int64_t* s_idx, int64_t n_s, const int64_t* c_idx, int64_t n_c,
                        double a, uint64_t seed, uint64_t iteration, int32_t split, int64_t gid0, double* q, double* logfac,
                        double* lnp_q) {
  NvtxRange nvtx_range("ms_stretch_evaluate");
  if (!h) return MS_ERR_INVALID;
  if (n_s < 0 || n_c < 1 || !state || !s_idx || !c_idx || !q || !logfac || !lnp_q) {
    set_err(h, "ms_stretch_evaluate: invalid argument"); return MS_ERR_INVALID;
  }
  if (n_s == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  if (h->shard_world > 1) {
    if (!ms_group_will_run(h, n_s)) { set_err(h, "ms_stretch_evaluate: a group shard is set but this batch does not take the group kernel"); return MS_ERR_UNSUPPORTED; }
    CK(cudaMemsetAsync(lnp_q, 0, (size_t)n_s * sizeof(double), h->stream));        // the other ranks' groups: summed in by the caller
  }
  const ProposeIn prop{state, (const long long*)s_idx, (const long long*)c_idx, n_c, a, seed, iteration, split, (long long)gid0, q, logfac};
  const PartsOut none{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  h->shard_active = h->shard_world > 1;
  const int rc = evaluate_device(h, q, n_s, lnp_q, none, -1, OptIn{nullptr, nullptr, nullptr, nullptr, nullptr}, prop);
  h->shard_active = false;
  return rc;
}

int ms_stretch_accept_inplace(ms_handle* h, double* state, int32_t* naccepted, const int64_t* s_idx, int64_t n_s, const double* q,
                              const double* lnp_q, const double* logfac, uint64_t seed, uint64_t iteration, int32_t split,
                              int64_t gid0) {
  if (!h) return MS_ERR_INVALID;
  if (n_s < 0 || !state || !s_idx || !q || !lnp_q || !logfac) { set_err(h, "ms_stretch_accept_inplace: invalid argument"); return MS_ERR_INVALID; }
  if (n_s == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  stretch_accept_inplace_kernel<<<(unsigned)((n_s + 127) / 128), 128, 0, h->stream>>>(state, naccepted, (const long long*)s_idx, q, lnp_q,
                                                                                     logfac, n_s, h->T.ncol, seed, iteration, split, gid0);
  CK(cudaGetLastError());
  h->launches += 1;
  return MS_OK;
}

int ms_split_keys(ms_handle* h, uint64_t seed, uint64_t iteration, int64_t n, int64_t* keys) {
  if (!h) return MS_ERR_INVALID;
  if (n < 0 || (n > 0 && !keys)) { set_err(h, "ms_split_keys: invalid argument"); return MS_ERR_INVALID; }
  if (n == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  split_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(seed, iteration, n, (long long*)keys);
  CK(cudaGetLastError());
  h->launches += 1;
  return MS_OK;
}

int ms_stretch_propose(ms_handle* h, const double* state, const int64_t* s_idx, int64_t n_s, const int64_t* c_idx,
                       int64_t n_c, int32_t ndim, double a, uint64_t seed, uint64_t iteration, int32_t split,
                       int64_t gid0, double* q, double* logfac) {
  if (!h) return MS_ERR_INVALID;
  if (n_s < 0 || n_c < 1 || ndim < 1 || !state || !s_idx || !c_idx || !q || !logfac) {
    set_err(h, "ms_stretch_propose: invalid argument"); return MS_ERR_INVALID;
  }
  if (n_s == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  stretch_propose_kernel<<<(unsigned)((n_s + 127) / 128), 128, 0, h->stream>>>(
      state, (const long long*)s_idx, n_s, (const long long*)c_idx, n_c, ndim, a, seed, iteration, split, gid0, q, logfac);
  CK(cudaGetLastError());
  h->launches += 1;
  return MS_OK;
}

int ms_stretch_accept(ms_handle* h, const double* state, const int64_t* s_idx, const double* q, const double* lnp_q,
                      const double* logfac, int64_t n_s, int32_t ndim, uint64_t seed, uint64_t iteration,
                      int32_t split, int64_t gid0, double* packed) {
  if (!h) return MS_ERR_INVALID;
  if (n_s < 0 || ndim < 1 || !state || !s_idx || !q || !lnp_q || !logfac || !packed) {
    set_err(h, "ms_stretch_accept: invalid argument"); return MS_ERR_INVALID;
  }
  if (n_s == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  stretch_accept_kernel<<<(unsigned)((n_s + 127) / 128), 128, 0, h->stream>>>(
      state, (const long long*)s_idx, q, lnp_q, logfac, n_s, ndim, seed, iteration, split, gid0, packed);
  CK(cudaGetLastError());
  h->launches += 1;
  return MS_OK;
}

int ms_stretch_scatter(ms_handle* h, double* state, int32_t* naccepted, const int64_t* idx, const double* packed,
                       int64_t n, int32_t ndim) {
  if (!h) return MS_ERR_INVALID;
  if (n < 0 || ndim < 1 || !state || !idx || !packed) { set_err(h, "ms_stretch_scatter: invalid argument"); return MS_ERR_INVALID; }
  if (n == 0) return MS_OK;
  CK(cudaSetDevice(h->device));
  stretch_scatter_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(state, naccepted, (const long long*)idx,
                                                                            packed, n, ndim);
  CK(cudaGetLastError());
  h->launches += 1;
  return MS_OK;
}

}  // extern "C"
#endif  // MS_HAS_MAIN
