// group_kernel.cuh -- lnprob_group_kernel: the large-batch form of the streaming evaluation (included by
// mcmcspec.cu; instantiated in translation unit MS_TU = 4).
//
// One persistent CTA evaluates *groups* of up to 32 walkers that sit in the same grid cell, i.e. whose 2 * 2^D
// bracketing grid rows (mft6.py:439-519) are the same rows.  What a walker needs of a row is, per observed pixel, the
// row's two model pixels that bracket it (mft6.py:1169: interp1d onto the observed wavelengths) -- the bracket table
// T.bracket[row][pixel] built once in ms_create -- and the row's integrals through the passbands (T.bandsum; np.trapz
// of mft6.py:731 resp. pyphot get_flux, mft6.py:776, both linear in the spectrum, so a walker's band fluxes are the
// same linear combination of the row integrals as its spectrum is of the rows, mft6.py:507-519, :689-700).  The
// bracket rows are streamed ONCE per group, KP observed pixels (one stage) at a time, through a shared-memory ring:
//
//   copies                 -- per stage ONE tiled tensor copy per star (cp.async.bulk.tensor, SASS UTMALDG: 3-D box =
//                             KP bracket pairs x 2 Teff nodes x 2 log g nodes = the star's four bracketing rows; 4-D
//                             with two [Fe/H] nodes = its eight rows on grids with that axis) and one
//                             bulk copy (UBLKCP) of the stage's per-observed-pixel records; completion is counted in
//                             bytes on the slot's `full` mbarrier.  There is no producer warp: the walker warps count
//                             themselves out of a slot (one shared-memory atomic per warp and stage) and the last one
//                             out issues the copies of the stage NST ahead into it.  The chi-square pass's per-pixel
//                             records go through the same ring in the same way.
//   all warps   walkers    -- LANE = WALKER: lane l of every walker warp serves walker l of the group and keeps its
//                             eight row coefficients and Av in registers; the warps split the observed pixels of
//                             each stage between them.  The staged table and the records are the same for all lanes, so
//                             every shared-memory read is a one-wavefront broadcast.  Per pixel and lane: composite at
//                             the two bracketing pixels (FMA chain over the rows, then the sum over the stars,
//                             mft6.py:744), extinction of both (mft6.py:1162), linear rebin (mft6.py:1169),
//                             norm_spec's moments (mft6.py:193-196), key range and a scale-free 64-bin histogram for
//                             the median; the model value goes to a transposed scratch row [pixel][lane] (coalesced).
//                             After the stream the partial sums of the warps are combined in a fixed order, the
//                             warps split one pass over the scratch for the chi-square (mft6.py:1178-1179) that also
//                             collects the keys of the median's histogram bin, and the first walker warp finishes:
//                             exact selection on those keys (mft6.py:1173, np.median convention), contrasts,
//                             magnitudes, priors, final sum (mft6.py:1182-1205).  No shuffles; the only atomics
//                             are the slot counters.
//
// Extinction: exp2(Av * ext(p)) = A_s * exp(Av * x_p) with A_s = exp2(Av * anchor_s) per stage and walker and an
// 8th-order polynomial on |Av * x_p| <= 0.045 (x^9/9! < 2.1e-18); walkers with larger exponents take exp2 of the whole
// exponent.  The upper bracket's factor is the lower one's times a 4th-order series in Av (x_hi - x_lo) (neighbouring
// model pixels).  A_s itself is taken exactly on every 8th stage and follows from the previous stage's by the
// polynomial in between.  Every product / sum whose fusion is optional is spelled out, so that all instantiations
// round alike.

namespace grp {

// -DMS_GROUP_TIMING: every warp of the first four CTAs prints, when its CTA runs out of groups, where its cycles went:
// waiting for a group, phase A, the barrier after it, phase B (bin of the median, chi-square pass, lists), candidate
// ranking and final sum (first warp only), and the time inside mbarrier waits
#ifdef MS_GROUP_TIMING
#define GTIME(i) { const long long t__ = clock64(); gt_acc[i] += t__ - gt_last; gt_last = t__; }
#define GDBG_WAIT(stmt) { const long long t__ = clock64(); stmt; gt_wait += clock64() - t__; }
#else
#define GTIME(i)
#define GDBG_WAIT(stmt) stmt;
#endif

#ifndef MS_GROUP_KP
#define MS_GROUP_KP 64
#endif
#ifndef MS_GROUP_NST
#define MS_GROUP_NST 5
#endif
#ifndef MS_GROUP_VW
#define MS_GROUP_VW 16
#endif
constexpr int KP = MS_GROUP_KP;            // observed pixels per stage (2-D grids: 4 rows per star)
constexpr int KP8 = MS_GROUP_KP / 2;       // ... with an [Fe/H] axis (8 rows per star): the same bytes per stage
constexpr int kp_of(int nr) { return nr == 8 ? KP8 : KP; }
constexpr int NST = MS_GROUP_NST;          // ring slots: stages in flight ahead of the walker warps
constexpr int ANCHOR_REFRESH = 8;          // the stage's common extinction factor is taken exactly on every 8th stage
constexpr int NWW_MAX = MS_GROUP_VW;       // walker warps per CTA (template parameter NWW: 8 with two CTAs per SM, 16 with
                                           // one): they split the observed pixels of the group's 32 walkers
constexpr int GMAX = 32;                   // walkers per group (one per lane)
constexpr int HBITS = 6;
constexpr int HBINS = 1 << HBITS;          // per-lane histogram bins
constexpr int LISTCAP = 1024;              // keys of the median's bin a lane collects (per walker warp)
constexpr int VW = MS_GROUP_VW;            // partial sums per walker: the quadruples of observed pixels resp. the blocks of a
                                           // chi-square chunk are dealt to partial sum (index mod VW) whatever the number of
                                           // walker warps, so that a walker's value does not depend on the kernel variant
constexpr int PIX = 4;                     // observed pixels a walker lane keeps in flight (independent FMA chains)
constexpr int CHI_U = 4;                   // observed pixels per block of the chi-square pass
constexpr int CHI_SB = VW * CHI_U;         // a super-block: one block for every partial sum
constexpr int CHI_NSB = 7;
constexpr int CHI_CHUNK = CHI_NSB * CHI_SB;   // observed pixels per chunk of the chi-square pass's 32-byte records
constexpr int CHI_BYTES = CHI_CHUNK * 32;
constexpr int CAND = 8;                    // keys of the median's second-level bin per lane and walker warp (fast finish)
constexpr double POLY_XMAX = 0.045;
constexpr double STEP_XMAX = 5.0e-4;
static_assert(KP % PIX == 0 && KP8 % PIX == 0, "a stage is a whole number of quadruples");

// One ring slot: the bracket table of the stage {star 0: [log g][Teff][KP] double2 | star 1} and its records.
template <int NR> struct StageLayout {
  static constexpr int NROW = 2 * NR;
  static constexpr int kp = kp_of(NR);                     // observed pixels per stage
  static constexpr int rowb = kp * 16;                     // one row of the stage
  static constexpr int box = NR * rowb;                    // one star's rows, written by one tensor copy
  static constexpr int B = 0;
  static constexpr int rec = 2 * box;                      // {t, x_lo}[KP] | {x_hi - x_lo, d q0}[KP] | {d q1, d q2}[KP]
  static constexpr int bytes = rec + kp * 48;
  static constexpr int ring = NST * bytes;
  static_assert(box % 128 == 0 && bytes % 128 == 0, "bulk copies need aligned destinations");
  static constexpr int cslots = ring / CHI_BYTES;          // the chi-square pass cuts the same ring into chunk slots
  static_assert(CHI_BYTES % 128 == 0 && cslots >= 2, "chi-square chunks: aligned, at least two in flight");
};

// What the group kernel needs beyond DevTables (built once in ms_create).
struct GroupTables {
  const double2* recblk;        // per stage: {t, x_lo}[KP] | {x_hi - x_lo, d q0}[KP] | {d q1, d q2}[KP]
  const double2* anchor;        // [n_stages] {extinction exponent the stage's x values are relative to, (its step from the
                                //  previous stage) ln 2, or +inf: take the exact factor here}
  const double* drec2;          // [n_data][4] {q1(x), q2(x), flux, 1/err^2}: phase B
  double xmax;                  // max |x| over all records (x = (ext - anchor) ln 2)
  double dxmax;                 // max |step|
  double pxmax;                 // max |x_hi - x_lo| over all records
  int n_stages;
  int n_teff, n_logg;
  int shard_rank, shard_world;  // this launch evaluates the groups shard_rank, shard_rank + shard_world, ... (ms_set_group_shard)
};

struct GroupDesc { int start, count; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" :: "r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// the same with a fourth coordinate ([Fe/H] node): the eight bracketing rows of one star in one request
__device__ __forceinline__ void tensor_g2s_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               :: "r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
// 3-D tiled tensor copy (pixel, Teff node, log g node) global -> shared: the four bracketing rows of one star in
// one request (TMA, SASS UTMALDG); out-of-range coordinates are filled with zeros
__device__ __forceinline__ void tensor_g2s_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               :: "r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// ++(uint16 counter in shared memory), invisible to the compiler's alias analysis (the caller brackets the phase with
// barriers that are compiler fences)
__device__ __forceinline__ void hist_inc(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  v = (unsigned short)(v + 1);
  asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"(v));
}

// *p = v for the lanes whose predicate is set, as one predicated store (no branch)
__device__ __forceinline__ void store_if(unsigned long long* p, double v, bool pred) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f64 [%0], %1;\n\t}" :: "l"(p), "d"(v), "r"((unsigned)pred));
}

// exp(y) for |y| <= POLY_XMAX: Taylor to y^8 (remainder y^9/9! < 2.1e-18 relative at 0.045).  The coefficients sit in
// constant memory so that every FMA takes its coefficient as a constant-bank operand.
__constant__ double EXP_C[7] = {1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
__device__ __forceinline__ double exp_small(double y) {
  double p = fma(y, EXP_C[0], EXP_C[1]);
  p = fma(p, y, EXP_C[2]); p = fma(p, y, EXP_C[3]); p = fma(p, y, EXP_C[4]); p = fma(p, y, EXP_C[5]);
  p = fma(p, y, EXP_C[6]); p = fma(p, y, 1.0); p = fma(p, y, 1.0);
  return p;
}

// exp(y) for |y| <= STEP_XMAX = 5e-4: Taylor to y^4 (remainder y^5/120 < 2.7e-19 relative)
__device__ __forceinline__ double exp_step(double y) {
  double p = fma(y, 1.0 / 24.0, EXP_C[5]);
  p = fma(p, y, EXP_C[6]); p = fma(p, y, 1.0); p = fma(p, y, 1.0);
  return p;
}

// Ordered 32-bit key of the high word of a double (same order as ordered_key() on the full value).
__device__ __forceinline__ unsigned hi_key(double v) {
  const int h = __double2hiint(v);
  return (unsigned)h ^ ((unsigned)(h >> 31) | 0x80000000u);
}

// ---- per-lane exact selection --------------------------------------------------------------------------
// Every lane selects on its own column of a transposed array (element t of lane l at [t * 32 + l]) with its own
// histogram column hist[bin * 32 + l] (uint16 counts: at most 65 535 elements), so there is nothing to share.
struct LaneSel {
  unsigned long long prefix;    // the top `bits` bits of the wanted key
  unsigned long long above;     // smallest key seen so far that is larger than every key with this prefix
  unsigned rank;                // rank of the wanted key among the keys with this prefix
  unsigned cnt;                 // how many keys have this prefix
  int bits;
};

__device__ __forceinline__ void lane_hist_zero(unsigned short* hist, int lane) {
#pragma unroll 8
  for (int b = 0; b < HBINS; ++b) hist[b * 32 + lane] = 0;
}

// Scan the bins in the order first + t (mod HBINS), t = 0, 1, ..., for the one holding `rank`.
__device__ __forceinline__ void lane_locate(const unsigned short* hist, int lane, unsigned first, unsigned rank, unsigned& bin_t,
                                            unsigned& below, unsigned& in_bin) {
  unsigned cum = 0;
  bool found = false;
  bin_t = 0; below = 0; in_bin = 0;
#pragma unroll 4
  for (unsigned t = 0; t < (unsigned)HBINS; ++t) {
    const unsigned c = hist[((first + t) & (HBINS - 1)) * 32 + lane];
    if (!found && rank < cum + c) { found = true; bin_t = t; below = cum; in_bin = c; }
    cum += c;
  }
}

// One more HBITS-bit digit on the elements [0, n_l) of the lane's column: histogram of the keys that carry the lane's
// prefix, then new prefix / rank / count.  VALUES: the column holds doubles (keys are derived), else it holds keys.
// Lanes with n_l == 0 or all 64 bits fixed only take part in the warp-wide trip count.
template <bool VALUES>
__device__ __forceinline__ void lane_narrow(const void* col, unsigned n_l, LaneSel& S, unsigned short* hist, int lane) {
  const bool active = S.bits < 64 && n_l > 0;
  const int width = min(HBITS, 64 - S.bits);
  const int shift = max(64 - S.bits - width, 0);
  const unsigned long long himask = (S.bits == 0) ? 0ull : (~0ull << (64 - S.bits));
  lane_hist_zero(hist, lane);
  const unsigned n_max = __reduce_max_sync(0xffffffffu, active ? n_l : 0u);
  for (unsigned t = 0; t < n_max; ++t) {
    if (active && t < n_l) {
      const unsigned long long k = VALUES ? ordered_key(reinterpret_cast<const double*>(col)[(size_t)t * 32 + lane])
                                          : reinterpret_cast<const unsigned long long*>(col)[(size_t)t * 32 + lane];
      if ((k & himask) == S.prefix) hist[(unsigned)((k >> shift) & ((1u << width) - 1u)) * 32 + lane] += 1;
    }
  }
  if (active) {
    unsigned bin, below, in_bin;
    lane_locate(hist, lane, 0u, S.rank, bin, below, in_bin);
    S.rank -= below; S.cnt = in_bin;
    S.prefix |= ((unsigned long long)bin) << shift;
    S.bits += width;
  }
}

// In-place compaction of a lane's key column to the keys carrying its (new) prefix; dropped keys above the prefix
// feed `above`.
__device__ __forceinline__ unsigned lane_compact(unsigned long long* col, unsigned n_l, LaneSel& S, int lane) {
  const unsigned long long himask = (S.bits >= 64) ? ~0ull : (~0ull << (64 - S.bits));
  const unsigned n_max = __reduce_max_sync(0xffffffffu, n_l);
  unsigned out = 0;
  for (unsigned t = 0; t < n_max; ++t) {
    if (t < n_l) {
      const unsigned long long k = col[(size_t)t * 32 + lane];
      if ((k & himask) == S.prefix) { col[(size_t)out * 32 + lane] = k; ++out; }
      else if (k > S.prefix) S.above = min(S.above, k);
    }
  }
  return out;
}

// The sum of the NWW walker warps' histograms, formed in place in warp 0's: warp wj adds up the bins [wj B, (wj + 1) B),
// B = HBINS / NWW, over all warps and writes them to warp 0's columns (a lane's counts add up to at most n <= 65 535).
// After a barrier every warp scans the one merged histogram (lane_locate) instead of NWW of them.
template <int NWW>
__device__ __forceinline__ void lane_merge(unsigned short* hist_all, int wj, int lane) {
  static_assert(HBINS % NWW == 0, "bins are dealt evenly to the warps");
  constexpr int BPW = HBINS / NWW;
#pragma unroll
  for (int k = 0; k < BPW; ++k) {
    const int b = wj * BPW + k;
    unsigned c = 0;
#pragma unroll
    for (int j = 0; j < NWW; ++j) c += hist_all[(j * HBINS + b) * 32 + lane];
    hist_all[b * 32 + lane] = (unsigned short)c;
  }
}

// ++counter in shared memory.  Relaxed: what must come first are the warp's READS of the slot, and those have returned
// their values (they were consumed by the arithmetic before this point); a release fence would also wait for the warp's
// global stores of the model values, hundreds of cycles per stage.  The counter is never reset (the NWW-th, 2 NWW-th, ...
// arrival is the last one out of a stage).
__device__ __forceinline__ unsigned count_out(uint32_t addr) {
  unsigned old;
  asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
  return old;
}

template <int NWW>
__device__ __forceinline__ void walker_bar() { asm volatile("bar.sync 1, %0;" :: "n"(NWW * 32) : "memory"); }

// ------------------------------------------------------------------------------------------
// NR: grid rows per star (4, or 8 with an [Fe/H] axis: a 4-D box per star).  Binaries, posterior mode.
// NWW walker warps per CTA.
// ------------------------------------------------------------------------------------------
template <int NR, int NWW, int CTAS_PER_SM>
__global__ void __launch_bounds__(NWW * 32, CTAS_PER_SM)
lnprob_group_kernel(const __grid_constant__ DevTables T, const __grid_constant__ GroupTables GT,
                    const __grid_constant__ CUtensorMap bracket_map, const WalkerPlan* __restrict__ plans,
                    const int* __restrict__ order, const GroupDesc* __restrict__ groups, const int* __restrict__ n_groups,
                    int* __restrict__ group_counter, double* __restrict__ lnprob, const PartsOut parts,
                    double* __restrict__ scratch, unsigned long long* __restrict__ list_scratch) {
  constexpr int NS = 2;
  constexpr int NROW = NS * NR;
  constexpr int NTHREADS = NWW * 32;
  static_assert(NWW <= NWW_MAX && VW % NWW == 0, "scratch is sized for NWW_MAX walker warps; partial sums are dealt to the warps");
  constexpr int NV = VW / NWW;                                     // partial sums a walker warp carries
  using L = StageLayout<NR>;
  constexpr int KPn = L::kp;
  constexpr int CSLOTS = L::cslots;
  extern __shared__ __align__(128) unsigned char gsm[];
  unsigned short* const hist_all = reinterpret_cast<unsigned short*>(gsm + L::ring);   // [NWW][HBINS][32]
  __shared__ __align__(8) unsigned long long s_full[NST], s_cfull[CSLOTS];
  __shared__ unsigned s_out[NST], s_cout[CSLOTS];        // walker warps through with the slot's stage / chunk
  __shared__ double s_band[BAND_COLS * NROW];            // row integrals through every band
  __shared__ double s_part[VW][4][32];                   // per partial sum and lane: three moments / chi-square
  __shared__ unsigned long long s_above[NWW][32];
  __shared__ unsigned s_kmin[NWW][32], s_kmax[NWW][32], s_nl[NWW][32], s_nc[NWW][32];
  __shared__ int s_row[NROW];
  __shared__ int s_box[NS][3];                           // (Teff, log g, [Fe/H]) origin of each star's box in the bracket table
  __shared__ int s_group, s_status;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(&s_full[s]), 1); s_out[s] = 0u; }
    for (int s = 0; s < CSLOTS; ++s) { mbar_init(smem_u32(&s_cfull[s]), 1); s_cout[s] = 0u; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int n_stages = GT.n_stages;
  const int n = T.n_data;
  const int n_chunks = (n + CHI_CHUNK - 1) / CHI_CHUNK;
  const int total_groups = *n_groups;
  const uint32_t ring = smem_u32(gsm);
  const int wj = warp;                                             // walker warp index
  // the CTA's scratch, reused by every group it takes: model values [pixel][lane], key lists [warp][slot][lane]
  double* const sMT = scratch + (size_t)blockIdx.x * (size_t)n * 32;
  unsigned long long* const list0 = list_scratch + (size_t)blockIdx.x * (NWW + 1) * (size_t)LISTCAP * 32;       // + a spare list
  unsigned long long* const listT = list0 + (size_t)wj * LISTCAP * 32;
  unsigned short* const hist = hist_all + (size_t)wj * HBINS * 32;
  unsigned it0 = 0, cit0 = 0;               // ring uses so far (stages / chi-square chunks): slot and mbarrier parity go on

#ifdef MS_GROUP_TIMING
  long long gt_acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gt_last = clock64();
  int gt_groups = 0;
  long long gt_wait = 0, gt_wait_a = 0;
#endif
  // Persistent CTA: groups are handed out by an atomic counter (reset by group_scan_kernel).
  while (true) {
    if (tid == 0) s_group = atomicAdd(group_counter, 1) * GT.shard_world + GT.shard_rank;
    __syncthreads();                                                // also: the previous group is completely done
    const int g = s_group;
    GTIME(0)
    if (g >= total_groups) break;
#ifdef MS_GROUP_TIMING
    ++gt_groups;
#endif
    const GroupDesc gd = groups[g];
    const int count = gd.count;

    if (tid == 0) {
      const WalkerPlan& p = plans[order[gd.start]];
      s_status = p.status;
#pragma unroll
      for (int st = 0; st < NS; ++st) {
#pragma unroll
        for (int r = 0; r < NR; ++r) s_row[st * NR + r] = p.row[st][r];
        // box origin: the star's lower (Teff, log g) nodes (the plan's corners are in box order)
        s_box[st][0] = p.row[st][0] % GT.n_teff; s_box[st][1] = (p.row[st][0] / GT.n_teff) % GT.n_logg;
        s_box[st][2] = p.row[st][0] / (GT.n_teff * GT.n_logg);
      }
    }
    __syncthreads();

    if (s_status != 0) {
      // the walkers that never reach the grid are sorted into their own cell: nothing to stream (mft6.py:1465-1466)
      for (int k = tid; k < count; k += NTHREADS) {
        const int w = order[gd.start + k];
        const WalkerPlan& p = plans[w];
        lnprob[w] = -CUDART_INF;
        if (parts.status) parts.status[w] = p.status;
        if (parts.lnprior) parts.lnprior[w] = -CUDART_INF;
        if (parts.chi2_spec) parts.chi2_spec[w] = CUDART_NAN;
        if (parts.logg) for (int s = 0; s < NS; ++s) parts.logg[w * NS + s] = p.logg[s];
        if (parts.contrast) for (int c = 0; c < T.n_contrast; ++c) parts.contrast[w * T.n_contrast + c] = CUDART_NAN;
        if (parts.phot) for (int f = 0; f < T.n_phot; ++f) parts.phot[w * T.n_phot + f] = CUDART_NAN;
        if (parts.comp_mag) for (int f = 0; f < T.n_contrast * NS; ++f) parts.comp_mag[w * T.n_contrast * NS + f] = CUDART_NAN;
        if (parts.model) for (int i = 0; i < n; ++i) parts.model[(long long)w * n + i] = CUDART_NAN;
      }
      continue;
    }
    // the rows' integrals through the passbands (read by the finishing warp, after the barriers below)
    for (int k = tid; k < BAND_COLS * NROW; k += NTHREADS)
      s_band[k] = __ldg(T.bandsum + (long long)s_row[k % NROW] * BAND_COLS + k / NROW);

    // ---- walker-lane state (walker warps only): lane l serves walker l of the group in every walker warp; idle
    // lanes shadow walker 0 (same arithmetic, nothing stored)
    const bool live = lane < count;
    const int w = order[gd.start + (live ? lane : 0)];
    double cf[NROW];
#pragma unroll
    for (int r = 0; r < NROW; ++r) cf[r] = 0.0;
    double av = 0.0;

    // the copies of stage s into its ring slot (one thread): the two stars' bracket boxes and the record block
    auto issue_stage = [&](int s) {
      const unsigned gs = it0 + (unsigned)s;
      const int slot = (int)(gs % NST);
      const uint32_t st = ring + (uint32_t)slot * L::bytes;
      const uint32_t bar = smem_u32(&s_full[slot]);
      mbar_expect_tx(bar, (unsigned)L::bytes);                        // pixels past the end of the spectrum are zero-filled
#pragma unroll
      for (int stx = 0; stx < NS; ++stx) {
        if constexpr (NR == 8)
          tensor_g2s_4d(st + L::B + stx * L::box, &bracket_map, 2 * KPn * s, s_box[stx][0], s_box[stx][1], s_box[stx][2], bar);
        else
          tensor_g2s_3d(st + L::B + stx * L::box, &bracket_map, 2 * KPn * s, s_box[stx][0], s_box[stx][1], bar);
      }
      bulk_g2s(st + L::rec, GT.recblk + (size_t)s * 3 * KPn, KPn * 48, bar);
    };
    if (lane == 0 && wj < min(NST, n_stages)) issue_stage(wj);       // the first NST stages, one per warp
    {
      // ================================================================= walker warps, phase A: lane = walker
      double momv[NV][3];
#pragma unroll
      for (int v = 0; v < NV; ++v) { momv[v][0] = 0.0; momv[v][1] = 0.0; momv[v][2] = 0.0; }
      unsigned kmin = 0xffffffffu, kmax = 0u;
      {
        const WalkerPlan& p = plans[w];
#pragma unroll
        for (int st = 0; st < NS; ++st) {
#pragma unroll
          for (int r = 0; r < NR; ++r) cf[st * NR + r] = p.coef[st][r];
        }
        av = p.av;
        lane_hist_zero(hist, lane);
      }
      // which form of the extinction factor a walker gets depends on its own Av only (never on its companions in the
      // group): the polynomial where it is accurate, exp2 of the whole exponent otherwise
      const bool lane_poly = av * GT.xmax <= POLY_XMAX, lane_inc = av * GT.dxmax <= POLY_XMAX;
      const bool lane_step = av * GT.pxmax <= STEP_XMAX;             // the upper bracket's factor from the lower one's
      const bool poly_ok = __all_sync(0xffffffffu, lane_poly && lane_step);
      double A = 1.0;
      double2 anc_next = __ldg(GT.anchor);
      int qrot = 0;                       // quadruples of the stages so far, mod VW: where this stage's first one is dealt
      const uint32_t hist_s = smem_u32(hist);
      for (int s = 0; s < n_stages; ++s) {
        const unsigned gs = it0 + (unsigned)s;
        const int slot = (int)(gs % NST);
        const double2 anc2 = anc_next;
        // the next stage's anchor, in flight during this stage (volatile: issued here, not sunk to its use)
        asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(anc_next.x), "=d"(anc_next.y) : "l"(GT.anchor + min(s + 1, n_stages - 1)));
        // the stage's common extinction factor exp2(av * anchor): exact on every ANCHOR_REFRESH-th stage, from the
        // previous stage's by the polynomial in between (the same sequence in every warp and kernel variant)
        const double anc = anc2.x;
        if (lane_inc && anc2.y < CUDART_INF) A = __dmul_rn(A, exp_small(__dmul_rn(av, anc2.y))); else A = exp2(__dmul_rn(av, anc));
        const int n_obs = min(KPn, n - s * KPn);
        GDBG_WAIT(mbar_wait(smem_u32(&s_full[slot]), (gs / NST) & 1u))
        const unsigned char* stB = gsm + (size_t)slot * L::bytes + L::B;
        {
          double* dst = sMT + (size_t)s * KPn * 32 + lane;
          const double2* recA = reinterpret_cast<const double2*>(gsm + (size_t)slot * L::bytes + L::rec);
          const double2* recB = recA + KPn;
          const double2* recC = recB + KPn;
          // one observed pixel for this lane's walker; every shared-memory operand is the same address in all
          // lanes (one-wavefront broadcasts)
          auto pixel = [&](auto poly, int k, double& rm_out) -> double {
            const double2 ra = recA[k];                                      // t, x_lo
            const double dx = recB[k].x;                                     // x_hi - x_lo
            double flo = 0.0, fhi = 0.0;
#pragma unroll
            for (int stx = 0; stx < NS; ++stx) {
              const double2 b0 = *reinterpret_cast<const double2*>(stB + stx * L::box + k * 16);
              double l = __dmul_rn(cf[stx * NR], b0.x), u = __dmul_rn(cf[stx * NR], b0.y);
#pragma unroll
              for (int r = 1; r < NR; ++r) {
                const double2 b = *reinterpret_cast<const double2*>(stB + stx * L::box + r * L::rowb + k * 16);
                l = fma(cf[stx * NR + r], b.x, l); u = fma(cf[stx * NR + r], b.y, u);
              }
              flo = (stx == 0) ? l : __dadd_rn(flo, l); fhi = (stx == 0) ? u : __dadd_rn(fhi, u);        // mft6.py:744
            }
            double elo, ehi, scale_a = A;
            if constexpr (decltype(poly)::value) {
              // the two bracketing model pixels are neighbours: exp(av x_hi) = exp(av x_lo) * exp(av dx), |av dx| <= 5e-4
              elo = exp_small(__dmul_rn(av, ra.y)); ehi = __dmul_rn(elo, exp_step(__dmul_rn(av, dx)));
            } else {
              // some lane's exponent is too large for the polynomial: those lanes take exp2 of the whole exponent
              // (1/ln 2 = 1.4426950408889634), the others the same arithmetic as on the fast path
              const double e2lo = exp2(av * fma(ra.y, 1.4426950408889634, anc));
              const double xhi = __dadd_rn(ra.y, dx);
              const double e2hi = exp2(av * fma(xhi, 1.4426950408889634, anc));
              elo = lane_poly ? exp_small(__dmul_rn(av, ra.y)) : e2lo;
              ehi = lane_poly ? (lane_step ? __dmul_rn(elo, exp_step(__dmul_rn(av, dx))) : exp_small(__dmul_rn(av, xhi))) : e2hi;
              scale_a = lane_poly ? A : 1.0;
            }
            // every product and sum is spelled out (no mul/add pair is left for the assembler to fuse or not): a walker's
            // value must not depend on the kernel variant
            const double ylo = __dmul_rn(flo, elo), yhi = __dmul_rn(fhi, ehi);     // extinct() at both brackets, mft6.py:1162
            const double m = __dmul_rn(scale_a, fma(__dsub_rn(yhi, ylo), ra.x, ylo));   // interp1d onto the observed pixel, mft6.py:1169
            rm_out = fast_rcp(m);
            return m;
          };
          auto tally = [&](int k, double m, double rm, double (&mom)[3]) {
            const double dq0 = recB[k].y; const double2 rc = recC[k];      // d q0, d q1, d q2
            mom[0] = fma(dq0, rm, mom[0]); mom[1] = fma(rc.x, rm, mom[1]); mom[2] = fma(rc.y, rm, mom[2]);   // norm_spec moments, mft6.py:194-195
            const unsigned hk = hi_key(m);
            kmin = min(kmin, hk); kmax = max(kmax, hk);
            // 3 exponent bits + HBITS - 3 mantissa bits.  The counter update goes through asm without a memory clobber so
            // that the compiler does not order the next pixels' (plain) shared-memory loads behind it
            hist_inc(hist_s + (((hk >> (20 - (HBITS - 3))) & (HBINS - 1)) * 32 + lane) * 2);
            dst[(size_t)k * 32] = m;
          };
          auto run = [&](auto poly) {
            // quadruple q of the stage (pixels 4q .. 4q+3) belongs to partial sum (qrot + q) mod VW: the quadruples are
            // dealt round robin along the whole vector, so the warps' shares even out over the stages; this warp
            // carries the partials wj, wj + NWW, ...
            for (int q0 = -qrot; q0 * PIX < n_obs; q0 += VW) {
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                const int k = (q0 + wj + v * NWW) * PIX;
                if (k >= 0 && k + PIX <= n_obs) {                            // warp-uniform: a whole quadruple
                  double mv[PIX], rmv[PIX];
#pragma unroll
                  for (int u = 0; u < PIX; ++u) mv[u] = pixel(poly, k + u, rmv[u]);      // no branch: the chains interleave
#pragma unroll
                  for (int u = 0; u < PIX; ++u) tally(k + u, mv[u], rmv[u], momv[v]);
                } else if (k >= 0 && k < n_obs) {                            // the last, partial quadruple of the vector
                  for (int u = 0; k + u < n_obs; ++u) {
                    double rm;
                    const double m = pixel(poly, k + u, rm);
                    tally(k + u, m, rm, momv[v]);
                  }
                }
              }
            }
          };
          if (poly_ok) run(std::true_type{}); else run(std::false_type{});
        }
        qrot = (qrot + (n_obs + PIX - 1) / PIX) % VW;
        __syncwarp();
        // count this warp out of the slot; the last one out refills it with the stage NST ahead
        if (lane == 0 && count_out(smem_u32(&s_out[slot])) % NWW == (unsigned)(NWW - 1)) {
          if (s + NST < n_stages) issue_stage(s + NST);
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s_part[wj + v * NWW][c][lane] = momv[v][c];
      }
      s_kmin[wj][lane] = kmin; s_kmax[wj][lane] = kmax;
    }
    it0 += (unsigned)n_stages;
    GTIME(1)
#ifdef MS_GROUP_TIMING
    gt_wait_a += gt_wait; gt_wait = 0;
#endif

    __syncthreads();          // band integrals and the walker warps' partial sums complete
    GTIME(2)

    // ================================================================= the chi-square pass's records {q1, q2, d, 1/e^2}
    // through the ring, CHI_CHUNK observed pixels per slot (the model values are read straight from the scratch)
    auto issue_chunk = [&](int c) {
      const unsigned gs = cit0 + (unsigned)c;
      const int slot = (int)(gs % CSLOTS);
      const unsigned npx = (unsigned)min(CHI_CHUNK, n - c * CHI_CHUNK);
      mbar_expect_tx(smem_u32(&s_cfull[slot]), npx * 32u);
      bulk_g2s(ring + (uint32_t)slot * CHI_BYTES, GT.drec2 + (size_t)c * CHI_CHUNK * 4, npx * 32u, smem_u32(&s_cfull[slot]));
    };
    if (lane == 0 && wj < min(CSLOTS, n_chunks)) issue_chunk(wj);
    {
      // ================================================================= walker warps, phase B: lane = walker
      // ---- the walker's sums over all its pixels, combined in a fixed order (every warp ends up with the same values)
      double mom[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double a = s_part[0][c][lane];
#pragma unroll
        for (int j = 1; j < VW; ++j) a += s_part[j][c][lane];
        mom[c] = a;
      }
      unsigned kmin = s_kmin[0][lane], kmax = s_kmax[0][lane];
#pragma unroll
      for (int j = 1; j < NWW; ++j) { kmin = min(kmin, s_kmin[j][lane]); kmax = max(kmax, s_kmax[j][lane]); }
      // ---- the bin of the median.  The phase-A histograms are indexed by 3 exponent + HBITS - 3 mantissa bits of the
      // key; their sum orders the keys as long as a walker's values span fewer than 8 binades and are finite
      LaneSel S;
      S.rank = (unsigned)((n - 1) >> 1); S.above = ~0ull; S.cnt = (unsigned)n; S.prefix = 0ull; S.bits = 0;
      int has_nan = 0;
      const bool suspicious = kmax >= 0xfff00000u || kmin <= 0x000fffffu;   // an infinity or a NaN may be present
      if (__any_sync(0xffffffffu, suspicious)) {
        for (int i = 0; i < n; ++i) { const double m = sMT[(size_t)i * 32 + lane]; has_nan |= (m != m); }   // mft6.py:1202
      }
      const bool fast_hist = ((kmax >> 20) - (kmin >> 20)) < 8u && !suspicious;
      const bool all_fast = __all_sync(0xffffffffu, fast_hist);
      lane_merge<NWW>(hist_all, wj, lane);
      walker_bar<NWW>();
      if (all_fast) {
        constexpr int SH = 20 - (HBITS - 3);                                // hi-key bits below the histogram digit
        unsigned bin_t, below, in_bin;
        const unsigned first = (kmin >> SH) & ~((1u << (HBITS - 3)) - 1u);     // lowest binade, mantissa bits 0
        lane_locate(hist_all, lane, first, S.rank, bin_t, below, in_bin);
        S.prefix = (unsigned long long)(first + bin_t) << (32 + SH); S.bits = 32 - SH;
        S.rank -= below; S.cnt = in_bin;
      }
      walker_bar<NWW>();                                                         // every warp has read the phase-A histograms
      if (!all_fast) {
        // general case: digits below the common prefix of the lane's key range, from the scratch column
        const unsigned long long k0 = (unsigned long long)kmin << 32, k1 = ((unsigned long long)kmax << 32) | 0xffffffffull;
        S.bits = __clzll((long long)(k0 ^ k1));
        S.prefix = (S.bits == 0) ? 0ull : (k0 & (~0ull << (64 - S.bits)));
        lane_narrow<true>(sMT, (unsigned)n, S, hist, lane);
      }
      while (__any_sync(0xffffffffu, S.cnt > (unsigned)LISTCAP && S.bits < 64)) {
        LaneSel S2 = S;
        if (!(S.cnt > (unsigned)LISTCAP && S.bits < 64)) S2.bits = 64;     // this lane is done: trip count only
        lane_narrow<true>(sMT, (unsigned)n, S2, hist, lane);
        if (S.cnt > (unsigned)LISTCAP && S.bits < 64) S = S2;
      }
      // ---- one pass over the model vector, split between the warps in blocks of 4 pixels: chi-square terms
      // (mft6.py:196, :120 on the unscaled model) + the keys of the median's bin into the warp's list and the smallest
      // key above the bin.  The per-pixel records {q1, q2, d, 1/e^2} arrive through the ring (bulk copies), the model
      // values are read one super-block ahead.  Branch free: eight independent pixels are in flight per lane.
      const double a0q0 = __dmul_rn(mom[0], T.basis[0]), a1 = mom[1], a2 = mom[2];
      double chiv[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) chiv[v] = 0.0;
      unsigned nl = 0;
      unsigned long long above = ~0ull;                                     // smallest key above the wanted (sub-)bin seen so far
      const bool collect = S.cnt <= (unsigned)LISTCAP && S.bits < 64;      // per lane; otherwise the slow path below
      const int width2 = min(HBITS, 64 - S.bits);
      const int shift2 = max(64 - S.bits - width2, 0);
      lane_hist_zero(hist, lane);
      GTIME(6)
      {
        constexpr int U = CHI_U;
        // HI32: every lane's prefix sits in the high word of the key (always so on the histogram path): 32-bit tests
        auto chi_pass = [&](auto hi32) {
          constexpr bool HI32 = decltype(hi32)::value;
          const unsigned long long himask = (S.bits >= 64) ? ~0ull : (S.bits == 0 ? 0ull : (~0ull << (64 - S.bits)));
          const unsigned long long prefix = S.prefix;
          const unsigned himask_h = (unsigned)(himask >> 32), prefix_h = (unsigned)(prefix >> 32);
          const unsigned bound_h = prefix_h | ~himask_h;                   // high word of the bin's largest key
          // a lane that does not collect (its bin is too large or empty: the slow path below) never hits
          const unsigned long long hitmask = collect ? himask : 0ull, hitprefix = collect ? prefix : 1ull;
          const unsigned hitmask_h = collect ? himask_h : 0u, hitprefix_h = collect ? prefix_h : 1u;
          unsigned above_h = 0xffffffffu;
          unsigned long long* lp = listT + lane;                            // next free slot of the lane's list column
          // this warp's pixels of super-block sb: blocks wj, wj + NWW, ... (block b belongs to partial sum b); the model
          // values are read from the scratch one super-block ahead (coalesced: the lanes' walkers are adjacent)
          auto load_m = [&](int sb, double (&mv)[NV][U]) {
            if (sb * CHI_SB + CHI_SB <= n) {
              const double* mp = sMT + ((size_t)sb * CHI_SB + wj * U) * 32 + lane;
#pragma unroll
              for (int v = 0; v < NV; ++v) {
#pragma unroll
                for (int u = 0; u < U; ++u) mv[v][u] = __ldcg(mp + (v * NWW * U + u) * 32);
              }
            } else {
#pragma unroll
              for (int v = 0; v < NV; ++v) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                  const int px = min(sb * CHI_SB + (wj + v * NWW) * U + u, n - 1);
                  mv[v][u] = __ldcg(sMT + (size_t)px * 32 + lane);
                }
              }
            }
          };
          // one pixel after its chi-square term: a key of the median's bin goes to the list by a predicated store (no branch:
          // a pixel is a hit for a few lanes at most; a branch would run its body once per hit with one lane on); `touch`:
          // a value between the bin and `above` was seen
          auto collect_px = [&](double m, bool valid, bool& touch) {
            bool hit;
            if constexpr (HI32) {
              const unsigned hk = hi_key(m);
              hit = valid && (hk & hitmask_h) == hitprefix_h;
              touch = touch || (valid && hk > bound_h && hk <= above_h);
            } else {
              const unsigned long long key = ordered_key(m);
              hit = valid && (key & hitmask) == hitprefix;
              touch = touch || (valid && (key & himask) != prefix && key > prefix && key < above);
            }
            store_if(lp, m, hit);
            lp += hit ? 32 : 0;
          };
          // smallest key above the bin (the upper median of an even-sized vector when the lower one is the bin's largest
          // key): the running minimum rarely changes, one warp-uniform branch per super-block
          auto above_px = [&](double m, bool valid) {
            const unsigned long long key = ordered_key(m);
            bool over;
            if constexpr (HI32) over = (unsigned)(key >> 32) > bound_h;
            else over = (key & himask) != prefix && key > prefix;
            if (valid && over) above = min(above, key);
          };
          double mnext[NV][U];
          load_m(0, mnext);
          for (int c = 0; c < n_chunks; ++c) {
            const unsigned gs = cit0 + (unsigned)c;
            const int slot = (int)(gs % CSLOTS);
            GDBG_WAIT(mbar_wait(smem_u32(&s_cfull[slot]), (gs / CSLOTS) & 1u))
            const double2* rec = reinterpret_cast<const double2*>(gsm + (size_t)slot * CHI_BYTES);   // {q1, q2}, {d, 1/e^2} per pixel
            const int npx = min(CHI_CHUNK, n - c * CHI_CHUNK);
            for (int j = 0; j < CHI_NSB && j * CHI_SB < npx; ++j) {
              const int sb = c * CHI_NSB + j;
              double mval[NV][U];
#pragma unroll
              for (int v = 0; v < NV; ++v) {
#pragma unroll
                for (int u = 0; u < U; ++u) mval[v][u] = mnext[v][u];
              }
              load_m(sb + 1, mnext);                                         // clamped past the end, never used there
              bool touch = false;
              if ((j + 1) * CHI_SB <= npx) {
                // a whole super-block (all but the last one of the vector): the reciprocals of the fits first, eight
                // independent chains, then the terms
                const double2* rq = rec + 2 * (j * CHI_SB + wj * U);
                double rcv[NV][U];
#pragma unroll
                for (int v = 0; v < NV; ++v) {
#pragma unroll
                  for (int u = 0; u < U; ++u) {
                    const double2 q = rq[2 * (v * NWW * U + u)];            // same address in every lane
                    rcv[v][u] = fast_rcp(fma(a2, q.y, fma(a1, q.x, a0q0)));
                  }
                }
#pragma unroll
                for (int v = 0; v < NV; ++v) {
#pragma unroll
                  for (int u = 0; u < U; ++u) {
                    const double2 dw = rq[2 * (v * NWW * U + u) + 1];
                    const double r = fma(-dw.x, rcv[v][u], mval[v][u]);
                    chiv[v] = __dadd_rn(chiv[v], __dmul_rn(__dmul_rn(r, r), dw.y));
                    collect_px(mval[v][u], true, touch);
                  }
                }
                if (__any_sync(0xffffffffu, touch)) {
#pragma unroll
                  for (int v = 0; v < NV; ++v) {
#pragma unroll
                    for (int u = 0; u < U; ++u) above_px(mval[v][u], true);
                  }
                  above_h = (unsigned)(above >> 32);
                }
              } else {
                // the last, partial super-block: pixel by pixel
#pragma unroll
                for (int v = 0; v < NV; ++v) {
#pragma unroll
                  for (int u = 0; u < U; ++u) {
                    const int px = j * CHI_SB + (wj + v * NWW) * U + u;      // within the chunk
                    const bool valid = px < npx;
                    const int pp = min(px, npx - 1);
                    const double2 q = rec[2 * pp], dw = rec[2 * pp + 1];
                    const double m = mval[v][u];
                    const double r = fma(-dw.x, fast_rcp(fma(a2, q.y, fma(a1, q.x, a0q0))), m);
                    chiv[v] = __dadd_rn(chiv[v], valid ? __dmul_rn(__dmul_rn(r, r), dw.y) : 0.0);
                    collect_px(m, valid, touch);
                  }
                }
                if (__any_sync(0xffffffffu, touch)) {
#pragma unroll
                  for (int v = 0; v < NV; ++v) {
#pragma unroll
                    for (int u = 0; u < U; ++u) above_px(mval[v][u], j * CHI_SB + (wj + v * NWW) * U + u < npx);
                  }
                  above_h = (unsigned)(above >> 32);
                }
              }
            }
            __syncwarp();
            // count this warp out of the chunk's slot; the last one out refills it
            if (lane == 0 && count_out(smem_u32(&s_cout[slot])) % NWW == (unsigned)(NWW - 1)) {
              if (c + CSLOTS < n_chunks) issue_chunk(c + CSLOTS);
            }
          }
          nl = (unsigned)((lp - (listT + lane)) >> 5);
        };
        if (__all_sync(0xffffffffu, S.bits <= 32)) chi_pass(std::true_type{}); else chi_pass(std::false_type{});
      }
      GTIME(7)
      // ---- the warp's list: raw values become ordered keys in place, second-level histogram of the next key bits
      // (every lane walks its own column)
      {
        const unsigned nl_eff = collect ? nl : 0u;
        const unsigned nmax = __reduce_max_sync(0xffffffffu, nl_eff);
        for (unsigned t0 = 0; t0 < nmax; t0 += 8) {
          unsigned long long kb[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) kb[u] = (t0 + u < nl_eff) ? listT[(size_t)(t0 + u) * 32 + lane] : 0ull;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (t0 + u < nl_eff) {
              const unsigned long long key = ordered_key(__longlong_as_double((long long)kb[u]));
              listT[(size_t)(t0 + u) * 32 + lane] = key;
              hist[(unsigned)((key >> shift2) & ((1u << width2) - 1u)) * 32 + lane] += 1;
            }
          }
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) s_part[wj + v * NWW][3][lane] = chiv[v];
      s_nl[wj][lane] = collect ? nl : 0u;
      walker_bar<NWW>();
      // ---- second and third level: every warp locates the sub-bin in the summed histograms of its level, then filters
      // its own list into a small candidate buffer
      // [NWW][CAND][32], on the ring: every walker warp is past the chi-square pass, its chunks are consumed
      static_assert(NWW * CAND * 32 * 8 <= L::ring, "the candidate buffers fit the ring");
      unsigned long long* const cand_all = reinterpret_cast<unsigned long long*>(gsm);
      unsigned long long* const cand = cand_all + (size_t)wj * CAND * 32;
      LaneSel S2 = S;
      unsigned nc = 0;
      bool overflow = !collect;
      {
        unsigned bin, below, in_bin;
        lane_merge<NWW>(hist_all, wj, lane);
        walker_bar<NWW>();
        lane_locate(hist_all, lane, 0u, S.rank, bin, below, in_bin);
        S2.rank = S.rank - below; S2.cnt = in_bin;
        S2.prefix = S.prefix | (((unsigned long long)bin) << shift2);
        S2.bits = S.bits + width2;
        const unsigned nl_eff = collect ? nl : 0u;
        const unsigned nmax = __reduce_max_sync(0xffffffffu, nl_eff);
        {
          // ---- third level: the same once more on the next key bits (a spectrum's values crowd around the continuum,
          // the median's second-level bin may still hold dozens of keys; after this one it holds one or two)
          const unsigned long long hmb = (S2.bits >= 64) ? ~0ull : (~0ull << (64 - S2.bits));
          const int width3 = min(HBITS, 64 - S2.bits);
          const int shift3 = max(64 - S2.bits - width3, 0);
          walker_bar<NWW>();                                               // every warp has read the second-level histograms
          lane_hist_zero(hist, lane);
          for (unsigned t0 = 0; t0 < nmax; t0 += 8) {
            unsigned long long kb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) kb[u] = (t0 + u < nl_eff) ? listT[(size_t)(t0 + u) * 32 + lane] : 0ull;
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (t0 + u < nl_eff && (kb[u] & hmb) == S2.prefix)
                hist[(unsigned)((kb[u] >> shift3) & ((1u << width3) - 1u)) * 32 + lane] += 1;
          }
          walker_bar<NWW>();
          lane_merge<NWW>(hist_all, wj, lane);
          walker_bar<NWW>();
          lane_locate(hist_all, lane, 0u, S2.rank, bin, below, in_bin);
          S2.rank -= below; S2.cnt = in_bin;
          S2.prefix |= ((unsigned long long)bin) << shift3;
          S2.bits += width3;
        }
        const unsigned long long hm2 = (S2.bits >= 64) ? ~0ull : (~0ull << (64 - S2.bits));
        for (unsigned t0 = 0; t0 < nmax; t0 += 8) {                        // eight loads in flight
          unsigned long long kb[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) kb[u] = (t0 + u < nl_eff) ? listT[(size_t)(t0 + u) * 32 + lane] : 0ull;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (t0 + u < nl_eff) {
              const unsigned long long k = kb[u];
              if ((k & hm2) == S2.prefix) {
                if (nc < (unsigned)CAND) cand[(size_t)nc * 32 + lane] = k; else overflow = true;
                ++nc;
              } else if (k > S2.prefix) {
                above = min(above, k);
              }
            }
          }
        }
      }
      s_above[wj][lane] = above; s_nc[wj][lane] = overflow ? 0xffffffffu : nc;
      __threadfence_block();
      GTIME(3)
      walker_bar<NWW>();
      GTIME(4)
      if (wj == 0) {
        // ---- the first walker warp finishes: merge, rank the candidates, final sum
        double chi = s_part[0][3][lane];
#pragma unroll
        for (int j = 1; j < VW; ++j) chi += s_part[j][3][lane];
#pragma unroll
        for (int j = 1; j < NWW; ++j) above = min(above, s_above[j][lane]);
        bool any_overflow = false;
#pragma unroll
        for (int j = 0; j < NWW; ++j) any_overflow = any_overflow || s_nc[j][lane] == 0xffffffffu;
        double v1, v2;
        if (!__any_sync(0xffffffffu, any_overflow)) {
          // at most NWW * CAND keys carry the prefix: rank them by counting (ties by position)
          unsigned long long r1 = 0ull, r2 = above;
          unsigned total = 0;
#pragma unroll
          for (int j = 0; j < NWW; ++j) total += s_nc[j][lane];
          const unsigned tmax = __reduce_max_sync(0xffffffffu, total);
          // candidate t of this lane: (warp ja, index ia) in the concatenation of the four buffers
          for (unsigned t = 0; t < tmax; ++t) {
            if (t < total) {
              unsigned rem = t; int ja = 0;
#pragma unroll
              for (int j = 0; j < NWW - 1; ++j) { const unsigned cj = s_nc[j][lane]; if (ja == j && rem >= cj) { rem -= cj; ja = j + 1; } }
              const unsigned long long mine = cand_all[((size_t)ja * CAND + rem) * 32 + lane];
              unsigned below = 0, pos = 0;
#pragma unroll
              for (int j = 0; j < NWW; ++j) {
                const unsigned cj = s_nc[j][lane];
                for (unsigned i = 0; i < cj; ++i, ++pos) {
                  const unsigned long long o = cand_all[((size_t)j * CAND + i) * 32 + lane];
                  below += (o < mine) || (o == mine && pos < t);
                }
              }
              if (below == S2.rank) r1 = mine;
              if (below == S2.rank + 1) r2 = mine;
            }
          }
          v1 = key_value(r1); v2 = key_value(r2);
        } else {
          // ---- slow path (a lane's second-level bin holds more keys than the buffers take, or all its values are
          // equal): the keys with the second-level prefix of all four lists go to a spare list, then digit by digit
          unsigned long long* const spare = list0 + (size_t)NWW * LISTCAP * 32;
          unsigned cnt = 0u;
          const unsigned long long hm2 = (S2.bits >= 64) ? ~0ull : (~0ull << (64 - S2.bits));
          for (int j = 0; j < NWW; ++j) {
            const unsigned long long* lj = list0 + (size_t)j * LISTCAP * 32;
            const unsigned nj = collect ? s_nl[j][lane] : 0u;
            const unsigned nmax = __reduce_max_sync(0xffffffffu, nj);
            for (unsigned t0 = 0; t0 < nmax; t0 += 8) {
              unsigned long long kb[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) kb[u] = (t0 + u < nj) ? lj[(size_t)(t0 + u) * 32 + lane] : 0ull;
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (t0 + u < nj && (kb[u] & hm2) == S2.prefix) { spare[(size_t)cnt * 32 + lane] = kb[u]; ++cnt; }
            }
          }
          __syncwarp();
          if (collect) { S = S2; S.above = above; }                          // `above` already has the keys dropped so far
          else S.above = above;
          unsigned long long* const listT = spare;
          if (S.bits >= 64) cnt = 0u;
          while (__any_sync(0xffffffffu, cnt > 8u && S.bits < 64)) {
            const bool go = cnt > 8u && S.bits < 64;
            LaneSel S3 = S;
            if (!go) S3.bits = 64;                                          // trip count only
            lane_narrow<false>(listT, go ? cnt : 0u, S3, hist, lane);
            const unsigned out = lane_compact(listT, go ? cnt : 0u, S3, lane);
            if (go) { S = S3; cnt = out; }
          }
          if (S.bits >= 64) {
            v1 = key_value(S.prefix);                                       // S.cnt copies of one value
            v2 = (S.rank + 1 < S.cnt) ? v1 : key_value(S.above);
          } else {
            unsigned long long r1 = 0ull, r2 = S.above;
            for (unsigned t = 0; t < cnt; ++t) {
              const unsigned long long mine = listT[(size_t)t * 32 + lane];
              unsigned b2 = 0;
              for (unsigned j = 0; j < cnt; ++j) {
                const unsigned long long o = listT[(size_t)j * 32 + lane];
                b2 += (o < mine) || (o == mine && j < t);
              }
              if (b2 == S.rank) r1 = mine;
              if (b2 == S.rank + 1) r2 = mine;
            }
            v1 = key_value(r1); v2 = key_value(r2);
          }
        }
        GTIME(8)
        const double med = has_nan ? CUDART_NAN : ((n & 1) ? v1 : (v1 + v2) / 2.0);   // np.median, mft6.py:1173
        const double chi_un = has_nan ? CUDART_NAN : chi;
        const double scale = T.data_median / med;
        if (parts.model && live) for (int i = 0; i < n; ++i) parts.model[(long long)w * n + i] = sMT[(size_t)i * 32 + lane] * scale;

        // ---- final sum (mft6.py:1178-1205), every lane for its own walker
        const double iic = (scale * scale) * chi_un / (double)n;            // mft6.py:1173-1179
        double icontrast = 0.0, iphot = 0.0;
        for (int f = 0; f < T.n_contrast; ++f) {
          double fl[NS];
#pragma unroll
          for (int stx = 0; stx < NS; ++stx) {
            double a = 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) a = fma(cf[stx * NR + r], s_band[f * NROW + stx * NR + r], a);
            fl[stx] = a;
          }
          const double m0 = __dmul_rn(-2.5, log10(fl[0])), m1 = __dmul_rn(-2.5, log10(fl[1]));
          const double con = __dsub_rn(m1, m0);                              // mft6.py:741
          const double d = con - T.c_mag[f];
          icontrast += (d * d) / (T.c_err[f] * T.c_err[f]);
          if (live && parts.contrast) parts.contrast[w * T.n_contrast + f] = con;
          if (live && parts.comp_mag) { parts.comp_mag[(w * T.n_contrast + f) * NS] = m0; parts.comp_mag[(w * T.n_contrast + f) * NS + 1] = m1; }
        }
        for (int f = 0; f < T.n_phot; ++f) {
          double tot = 0.0;
#pragma unroll
          for (int stx = 0; stx < NS; ++stx) {
            double a = 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) a = fma(cf[stx * NR + r], s_band[(MS_MAX_CONTRAST + f) * NROW + stx * NR + r], a);
            tot += a;
          }
          double mag = __dmul_rn(-2.5, log10(__dmul_rn(tot, T.p_norm[f])));   // mft6.py:780-782
          mag = __dadd_rn(mag, __dmul_rn(av, T.p_ext[f]));                  // mft6.py:1163 (av == 0 when skipped)
          const double d = mag - T.p_mag[f];
          iphot += (d * d) / (T.p_err[f] * T.p_err[f]);
          if (live && parts.phot) parts.phot[w * T.n_phot + f] = mag;
        }
        if (live) {
          const WalkerPlan& p = plans[w];
          const double init_cs = __dadd_rn(__dadd_rn(__dmul_rn(iic, (double)(T.n_contrast + T.n_phot)), icontrast), iphot);   // mft6.py:1191
          const double lnl = (init_cs != init_cs) ? -CUDART_INF : -0.5 * init_cs;                // mft6.py:1202-1205
          double out = p.lp + lnl;
          if (out != out) out = -CUDART_INF;
          lnprob[w] = out;
          if (parts.status) parts.status[w] = (init_cs != init_cs) ? MS_W_NONFINITE : 0;
          if (parts.lnprior) parts.lnprior[w] = p.lp;
          if (parts.chi2_spec) parts.chi2_spec[w] = iic;
          if (parts.logg) for (int s = 0; s < NS; ++s) parts.logg[w * NS + s] = p.logg[s];
        }
      }
      GTIME(5)
    }
    cit0 += (unsigned)n_chunks;
  }
#ifdef MS_GROUP_TIMING
  if (blockIdx.x < 4 && lane == 0)
    printf("cta %d warp %d groups %d: idle %lld A %lld sync %lld B: locate %lld chi %lld lists %lld bar %lld rank %lld finish %lld | mbarrier waits A %lld B %lld\n", blockIdx.x, warp,
           gt_groups, gt_acc[0], gt_acc[1], gt_acc[2], gt_acc[6], gt_acc[7], gt_acc[3], gt_acc[4], gt_acc[8], gt_acc[5], gt_wait_a, gt_wait);
#endif
}

// ---- group table: contiguous runs of at most GMAX cell-sorted walkers that share a cell --------------------
// cell_start: exclusive prefix of the cell histogram (n_cells + 1 entries); group_off: exclusive prefix of
// ceil(count / GMAX).  One thread per walker (after the counting sort has placed it): the first walker of every
// group writes the descriptor.
__global__ void group_table_kernel(const WalkerPlan* __restrict__ plans, const int* __restrict__ order, long long nw,
                                   const unsigned* __restrict__ cell_start, const unsigned* __restrict__ group_off,
                                   GroupDesc* __restrict__ groups) {
  const long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pos >= nw) return;
  const int cell = plans[order[pos]].cell;
  const unsigned local = (unsigned)pos - cell_start[cell];
  if (local % GMAX == 0) {
    const unsigned cnt = cell_start[cell + 1] - cell_start[cell];
    GroupDesc g; g.start = (int)pos; g.count = (int)min((unsigned)GMAX, cnt - local);
    groups[group_off[cell] + local / GMAX] = g;
  }
}

// Exclusive scans of the cell histogram (-> cell_start[n_cells + 1] and the mutable copy `offsets` the scatter
// consumes) and of the per-cell group counts (-> group_off[n_cells], *n_groups).  One CTA; each warp scans 256
// consecutive cells per pass (coalesced rows of 32), then the warps' totals are scanned.
__global__ void __launch_bounds__(1024) group_scan_kernel(unsigned* __restrict__ hist_offsets, int n_cells,
                                                          unsigned* __restrict__ cell_start, unsigned* __restrict__ group_off,
                                                          int* __restrict__ n_groups, int* __restrict__ group_counter) {
  __shared__ unsigned wa[32], wb[32];
  __shared__ unsigned carry_a, carry_b;
  constexpr int ROWS = 8;                        // 32-cell rows per warp and pass: 256 consecutive cells per warp, 8192 per pass
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_a = 0; carry_b = 0; }
  __syncthreads();
  for (int base = 0; base < n_cells; base += 1024 * ROWS) {
    const int w0 = base + warp * 32 * ROWS;
    unsigned v[ROWS], ea[ROWS], eb[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {               // coalesced, all in flight
      const int i = w0 + k * 32 + lane;
      v[k] = (i < n_cells) ? hist_offsets[i] : 0u;
    }
    unsigned run_a = 0, run_b = 0;                  // the warp's running totals
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const unsigned g = (v[k] + GMAX - 1) / GMAX;
      unsigned ia = v[k], ib = g;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb; }
      }
      ea[k] = run_a + ia - v[k]; eb[k] = run_b + ib - g;          // exclusive, within the warp's 256 cells
      run_a += __shfl_sync(0xffffffffu, ia, 31); run_b += __shfl_sync(0xffffffffu, ib, 31);
    }
    if (lane == 31) { wa[warp] = run_a; wb[warp] = run_b; }
    __syncthreads();
    if (warp == 0) {
      unsigned xa = wa[lane], xb = wb[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned ta = __shfl_up_sync(0xffffffffu, xa, o), tb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= o) { xa += ta; xb += tb; }
      }
      wa[lane] = xa; wb[lane] = xb;                 // inclusive over the warps
    }
    __syncthreads();
    const unsigned off_a = carry_a + (warp ? wa[warp - 1] : 0u), off_b = carry_b + (warp ? wb[warp - 1] : 0u);
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const int i = w0 + k * 32 + lane;
      if (i < n_cells) { hist_offsets[i] = off_a + ea[k]; cell_start[i] = off_a + ea[k]; group_off[i] = off_b + eb[k]; }
    }
    __syncthreads();
    if (tid == 0) { carry_a += wa[31]; carry_b += wb[31]; }
    __syncthreads();
  }
  if (tid == 0) { cell_start[n_cells] = carry_a; *n_groups = (int)carry_b; *group_counter = 0; }
}

}  // namespace grp
