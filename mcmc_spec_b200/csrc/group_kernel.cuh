// group_kernel.cuh -- lnprob_group_kernel: the large-batch form of the streaming evaluation (included by
// mcmcspec.cu; translation unit MS_TU = 4).
//
// One CTA evaluates a *group* of up to GW walkers that sit in the same grid cell, i.e. whose 2 * 2^D bracketing
// grid rows (mft6.py:439-519) are the same rows.  The rows are streamed ONCE per group:
//
//   warp 0      producer   -- per 128-pixel stage one elected set of lanes issues 1-D bulk tensor copies
//                             (cp.async.bulk ... mbarrier::complete_tx: SASS UBLKCP) of the bracketing row segments,
//                             of the weight tiles of the passbands that touch the stage and of the per-observed-pixel
//                             records into a 4-stage shared-memory ring; `full` / `empty` mbarriers per stage.
//   warp 1      gather     -- picks, for every observed pixel whose upper bracketing model pixel lies in the stage,
//                             the two bracketing pixels of every row out of the staged segments (mft6.py:1169:
//                             interp1d onto the observed wavelengths) into a compact table B[row][pixel] that all
//                             walkers of the group read coalesced; `bready` mbarrier.
//   warps 2,3   band       -- integrate every staged ROW through the passbands that touch the stage
//                             (sum_p w_f[p] * row[p]: np.trapz of mft6.py:731 resp. pyphot get_flux, mft6.py:776, which
//                             are linear in the spectrum); a walker's band fluxes are the same linear combination of
//                             these row integrals as its spectrum is of the rows (mft6.py:507-519, :689-700).
//   warps 4..   walkers    -- one warp per walker: forms the composite at the two bracketing pixels of each observed
//                             pixel (FMA chain over the rows, then the sum over the stars, mft6.py:744), extincts both
//                             (mft6.py:1162), rebins (mft6.py:1169), accumulates norm_spec's moments (mft6.py:193-196)
//                             and the key range for the median, and writes the model vector to its scratch row.
//                             After the stream the same warp finishes alone: exact median by MSD radix select fused
//                             with the chi-square pass (mft6.py:1173-1179), contrasts, magnitudes, priors, final sum
//                             (mft6.py:1182-1205).  No block-wide barrier inside either phase.
//
// Consumer warps never touch global memory for grid rows, weights or records (LDS only), so their stalls are
// shared-memory latency (~30 cycles) instead of L2 latency (~800), and L2 -> SM traffic drops by the group size.
// Extinction: exp2(Av * ext(p)) = A_s * exp(Av * x_p) with A_s = exp2(Av * anchor_s) computed once per stage and
// walker (32 stages at a time, one per lane) and a 7th-order polynomial on |Av * x_p| < 0.02 (x^8/8! < 7e-19);
// larger exponents take the exp2 path.

namespace grp {

constexpr int TP = 128;                    // model pixels per stage
constexpr int ROWPX = TP + 2;              // staged pixels per row: the tile and the two pixels before it (16-byte
                                           // alignment of the bulk copy; the second one is the lower bracket of an
                                           // observed pixel whose upper bracket is the first pixel of the tile)
constexpr int ROWB = ROWPX * 8;            // 1040 bytes = 16 (mod 128): the 8 rows of one gather land in distinct banks
constexpr int WS = MS_GROUP_WSLOTS;        // weight slots: passbands that may overlap on one stage
constexpr int RCAP = MS_GROUP_RCAP;        // observed pixels per stage the ring is sized for
constexpr int NSTAGE = 4;
constexpr int GW = MS_GROUP_WALKERS;       // walkers (= walker warps) per CTA
constexpr int NSERVICE = 4;                // producer, gather, 2 x band
constexpr int NTHREADS = (GW + NSERVICE) * 32;
constexpr int BROWB = RCAP * 16 + 16;      // stride of a B row (one pad slot: conflict-free stores from 8 rows)
constexpr int LIST_MAX = 1024;             // keys of the median bin a warp keeps in shared memory
constexpr double POLY_XMAX = 0.02;

template <int NROW> struct StageLayout {
  static constexpr int rows = 0;
  static constexpr int weights = NROW * ROWB;
  static constexpr int recA = weights + WS * TP * 8;       // double2 {t, x_lo}
  static constexpr int recB = recA + RCAP * 16;            // double2 {x_hi, d q0}
  static constexpr int recC = recB + RCAP * 16;            // double2 {d q1, d q2}
  static constexpr int hl = recC + RCAP * 16;              // int32[RCAP + 4]: upper bracket, tile-local
  static constexpr int meta = hl + (RCAP + 4) * 4;
  static constexpr int B = meta + 16;
  static constexpr int bytes = B + NROW * BROWB;
  static_assert(weights % 16 == 0 && meta % 16 == 0 && B % 16 == 0 && bytes % 16 == 0, "bulk copies need 16-byte alignment");
};

// What the group kernel needs beyond DevTables (built once in ms_create).
struct GroupTables {
  const int4* stage_meta;       // [n_stages] {d0, d1, 4 x uint8 band column per weight slot (0xff: free), flush mask}
  const double2* recA; const double2* recB; const double2* recC;   // [n_data (+pad)]
  const int* hl;                // [n_data (+pad)]
  const double* anchor;         // [n_stages + 32] extinction exponent the stage's x values are relative to
  const double* drec2;          // [n_data][4] {q1(x), q2(x), flux, 1/err^2}: phase B
  const double* wbase[BAND_COLS];   // weight of region pixel p for band column c: wbase[c][p]
  double xmax;                  // max |x| over all records (x = (ext - anchor) ln 2)
  int n_stages;
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
// 1-D bulk tensor copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

// exp(y) for |y| <= POLY_XMAX: Taylor to y^7 (remainder y^8/40320 < 7e-19 relative)
__device__ __forceinline__ double exp_small(double y) {
  double p = fma(y, 1.0 / 5040.0, 1.0 / 720.0);
  p = fma(p, y, 1.0 / 120.0); p = fma(p, y, 1.0 / 24.0); p = fma(p, y, 1.0 / 6.0);
  p = fma(p, y, 0.5); p = fma(p, y, 1.0); p = fma(p, y, 1.0);
  return p;
}

// Ordered 32-bit key of the high word of a double (same order as ordered_key() on the full value).
__device__ __forceinline__ unsigned hi_key(double v) {
  const int h = __double2hiint(v);
  return (unsigned)h ^ ((unsigned)(h >> 31) | 0x80000000u);
}

// ---- warp-level exact selection --------------------------------------------------------------
// Locate the bin of `hist` (256 bins in shared memory) that holds rank `rank`; every lane returns the same
// (bin, elements below the bin, elements in the bin).
__device__ __forceinline__ void warp_locate_bin(const unsigned* hist, unsigned rank, int lane, unsigned& bin, unsigned& below,
                                                unsigned& in_bin) {
  unsigned c[8]; unsigned tot = 0;
  const uint4 a = *reinterpret_cast<const uint4*>(hist + lane * 8);
  const uint4 b = *reinterpret_cast<const uint4*>(hist + lane * 8 + 4);
  c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
#pragma unroll
  for (int j = 0; j < 8; ++j) tot += c[j];
  unsigned incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  const unsigned excl = incl - tot;
  const bool mine = rank >= excl && rank < incl;
  unsigned mb = 0, mbelow = 0, mcnt = 0;
  if (mine) {
    unsigned run = excl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (rank >= run && rank < run + c[j]) { mb = (unsigned)(lane * 8 + j); mbelow = run; mcnt = c[j]; }
      run += c[j];
    }
  }
  const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
  bin = __shfl_sync(0xffffffffu, mb, src);
  below = __shfl_sync(0xffffffffu, mbelow, src);
  in_bin = __shfl_sync(0xffffffffu, mcnt, src);
}

__device__ __forceinline__ void warp_hist_add(unsigned* hist, unsigned digit, int lane) {
  // warp-aggregated shared-memory atomics (neighbouring pixels of a smooth spectrum share a bin); digit
  // 0xffffffff = no element
  const unsigned peers = __match_any_sync(0xffffffffu, digit);
  if (digit != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(hist + digit, (unsigned)__popc(peers));
}

// Median of sM[0..n) (np.median convention) by one warp; also returns through *chi_out
// sum_i (m_i - d_i / p(x_i))^2 / e_i^2 with p = a0 q0 + a1 q1 + a2 q2 (mft6.py:196, :120 on the unscaled model).
// hist: 256 unsigned, list: LIST_MAX keys, both warp-private shared memory.  [kmin32, kmax32]: range of hi_key().
__device__ double warp_median_chi(const double* __restrict__ sM, int n, unsigned* hist, unsigned long long* list,
                                  unsigned kmin32, unsigned kmax32, const double* __restrict__ drec2, double a0q0, double a1,
                                  double a2, double* chi_out, int lane) {
  const unsigned long long kmin = (unsigned long long)kmin32 << 32, kmax = ((unsigned long long)kmax32 << 32) | 0xffffffffull;
  unsigned rank = (unsigned)((n - 1) >> 1);
  int bits_done = __clzll((long long)(kmin ^ kmax));                 // < 64: the low words differ
  unsigned long long prefix = (bits_done == 0) ? 0ull : (kmin & (~0ull << (64 - bits_done)));
  int width = min(8, 64 - bits_done);
  int shift = 64 - bits_done - width;

  // ---- pass 1: chi-square terms + histogram of the first differing key bits
#pragma unroll
  for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0u;
  __syncwarp();
  double chi = 0.0;
  constexpr int U = 4;
  for (int i0 = 0; i0 < n; i0 += 32 * U) {
    double m[U]; dbl4 rec[U]; bool v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 32 + lane;
      v[u] = i < n;
      m[u] = 0.0; rec[u].x = rec[u].y = rec[u].z = rec[u].w = 0.0;
      if (v[u]) { m[u] = sM[i]; rec[u] = ldg_cached4(drec2 + 4 * (long long)i); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned dg = 0xffffffffu;
      if (v[u]) {
        const double pfit = fma(a2, rec[u].y, fma(a1, rec[u].x, a0q0));
        const double res = m[u] - rec[u].z * fast_rcp(pfit);
        chi = fma(res * res, rec[u].w, chi);
        dg = (unsigned)((ordered_key(m[u]) >> shift) & ((1u << width) - 1u));
      }
      warp_hist_add(hist, dg, lane);
    }
  }
  *chi_out = warp_sum(chi);
  __syncwarp();

  // ---- narrow down on the full vector until the bin fits the list
  unsigned bin, below, in_bin;
  while (true) {
    warp_locate_bin(hist, rank, lane, bin, below, in_bin);
    rank -= below;
    prefix |= ((unsigned long long)bin) << shift;
    bits_done += width;
    if (in_bin <= (unsigned)LIST_MAX || bits_done >= 64) break;
    width = min(8, 64 - bits_done);
    shift = 64 - bits_done - width;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0u;
    __syncwarp();
    const unsigned long long himask = ~0ull << (64 - bits_done);
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      unsigned dg = 0xffffffffu;
      if (i < n) {
        const unsigned long long k = ordered_key(sM[i]);
        if ((k & himask) == prefix) dg = (unsigned)((k >> shift) & ((1u << width) - 1u));
      }
      warp_hist_add(hist, dg, lane);
    }
    __syncwarp();
  }

  double v1, v2;
  bool need_next;
  if (bits_done >= 64) {
    v1 = key_value(prefix); v2 = v1;                                 // in_bin copies of one value
    need_next = (rank + 1 >= in_bin);
  } else {
    // ---- compact the bin's keys (in index order), then keep narrowing on the list
    const unsigned long long himask = ~0ull << (64 - bits_done);
    unsigned cnt = 0;
    for (int i0 = 0; i0 < n; i0 += 32 * U) {
      unsigned long long k[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 32 + lane;
        k[u] = (i < n) ? ordered_key(sM[i]) : ~prefix;               // ~prefix never matches
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool hit = (k[u] & himask) == prefix;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) list[cnt + __popc(bal & ((1u << lane) - 1u))] = k[u];
        cnt += __popc(bal);
      }
    }
    __syncwarp();
    while (cnt > 64u && bits_done < 64) {
      width = min(8, 64 - bits_done);
      shift = 64 - bits_done - width;
#pragma unroll
      for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0u;
      __syncwarp();
      for (unsigned t0 = 0; t0 < cnt; t0 += 32) {
        const unsigned t = t0 + lane;
        warp_hist_add(hist, t < cnt ? (unsigned)((list[t] >> shift) & ((1u << width) - 1u)) : 0xffffffffu, lane);
      }
      __syncwarp();
      warp_locate_bin(hist, rank, lane, bin, below, in_bin);
      rank -= below;
      prefix |= ((unsigned long long)bin) << shift;
      bits_done += width;
      const unsigned long long hm = (bits_done >= 64) ? ~0ull : (~0ull << (64 - bits_done));
      unsigned out = 0;                                              // in-place compaction: out <= t0 always
      for (unsigned t0 = 0; t0 < cnt; t0 += 32) {
        const unsigned t = t0 + lane;
        const unsigned long long k = t < cnt ? list[t] : ~prefix;
        const bool hit = t < cnt && (k & hm) == prefix;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        __syncwarp();
        if (hit) list[out + __popc(bal & ((1u << lane) - 1u))] = k;
        out += __popc(bal);
        __syncwarp();
      }
      cnt = out;
    }
    // ---- rank the remaining keys by counting
    unsigned long long r1 = 0ull, r2 = 0ull;
    bool f1 = false, f2 = false;
    for (unsigned t = lane; t < cnt; t += 32) {
      const unsigned long long mine = list[t];
      unsigned b = 0;
      for (unsigned j = 0; j < cnt; ++j) {
        const unsigned long long o = list[j];
        b += (o < mine) || (o == mine && j < t);
      }
      if (b == rank) { r1 = mine; f1 = true; }
      if (b == rank + 1) { r2 = mine; f2 = true; }
    }
    const unsigned b1 = __ballot_sync(0xffffffffu, f1), b2 = __ballot_sync(0xffffffffu, f2);
    r1 = __shfl_sync(0xffffffffu, r1, __ffs(b1) - 1);
    need_next = (b2 == 0u);
    r2 = __shfl_sync(0xffffffffu, r2, need_next ? 0 : __ffs(b2) - 1);
    v1 = key_value(r1); v2 = key_value(r2);
  }
  if (n & 1) return v1;
  if (need_next) {
    // rare: the lower median is the largest element of its bin -> upper median = min{key > key(v1)}
    const unsigned long long k1 = ordered_key(v1);
    unsigned long long kgt = ~0ull;
    for (int i = lane; i < n; i += 32) {
      const unsigned long long k = ordered_key(sM[i]);
      if (k > k1) kgt = min(kgt, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kgt = min(kgt, __shfl_xor_sync(0xffffffffu, kgt, o));
    v2 = key_value(kgt);
  }
  return (v1 + v2) / 2.0;
}

// ------------------------------------------------------------------------------------------
// NR: grid rows per star (4 or 8).  Binaries, posterior mode.
// ------------------------------------------------------------------------------------------
template <int NR>
__global__ void __launch_bounds__(NTHREADS, (NR == 4) ? 2 : 1)
lnprob_group_kernel(const __grid_constant__ DevTables T, const __grid_constant__ GroupTables GT,
                    const WalkerPlan* __restrict__ plans, const int* __restrict__ order,
                    const GroupDesc* __restrict__ groups, const int* __restrict__ n_groups,
                    int* __restrict__ group_counter, double* __restrict__ lnprob, const PartsOut parts,
                    double* __restrict__ scratch) {
  constexpr int NS = 2;
  constexpr int NROW = NS * NR;
  using L = StageLayout<NROW>;
  extern __shared__ __align__(128) unsigned char gsm[];
  __shared__ __align__(8) unsigned long long s_full[NSTAGE], s_empty[NSTAGE], s_bready[NSTAGE];
  __shared__ double s_band[BAND_COLS * NROW];            // row integrals through every band
  __shared__ double s_cf[GW][NROW];
  __shared__ double s_av[GW], s_lp[GW], s_logg[GW][NS];
  __shared__ int s_w[GW];
  __shared__ int s_row[NROW];
  __shared__ int s_group, s_status;
  static_assert(GW * (1024 + LIST_MAX * 8) <= NSTAGE * L::bytes, "phase-B scratch aliases the ring");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 32) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), (unsigned)(GW + 3));        // every walker warp + gather + 2 band warps
      mbar_init(smem_u32(&s_bready[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int n_stages = GT.n_stages;
  const int n_data = T.n_data;
  const int total_groups = *n_groups;
  const uint32_t ring = smem_u32(gsm);
  // this CTA's model vectors: one scratch row per walker warp, reused by every group the CTA takes
  double* const sM = scratch + ((size_t)blockIdx.x * GW + (size_t)max(warp - NSERVICE, 0)) * (size_t)n_data;
  unsigned it0 = 0;                         // stages this CTA has streamed so far: ring slot and mbarrier parity go on

  // Persistent CTA: groups are handed out by an atomic counter (reset by group_scan_kernel).
  while (true) {
    if (tid == 0) s_group = atomicAdd(group_counter, 1);
    __syncthreads();                                                // also: the previous group's phase B is over
    const int g = s_group;
    if (g >= total_groups) break;
    const GroupDesc gd = groups[g];
    const int count = gd.count;

    // ---- group prologue: plans of the members, rows of the cell
    if (tid < GW) {
      const bool live = tid < count;
      const int w = live ? order[gd.start + tid] : -1;
      s_w[tid] = w;
      if (live) {
        const WalkerPlan& p = plans[w];
#pragma unroll
        for (int st = 0; st < NS; ++st) {
#pragma unroll
          for (int r = 0; r < NR; ++r) s_cf[tid][st * NR + r] = p.coef[st][r];
          s_logg[tid][st] = p.logg[st];
        }
        s_av[tid] = p.av; s_lp[tid] = p.lp;
        if (tid == 0) {
          s_status = p.status;
#pragma unroll
          for (int st = 0; st < NS; ++st) {
#pragma unroll
            for (int r = 0; r < NR; ++r) s_row[st * NR + r] = p.row[st][r];
          }
        }
      }
    }
    for (int k = tid; k < BAND_COLS * NROW; k += NTHREADS) s_band[k] = 0.0;
    __syncthreads();

    if (s_status != 0) {
      // the walkers that never reach the grid are sorted into their own cell: nothing to stream (mft6.py:1465-1466)
      if (tid < count) {
        const int w = s_w[tid];
        const WalkerPlan& p = plans[w];
        lnprob[w] = -CUDART_INF;
        if (parts.status) parts.status[w] = p.status;
        if (parts.lnprior) parts.lnprior[w] = -CUDART_INF;
        if (parts.chi2_spec) parts.chi2_spec[w] = CUDART_NAN;
        if (parts.logg) for (int s = 0; s < NS; ++s) parts.logg[w * NS + s] = p.logg[s];
        if (parts.contrast) for (int c = 0; c < T.n_contrast; ++c) parts.contrast[w * T.n_contrast + c] = CUDART_NAN;
        if (parts.phot) for (int f = 0; f < T.n_phot; ++f) parts.phot[w * T.n_phot + f] = CUDART_NAN;
        if (parts.comp_mag) for (int f = 0; f < T.n_contrast * NS; ++f) parts.comp_mag[w * T.n_contrast * NS + f] = CUDART_NAN;
      }
      if (parts.model)
        for (int k = 0; k < count; ++k)
          for (int i = tid; i < n_data; i += NTHREADS) parts.model[(long long)s_w[k] * n_data + i] = CUDART_NAN;
      continue;
    }

    // walker-warp state that survives the stream
    double mom0 = 0.0, mom1 = 0.0, mom2 = 0.0;
    unsigned kmin = 0xffffffffu, kmax = 0u;

    if (warp == 0) {
      // ================================================================= producer
      const double* rowp = nullptr;
      if (lane < NROW) rowp = T.grid + (long long)s_row[lane] * T.stride - 2;    // tile pixel 0 = region pixel px0 - 2
      for (int s = 0; s < n_stages; ++s) {
        const unsigned gs = it0 + (unsigned)s;
        const int slot = (int)(gs & (NSTAGE - 1));
        const unsigned use = gs / NSTAGE;
        mbar_wait(smem_u32(&s_empty[slot]), (use & 1u) ^ 1u);
        const int4 meta = __ldg(GT.stage_meta + s);
        const int n_obs = meta.y - meta.x;
        const uint32_t st = ring + (uint32_t)slot * L::bytes;
        const uint32_t bar = smem_u32(&s_full[slot]);
        // bytes of this stage (uniform): rows, active weight tiles, records
        const int hl_lo = meta.x & ~3;
        const int hl_n = n_obs > 0 ? ((meta.y - hl_lo + 3) & ~3) : 0;
        int n_w = 0;
#pragma unroll
        for (int q = 0; q < WS; ++q) n_w += (((unsigned)meta.z >> (8 * q)) & 0xffu) != 0xffu;
        const unsigned total = (unsigned)(NROW * ROWB + n_w * TP * 8 + 3 * n_obs * 16 + hl_n * 4);
        if (lane == 0) {
          *reinterpret_cast<int4*>(gsm + (size_t)slot * L::bytes + L::meta) = meta;
          mbar_expect_tx(bar, total);
        }
        __syncwarp();
        const long long px0 = (long long)s * TP;
        if (lane < NROW) {
          bulk_g2s(st + L::rows + lane * ROWB, rowp + px0, ROWB, bar);
        } else if (lane < NROW + WS) {
          const int q = lane - NROW;
          const unsigned col = ((unsigned)meta.z >> (8 * q)) & 0xffu;
          if (col != 0xffu) bulk_g2s(st + L::weights + q * TP * 8, GT.wbase[col] + px0, TP * 8, bar);
        } else if (n_obs > 0) {
          const int j = lane - NROW - WS;
          if (j == 0) bulk_g2s(st + L::recA, GT.recA + meta.x, n_obs * 16, bar);
          else if (j == 1) bulk_g2s(st + L::recB, GT.recB + meta.x, n_obs * 16, bar);
          else if (j == 2) bulk_g2s(st + L::recC, GT.recC + meta.x, n_obs * 16, bar);
          else if (j == 3) bulk_g2s(st + L::hl, GT.hl + hl_lo, hl_n * 4, bar);
        }
      }
    } else if (warp == 1) {
      // ================================================================= gather
      for (int s = 0; s < n_stages; ++s) {
        const unsigned gs = it0 + (unsigned)s;
        const int slot = (int)(gs & (NSTAGE - 1));
        const unsigned par = (gs / NSTAGE) & 1u;
        mbar_wait(smem_u32(&s_full[slot]), par);
        const uint32_t st = ring + (uint32_t)slot * L::bytes;
        const int4 meta = *reinterpret_cast<const int4*>(gsm + (size_t)slot * L::bytes + L::meta);
        const int n_obs = meta.y - meta.x;
        const int hoff = meta.x & 3;
        const int r = lane & 7;
        for (int k0 = 0; k0 < n_obs; k0 += 4) {
          const int k = k0 + (lane >> 3);
          if (k < n_obs) {
            const int hl = *reinterpret_cast<const int*>(gsm + (size_t)slot * L::bytes + L::hl + 4 * (k + hoff));
            const int j1 = hl + 1;                          // tile index of the lower bracket (upper: hl + 2)
            const int e = j1 & ~1;
#pragma unroll
            for (int rr = 0; rr < NROW; rr += 8) {
              const uint32_t rp = st + L::rows + (uint32_t)(rr + r) * ROWB + (uint32_t)e * 8;
              const double2 v01 = lds128(rp);
              const double v2 = lds64(rp + 16);
              double2 pr;
              pr.x = (j1 & 1) ? v01.y : v01.x;
              pr.y = (j1 & 1) ? v2 : v01.y;
              sts128(st + L::B + (uint32_t)(rr + r) * BROWB + (uint32_t)k * 16, pr);
            }
          }
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(smem_u32(&s_bready[slot])); mbar_arrive(smem_u32(&s_empty[slot])); }
      }
    } else if (warp < NSERVICE) {
      // ================================================================= band integration (rows [r0, r0 + NROW/2))
      constexpr int HR = NROW / 2;
      const int r0 = (warp - 2) * HR;
      double acc[HR][WS];
#pragma unroll
      for (int r = 0; r < HR; ++r) {
#pragma unroll
        for (int q = 0; q < WS; ++q) acc[r][q] = 0.0;
      }
      for (int s = 0; s < n_stages; ++s) {
        const unsigned gs = it0 + (unsigned)s;
        const int slot = (int)(gs & (NSTAGE - 1));
        const unsigned par = (gs / NSTAGE) & 1u;
        mbar_wait(smem_u32(&s_full[slot]), par);
        const uint32_t st = ring + (uint32_t)slot * L::bytes;
        const int4 meta = *reinterpret_cast<const int4*>(gsm + (size_t)slot * L::bytes + L::meta);
        const unsigned slots = (unsigned)meta.z, flush = (unsigned)meta.w;
        if (slots != 0xffffffffu) {
          // this lane's pixels of the tile: {2l, 2l+1} and {64 + 2l, 65 + 2l} (16-byte accesses, lanes contiguous)
#pragma unroll
          for (int q = 0; q < WS; ++q) {
            const unsigned col = (slots >> (8 * q)) & 0xffu;
            if (col != 0xffu) {                                // warp-uniform
              const double2 w0 = lds128(st + L::weights + q * TP * 8 + lane * 16);
              const double2 w1 = lds128(st + L::weights + q * TP * 8 + 512 + lane * 16);
#pragma unroll
              for (int r = 0; r < HR; ++r) {
                const uint32_t rp = st + L::rows + (uint32_t)(r0 + r) * ROWB + 16 + lane * 16;
                const double2 v0 = lds128(rp);
                const double2 v1 = lds128(rp + 512);
                double a = acc[r][q];
                a = fma(w0.x, v0.x, a); a = fma(w0.y, v0.y, a); a = fma(w1.x, v1.x, a); a = fma(w1.y, v1.y, a);
                acc[r][q] = a;
              }
              if ((flush >> q) & 1u) {
#pragma unroll
                for (int r = 0; r < HR; ++r) {
                  const double tot = warp_sum(acc[r][q]);
                  if (lane == 0) s_band[col * NROW + r0 + r] = tot;
                  acc[r][q] = 0.0;
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_empty[slot]));
      }
    } else {
      // ================================================================= walker warp: phase A
      const int j = warp - NSERVICE;
      const bool live = j < count;                         // idle walker warps only keep the ring's arrival counts
      double cf[NROW];
#pragma unroll
      for (int r = 0; r < NROW; ++r) cf[r] = live ? s_cf[j][r] : 0.0;
      const double av = live ? s_av[j] : 0.0;
      const bool poly_ok = av * GT.xmax <= POLY_XMAX;
      double a_lane = 1.0;
      for (int s = 0; s < n_stages; ++s) {
        const unsigned gs = it0 + (unsigned)s;
        const int slot = (int)(gs & (NSTAGE - 1));
        const unsigned par = (gs / NSTAGE) & 1u;
        if (live && (s & 31) == 0) a_lane = exp2(av * __ldg(GT.anchor + s + lane));   // A of stages s .. s+31, one per lane
        mbar_wait(smem_u32(&s_bready[slot]), par);          // B is ready (the gather warp saw `full`)
        mbar_wait(smem_u32(&s_full[slot]), par);            // records: acquire on the barrier the bulk copies signal
        const uint32_t st = ring + (uint32_t)slot * L::bytes;
        const int4 meta = *reinterpret_cast<const int4*>(gsm + (size_t)slot * L::bytes + L::meta);
        const int n_obs = meta.y - meta.x;
        if (live && n_obs > 0) {
          const double A = __shfl_sync(0xffffffffu, a_lane, s & 31);
          for (int k = lane; k < n_obs; k += 32) {
            const double2 ra = lds128(st + L::recA + k * 16);        // t, x_lo
            const double2 rb = lds128(st + L::recB + k * 16);        // x_hi, d q0
            const double2 rc = lds128(st + L::recC + k * 16);        // d q1, d q2
            double flo = 0.0, fhi = 0.0;
#pragma unroll
            for (int stx = 0; stx < NS; ++stx) {
              const double2 b0 = lds128(st + L::B + (uint32_t)(stx * NR) * BROWB + k * 16);
              double l = cf[stx * NR] * b0.x, u = cf[stx * NR] * b0.y;
#pragma unroll
              for (int r = 1; r < NR; ++r) {
                const double2 b = lds128(st + L::B + (uint32_t)(stx * NR + r) * BROWB + k * 16);
                l = fma(cf[stx * NR + r], b.x, l); u = fma(cf[stx * NR + r], b.y, u);
              }
              flo = (stx == 0) ? l : flo + l; fhi = (stx == 0) ? u : fhi + u;    // mft6.py:744
            }
            double elo, ehi, scale_a = A;
            if (poly_ok) {
              elo = exp_small(av * ra.y); ehi = exp_small(av * rb.x);
            } else {
              // exponent too large for the polynomial: exp2 of the whole exponent (1/ln 2 = 1.4426950408889634)
              const double anc = __ldg(GT.anchor + s);
              elo = exp2(av * fma(ra.y, 1.4426950408889634, anc));
              ehi = exp2(av * fma(rb.x, 1.4426950408889634, anc));
              scale_a = 1.0;
            }
            const double ylo = flo * elo, yhi = fhi * ehi;                 // extinct() at both brackets, mft6.py:1162
            const double m = scale_a * fma(yhi - ylo, ra.x, ylo);          // interp1d onto the observed pixel, mft6.py:1169
            const double rm = fast_rcp(m);                                 // norm_spec moments, mft6.py:194-195
            mom0 = fma(rb.y, rm, mom0); mom1 = fma(rc.x, rm, mom1); mom2 = fma(rc.y, rm, mom2);
            const unsigned hk = hi_key(m);
            kmin = min(kmin, hk); kmax = max(kmax, hk);
            sM[meta.x + k] = m;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_empty[slot]));
      }
    }
    it0 += (unsigned)n_stages;

    __syncthreads();          // band integrals complete; the ring is idle and becomes phase-B scratch

    // ================================================================= walker warp: phase B + final sum
    if (warp >= NSERVICE && warp - NSERVICE < count) {
      const int j = warp - NSERVICE;
      const int w = s_w[j];
      const int n = n_data;
      unsigned* hist = reinterpret_cast<unsigned*>(gsm + (size_t)j * (1024 + LIST_MAX * 8));
      unsigned long long* list = reinterpret_cast<unsigned long long*>(gsm + (size_t)j * (1024 + LIST_MAX * 8) + 1024);
      mom0 = warp_sum(mom0); mom1 = warp_sum(mom1); mom2 = warp_sum(mom2);
      kmin = __reduce_min_sync(0xffffffffu, kmin); kmax = __reduce_max_sync(0xffffffffu, kmax);
      __syncwarp();
      bool has_nan = false;
      if (kmax >= 0xfff00000u || kmin <= 0x000fffffu) {
        // an infinity or a NaN is in the vector: look for NaNs exactly (mft6.py:1202)
        int f = 0;
        for (int i = lane; i < n; i += 32) { const double m = sM[i]; f |= (m != m); }
        has_nan = __any_sync(0xffffffffu, f);
      }
      double chi_un = CUDART_NAN, med = CUDART_NAN;
      if (!has_nan) med = warp_median_chi(sM, n, hist, list, kmin, kmax, GT.drec2, mom0 * T.basis[0], mom1, mom2, &chi_un, lane);
      const double scale = T.data_median / med;                            // mft6.py:1173
      if (parts.model) for (int i = lane; i < n; i += 32) parts.model[(long long)w * n + i] = sM[i] * scale;

      if (lane == 0) {
        const double av = s_av[j], lp = s_lp[j];
        double cf[NROW];
#pragma unroll
        for (int r = 0; r < NROW; ++r) cf[r] = s_cf[j][r];
        const double iic = (scale * scale) * chi_un / (double)n;          // mft6.py:1173-1179
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
          const double m0 = -2.5 * log10(fl[0]), m1 = -2.5 * log10(fl[1]);
          const double con = m1 - m0;                                      // mft6.py:741
          const double d = con - T.c_mag[f];
          icontrast += (d * d) / (T.c_err[f] * T.c_err[f]);
          if (parts.contrast) parts.contrast[w * T.n_contrast + f] = con;
          if (parts.comp_mag) { parts.comp_mag[(w * T.n_contrast + f) * NS] = m0; parts.comp_mag[(w * T.n_contrast + f) * NS + 1] = m1; }
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
          double mag = -2.5 * log10(tot * T.p_norm[f]);                   // mft6.py:780-782
          mag += av * T.p_ext[f];                                         // mft6.py:1163 (av == 0 when skipped)
          const double d = mag - T.p_mag[f];
          iphot += (d * d) / (T.p_err[f] * T.p_err[f]);
          if (parts.phot) parts.phot[w * T.n_phot + f] = mag;
        }
        const double init_cs = iic * (double)(T.n_contrast + T.n_phot) + icontrast + iphot;   // mft6.py:1191
        const double lnl = (init_cs != init_cs) ? -CUDART_INF : -0.5 * init_cs;                // mft6.py:1202-1205
        double out = lp + lnl;
        if (out != out) out = -CUDART_INF;
        lnprob[w] = out;
        if (parts.status) parts.status[w] = (init_cs != init_cs) ? MS_W_NONFINITE : 0;
        if (parts.lnprior) parts.lnprior[w] = lp;
        if (parts.chi2_spec) parts.chi2_spec[w] = iic;
        if (parts.logg) for (int s = 0; s < NS; ++s) parts.logg[w * NS + s] = s_logg[j][s];
      }
    }
  }
}

// ---- group table: contiguous runs of at most GW cell-sorted walkers that share a cell ----------------------
// cell_start: exclusive prefix of the cell histogram (n_cells + 1 entries); group_off: exclusive prefix of
// ceil(count / GW).  One thread per walker (after the counting sort has placed it): the first walker of every
// group writes the descriptor.
__global__ void group_table_kernel(const WalkerPlan* __restrict__ plans, const int* __restrict__ order, long long nw,
                                   const unsigned* __restrict__ cell_start, const unsigned* __restrict__ group_off,
                                   GroupDesc* __restrict__ groups) {
  const long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pos >= nw) return;
  const int cell = plans[order[pos]].cell;
  const unsigned local = (unsigned)pos - cell_start[cell];
  if (local % GW == 0) {
    const unsigned cnt = cell_start[cell + 1] - cell_start[cell];
    GroupDesc g; g.start = (int)pos; g.count = (int)min((unsigned)GW, cnt - local);
    groups[group_off[cell] + local / GW] = g;
  }
}

// Exclusive scans of the cell histogram (-> cell_start[n_cells + 1] and the mutable copy `offsets` the scatter
// consumes) and of the per-cell group counts (-> group_off[n_cells], *n_groups).  One CTA.
__global__ void __launch_bounds__(1024) group_scan_kernel(unsigned* __restrict__ hist_offsets, int n_cells,
                                                          unsigned* __restrict__ cell_start, unsigned* __restrict__ group_off,
                                                          int* __restrict__ n_groups, int* __restrict__ group_counter) {
  __shared__ unsigned wa[32], wb[32];
  __shared__ unsigned carry_a, carry_b;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_a = 0; carry_b = 0; }
  __syncthreads();
  for (int base = 0; base < n_cells; base += 1024) {
    const int i = base + tid;
    const unsigned v = (i < n_cells) ? hist_offsets[i] : 0u;
    const unsigned g = (v + GW - 1) / GW;
    unsigned ia = v, ib = g;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) { ia += ta; ib += tb; }
    }
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    if (warp == 0) {
      unsigned xa = wa[lane], xb = wb[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned ta = __shfl_up_sync(0xffffffffu, xa, o), tb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= o) { xa += ta; xb += tb; }
      }
      wa[lane] = xa; wb[lane] = xb;
    }
    __syncthreads();
    const unsigned ea = carry_a + (warp ? wa[warp - 1] : 0u) + ia - v;
    const unsigned eb = carry_b + (warp ? wb[warp - 1] : 0u) + ib - g;
    if (i < n_cells) { hist_offsets[i] = ea; cell_start[i] = ea; group_off[i] = eb; }
    __syncthreads();
    if (tid == 1023) { carry_a = ea + v; carry_b = eb + g; }
    __syncthreads();
  }
  if (tid == 0) { cell_start[n_cells] = carry_a; *n_groups = (int)carry_b; *group_counter = 0; }
}

}  // namespace grp
