// pwpp_fit_group.cuh — plane fitting (R-VPF + R-GPF) for GROUPS of patches: the fit stage of the path for every patch that
// fits on chip (<= GCAP points); larger patches (dense sensors) go to k_fit_big (pwpp_fit_big.cuh).
//
// Reference: cpp/patchworkpp/src/patchworkpp.cpp ("S:") extract_piecewiseground 467-549, extract_initial_seeds 77-149,
// estimate_plane 47-75, calc_point_to_plane_d 551-554. Same per-patch state machine, arithmetic contract and selection
// logic as pwpp_fit.cuh (which this kernel replaces for the classes S..L3); what changes is the decomposition:
//
//   * A work item is a GROUP: a run of consecutive bins of one frame (k_bin_scan_groups packs them) whose points are one
//     contiguous range of the bin-sorted array. The group is copied into shared memory ONCE, by a single bulk async copy
//     (cp.async.bulk global -> shared::cta, completion on an mbarrier: the TMA path, SASS UBLKCP), as float4 {x,y,z,idx}.
//   * All patches of a group advance in lock step through "rounds" (one pass over the points + one plane fit). The pass is
//     done by ALL warps over ALL points of the group: warp w owns a contiguous range of 32-point rows and walks the
//     segments (warp range x patch) inside it, so the per-patch quantities (plane, thresholds) are loaded once per segment.
//   * The 3x3 eigen-problems of a round are solved with ONE LANE PER PATCH (warp 0; in fused rounds warp 1 solves the
//     inner seed planes at the same time) instead of one warp (or one CTA) per patch redundantly: the ~1.5k-instruction
//     double-precision solve was 60-75 % of the issued instructions of the warp-per-patch kernels and the reason seven
//     warps of a CTA-per-patch kernel idled at a barrier.
//   * Several CTAs share an SM, so one group's solve phase overlaps the other groups' passes.
//
// Determinism: a lane accumulates its points in row order, lanes are combined by an xor butterfly, segments by the
// solver lane in warp order: the sums do not depend on scheduling (bit-reproducible run to run).
#pragma once
#include "pwpp_fit.cuh"

namespace pwpp {

// optional phase clocks (diagnostic builds only: -DPWPP_PHASE_CLOCKS): thread 0 of every CTA accumulates the cycles it spends
// in each phase of k_fit_group into g_phase_clk[class][phase]; read with pwpp_debug_phase_clocks()
#if defined(PWPP_PHASE_CLOCKS) && !defined(PWPP_SIMT_EMU)
__device__ unsigned long long g_phase_clk[3][16];
#define PW_CLK_DECL long long _clk_t = clock64(); unsigned long long _clk_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define PW_CLK(ph) do { if (threadIdx.x == 0) { const long long _n = clock64(); _clk_acc[ph] += (unsigned long long) (_n - _clk_t); _clk_t = _n; } } while (0)
#define PW_CNT(ph) do { if (threadIdx.x == 0) _clk_acc[ph] += 1; } while (0)
#define PW_CLK_FLUSH(cls) do { if (threadIdx.x == 0) for (int _q = 0; _q < 12; ++_q) atomicAdd(&g_phase_clk[cls][_q], _clk_acc[_q]); } while (0)
// event trace of CTA 0 of one class (PWPP_TRACE_CLS): (event id, warp, clock) per warp
#ifndef PWPP_TRACE_CLS
#define PWPP_TRACE_CLS 1
#endif
__device__ unsigned g_evn;
__device__ uint4 g_ev[16384];   // 16 warps x 1024 events: every warp of CTA 0 fills its own range (no atomics: ~30 cycles per event)
#define PW_EV_DECL unsigned _evi = 0
#define PW_EV(id) do { if (CLS == PWPP_TRACE_CLS && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && _evi < 1024u) { g_ev[((threadIdx.x >> 5) << 10) + _evi] = make_uint4((unsigned) (id), threadIdx.x >> 5, (unsigned) clock64(), 1u); ++_evi; } } while (0)
#else
#define PW_CLK_DECL
#define PW_CLK(ph)
#define PW_CNT(ph)
#define PW_CLK_FLUSH(cls)
#define PW_EV(id)
#define PW_EV_DECL
#endif
constexpr int GRP_CSEG = 64;    // candidates one segment of a multi-warp patch may hand to the selecting warp (a half of the warp's buffer)
constexpr int GRP_CBUF = 128;   // candidates the exact selection handles (4 keys per lane)

// work item of a group: x = (frame << 12) | first bin, y = (bins in the span << 14) | points, (z, w) = offset of the
// group's first point in the bin-sorted array
__device__ __forceinline__ int4 make_group_item(int frame, int bin0, int nspan, int npts, long long start) {
  return make_int4((frame << 12) | bin0, (nspan << 14) | npts, (int) (unsigned) (start & 0xffffffffll), (int) (start >> 32));
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map needed for a contiguous range) -----------------------
#if !defined(PWPP_SIMT_EMU)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// exact selection among cc (<= GRP_CBUF) candidate keys in cbuf: mean of the `target` smallest (S:99-103) by RANKING: a
// candidate's rank = number of candidates before it in (key, position) order; all compares are independent (no serial
// bisection). The sum of <= 32 floats in double is exact, so the result does not depend on the candidates' order. Uniform.
__device__ __forceinline__ double grp_rank_mean(const unsigned* cbuf, int cc, int target) {
  const int lane = lane_id();
  unsigned ck[GRP_CBUF / 32];
  int rank[GRP_CBUF / 32];
  const int nq = (cc + 31) >> 5;
#pragma unroll
  for (int q = 0; q < GRP_CBUF / 32; ++q) { const int i = lane + 32 * q; ck[q] = i < cc ? cbuf[i] : 0xffffffffu; rank[q] = 0; }
  // candidates are broadcast 8 at a time (independent shared-memory reads in flight); entries past cc hold 0xffffffff (or
  // stale keys) and are masked by the position test
  for (int j0 = 0; j0 < cc; j0 += 8) {
    unsigned v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = cbuf[(j0 + u) < GRP_CBUF ? (j0 + u) : (GRP_CBUF - 1)];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u;
      if (j < cc) {   // uniform
#pragma unroll
        for (int q = 0; q < GRP_CBUF / 32; ++q) { if (q >= nq) break; rank[q] += (v[u] < ck[q] || (v[u] == ck[q] && j < lane + 32 * q)) ? 1 : 0; }
      }
    }
  }
  double ps = 0.0;
#pragma unroll
  for (int q = 0; q < GRP_CBUF / 32; ++q) if (lane + 32 * q < cc && rank[q] < target) ps += (double) key_to_float(ck[q]);
  ps = warp_sum(ps);
  return ps / (double) target;
}

// Rare path (num_lpr > 32, or more candidates tie below the bound than the buffers hold): exact selection over the whole
// patch [ps, pe) of the group by one warp without any buffer: bisection on the order keys, one sweep of the patch per bit
// below the bits all candidates share. Kept out of line.
__device__ __noinline__ double grp_lpr_bisect(const float4* __restrict__ pts, int ps, int pe, bool use_alive, const unsigned* __restrict__ alive, bool zone0,
                                              double margin_z, int num_lpr) {
  const int lane = lane_id();
  auto valid_key = [&](int j0, unsigned& key) {
    const int j = j0 + lane;
    bool valid = j >= ps && j < pe;
    const float z = pts[valid ? j : ps].z;
    if (use_alive) valid = valid && ((alive[j0 >> 5] >> lane) & 1u);
    if (zone0 && ((double) z < margin_z)) valid = false;
    key = order_key(z);
    return valid;
  };
  unsigned kmn = 0xffffffffu, kmx = 0u;
  int nv = 0;
  for (int j0 = ps & ~31; j0 < pe; j0 += 32) {
    unsigned key;
    if (valid_key(j0, key)) { kmn = min(kmn, key); kmx = max(kmx, key); ++nv; }
  }
  nv = __reduce_add_sync(0xffffffffu, nv);
  const int target = nv < num_lpr ? nv : num_lpr;
  if (target == 0) return 0.0;
  kmn = __reduce_min_sync(0xffffffffu, kmn);
  kmx = __reduce_max_sync(0xffffffffu, kmx);
  const unsigned ans = kth_key(kmn, kmx, target, [&](unsigned cand) {
    int cnt = 0;
    for (int j0 = ps & ~31; j0 < pe; j0 += 32) {
      unsigned key;
      if (valid_key(j0, key) && key < cand) ++cnt;
    }
    return __reduce_add_sync(0xffffffffu, cnt);
  });
  double sum = 0.0;
  int c_lt = 0;
  for (int j0 = ps & ~31; j0 < pe; j0 += 32) {
    unsigned key;
    if (valid_key(j0, key) && key < ans) { sum += (double) key_to_float(key); ++c_lt; }
  }
  sum = warp_sum(sum);
  c_lt = __reduce_add_sync(0xffffffffu, c_lt);
  return (sum + (double) (target - c_lt) * (double) key_to_float(ans)) / (double) target;
}

__device__ __noinline__ void grp_solve(const double* a9, int n, const double* c3, double* plane10) {
  Moments m;
  m.n = n;
#pragma unroll
  for (int q = 0; q < 3; ++q) m.s1[q] = a9[q];
#pragma unroll
  for (int q = 0; q < 6; ++q) m.s2[q] = a9[3 + q];
  const double c[3] = {c3[0], c3[1], c3[2]};
  Plane pl;
  plane_from_moments(m, c, pl);
#pragma unroll
  for (int q = 0; q < 3; ++q) { plane10[q] = pl.mean[q]; plane10[3 + q] = pl.normal[q]; plane10[6 + q] = pl.sv[q]; }
  plane10[9] = pl.d;
}

// GCAP: points a group may hold (shared memory: 16 B each); MP: patches per group (<= 32: one solver lane each);
// NW: warps per CTA; CLS: index of the work queue this instantiation drains.
template <int GCAP, int MP, int NW, int MINB, int CLS>
__global__ void __launch_bounds__(NW * 32, MINB) k_fit_group(const float4* __restrict__ sorted, FrameTable ft, const StreamState* __restrict__ states, Geometry g,
                                                              AlgoParams ap, int nbp, const int* __restrict__ bin_off, WorkQueues wq, int* __restrict__ part,
                                                              BinFit* __restrict__ fits) {
  constexpr int NT = NW * 32;
  constexpr int NROWS = GCAP / 32;
  constexpr int MAXSEG = MP + NW;
  static_assert(2 * GRP_CSEG == GRP_CBUF, "two pool slots per warp buffer");
  static_assert(MP <= 32 && GCAP % 32 == 0 && NW >= 2, "one solver lane per patch; warp 1 solves the inner planes of fused rounds");
  PW_DYN_SHARED(float4, s_gpts);   // [GCAP] the group's points {x, y, z, idx}
  // per patch
  __shared__ int s_pstart[MP], s_pn[MP], s_pbin[MP], s_state[MP], s_rvpf_it[MP], s_gpf_it[MP], s_have[MP], s_nground[MP], s_rm[MP], s_anyrm[MP], s_wfirst[MP],
      s_wlast[MP], s_totn[MP], s_mni_tot[MP];
  __shared__ double s_plane[MP][10];    // mean[3] normal[3] sv[3] d of the current plane
  __shared__ double s_planeI[MP][10];   // fused round: plane of the inner (R-GPF seed) set, solved by warp 1
  __shared__ double s_mi[MP][9];        // fused round: moments of the inner set
  __shared__ double s_c[MP][3];         // reference point of the moment sums: first x, first y, LPR height
  __shared__ double s_tot[MP][9];       // running moment sums of the R-GPF phase
  __shared__ double s_lpr[MP];
  __shared__ float s_pf[MP][4];         // float copy of (normal, d) for the fp32 distance filter
  // per warp / per segment (segment = the part of one patch inside one warp's rows)
  __shared__ int s_wplo[NW], s_wphi[NW], s_wsegbase[NW];
  // the candidate buffers of the LPR selection and the per-segment partial moments of the pass are never live together
  constexpr int SCRATCH_BYTES = (NW * GRP_CBUF * 4 > 2 * MAXSEG * 9 * 8) ? NW * GRP_CBUF * 4 : 2 * MAXSEG * 9 * 8;
  __shared__ __align__(16) unsigned char s_scratch[SCRATCH_BYTES];
  double (*s_pa)[9] = reinterpret_cast<double (*)[9]>(s_scratch);
  double (*s_pb)[9] = reinterpret_cast<double (*)[9]>(s_scratch + MAXSEG * 9 * 8);
  unsigned (*s_cbuf)[GRP_CBUF] = reinterpret_cast<unsigned (*)[GRP_CBUF]>(s_scratch);   // a multi-warp patch's segment uses one half (GRP_CSEG) of its warp's buffer
  __shared__ int s_pmn[MAXSEG], s_pchg[MAXSEG], s_pmni[MAXSEG];
  __shared__ unsigned s_segT[MAXSEG];
  __shared__ int s_segnv[MAXSEG], s_segg[MAXSEG], s_segv[MAXSEG];
  __shared__ unsigned s_alive[NROWS];   // ballot word per row: alive = not removed by R-VPF (S:495-504); set membership lives in per-lane register bits
  __shared__ int s_poolcnt[NW][2];
  __shared__ int4 s_item;
  __shared__ int s_np;
#if !defined(PWPP_SIMT_EMU)
  __shared__ __align__(8) unsigned long long s_bar;
#endif
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned lt = lanemask_lt();
  const float thf = (float) ap.th_dist;
  const bool fuse_ok = ap.th_seeds <= ap.th_seeds_v;
  const int K = ap.num_lpr;
  const int4 no_item = make_int4(-1, 0, 0, 0);
  unsigned phase = 0;
#if !defined(PWPP_SIMT_EMU)
  if (tid == 0) { mbar_init(&s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif
  (void) phase;
  PW_CLK_DECL;
  PW_EV_DECL;

  for (;;) {
    __syncthreads();   // the previous group's shared state is dead; the mbarrier is initialised
    PW_CLK(5);
    if (tid == 0) {
      const int t = atomicAdd(&wq.head[CLS], 1);
      const int4 it = t < wq.count[CLS] ? wq.items[CLS][t] : no_item;
      s_item = it;
#if !defined(PWPP_SIMT_EMU)
      if (it.x >= 0) {   // one bulk copy of the whole group (16 B per point: size and both addresses are multiples of 16)
        const unsigned bytes = (unsigned) (it.y & 0x3fff) * 16u;
        fence_proxy_async();   // the generic-proxy reads of the previous group's points are ordered before the async write
        mbar_expect_tx(&s_bar, bytes);
        bulk_g2s(s_gpts, sorted + work_item_start(it), bytes, &s_bar);
      }
#endif
    }
    __syncthreads();
    const int4 cur = s_item;
    if (cur.x < 0) { PW_CLK_FLUSH(CLS); return; }
    PW_CNT(8); PW_EV(1);
    const int f = cur.x >> 12, b0 = cur.x & 0xfff, nspan = cur.y >> 14, n = cur.y & 0x3fff;
    const long long start = work_item_start(cur);
    int* out = part + start;
    const int* bo = bin_off + (size_t) f * (nbp + 1);
    const double margin_z = ap.adaptive_seed_selection_margin * states[f].sensor_height;   // S:90
    const int nrows = (n + 31) >> 5;
    const int RPW = (nrows + NW - 1) / NW;                  // rows per warp
    const int w_begin = min(n, w * RPW * 32), w_end = min(n, (w + 1) * RPW * 32);
#if defined(PWPP_SIMT_EMU)
    for (int j = tid; j < n; j += NT) s_gpts[j] = sorted[start + j];
#endif
    // ---- patch table: the fitted bins of the span, in bin order (S:191: bins below num_min_pts are not fitted) ----
    if (w == 0) {
      int np = 0;
      const int base = bo[b0];
      for (int c0 = 0; c0 < nspan; c0 += 32) {
        const bool act = c0 + lane < nspan;
        const int b = b0 + c0 + lane;
        const int o0 = act ? bo[b] : 0, o1 = act ? bo[b + 1] : 0;
        const bool fit = act && (o1 - o0) >= ap.num_min_pts && (o1 - o0) > 0;
        const unsigned m = __ballot_sync(0xffffffffu, fit);
        if (fit) {
          const int p = np + __popc(m & lt);
          if (p < MP) {
            s_pstart[p] = o0 - base; s_pn[p] = o1 - o0; s_pbin[p] = b;
            s_state[p] = (ap.enable_RVPF && b < g.bin_base[1]) ? ST_RVPF : ST_SEED;   // zone != 0: the R-VPF fit is dead code (see k_fit_stream)
            s_rvpf_it[p] = 0; s_gpf_it[p] = 0; s_have[p] = 0; s_nground[p] = 0; s_rm[p] = 0; s_anyrm[p] = 0; s_totn[p] = 0;
#pragma unroll
            for (int q = 0; q < 10; ++q) s_plane[p][q] = 0.0;
          }
        }
        np += __popc(m);
      }
      if (lane == 0) s_np = np < MP ? np : MP;   // (the packer never exceeds MP)
    }
    for (int r = tid; r < nrows; r += NT) s_alive[r] = 0xffffffffu;
    if (tid < NW) { s_wplo[tid] = MP; s_wphi[tid] = -1; }
    __syncthreads();
    const int np = s_np;
    if (tid < np) {
      const int ps = s_pstart[tid], pe = ps + s_pn[tid];
      const int wf = (ps >> 5) / RPW, wl = ((pe - 1) >> 5) / RPW;
      s_wfirst[tid] = wf; s_wlast[tid] = wl;
      for (int ww = wf; ww <= wl; ++ww) { atomicMin(&s_wplo[ww], tid); atomicMax(&s_wphi[ww], tid); }
    }
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int ww = 0; ww < NW; ++ww) { s_wsegbase[ww] = run; if (s_wphi[ww] >= s_wplo[ww]) run += s_wphi[ww] - s_wplo[ww] + 1; }
    }
#if !defined(PWPP_SIMT_EMU)
    mbar_wait(&s_bar, phase & 1u);   // the group's points have landed
    ++phase;
#endif
    __syncthreads();
    if (tid < np) { const float4 first = s_gpts[s_pstart[tid]]; s_c[tid][0] = (double) first.x; s_c[tid][1] = (double) first.y; s_c[tid][2] = 0.0; }
    const int plo = s_wplo[w], phi = s_wphi[w], segbase = s_wsegbase[w];
    __syncthreads();
    PW_CLK(0); PW_EV(2);

    // =========================================== rounds ===========================================
    // Per-lane row bits: bit k of `member` <-> the point (row w*RPW + k, lane) is in the set the current plane was fitted
    // to. The passes below contain no warp-synchronous operation per row, so consecutive rows overlap in the pipeline.
    unsigned member = 0u;
    const int row0 = w * RPW;
    for (;;) {
      // states of all patches, one per lane (uniform across the CTA)
      const int st_l = lane < np ? s_state[lane] : ST_DONE;
      const unsigned m_active = __ballot_sync(0xffffffffu, st_l != ST_DONE);
      if (m_active == 0u) break;
      const unsigned m_seed = __ballot_sync(0xffffffffu, st_l == ST_RVPF || st_l == ST_SEED);
      const unsigned m_fused = fuse_ok ? __ballot_sync(0xffffffffu, st_l == ST_RVPF) : 0u;
      const bool multi_l = lane < np && s_wfirst[lane] != s_wlast[lane];
      const unsigned m_multi_seed = __ballot_sync(0xffffffffu, multi_l) & m_seed;
      PW_CNT(9); PW_EV(10);

      // ------------------------------------ LPR selection (seed rounds) ------------------------------------
      // mean of the K lowest z among the points that are alive and, in zone 0, not below the adaptive margin (S:88-103).
      // Two-level: the K-th smallest of a segment's 32 lane minima bounds the K-th smallest point of the patch, so only the
      // few points not above it are candidates; the exact K smallest of the candidates are found by ranking.
      if (m_seed) {
        // patches inside this warp's rows first (they use the warp's whole candidate buffer), then the segments of multi-warp
        // patches (one half of the buffer each: slot 0 = the warp's first patch, slot 1 = its last), which stay for the pooling
        for (int sub = 0; sub < 2; ++sub)
        for (int p = plo; p <= phi; ++p) {
          if (!((m_seed >> p) & 1u)) continue;
          const bool multi = s_wfirst[p] != s_wlast[p];
          if (multi != (sub == 1)) continue;
          const int ps = s_pstart[p], pe = ps + s_pn[p];
          const int lo = max(ps, w_begin), hi = min(pe, w_end);
          const int kb = (lo >> 5) - row0, ke = ((hi - 1) >> 5) - row0;
          const bool rm = s_rm[p] != 0, zone0 = s_pbin[p] < g.bin_base[1];
          const bool use_alive = rm || s_anyrm[p] != 0;
          if (rm) {   // pending R-VPF removal (S:495-504): points within th_dist_v of the vertical plane leave the patch
            Plane rp;
#pragma unroll
            for (int q = 0; q < 3; ++q) { rp.mean[q] = s_plane[p][q]; rp.normal[q] = s_plane[p][3 + q]; rp.sv[q] = s_plane[p][6 + q]; }
            rp.d = s_plane[p][9];
            for (int k = kb; k <= ke; ++k) {
              const int j = ((row0 + k) << 5) + lane;
              const bool inr = j >= lo && j < hi;
              const float4 q4 = s_gpts[j];
              const unsigned word = s_alive[row0 + k];
              const bool keep = inr && ((word >> lane) & 1u) && !(fabs(point_plane_distance(rp, q4.x, q4.y, q4.z)) < ap.th_dist_v);   // S:499
              const unsigned bal = __ballot_sync(0xffffffffu, keep), segm = __ballot_sync(0xffffffffu, inr);
              __syncwarp();
              if (lane == 0) s_alive[row0 + k] = (word & ~segm) | bal;
            }
            __syncwarp();
          }
          PW_EV(11);
          // sweep 1: lane minima
          unsigned kmin = 0xffffffffu;
          int nv = 0;
#pragma unroll 4
          for (int k = kb; k <= ke; ++k) {
            const int j = ((row0 + k) << 5) + lane;
            bool valid = j >= lo && j < hi;
            const float z = s_gpts[j].z;
            if (use_alive) valid = valid && ((s_alive[row0 + k] >> lane) & 1u);
            if (zone0 && ((double) z < margin_z)) valid = false;
            if (valid) { kmin = min(kmin, order_key(z)); ++nv; }
          }
          nv = __reduce_add_sync(0xffffffffu, nv);
          // K-th smallest of the 32 lane minima by ranking (all shuffles independent); fewer than K lanes with a candidate: no bound
          unsigned T = 0xffffffffu;
          if (K <= 32) {
            int rank = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const unsigned o = __shfl_sync(0xffffffffu, kmin, i); rank += (o < kmin || (o == kmin && i < lane)) ? 1 : 0; }
            const unsigned holder = __ballot_sync(0xffffffffu, rank == K - 1);
            T = __shfl_sync(0xffffffffu, kmin, __ffs(holder) - 1);   // 0xffffffff when fewer than K lanes hold a candidate
          }
          PW_EV(12);
          const int sg = segbase + p - plo;
          if (lane == 0) { s_segT[sg] = T; s_segnv[sg] = nv; }
          // sweep 2: the candidates not above the segment's own bound, in (row, lane) order
          const int slot = (p == plo) ? 0 : 1;
          unsigned* dst = multi ? (s_cbuf[w] + slot * GRP_CSEG) : s_cbuf[w];
          const int cap = multi ? GRP_CSEG : GRP_CBUF;
          int cc = 0;
          if (nv > 0 && K <= 32) {
            for (int k = kb; k <= ke; ++k) {
              const int j = ((row0 + k) << 5) + lane;
              bool valid = j >= lo && j < hi;
              const float z = s_gpts[j].z;
              if (use_alive) valid = valid && ((s_alive[row0 + k] >> lane) & 1u);
              if (zone0 && ((double) z < margin_z)) valid = false;
              const unsigned key = order_key(z);
              const bool c = valid && key <= T;
              const unsigned bal = __ballot_sync(0xffffffffu, c);
              if (c) { const int pos = cc + __popc(bal & lt); if (pos < cap) dst[pos] = key; }
              cc += __popc(bal);
            }
            __syncwarp();
          }
          PW_EV(13);
          if (multi) {
            if (lane == 0) s_poolcnt[w][slot] = cc;
          } else {   // the whole patch lies in this warp's rows: select now
            const int target = nv < K ? nv : K;
            double lpr = 0.0;   // S:99-103 with no candidate: lpr_height stays 0
            if (target > 0) {
              if (K > 32 || cc > GRP_CBUF) lpr = grp_lpr_bisect(s_gpts, ps, pe, use_alive, s_alive, zone0, margin_z, K);
              else lpr = grp_rank_mean(s_cbuf[w], cc, target);
            }
            if (lane == 0) s_lpr[p] = lpr;
            __syncwarp();
          }
        }
        if (m_multi_seed) {
          PW_EV(14);
          __syncthreads();
          PW_EV(15);
          // patches spread over several warps: the first of them pools the segments' candidates (filtered by the tightest
          // bound) and selects
          for (int p = plo; p <= phi; ++p) {
            if (!((m_multi_seed >> p) & 1u) || s_wfirst[p] != w) continue;
            const int ps = s_pstart[p], pe = ps + s_pn[p];
            const bool zone0 = s_pbin[p] < g.bin_base[1];
            const bool use_alive = s_rm[p] != 0 || s_anyrm[p] != 0;
            const int wl = s_wlast[p];
            int nvalid = 0;
            unsigned T = 0xffffffffu;
            bool overflow = false;
            for (int ww = w; ww <= wl; ++ww) {
              const int sg = s_wsegbase[ww] + p - s_wplo[ww];
              nvalid += s_segnv[sg];
              T = min(T, s_segT[sg]);
              if (s_poolcnt[ww][(p == s_wplo[ww]) ? 0 : 1] > GRP_CSEG) overflow = true;
            }
            const int target = nvalid < K ? nvalid : K;
            double lpr = 0.0;
            if (target > 0) {
              // compact the candidates <= T of all segments into this warp's own half (its segment of p comes first: in place)
              unsigned* dst = s_cbuf[w] + ((p == plo) ? 0 : 1) * GRP_CSEG;
              int cc = 0;
              if (!overflow && K <= 32) {
                for (int ww = w; ww <= wl; ++ww) {
                  const int slot = (p == s_wplo[ww]) ? 0 : 1;
                  const int cs = s_poolcnt[ww][slot];
                  const unsigned* src = s_cbuf[ww] + slot * GRP_CSEG;
                  for (int i0 = 0; i0 < cs; i0 += 32) {
                    const int i = i0 + lane;
                    const unsigned v = i < cs ? src[i] : 0xffffffffu;
                    const bool c = i < cs && v <= T;
                    const unsigned bal = __ballot_sync(0xffffffffu, c);   // (also orders the reads of this chunk before the writes below)
                    const int pos = cc + __popc(bal & lt);
                    if (c && pos < GRP_CSEG) dst[pos] = v;
                    cc += __popc(bal);
                    __syncwarp();
                  }
                }
                if (cc > GRP_CSEG) overflow = true;
              }
              if (K > 32 || overflow) lpr = grp_lpr_bisect(s_gpts, ps, pe, use_alive, s_alive, zone0, margin_z, K);
              else lpr = grp_rank_mean(dst, cc, target);
            }
            if (lane == 0) s_lpr[p] = lpr;
            __syncwarp();
          }
        }
        PW_EV(16);
        __syncthreads();
        PW_CLK(1); PW_CNT(10); PW_EV(17);
      }

      // ------------------------------------ pass: predicate + moments ------------------------------------
      for (int p = plo; p <= phi; ++p) {
        if (!((m_active >> p) & 1u)) continue;
        const int st = s_state[p];
        const int ps = s_pstart[p], pe = ps + s_pn[p];
        const int lo = max(ps, w_begin), hi = min(pe, w_end);
        const int kb = (lo >> 5) - row0, ke = ((hi - 1) >> 5) - row0;
        const int sg = segbase + p - plo;
        const bool use_alive = s_anyrm[p] != 0 || s_rm[p] != 0;
        const double c0 = s_c[p][0], c1 = s_c[p][1];
        double a[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) a[q] = 0.0;
        int mn = 0;
        PW_EV(20);
        if (st == ST_RVPF || st == ST_SEED) {
          // seed rounds accumulate their whole set {alive, z < lpr + th} (S:107-111 / S:144-148); a fused R-VPF round also the
          // inner set {z < lpr + th_seeds} of the R-GPF seed fit that follows when nothing is removed (same LPR height)
          const bool fused = (st == ST_RVPF) && fuse_ok;
          const double lpr = s_lpr[p];
          const double zthr = lpr + (st == ST_RVPF ? ap.th_seeds_v : ap.th_seeds), zin = lpr + ap.th_seeds;
          const double c2 = lpr;
          double bi[9];
#pragma unroll
          for (int q = 0; q < 9; ++q) bi[q] = 0.0;
          int mni = 0;
#pragma unroll 2
          for (int k = kb; k <= ke; ++k) {
            const int j = ((row0 + k) << 5) + lane;
            const float4 q4 = s_gpts[j];
            bool in = j >= lo && j < hi && ((double) q4.z < zthr);
            if (use_alive) in = in && ((s_alive[row0 + k] >> lane) & 1u);
            const bool inner = in && ((double) q4.z < zin);
            const bool mem = fused ? inner : in;
            if (j >= lo && j < hi) member = (member & ~(1u << k)) | ((mem ? 1u : 0u) << k);
            if (in) {
              const double dx = (double) q4.x - c0, dy = (double) q4.y - c1, dz = (double) q4.z - c2;
              a[0] += dx; a[1] += dy; a[2] += dz;
              a[3] += dx * dx; a[4] += dx * dy; a[5] += dx * dz; a[6] += dy * dy; a[7] += dy * dz; a[8] += dz * dz;
              ++mn;
              if (fused && inner) {
                bi[0] += dx; bi[1] += dy; bi[2] += dz;
                bi[3] += dx * dx; bi[4] += dx * dy; bi[5] += dx * dz; bi[6] += dy * dy; bi[7] += dy * dz; bi[8] += dz * dz;
                ++mni;
              }
            }
          }
          PW_EV(21);
#pragma unroll
          for (int q = 0; q < 9; ++q) a[q] = warp_sum(a[q]);
          mn = __reduce_add_sync(0xffffffffu, mn);
          if (fused) {
#pragma unroll
            for (int q = 0; q < 9; ++q) bi[q] = warp_sum(bi[q]);
            mni = __reduce_add_sync(0xffffffffu, mni);
            if (lane == 0) {
#pragma unroll
              for (int q = 0; q < 9; ++q) s_pb[sg][q] = bi[q];
              s_pmni[sg] = mni;
            }
          }
          if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 9; ++q) s_pa[sg][q] = a[q];
            s_pmn[sg] = mn; s_pchg[sg] = 1;
          }
        } else {
          // R-GPF rounds (S:516-543) are incremental: every fit of the phase shares the reference point, a round adds (+) /
          // removes (-) only the points whose membership changed; no change at all = the fixpoint of the iteration
          const bool have = s_have[p] != 0;
          PlaneF pf;
          pf.n0 = s_pf[p][0]; pf.n1 = s_pf[p][1]; pf.n2 = s_pf[p][2]; pf.d = s_pf[p][3];
          const double c2 = s_c[p][2];
          int nchg = 0;
#pragma unroll 2
          for (int k = kb; k <= ke; ++k) {
            const int j = ((row0 + k) << 5) + lane;
            const float4 q4 = s_gpts[j];
            bool in = false;
            if (have) {
              int fl = dist_filter(pf, thf, q4.x, q4.y, q4.z);
              if (fl < 0) {   // inside the fp32 error bound of th_dist: decide in double (S:525 / S:529, exact)
                Plane pl;
#pragma unroll
                for (int q = 0; q < 3; ++q) pl.normal[q] = s_plane[p][3 + q];
                pl.d = s_plane[p][9];
                fl = (point_plane_distance(pl, q4.x, q4.y, q4.z) < ap.th_dist) ? 1 : 0;
              }
              in = fl != 0;
            }
            const bool inr = j >= lo && j < hi;
            if (use_alive) in = in && ((s_alive[row0 + k] >> lane) & 1u);
            const bool was = (member >> k) & 1u;
            if (inr && in != was) {
              member ^= 1u << k;
              const double wgt = in ? 1.0 : -1.0;
              const double dx = (double) q4.x - c0, dy = (double) q4.y - c1, dz = (double) q4.z - c2;
              const double wx = dx * wgt, wy = dy * wgt, wz = dz * wgt;
              a[0] += wx; a[1] += wy; a[2] += wz;
              a[3] += wx * dx; a[4] += wx * dy; a[5] += wx * dz; a[6] += wy * dy; a[7] += wy * dz; a[8] += wz * dz;
              mn += in ? 1 : -1;
              ++nchg;
            }
          }
          PW_EV(22);
          nchg = __reduce_add_sync(0xffffffffu, nchg);
          if (nchg) {   // uniform
#pragma unroll
            for (int q = 0; q < 9; ++q) a[q] = warp_sum(a[q]);
            mn = __reduce_add_sync(0xffffffffu, mn);
            if (lane == 0) {
#pragma unroll
              for (int q = 0; q < 9; ++q) s_pa[sg][q] = a[q];
              s_pmn[sg] = mn;
            }
          }
          if (lane == 0) s_pchg[sg] = nchg;
        }
      }
      PW_EV(23);
      __syncthreads();
      PW_CLK(2); PW_EV(24);

      // ------------------------------------ plane fits: one lane per patch ------------------------------------
      Plane pv;              // warp 0: the plane fitted this round
      double msum[9];
      int mn_p = 0, chg_p = 0;
      const bool mine = (w == 0) && ((m_active >> lane) & 1u);
      const bool my_seed = (m_seed >> lane) & 1u, my_fused = (m_fused >> lane) & 1u;
      bool solved = false;
      if (mine) {
        const int p = lane;
#pragma unroll
        for (int q = 0; q < 9; ++q) msum[q] = 0.0;
        for (int ww = s_wfirst[p]; ww <= s_wlast[p]; ++ww) {   // segments in warp order
          const int sg = s_wsegbase[ww] + p - s_wplo[ww];
          const int ch = s_pchg[sg];
          if (ch) {
#pragma unroll
            for (int q = 0; q < 9; ++q) msum[q] += s_pa[sg][q];
            mn_p += s_pmn[sg];
            chg_p += ch;
          }
        }
        PW_EV(30);
        int tn = mn_p;
        if (my_seed) s_c[p][2] = s_lpr[p];
        else if (chg_p != 0) {
          tn += s_totn[p];
#pragma unroll
          for (int q = 0; q < 9; ++q) msum[q] += s_tot[p][q];
        }
        if (my_seed || chg_p != 0) {
#pragma unroll
          for (int q = 0; q < 9; ++q) s_tot[p][q] = msum[q];
          s_totn[p] = tn;
          if (tn > 0) {
            Moments m;
            m.n = tn;
#pragma unroll
            for (int q = 0; q < 3; ++q) m.s1[q] = msum[q];
#pragma unroll
            for (int q = 0; q < 6; ++q) m.s2[q] = msum[3 + q];
            const double cc[3] = {s_c[p][0], s_c[p][1], s_c[p][2]};
            plane_from_moments(m, cc, pv);
            solved = true;
            PW_EV(31);
          }
        }
      }
      if (m_fused) {   // uniform: warp 1 solves the inner seed planes meanwhile
        if (w == 1 && ((m_fused >> lane) & 1u)) {
          const int p = lane;
          Moments m;
          m.n = 0;
#pragma unroll
          for (int q = 0; q < 3; ++q) m.s1[q] = 0.0;
#pragma unroll
          for (int q = 0; q < 6; ++q) m.s2[q] = 0.0;
          for (int ww = s_wfirst[p]; ww <= s_wlast[p]; ++ww) {
            const int sg = s_wsegbase[ww] + p - s_wplo[ww];
#pragma unroll
            for (int q = 0; q < 3; ++q) m.s1[q] += s_pb[sg][q];
#pragma unroll
            for (int q = 0; q < 6; ++q) m.s2[q] += s_pb[sg][3 + q];
            m.n += s_pmni[sg];
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) s_mi[p][q] = m.s1[q];
#pragma unroll
          for (int q = 0; q < 6; ++q) s_mi[p][3 + q] = m.s2[q];
          s_mni_tot[p] = m.n;
          if (m.n > 0) {
            const double cc[3] = {s_c[p][0], s_c[p][1], s_lpr[p]};
            Plane pi;
            plane_from_moments(m, cc, pi);
#pragma unroll
            for (int q = 0; q < 3; ++q) { s_planeI[p][q] = pi.mean[q]; s_planeI[p][3 + q] = pi.normal[q]; s_planeI[p][6 + q] = pi.sv[q]; }
            s_planeI[p][9] = pi.d;
          }
        }
        PW_EV(32);
        __syncthreads();
        PW_EV(33);
      }
      // ---- state transition (the machine of k_fit_cta / k_fit_resident) ----
      if (mine) {
        const int p = lane;
        int state = s_state[p];
        bool have_plane = s_have[p] != 0;
        if (solved) {   // S:49: an empty set keeps the previous plane
#pragma unroll
          for (int q = 0; q < 3; ++q) { s_plane[p][q] = pv.mean[q]; s_plane[p][3 + q] = pv.normal[q]; s_plane[p][6 + q] = pv.sv[q]; }
          s_plane[p][9] = pv.d;
          have_plane = true;
        }
        int tot_n = s_totn[p];
        s_rm[p] = 0;
        if (my_fused) {
          const double vz = solved ? pv.normal[2] : s_plane[p][5];
          const bool taken = !(have_plane && vz < ap.uprightness_thr);   // S:489 false -> S:506 break: nothing removed, the seed fit follows
          if (taken) {   // the R-GPF seed fit of S:513-514 from the same pass
            const int mni = s_mni_tot[p];
#pragma unroll
            for (int q = 0; q < 9; ++q) s_tot[p][q] = s_mi[p][q];
            s_totn[p] = mni;
            tot_n = mni;
            if (mni > 0) {
#pragma unroll
              for (int q = 0; q < 10; ++q) s_plane[p][q] = s_planeI[p][q];
              have_plane = true;
            }
            state = (ap.num_iter > 1) ? ST_GPF : ST_FINAL;
            s_gpf_it[p] = 0;
          } else {       // S:489: remove the vertical structure (at the start of the next round), iterate
            s_rm[p] = 1; s_anyrm[p] = 1;
            const int it = s_rvpf_it[p] + 1;
            s_rvpf_it[p] = it;
            if (it >= ap.num_iter) state = ST_SEED;
          }
        } else if (state == ST_RVPF) {
          if (have_plane && s_plane[p][5] < ap.uprightness_thr) {
            s_rm[p] = 1; s_anyrm[p] = 1;
            const int it = s_rvpf_it[p] + 1;
            s_rvpf_it[p] = it;
            if (it >= ap.num_iter) state = ST_SEED;
          } else state = ST_SEED;   // S:506 break
        } else if (state == ST_SEED) {
          state = (ap.num_iter > 1) ? ST_GPF : ST_FINAL;
          s_gpf_it[p] = 0;
        } else if (state == ST_GPF) {
          const int it = s_gpf_it[p] + 1;
          s_gpf_it[p] = it;
          if (it >= ap.num_iter - 1) state = ST_FINAL;
          if (chg_p == 0) state = ST_DONE;   // fixpoint: every later iteration reproduces this set and this plane
        } else {
          state = ST_DONE;   // ST_FINAL
        }
        if (state == ST_DONE) s_nground[p] = have_plane ? tot_n : 0;
        s_state[p] = state;
        s_have[p] = have_plane ? 1 : 0;
        s_pf[p][0] = (float) s_plane[p][3]; s_pf[p][1] = (float) s_plane[p][4]; s_pf[p][2] = (float) s_plane[p][5]; s_pf[p][3] = (float) s_plane[p][9];
      }
      PW_EV(34);
      __syncthreads();
      PW_CLK(3); PW_EV(35);
    }

    PW_EV(40);
    // ---- stable partition per patch: ground indices ascending, then non-ground indices ascending ----
    for (int p = plo; p <= phi; ++p) {
      const int ps = s_pstart[p], pe = ps + s_pn[p];
      const int lo = max(ps, w_begin), hi = min(pe, w_end);
      const int kb = (lo >> 5) - row0, ke = ((hi - 1) >> 5) - row0;
      int gcount = 0;
      if (s_have[p] != 0)
        for (int k = kb; k <= ke; ++k) {
          const int j = ((row0 + k) << 5) + lane;
          gcount += (j >= lo && j < hi && ((member >> k) & 1u)) ? 1 : 0;
        }
      gcount = __reduce_add_sync(0xffffffffu, gcount);
      if (lane == 0) { const int sg = segbase + p - plo; s_segg[sg] = gcount; s_segv[sg] = hi - lo; }
    }
    __syncthreads();
    for (int p = plo; p <= phi; ++p) {
      const int ps = s_pstart[p], pe = ps + s_pn[p];
      const int lo = max(ps, w_begin), hi = min(pe, w_end);
      const int kb = (lo >> 5) - row0, ke = ((hi - 1) >> 5) - row0;
      const bool have = s_have[p] != 0;
      const int n_ground = s_nground[p];
      int g_run = 0, ng_run = 0;
      for (int ww = s_wfirst[p]; ww < w; ++ww) { const int sg = s_wsegbase[ww] + p - s_wplo[ww]; g_run += s_segg[sg]; ng_run += s_segv[sg] - s_segg[sg]; }
      int* po = out + ps;
      for (int k = kb; k <= ke; ++k) {
        const int j = ((row0 + k) << 5) + lane;
        const bool v = j >= lo && j < hi;
        const bool isg = v && have && ((member >> k) & 1u);
        const unsigned bv = __ballot_sync(0xffffffffu, v);
        const unsigned bg = __ballot_sync(0xffffffffu, isg);
        const unsigned bn = bv & ~bg;
        if (v) {
          const int idx = __float_as_int(s_gpts[j].w);
          if (isg) po[g_run + __popc(bg & lt)] = idx;
          else po[n_ground + ng_run + __popc(bn & lt)] = idx;
        }
        g_run += __popc(bg);
        ng_run += __popc(bn);
      }
    }
    if (tid < np) {
      const int p = tid;
      BinFit& r = fits[(size_t) f * g.nbins + s_pbin[p]];
      r.n = s_pn[p]; r.n_ground = s_nground[p]; r.fitted = 1;
      r.verdict = s_have[p] ? 0 : PW_FIT_NO_PLANE;
#pragma unroll
      for (int q = 0; q < 3; ++q) { r.mean[q] = s_plane[p][q]; r.normal[q] = s_plane[p][3 + q]; r.sv[q] = s_plane[p][6 + q]; }
      r.d = s_plane[p][9];
    }
  }
}

}  // namespace pwpp
