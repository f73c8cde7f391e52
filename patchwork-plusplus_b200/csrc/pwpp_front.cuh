// pwpp_front.cuh — the front end of the path as ONE kernel: a THREAD-BLOCK CLUSTER per frame.
//
//   RNR + polar binning      reference reflected_noise_removal S:377-400, pc2czm S:578-622, flush_patches S:33-45
//   per-frame scan + queues  the growth of czm[k][i][j] (S:602-614) and the per-bin gate S:191
//   stable scatter           points of a bin contiguous, ascending point index inside a bin (the order pc2czm produces)
// ("S:" = cpp/patchworkpp/src/patchworkpp.cpp). Results are those of the three stand-alone kernels k_bin_hist / k_bin_scan /
// k_scatter (pwpp_kernels.cuh), which remain as the PWPP_FRONT=0 path; this kernel removes what made them expensive:
//   * the three kernels communicated through global memory (per-chunk histograms and scatter bases: chist / cbase) and
//     each streamed the whole batch, so the cloud was read from HBM twice (r01: 49 B/point of DRAM traffic, 1.54 ms);
//   * here the FC_CS CTAs of a cluster split one frame into slices. Pass 1: every CTA bins its slice tile by tile — tiles
//     arrive through the TMA engine (cp.async.bulk global -> shared, mbarrier completion, double buffered: SASS UBLKCP) —
//     and keeps its histogram in shared memory. After a cluster barrier every CTA reads the other CTAs' histograms through
//     DISTRIBUTED SHARED MEMORY (cluster.map_shared_rank), scans the bins and knows where its points of every bin go;
//     rank 0 also writes the bin offsets and fills the fit work queues. Pass 2: the CTA streams its slice again — 2 MB of a
//     frame read a few microseconds earlier by the same cluster: L2 hits — and scatters. HBM sees the cloud once (16 B/pt in,
//     16 B/pt out for the re-laid-out copy, 2 + 2 B/pt of bin ids that mostly stay in L2).
// The bin ids still go to global memory (pwpp_copy_bin_ids, and frames of any size: a dense frame's slice does not fit on chip).
#pragma once
#include "pwpp_common.cuh"
#include "pwpp_fit.cuh"

namespace pwpp {

constexpr int FC_CS = 8;                 // CTAs per cluster (portable maximum)
constexpr int FC_THREADS = 256;
constexpr int FC_TILE = 1024;            // points per tile: 16 KB, two buffers in flight (48 KB of shared memory per CTA in all: 4 CTAs per SM)
constexpr int FC_ROWS = FC_TILE / FC_THREADS;   // 32-point rows per warp and tile: 4

__host__ __device__ inline size_t front_cluster_smem_bytes(int nbp) {
  return (size_t) 2 * FC_TILE * sizeof(float4) + (size_t) 2 * nbp * sizeof(unsigned) + (size_t) (1 + FC_THREADS / 32) * nbp * sizeof(unsigned short) + (size_t) (nbp + 1) * sizeof(int) + 16 + 64 + 3 * NUM_CLASSES * sizeof(int);
}

template <bool FAST, int L2MAX>
__global__ void
#if !defined(PWPP_SIMT_EMU)
__cluster_dims__(FC_CS, 1, 1)
#endif
__launch_bounds__(FC_THREADS, 4) k_front_cluster(const float4* __restrict__ pts, FrameTable ft, const StreamState* __restrict__ states, Geometry g, AlgoParams ap,
                                                  int has_intensity, int nbp, int nbins, unsigned short* __restrict__ bin_ids, int* __restrict__ bin_off, WorkQueues wq,
                                                  BinFit* __restrict__ fits, float4* __restrict__ sorted) {
  PW_DYN_SHARED(unsigned char, s_raw);
  float4* s_tile = reinterpret_cast<float4*>(s_raw);                                   // [2][FC_TILE]
  unsigned* s_hist = reinterpret_cast<unsigned*>(s_raw + (size_t) 2 * FC_TILE * 16);   // [nbp] this CTA's bin counts (read by the whole cluster)
  unsigned* s_base = s_hist + nbp;                                                     // [nbp] where this CTA's next point of a bin goes
  unsigned short* s_wcnt = reinterpret_cast<unsigned short*>(s_base + nbp);            // [8][nbp] per-warp counts, then positions inside the tile's share of a bin
  unsigned short* s_tcnt = s_wcnt + (FC_THREADS / 32) * nbp;                           // [nbp] points of the current tile per bin (nbp is a multiple of 32: 4-byte aligned end)
  int* s_scan = reinterpret_cast<int*>(s_tcnt + nbp);                                  // [nbp + 1]
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(s_scan) + (((size_t) (nbp + 1) * 4 + 15) & ~(size_t) 15));   // [2] (+ 48 B pad)
  int* s_cls = reinterpret_cast<int*>(s_bar + 8);                                      // [3][NUM_CLASSES]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int f = blockIdx.y;
  const int rank = (int) pw_cluster_rank();
  const long long p0 = ft.pt_off[f];
  const int n = (int) (ft.pt_off[f + 1] - p0);
  const int ntiles = (n + FC_TILE - 1) / FC_TILE;
  const int t0 = (int) ((long long) ntiles * rank / FC_CS), t1 = (int) ((long long) ntiles * (rank + 1) / FC_CS);   // this CTA's tiles
  const float4* fp = pts + p0;
  const double sensor_height = states[f].sensor_height;
  const bool rnr_on = ap.enable_RNR && has_intensity;  // S:161, S:379-382

  for (int b = tid; b < nbp; b += FC_THREADS) s_hist[b] = 0;
#if !defined(PWPP_SIMT_EMU)
  if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif
  __syncthreads();
  unsigned ph[2] = {0u, 0u};   // mbarrier phase of each buffer
  (void) ph;
  auto issue = [&](int t, int buf) {   // thread 0: tile t -> buffer buf
#if !defined(PWPP_SIMT_EMU)
    const int cnt = min(FC_TILE, n - t * FC_TILE);
    fence_proxy_async();
    mbar_expect_tx(&s_bar[buf], (unsigned) cnt * 16u);
    bulk_g2s(s_tile + buf * FC_TILE, fp + (size_t) t * FC_TILE, (unsigned) cnt * 16u, &s_bar[buf]);
#else
    (void) t; (void) buf;
#endif
  };
  auto wait_tile = [&](int t, int buf) {
#if !defined(PWPP_SIMT_EMU)
    (void) t;
    mbar_wait(&s_bar[buf], ph[buf] & 1u);
    ++ph[buf];
#else
    const int cnt = min(FC_TILE, n - t * FC_TILE);
    for (int i = tid; i < cnt; i += FC_THREADS) s_tile[buf * FC_TILE + i] = fp[(size_t) t * FC_TILE + i];
    __syncthreads();
#endif
  };

  // ---------------- pass 1: bin ids + this CTA's histogram ----------------
  if (tid == 0 && t0 < t1) issue(t0, 0);
  for (int t = t0; t < t1; ++t) {
    const int buf = (t - t0) & 1;
    if (tid == 0 && t + 1 < t1) issue(t + 1, buf ^ 1);   // (every thread left that buffer at the barrier that ended tile t - 1)
    wait_tile(t, buf);
    const float4* tp = s_tile + buf * FC_TILE;
    const int base = t * FC_TILE + w * (FC_ROWS * 32);
#pragma unroll 2
    for (int r = 0; r < FC_ROWS; ++r) {
      const int li = w * (FC_ROWS * 32) + r * 32 + lane;   // index inside the tile
      const int i = base + r * 32 + lane;                  // index inside the frame
      int bin = -1;
      if (i < n) {
        const float4 p = tp[li];
        if (rnr_on && rnr_hit(p.x, p.y, p.z, p.w, sensor_height, ap)) bin = PW_BIN_RNR(g.nbins);
        else if (p.z == FLT_MIN) bin = PW_BIN_DROP(g.nbins);  // S:591
        else bin = FAST ? bin_of_point(p.x, p.y, p.z, g) : bin_of_point_exact(p.x, p.y, p.z, g);
        bin_ids[p0 + i] = (unsigned short) bin;
      }
      const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
      if (bin >= 0) {   // warp-aggregated histogram update: one shared atomic per distinct bin in the row
        const unsigned peers = __match_any_sync(act, bin);
        if ((peers & lanemask_lt()) == 0) atomicAdd(&s_hist[bin], __popc(peers));
      }
    }
    __syncthreads();
  }
  pw_cluster_sync();   // every histogram of the frame is complete and visible cluster-wide

  // ---------------- scan: totals over the cluster, bin offsets, this CTA's bases ----------------
  for (int b = tid; b < nbp; b += FC_THREADS) {
    unsigned tot = 0, before = 0;
#pragma unroll
    for (int c = 0; c < FC_CS; ++c) {
      const unsigned h = *pw_cluster_map(&s_hist[b], c);   // distributed shared memory
      if (c < rank) before += h;
      tot += h;
    }
    s_scan[b] = (int) tot;
    s_base[b] = before;
  }
  __syncthreads();
  if (tid < 32) {   // exclusive scan over nbp (<= 4096) bins by warp 0
    int carry = 0;
    for (int b0 = 0; b0 < nbp; b0 += 32) {
      const int b = b0 + tid;
      const int v = b < nbp ? s_scan[b] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += u; }
      if (b < nbp) s_scan[b] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (tid == 0) s_scan[nbp] = carry;
  }
  __syncthreads();
  for (int b = tid; b < nbp; b += FC_THREADS) s_base[b] += (unsigned) s_scan[b];
  pw_cluster_sync();   // nobody reads a remote histogram after this point (a CTA may exit before its neighbours)
  if (rank == 0) {
    // what k_bin_scan leaves for the later stages: bin offsets, the fit work queues (S:191: patches below num_min_pts are
    // not fitted), the records of the patches that will not be fitted
    int* bo = bin_off + (size_t) f * (nbp + 1);
    for (int b = tid; b <= nbp; b += FC_THREADS) bo[b] = s_scan[b];
    int* s_cnt = s_cls, *s_cbase = s_cls + NUM_CLASSES, *s_pos = s_cls + 2 * NUM_CLASSES;
    if (tid < NUM_CLASSES) { s_cnt[tid] = 0; s_pos[tid] = 0; }
    __syncthreads();
    auto cls_of = [](int m) { return m <= CLS_S_MAX ? 0 : m <= CLS_M_MAX ? 1 : m <= CLS_L1_MAX ? 2 : m <= L2MAX ? 3 : m <= CLS_L3_MAX ? 4 : 5; };
    for (int b = tid; b < nbins; b += FC_THREADS) {
      const int m = s_scan[b + 1] - s_scan[b];
      if (m >= ap.num_min_pts && m > 0) atomicAdd(&s_cnt[cls_of(m)], 1);
      else {
        BinFit& r = fits[(size_t) f * nbins + b];
        r.n = m; r.n_ground = 0; r.d = 0.0;
        for (int k = 0; k < 3; ++k) { r.mean[k] = 0.0; r.normal[k] = 0.0; r.sv[k] = 0.0; }
        r.fitted = (m >= ap.num_min_pts) ? 1 : 0;   // an EMPTY patch with num_min_pts <= 0 is "fitted" with the previous patch's plane (S:49)
        r.verdict = r.fitted ? PW_FIT_NO_PLANE : 0;
      }
    }
    __syncthreads();
    if (tid < NUM_CLASSES) s_cbase[tid] = s_cnt[tid] ? atomicAdd(&wq.count[tid], s_cnt[tid]) : 0;
    __syncthreads();
    for (int b = tid; b < nbins; b += FC_THREADS) {
      const int m = s_scan[b + 1] - s_scan[b];
      if (m >= ap.num_min_pts && m > 0) {
        const int c = cls_of(m);
        wq.items[c][s_cbase[c] + atomicAdd(&s_pos[c], 1)] = make_work_item(f, b, m, p0 + (long long) s_scan[b]);
      }
    }
  }
  __syncthreads();

  // ---------------- pass 2: stable scatter of this CTA's slice ----------------
  // position of a point = s_base[bin] (first free slot of this CTA for the bin) + points of the bin in lower warps of the
  // tile + earlier rows of this warp + lower lanes of the row
  float4* out = sorted + p0;
  if (tid == 0 && t0 < t1) issue(t0, 0);
  for (int t = t0; t < t1; ++t) {
    const int buf = (t - t0) & 1;
    if (tid == 0 && t + 1 < t1) issue(t + 1, buf ^ 1);
    const int base = t * FC_TILE + w * (FC_ROWS * 32);
    int bins[FC_ROWS];
#pragma unroll
    for (int r = 0; r < FC_ROWS; ++r) { const int i = base + r * 32 + lane; bins[r] = (i < n) ? (int) bin_ids[p0 + i] : -1; }   // written by this very thread in pass 1
    for (int b = tid; b < (FC_THREADS / 32) * nbp / 2; b += FC_THREADS) reinterpret_cast<unsigned*>(s_wcnt)[b] = 0u;
    __syncthreads();
    unsigned short* my = s_wcnt + w * nbp;
#pragma unroll
    for (int r = 0; r < FC_ROWS; ++r) {
      const int bin = bins[r];
      const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
      if (bin >= 0) {
        const unsigned peers = __match_any_sync(act, bin);
        if ((peers & lanemask_lt()) == 0) my[bin] = (unsigned short) (my[bin] + __popc(peers));   // only this warp writes its row
      }
      __syncwarp();
    }
    __syncthreads();
    for (int b = tid; b < nbp; b += FC_THREADS) {   // exclusive prefix over the 8 warps (positions relative to the CTA's running base of the bin)
      unsigned run = 0;
#pragma unroll
      for (int ww = 0; ww < FC_THREADS / 32; ++ww) { const unsigned v = s_wcnt[ww * nbp + b]; s_wcnt[ww * nbp + b] = (unsigned short) run; run += v; }
      s_tcnt[b] = (unsigned short) run;
    }
    wait_tile(t, buf);
    __syncthreads();
    const float4* tp = s_tile + buf * FC_TILE;
#pragma unroll
    for (int r = 0; r < FC_ROWS; ++r) {
      const int li = w * (FC_ROWS * 32) + r * 32 + lane;
      const int i = base + r * 32 + lane;
      const int bin = bins[r];
      const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
      if (bin >= 0) {
        const unsigned peers = __match_any_sync(act, bin);
        const unsigned pos = s_base[bin] + my[bin] + __popc(peers & lanemask_lt());
        float4 p = tp[li];
        p.w = __int_as_float(i);
        out[pos] = p;
        __syncwarp(peers);
        if ((peers & lanemask_lt()) == 0) my[bin] = (unsigned short) (my[bin] + __popc(peers));
      }
      __syncwarp();
    }
    __syncthreads();
    for (int b = tid; b < nbp; b += FC_THREADS) s_base[b] += s_tcnt[b];   // the CTA's running base moves past this tile (ordered before the next tile's placement by its barriers)
  }
}

}  // namespace pwpp
