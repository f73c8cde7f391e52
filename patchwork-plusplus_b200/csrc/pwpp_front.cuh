// pwpp_front.cuh — the front end of the path (RNR + polar binning + per-frame scan + stable scatter) as ONE persistent
// kernel, software-pipelined through L2 (PWPP_FRONT=1; round-2 candidate: checked on the SIMT twin, not yet measured).
//
// Why: k_bin_hist and k_scatter each stream the whole batch from HBM (r01: 0.66 + 0.83 ms of a 5.76 ms step, 52 B/point
// of traffic between them); by the time k_scatter runs, a 1024-frame batch (1.9 GB) has long left the 126 MB L2, so the
// cloud is read from HBM twice. Here the three stages are items of one ordered work list
//     H(f,c)  bin ids + histogram of chunk c of frame f          (body of k_bin_hist)
//     S(f)    offsets, chunk bases, work queues of frame f        (body of k_bin_scan)
//     P(f,c)  stable scatter of chunk c of frame f                (body of k_scatter)
// interleaved so that P(f - W, *) follows H(f, *), S(f): the scatter of a frame runs W frames (~32 MB of points) after
// its binning pass and re-reads the cloud and the bin ids from L2; HBM sees 16 B/point in, 16 B/point out.
// Persistent CTAs claim items in list order. Dependencies are counters in global memory: S(f) waits until all H(f,*)
// have finished, P(f,c) until S(f) has. A waiting CTA never blocks progress: every item before it in the list has
// already been claimed by a CTA that is running it, and H items wait for nothing — so no co-residency assumption and no
// deadlock for any grid size >= 1. Data written by one CTA and read by another (histogram rows, chunk bases, bin ids)
// is published with __threadfence() + an atomic and read with ld.global.cg (L2, never a stale L1 line).
// The bodies are copies of the stand-alone kernels (which stay untouched as the default path); results are identical
// by construction: same per-chunk histograms, same scan, same stable ranks.
#pragma once
#include "pwpp_kernels.cuh"

namespace pwpp {

// one work item: type in the top 2 bits (0 = H, 1 = S, 2 = P), frame in bits 8..29, chunk in bits 0..7 is too small
// for dense frames, so two ints: x = type | frame << 2, y = chunk
struct FrontItem { int tf, chunk; };
constexpr int FRONT_H = 0, FRONT_S = 1, FRONT_P = 2;
constexpr int FRONT_THREADS = CHUNK_THREADS;

#if defined(PWPP_SIMT_EMU)
__device__ __forceinline__ int ld_acquire_i(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ unsigned short ldcg_u16(const unsigned short* p) { return *p; }
__device__ __forceinline__ unsigned int ldcg_u32(const unsigned int* p) { return *p; }
__device__ __forceinline__ float4 ldcg_f4(const float4* p) { return *p; }
__device__ __forceinline__ void __threadfence() {}
__device__ __forceinline__ void front_spin_pause() {   // concurrent CTAs: let the others run; sequential CTAs: nobody could ever satisfy the wait
  if (simt::g_concurrent) simt::yield_(); else simt::deadlock("k_front: an item waits for a later one (the work list is out of order)");
}
#else
__device__ __forceinline__ int ld_acquire_i(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned short ldcg_u16(const unsigned short* p) { unsigned short v; asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p)); return v; }
__device__ __forceinline__ unsigned int ldcg_u32(const unsigned int* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg_f4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ void front_spin_pause() { __nanosleep(64); }
#endif

// shared-memory words of k_front: the largest body scratch (8 x nbp for the scatter, nbp + 1 for the scan) + 1 + 3 x classes
__host__ __device__ inline int front_body_words(int nbp) { return (CHUNK_THREADS / 32) * nbp > nbp + 1 ? (CHUNK_THREADS / 32) * nbp : nbp + 1; }
__host__ __device__ inline size_t front_smem_bytes(int nbp) { return (size_t) (front_body_words(nbp) + 1 + 3 * NUM_CLASSES) * sizeof(unsigned int); }

struct FrontArgs {
  const float4* pts;
  FrameTable ft;
  const StreamState* states;
  Geometry g;
  AlgoParams ap;
  int has_intensity, nbp, nbins, fast_bin;
  unsigned short* bin_ids;
  unsigned short* chist;
  unsigned int* cbase;
  int* bin_off;
  WorkQueues wq;
  BinFit* fits;
  float4* sorted;
  const FrontItem* items;
  int nitems;
  int l2max;        // upper bound of the L2 patch class (CLS_L2_MAX, or CLS_L2_WIDE_MAX with PWPP_L2_WIDE)
  int mmax;         // upper bound of the M patch class (CLS_M_MAX, or CLS_M_HALF_MAX with PWPP_M_HALF)
  int* ctr;         // [0] next item, [1 .. 1+F) hist_done per frame, [1+F .. 1+2F) scan_done per frame (zeroed per call)
  int nframes;
};

// ---- H: k_bin_hist's body for one chunk (PIPE = 2 variant). smem: nbp unsigned ----
__device__ __forceinline__ void front_hist_chunk(const FrontArgs& a, int f, int chunk, unsigned int* s_hist) {
  const long long p0 = a.ft.pt_off[f];
  const int n = (int) (a.ft.pt_off[f + 1] - p0);
  for (int b = threadIdx.x; b < a.nbp; b += FRONT_THREADS) s_hist[b] = 0;
  __syncthreads();
  const double sensor_height = a.states[f].sensor_height;
  const bool rnr_on = a.ap.enable_RNR && a.has_intensity;  // S:161, S:379-382
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int base = chunk * CHUNK_PTS + warp * WARP_PTS;
  constexpr int HB = 2;
  const int last = n - 1;
  float4 q[HB], qn[HB];
#pragma unroll
  for (int u = 0; u < HB; ++u) { const int i = base + u * 32 + lane; q[u] = a.pts[p0 + (i < n ? i : last)]; }   // normal loads: the lines should stay in L2 for P
#pragma unroll 1
  for (int h = 0; h < WARP_ITERS; h += HB) {
    if (h + HB < WARP_ITERS) {
#pragma unroll
      for (int u = 0; u < HB; ++u) { const int i = base + (h + HB + u) * 32 + lane; qn[u] = a.pts[p0 + (i < n ? i : last)]; }
    }
#pragma unroll
    for (int u = 0; u < HB; ++u) {
      const int i = base + (h + u) * 32 + lane;
      const float4 p = q[u];
      int bin = -1;
      if (i < n) {
        if (rnr_on && rnr_hit(p.x, p.y, p.z, p.w, sensor_height, a.ap)) bin = PW_BIN_RNR(a.g.nbins);
        else if (p.z == FLT_MIN) bin = PW_BIN_DROP(a.g.nbins);  // S:591
        else bin = a.fast_bin ? bin_of_point(p.x, p.y, p.z, a.g) : bin_of_point_exact(p.x, p.y, p.z, a.g);
        a.bin_ids[p0 + i] = (unsigned short) bin;
      }
      const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
      if (bin >= 0) {
        const unsigned peers = __match_any_sync(act, bin);
        if ((peers & lanemask_lt()) == 0) atomicAdd(&s_hist[bin], __popc(peers));
      }
    }
#pragma unroll
    for (int u = 0; u < HB; ++u) q[u] = qn[u];
  }
  __syncthreads();
  unsigned short* row = a.chist + (size_t) (a.ft.chunk_off[f] + chunk) * a.nbp;
  for (int b = threadIdx.x; b < a.nbp; b += FRONT_THREADS) row[b] = (unsigned short) s_hist[b];
}

// ---- S: k_bin_scan's body for one frame. smem: nbp + 1 ints ----
__device__ __forceinline__ void front_scan_frame(const FrontArgs& a, int f, int* s_scan, int* s_cls) {
  int* s_cls_cnt = s_cls;                       // [NUM_CLASSES] each
  int* s_cls_base = s_cls + NUM_CLASSES;
  int* s_cls_pos = s_cls + 2 * NUM_CLASSES;
  const int nbp = a.nbp, nbins = a.nbins;
  const int c0 = a.ft.chunk_off[f], c1 = a.ft.chunk_off[f + 1];
  if (threadIdx.x < NUM_CLASSES) { s_cls_cnt[threadIdx.x] = 0; s_cls_pos[threadIdx.x] = 0; }
  for (int b = threadIdx.x; b < nbp; b += blockDim.x) {
    int tot = 0;
    for (int c = c0; c < c1; ++c) tot += ldcg_u16(a.chist + (size_t) c * nbp + b);
    s_scan[b] = tot;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    int carry = 0;
    for (int b0 = 0; b0 < nbp; b0 += 32) {
      const int b = b0 + threadIdx.x;
      int v = b < nbp ? s_scan[b] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int) threadIdx.x >= o) incl += t; }
      if (b < nbp) s_scan[b] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) s_scan[nbp] = carry;
  }
  __syncthreads();
  int* bo = a.bin_off + (size_t) f * (nbp + 1);
  for (int b = threadIdx.x; b <= nbp; b += blockDim.x) bo[b] = s_scan[b];
  for (int b = threadIdx.x; b < nbp; b += blockDim.x) {
    unsigned int run = (unsigned int) s_scan[b];
    for (int c = c0; c < c1; ++c) {
      const unsigned int v = ldcg_u16(a.chist + (size_t) c * nbp + b);
      a.cbase[(size_t) c * nbp + b] = run;
      run += v;
    }
  }
  const int l2max = a.l2max, mmax = a.mmax;
  auto cls_of = [l2max, mmax](int n) { return n <= CLS_S_MAX ? 0 : n <= mmax ? 1 : n <= CLS_L1_MAX ? 2 : n <= l2max ? 3 : n <= CLS_L3_MAX ? 4 : 5; };
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
    const int n = s_scan[b + 1] - s_scan[b];
    if (n >= a.ap.num_min_pts && n > 0) atomicAdd(&s_cls_cnt[cls_of(n)], 1);
    else {
      BinFit& r = a.fits[(size_t) f * nbins + b];
      r.n = n; r.n_ground = 0; r.d = 0.0;
      for (int k = 0; k < 3; ++k) { r.mean[k] = 0.0; r.normal[k] = 0.0; r.sv[k] = 0.0; }
      r.fitted = (n >= a.ap.num_min_pts) ? 1 : 0;
      r.verdict = r.fitted ? PW_FIT_NO_PLANE : 0;
    }
  }
  __syncthreads();
  if (threadIdx.x < NUM_CLASSES) s_cls_base[threadIdx.x] = s_cls_cnt[threadIdx.x] ? atomicAdd(&a.wq.count[threadIdx.x], s_cls_cnt[threadIdx.x]) : 0;
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
    const int n = s_scan[b + 1] - s_scan[b];
    if (n >= a.ap.num_min_pts && n > 0) {
      const int c = cls_of(n);
      a.wq.items[c][s_cls_base[c] + atomicAdd(&s_cls_pos[c], 1)] = make_work_item(f, b, n, a.ft.pt_off[f] + (long long) s_scan[b]);
    }
  }
}

// ---- P: k_scatter's body (PIPE = false) for one chunk. smem: 8 x nbp unsigned ----
__device__ __forceinline__ void front_scatter_chunk(const FrontArgs& a, int f, int chunk, unsigned int* s_wcnt) {
  const int nbp = a.nbp;
  const long long p0 = a.ft.pt_off[f];
  const int n = (int) (a.ft.pt_off[f + 1] - p0);
  const int nwarps = FRONT_THREADS / 32;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int base = chunk * CHUNK_PTS + warp * WARP_PTS;
  const int last = n - 1;
  constexpr int SB = 4;
  float4 qa[SB];
  for (int b = threadIdx.x; b < nwarps * nbp; b += FRONT_THREADS) s_wcnt[b] = 0;
  __syncthreads();
  unsigned int* my = s_wcnt + warp * nbp;
  int bins[WARP_ITERS];
#pragma unroll
  for (int it = 0; it < WARP_ITERS; ++it) {
    const int i = base + it * 32 + lane;
    bins[it] = (i < n) ? (int) ldcg_u16(a.bin_ids + p0 + i) : -1;
  }
#pragma unroll
  for (int it = 0; it < WARP_ITERS; ++it) {
    const int bin = bins[it];
    const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
    if (bin >= 0) {
      const unsigned peers = __match_any_sync(act, bin);
      if ((peers & lanemask_lt()) == 0) my[bin] += __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  const unsigned int* cb = a.cbase + (size_t) (a.ft.chunk_off[f] + chunk) * nbp;
  for (int b = threadIdx.x; b < nbp; b += FRONT_THREADS) {
    unsigned int run = ldcg_u32(cb + b);
#pragma unroll
    for (int w = 0; w < nwarps; ++w) { const unsigned int v = s_wcnt[w * nbp + b]; s_wcnt[w * nbp + b] = run; run += v; }
  }
  __syncthreads();
  float4* out = a.sorted + p0;
#pragma unroll
  for (int h = 0; h < WARP_ITERS; h += SB) {
#pragma unroll
    for (int u = 0; u < SB; ++u) { const int i = base + (h + u) * 32 + lane; qa[u] = ldcg_f4(a.pts + p0 + (i < n ? i : last)); }   // L2 hit: read by H a moment ago
#pragma unroll
    for (int u = 0; u < SB; ++u) {
      const int i = base + (h + u) * 32 + lane;
      const int bin = bins[h + u];
      const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
      if (bin >= 0) {
        const unsigned peers = __match_any_sync(act, bin);
        const unsigned int pos = my[bin] + __popc(peers & lanemask_lt());
        float4 p = qa[u];
        p.w = __int_as_float(i);
        out[pos] = p;
        __syncwarp(peers);
        if ((peers & lanemask_lt()) == 0) my[bin] += __popc(peers);
      }
      __syncwarp();
    }
  }
}

// The ordered work list, one thread per step s in [0, F + W): step s holds H(s,*), S(s) (s < F) and P(s - W,*) (s >= W).
// Closed-form positions from the chunk table: items of kinds H and S before step s: co(min(s,F)) + min(s,F); of kind P:
// co(clamp(s - W, 0, F)), with co(j) = chunk_off[j] - chunk_off[0].
__global__ void k_front_plan(const int* __restrict__ chunk_off, int F, int W, FrontItem* __restrict__ items) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= F + W) return;
  const int c0 = chunk_off[0];
  const int sf = s < F ? s : F;
  int pw = s - W;
  pw = pw < 0 ? 0 : (pw > F ? F : pw);
  int pos = (chunk_off[sf] - c0) + sf + (chunk_off[pw] - c0);
  if (s < F) {
    const int nc = chunk_off[s + 1] - chunk_off[s];
    for (int c = 0; c < nc; ++c) { items[pos].tf = FRONT_H | (s << 2); items[pos].chunk = c; ++pos; }
    items[pos].tf = FRONT_S | (s << 2); items[pos].chunk = 0; ++pos;
  }
  if (s >= W && s - W < F) {
    const int f = s - W;
    const int nc = chunk_off[f + 1] - chunk_off[f];
    for (int c = 0; c < nc; ++c) { items[pos].tf = FRONT_P | (f << 2); items[pos].chunk = c; ++pos; }
  }
}

__global__ void __launch_bounds__(FRONT_THREADS, 4) k_front(FrontArgs a) {
  // dynamic shared memory only (the SIMT twin can then run several CTAs of this kernel concurrently):
  // [0, front_body_words) the bodies' scratch, then the claimed item index, then the scan's class counters
  PW_DYN_SHARED(unsigned int, s_dyn);
  const int body_words = front_body_words(a.nbp);
  int& s_k = *reinterpret_cast<int*>(s_dyn + body_words);
  int* s_cls = reinterpret_cast<int*>(s_dyn + body_words + 1);
  int* hist_done = a.ctr + 1;
  int* scan_done = a.ctr + 1 + a.nframes;
  for (;;) {
    if (threadIdx.x == 0) s_k = atomicAdd(&a.ctr[0], 1);
    __syncthreads();
    const int k = s_k;
    if (k >= a.nitems) return;
    const FrontItem it = a.items[k];
    const int type = it.tf & 3, f = it.tf >> 2;
    if (type == FRONT_H) {
      front_hist_chunk(a, f, it.chunk, s_dyn);
      __syncthreads();                       // every thread's bin ids and the histogram row are written
      if (threadIdx.x == 0) { __threadfence(); atomicAdd(&hist_done[f], 1); }
    } else if (type == FRONT_S) {
      const int need = a.ft.chunk_off[f + 1] - a.ft.chunk_off[f];
      if (threadIdx.x == 0) { while (ld_acquire_i(&hist_done[f]) < need) front_spin_pause(); }
      __syncthreads();
      front_scan_frame(a, f, reinterpret_cast<int*>(s_dyn), s_cls);
      __syncthreads();
      if (threadIdx.x == 0) { __threadfence(); atomicAdd(&scan_done[f], 1); }
    } else {
      if (threadIdx.x == 0) { while (ld_acquire_i(&scan_done[f]) < 1) front_spin_pause(); }
      __syncthreads();
      front_scatter_chunk(a, f, it.chunk, s_dyn);
    }
    __syncthreads();   // s_k and the dynamic shared memory are reused by the next item
  }
}

}  // namespace pwpp
