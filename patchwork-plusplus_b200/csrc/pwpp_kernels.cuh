// pwpp_kernels.cuh — the sm_100a kernels of the estimateGround() path.
//
//   k_bin_hist   RNR predicate + polar bin id per point + per-chunk bin histogram
//                (reference reflected_noise_removal S:377-400, pc2czm S:578-622, flush_patches S:33-45)
//   k_bin_scan   per frame: bin offsets + per-(chunk,bin) scatter bases  (replaces the emplace_back
//                growth of czm[k][i][j], S:602-614)
//   k_scatter    stable, deterministic scatter of (x,y,z,idx) into bin-contiguous order, ascending
//                point index inside a bin — the order pc2czm produces
//   k_fit        per bin: LPR seed selection, R-VPF, R-GPF plane fits, ground/non-ground split
//                (sort S:199 [not needed, see below], extract_piecewiseground S:467-549,
//                 extract_initial_seeds S:77-149, estimate_plane S:47-75, calc_point_to_plane_d S:551-554)
//   k_gle        per frame: A-GLE verdicts, TGR, adaptive threshold + sensor-height update, output
//                segment layout in the reference's emission order
//                (S:211-311, temporal_ground_revert S:402-464, update_* S:338-375)
//   k_emit       copies every bin's ground / non-ground part to the final index lists
//                (addCloud S:28-31 + toIndices S:18-26)
//   k_gather_xyz toEigenCloud S:8-16 on demand
//
// "S:" = reference cpp/patchworkpp/src/patchworkpp.cpp, "H:" = .../include/patchwork/patchworkpp.h.
#pragma once
#include <cuda_runtime.h>

#include "pwpp_math.cuh"
#include "pwpp_gle.cuh"

namespace pwpp {

constexpr int CHUNK_PTS = 4096;      // points per CTA in k_bin_hist / k_scatter
constexpr int CHUNK_THREADS = 256;   // 8 warps, each owns 512 consecutive points
constexpr int WARP_PTS = CHUNK_PTS / (CHUNK_THREADS / 32);  // 512
constexpr int WARP_ITERS = WARP_PTS / 32;                   // 16
constexpr int MAX_LPR = 64;          // num_lpr supported by the warp selection buffer
constexpr int MAX_RVPF = 8;          // num_iter supported (R-VPF planes kept in registers)

struct FrameTable {            // per call, device arrays indexed by frame
  const long long* pt_off;     // [F+1] first point of each frame in the packed point array
  const int* chunk_off;        // [F+1] first chunk of each frame
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  // read-once data: bypass L1 allocation, keep L2 normal
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------------------------------
// k_bin_hist: grid (max_chunks_per_frame, F), 256 threads. Each warp owns 512 consecutive points.
// Writes bin ids (u16) and the chunk's histogram row (u16[nbp]).
template <bool FAST>
__global__ void __launch_bounds__(CHUNK_THREADS) k_bin_hist(const float4* __restrict__ pts, FrameTable ft, const StreamState* __restrict__ states,
                                                             Geometry g, AlgoParams ap, int has_intensity, int nbp,
                                                             unsigned short* __restrict__ bin_ids, unsigned short* __restrict__ chist) {
  extern __shared__ unsigned int s_hist[];  // [nbp]
  const int f = blockIdx.y;
  const long long p0 = ft.pt_off[f];
  const int n = (int) (ft.pt_off[f + 1] - p0);
  const int nchunks = (n + CHUNK_PTS - 1) / CHUNK_PTS;
  if ((int) blockIdx.x >= nchunks) return;
  for (int b = threadIdx.x; b < nbp; b += CHUNK_THREADS) s_hist[b] = 0;
  __syncthreads();
  const double sensor_height = states[f].sensor_height;
  const bool rnr_on = ap.enable_RNR && has_intensity;  // S:161, S:379-382
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int base = blockIdx.x * CHUNK_PTS + warp * WARP_PTS;
#pragma unroll 4
  for (int it = 0; it < WARP_ITERS; ++it) {
    const int i = base + it * 32 + lane;
    int bin = -1;
    if (i < n) {
      const float4 p = ld_stream_f4(pts + p0 + i);
      if (rnr_on && rnr_hit(p.x, p.y, p.z, p.w, sensor_height, ap)) bin = PW_BIN_RNR(g.nbins);
      else if (p.z == FLT_MIN) bin = PW_BIN_DROP(g.nbins);  // S:591
      else bin = FAST ? bin_of_point(p.x, p.y, p.z, g) : bin_of_point_exact(p.x, p.y, p.z, g);
      bin_ids[p0 + i] = (unsigned short) bin;
    }
    // warp-aggregated histogram update: one shared atomic per distinct bin in the warp
    const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
    if (bin >= 0) {
      const unsigned peers = __match_any_sync(act, bin);
      if ((peers & lanemask_lt()) == 0) atomicAdd(&s_hist[bin], __popc(peers));
    }
  }
  __syncthreads();
  unsigned short* row = chist + (size_t) (ft.chunk_off[f] + blockIdx.x) * nbp;
  for (int b = threadIdx.x; b < nbp; b += CHUNK_THREADS) row[b] = (unsigned short) s_hist[b];
}

// ---------------------------------------------------------------------------------------------------
// k_bin_scan: one CTA per frame, thread b <-> bin b (nbp <= blockDim.x * ITEMS handled by striding).
// bin_off[f][b] = first position of bin b inside the frame's sorted region ([nbp+1] entries);
// cbase[chunk][b] = position where chunk's first point of bin b goes.
__global__ void k_bin_scan(FrameTable ft, int nbp, const unsigned short* __restrict__ chist, unsigned int* __restrict__ cbase, int* __restrict__ bin_off) {
  extern __shared__ int s_scan[];  // [nbp + 1]
  const int f = blockIdx.x;
  const int c0 = ft.chunk_off[f], c1 = ft.chunk_off[f + 1];
  for (int b = threadIdx.x; b < nbp; b += blockDim.x) {
    int tot = 0;
    for (int c = c0; c < c1; ++c) tot += chist[(size_t) c * nbp + b];
    s_scan[b] = tot;
  }
  __syncthreads();
  // exclusive scan over nbp (<= 4096) values by warp 0: simple and tiny
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
  int* bo = bin_off + (size_t) f * (nbp + 1);
  for (int b = threadIdx.x; b <= nbp; b += blockDim.x) bo[b] = s_scan[b];
  for (int b = threadIdx.x; b < nbp; b += blockDim.x) {
    unsigned int run = (unsigned int) s_scan[b];
    for (int c = c0; c < c1; ++c) {
      const unsigned int v = chist[(size_t) c * nbp + b];
      cbase[(size_t) c * nbp + b] = run;
      run += v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// k_scatter: same decomposition as k_bin_hist. Stable: a point's position inside its bin is its rank
// among the frame's points of that bin in ascending point index.
//   rank = cbase[chunk][bin] + (#points of bin in lower warps of the chunk)
//        + (#points of bin in earlier iterations of this warp) + (#lower lanes with the same bin)
__global__ void __launch_bounds__(CHUNK_THREADS) k_scatter(const float4* __restrict__ pts, FrameTable ft, int nbp,
                                                            const unsigned short* __restrict__ bin_ids, const unsigned int* __restrict__ cbase,
                                                            float4* __restrict__ sorted) {
  extern __shared__ unsigned int s_wcnt[];  // [8][nbp]: per-warp histograms, then per-warp running positions
  const int f = blockIdx.y;
  const long long p0 = ft.pt_off[f];
  const int n = (int) (ft.pt_off[f + 1] - p0);
  const int nchunks = (n + CHUNK_PTS - 1) / CHUNK_PTS;
  if ((int) blockIdx.x >= nchunks) return;
  const int nwarps = CHUNK_THREADS / 32;
  for (int b = threadIdx.x; b < nwarps * nbp; b += CHUNK_THREADS) s_wcnt[b] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int base = blockIdx.x * CHUNK_PTS + warp * WARP_PTS;
  unsigned int* my = s_wcnt + warp * nbp;
  int bins[WARP_ITERS];
#pragma unroll
  for (int it = 0; it < WARP_ITERS; ++it) {
    const int i = base + it * 32 + lane;
    const int bin = (i < n) ? (int) bin_ids[p0 + i] : -1;
    bins[it] = bin;
    const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
    if (bin >= 0) {
      const unsigned peers = __match_any_sync(act, bin);
      if ((peers & lanemask_lt()) == 0) my[bin] += __popc(peers);  // only this warp writes its row
    }
    __syncwarp();
  }
  __syncthreads();
  // per bin: exclusive prefix over the 8 warps, offset by the chunk's base
  const unsigned int* cb = cbase + (size_t) (ft.chunk_off[f] + blockIdx.x) * nbp;
  for (int b = threadIdx.x; b < nbp; b += CHUNK_THREADS) {
    unsigned int run = cb[b];
#pragma unroll
    for (int w = 0; w < nwarps; ++w) { const unsigned int v = s_wcnt[w * nbp + b]; s_wcnt[w * nbp + b] = run; run += v; }
  }
  __syncthreads();
  float4* out = sorted + p0;
#pragma unroll
  for (int it = 0; it < WARP_ITERS; ++it) {
    const int i = base + it * 32 + lane;
    const int bin = bins[it];
    const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
    if (bin >= 0) {
      const unsigned peers = __match_any_sync(act, bin);
      const unsigned int pos = my[bin] + __popc(peers & lanemask_lt());
      float4 p = ld_stream_f4(pts + p0 + i);
      p.w = __int_as_float(i);
      out[pos] = p;
      __syncwarp(peers);
      if ((peers & lanemask_lt()) == 0) my[bin] += __popc(peers);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// k_fit helpers (one warp per bin)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(0xffffffffu, v); }

// Bitonic sort of 128 floats in shared memory by one warp (ascending).
__device__ __forceinline__ void warp_sort128(float* buf) {
  const int lane = lane_id();
  for (int k = 2; k <= 128; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int idx = lane + 32 * t;                       // 0..63: pair index
        const int i = ((idx & ~(j - 1)) << 1) | (idx & (j - 1));  // lower element of the pair
        const int l = i | j;
        const bool up = ((i & k) == 0);
        const float a = buf[i], b = buf[l];
        if ((a > b) == up) { buf[i] = b; buf[l] = a; }
      }
      __syncwarp();
    }
  }
}

// Streaming selection of the K smallest keys: candidates below the current bound are appended to a
// 128-slot shared buffer; when it could overflow it is sorted and truncated to K.
struct LprSelector {
  float* buf;   // [128]
  int m;        // valid entries
  float tau;    // current bound: the K-th smallest so far once K are known, else +inf
  int K;
  __device__ __forceinline__ void init(float* b, int k) { buf = b; m = 0; tau = INFINITY; K = k; }
  __device__ __forceinline__ void prune() {
    const int lane = lane_id();
    for (int i = m + lane; i < 128; i += 32) buf[i] = INFINITY;
    __syncwarp();
    warp_sort128(buf);
    if (m > K) m = K;
    if (m == K) tau = buf[K - 1];
    __syncwarp();
  }
  // every lane calls with its candidate (valid == false for lanes without one)
  __device__ __forceinline__ void push(bool valid, float key) {
    const bool c = valid && (key < tau);
    const unsigned bal = __ballot_sync(0xffffffffu, c);
    if (bal == 0) return;
    if (c) buf[m + __popc(bal & lanemask_lt())] = key;
    m += __popc(bal);
    __syncwarp();
    if (m > 96) prune();
  }
};

// extract_initial_seeds (S:77-149) over the currently alive points of the bin: returns lpr_height.
// alive(p) = not removed by an earlier R-VPF iteration.
struct RvpfPlanes {
  Plane pl[MAX_RVPF];
  int n;
};

__device__ __forceinline__ bool is_alive(const RvpfPlanes& rv, double th_dist_v, float x, float y, float z) {
  bool alive = true;
  for (int k = 0; k < rv.n; ++k) alive = alive && !(fabs(point_plane_distance(rv.pl[k], x, y, z)) < th_dist_v);  // S:499
  return alive;
}

__device__ double select_lpr(const float4* __restrict__ P, int n, bool zone0, double margin_z, int num_lpr, const RvpfPlanes& rv, double th_dist_v,
                             float* sel_buf) {
  LprSelector sel;
  sel.init(sel_buf, num_lpr);
  const int lane = lane_id();
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    bool valid = false;
    float z = 0.f;
    if (i < n) {
      const float4 p = P[i];
      z = p.z;
      valid = (rv.n == 0) || is_alive(rv, th_dist_v, p.x, p.y, p.z);
      if (zone0 && ((double) z < margin_z)) valid = false;  // S:88-96: the sorted prefix below the margin is skipped
    }
    sel.push(valid, z);
  }
  sel.prune();
  // S:99-103: double sum of the (<= num_lpr) lowest z in ascending order
  double lpr = 0.0;
  if (lane == 0) {
    double sum = 0.0;
    const int cnt = sel.m;
    for (int i = 0; i < cnt; ++i) sum += (double) sel_buf[i];
    lpr = cnt != 0 ? sum / cnt : 0.0;
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, lpr, 0);
}

// Moment sums over {alive && pred}, pred = (z < z_thr) for seeds or (dist(plane) < th) for R-GPF.
// MODE 0: seeds (z < zthr); MODE 1: signed distance to `pl` below th_dist.
template <int MODE>
__device__ __forceinline__ Moments accumulate(const float4* __restrict__ P, int n, const RvpfPlanes& rv, double th_dist_v, double zthr, const Plane& pl,
                                              double th_dist, const double c[3]) {
  Moments m;
  m.n = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) m.s1[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) m.s2[k] = 0.0;
  const int lane = lane_id();
  for (int i = lane; i < n; i += 32) {
    const float4 p = P[i];
    bool in = (rv.n == 0) || is_alive(rv, th_dist_v, p.x, p.y, p.z);
    if (MODE == 0) in = in && ((double) p.z < zthr);                        // S:108 / S:145
    else in = in && (point_plane_distance(pl, p.x, p.y, p.z) < th_dist);     // S:525 / S:529
    if (in) {
      const double dx = (double) p.x - c[0], dy = (double) p.y - c[1], dz = (double) p.z - c[2];
      m.s1[0] += dx; m.s1[1] += dy; m.s1[2] += dz;
      m.s2[0] += dx * dx; m.s2[1] += dx * dy; m.s2[2] += dx * dz;
      m.s2[3] += dy * dy; m.s2[4] += dy * dz; m.s2[5] += dz * dz;
      m.n += 1;
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) m.s1[k] = warp_sum(m.s1[k]);
#pragma unroll
  for (int k = 0; k < 6; ++k) m.s2[k] = warp_sum(m.s2[k]);
  m.n = warp_sum_i(m.n);
  return m;
}

// k_fit: one warp per (bin, frame) item, items ordered bin-major so that neighbouring warps get bins
// of similar size (the large zone-0 bins of all frames come first).
__global__ void __launch_bounds__(128) k_fit(const float4* __restrict__ sorted, FrameTable ft, const StreamState* __restrict__ states, Geometry g, AlgoParams ap,
                                             int nframes, int nbp, const int* __restrict__ bin_off, int* __restrict__ part, BinFit* __restrict__ fits) {
  __shared__ float s_sel[4][128];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long item = (long long) blockIdx.x * 4 + warp;
  const int nb_all = g.nbins + PW_NUM_PSEUDO;
  if (item >= (long long) nframes * nb_all) return;
  const int bin = (int) (item / nframes), f = (int) (item % nframes);
  const int* bo = bin_off + (size_t) f * (nbp + 1);
  const int off = bo[bin], n = bo[bin + 1] - off;
  const long long p0 = ft.pt_off[f];
  const float4* P = sorted + p0 + off;
  int* out = part + p0 + off;
  if (bin >= g.nbins || n < ap.num_min_pts || n == 0) {
    // pseudo-bins and patches below num_min_pts: every point non-ground, ascending index (S:191-195)
    for (int i = lane; i < n; i += 32) out[i] = __float_as_int(P[i].w);
    if (bin < g.nbins && lane == 0) {
      BinFit& r = fits[(size_t) f * g.nbins + bin];
      r.n = n; r.n_ground = 0; r.fitted = (n >= ap.num_min_pts) ? 1 : 0;  // n == 0 with num_min_pts <= 0: "fitted" with the stale plane
      r.verdict = 0;
      for (int k = 0; k < 3; ++k) { r.mean[k] = 0; r.normal[k] = 0; r.sv[k] = 0; }
      r.d = 0;
      if (r.fitted) r.verdict = PW_FIT_NO_PLANE;  // plane must be taken from the stale carry in k_gle
    }
    return;
  }
  const int zone = (bin >= g.bin_base[3]) ? 3 : (bin >= g.bin_base[2]) ? 2 : (bin >= g.bin_base[1]) ? 1 : 0;
  const bool zone0 = (zone == 0);
  const double margin_z = ap.adaptive_seed_selection_margin * states[f].sensor_height;  // S:90
  float* sel_buf = s_sel[warp];

  RvpfPlanes rv;
  rv.n = 0;
  Plane pl;  // the "member" plane: normal_, pc_mean_, singular_values_, d_
  bool have_plane = false;
  const float4 first = P[0];
  double c[3] = {(double) first.x, (double) first.y, 0.0};

  // 1. R-VPF (S:482-508). For zone != 0 the fitted plane can never be used (the loop breaks at once
  //    and the R-GPF seed fit below overwrites it because its seed set is non-empty for th_seeds > 0,
  //    which pwpp_create enforces), so the fit is skipped there.
  if (ap.enable_RVPF && zone0) {
    for (int it = 0; it < ap.num_iter; ++it) {
      const double lpr = select_lpr(P, n, true, margin_z, ap.num_lpr, rv, ap.th_dist_v, sel_buf);
      c[2] = lpr;
      const Moments m = accumulate<0>(P, n, rv, ap.th_dist_v, lpr + ap.th_seeds_v, pl, 0.0, c);
      if (m.n > 0) { plane_from_moments(m, c, pl); have_plane = true; }
      if (have_plane && pl.normal[2] < ap.uprightness_thr) {  // S:489
        if (rv.n < MAX_RVPF) rv.pl[rv.n++] = pl;
      } else break;
    }
  }
  // 2. R-GPF (S:513-543)
  {
    const double lpr = select_lpr(P, n, zone0, margin_z, ap.num_lpr, rv, ap.th_dist_v, sel_buf);
    c[2] = lpr;
    const Moments m = accumulate<0>(P, n, rv, ap.th_dist_v, lpr + ap.th_seeds, pl, 0.0, c);
    if (m.n > 0) { plane_from_moments(m, c, pl); have_plane = true; }
  }
  for (int it = 0; it < ap.num_iter - 1; ++it) {
    if (!have_plane) break;
    const double cc[3] = {pl.mean[0], pl.mean[1], pl.mean[2]};
    const Moments m = accumulate<1>(P, n, rv, ap.th_dist_v, 0.0, pl, ap.th_dist, cc);
    if (m.n > 0) plane_from_moments(m, cc, pl);
  }
  // last iteration: split into ground / non-ground and refit on the ground part (S:528-542).
  // Ground indices are written from the front, non-ground from the back (k_emit un-reverses).
  int n_ground = 0;
  {
    const double cc[3] = {have_plane ? pl.mean[0] : c[0], have_plane ? pl.mean[1] : c[1], have_plane ? pl.mean[2] : c[2]};
    Moments m;
    m.n = 0;
    for (int k = 0; k < 3; ++k) m.s1[k] = 0.0;
    for (int k = 0; k < 6; ++k) m.s2[k] = 0.0;
    int g_run = 0, ng_run = 0;
    const bool any_iter = ap.num_iter >= 1;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      bool valid = i < n, is_g = false;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        p = P[i];
        const bool alive = (rv.n == 0) || is_alive(rv, ap.th_dist_v, p.x, p.y, p.z);
        // num_iter == 0: the R-GPF loop body never runs, dst stays empty (S:516)
        is_g = alive && any_iter && have_plane && (point_plane_distance(pl, p.x, p.y, p.z) < ap.th_dist);
        if (is_g) {
          const double dx = (double) p.x - cc[0], dy = (double) p.y - cc[1], dz = (double) p.z - cc[2];
          m.s1[0] += dx; m.s1[1] += dy; m.s1[2] += dz;
          m.s2[0] += dx * dx; m.s2[1] += dx * dy; m.s2[2] += dx * dz;
          m.s2[3] += dy * dy; m.s2[4] += dy * dz; m.s2[5] += dz * dz;
          m.n += 1;
        }
      }
      const unsigned bg = __ballot_sync(0xffffffffu, valid && is_g);
      const unsigned bn = __ballot_sync(0xffffffffu, valid && !is_g);
      if (valid) {
        const int idx = __float_as_int(p.w);
        if (is_g) out[g_run + __popc(bg & lanemask_lt())] = idx;
        else out[n - 1 - (ng_run + __popc(bn & lanemask_lt()))] = idx;
      }
      g_run += __popc(bg);
      ng_run += __popc(bn);
    }
    n_ground = g_run;
    for (int k = 0; k < 3; ++k) m.s1[k] = warp_sum(m.s1[k]);
    for (int k = 0; k < 6; ++k) m.s2[k] = warp_sum(m.s2[k]);
    m.n = warp_sum_i(m.n);
    if (m.n > 0 && any_iter) plane_from_moments(m, cc, pl);
  }
  if (lane == 0) {
    BinFit& r = fits[(size_t) f * g.nbins + bin];
    r.n = n; r.n_ground = n_ground; r.fitted = 1;
    r.verdict = have_plane ? 0 : PW_FIT_NO_PLANE;
    for (int k = 0; k < 3; ++k) { r.mean[k] = pl.mean[k]; r.normal[k] = pl.normal[k]; r.sv[k] = pl.sv[k]; }
    r.d = pl.d;
  }
}

// ---------------------------------------------------------------------------------------------------
// k_gle: one thread block of one warp per frame; the sequential A-GLE / TGR / threshold logic lives in
// pwpp_gle.cuh (host+device) and is walked by lane 0 in the reference's loop order (S:184-311).
__global__ void __launch_bounds__(32) k_gle(FrameTable ft, StreamState* __restrict__ states, double* __restrict__ hist, int hcap, Geometry g, AlgoParams ap,
                                            int nbp, const int* __restrict__ bin_off, BinFit* __restrict__ fits, BinSeg* __restrict__ segs,
                                            int* __restrict__ num_ground, int* __restrict__ num_patches, float* __restrict__ centers, float* __restrict__ normals,
                                            int* __restrict__ num_dropped) {
  __shared__ GleScratch scratch;
  const int f = blockIdx.x;
  if (threadIdx.x != 0) return;
  double* h_elev = hist + ((size_t) f * 2 + 0) * 4 * hcap;
  double* h_flat = hist + ((size_t) f * 2 + 1) * 4 * hcap;
  int ng = 0, np = 0, nd = 0;
  gle_frame(g, ap, states[f], h_elev, h_flat, hcap, bin_off + (size_t) f * (nbp + 1), fits + (size_t) f * g.nbins,
            segs + (size_t) f * (g.nbins + PW_NUM_PSEUDO), centers + (size_t) f * g.nbins * 3, normals + (size_t) f * g.nbins * 3, scratch, ng, np, nd);
  update_thresholds(ap, states[f], h_elev, h_flat, hcap);
  num_ground[f] = ng;
  num_patches[f] = np;
  num_dropped[f] = nd;
}

// ---------------------------------------------------------------------------------------------------
// k_emit: grid (chunks, F): thread per sorted position; copies part[] into the final lists.
__global__ void __launch_bounds__(256) k_emit(FrameTable ft, Geometry g, int nbp, const int* __restrict__ bin_off, const BinFit* __restrict__ fits,
                                              const BinSeg* __restrict__ segs, const int* __restrict__ part, int* __restrict__ out_idx) {
  extern __shared__ int s_off[];  // [nb_all + 1]
  const int f = blockIdx.y;
  const long long p0 = ft.pt_off[f];
  const int n = (int) (ft.pt_off[f + 1] - p0);
  const int base = blockIdx.x * CHUNK_PTS;
  if (base >= n) return;
  const int nb_all = g.nbins + PW_NUM_PSEUDO;
  const int* bo = bin_off + (size_t) f * (nbp + 1);
  for (int b = threadIdx.x; b <= nb_all; b += blockDim.x) s_off[b] = bo[b];
  __syncthreads();
  const BinSeg* seg = segs + (size_t) f * nb_all;
  const BinFit* fit = fits + (size_t) f * g.nbins;
  for (int i = base + threadIdx.x; i < n && i < base + CHUNK_PTS; i += blockDim.x) {
    if (i >= s_off[nb_all]) continue;
    // binary search: largest b with s_off[b] <= i
    int lo = 0, hi = nb_all;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= i) lo = mid; else hi = mid; }
    const int b = lo;
    const int j = i - s_off[b];
    const int nbin = s_off[b + 1] - s_off[b];
    const int ng = (b < g.nbins) ? fit[b].n_ground : 0;
    const BinSeg sg = seg[b];
    int dst;
    int src = i;
    if (j < ng) dst = sg.g_dst + j;
    else {
      if (sg.ng_dst < 0) continue;  // dropped points (S:591)
      dst = sg.ng_dst + (j - ng);
      // fitted bins store their non-ground part reversed (k_fit); skipped and pseudo bins ascending
      if (b < g.nbins && fit[b].fitted) src = s_off[b] + (nbin - 1 - (j - ng));
    }
    out_idx[p0 + dst] = part[p0 + src];
  }
}

// k_gather_xyz: toEigenCloud (S:8-16): xyz of the listed points of one frame.
__global__ void k_gather_xyz(const float4* __restrict__ pts, const int* __restrict__ idx, int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float4 p = pts[idx[i]];
    dst[3 * i] = p.x; dst[3 * i + 1] = p.y; dst[3 * i + 2] = p.z;
  }
}

}  // namespace pwpp
