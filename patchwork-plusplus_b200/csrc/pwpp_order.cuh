// pwpp_order.cuh — reference emission order inside a patch (optional stage, pwpp_set_output_order(PWPP_ORDER_REFERENCE)).
//
// The reference sorts every bin by z before fitting (std::sort, cpp/patchworkpp/src/patchworkpp.cpp:199 "S:199") and all
// its later passes keep that order, so a fitted bin contributes to the output lists (S:264-284)
//     ground part     : its ground points in ascending z,
//     non-ground part : the points removed by R-VPF iteration 1 in ascending z, those of iteration 2, ... (S:495-504),
//                       then the points rejected by the final distance test in ascending z (S:529-541).
// The fit kernels do not sort (a K-smallest selection is all the algorithm needs, pwpp_fit.cuh) and emit both parts in
// ascending point index. This kernel re-orders the two parts of every fitted patch in `part` to the reference's order;
// ties in z keep ascending point index, which is what a stable sort of the reference's bins gives
// (oracle/_ref/libpwref_stable.so; with std::sort the order of equal z is unspecified in the reference itself).
// Sort key of a point: (group, z, position in the bin) with group 0 = ground, 1..num_iter = R-VPF iteration, 9 = final
// reject — one sort per patch yields ground part + non-ground part at once. Positions come from the labels the fit kernels
// leave (WorkQueues::labels).
//
// The sort is a bitonic network with every compare-exchange ascending (merge level k first pairs i with i ^ (k - 1), i in the lower
// half of its k-block, then i with i + j for j = k/4 .. 1), so the virtual +inf padding above n never moves. A thread owns 16
// consecutive keys in registers: steps with partner distance < 16 never leave them.
//   * k_order_warp — patches of at most 512 points (classes S and M: 83 % of the patches of a KITTI frame): ONE WARP per patch,
//     distances 16..256 are warp shuffles, no shared memory, no block barrier (the first form of this stage gave each of them a
//     512-thread CTA and ~25 barrier intervals);
//   * k_order_cta<NT> — one CTA of NT = 128 / 256 / 512 threads per patch of class L1 / L2 / L3 (16 NT keys): distances >= 16 go
//     through shared memory behind a block barrier (54 barrier intervals for 8192 keys; the plain shared-memory network has 91);
//     the CTA is sized to the class so that no warp idles at the barriers;
//   * class X (dense sensors, more than 8192 points): sorted in place in global memory through an index array (k_order_cta<512>).
// r02 on 1024 KITTI-shaped frames (profiles/README.md): one 512-thread CTA per patch for everything 8.0 ms; 512-key register
// blocks per warp with 16-warp CTAs for every patch above 512 points 14.5 ms (one CTA per SM at 92 registers, 4 of 16 warps busy on
// an L1 patch: rejected); this form: see profiles/README.md.
#pragma once
#include "pwpp_fit.cuh"

namespace pwpp {

constexpr int ORD_CAP = 8192;       // keys in shared memory
constexpr int ORD_NUM_HEADS = 5;    // ticket counters of the five launches: classes X, L3, L2, L1 and the warp-sorted classes M + S
__host__ __device__ constexpr size_t ord_cta_smem_bytes(int nt) { return (size_t) nt * 17 * sizeof(unsigned long long); }   // 16 keys + one pad key per thread

__device__ __forceinline__ unsigned long long order_sort_key(float z, unsigned char label, unsigned pos) {
  const unsigned grp = label == PW_LABEL_GROUND ? 0u : (label == PW_LABEL_REJECT ? 9u : (unsigned) label);
  // (z + 0.0f: -0.0 and +0.0 compare equal in the reference's float comparison, S:6)
  return ((unsigned long long) grp << 56) | ((unsigned long long) order_key(z + 0.0f) << 24) | (unsigned long long) (pos & 0xffffffu);
}

typedef unsigned long long OrdKey;

__device__ __forceinline__ void ord_cex(OrdKey& a, OrdKey& b) { if (a > b) { const OrdKey t = a; a = b; b = t; } }

// steps j = 8, 4, 2, 1 of a level: inside a lane's 16 keys
__device__ __forceinline__ void ord_local_tail(OrdKey (&r)[16]) {
#pragma unroll
  for (int j = 8; j > 0; j >>= 1) {
#pragma unroll
    for (int e = 0; e < 16; ++e) if (!(e & j)) ord_cex(r[e], r[e | j]);
  }
}
// levels 2..16 entirely inside a lane
__device__ __forceinline__ void ord_local_presort(OrdKey (&r)[16]) {
#pragma unroll
  for (int k = 2; k <= 16; k <<= 1) {
#pragma unroll
    for (int e = 0; e < 16; ++e) { const int l = e ^ (k - 1); if (l > e && (e & (k - 1)) < (k >> 1)) ord_cex(r[e], r[l]); }
#pragma unroll
    for (int j = k >> 2; j > 0; j >>= 1) {
#pragma unroll
      for (int e = 0; e < 16; ++e) if (!(e & j)) ord_cex(r[e], r[e | j]);
    }
  }
}
// step "i with i + j" for j = 16 m (m = 1..16): key e of a lane meets key e of lane ^ m; the lane with bit m clear keeps the smaller
__device__ __forceinline__ void ord_lane_step(OrdKey (&r)[16], int m, int lane) {
  const bool upper = (lane & m) != 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const OrdKey o = __shfl_xor_sync(0xffffffffu, r[e], m);
    const bool take = upper ? (o > r[e]) : (o < r[e]);
    r[e] = take ? o : r[e];
  }
}
// first step of level k (32 <= k <= 512): i meets i ^ (k - 1) = key 15 - e of lane ^ (k/16 - 1); the upper half of a k-block keeps the larger
__device__ __forceinline__ void ord_lane_flip(OrdKey (&r)[16], int k, int lane) {
  const int mm = (k >> 4) - 1;
  const bool upper = (lane & (k >> 5)) != 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const OrdKey a = __shfl_xor_sync(0xffffffffu, r[15 - e], mm);   // the partner's key 15 - e
    const OrdKey b = __shfl_xor_sync(0xffffffffu, r[e], mm);        // the partner's key e
    const bool ta = upper ? (a > r[e]) : (a < r[e]);
    const bool tb = upper ? (b > r[15 - e]) : (b < r[15 - e]);
    r[e] = ta ? a : r[e];
    r[15 - e] = tb ? b : r[15 - e];
  }
}
// a warp's 512 keys, sorted up to level kmax (a power of two in [16, 512]: blocks of kmax keys come out sorted)
__device__ __forceinline__ void ord_warp_sort(OrdKey (&r)[16], int kmax, int lane) {
  ord_local_presort(r);
  for (int k = 32; k <= kmax; k <<= 1) {
    ord_lane_flip(r, k, lane);
    for (int m = k >> 6; m > 0; m >>= 1) ord_lane_step(r, m, lane);
    ord_local_tail(r);
  }
}

// One CTA of NT threads per patch of queue CLS (items: make_work_item format), persistent; `head` is this launch's ticket counter.
template <int NT, int CLS>
__global__ void __launch_bounds__(NT) k_order_cta(const float4* __restrict__ sorted, WorkQueues wq, int* __restrict__ head, int* __restrict__ part) {
  PW_DYN_SHARED(unsigned long long, s_key);   // [NT * 17]
  __shared__ int s_t;
  constexpr int CAP = NT * 16;
  const int tid = threadIdx.x;
  const int total = wq.count[CLS];
  auto at = [&](int i) -> OrdKey& { return s_key[i + (i >> 4)]; };   // the pad keeps a thread's 16-key block off its neighbours' banks
  for (;;) {
    __syncthreads();
    if (tid == 0) s_t = atomicAdd(head, 1);
    __syncthreads();
    const int t = s_t;
    if (t >= total) return;
    const int4 wi = wq.items[CLS][t];
    const int n = wi.y;
    const long long start = work_item_start(wi);
    const float4* P = sorted + start;
    const unsigned char* L = wq.labels + start;
    int* out = part + start;
    if (n <= CAP) {
      for (int i = tid; i < n; i += NT) at(i) = order_sort_key(P[i].z, L[i], (unsigned) i);   // coalesced
      __syncthreads();
      int n2 = 16;
      while (n2 < n) n2 <<= 1;
      const int npairs = n2 >> 1;
      const int base = tid << 4;
      const bool own = base < n;   // blocks at or above n hold only the virtual +inf padding, which never moves
      OrdKey r[16];
      auto load_block = [&]() {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = (base + e < n) ? at(base + e) : ~0ull;
      };
      auto store_block = [&]() {
#pragma unroll
        for (int e = 0; e < 16; ++e) if (base + e < n) at(base + e) = r[e];
      };
      if (own) { load_block(); ord_local_presort(r); store_block(); }
      __syncthreads();
      auto cex = [&](int i, int l) {
        if (l < n) {
          OrdKey &pa = at(i), &pb = at(l);
          const OrdKey a = pa, b = pb;
          if (a > b) { pa = b; pb = a; }
        }
      };
      for (int k = 32, lk = 5; k <= n2; k <<= 1, ++lk) {
        const int hk = k >> 1;
#pragma unroll 4
        for (int q = tid; q < npairs; q += NT) { const int i = ((q >> (lk - 1)) << lk) | (q & (hk - 1)); cex(i, i ^ (k - 1)); }
        __syncthreads();
        for (int j = k >> 2; j >= 16; j >>= 1) {
#pragma unroll 4
          for (int q = tid; q < npairs; q += NT) { const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1)); cex(i, i | j); }
          __syncthreads();
        }
        if (own) { load_block(); ord_local_tail(r); store_block(); }
        __syncthreads();
      }
      for (int i = tid; i < n; i += NT) out[i] = __float_as_int(P[(int) (at(i) & 0xffffffull)].w);
    } else {
      // in place in global memory: `out` holds positions, compared through their keys; same network
      for (int i = tid; i < n; i += NT) out[i] = i;
      __syncthreads();
      for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int j = k - 1; j > 0; j = (j == k - 1) ? (k >> 2) : (j >> 1)) {
          for (int i = tid; i < n; i += NT) {
            const int l = i ^ j;
            if (l > i && l < n) {
              const int pa = out[i], pb = out[l];
              const unsigned long long a = order_sort_key(P[pa].z, L[pa], (unsigned) pa), b = order_sort_key(P[pb].z, L[pb], (unsigned) pb);
              if (a > b) { out[i] = pb; out[l] = pa; }
            }
          }
          __syncthreads();
          if (j == 0) break;
        }
      }
      for (int i = tid; i < n; i += NT) out[i] = __float_as_int(P[out[i]].w);
    }
  }
}

// Classes M and S: every warp on its own, one patch per ticket, keys straight into registers.
constexpr int ORD_WARP_THREADS = 128;
__global__ void __launch_bounds__(ORD_WARP_THREADS) k_order_warp(const float4* __restrict__ sorted, WorkQueues wq, int* __restrict__ head, int* __restrict__ part) {
  static_assert(CLS_M_MAX <= 512, "classes S and M are sorted by one warp per patch (16 keys per lane)");
  const int lane = threadIdx.x & 31;
  const int n_m = wq.count[1], total = n_m + wq.count[0];
  OrdKey r[16];
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(head, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= total) return;
    const int4 wi = t < n_m ? wq.items[1][t] : wq.items[0][t - n_m];
    const int n = wi.y;
    const long long start = work_item_start(wi);
    const float4* P = sorted + start;
    const unsigned char* L = wq.labels + start;
    int* out = part + start;
    const int base = lane << 4;
#pragma unroll
    for (int e = 0; e < 16; ++e) { const int i = base + e; r[e] = i < n ? order_sort_key(P[i].z, L[i], (unsigned) i) : ~0ull; }
    int n2 = 16;
    while (n2 < n) n2 <<= 1;
    ord_warp_sort(r, n2, lane);
#pragma unroll
    for (int e = 0; e < 16; ++e) { const int i = base + e; if (i < n) out[i] = __float_as_int(P[(int) (r[e] & 0xffffffull)].w); }
  }
}

}  // namespace pwpp
