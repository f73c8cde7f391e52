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
// leave (WorkQueues::labels). Patches up to ORD_CAP points are sorted in shared memory (64-bit keys, bitonic network with
// all compare-exchanges ascending, so the virtual +inf padding above n never moves and is never stored); larger patches
// (class X, dense sensors) are sorted in place in global memory through an index array with the same network.
#pragma once
#include "pwpp_fit.cuh"

namespace pwpp {

constexpr int ORD_CAP = 8192;       // keys in shared memory
constexpr int ORD_SMEM_KEYS = ORD_CAP + ORD_CAP / 16;   // one pad key per 16-key block (68 KB)
constexpr int ORD_THREADS = 512;

__device__ __forceinline__ unsigned long long order_sort_key(float z, unsigned char label, unsigned pos) {
  const unsigned grp = label == PW_LABEL_GROUND ? 0u : (label == PW_LABEL_REJECT ? 9u : (unsigned) label);
  // (z + 0.0f: -0.0 and +0.0 compare equal in the reference's float comparison, S:6)
  return ((unsigned long long) grp << 56) | ((unsigned long long) order_key(z + 0.0f) << 24) | (unsigned long long) (pos & 0xffffffu);
}

// One CTA per fitted patch, persistent over all class queues (items: make_work_item format).
__global__ void __launch_bounds__(ORD_THREADS) k_order(const float4* __restrict__ sorted, WorkQueues wq, int* __restrict__ order_head, int* __restrict__ part) {
  PW_DYN_SHARED(unsigned long long, s_key);   // [ORD_SMEM_KEYS]
  __shared__ int s_t;
  const int tid = threadIdx.x;
  int cum[NUM_CLASSES + 1];   // tickets run over the queues from the largest size class to the smallest (long sorts first)
  cum[0] = 0;
#pragma unroll
  for (int c = 0; c < NUM_CLASSES; ++c) cum[c + 1] = cum[c] + wq.count[NUM_CLASSES - 1 - c];
  const int total = cum[NUM_CLASSES];
  for (;;) {
    __syncthreads();
    if (tid == 0) s_t = atomicAdd(order_head, 1);
    __syncthreads();
    const int t = s_t;
    if (t >= total) return;
    int c = 0;
#pragma unroll
    for (int q = 1; q < NUM_CLASSES; ++q) if (t >= cum[q]) c = q;
    const int4 wi = wq.items[NUM_CLASSES - 1 - c][t - cum[c]];
    const int n = wi.y;
    const long long start = work_item_start(wi);
    const float4* P = sorted + start;
    const unsigned char* L = wq.labels + start;
    int* out = part + start;
    if (n <= ORD_CAP) {
      // Bitonic network with every compare-exchange ascending: merge level k first pairs i with i ^ (k - 1) (i in the lower half
      // of its k-block), then i with i + j for j = k/4 .. 1. Thread t OWNS the 16 consecutive keys [16 t, 16 t + 16): every step
      // whose partner distance is below 16 — all of levels 2..16 and the last four steps of every later level — runs in its
      // registers without a barrier, the other steps go through shared memory (one loop iteration per PAIR). 8192 keys: 54
      // barrier intervals instead of the 91 of the plain network (a one-frame call waits for the sort of its largest patch).
      // Shared-memory index of key i is i + (i >> 4): the pad keeps a thread's 16-key block off its neighbours' banks.
      auto at = [&](int i) -> unsigned long long& { return s_key[i + (i >> 4)]; };
      for (int i = tid; i < n; i += ORD_THREADS) at(i) = order_sort_key(P[i].z, L[i], (unsigned) i);
      __syncthreads();
      int n2 = 16;
      while (n2 < n) n2 <<= 1;
      const int npairs = n2 >> 1;
      const int base = tid << 4;
      const bool own = base < n;   // blocks at or above n hold only the virtual +inf padding, which never moves
      unsigned long long r[16];
      auto cexr = [&](int a, int b) { if (r[a] > r[b]) { const unsigned long long t = r[a]; r[a] = r[b]; r[b] = t; } };
      auto load_block = [&]() {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = (base + e < n) ? at(base + e) : ~0ull;
      };
      auto store_block = [&]() {
#pragma unroll
        for (int e = 0; e < 16; ++e) if (base + e < n) at(base + e) = r[e];
      };
      auto local_tail = [&]() {   // steps j = 8, 4, 2, 1 of a level
#pragma unroll
        for (int j = 8; j > 0; j >>= 1) {
#pragma unroll
          for (int e = 0; e < 16; ++e) if (!(e & j)) cexr(e, e | j);
        }
      };
      if (own) {
        load_block();
#pragma unroll
        for (int k = 2; k <= 16; k <<= 1) {   // levels 2..16 entirely in registers
#pragma unroll
          for (int e = 0; e < 16; ++e) { const int l = e ^ (k - 1); if (l > e && (e & (k - 1)) < (k >> 1)) cexr(e, l); }
#pragma unroll
          for (int j = k >> 2; j > 0; j >>= 1) {
#pragma unroll
            for (int e = 0; e < 16; ++e) if (!(e & j)) cexr(e, e | j);
          }
        }
        store_block();
      }
      __syncthreads();
      auto cex = [&](int i, int l) {
        if (l < n) {
          unsigned long long &pa = at(i), &pb = at(l);
          const unsigned long long a = pa, b = pb;
          if (a > b) { pa = b; pb = a; }
        }
      };
      for (int k = 32, lk = 5; k <= n2; k <<= 1, ++lk) {
        const int hk = k >> 1;
#pragma unroll 4
        for (int q = tid; q < npairs; q += ORD_THREADS) { const int i = ((q >> (lk - 1)) << lk) | (q & (hk - 1)); cex(i, i ^ (k - 1)); }
        __syncthreads();
        for (int j = k >> 2; j >= 16; j >>= 1) {
#pragma unroll 4
          for (int q = tid; q < npairs; q += ORD_THREADS) { const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1)); cex(i, i | j); }
          __syncthreads();
        }
        if (own) { load_block(); local_tail(); store_block(); }
        __syncthreads();
      }
      for (int i = tid; i < n; i += ORD_THREADS) out[i] = __float_as_int(P[(int) (at(i) & 0xffffffull)].w);
    } else {
      // in place in global memory: `out` holds positions, compared through their keys; same network
      for (int i = tid; i < n; i += ORD_THREADS) out[i] = i;
      __syncthreads();
      for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int j = k - 1; j > 0; j = (j == k - 1) ? (k >> 2) : (j >> 1)) {
          for (int i = tid; i < n; i += ORD_THREADS) {
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
      for (int i = tid; i < n; i += ORD_THREADS) out[i] = __float_as_int(P[out[i]].w);
    }
  }
}

}  // namespace pwpp
