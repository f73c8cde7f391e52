// tests/simt/cuda_runtime.h — TEST-ONLY stand-in for <cuda_runtime.h>: runs the product's CUDA kernels on the CPU.
//
// The build container has no GPU. This header lets g++ compile csrc/*.cuh unchanged (it is found first on the include
// path of tests/simt/simt_twin.cpp) and executes a kernel launch as follows:
//   * CTAs run one after another on the calling OS thread (so `__shared__` variables may be plain statics); a kernel whose
//     CTAs wait for each other (k_front) is launched with simt::launch_concurrent: all CTAs live at once, fibers of all
//     of them scheduled round-robin — such a kernel must keep everything in dynamic shared memory;
//   * every thread of a CTA is a fiber with its own stack (hand-written x86-64 context switch); a fiber runs until it
//     reaches a warp collective or a block barrier, where it waits for the other participants round-robin;
//   * __shfl*_sync / __ballot_sync / __any_sync / __match_any_sync / __reduce_*_sync / __syncwarp are rendezvous of the
//     lanes named in the mask (several disjoint masks per warp may be in flight, as after __match_any_sync);
//     __syncthreads is a rendezvous of all threads of the CTA that have not returned;
//   * a rendezvous that can never complete (a lane returned or waits elsewhere) is reported as a deadlock, with the
//     kernel name, instead of hanging.
// What this checks: the kernels' control flow, index arithmetic, queue protocols and numerics (the same double-precision
// code as on the device: csrc/pwpp_math.cuh has host branches for every device intrinsic it uses).
// What it does not check: memory-model races (fibers only switch at collectives), launch bounds, occupancy, speed.
// Nothing in the product includes this file; lib/libpwpp_b200.so is built by nvcc against the real CUDA headers.
#pragma once
#define PWPP_SIMT_EMU 1

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <string>
#include <vector>
// (every standard header the twin needs is included above: the CUDA spelling __noinline__ defined below would break
//  libstdc++'s own __attribute__((__noinline__)) in headers included later)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
struct int4 { int x, y, z, w; } __attribute__((aligned(16)));
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

namespace simt {

constexpr int MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 256 * 1024;

struct Fiber {
  void* sp = nullptr;     // saved stack pointer
  char* stack = nullptr;
  bool done = true;
};

// one in-flight rendezvous of a warp, keyed by its lane mask
struct Slot {
  unsigned mask = 0;
  int arrived = 0, departed = 0;
  bool draining = false;
  unsigned long long gen = 0;
  unsigned long long in[32];
  int aux[32];
  unsigned long long out[32];
};
struct WarpState {
  Slot slots[33];
  int nslots = 0;
};

struct Cta {
  Fiber fib[MAX_THREADS];
  WarpState warps[MAX_THREADS / 32];
  int nthreads = 0, alive = 0;
  int cur = 0;                 // running fiber
  void* main_sp = nullptr;
  // block barrier
  int bar_arrived = 0;
  unsigned long long bar_gen = 0;
  // progress watchdog
  unsigned long long progress = 0;
  const char* kernel_name = "";
  std::function<void()> body;
  std::vector<char> dyn_smem;
  uint3 bid = {0, 0, 0};       // blockIdx of this CTA (concurrent launches)
};

extern Cta* g_cta_p;            // the CTA of the running fiber
#define g_cta (*simt::g_cta_p)
extern uint3 g_tid[MAX_THREADS];
extern bool g_concurrent;       // launch_concurrent(): all CTAs of the grid are live at once (kernels that wait for each other)

extern "C" void simt_switch(void** save_sp, void* load_sp);
void yield_();
[[noreturn]] void deadlock(const char* what);
void launch_impl(const char* name, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void launch_concurrent_impl(const char* name, int nctas, dim3 block, size_t smem, const std::function<void()>& body, unsigned block_y = 0);
void cluster_sync_();                       // barrier over all threads of all live CTAs (a cluster = the CTAs of one concurrent launch)
void* cluster_map_(void* p, int rank);      // the same dynamic-shared-memory location in CTA `rank`

}  // namespace simt

// ---- built-in variables ----
extern uint3 threadIdx;   // rewritten on every fiber switch
extern uint3 blockIdx;
extern dim3 blockDim, gridDim;

static inline void* simt_dyn_smem() { return g_cta.dyn_smem.data(); }
// `extern __shared__ T name[];` of the device build
#define PW_DYN_SHARED(T, name) T* name = reinterpret_cast<T*>(simt_dyn_smem())

// ---- rendezvous core ----
namespace simt {
enum Op { OP_SYNC, OP_SHFL, OP_BALLOT, OP_MATCH, OP_ADD, OP_MIN, OP_MAX, OP_OR, OP_MINS, OP_MAXS };

static inline unsigned long long rendezvous(unsigned mask, Op op, unsigned long long val, int aux) {
  Cta& c = g_cta;
  const int t = c.cur, lane = t & 31;
  if (!((mask >> lane) & 1u)) { std::fprintf(stderr, "simt: lane %d calls a collective whose mask 0x%08x excludes it (%s)\n", lane, mask, c.kernel_name); std::abort(); }
  WarpState& w = c.warps[t >> 5];
  Slot* s = nullptr;
  for (;;) {
    s = nullptr;
    for (int i = 0; i < w.nslots; ++i) if (w.slots[i].mask == mask) { s = &w.slots[i]; break; }
    if (s && s->draining) { yield_(); continue; }   // the previous rendezvous on this mask has not been read by all lanes yet
    break;
  }
  if (!s) {
    for (int i = 0; i < w.nslots; ++i) if (w.slots[i].arrived == 0 && !w.slots[i].draining) { s = &w.slots[i]; break; }
    if (!s) { if (w.nslots >= 33) deadlock("too many distinct masks in flight"); s = &w.slots[w.nslots++]; }
    s->mask = mask; s->arrived = 0; s->departed = 0; s->draining = false;
  }
  s->in[lane] = val; s->aux[lane] = aux;
  const int need = __builtin_popcount(mask);
  ++c.progress;
  if (++s->arrived == need) {
    // the last lane computes everybody's result
    unsigned long long acc = 0;
    bool first = true;
    for (int l = 0; l < 32; ++l) {
      if (!((mask >> l) & 1u)) continue;
      const unsigned long long v = s->in[l];
      switch (op) {
        case OP_ADD: acc = first ? v : (unsigned long long) ((unsigned) acc + (unsigned) v); break;
        case OP_MIN: acc = first ? v : ((unsigned) v < (unsigned) acc ? v : acc); break;
        case OP_MAX: acc = first ? v : ((unsigned) v > (unsigned) acc ? v : acc); break;
        case OP_MINS: acc = first ? v : ((int) v < (int) acc ? v : acc); break;
        case OP_MAXS: acc = first ? v : ((int) v > (int) acc ? v : acc); break;
        case OP_OR: acc = first ? v : (acc | v); break;
        case OP_BALLOT: if (v) acc |= 1ull << l; break;
        default: break;
      }
      first = false;
    }
    for (int l = 0; l < 32; ++l) {
      if (!((mask >> l) & 1u)) continue;
      switch (op) {
        case OP_SYNC: s->out[l] = 0; break;
        case OP_SHFL: {
          const int src = s->aux[l];
          s->out[l] = (src >= 0 && src < 32 && ((mask >> src) & 1u)) ? s->in[src] : s->in[l];
          break;
        }
        case OP_MATCH: {
          unsigned m = 0;
          for (int k = 0; k < 32; ++k) if (((mask >> k) & 1u) && s->in[k] == s->in[l]) m |= 1u << k;
          s->out[l] = m;
          break;
        }
        default: s->out[l] = acc; break;
      }
    }
    s->draining = true; s->departed = 0; ++s->gen;
  } else {
    const unsigned long long g = s->gen;
    while (s->gen == g) yield_();
  }
  const unsigned long long r = s->out[lane];
  if (++s->departed == need) { s->draining = false; s->arrived = 0; s->mask = 0; }
  return r;
}
}  // namespace simt

// ---- warp collectives ----
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::rendezvous(mask, simt::OP_SYNC, 0, 0); }
static inline unsigned __ballot_sync(unsigned mask, int pred) { return (unsigned) simt::rendezvous(mask, simt::OP_BALLOT, pred ? 1 : 0, 0); }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __match_any_sync(unsigned mask, int v) { return (unsigned) simt::rendezvous(mask, simt::OP_MATCH, (unsigned long long) (unsigned) v, 0); }
static inline unsigned __match_any_sync(unsigned mask, unsigned v) { return (unsigned) simt::rendezvous(mask, simt::OP_MATCH, v, 0); }
static inline int __reduce_add_sync(unsigned mask, int v) { return (int) (unsigned) simt::rendezvous(mask, simt::OP_ADD, (unsigned) v, 0); }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return (unsigned) simt::rendezvous(mask, simt::OP_ADD, v, 0); }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return (unsigned) simt::rendezvous(mask, simt::OP_MIN, v, 0); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return (unsigned) simt::rendezvous(mask, simt::OP_MAX, v, 0); }
static inline int __reduce_min_sync(unsigned mask, int v) { return (int) (unsigned) simt::rendezvous(mask, simt::OP_MINS, (unsigned) v, 0); }
static inline int __reduce_max_sync(unsigned mask, int v) { return (int) (unsigned) simt::rendezvous(mask, simt::OP_MAXS, (unsigned) v, 0); }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return (unsigned) simt::rendezvous(mask, simt::OP_OR, v, 0); }

namespace simt {
template <typename T>
static inline T shfl_bits(unsigned mask, T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  unsigned long long b = 0;
  std::memcpy(&b, &v, sizeof(T));
  b = rendezvous(mask, OP_SHFL, b, src);
  T r;
  std::memcpy(&r, &b, sizeof(T));
  return r;
}
static inline int cur_lane() { return g_cta.cur & 31; }
}  // namespace simt
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = simt::cur_lane();
  return simt::shfl_bits(mask, v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  (void) width;
  return simt::shfl_bits(mask, v, simt::cur_lane() ^ lanemask);
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = simt::cur_lane();
  const int src = lane - (int) delta;
  return simt::shfl_bits(mask, v, (src < (lane & ~(width - 1))) ? lane : src);
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = simt::cur_lane();
  const int src = lane + (int) delta;
  return simt::shfl_bits(mask, v, (src > (lane | (width - 1))) ? lane : src);
}

// ---- block barrier: all threads of the CTA that have not returned ----
static inline void __syncthreads() {
  simt::Cta& c = g_cta;
  ++c.progress;
  const unsigned long long g = c.bar_gen;
  if (++c.bar_arrived >= c.alive) { c.bar_arrived = 0; ++c.bar_gen; return; }
  while (c.bar_gen == g) {
    simt::yield_();
    // threads that returned since we arrived no longer count
    if (c.bar_gen == g && c.bar_arrived >= c.alive) { c.bar_arrived = 0; ++c.bar_gen; }
  }
}

// ---- atomics (one OS thread: plain read-modify-write) ----
static inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline int atomicMax(int* p, int v) { const int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int* p, int v) { const int o = *p; if (v < o) *p = v; return o; }
static inline int atomicExch(int* p, int v) { const int o = *p; *p = v; return o; }
static inline int atomicCAS(int* p, int cmp, int v) { const int o = *p; if (o == cmp) *p = v; return o; }

// ---- bit / conversion intrinsics ----
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned) v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long u; std::memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; std::memcpy(&d, &u, 8); return d; }
template <typename T> static inline T __ldg(const T* p) { return *p; }

using std::fabs;
using std::fmaf;
using std::sqrt;
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }

// launch: simt::launch("name", grid, block, dyn_smem_bytes, [&] { kernel(args...); });
namespace simt {
template <typename F>
static inline void launch(const char* name, dim3 grid, dim3 block, size_t smem, F&& f) { launch_impl(name, grid, block, smem, std::function<void()>(f)); }
template <typename F>
static inline void launch_concurrent(const char* name, int nctas, dim3 block, size_t smem, F&& f, unsigned block_y = 0) { launch_concurrent_impl(name, nctas, block, smem, std::function<void()>(f), block_y); }
}
