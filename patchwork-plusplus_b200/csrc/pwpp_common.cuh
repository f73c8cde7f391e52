// pwpp_common.cuh — small device helpers and launch geometry shared by all kernels.
#pragma once
#include <cuda_runtime.h>

#include "pwpp_gle.cuh"
#include "pwpp_math.cuh"

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
// ---- thread-block clusters (k_front_cluster): rank, barrier, distributed shared memory ----
#if defined(PWPP_SIMT_EMU)
__device__ __forceinline__ unsigned pw_cluster_rank() { return blockIdx.x; }                 // the twin launches one cluster at a time, blockIdx.x = rank
__device__ __forceinline__ void pw_cluster_sync() { simt::cluster_sync_(); }
template <typename T> __device__ __forceinline__ T* pw_cluster_map(T* p, unsigned rank) { return static_cast<T*>(simt::cluster_map_(p, (int) rank)); }
#else
}  // namespace pwpp
#include <cooperative_groups.h>
namespace pwpp {
__device__ __forceinline__ unsigned pw_cluster_rank() { return cooperative_groups::this_cluster().block_rank(); }
__device__ __forceinline__ void pw_cluster_sync() { cooperative_groups::this_cluster().sync(); }
template <typename T> __device__ __forceinline__ T* pw_cluster_map(T* p, unsigned rank) { return cooperative_groups::this_cluster().map_shared_rank(p, rank); }
#endif

#if defined(PWPP_SIMT_EMU)   // tests/simt: the kernels compiled by g++ and run lane by lane on the CPU (test infrastructure only)
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) { return *p; }
__device__ __forceinline__ void prefetch_l2(const void*) {}
#else
// dynamic shared memory of a kernel, typed (tests/simt/cuda_runtime.h defines the CPU counterpart)
#define PW_DYN_SHARED(T, name) extern __shared__ T name[]
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  // read-once data: bypass L1 allocation, keep L2 normal
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

// ---- bulk async copy global -> shared::cta (the TMA engine; a contiguous range needs no tensor map; SASS: UBLKCP), completion
// on an mbarrier ----
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


}  // namespace pwpp
