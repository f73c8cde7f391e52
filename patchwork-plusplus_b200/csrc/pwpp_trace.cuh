// pwpp_trace.cuh — DIAGNOSTIC instrumentation, compiled in only with -DPWPP_PHASE_CLOCKS (tools/gpu_phase_probe.py): per-phase
// cycle counters and a per-warp event trace of CTA 0 of one kernel class. In product builds every macro is empty.
#pragma once
#include <cuda_runtime.h>

namespace pwpp {
// optional phase clocks (diagnostic builds only: -DPWPP_PHASE_CLOCKS): thread 0 of every CTA accumulates the cycles it spends
// in each phase of k_fit_group into g_phase_clk[class][phase]; read with pwpp_debug_phase_clocks()
#if defined(PWPP_PHASE_CLOCKS) && !defined(PWPP_SIMT_EMU)
__device__ unsigned long long g_phase_clk[8][16];
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
}  // namespace pwpp
