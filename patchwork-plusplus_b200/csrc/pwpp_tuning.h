// pwpp_tuning.h — compiled-in defaults of the kernel-variant switches (each can be overridden by the environment variable
// of the same name without the _DEFAULT suffix when a context is created). tools/apply_chosen.py writes the values of a
// gpurun_out/chosen.env measured by tools/gpu_tune.py on a B200. r01: chosen.env was FUSE_SEED=1 (then one switch for all
// classes), L2_MINB=3, X_KERNEL=1, X_MINB=1 and the GPU parity suite passed under it; the per-stage times of that run
// (profiles/r01_tune.jsonl) show the fused WARP kernels slower than the unfused ones, so fusion is on for the CTA kernels only.
#pragma once
#define PWPP_FIT_GROUP_DEFAULT 0   /* 1: the group fit kernel (pwpp_fit_group.cuh) serves every patch up to 8192 points; 0: the size-classed kernels of pwpp_fit.cuh */
#define PWPP_HIST_PIPE_DEFAULT 2    /* k_bin_hist load pipelining: 0 none, 1 groups of 4, 2 groups of 2 */
#define PWPP_SCATTER_V_DEFAULT 0    /* 1: software-pipelined k_scatter at 3 CTAs/SM */
#define PWPP_SERIAL_FIT_DEFAULT 0   /* 1: the fit kernels run one after another on the call's stream instead of forked onto side streams */
#define PWPP_S_MINB_DEFAULT 2       /* launch-bounds CTAs/SM of the class-S kernel */
#define PWPP_M_MINB_DEFAULT 2
#define PWPP_L1_MINB_DEFAULT 2
#define PWPP_L2_MINB_DEFAULT 3
#define PWPP_L2_NW_DEFAULT 8        /* warps per patch of the class-L2 CTA kernel */
#define PWPP_L3_NW_DEFAULT 8
#define PWPP_FUSE_SEED_DEFAULT 1    /* bit mask: R-VPF + R-GPF seed fit of zone-0 patches from one selection and one pass; 1 = CTA kernels (L2, L3, X), 2 = warp kernels (M, L1) */
#define PWPP_SOLVE_CALL_DEFAULT 0   /* 1: the warp kernels call one out-of-line plane solver (instruction-cache footprint; not yet measured) */
#define PWPP_FRONT_DEFAULT 0        /* 1: binning + scan + scatter as one persistent kernel pipelined through L2 (pwpp_front.cuh); checked on the SIMT twin, not yet measured */
#define PWPP_L2_WIDE_DEFAULT 0      /* 1: class L2 covers 2049..5888 points (k_fit_cta<5888> still at 3 CTAs/SM), class L3 the rest up to 8192; checked on the SIMT twin, not yet measured */
#define PWPP_M_RESIDENT_DEFAULT 0   /* 1: class M on the register-resident kernel k_fit_resident<32,16>; checked on the SIMT twin, not yet measured */
#define PWPP_M_HALF_DEFAULT 0       /* 1: class M = 65..256 points on k_fit_resident<16,16> (two patches per warp), 257..512 joins class L1; checked on the SIMT twin, not yet measured */
#define PWPP_L1_CTA_DEFAULT 0       /* 1: class L1 on the fused CTA kernel k_fit_cta<2048>; checked on the SIMT twin, not yet measured */
#define PWPP_L2_PLS_DEFAULT 0       /* 1: class-L2 kernel keeps the current plane in shared memory (fewer spills); checked on the SIMT twin, not yet measured */
#define PWPP_PART_ILP_DEFAULT 0     /* 1: four index loads in flight in the final partition of the M/L1/L2/L3 kernels; checked on the SIMT twin, not yet measured */
#define PWPP_EMIT_SPLIT_DEFAULT 1    /* k_emit: slices per bin (grid.z); > 1 written for dense frames, checked on the SIMT twin, not yet measured */
#define PWPP_X_KERNEL_DEFAULT 1     /* class X (> 8192 points): 1 = CTA per patch (k_fit_big), 0 = one warp per patch */
#define PWPP_X_FIXPOINT_DEFAULT 0   /* 1: k_fit_big ends the R-GPF passes at the exact fixpoint (set recorded as ballot words); checked on the SIMT twin, not yet measured */
#define PWPP_X_NW_DEFAULT 16
#define PWPP_X_MINB_DEFAULT 1
