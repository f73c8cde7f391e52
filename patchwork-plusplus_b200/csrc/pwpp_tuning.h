// pwpp_tuning.h — compiled-in defaults of the kernel-variant switches (each can be overridden by the environment variable
// of the same name without the _DEFAULT suffix when a context is created). Written by tools/apply_chosen.py from the
// gpurun_out/chosen.env that tools/gpu_tune.py measured on a B200; the GPU parity suite ran under exactly these values.
#pragma once
#define PWPP_HIST_PIPE_DEFAULT 2    /* k_bin_hist load pipelining: 0 none, 1 groups of 4, 2 groups of 2 */
#define PWPP_SCATTER_V_DEFAULT 0    /* 1: software-pipelined k_scatter at 3 CTAs/SM */
#define PWPP_S_MINB_DEFAULT 2       /* launch-bounds CTAs/SM of the class-S kernel */
#define PWPP_M_MINB_DEFAULT 2
#define PWPP_L1_MINB_DEFAULT 2
#define PWPP_L2_MINB_DEFAULT 4
#define PWPP_L2_NW_DEFAULT 8        /* warps per patch of the class-L2 CTA kernel */
#define PWPP_L3_NW_DEFAULT 8
#define PWPP_FUSE_SEED_DEFAULT 0    /* 1: R-VPF + R-GPF seed fit of zone-0 patches from one selection and one pass */
#define PWPP_X_KERNEL_DEFAULT 1     /* class X (> 8192 points): 1 = CTA per patch (k_fit_big), 0 = one warp per patch */
#define PWPP_X_NW_DEFAULT 16
#define PWPP_X_MINB_DEFAULT 2
