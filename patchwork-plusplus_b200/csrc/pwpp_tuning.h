// pwpp_tuning.h — compiled-in defaults of the few remaining switches (each can be overridden by the environment variable of
// the same name without the _DEFAULT suffix when a context is created). Everything else that was a switch in round 1 was
// measured on a B200 in round 2 (profiles/r02_baseline/tune.jsonl) and either became the one code path or was deleted.
#pragma once
#define PWPP_FRONT_DEFAULT 1       /* 1: cluster-per-frame front end (pwpp_front.cuh); 0: k_bin_hist + k_bin_scan + k_scatter */
#define PWPP_FIT_PATCH_DEFAULT 0   /* 1: patches above 512 points on k_fit_patch (pwpp_fit_patch.cuh) instead of k_fit_warp<L1> / k_fit_cta */
#define PWPP_SMALL_CALL_DEFAULT 4  /* calls of at most this many frames: CTA-per-patch fit kernels above 512 points and the three stand-alone front-end kernels (latency path) */
