// pwpp_gle.cuh — data of the per-frame sequential stage of estimateGround() (A-GLE verdicts, Temporal Ground
// Revert, adaptive threshold / sensor-height update, layout of the output lists) and its two scalar helpers.
// k_gle (pwpp_kernels.cuh) is the product implementation; the plain sequential statement of the same stage that
// the CPU twin runs lives with the tests (tests/gle_sequential.cuh).
//
// Reference: cpp/patchworkpp/src/patchworkpp.cpp ("S:") 211-311 (A-GLE block and per-ring TGR call),
// 402-464 (temporal_ground_revert), 338-375 (update_elevation_thr / update_flatness_thr),
// 557-566 (calc_mean_stdev). Every quirk listed in SURVEY.md Appendix C is kept.
#pragma once
#include "pwpp_math.cuh"

namespace pwpp {

// Temporal state of one stream (reference members mutated between frames, H:169,174-175)
struct StreamState {
  double sensor_height;        // params_.sensor_height (adaptive, S:348)
  double elevation_thr[4];     // params_.elevation_thr (S:347,350)
  double flatness_thr[4];      // params_.flatness_thr  (S:368)
  int n_elev[4], n_flat[4];    // sizes of update_elevation_/update_flatness_
  double stale_mean[3], stale_normal[3], stale_sv[3];  // last estimate_plane() result (S:49 carry-over)
};

// What k_fit leaves for k_gle, one per (frame, bin)
struct BinFit {   // same layout as pwpp_bin_result (include/pwpp.h)
  double mean[3], normal[3], sv[3], d;
  int n, n_ground, verdict, fitted;
};
#define PW_FIT_NO_PLANE (-1)  /* BinFit.verdict marker from k_fit: estimate_plane never saw a non-empty set */

// Where k_emit must copy a bin's two parts (positions inside the frame's output region)
struct BinSeg {
  int g_dst;    // destination of part[0, n_ground)           (< 0: nothing to copy)
  int ng_dst;   // destination of part[n_ground, n) (un-reversed; < 0: dropped)
};

// calc_mean_stdev S:557-566: mean / stdev are left untouched when n <= 1
PW_HD void calc_mean_stdev(const double* v, int n, double& mean, double& stdev) {
  if (n <= 1) return;
  double s = 0.0;
  for (int i = 0; i < n; ++i) s = dadd(s, v[i]);
  mean = ddiv(s, (double) n);
  for (int i = 0; i < n; ++i) { const double d = dsub(v[i], mean); stdev = dadd(stdev, dmul(d, d)); }
  stdev = ddiv(stdev, (double) (n - 1));
  stdev = dsqrt(stdev);
}

PW_HD void history_push(double* row, int& n, int hcap, double v) {
  // The reference's vectors are unbounded; a history only grows past max_*_storage + one frame while a
  // lower ring blocks the flatness update (S:363-364). When the fixed-capacity row is full the oldest
  // sample is dropped.
  if (n == hcap) { for (int q = 1; q < n; ++q) row[q - 1] = row[q]; --n; }
  row[n++] = v;
}

}  // namespace pwpp
