// gle_sequential.cuh — TEST-ONLY sequential statement of the per-frame A-GLE / TGR / threshold stage, on the product's
// data structures and scalar helpers (csrc/pwpp_gle.cuh, csrc/pwpp_math.cuh). tests/host_twin.cu runs it on the CPU;
// the product runs k_gle (csrc/pwpp_kernels.cuh), which must take the same decisions.
//
// Reference: cpp/patchworkpp/src/patchworkpp.cpp ("S:") 211-311, 402-464, 338-375, 557-566.
#pragma once
#include "pwpp_gle.cuh"

namespace pwpp {

struct GleScratch {
  double ringflat[4096];   // ringwise_flatness (S:182): <= 4 near rings x <= 1024 sectors
  double cand_lv[1024];    // line_variable of the ring's RevertCandidates (H:34)
  int cand_bin[1024];      // their bins
};

// One frame. bo = bin offsets [nb_all+1], fit/seg per bin. Output destinations are relative to the frame's
// output region: the ground list first, the non-ground list right behind it.
// Emission order of the non-ground list (S:393,618,193,264,272,284,298,458): RNR hits, out-of-range points,
// then bins in loop order (rejected ground part before the bin's non-ground part); rejected candidates at
// the end of their ring. Ground list (S:268,276,450): accepted bins in loop order, reverted candidates at
// the end of their ring.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void gle_frame(const Geometry& g, const AlgoParams& ap, StreamState& st, double* h_elev, double* h_flat, int hcap, const int* bo, BinFit* fit,
                      BinSeg* seg, float* cen, float* nor, GleScratch& sc, int& out_num_ground, int& out_num_patches, int& out_num_dropped) {
  const int nb = g.nbins;
  int g_run = 0, ng_run = 0;
  const int n_rnr = bo[PW_BIN_RNR(nb) + 1] - bo[PW_BIN_RNR(nb)];
  const int n_oor = bo[PW_BIN_OOR(nb) + 1] - bo[PW_BIN_OOR(nb)];
  const int n_drop = bo[PW_BIN_DROP(nb) + 1] - bo[PW_BIN_DROP(nb)];
  seg[PW_BIN_RNR(nb)].g_dst = -1; seg[PW_BIN_RNR(nb)].ng_dst = ng_run; ng_run += n_rnr;
  seg[PW_BIN_OOR(nb)].g_dst = -1; seg[PW_BIN_OOR(nb)].ng_dst = ng_run; ng_run += n_oor;
  seg[PW_BIN_DROP(nb)].g_dst = -1; seg[PW_BIN_DROP(nb)].ng_dst = -1;

  int concentric_idx = 0, npatch = 0, n_ringflat = 0, ncand = 0;
  for (int zone = 0; zone < 4; ++zone) {
    for (int ring = 0; ring < g.num_rings[zone]; ++ring) {
      for (int sector = 0; sector < g.num_sectors[zone]; ++sector) {
        const int b = g.bin_base[zone] + ring * g.num_sectors[zone] + sector;
        BinFit& r = fit[b];
        if (!r.fitted) {  // S:191-195
          r.verdict = 0;  // PWPP_VERDICT_SKIPPED
          seg[b].g_dst = -1; seg[b].ng_dst = ng_run; ng_run += r.n;
          continue;
        }
        if (r.verdict == PW_FIT_NO_PLANE) {  // members keep the last plane (S:49)
          for (int k = 0; k < 3; ++k) { r.mean[k] = st.stale_mean[k]; r.normal[k] = st.stale_normal[k]; r.sv[k] = st.stale_sv[k]; }
        } else {
          for (int k = 0; k < 3; ++k) { st.stale_mean[k] = r.mean[k]; st.stale_normal[k] = r.normal[k]; st.stale_sv[k] = r.sv[k]; }
        }
        for (int k = 0; k < 3; ++k) { cen[npatch * 3 + k] = (float) r.mean[k]; nor[npatch * 3 + k] = (float) r.normal[k]; }  // S:211-212
        ++npatch;
        const double ground_uprightness = r.normal[2];  // S:217-223
        const double ground_elevation = r.mean[2];
        double ground_flatness = r.sv[0];
        if (r.sv[1] < ground_flatness) ground_flatness = r.sv[1];
        if (r.sv[2] < ground_flatness) ground_flatness = r.sv[2];
        const double line_variable = r.sv[1] != 0 ? ddiv(r.sv[0], r.sv[1]) : DBL_MAX;
        double heading = 0.0;
        for (int k = 0; k < 3; ++k) heading = dadd(heading, dmul(r.mean[k], r.normal[k]));
        const bool is_upright = ground_uprightness > ap.uprightness_thr;  // S:235-246
        const bool is_near_zone = concentric_idx < ap.num_rings_of_interest;
        const bool is_heading_outside = heading < 0.0;
        bool is_not_elevated = false, is_flat = false;
        if (is_near_zone) {
          is_not_elevated = ground_elevation < st.elevation_thr[concentric_idx];
          is_flat = ground_flatness < st.flatness_thr[concentric_idx];
        }
        if (is_upright && is_not_elevated && is_near_zone) {  // S:253-259
          history_push(h_elev + concentric_idx * hcap, st.n_elev[concentric_idx], hcap, ground_elevation);
          history_push(h_flat + concentric_idx * hcap, st.n_flat[concentric_idx], hcap, ground_flatness);
          if (n_ringflat < 4096) sc.ringflat[n_ringflat++] = ground_flatness;
        }
        bool ground_to_g = false, is_cand = false;
        if (!is_upright) { r.verdict = 1; }                                   // S:262-265
        else if (!is_near_zone) { r.verdict = 2; ground_to_g = true; }        // S:266-269
        else if (!is_heading_outside) { r.verdict = 3; }                      // S:270-273
        else if (is_not_elevated || is_flat) { r.verdict = 4; ground_to_g = true; }  // S:274-277
        else { is_cand = true; r.verdict = 6; sc.cand_bin[ncand] = b; sc.cand_lv[ncand] = line_variable; ++ncand; }  // S:278-282
        if (is_cand) { seg[b].g_dst = -2; }                                   // resolved at the end of the ring
        else if (ground_to_g) { seg[b].g_dst = g_run; g_run += r.n_ground; }
        else { seg[b].g_dst = -3 - ng_run; ng_run += r.n_ground; }            // ground part -> non-ground list (encoded)
        seg[b].ng_dst = ng_run; ng_run += r.n - r.n_ground;                   // S:284
      }
      if (ncand > 0) {  // S:292-304
        double mean_flatness = 0.0, stdev_flatness = 0.0;
        if (ap.enable_TGR) calc_mean_stdev(sc.ringflat, n_ringflat, mean_flatness, stdev_flatness);  // S:407-408
        for (int c = 0; c < ncand; ++c) {
          const int b = sc.cand_bin[c];
          BinFit& r = fit[b];
          bool revert = false;
          if (ap.enable_TGR) {  // temporal_ground_revert S:416-461
            double flat = r.sv[0];
            if (r.sv[1] < flat) flat = r.sv[1];
            if (r.sv[2] < flat) flat = r.sv[2];
            const double mu_flatness = dadd(mean_flatness, dmul(1.5, stdev_flatness));
            double prob_flatness = ddiv(1.0, dadd(1.0, exp(ddiv(dsub(flat, mu_flatness), ddiv(mu_flatness, 10.0)))));
            if (r.n_ground > 1500 && flat < dmul(ap.th_dist, ap.th_dist)) prob_flatness = 1.0;
            double prob_line = 1.0;
            if (sc.cand_lv[c] > 8.0) prob_line = 0.0;
            revert = dmul(prob_line, prob_flatness) > 0.5;
          }
          if (revert) { r.verdict = 5; seg[b].g_dst = g_run; g_run += r.n_ground; }
          else { r.verdict = 6; seg[b].g_dst = -3 - ng_run; ng_run += r.n_ground; }
        }
        ncand = 0;
        n_ringflat = 0;
      }
      concentric_idx++;
    }
  }
  // non-ground destinations sit behind the ground list; decode the "ground part -> non-ground list" marker
  const int nb_all = nb + PW_NUM_PSEUDO;
  for (int b = 0; b < nb_all; ++b) {
    if (seg[b].ng_dst >= 0) seg[b].ng_dst += g_run;
    if (seg[b].g_dst <= -3) seg[b].g_dst = (-3 - seg[b].g_dst) + g_run;
  }
  out_num_ground = g_run;
  out_num_patches = npatch;
  out_num_dropped = n_drop;
}

// update_elevation_thr S:338-357 then update_flatness_thr S:359-375
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void update_thresholds(const AlgoParams& ap, StreamState& st, double* h_elev, double* h_flat, int hcap) {
  for (int i = 0; i < ap.num_rings_of_interest; ++i) {
    if (st.n_elev[i] == 0) continue;  // S:342
    double m = 0.0, sd = 0.0;
    double* a = h_elev + i * hcap;
    calc_mean_stdev(a, st.n_elev[i], m, sd);
    if (i == 0) { st.elevation_thr[0] = dadd(m, dmul(3.0, sd)); st.sensor_height = -m; }  // S:346-349
    else st.elevation_thr[i] = dadd(m, dmul(2.0, sd));                                     // S:350
    const int exceed = st.n_elev[i] - ap.max_elevation_storage;                            // S:354-355
    if (exceed > 0) { for (int q = exceed; q < st.n_elev[i]; ++q) a[q - exceed] = a[q]; st.n_elev[i] -= exceed; }
  }
  for (int i = 0; i < ap.num_rings_of_interest; ++i) {
    if (st.n_flat[i] == 0) break;   // S:363
    if (st.n_flat[i] <= 1) break;   // S:364
    double m = 0.0, sd = 0.0;
    double* a = h_flat + i * hcap;
    calc_mean_stdev(a, st.n_flat[i], m, sd);
    st.flatness_thr[i] = dadd(m, sd);  // S:368
    const int exceed = st.n_flat[i] - ap.max_flatness_storage;  // S:372-373
    if (exceed > 0) { for (int q = exceed; q < st.n_flat[i]; ++q) a[q - exceed] = a[q]; st.n_flat[i] -= exceed; }
  }
}

}  // namespace pwpp
