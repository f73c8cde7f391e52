/*
 * pwpp_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the per-frame hot path of url-kaist/patchwork-plusplus,
 *   patchwork::PatchWorkpp::estimateGround()   cpp/patchworkpp/src/patchworkpp.cpp:151-336
 * ("S:" below = that file, "H:" = cpp/patchworkpp/include/patchwork/patchworkpp.h).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (libpwpp_b200.so) never links, imports or falls back to it.
 *
 * PINNING. The reference ships no golden outputs and no tests. This restatement is
 * pinned against the reference's OWN source instead: oracle/_ref/libpwref_stable.so is
 * S compiled unmodified against oracle/eigen_shim (Eigen 3.4.0, the reference's third-party math
 * dependency pinned at cpp/cmake/eigen.cmake:31, is absent from the container), and
 * tests/test_oracle_vs_reference.py requires arith=REF32 below to agree with it BIT FOR BIT —
 * index lists in emission order, centers, normals, thresholds, histories — on the six fixture
 * scans (fresh and sequential), on synthetic scans and on alternative parameter sets.
 * What remains unpinned is Eigen's last-bit rounding order inside its reductions, which is not
 * canonical even for the real library (SIMD width / cache-size dependent):
 * "parity unpinned" in that narrow sense; DESIGN.md §3 states it.
 *
 * TWO ARITHMETIC MODES (same control flow, selected at create time):
 *   PWO_ARITH_REF32  (0): the reference's precisions literally — fp32 mean / covariance with
 *       sequential sums in z-sorted order, fp32 two-sided Jacobi SVD as published for Eigen 3.4.0
 *       JacobiSVD (S:47-75), fp32 point-plane dot product plus double d (S:551-554).
 *   PWO_ARITH_CANON64 (1): the same formulas evaluated in IEEE double: sums of the plane fit in
 *       long double rounded once to double (=> independent of summation order to ~1e-19), the same
 *       Jacobi SVD in double, point-plane distance in double. This is the arithmetic the CUDA path
 *       implements (it cannot reproduce a sequential fp32 sum order in parallel; an order-free
 *       definition makes ground/non-ground index SETS comparable bit-exactly). The two modes are
 *       compared in tests/test_oracle_golden.py (identical index sets on all fixtures; plane
 *       parameters within fp32 noise).
 *
 * Defined behaviour where the reference has none:
 *   - a point whose z is NaN/Inf would reach std::sort with an inconsistent comparator (UB, S:199);
 *     here (and in the CUDA path) such a point is treated like an out-of-range point (non-ground).
 *   - estimate_plane on ONE point divides 0 by 0 (S:57); Eigen 3.4 then reports InvalidInput with
 *     unspecified outputs. Here: U = I (normal (0,0,1)), singular values NaN (Eigen 3.3 behaviour).
 *   - std::sort's order among equal z (S:199) is implementation-defined; here ties keep ascending
 *     point index (stable sort). It only affects fp32 summation order and within-bin output order.
 *   - num_min_pts <= 0 makes the reference fit empty bins with the PREVIOUS bin's plane (S:49) and,
 *     on the very first bin, read empty Eigen vectors (UB). The stale-plane carry is reproduced;
 *     the initial plane is all zeros.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "pwpp.h"

#define PWO_ARITH_REF32 0
#define PWO_ARITH_CANON64 1

typedef struct { float x, y, z; int32_t idx; } pt_t; /* H:20-27 */
typedef struct { pt_t* p; size_t n, cap; } pvec;
typedef struct { double* v; size_t n, cap; } dvec;

static void pv_push(pvec* v, pt_t q) {
  if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 64; v->p = (pt_t*) realloc(v->p, v->cap * sizeof(pt_t)); }
  v->p[v->n++] = q;
}
static void pv_append(pvec* dst, const pvec* src) { for (size_t i = 0; i < src->n; ++i) pv_push(dst, src->p[i]); } /* addCloud S:28-31 */
static void pv_assign(pvec* dst, const pvec* src) { dst->n = 0; pv_append(dst, src); }
static void dv_push(dvec* v, double x) {
  if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 64; v->v = (double*) realloc(v->v, v->cap * sizeof(double)); }
  v->v[v->n++] = x;
}
static void dv_erase_front(dvec* v, size_t k) { memmove(v->v, v->v + k, (v->n - k) * sizeof(double)); v->n -= k; }

typedef struct {
  int concentric_idx, sector_idx; double flatness, line_variable; pvec ground; int bin;
} candidate_t; /* RevertCandidate H:29-40 */

typedef struct pwo {
  pwpp_params prm;            /* params_ (H:169) — sensor_height / *_thr are adaptive */
  int arith;
  double min_ranges[4], ring_sizes[4], sector_sizes[4]; /* H:122-134 */
  int bin_base[5], nbins;
  pvec* czm;                  /* ConcentricZoneModel_ flattened (zone,ring,sector) */
  dvec upd_flat[4], upd_elev[4]; /* H:174-175 */
  /* last estimate_plane() results (members normal_, pc_mean_, singular_values_, d_; H:177-182).
   * In REF32 mode they hold float values exactly. */
  double normal[3], mean[3], sv[3], d;
  pvec ground_pc, rw_ground, rw_nonground, src_wo, src_tmp; /* H:190-191 and locals of S:467-549 */
  pvec cloud_ground, cloud_nonground;  /* H:193 */
  pvec centers, normals;               /* H:195 (idx unused) */
  /* diagnostics */
  uint16_t* bin_ids; int64_t n_pts;
  pwpp_bin_result* bres;
  int32_t* min_fit_n; int cur_bin; /* per bin: smallest non-empty point set handed to estimate_plane */
  pt_t* msort_tmp; size_t msort_cap;
} pwo;

/* ------------------------------------------------------------------------------------------ */
/* 3x3 Jacobi SVD exactly as oracle/eigen_shim restates Eigen 3.4.0's JacobiSVD, for float and  */
/* double (the macro instantiates both; every statement is one rounded operation, compiled with */
/* -ffp-contract=off).                                                                           */
#define DEFINE_JSVD(R, NAME, SQRT, FABS, RMIN, REPS, RNAN)                                         \
  static void NAME##_rot_apply(R* x, int incx, R* y, int incy, int n, R c, R s) {                  \
    if (c == (R) 1 && s == (R) 0) return;                                                          \
    for (int i = 0; i < n; ++i) {                                                                  \
      R xi = x[i * incx], yi = y[i * incy];                                                        \
      R a = c * xi, b = s * yi, e = s * xi, f = c * yi;                                            \
      x[i * incx] = a + b;                                                                         \
      y[i * incy] = -e + f;                                                                        \
    }                                                                                              \
  }                                                                                                \
  /* W, U column-major 3x3: M(i,j) = M[i + 3*j] */                                                 \
  static void NAME(const R* cov, R* sv, R* U) {                                                    \
    const R precision = (R) 2 * REPS, considerAsZero = RMIN;                                       \
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) U[i + 3 * j] = (i == j) ? (R) 1 : (R) 0; \
    R scale = (R) 0; int finite = 1;                                                               \
    for (int i = 0; i < 9; ++i) { R v = FABS(cov[i]); if (!(v == v) || isinf(v)) finite = 0; if (v > scale) scale = v; } \
    if (!finite) { sv[0] = sv[1] = sv[2] = RNAN; return; }                                         \
    if (scale == (R) 0) scale = (R) 1;                                                             \
    R W[9];                                                                                        \
    for (int i = 0; i < 9; ++i) W[i] = cov[i] / scale;                                             \
    R maxDiag = (R) 0;                                                                             \
    for (int i = 0; i < 3; ++i) { R v = FABS(W[i + 3 * i]); if (v > maxDiag) maxDiag = v; }        \
    int finished = 0;                                                                              \
    while (!finished) {                                                                            \
      finished = 1;                                                                                \
      for (int p = 1; p < 3; ++p) for (int q = 0; q < p; ++q) {                                    \
        R thr = precision * maxDiag; if (considerAsZero > thr) thr = considerAsZero;               \
        if (FABS(W[p + 3 * q]) > thr || FABS(W[q + 3 * p]) > thr) {                                \
          finished = 0;                                                                            \
          /* real_2x2_jacobi_svd */                                                                \
          R m00 = W[p + 3 * p], m01 = W[p + 3 * q], m10 = W[q + 3 * p], m11 = W[q + 3 * q];        \
          R r1c, r1s;                                                                              \
          R t = m00 + m11, dd = m10 - m01;                                                         \
          if (FABS(dd) < RMIN) { r1s = (R) 0; r1c = (R) 1; }                                       \
          else { R u = t / dd; R uu = u * u; R tmp = SQRT((R) 1 + uu); r1s = (R) 1 / tmp; r1c = u / tmp; } \
          if (!(r1c == (R) 1 && r1s == (R) 0)) {                                                   \
            R p1 = r1c * m00, p2 = r1s * m10, p3 = r1s * m00, p4 = r1c * m10;                      \
            R p5 = r1c * m01, p6 = r1s * m11, p7 = r1s * m01, p8 = r1c * m11;                      \
            m00 = p1 + p2; m10 = -p3 + p4; m01 = p5 + p6; m11 = -p7 + p8;                          \
          }                                                                                        \
          /* j_right.makeJacobi(m00, m01, m11) */                                                  \
          R jrc, jrs;                                                                              \
          R deno = (R) 2 * FABS(m01);                                                              \
          if (deno < RMIN) { jrc = (R) 1; jrs = (R) 0; }                                           \
          else {                                                                                   \
            R tau = (m00 - m11) / deno; R tt = tau * tau; R w = SQRT(tt + (R) 1); R tn;            \
            if (tau > (R) 0) tn = (R) 1 / (tau + w); else tn = (R) 1 / (tau - w);                  \
            R sign_t = tn > (R) 0 ? (R) 1 : (R) -1;                                                \
            R t2 = tn * tn; R nn = (R) 1 / SQRT(t2 + (R) 1);                                       \
            jrs = -sign_t * (m01 / FABS(m01)) * FABS(tn) * nn; jrc = nn;                           \
          }                                                                                        \
          /* j_left = rot1 * j_right.transpose()  (transpose: (c, -s)) */                          \
          R oc = jrc, os = -jrs;                                                                   \
          R q1 = r1c * oc, q2 = r1s * os, q3 = r1c * os, q4 = r1s * oc;                            \
          R jlc = q1 - q2, jls = q3 + q4;                                                          \
          /* W.applyOnTheLeft(p,q,j_left): rows p,q */                                             \
          NAME##_rot_apply(&W[p], 3, &W[q], 3, 3, jlc, jls);                                       \
          /* U.applyOnTheRight(p,q,j_left.transpose()) -> rotation in the plane with transpose of that */ \
          NAME##_rot_apply(&U[3 * p], 1, &U[3 * q], 1, 3, jlc, jls);                               \
          /* W.applyOnTheRight(p,q,j_right) -> uses j_right.transpose() */                         \
          NAME##_rot_apply(&W[3 * p], 1, &W[3 * q], 1, 3, jrc, -jrs);                              \
          R a1 = FABS(W[p + 3 * p]), a2 = FABS(W[q + 3 * q]);                                      \
          R mx = a1 > a2 ? a1 : a2; if (mx > maxDiag) maxDiag = mx;                                \
        }                                                                                          \
      }                                                                                            \
    }                                                                                              \
    for (int i = 0; i < 3; ++i) {                                                                  \
      R a = W[i + 3 * i]; sv[i] = FABS(a);                                                         \
      if (a < (R) 0) for (int k = 0; k < 3; ++k) U[k + 3 * i] = U[k + 3 * i] * (R) -1;             \
    }                                                                                              \
    for (int i = 0; i < 3; ++i) sv[i] = sv[i] * scale;                                             \
    for (int i = 0; i < 3; ++i) {                                                                  \
      int pos = 0; R mx = sv[i];                                                                   \
      for (int k = 1; k < 3 - i; ++k) if (sv[i + k] > mx) { mx = sv[i + k]; pos = k; }             \
      if (mx == (R) 0) break;                                                                      \
      if (pos) { pos += i; R ts = sv[i]; sv[i] = sv[pos]; sv[pos] = ts;                            \
        for (int k = 0; k < 3; ++k) { R tu = U[k + 3 * i]; U[k + 3 * i] = U[k + 3 * pos]; U[k + 3 * pos] = tu; } } \
    }                                                                                              \
  }

DEFINE_JSVD(float, jsvd3f, sqrtf, fabsf, FLT_MIN, FLT_EPSILON, NAN)
DEFINE_JSVD(double, jsvd3d, sqrt, fabs, DBL_MIN, DBL_EPSILON, NAN)

/* ------------------------------------------------------------------------------------------ */
/* estimate_plane  S:47-75                                                                     */
static void estimate_plane(pwo* o, const pvec* g) {
  if (g->n == 0) return; /* S:49: members keep the previous plane */
  const size_t n = g->n;
  if (o->cur_bin >= 0 && (int32_t) n < o->min_fit_n[o->cur_bin]) o->min_fit_n[o->cur_bin] = (int32_t) n;
  if (o->arith == PWO_ARITH_REF32) {
    float mean[3];
    for (int c = 0; c < 3; ++c) { /* colwise().mean(): sequential fp32 sum / float(n) */
      float s = 0.0f;
      for (size_t i = 0; i < n; ++i) { const float* q = &g->p[i].x; s = s + q[c]; }
      mean[c] = s / (float) n;
    }
    float cov[9];
    const float dn = (float) (double) (n - 1); /* S:57 "/ double(rows-1)", demoted to float */
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) {
      float s = 0.0f;
      for (size_t k = 0; k < n; ++k) {
        const float* q = &g->p[k].x;
        float a = q[i] - mean[i], b = q[j] - mean[j];
        float pr = a * b; s = s + pr;
      }
      cov[i + 3 * j] = s / dn;
    }
    float sv[3], U[9];
    jsvd3f(cov, sv, U);
    float nrm[3] = { U[0 + 6], U[1 + 6], U[2 + 6] }; /* U.col(2) */
    if (nrm[2] < 0) for (int i = 0; i < 3; ++i) nrm[i] *= -1;
    float x0 = nrm[0] * mean[0], x1 = nrm[1] * mean[1], x2 = nrm[2] * mean[2];
    float t = x1 + x2; float dot = x0 + t;
    for (int i = 0; i < 3; ++i) { o->normal[i] = nrm[i]; o->mean[i] = mean[i]; o->sv[i] = sv[i]; }
    o->d = -dot;
  } else {
    long double s1[3] = { 0, 0, 0 };
    for (size_t i = 0; i < n; ++i) { s1[0] += g->p[i].x; s1[1] += g->p[i].y; s1[2] += g->p[i].z; }
    double mean[3];
    for (int c = 0; c < 3; ++c) mean[c] = (double) (s1[c] / (long double) n);
    long double s2[6] = { 0, 0, 0, 0, 0, 0 }; /* xx xy xz yy yz zz, centred on the rounded mean */
    for (size_t k = 0; k < n; ++k) {
      long double a = (long double) g->p[k].x - mean[0], b = (long double) g->p[k].y - mean[1], c = (long double) g->p[k].z - mean[2];
      s2[0] += a * a; s2[1] += a * b; s2[2] += a * c; s2[3] += b * b; s2[4] += b * c; s2[5] += c * c;
    }
    const long double dn = (long double) (n - 1);
    double cxx = (double) (s2[0] / dn), cxy = (double) (s2[1] / dn), cxz = (double) (s2[2] / dn);
    double cyy = (double) (s2[3] / dn), cyz = (double) (s2[4] / dn), czz = (double) (s2[5] / dn);
    double cov[9] = { cxx, cxy, cxz, cxy, cyy, cyz, cxz, cyz, czz };
    double sv[3], U[9];
    jsvd3d(cov, sv, U);
    double nrm[3] = { U[6], U[7], U[8] };
    if (nrm[2] < 0) for (int i = 0; i < 3; ++i) nrm[i] *= -1;
    double x0 = nrm[0] * mean[0], x1 = nrm[1] * mean[1], x2 = nrm[2] * mean[2];
    double t = x1 + x2; double dot = x0 + t;
    for (int i = 0; i < 3; ++i) { o->normal[i] = nrm[i]; o->mean[i] = mean[i]; o->sv[i] = sv[i]; }
    o->d = -dot;
  }
}

/* calc_point_to_plane_d  S:551-554 */
static double point_plane_d(const pwo* o, pt_t p) {
  if (o->arith == PWO_ARITH_REF32) {
    float n0 = (float) o->normal[0], n1 = (float) o->normal[1], n2 = (float) o->normal[2];
    float a = n0 * p.x, b = n1 * p.y, c = n2 * p.z;
    float s = a + b; s = s + c;
    return (double) s + o->d;
  } else {
    double a = o->normal[0] * (double) p.x, b = o->normal[1] * (double) p.y, c = o->normal[2] * (double) p.z;
    double s = a + b; s = s + c;
    return s + o->d;
  }
}

/* extract_initial_seeds  S:77-112 / S:114-149 (the two overloads differ only in the threshold) */
static void extract_initial_seeds(pwo* o, int zone_idx, const pvec* sorted, pvec* seeds, double th_seed) {
  seeds->n = 0;
  double sum = 0; int cnt = 0;
  size_t init_idx = 0;
  if (zone_idx == 0) {
    for (size_t i = 0; i < sorted->n; ++i) {
      if ((double) sorted->p[i].z < o->prm.adaptive_seed_selection_margin * o->prm.sensor_height) ++init_idx; else break;
    }
  }
  for (size_t i = init_idx; i < sorted->n && cnt < o->prm.num_lpr; ++i) { sum += sorted->p[i].z; cnt++; }
  double lpr_height = cnt != 0 ? sum / cnt : 0;
  for (size_t i = 0; i < sorted->n; ++i) if ((double) sorted->p[i].z < lpr_height + th_seed) pv_push(seeds, sorted->p[i]);
}

/* extract_piecewiseground  S:467-549. Returns the number of R-VPF removals (diagnostic). */
static void extract_piecewiseground(pwo* o, int zone_idx, const pvec* src, pvec* dst, pvec* non_ground_dst) {
  o->ground_pc.n = 0; dst->n = 0; non_ground_dst->n = 0;
  pv_assign(&o->src_wo, src);
  if (o->prm.enable_RVPF) {
    for (int i = 0; i < o->prm.num_iter; ++i) {
      extract_initial_seeds(o, zone_idx, &o->src_wo, &o->ground_pc, o->prm.th_seeds_v);
      estimate_plane(o, &o->ground_pc);
      if (zone_idx == 0 && o->normal[2] < o->prm.uprightness_thr) {
        pv_assign(&o->src_tmp, &o->src_wo);
        o->src_wo.n = 0;
        for (size_t k = 0; k < o->src_tmp.n; ++k) {
          double distance = point_plane_d(o, o->src_tmp.p[k]);
          if (fabs(distance) < o->prm.th_dist_v) pv_push(non_ground_dst, o->src_tmp.p[k]);
          else pv_push(&o->src_wo, o->src_tmp.p[k]);
        }
      } else break;
    }
  }
  extract_initial_seeds(o, zone_idx, &o->src_wo, &o->ground_pc, o->prm.th_seeds);
  estimate_plane(o, &o->ground_pc);
  for (int i = 0; i < o->prm.num_iter; ++i) {
    o->ground_pc.n = 0;
    for (size_t k = 0; k < o->src_wo.n; ++k) {
      double distance = point_plane_d(o, o->src_wo.p[k]);
      if (i < o->prm.num_iter - 1) { if (distance < o->prm.th_dist) pv_push(&o->ground_pc, o->src_wo.p[k]); }
      else { if (distance < o->prm.th_dist) pv_push(dst, o->src_wo.p[k]); else pv_push(non_ground_dst, o->src_wo.p[k]); }
    }
    if (i < o->prm.num_iter - 1) estimate_plane(o, &o->ground_pc); else estimate_plane(o, dst);
  }
}

/* calc_mean_stdev  S:557-566 (mean, stdev untouched when size <= 1) */
static void calc_mean_stdev(const double* v, size_t n, double* mean, double* stdev) {
  if (n <= 1) return;
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += v[i];
  *mean = s / (double) n;
  for (size_t i = 0; i < n; ++i) *stdev += (v[i] - *mean) * (v[i] - *mean);
  *stdev /= (double) (n - 1);
  *stdev = sqrt(*stdev);
}

/* update_elevation_thr S:338-357, update_flatness_thr S:359-375 */
static void update_elevation_thr(pwo* o) {
  for (int i = 0; i < o->prm.num_rings_of_interest; ++i) {
    if (o->upd_elev[i].n == 0) continue;
    double m = 0.0, sd = 0.0;
    calc_mean_stdev(o->upd_elev[i].v, o->upd_elev[i].n, &m, &sd);
    if (i == 0) { o->prm.elevation_thr[i] = m + 3 * sd; o->prm.sensor_height = -m; }
    else o->prm.elevation_thr[i] = m + 2 * sd;
    int exceed = (int) o->upd_elev[i].n - o->prm.max_elevation_storage;
    if (exceed > 0) dv_erase_front(&o->upd_elev[i], (size_t) exceed);
  }
}
static void update_flatness_thr(pwo* o) {
  for (int i = 0; i < o->prm.num_rings_of_interest; ++i) {
    if (o->upd_flat[i].n == 0) break;
    if (o->upd_flat[i].n <= 1) break;
    double m = 0.0, sd = 0.0;
    calc_mean_stdev(o->upd_flat[i].v, o->upd_flat[i].n, &m, &sd);
    o->prm.flatness_thr[i] = m + sd;
    int exceed = (int) o->upd_flat[i].n - o->prm.max_flatness_storage;
    if (exceed > 0) dv_erase_front(&o->upd_flat[i], (size_t) exceed);
  }
}

/* temporal_ground_revert S:402-464 */
static void temporal_ground_revert(pwo* o, const dvec* ring_flatness, candidate_t* cands, size_t nc, int concentric_idx) {
  double mean_flatness = 0.0, stdev_flatness = 0.0;
  calc_mean_stdev(ring_flatness->v, ring_flatness->n, &mean_flatness, &stdev_flatness);
  for (size_t c = 0; c < nc; ++c) {
    double mu_flatness = mean_flatness + 1.5 * stdev_flatness;
    double prob_flatness = 1 / (1 + exp((cands[c].flatness - mu_flatness) / (mu_flatness / 10)));
    if (cands[c].ground.n > 1500 && cands[c].flatness < o->prm.th_dist * o->prm.th_dist) prob_flatness = 1.0;
    double prob_line = 1.0;
    if (cands[c].line_variable > 8.0) prob_line = 0.0;
    int revert = prob_line * prob_flatness > 0.5;
    if (concentric_idx < o->prm.num_rings_of_interest) {
      if (revert) { pv_append(&o->cloud_ground, &cands[c].ground); o->bres[cands[c].bin].verdict = PWPP_VERDICT_TGR_REVERTED; }
      else { pv_append(&o->cloud_nonground, &cands[c].ground); o->bres[cands[c].bin].verdict = PWPP_VERDICT_TGR_REJECTED; }
    }
  }
}

/* stable merge sort by z ascending == std::stable_sort with point_z_cmp (S:6, S:199) */
static void msort(pwo* o, pt_t* a, size_t n) {
  if (n < 2) return;
  if (o->msort_cap < n) { o->msort_cap = n * 2; o->msort_tmp = (pt_t*) realloc(o->msort_tmp, o->msort_cap * sizeof(pt_t)); }
  pt_t* src = a; pt_t* dst = o->msort_tmp;
  for (size_t w = 1; w < n; w *= 2) {
    for (size_t lo = 0; lo < n; lo += 2 * w) {
      size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      size_t i = lo, j = mid, k = lo;
      while (i < mid && j < hi) { if (src[j].z < src[i].z) dst[k++] = src[j++]; else dst[k++] = src[i++]; }
      while (i < mid) dst[k++] = src[i++];
      while (j < hi) dst[k++] = src[j++];
    }
    pt_t* t = src; src = dst; dst = t;
  }
  if (src != a) memcpy(a, src, n * sizeof(pt_t));
}

/* ------------------------------------------------------------------------------------------ */
void* pwo_create(const pwpp_params* p, int arith) {
  if (p->num_zones != 4 || p->num_rings_of_interest > 4) return NULL;
  pwo* o = (pwo*) calloc(1, sizeof(pwo));
  o->prm = *p; o->arith = arith;
  /* H:122-134 */
  double z2 = (7 * p->min_range + p->max_range) / 8.0, z3 = (3 * p->min_range + p->max_range) / 4.0, z4 = (p->min_range + p->max_range) / 2.0;
  o->min_ranges[0] = p->min_range; o->min_ranges[1] = z2; o->min_ranges[2] = z3; o->min_ranges[3] = z4;
  o->ring_sizes[0] = (z2 - p->min_range) / p->num_rings_each_zone[0];
  o->ring_sizes[1] = (z3 - z2) / p->num_rings_each_zone[1];
  o->ring_sizes[2] = (z4 - z3) / p->num_rings_each_zone[2];
  o->ring_sizes[3] = (p->max_range - z4) / p->num_rings_each_zone[3];
  for (int k = 0; k < 4; ++k) o->sector_sizes[k] = 2 * M_PI / p->num_sectors_each_zone[k];
  o->bin_base[0] = 0;
  for (int k = 0; k < 4; ++k) o->bin_base[k + 1] = o->bin_base[k] + p->num_rings_each_zone[k] * p->num_sectors_each_zone[k];
  o->nbins = o->bin_base[4];
  o->czm = (pvec*) calloc((size_t) o->nbins, sizeof(pvec));
  o->bres = (pwpp_bin_result*) calloc((size_t) o->nbins, sizeof(pwpp_bin_result));
  o->min_fit_n = (int32_t*) calloc((size_t) o->nbins, sizeof(int32_t));
  o->cur_bin = -1;
  return o;
}
void pwo_destroy(void* h) {
  pwo* o = (pwo*) h;
  for (int b = 0; b < o->nbins; ++b) free(o->czm[b].p);
  free(o->czm); free(o->bres); free(o->min_fit_n); free(o->bin_ids); free(o->msort_tmp);
  for (int i = 0; i < 4; ++i) { free(o->upd_flat[i].v); free(o->upd_elev[i].v); }
  free(o->ground_pc.p); free(o->rw_ground.p); free(o->rw_nonground.p); free(o->src_wo.p); free(o->src_tmp.p);
  free(o->cloud_ground.p); free(o->cloud_nonground.p); free(o->centers.p); free(o->normals.p);
  free(o);
}
int pwo_num_bins(void* h) { return ((pwo*) h)->nbins; }

/* estimateGround S:151-336. pts row-major n x cols (cols 3 or 4). */
void pwo_estimate(void* h, const float* pts, int64_t n, int cols) {
  pwo* o = (pwo*) h;
  const pwpp_params* P = &o->prm;
  o->cloud_ground.n = 0; o->cloud_nonground.n = 0; /* S:153-154 */
  o->bin_ids = (uint16_t*) realloc(o->bin_ids, (size_t) (n > 0 ? n : 1) * sizeof(uint16_t));
  o->n_pts = n;
  float* zcopy = (float*) malloc((size_t) (n > 0 ? n : 1) * sizeof(float)); /* the by-value copy's z column (S:394) */
  for (int64_t i = 0; i < n; ++i) { zcopy[i] = pts[i * cols + 2]; o->bin_ids[i] = 0xFFFF; }

  /* 1. reflected_noise_removal S:377-400 */
  if (P->enable_RNR && cols >= 4) {
    for (int64_t i = 0; i < n; ++i) {
      float x = pts[i * cols], y = pts[i * cols + 1];
      float xx = x * x, yy = y * y; float rr = xx + yy;
      double r = (double) sqrtf(rr);                  /* S:387: float arithmetic, std::sqrt(float) */
      double z = pts[i * cols + 2];
      double ver_angle_in_deg = atan2(z, r) * 180 / M_PI;
      if (ver_angle_in_deg < P->RNR_ver_angle_thr && z < -P->sensor_height - 0.8 && (double) pts[i * cols + 3] < P->RNR_intensity_thr) {
        pt_t q = { x, y, pts[i * cols + 2], (int32_t) i };
        pv_push(&o->cloud_nonground, q);
        zcopy[i] = FLT_MIN; /* tombstone, S:394 */
        o->bin_ids[i] = (uint16_t) o->nbins; /* RNR */
      }
    }
  }
  /* 2. flush_patches S:33-45 + pc2czm S:578-622 */
  for (int b = 0; b < o->nbins; ++b) o->czm[b].n = 0;
  for (int64_t i = 0; i < n; ++i) {
    float x = pts[i * cols], y = pts[i * cols + 1], z = zcopy[i];
    if (z == FLT_MIN) { if (o->bin_ids[i] == 0xFFFF) o->bin_ids[i] = (uint16_t) (o->nbins + 2); continue; } /* S:591 */
    double xd = x, yd = y;
    double r = sqrt(xd * xd + yd * yd);           /* xy2radius S:573-576 */
    if ((r <= P->max_range) && (r > P->min_range) && isfinite(z)) {
      double theta = atan2(yd, xd);               /* xy2theta S:568-571 */
      theta = theta > 0 ? theta : 2 * M_PI + theta;
      int k = (r < o->min_ranges[1]) ? 0 : (r < o->min_ranges[2]) ? 1 : (r < o->min_ranges[3]) ? 2 : 3;
      int ring_idx = (int) ((r - o->min_ranges[k]) / o->ring_sizes[k]);
      if (ring_idx > P->num_rings_each_zone[k] - 1) ring_idx = P->num_rings_each_zone[k] - 1;
      int sector_idx = (int) (theta / o->sector_sizes[k]);
      if (sector_idx > P->num_sectors_each_zone[k] - 1) sector_idx = P->num_sectors_each_zone[k] - 1;
      int b = o->bin_base[k] + ring_idx * P->num_sectors_each_zone[k] + sector_idx;
      pt_t q = { x, y, z, (int32_t) i };
      pv_push(&o->czm[b], q);
      o->bin_ids[i] = (uint16_t) b;
    } else {
      pt_t q = { x, y, z, (int32_t) i };
      pv_push(&o->cloud_nonground, q); /* S:618 */
      o->bin_ids[i] = (uint16_t) (o->nbins + 1);
    }
  }
  free(zcopy);

  int concentric_idx = 0;
  o->centers.n = 0; o->normals.n = 0; /* S:176-177 */
  candidate_t* cands = NULL; size_t ncand = 0, capcand = 0;
  dvec ringwise_flatness = { 0, 0, 0 };
  memset(o->bres, 0, (size_t) o->nbins * sizeof(pwpp_bin_result));
  for (int b = 0; b < o->nbins; ++b) o->min_fit_n[b] = INT32_MAX;

  for (int zone_idx = 0; zone_idx < P->num_zones; ++zone_idx) {
    for (int ring_idx = 0; ring_idx < P->num_rings_each_zone[zone_idx]; ++ring_idx) {
      for (int sector_idx = 0; sector_idx < P->num_sectors_each_zone[zone_idx]; ++sector_idx) {
        const int b = o->bin_base[zone_idx] + ring_idx * P->num_sectors_each_zone[zone_idx] + sector_idx;
        pvec* bin = &o->czm[b];
        pwpp_bin_result* br = &o->bres[b];
        br->n = (int32_t) bin->n;
        if (bin->n < (size_t) P->num_min_pts) { /* S:191-195 (size_t < int: negative num_min_pts converts to huge) */
          pv_append(&o->cloud_nonground, bin);
          br->verdict = PWPP_VERDICT_SKIPPED;
          continue;
        }
        msort(o, bin->p, bin->n); /* S:199 */
        o->cur_bin = b;
        extract_piecewiseground(o, zone_idx, bin, &o->rw_ground, &o->rw_nonground); /* S:206 */
        o->cur_bin = -1;
        pt_t cq = { (float) o->mean[0], (float) o->mean[1], (float) o->mean[2], -1 };
        pt_t nq = { (float) o->normal[0], (float) o->normal[1], (float) o->normal[2], -1 };
        pv_push(&o->centers, cq); pv_push(&o->normals, nq); /* S:211-212 */
        br->fitted = 1; br->n_ground = (int32_t) o->rw_ground.n; br->d = o->d;
        for (int i = 0; i < 3; ++i) { br->mean[i] = o->mean[i]; br->normal[i] = o->normal[i]; br->sv[i] = o->sv[i]; }

        /* S:217-223 */
        const double ground_uprightness = o->normal[2];
        const double ground_elevation = o->mean[2];
        double ground_flatness, line_variable, heading = 0.0;
        if (o->arith == PWO_ARITH_REF32) {
          float m = (float) o->sv[0]; for (int i = 1; i < 3; ++i) if ((float) o->sv[i] < m) m = (float) o->sv[i];
          ground_flatness = m;
          float s0 = (float) o->sv[0], s1 = (float) o->sv[1];
          line_variable = s1 != 0 ? (double) (s0 / s1) : DBL_MAX;
          for (int i = 0; i < 3; ++i) { float pr = (float) o->mean[i] * (float) o->normal[i]; heading += pr; }
        } else {
          double m = o->sv[0]; for (int i = 1; i < 3; ++i) if (o->sv[i] < m) m = o->sv[i];
          ground_flatness = m;
          line_variable = o->sv[1] != 0 ? o->sv[0] / o->sv[1] : DBL_MAX;
          for (int i = 0; i < 3; ++i) { double pr = o->mean[i] * o->normal[i]; heading += pr; }
        }
        int is_upright = ground_uprightness > P->uprightness_thr;
        int is_near_zone = concentric_idx < P->num_rings_of_interest;
        int is_heading_outside = heading < 0.0;
        int is_not_elevated = 0, is_flat = 0;
        if (concentric_idx < P->num_rings_of_interest) {
          is_not_elevated = ground_elevation < P->elevation_thr[concentric_idx];
          is_flat = ground_flatness < P->flatness_thr[concentric_idx];
        }
        if (is_upright && is_not_elevated && is_near_zone) { /* S:253-259 */
          dv_push(&o->upd_elev[concentric_idx], ground_elevation);
          dv_push(&o->upd_flat[concentric_idx], ground_flatness);
          dv_push(&ringwise_flatness, ground_flatness);
        }
        if (!is_upright) { pv_append(&o->cloud_nonground, &o->rw_ground); br->verdict = PWPP_VERDICT_NOT_UPRIGHT; }
        else if (!is_near_zone) { pv_append(&o->cloud_ground, &o->rw_ground); br->verdict = PWPP_VERDICT_FAR_GROUND; }
        else if (!is_heading_outside) { pv_append(&o->cloud_nonground, &o->rw_ground); br->verdict = PWPP_VERDICT_HEADING; }
        else if (is_not_elevated || is_flat) { pv_append(&o->cloud_ground, &o->rw_ground); br->verdict = PWPP_VERDICT_NEAR_GROUND; }
        else {
          if (ncand == capcand) { capcand = capcand ? capcand * 2 : 8; cands = (candidate_t*) realloc(cands, capcand * sizeof(candidate_t)); }
          candidate_t* c = &cands[ncand++];
          memset(c, 0, sizeof(*c));
          c->concentric_idx = concentric_idx; c->sector_idx = sector_idx; c->flatness = ground_flatness; c->line_variable = line_variable; c->bin = b;
          pv_assign(&c->ground, &o->rw_ground);
        }
        pv_append(&o->cloud_nonground, &o->rw_nonground); /* S:284 */
      }
      if (ncand) { /* S:292-304 */
        if (P->enable_TGR) temporal_ground_revert(o, &ringwise_flatness, cands, ncand, concentric_idx);
        else for (size_t c = 0; c < ncand; ++c) { pv_append(&o->cloud_nonground, &cands[c].ground); o->bres[cands[c].bin].verdict = PWPP_VERDICT_TGR_REJECTED; }
        for (size_t c = 0; c < ncand; ++c) free(cands[c].ground.p);
        ncand = 0;
        ringwise_flatness.n = 0;
      }
      concentric_idx++;
    }
  }
  free(cands); free(ringwise_flatness.v);
  update_elevation_thr(o); /* S:314 */
  update_flatness_thr(o);  /* S:315 */
}

/* ---- getters (H:154-163, S:8-26) ---- */
int64_t pwo_num_ground(void* h) { return (int64_t) ((pwo*) h)->cloud_ground.n; }
int64_t pwo_num_nonground(void* h) { return (int64_t) ((pwo*) h)->cloud_nonground.n; }
static void copy_idx(const pvec* v, int32_t* dst) { for (size_t i = 0; i < v->n; ++i) dst[i] = v->p[i].idx; }
static void copy_xyz(const pvec* v, float* dst) { for (size_t i = 0; i < v->n; ++i) { dst[3 * i] = v->p[i].x; dst[3 * i + 1] = v->p[i].y; dst[3 * i + 2] = v->p[i].z; } }
void pwo_ground_indices(void* h, int32_t* dst) { copy_idx(&((pwo*) h)->cloud_ground, dst); }
void pwo_nonground_indices(void* h, int32_t* dst) { copy_idx(&((pwo*) h)->cloud_nonground, dst); }
void pwo_ground_xyz(void* h, float* dst) { copy_xyz(&((pwo*) h)->cloud_ground, dst); }
void pwo_nonground_xyz(void* h, float* dst) { copy_xyz(&((pwo*) h)->cloud_nonground, dst); }
int pwo_num_patches(void* h) { return (int) ((pwo*) h)->centers.n; }
void pwo_centers(void* h, float* dst) { copy_xyz(&((pwo*) h)->centers, dst); }
void pwo_normals(void* h, float* dst) { copy_xyz(&((pwo*) h)->normals, dst); }
double pwo_height(void* h) { return ((pwo*) h)->prm.sensor_height; }
void pwo_get_state(void* h, pwpp_state* st) {
  pwo* o = (pwo*) h;
  st->sensor_height = o->prm.sensor_height;
  for (int i = 0; i < 4; ++i) {
    st->elevation_thr[i] = o->prm.elevation_thr[i]; st->flatness_thr[i] = o->prm.flatness_thr[i];
    st->n_elevation[i] = (int32_t) o->upd_elev[i].n; st->n_flatness[i] = (int32_t) o->upd_flat[i].n;
  }
}
void pwo_history(void* h, int ring, int which, double* dst) {
  pwo* o = (pwo*) h;
  const dvec* v = which ? &o->upd_flat[ring] : &o->upd_elev[ring];
  if (v->n) memcpy(dst, v->v, v->n * sizeof(double));
}
void pwo_bin_ids(void* h, uint16_t* dst) { pwo* o = (pwo*) h; memcpy(dst, o->bin_ids, (size_t) o->n_pts * sizeof(uint16_t)); }
/* Diagnostic: per bin, the smallest non-empty point set that estimate_plane (S:47-75) was given. Fewer than 3
 * points make the covariance rank deficient: the "normal" is then a null-space vector chosen by rounding noise
 * (in the reference's fp32 arithmetic as well), so such patches are excluded from cross-arithmetic comparisons. */
void pwo_bin_min_fit_n(void* h, int32_t* dst) { pwo* o = (pwo*) h; memcpy(dst, o->min_fit_n, (size_t) o->nbins * sizeof(int32_t)); }
void pwo_bin_results(void* h, pwpp_bin_result* dst) { pwo* o = (pwo*) h; memcpy(dst, o->bres, (size_t) o->nbins * sizeof(pwpp_bin_result)); }
