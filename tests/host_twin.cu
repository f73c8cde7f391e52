// host_twin.cu — TEST-ONLY CPU twin of the CUDA pipeline.
//
// Runs the product's __host__ __device__ code (csrc/pwpp_math.cuh: binning filter, Jacobi SVD, plane from
// shifted moments, point-plane distance) and the sequential A-GLE / TGR / thresholds stage (tests/gle_sequential.cuh) on the
// CPU, with the warp-parallel glue of the kernels replaced by plain sequential loops that follow the same
// algorithmic restructuring as k_fit (no z-sort, K-smallest LPR selection, R-VPF "alive" test by stored
// planes, one-pass moments about a reference point). tests/test_host_twin.py compares it with the oracle,
// which lets the math and the sequential logic be validated in the GPU-less build container.
// Built by tests/conftest.py with: nvcc -x cu (host pass only; no device is needed to run it).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "pwpp.h"
#include "pwpp_host.hpp"
#include "gle_sequential.cuh"

using namespace pwpp;

namespace {
struct Pt { float x, y, z, w; int idx; };

bool alive(const std::vector<Plane>& rv, double th_dist_v, const Pt& p) {
  bool a = true;
  for (const Plane& pl : rv) a = a && !(fabs(point_plane_distance(pl, p.x, p.y, p.z)) < th_dist_v);
  return a;
}
double select_lpr(const std::vector<Pt>& P, bool zone0, double margin_z, int num_lpr, const std::vector<Plane>& rv, double th_dist_v) {
  std::vector<float> zs;
  for (const Pt& p : P) {
    bool valid = alive(rv, th_dist_v, p);
    if (zone0 && ((double) p.z < margin_z)) valid = false;
    if (valid) zs.push_back(p.z);
  }
  std::sort(zs.begin(), zs.end());
  const int cnt = (int) std::min<size_t>(zs.size(), (size_t) num_lpr);
  double sum = 0.0;
  for (int i = 0; i < cnt; ++i) sum += (double) zs[i];
  return cnt != 0 ? sum / cnt : 0.0;
}
template <int MODE>
Moments accumulate(const std::vector<Pt>& P, const std::vector<Plane>& rv, double th_dist_v, double zthr, const Plane& pl, double th_dist, const double c[3]) {
  Moments m;
  std::memset(&m, 0, sizeof(m));
  for (const Pt& p : P) {
    bool in = alive(rv, th_dist_v, p);
    if (MODE == 0) in = in && ((double) p.z < zthr);
    else in = in && (point_plane_distance(pl, p.x, p.y, p.z) < th_dist);
    if (in) {
      const double dx = (double) p.x - c[0], dy = (double) p.y - c[1], dz = (double) p.z - c[2];
      m.s1[0] += dx; m.s1[1] += dy; m.s1[2] += dz;
      m.s2[0] += dx * dx; m.s2[1] += dx * dy; m.s2[2] += dx * dz; m.s2[3] += dy * dy; m.s2[4] += dy * dz; m.s2[5] += dz * dz;
      m.n += 1;
    }
  }
  return m;
}
}  // namespace

struct Twin {
  pwpp_params prm;
  Geometry g;
  AlgoParams ap;
  bool fast;
  int hcap;
  StreamState st;
  std::vector<double> hist;  // [2][4][hcap]
  // results of the last frame
  std::vector<int> ground, nonground;
  std::vector<uint16_t> bin_ids;
  std::vector<BinFit> fits;
  std::vector<float> centers, normals;
  int npatch = 0;
  long long fast_mismatch = 0;
};

extern "C" {

void* twin_create(const pwpp_params* p) {
  Twin* t = new Twin();
  t->prm = *p;
  build_geometry(*p, t->g, t->ap, t->fast);
  int max_sectors = 0;
  for (int k = 0; k < 4; ++k) max_sectors = std::max(max_sectors, t->g.num_sectors[k]);
  t->hcap = std::max(p->max_elevation_storage, p->max_flatness_storage) + 4 * max_sectors + 64;
  t->hist.assign((size_t) 2 * 4 * t->hcap, 0.0);
  init_state(*p, t->st);
  return t;
}
void twin_destroy(void* h) { delete (Twin*) h; }
int twin_num_bins(void* h) { return ((Twin*) h)->g.nbins; }
int twin_uses_fast_binning(void* h) { return ((Twin*) h)->fast ? 1 : 0; }

void twin_estimate(void* h, const float* pts, int64_t n, int cols) {
  Twin* t = (Twin*) h;
  const Geometry& g = t->g;
  const AlgoParams& ap = t->ap;
  const int nb = g.nbins, nb_all = nb + PW_NUM_PSEUDO;
  // k_bin_hist
  t->bin_ids.assign((size_t) n, 0);
  std::vector<std::vector<Pt>> bins(nb_all);
  const bool rnr_on = ap.enable_RNR && cols >= 4;
  for (int64_t i = 0; i < n; ++i) {
    Pt p{pts[i * cols], pts[i * cols + 1], pts[i * cols + 2], cols >= 4 ? pts[i * cols + 3] : 0.f, (int) i};
    int bin;
    if (rnr_on && rnr_hit(p.x, p.y, p.z, p.w, t->st.sensor_height, ap)) bin = PW_BIN_RNR(nb);
    else if (p.z == FLT_MIN) bin = PW_BIN_DROP(nb);
    else {
      const int be = bin_of_point_exact(p.x, p.y, p.z, g);
      bin = t->fast ? bin_of_point(p.x, p.y, p.z, g) : be;
      if (bin != be) t->fast_mismatch++;
    }
    t->bin_ids[(size_t) i] = (uint16_t) bin;
    bins[bin].push_back(p);  // k_scatter: stable, ascending index
  }
  std::vector<int> bo(nb_all + 1, 0);
  for (int b = 0; b < nb_all; ++b) bo[b + 1] = bo[b] + (int) bins[b].size();
  // k_fit
  t->fits.assign((size_t) nb, BinFit());
  std::vector<std::vector<int>> part(nb_all);
  for (int b = 0; b < nb_all; ++b) {
    const std::vector<Pt>& P = bins[b];
    const int cnt = (int) P.size();
    if (b >= nb || cnt < ap.num_min_pts || cnt == 0) {
      for (const Pt& p : P) part[b].push_back(p.idx);
      if (b < nb) {
        BinFit& r = t->fits[b];
        std::memset(&r, 0, sizeof(r));
        r.n = cnt; r.fitted = (cnt >= ap.num_min_pts) ? 1 : 0;
        if (r.fitted) r.verdict = PW_FIT_NO_PLANE;
      }
      continue;
    }
    const int zone = (b >= g.bin_base[3]) ? 3 : (b >= g.bin_base[2]) ? 2 : (b >= g.bin_base[1]) ? 1 : 0;
    const bool zone0 = zone == 0;
    const double margin_z = ap.adaptive_seed_selection_margin * t->st.sensor_height;
    std::vector<Plane> rv;
    Plane pl;
    std::memset(&pl, 0, sizeof(pl));
    bool have_plane = false;
    double c[3] = {(double) P[0].x, (double) P[0].y, 0.0};
    if (ap.enable_RVPF && zone0) {
      for (int it = 0; it < ap.num_iter; ++it) {
        const double lpr = select_lpr(P, true, margin_z, ap.num_lpr, rv, ap.th_dist_v);
        c[2] = lpr;
        const Moments m = accumulate<0>(P, rv, ap.th_dist_v, lpr + ap.th_seeds_v, pl, 0.0, c);
        if (m.n > 0) { plane_from_moments(m, c, pl); have_plane = true; }
        if (have_plane && pl.normal[2] < ap.uprightness_thr) rv.push_back(pl); else break;
      }
    }
    {
      const double lpr = select_lpr(P, zone0, margin_z, ap.num_lpr, rv, ap.th_dist_v);
      c[2] = lpr;
      const Moments m = accumulate<0>(P, rv, ap.th_dist_v, lpr + ap.th_seeds, pl, 0.0, c);
      if (m.n > 0) { plane_from_moments(m, c, pl); have_plane = true; }
    }
    for (int it = 0; it < ap.num_iter - 1; ++it) {
      if (!have_plane) break;
      const double cc[3] = {pl.mean[0], pl.mean[1], pl.mean[2]};
      const Moments m = accumulate<1>(P, rv, ap.th_dist_v, 0.0, pl, ap.th_dist, cc);
      if (m.n > 0) plane_from_moments(m, cc, pl);
    }
    std::vector<int> gi, ni;
    {
      const double cc[3] = {have_plane ? pl.mean[0] : c[0], have_plane ? pl.mean[1] : c[1], have_plane ? pl.mean[2] : c[2]};
      Moments m;
      std::memset(&m, 0, sizeof(m));
      for (const Pt& p : P) {
        const bool is_g = alive(rv, ap.th_dist_v, p) && have_plane && (point_plane_distance(pl, p.x, p.y, p.z) < ap.th_dist);
        if (is_g) {
          const double dx = (double) p.x - cc[0], dy = (double) p.y - cc[1], dz = (double) p.z - cc[2];
          m.s1[0] += dx; m.s1[1] += dy; m.s1[2] += dz;
          m.s2[0] += dx * dx; m.s2[1] += dx * dy; m.s2[2] += dx * dz; m.s2[3] += dy * dy; m.s2[4] += dy * dz; m.s2[5] += dz * dz;
          m.n += 1;
          gi.push_back(p.idx);
        } else ni.push_back(p.idx);
      }
      if (m.n > 0) plane_from_moments(m, cc, pl);
    }
    part[b] = gi;
    part[b].insert(part[b].end(), ni.begin(), ni.end());  // (the kernel stores this part reversed; k_emit un-reverses)
    BinFit& r = t->fits[b];
    r.n = cnt; r.n_ground = (int) gi.size(); r.fitted = 1; r.verdict = have_plane ? 0 : PW_FIT_NO_PLANE;
    for (int k = 0; k < 3; ++k) { r.mean[k] = pl.mean[k]; r.normal[k] = pl.normal[k]; r.sv[k] = pl.sv[k]; }
    r.d = pl.d;
  }
  // k_gle
  std::vector<BinSeg> seg(nb_all);
  t->centers.assign((size_t) nb * 3, 0.f);
  t->normals.assign((size_t) nb * 3, 0.f);
  static GleScratch sc;
  int ng = 0, np = 0, nd = 0;
  double* h_elev = t->hist.data();
  double* h_flat = t->hist.data() + (size_t) 4 * t->hcap;
  gle_frame(g, ap, t->st, h_elev, h_flat, t->hcap, bo.data(), t->fits.data(), seg.data(), t->centers.data(), t->normals.data(), sc, ng, np, nd);
  update_thresholds(ap, t->st, h_elev, h_flat, t->hcap);
  t->npatch = np;
  // k_emit
  std::vector<int> out((size_t) n, -1);
  for (int b = 0; b < nb_all; ++b) {
    const int ngb = (b < nb) ? t->fits[b].n_ground : 0;
    for (int j = 0; j < (int) part[b].size(); ++j) {
      int dst;
      if (j < ngb) dst = seg[b].g_dst + j;
      else { if (seg[b].ng_dst < 0) continue; dst = seg[b].ng_dst + (j - ngb); }
      out[(size_t) dst] = part[b][j];
    }
  }
  t->ground.assign(out.begin(), out.begin() + ng);
  t->nonground.assign(out.begin() + ng, out.begin() + (n - nd));
}

int64_t twin_num_ground(void* h) { return (int64_t) ((Twin*) h)->ground.size(); }
int64_t twin_num_nonground(void* h) { return (int64_t) ((Twin*) h)->nonground.size(); }
void twin_ground_indices(void* h, int32_t* dst) { Twin* t = (Twin*) h; std::memcpy(dst, t->ground.data(), t->ground.size() * 4); }
void twin_nonground_indices(void* h, int32_t* dst) { Twin* t = (Twin*) h; std::memcpy(dst, t->nonground.data(), t->nonground.size() * 4); }
int twin_num_patches(void* h) { return ((Twin*) h)->npatch; }
void twin_centers(void* h, float* dst) { Twin* t = (Twin*) h; std::memcpy(dst, t->centers.data(), (size_t) t->npatch * 12); }
void twin_normals(void* h, float* dst) { Twin* t = (Twin*) h; std::memcpy(dst, t->normals.data(), (size_t) t->npatch * 12); }
void twin_bin_ids(void* h, uint16_t* dst) { Twin* t = (Twin*) h; std::memcpy(dst, t->bin_ids.data(), t->bin_ids.size() * 2); }
void twin_bin_results(void* h, pwpp_bin_result* dst) { Twin* t = (Twin*) h; std::memcpy(dst, t->fits.data(), t->fits.size() * sizeof(BinFit)); }
long long twin_fast_mismatches(void* h) { return ((Twin*) h)->fast_mismatch; }
void twin_get_state(void* h, pwpp_state* out) {
  Twin* t = (Twin*) h;
  out->sensor_height = t->st.sensor_height;
  for (int i = 0; i < 4; ++i) {
    out->elevation_thr[i] = t->st.elevation_thr[i]; out->flatness_thr[i] = t->st.flatness_thr[i];
    out->n_elevation[i] = t->st.n_elev[i]; out->n_flatness[i] = t->st.n_flat[i];
  }
}
void twin_history(void* h, int ring, int which, double* dst) {
  Twin* t = (Twin*) h;
  const int n = which ? t->st.n_flat[ring] : t->st.n_elev[ring];
  std::memcpy(dst, t->hist.data() + ((size_t) which * 4 + ring) * t->hcap, (size_t) n * 8);
}
// batch access to the two 3x3 solvers for tests/test_host_twin.py: cov = n x 6 (xx xy xz yy yz zz), out = n x 6 (sv, vector)
void twin_sym_eig3(const double* cov, long long n, double* out) {
  for (long long i = 0; i < n; ++i) sym_eig3(cov[i * 6], cov[i * 6 + 1], cov[i * 6 + 2], cov[i * 6 + 3], cov[i * 6 + 4], cov[i * 6 + 5], out + i * 6, out + i * 6 + 3);
}
void twin_jacobi_svd3(const double* cov, long long n, double* out) {
  for (long long i = 0; i < n; ++i) jacobi_svd3(cov[i * 6], cov[i * 6 + 1], cov[i * 6 + 2], cov[i * 6 + 3], cov[i * 6 + 4], cov[i * 6 + 5], out + i * 6, out + i * 6 + 3);
}
}  // extern "C"
