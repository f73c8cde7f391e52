// TEST INFRASTRUCTURE ONLY — builds into oracle/_ref/ (git-ignored), never into the product.
//
// Compiles the reference's own translation unit
//   /root/reference/cpp/patchworkpp/src/patchworkpp.cpp   (included below, where it lies)
// against oracle/eigen_shim (Eigen is absent from this container, see eigen_shim/Eigen/Dense)
// and exposes patchwork::PatchWorkpp (reference patchworkpp.h:114-163) through a small C API so
// that tests and bench.py can drive it with ctypes.
//
// -DPWREF_STABLE_SORT: the reference sorts each bin with std::sort (patchworkpp.cpp:199), whose
// order among equal z is implementation-defined. With this macro an overload found by ordinary
// lookup before std::sort makes that call a stable sort, i.e. ties keep ascending point index.
// This is the variant the restated oracle is compared against bit for bit; the variant WITHOUT
// the macro is the unmodified reference and is the one that is timed as the CPU baseline.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <numeric>
#include <vector>
#include <time.h>

#include <Eigen/Dense>

#define private public  // read-only access to params_ / update_*_ for state parity checks
#include "patchwork/patchworkpp.h"
#undef private

#ifdef PWREF_STABLE_SORT
namespace patchwork {
inline void sort(std::vector<PointXYZ>::iterator first, std::vector<PointXYZ>::iterator last,
                 bool (*cmp)(PointXYZ, PointXYZ)) {
  std::stable_sort(first, last, cmp);
}
}  // namespace patchwork
#endif

#include PWREF_SOURCE  // "/root/reference/cpp/patchworkpp/src/patchworkpp.cpp"

#include "pwpp.h"

namespace {
struct RefHandle {
  patchwork::PatchWorkpp* obj;
};
patchwork::Params to_params(const pwpp_params* p) {
  patchwork::Params q;
  q.verbose = p->verbose != 0;
  q.enable_RNR = p->enable_RNR != 0;
  q.enable_RVPF = p->enable_RVPF != 0;
  q.enable_TGR = p->enable_TGR != 0;
  q.num_iter = p->num_iter;
  q.num_lpr = p->num_lpr;
  q.num_min_pts = p->num_min_pts;
  q.num_zones = p->num_zones;
  q.num_rings_of_interest = p->num_rings_of_interest;
  q.RNR_ver_angle_thr = p->RNR_ver_angle_thr;
  q.RNR_intensity_thr = p->RNR_intensity_thr;
  q.sensor_height = p->sensor_height;
  q.th_seeds = p->th_seeds;
  q.th_dist = p->th_dist;
  q.th_seeds_v = p->th_seeds_v;
  q.th_dist_v = p->th_dist_v;
  q.max_range = p->max_range;
  q.min_range = p->min_range;
  q.uprightness_thr = p->uprightness_thr;
  q.adaptive_seed_selection_margin = p->adaptive_seed_selection_margin;
  q.intensity_thr = p->intensity_thr;
  q.num_sectors_each_zone.assign(p->num_sectors_each_zone, p->num_sectors_each_zone + 4);
  q.num_rings_each_zone.assign(p->num_rings_each_zone, p->num_rings_each_zone + 4);
  q.max_flatness_storage = p->max_flatness_storage;
  q.max_elevation_storage = p->max_elevation_storage;
  q.elevation_thr.assign(p->elevation_thr, p->elevation_thr + 4);
  q.flatness_thr.assign(p->flatness_thr, p->flatness_thr + 4);
  return q;
}
}  // namespace

extern "C" {

// The reference constructor prints an unconditional banner (patchworkpp.h:149). This library is the only user of std::cout in
// the processes that load it (tests, bench.py), so the stream is switched off once when the library is loaded: insertions into
// a stream whose failbit is set do nothing, and no constructor has to be serialised behind a lock (bench.py creates one
// instance per frame from every host thread: a lock around the ~90 us constructor would cap the reference arm's parallelism).
namespace {
struct SilenceCout { SilenceCout() { std::cout.setstate(std::ios_base::failbit); } } g_silence_cout;
}

void* pwref_create(const pwpp_params* p) {
  RefHandle* h = new RefHandle;
  h->obj = new patchwork::PatchWorkpp(to_params(p));
  return h;
}
void pwref_destroy(void* hv) {
  RefHandle* h = static_cast<RefHandle*>(hv);
  delete h->obj;
  delete h;
}
// pts: row-major n x cols (numpy C order), converted to the column-major MatrixXf the
// reference takes by value (patchworkpp.h:152), like pybind11/eigen.h does for the binding.
void pwref_estimate(void* hv, const float* pts, int64_t n, int cols) {
  RefHandle* h = static_cast<RefHandle*>(hv);
  Eigen::MatrixXf cloud(n, cols);
  for (int c = 0; c < cols; ++c)
    for (int64_t i = 0; i < n; ++i) cloud(i, c) = pts[i * cols + c];
  h->obj->estimateGround(cloud);
}
int64_t pwref_num_ground(void* hv) { return (int64_t) static_cast<RefHandle*>(hv)->obj->cloud_ground_.size(); }
int64_t pwref_num_nonground(void* hv) { return (int64_t) static_cast<RefHandle*>(hv)->obj->cloud_nonground_.size(); }
void pwref_ground_indices(void* hv, int32_t* dst) {
  Eigen::VectorXi v = static_cast<RefHandle*>(hv)->obj->getGroundIndices();
  for (int64_t i = 0; i < v.rows(); ++i) dst[i] = v(i);
}
void pwref_nonground_indices(void* hv, int32_t* dst) {
  Eigen::VectorXi v = static_cast<RefHandle*>(hv)->obj->getNongroundIndices();
  for (int64_t i = 0; i < v.rows(); ++i) dst[i] = v(i);
}
static void copy_x3(const Eigen::MatrixX3f& m, float* dst) {
  for (int64_t i = 0; i < m.rows(); ++i) for (int c = 0; c < 3; ++c) dst[i * 3 + c] = m(i, c);
}
void pwref_ground_xyz(void* hv, float* dst) { copy_x3(static_cast<RefHandle*>(hv)->obj->getGround(), dst); }
void pwref_nonground_xyz(void* hv, float* dst) { copy_x3(static_cast<RefHandle*>(hv)->obj->getNonground(), dst); }
int pwref_num_patches(void* hv) { return (int) static_cast<RefHandle*>(hv)->obj->centers_.size(); }
void pwref_centers(void* hv, float* dst) { copy_x3(static_cast<RefHandle*>(hv)->obj->getCenters(), dst); }
void pwref_normals(void* hv, float* dst) { copy_x3(static_cast<RefHandle*>(hv)->obj->getNormals(), dst); }
double pwref_height(void* hv) { return static_cast<RefHandle*>(hv)->obj->getHeight(); }
double pwref_time_taken(void* hv) { return static_cast<RefHandle*>(hv)->obj->getTimeTaken(); }
void pwref_get_state(void* hv, pwpp_state* st) {
  patchwork::PatchWorkpp* o = static_cast<RefHandle*>(hv)->obj;
  st->sensor_height = o->params_.sensor_height;
  for (int i = 0; i < 4; ++i) {
    st->elevation_thr[i] = o->params_.elevation_thr[i];
    st->flatness_thr[i] = o->params_.flatness_thr[i];
    st->n_elevation[i] = (int32_t) o->update_elevation_[i].size();
    st->n_flatness[i] = (int32_t) o->update_flatness_[i].size();
  }
}
void pwref_history(void* hv, int ring, int which, double* dst) {
  patchwork::PatchWorkpp* o = static_cast<RefHandle*>(hv)->obj;
  const std::vector<double>& v = which ? o->update_flatness_[ring] : o->update_elevation_[ring];
  if (!v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(double));
}

}  // extern "C"
