// pypatchworkpp — Python module with the surface of the reference binding
// (reference python/patchworkpp/pybinding.cpp:9-55): class `Parameters` (same fields) and class
// `patchworkpp` (same nine methods). numpy in, numpy out: estimateGround takes any float-convertible
// 2-D array (n,3|4) in C or Fortran order without an extra copy; getters return float32 (n,3)
// Fortran-ordered arrays / int32 (n,) arrays like pybind11/eigen.h produces for the reference.
// The GIL is released while the GPU works.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "patchwork/patchworkpp.h"

namespace py = pybind11;

namespace {
py::array_t<float> x3(const std::vector<float>& v) {
  const py::ssize_t n = (py::ssize_t) (v.size() / 3);
  py::array_t<float, py::array::f_style> a({n, (py::ssize_t) 3});
  auto r = a.mutable_unchecked<2>();
  for (py::ssize_t i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) r(i, c) = v[3 * i + c];
  return a;
}
py::array_t<int> ivec(const std::vector<int>& v) {
  py::array_t<int> a((py::ssize_t) v.size());
  if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(int));
  return a;
}
}  // namespace

PYBIND11_MODULE(pypatchworkpp, m) {
  m.doc() = "Python Patchwork++ (B200 engine)";
  m.attr("__version__") = "0.0.1";

  py::class_<patchwork::Params>(m, "Parameters")
      .def(py::init<>())
      .def_readwrite("sensor_height", &patchwork::Params::sensor_height)
      .def_readwrite("verbose", &patchwork::Params::verbose)
      .def_readwrite("enable_RNR", &patchwork::Params::enable_RNR)
      .def_readwrite("enable_RVPF", &patchwork::Params::enable_RVPF)
      .def_readwrite("enable_TGR", &patchwork::Params::enable_TGR)
      .def_readwrite("num_iter", &patchwork::Params::num_iter)
      .def_readwrite("num_lpr", &patchwork::Params::num_lpr)
      .def_readwrite("num_min_pts", &patchwork::Params::num_min_pts)
      .def_readwrite("num_zones", &patchwork::Params::num_zones)
      .def_readwrite("num_rings_of_interest", &patchwork::Params::num_rings_of_interest)
      .def_readwrite("RNR_ver_angle_thr", &patchwork::Params::RNR_ver_angle_thr)
      .def_readwrite("RNR_intensity_thr", &patchwork::Params::RNR_intensity_thr)
      .def_readwrite("th_seeds", &patchwork::Params::th_seeds)
      .def_readwrite("th_dist", &patchwork::Params::th_dist)
      .def_readwrite("th_seeds_v", &patchwork::Params::th_seeds_v)
      .def_readwrite("th_dist_v", &patchwork::Params::th_dist_v)
      .def_readwrite("max_range", &patchwork::Params::max_range)
      .def_readwrite("min_range", &patchwork::Params::min_range)
      .def_readwrite("uprightness_thr", &patchwork::Params::uprightness_thr)
      .def_readwrite("adaptive_seed_selection_margin", &patchwork::Params::adaptive_seed_selection_margin)
      .def_readwrite("intensity_thr", &patchwork::Params::intensity_thr)
      .def_readwrite("num_sectors_each_zone", &patchwork::Params::num_sectors_each_zone)
      .def_readwrite("num_rings_each_zone", &patchwork::Params::num_rings_each_zone)
      .def_readwrite("max_flatness_storage", &patchwork::Params::max_flatness_storage)
      .def_readwrite("max_elevation_storage", &patchwork::Params::max_elevation_storage)
      .def_readwrite("elevation_thr", &patchwork::Params::elevation_thr)
      .def_readwrite("flatness_thr", &patchwork::Params::flatness_thr);

  py::class_<patchwork::PatchWorkpp>(m, "patchworkpp")
      .def(py::init<patchwork::Params>())
      .def(py::init<patchwork::Params, int>(), py::arg("params"), py::arg("device"))
      .def("getHeight", &patchwork::PatchWorkpp::getHeight)
      .def("getTimeTaken", &patchwork::PatchWorkpp::getTimeTaken)
      .def("getGround", [](patchwork::PatchWorkpp& s) { return x3(s.getGroundVec()); })
      .def("getNonground", [](patchwork::PatchWorkpp& s) { return x3(s.getNongroundVec()); })
      .def("getCenters", [](patchwork::PatchWorkpp& s) { return x3(s.getCentersVec()); })
      .def("getGroundIndices", [](patchwork::PatchWorkpp& s) { return ivec(s.getGroundIndicesVec()); })
      .def("getNongroundIndices", [](patchwork::PatchWorkpp& s) { return ivec(s.getNongroundIndicesVec()); })
      .def("getNormals", [](patchwork::PatchWorkpp& s) { return x3(s.getNormalsVec()); })
      .def("estimateGround", [](patchwork::PatchWorkpp& s, py::array_t<float, py::array::forcecast> cloud) {
        if (cloud.ndim() != 2) throw std::runtime_error("estimateGround: expected a 2-D array (n, 3|4)");
        const float* data = cloud.data();
        const int64_t n = cloud.shape(0);
        const int cols = (int) cloud.shape(1);
        const int64_t rs = cloud.strides(0) / (py::ssize_t) sizeof(float), cs = cloud.strides(1) / (py::ssize_t) sizeof(float);
        py::gil_scoped_release nogil;
        s.estimateGround(data, n, cols, rs, cs);
      });
}
