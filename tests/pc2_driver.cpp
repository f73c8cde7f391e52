// TEST-ONLY driver for include/patchwork/pointcloud2.hpp (tests/test_examples.py): reads a scan (float32 x,y,z,i records),
// lays it out as several PointCloud2-style messages (different point_step / field offsets), runs each through
// patchwork::estimateGround(pw, view) and prints "layout zero_copy ground nonground payload_bytes".
#include <patchwork/pointcloud2.hpp>

#include <cstdio>
#include <cstring>
#include <vector>

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<float> scan(4 * 200000);
  const size_t n = std::fread(scan.data(), 16, 200000, f);
  std::fclose(f);
  struct Layout { const char* name; uint32_t step; int ox, oy, oz, oi; };
  const Layout layouts[] = {
      {"xyz12", 12, 0, 4, 8, -1},            // what the reference node consumes
      {"xyzi16", 16, 0, 4, 8, 12},           // packed with intensity
      {"pcl_xyzi32", 32, 0, 4, 8, 16},       // PCL PointXYZI: padding before intensity -> gather
      {"velodyne22", 22, 0, 4, 8, 12},       // x,y,z,intensity,ring(u16),time... unaligned step -> gather
      {"ouster48_noint", 48, 16, 20, 24, -1} // fields in the middle of a wide point, no intensity -> strided
  };
  patchwork::Params params;
  params.verbose = false;
  if (argc > 2 && std::strcmp(argv[2], "ros") == 0) {   // reference ros/launch/patchworkpp.launch.py:50-64 + GroundSegmentationServer.cpp:47
    params.sensor_height = 1.88; params.num_iter = 3; params.num_lpr = 20; params.num_min_pts = 0; params.th_seeds = 0.3; params.th_dist = 0.125;
    params.th_seeds_v = 0.25; params.th_dist_v = 0.9; params.max_range = 80.0; params.min_range = 1.0; params.uprightness_thr = 0.101;
    params.enable_RNR = false;
  }
  for (const Layout& L : layouts) {
    std::vector<uint8_t> msg((size_t) n * L.step + 8, 0xAB);
    for (size_t i = 0; i < n; ++i) {
      uint8_t* p = msg.data() + i * L.step;
      std::memcpy(p + L.ox, &scan[4 * i], 4); std::memcpy(p + L.oy, &scan[4 * i + 1], 4); std::memcpy(p + L.oz, &scan[4 * i + 2], 4);
      if (L.oi >= 0) std::memcpy(p + L.oi, &scan[4 * i + 3], 4);
    }
    patchwork::PatchWorkpp pw(params);
    patchwork::PointCloud2View v;
    v.data = msg.data(); v.num_points = (int64_t) n; v.point_step = L.step; v.off_x = L.ox; v.off_y = L.oy; v.off_z = L.oz; v.off_intensity = L.oi;
    const bool zc = patchwork::estimateGround(pw, v);
    const patchwork::PointCloud2Payload g = patchwork::makeCloudPayload(pw, true), ng = patchwork::makeCloudPayload(pw, false);
    long long chk = 0;   // order-independent checksum of the ground index set
    for (int v : pw.getGroundIndicesVec()) chk += (long long) v * (long long) (v % 97 + 1);
    std::printf("%s %d %u %u %zu %lld\n", L.name, zc ? 1 : 0, g.width, ng.width, g.data.size() + ng.data.size(), chk);
  }
  return 0;
}
