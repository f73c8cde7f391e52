// pwpp_latency — single-frame latency of the drop-in class (BASELINE config 2): the call sequence of the reference's demos
// (cpp/patchworkpp/examples/demo_visualize.cpp:75-93: estimateGround, then the index getters) on ONE frame per call,
// timed with the host clock around the whole sequence, so host->device and device->host copies are inside.
//
//   pwpp_latency <scan.bin> [iterations=300] [warmup=30]
//
// Prints one JSON line: median / p10 / p90 / min in microseconds for a pageable caller buffer (what an Eigen matrix is)
// and for a page-locked one (pwpp_host_alloc), plus getTimeTaken() of the last call. bench.py runs it and reports the
// numbers as "latency_us".
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "patchwork/patchworkpp.h"

namespace {
struct Stats { double med, p10, p90, mn; };
Stats stats(std::vector<double> v) {
  std::sort(v.begin(), v.end());
  auto at = [&](double q) { return v[std::min(v.size() - 1, (size_t) (q * v.size()))]; };
  return {at(0.5), at(0.1), at(0.9), v.front()};
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: pwpp_latency <scan.bin> [iterations] [warmup]\n"); return 2; }
  const int iters = argc > 2 ? std::atoi(argv[2]) : 300, warm = argc > 3 ? std::atoi(argv[3]) : 30;
  try {
    std::FILE* fp = std::fopen(argv[1], "rb");
    if (!fp) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    std::fseek(fp, 0, SEEK_END);
    const long bytes = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    const int64_t n = bytes / 16;
    std::vector<float> pageable((size_t) n * 4);
    if (std::fread(pageable.data(), 16, (size_t) n, fp) != (size_t) n) { std::fclose(fp); return 2; }
    std::fclose(fp);
    float* pinned = static_cast<float*>(pwpp_host_alloc((size_t) n * 16));
    if (!pinned) { std::fprintf(stderr, "pwpp_host_alloc failed\n"); return 1; }
    std::memcpy(pinned, pageable.data(), (size_t) n * 16);

    patchwork::Params params;
    params.verbose = false;
    std::fflush(stdout);
    std::FILE* real_out = stdout;
    (void) real_out;
    patchwork::PatchWorkpp pw(params);
    size_t ng = 0, nn = 0;
    auto run = [&](const float* data, std::vector<double>& out) {
      // every iteration is the reference's per-frame pattern on a FRESH temporal state would need a new object; the demos
      // keep one object per sequence, so does this loop (the adaptive state converges after the first few calls)
      for (int i = 0; i < warm + iters; ++i) {
        const auto t0 = std::chrono::steady_clock::now();
        pw.estimateGround(data, n, 4, 4, 1);
        const std::vector<int> g = pw.getGroundIndicesVec();
        const std::vector<int> q = pw.getNongroundIndicesVec();
        const auto t1 = std::chrono::steady_clock::now();
        ng = g.size(); nn = q.size();
        if (i >= warm) out.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
      }
    };
    std::vector<double> a, b;
    run(pageable.data(), a);
    run(pinned, b);
    const Stats sa = stats(a), sb = stats(b);
    std::printf("{\"points\": %lld, \"ground\": %zu, \"nonground\": %zu, \"iterations\": %d, "
                "\"pageable\": {\"median_us\": %.1f, \"p10_us\": %.1f, \"p90_us\": %.1f, \"min_us\": %.1f}, "
                "\"pinned\": {\"median_us\": %.1f, \"p10_us\": %.1f, \"p90_us\": %.1f, \"min_us\": %.1f}, \"time_taken_us\": %.1f}\n",
                (long long) n, ng, nn, iters, sa.med, sa.p10, sa.p90, sa.mn, sb.med, sb.p10, sb.p90, sb.mn, pw.getTimeTaken());
    pwpp_host_free(pinned);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "pwpp_latency: %s\n", e.what());
    return 1;
  }
}
