// pwpp_capi.cu — implementation of the C-ABI declared in include/pwpp.h (libpwpp_b200.so).
// Host side only orchestrates: every stage of estimateGround() runs in the kernels of pwpp_kernels.cuh.
// There is no CPU fallback; without a CUDA device pwpp_create() fails with PWPP_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <sched.h>

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pwpp.h"
#include "pwpp_kernels.cuh"
#include "pwpp_tuning.h"
#include "pwpp_host.hpp"

using namespace pwpp;

static_assert(sizeof(BinFit) == sizeof(pwpp_bin_result), "BinFit must mirror pwpp_bin_result");

namespace {

thread_local std::string g_last_error;

// integer experiment switch from the environment, clamped to [lo, hi]
int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = std::getenv(name);
  if (!e || !*e) return dflt;
  const long v = std::strtol(e, nullptr, 10);
  return (int) (v < lo ? lo : (v > hi ? hi : v));
}

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CU_TRY(expr)                                                                                         \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess) {                                                                                 \
      return fail(PWPP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    }                                                                                                        \
  } while (0)

unsigned long long g_alloc_gen = 0;   // bumped whenever a device buffer moves: captured CUDA graphs hold the old pointers

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    ++g_alloc_gen;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace

// N x cols column-major (an Eigen::MatrixXf, reference patchworkpp.h:152) -> packed float4 {x, y, z, intensity | 0}
__global__ void k_repack_colmajor(const float* __restrict__ src, long long n, int cols, float4* __restrict__ dst) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = make_float4(src[i], src[n + i], src[2 * n + i], cols == 4 ? src[3 * n + i] : 0.f);
}

// packed N x 3 rows on the device (the ROS node's conversion, reference ros/src/Utils.hpp:158-172) -> float4 {x, y, z, 0}
__global__ void k_pad_xyz(const float* __restrict__ src, long long n, float4* __restrict__ dst) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}

typedef void (*FitKernel)(const float4*, FrameTable, const StreamState*, Geometry, AlgoParams, int, const int*, WorkQueues, int*, BinFit*);
struct FitLaunch {
  FitKernel fn = nullptr;
  int grid = 0, threads = 0;
  size_t smem = 0;
};

constexpr int NUM_SIDE = 6;

struct pwpp_ctx {
  pwpp_params prm;
  Geometry g;
  AlgoParams ap;
  int device = 0;
  int num_streams = 0;
  int nbp = 0;      // padded number of bins incl. pseudo-bins
  int hcap = 0;     // history row capacity (doubles)
  bool fast_bin = true;
  // kernel-variant switches, read from the environment when the context is created (see pwpp_create)
  int sw_front = 1, sw_patch = 0, small_call_frames = 0;
  bool sw_serial_fit = false, front_dense_ok = false;
  cudaStream_t stream = nullptr, stream_h2d = nullptr, stream_d2h = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_begin = nullptr, ev_end = nullptr;
  bool call_times_valid = false;
  cudaEvent_t stage_ev[PWPP_NUM_STAGES + 1] = {};
  cudaStream_t side[NUM_SIDE] = {};       // one side stream per patch-size class: its fit kernel, then its sort (reference order)
  cudaEvent_t ev_fork = nullptr, ev_fit[NUM_SIDE] = {}, ev_join[NUM_SIDE] = {};
  bool profiling = false, stage_valid = false;
  long long launches = 0;

  // persistent per-stream state
  DevBuf<StreamState> d_states;
  DevBuf<StreamState> d_states_init;  // constructor state of every stream (reset source)
  DevBuf<double> d_hist;

  // per-call work buffers
  DevBuf<float4> d_in;            // host path only
  DevBuf<float> d_cm;             // host path: column-major frames as uploaded, repacked on the device
  DevBuf<long long> d_pt_off;     // [F+1]
  DevBuf<int> d_chunk_off;        // [F+1]
  DevBuf<unsigned short> d_bin_ids;
  DevBuf<unsigned short> d_chist;
  DevBuf<unsigned int> d_cbase;
  DevBuf<int> d_bin_off;          // [F][nbp+1]
  DevBuf<float4> d_sorted;
  DevBuf<int> d_part;
  DevBuf<BinFit> d_fits;          // [F][nbins]
  DevBuf<BinSeg> d_segs;          // [F][nbins+3]
  DevBuf<int4> d_wq_items[NUM_CLASSES];  // fit work queues
  DevBuf<int> d_wq_ctr;           // [2*NUM_CLASSES + ORD_NUM_HEADS]: counts, heads, the heads of the five k_order launches
  DevBuf<unsigned char> d_labels; // reference-order output only: what became of every point of a fitted patch
  int order_mode = 0;             // PWPP_ORDER_*
  FitLaunch fit[NUM_CLASSES];   // persistent fit kernel of every patch-size class (variant chosen in pwpp_create)
  FitLaunch fit_small[NUM_CLASSES];   // the same for calls of at most small_call_frames frames (one CTA per patch above 512 points)
  int max_sectors = 0;
  FitLaunch order_k[ORD_NUM_HEADS];   // k_order_cta<512, X>, <512, L3>, <256, L2>, <128, L1>, k_order_warp (same argument list as the fit kernels' type is not needed: launched by name)
  DevBuf<int> d_out_idx;
  DevBuf<int> d_counts;           // [3][F]: num_ground, num_patches, num_dropped
  DevBuf<float> d_centers, d_normals;  // [F][nbins][3]
  DevBuf<float> d_xyz;            // gather scratch

  PinBuf<float4> h_in;
  PinBuf<long long> h_pt_off_buf[2];      // double-buffered: a call never waits for the previous call's upload
  PinBuf<int> h_chunk_off_buf[2];
  cudaEvent_t tab_ev[2] = {nullptr, nullptr};
  int tab_cur = 0;
  std::vector<int> chunk_off;             // host copy of the current call's chunk table
  PinBuf<int> h_out_idx;
  PinBuf<int> h_counts;
  PinBuf<float> h_centers, h_normals;

  // description of the last call
  int last_nframes = 0;
  long long last_total = 0;
  std::vector<long long> pt_off;  // host copy
  const float4* last_pts = nullptr;  // device pointer of the input of the last call
  bool counts_fetched = false, idx_fetched = false, patches_fetched = false;
  double last_time_us = 0.0;
  cudaStream_t last_stream = nullptr;

  // small calls (the reference's one-frame-per-call pattern): the launch sequence replayed as a CUDA graph
  struct GraphKey { int nf, has_intensity, chunks; unsigned long long gen; const void* pts; };
  cudaGraphExec_t gexec[2] = {nullptr, nullptr};
  GraphKey gkey[2] = {};
  long long glaunches[2] = {0, 0};
  int sw_graph = 1;
};

namespace {

int validate_params(const pwpp_params* p) {
  if (!p) return fail(PWPP_ERR_INVALID_ARG, "params is NULL");
  if (p->num_zones != PWPP_NUM_ZONES) return fail(PWPP_ERR_UNSUPPORTED, "num_zones must be 4 (the reference hard-wires four zones, patchworkpp.h:127-134)");
  if (p->num_rings_of_interest < 0 || p->num_rings_of_interest > PWPP_MAX_RINGS_OF_INTEREST)
    return fail(PWPP_ERR_UNSUPPORTED, "num_rings_of_interest must be in [0,4] (patchworkpp.h:174-175 holds 4 histories)");
  if (p->num_min_pts < 0) return fail(PWPP_ERR_UNSUPPORTED, "num_min_pts must be >= 0");
  if (p->num_iter < 1 || p->num_iter > MAX_RVPF) return fail(PWPP_ERR_UNSUPPORTED, "num_iter must be in [1,8]");
  if (p->num_lpr < 1 || p->num_lpr > MAX_LPR) return fail(PWPP_ERR_UNSUPPORTED, "num_lpr must be in [1,64]");
  if (!(p->th_seeds > 0) || !(p->th_seeds_v > 0)) return fail(PWPP_ERR_UNSUPPORTED, "th_seeds and th_seeds_v must be > 0");
  if (!(p->max_range > p->min_range) || !(p->min_range >= 0)) return fail(PWPP_ERR_INVALID_ARG, "need 0 <= min_range < max_range");
  if (p->max_flatness_storage < 1 || p->max_elevation_storage < 1) return fail(PWPP_ERR_INVALID_ARG, "max_*_storage must be >= 1");
  long long nb = 0;
  for (int k = 0; k < 4; ++k) {
    if (p->num_rings_each_zone[k] < 1 || p->num_sectors_each_zone[k] < 1 || p->num_sectors_each_zone[k] > 1024)
      return fail(PWPP_ERR_UNSUPPORTED, "rings per zone must be >= 1 and sectors per zone in [1,1024]");
    nb += (long long) p->num_rings_each_zone[k] * p->num_sectors_each_zone[k];
  }
  if (nb + PW_NUM_PSEUDO > 4096) return fail(PWPP_ERR_UNSUPPORTED, "more than 4093 bins are not supported");
  return PWPP_OK;
}

int bind_device(pwpp_ctx* ctx) {
  CU_TRY(cudaSetDevice(ctx->device));
  return PWPP_OK;
}

// Sizes the work buffers for a call over nframes frames (ctx->pt_off filled) and uploads the frame tables.
int prepare_call(pwpp_ctx* ctx, int nframes, cudaStream_t s) {
  const long long total = ctx->pt_off[nframes];
  int total_chunks = 0;
  ctx->tab_cur ^= 1;
  const int tb = ctx->tab_cur;
  PinBuf<long long>& h_pt_off = ctx->h_pt_off_buf[tb];
  PinBuf<int>& h_chunk_off = ctx->h_chunk_off_buf[tb];
  CU_TRY(cudaEventSynchronize(ctx->tab_ev[tb]));   // upload issued two calls ago: long finished
  CU_TRY(h_pt_off.reserve(nframes + 1));
  CU_TRY(h_chunk_off.reserve(nframes + 1));
  ctx->chunk_off.assign(nframes + 1, 0);
  for (int f = 0; f < nframes; ++f) {
    const long long n = ctx->pt_off[f + 1] - ctx->pt_off[f];
    if (n < 0 || n > 0x7fffffffLL - CHUNK_PTS) return fail(PWPP_ERR_INVALID_ARG, "frame size out of range");
    h_pt_off.p[f] = ctx->pt_off[f];
    h_chunk_off.p[f] = total_chunks;
    ctx->chunk_off[f] = total_chunks;
    total_chunks += (int) ((n + CHUNK_PTS - 1) / CHUNK_PTS);
  }
  h_pt_off.p[nframes] = total;
  h_chunk_off.p[nframes] = total_chunks;
  ctx->chunk_off[nframes] = total_chunks;
  const int nb = ctx->g.nbins, nbp = ctx->nbp, nb_all = nb + PW_NUM_PSEUDO;
  CU_TRY(ctx->d_pt_off.reserve(nframes + 1));
  CU_TRY(ctx->d_chunk_off.reserve(nframes + 1));
  CU_TRY(ctx->d_bin_ids.reserve((size_t) total));
  CU_TRY(ctx->d_chist.reserve((size_t) total_chunks * nbp));
  CU_TRY(ctx->d_cbase.reserve((size_t) total_chunks * nbp));
  CU_TRY(ctx->d_bin_off.reserve((size_t) nframes * (nbp + 1)));
  CU_TRY(ctx->d_sorted.reserve((size_t) total));
  CU_TRY(ctx->d_part.reserve((size_t) total));
  if (ctx->order_mode) CU_TRY(ctx->d_labels.reserve((size_t) total));
  CU_TRY(ctx->d_fits.reserve((size_t) nframes * nb));
  CU_TRY(ctx->d_segs.reserve((size_t) nframes * nb_all));
  for (int c = 0; c < NUM_CLASSES; ++c) CU_TRY(ctx->d_wq_items[c].reserve((size_t) nframes * nb));
  CU_TRY(ctx->d_out_idx.reserve((size_t) total));
  CU_TRY(ctx->d_counts.reserve((size_t) 3 * ctx->num_streams));
  CU_TRY(ctx->d_centers.reserve((size_t) nframes * nb * 3));
  CU_TRY(ctx->d_normals.reserve((size_t) nframes * nb * 3));
  CU_TRY(cudaMemcpyAsync(ctx->d_pt_off.p, h_pt_off.p, (nframes + 1) * sizeof(long long), cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(ctx->d_chunk_off.p, h_chunk_off.p, (nframes + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
  CU_TRY(cudaEventRecord(ctx->tab_ev[tb], s));
  ctx->last_nframes = nframes;
  ctx->last_total = total;
  ctx->counts_fetched = ctx->idx_fetched = ctx->patches_fetched = false;
  ctx->stage_valid = false;
  return PWPP_OK;
}

// Launches the whole path for frames [f0, f0 + nf) of the prepared call on stream s. A frame range is the same
// launch sequence over per-frame arrays offset by f0 (frame tables hold absolute point / chunk positions), which is
// what lets pwpp_estimate_host pipeline chunks of frames against their H2D / D2H copies.
int launch_range_impl(pwpp_ctx* ctx, int f0, int nf, const float4* d_pts, int has_intensity, cudaStream_t s, bool prof, int chunks_override) {
  int max_chunks = 0;
  for (int f = f0; f < f0 + nf; ++f) max_chunks = std::max(max_chunks, ctx->chunk_off[f + 1] - ctx->chunk_off[f]);
  if (chunks_override > 0) max_chunks = chunks_override;   // graph capture: grids sized for a range of frame sizes (surplus CTAs exit at once)
  const int nb = ctx->g.nbins, nbp = ctx->nbp, nb_all = nb + PW_NUM_PSEUDO;
  const int nframes = nf;
  FrameTable ft{ctx->d_pt_off.p + f0, ctx->d_chunk_off.p + f0};
  StreamState* states = ctx->d_states.p + f0;
  double* hist = ctx->d_hist.p + (size_t) f0 * 2 * 4 * ctx->hcap;
  int* bin_off = ctx->d_bin_off.p + (size_t) f0 * (nbp + 1);
  BinFit* fits = ctx->d_fits.p + (size_t) f0 * nb;
  BinSeg* segs = ctx->d_segs.p + (size_t) f0 * nb_all;
  float* centers = ctx->d_centers.p + (size_t) f0 * nb * 3;
  float* normals = ctx->d_normals.p + (size_t) f0 * nb * 3;
  CU_TRY(cudaMemsetAsync(ctx->d_wq_ctr.p, 0, (2 * NUM_CLASSES + ORD_NUM_HEADS) * sizeof(int), s));
  int stage = 0;
#define STAGE_MARK() do { if (prof) CU_TRY(cudaEventRecord(ctx->stage_ev[stage], s)); ++stage; } while (0)
  STAGE_MARK();
  WorkQueues wq;
  for (int c = 0; c < NUM_CLASSES; ++c) wq.items[c] = ctx->d_wq_items[c].p;
  wq.count = ctx->d_wq_ctr.p;
  wq.head = ctx->d_wq_ctr.p + NUM_CLASSES;
  wq.labels = ctx->order_mode ? ctx->d_labels.p : nullptr;
  // A call of a few frames (the reference's pattern is one frame per call) cannot fill the GPU with warp-sized work items: there the
  // latency of the longest patch counts, so patches above 512 points go to the CTA-per-patch kernels, and one frame's cloud is
  // binned by ~30 independent CTAs (the three stand-alone kernels) rather than by the 8 CTAs of one cluster. r02, one KITTI frame:
  // longest fit kernel 80 -> 39 us, front end 61 -> 41 us (profiles/r02/ab2_front_bids_latency_solve.log).
  const bool small_call = nframes <= ctx->small_call_frames;
  const FitLaunch* fitk = small_call ? ctx->fit_small : ctx->fit;
  if (ctx->sw_front && !small_call) {
    // one thread-block cluster per frame: binning, scan and stable scatter in one kernel (pwpp_front.cuh)
    if (nframes > 0) {
      // 64 warps per frame for KITTI-sized frames, 128 for dense ones (pwpp_front.cuh)
      const long long mean_front = (ctx->pt_off[f0 + nf] - ctx->pt_off[f0]) / std::max(nf, 1);
      const bool dense = mean_front > 400000 && ctx->front_dense_ok;
      const int nt = dense ? FC_THREADS_DENSE : FC_THREADS;
      const size_t sm_f = front_cluster_smem_bytes(nbp, nt);
      dim3 grid(FC_CS, nframes);
#define FC_ARGS d_pts, ft, states, ctx->g, ctx->ap, has_intensity, nbp, nb, ctx->d_bin_ids.p, bin_off, wq, fits, ctx->d_sorted.p
      if (dense) {
        if (ctx->fast_bin) k_front_cluster<true, CLS_L2_MAX, FC_THREADS_DENSE><<<grid, nt, sm_f, s>>>(FC_ARGS);
        else k_front_cluster<false, CLS_L2_MAX, FC_THREADS_DENSE><<<grid, nt, sm_f, s>>>(FC_ARGS);
      } else {
        if (ctx->fast_bin) k_front_cluster<true, CLS_L2_MAX, FC_THREADS><<<grid, nt, sm_f, s>>>(FC_ARGS);
        else k_front_cluster<false, CLS_L2_MAX, FC_THREADS><<<grid, nt, sm_f, s>>>(FC_ARGS);
      }
#undef FC_ARGS
      ++ctx->launches;
    }
    STAGE_MARK(); STAGE_MARK(); STAGE_MARK();
  } else {
    // the three stand-alone kernels (PWPP_FRONT=0)
    if (max_chunks > 0) {
      dim3 grid(max_chunks, nframes);
      const size_t sm_h = nbp * sizeof(unsigned int);
#define HIST_ARGS d_pts, ft, states, ctx->g, ctx->ap, has_intensity, nbp, ctx->d_bin_ids.p, ctx->d_chist.p
      if (!ctx->fast_bin) k_bin_hist<false, 0><<<grid, CHUNK_THREADS, sm_h, s>>>(HIST_ARGS);
      else k_bin_hist<true, 2><<<grid, CHUNK_THREADS, sm_h, s>>>(HIST_ARGS);
#undef HIST_ARGS
      ++ctx->launches;
    }
    STAGE_MARK();
    k_bin_scan<CLS_L2_MAX><<<nframes, 512, (nbp + 1) * sizeof(int), s>>>(ft, nbp, nb, ctx->ap.num_min_pts, ctx->d_chist.p, ctx->d_cbase.p, bin_off, wq, fits);
    ++ctx->launches;
    STAGE_MARK();
    if (max_chunks > 0) {
      dim3 grid(max_chunks, nframes);
      const size_t sm_sc = (size_t) (CHUNK_THREADS / 32) * nbp * sizeof(unsigned int);
      k_scatter<false, 4><<<grid, CHUNK_THREADS, sm_sc, s>>>(d_pts, ft, nbp, ctx->d_bin_ids.p, ctx->d_cbase.p, ctx->d_sorted.p);
      ++ctx->launches;
    }
    STAGE_MARK();
  }
  // persistent fit kernels, one per patch-size class (queues were filled by k_bin_scan). The classes are independent:
  // unless per-stage timing is requested they run on side streams so that the tail of one class (few long patches
  // left) overlaps the start of the next.
#define FIT_ARGS ctx->d_sorted.p, ft, states, ctx->g, ctx->ap, nbp, bin_off, wq, ctx->d_part.p, fits
  const bool serial_fit = ctx->sw_serial_fit;   // PWPP_SERIAL_FIT: diagnostic switch
  // persistent grids are sized for a full GPU; a small call (one frame per call is the reference's pattern) would launch hundreds
  // of CTAs per class that find their queue empty and, worse, keep the six classes from running side by side: cap the grid
  // by what the call can hold (a frame has at most a few hundred patches)
  auto launch_fit = [&](int c, cudaStream_t st) {
    const FitLaunch& k = fitk[c];
    if (!k.fn) return;
    const long long cap = (long long) nframes * (k.threads <= 128 ? 96 : 48);
    const int grid = (int) std::min<long long>(k.grid, std::max<long long>(1, cap));
    k.fn<<<grid, k.threads, k.smem, st>>>(FIT_ARGS);
    ++ctx->launches;
  };
  // reference emission order inside every fitted patch (pwpp_order.cuh): one launch per class (X, L3, L2, L1 with a CTA sized to the
  // class; M + S one warp per patch). A sort only permutes `part` inside the patches of its own class, so it runs on its class's
  // side stream right behind the fit kernel, beside the other classes' fits and beside k_gle (which only reads the patch records);
  // k_emit joins everything. On a one-frame call the sort of the largest patch (~50 us) and the ring walk of k_gle (~50 us) are
  // both pure latency.
  int* heads = ctx->d_wq_ctr.p + 2 * NUM_CLASSES;
  auto launch_order = [&](int q, cudaStream_t so) {   // q: 0 = X, 1 = L3, 2 = L2, 3 = L1, 4 = M + S
    const int grid = (int) std::min<long long>(ctx->order_k[q].grid, std::max<long long>(1, (long long) nframes * (q == 4 ? 64 : 32)));
    const int th = ctx->order_k[q].threads;
    const size_t sm = ctx->order_k[q].smem;
    switch (q) {
      case 0: k_order_cta<512, 5><<<grid, th, sm, so>>>(ctx->d_sorted.p, wq, heads + q, ctx->d_part.p); break;
      case 1: k_order_cta<512, 4><<<grid, th, sm, so>>>(ctx->d_sorted.p, wq, heads + q, ctx->d_part.p); break;
      case 2: k_order_cta<256, 3><<<grid, th, sm, so>>>(ctx->d_sorted.p, wq, heads + q, ctx->d_part.p); break;
      case 3: k_order_cta<128, 2><<<grid, th, sm, so>>>(ctx->d_sorted.p, wq, heads + q, ctx->d_part.p); break;
      default: k_order_warp<<<grid, th, 0, so>>>(ctx->d_sorted.p, wq, heads + q, ctx->d_part.p); break;
    }
    ++ctx->launches;
  };
  const bool aside = !(prof || serial_fit);
  if (!aside) {
    static const int order[NUM_CLASSES] = {0, 4, 3, 2, 1, 5};   // stage slots: S, L3, L2, L1, M, X
    for (int q = 0; q < NUM_CLASSES; ++q) { launch_fit(order[q], s); STAGE_MARK(); }
    if (ctx->order_mode) for (int q = 0; q < ORD_NUM_HEADS; ++q) launch_order(q, s);
  } else {
    // side stream of every class, longest patches first: L3, L2, L1, M, S, X
    static const int cls_of_side[NUM_SIDE] = {4, 3, 2, 1, 0, 5};
    static const int order_of_side[NUM_SIDE] = {1, 2, 3, 4, -1, 0};   // the sort that follows the fit on that stream (M + S: behind M)
    CU_TRY(cudaEventRecord(ctx->ev_fork, s));
    for (int q = 0; q < NUM_SIDE; ++q) {
      CU_TRY(cudaStreamWaitEvent(ctx->side[q], ctx->ev_fork, 0));
      launch_fit(cls_of_side[q], ctx->side[q]);
      CU_TRY(cudaEventRecord(ctx->ev_fit[q], ctx->side[q]));
    }
    if (ctx->order_mode) {
      for (int q = 0; q < NUM_SIDE; ++q) {
        if (order_of_side[q] < 0) continue;
        if (order_of_side[q] == 4) CU_TRY(cudaStreamWaitEvent(ctx->side[q], ctx->ev_fit[4], 0));   // the warp sort also covers class S (side 4)
        launch_order(order_of_side[q], ctx->side[q]);
        CU_TRY(cudaEventRecord(ctx->ev_join[q], ctx->side[q]));
      }
    }
    for (int q = 0; q < NUM_SIDE; ++q) CU_TRY(cudaStreamWaitEvent(s, ctx->ev_fit[q], 0));   // k_gle needs every fit, no sort
    stage += 6;
  }
#undef FIT_ARGS
  int* d_ng = ctx->d_counts.p + f0;
  int* d_np = ctx->d_counts.p + ctx->num_streams + f0;
  int* d_nd = ctx->d_counts.p + 2 * ctx->num_streams + f0;
  {
    const size_t gle_smem = gle_smem_bytes(ctx->max_sectors);
    k_gle<<<nframes, 32, gle_smem, s>>>(ft, states, hist, ctx->hcap, ctx->g, ctx->ap, nbp, ctx->max_sectors, bin_off, fits, segs, d_ng, d_np, centers, normals, d_nd);
    ++ctx->launches;
  }
  STAGE_MARK();
  if (aside && ctx->order_mode) for (int q = 0; q < NUM_SIDE; ++q) if (q != 4) CU_TRY(cudaStreamWaitEvent(s, ctx->ev_join[q], 0));
  if (max_chunks > 0) {
    const long long max_pts = (long long) max_chunks * CHUNK_PTS;   // upper bound of the largest frame of the range
    dim3 grid((unsigned) ((max_pts + (long long) EMIT_TILE * EMIT_WARPS - 1) / ((long long) EMIT_TILE * EMIT_WARPS)), nframes);
    k_emit<<<grid, EMIT_WARPS * 32, 0, s>>>(ft, ctx->g, nbp, bin_off, fits, segs, ctx->d_part.p, ctx->d_sorted.p, ctx->d_out_idx.p);
    ++ctx->launches;
  }
  STAGE_MARK();
#undef STAGE_MARK
  if (prof) ctx->stage_valid = true;
  CU_TRY(cudaGetLastError());
  ctx->last_pts = d_pts;
  ctx->last_stream = s;
  return PWPP_OK;
}

// Small calls are launch-bound (ten kernels of a few microseconds each plus the fork / join events of the fit kernels): the
// sequence is captured once per (frames, intensity flag, grid size class, buffer generation) and replayed as one CUDA
// graph. Only for calls whose input sits in the ctx's own upload buffer (the host entry point), so the captured pointers
// stay valid; PWPP_GRAPH=0 switches it off.
constexpr int GRAPH_MAX_FRAMES = 16;
int launch_range(pwpp_ctx* ctx, int f0, int nf, const float4* d_pts, int has_intensity, cudaStream_t s, bool prof) {
  if (prof || !ctx->sw_graph || nf > GRAPH_MAX_FRAMES || f0 != 0 || d_pts != ctx->d_in.p) return launch_range_impl(ctx, f0, nf, d_pts, has_intensity, s, prof, 0);
  int max_chunks = 0;
  for (int f = 0; f < nf; ++f) max_chunks = std::max(max_chunks, ctx->chunk_off[f + 1] - ctx->chunk_off[f]);
  const int capc = std::max(8, (max_chunks + 7) & ~7);
  const int slot = has_intensity ? 1 : 0;
  const pwpp_ctx::GraphKey key{nf, has_intensity, capc, g_alloc_gen, (const void*) d_pts};
  pwpp_ctx::GraphKey& have = ctx->gkey[slot];
  if (!ctx->gexec[slot] || have.nf != key.nf || have.chunks != key.chunks || have.gen != key.gen || have.pts != key.pts) {
    if (ctx->gexec[slot]) { cudaGraphExecDestroy(ctx->gexec[slot]); ctx->gexec[slot] = nullptr; }
    const long long l0 = ctx->launches;
    CU_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int rc = launch_range_impl(ctx, 0, nf, d_pts, has_intensity, s, false, capc);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (rc != PWPP_OK || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      ctx->sw_graph = 0;   // capture is not available here: fall back to plain launches for good
      ctx->launches = l0;
      return launch_range_impl(ctx, 0, nf, d_pts, has_intensity, s, false, 0);
    }
    const cudaError_t e2 = cudaGraphInstantiate(&ctx->gexec[slot], graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) { ctx->gexec[slot] = nullptr; ctx->sw_graph = 0; ctx->launches = l0; cudaGetLastError(); return launch_range_impl(ctx, 0, nf, d_pts, has_intensity, s, false, 0); }
    ctx->glaunches[slot] = ctx->launches - l0;
    ctx->launches = l0;
    have = key;
  }
  CU_TRY(cudaGraphLaunch(ctx->gexec[slot], s));
  ctx->launches += ctx->glaunches[slot];
  ctx->last_pts = d_pts;
  ctx->last_stream = s;
  return PWPP_OK;
}

// Optional sub-batching of a big call (PWPP_SUBBATCH_POINTS=n: consecutive launch sequences of at most n points).
// Measured on B200 with correctly placed CUDA events: slower than one sequence per call at every size tried (each
// persistent fit kernel already fills the GPU; splitting only adds tails), so the default is one sequence.
long long subbatch_points() {
  static long long v = [] {
    const char* e = std::getenv("PWPP_SUBBATCH_POINTS");
    return e ? std::atoll(e) : (1LL << 62);   // default: one sequence per call (sub-batching measured slower on B200)
  }();
  return v;
}

int run_path(pwpp_ctx* ctx, int nframes, const float4* d_pts, int has_intensity, cudaStream_t s) {
  int rc = prepare_call(ctx, nframes, s);
  if (rc) return rc;
  if (ctx->profiling) return launch_range(ctx, 0, nframes, d_pts, has_intensity, s, true);
  const long long cap = subbatch_points();
  int f0 = 0;
  while (f0 < nframes) {
    int f1 = f0 + 1;
    while (f1 < nframes && ctx->pt_off[f1 + 1] - ctx->pt_off[f0] <= cap) ++f1;
    rc = launch_range(ctx, f0, f1 - f0, d_pts, has_intensity, s, false);
    if (rc) return rc;
    f0 = f1;
  }
  return PWPP_OK;
}

int check_frame(pwpp_ctx* ctx, int f) {
  if (!ctx) return fail(PWPP_ERR_INVALID_ARG, "ctx is NULL");
  if (f < 0 || f >= ctx->last_nframes) return fail(PWPP_ERR_INVALID_ARG, "frame index outside the last estimate call");
  return PWPP_OK;
}

int fetch_counts(pwpp_ctx* ctx) {
  if (ctx->counts_fetched) return PWPP_OK;
  int rc = bind_device(ctx);
  if (rc) return rc;
  CU_TRY(ctx->h_counts.reserve((size_t) 3 * ctx->num_streams));
  CU_TRY(cudaMemcpyAsync(ctx->h_counts.p, ctx->d_counts.p, (size_t) 3 * ctx->num_streams * sizeof(int), cudaMemcpyDeviceToHost, ctx->last_stream));
  CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  ctx->counts_fetched = true;
  return PWPP_OK;
}
int fetch_indices(pwpp_ctx* ctx) {
  if (ctx->idx_fetched) return PWPP_OK;
  int rc = fetch_counts(ctx);
  if (rc) return rc;
  CU_TRY(ctx->h_out_idx.reserve((size_t) std::max<long long>(ctx->last_total, 1)));
  if (ctx->last_total > 0)
    CU_TRY(cudaMemcpyAsync(ctx->h_out_idx.p, ctx->d_out_idx.p, (size_t) ctx->last_total * sizeof(int), cudaMemcpyDeviceToHost, ctx->last_stream));
  CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  ctx->idx_fetched = true;
  return PWPP_OK;
}
int fetch_patches(pwpp_ctx* ctx) {
  if (ctx->patches_fetched) return PWPP_OK;
  int rc = fetch_counts(ctx);
  if (rc) return rc;
  const size_t n = (size_t) ctx->last_nframes * ctx->g.nbins * 3;
  CU_TRY(ctx->h_centers.reserve(n));
  CU_TRY(ctx->h_normals.reserve(n));
  CU_TRY(cudaMemcpyAsync(ctx->h_centers.p, ctx->d_centers.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->last_stream));
  CU_TRY(cudaMemcpyAsync(ctx->h_normals.p, ctx->d_normals.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->last_stream));
  CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  ctx->patches_fetched = true;
  return PWPP_OK;
}

inline int frame_n(const pwpp_ctx* ctx, int f) { return (int) (ctx->pt_off[f + 1] - ctx->pt_off[f]); }

}  // namespace

extern "C" {

void pwpp_params_default(pwpp_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->verbose = 0; p->enable_RNR = 1; p->enable_RVPF = 1; p->enable_TGR = 1;            // patchworkpp.h:80-83
  p->num_iter = 3; p->num_lpr = 20; p->num_min_pts = 10; p->num_zones = 4; p->num_rings_of_interest = 4;  // :85-89
  p->RNR_ver_angle_thr = -15.0; p->RNR_intensity_thr = 0.2;                             // :91-92
  p->sensor_height = 1.723; p->th_seeds = 0.125; p->th_dist = 0.125; p->th_seeds_v = 0.25; p->th_dist_v = 0.1;  // :94-98
  p->max_range = 80.0; p->min_range = 2.7; p->uprightness_thr = 0.707; p->adaptive_seed_selection_margin = -1.2;  // :99-102
  p->intensity_thr = 0.0;
  const int sectors[4] = {16, 32, 54, 32}, rings[4] = {2, 4, 4, 4};                     // :104-105
  for (int k = 0; k < 4; ++k) { p->num_sectors_each_zone[k] = sectors[k]; p->num_rings_each_zone[k] = rings[k]; p->elevation_thr[k] = 0; p->flatness_thr[k] = 0; }
  p->max_flatness_storage = 1000; p->max_elevation_storage = 1000;                      // :107-108
}

const char* pwpp_last_error(void) { return g_last_error.c_str(); }
int pwpp_abi_version(void) { return PWPP_ABI_VERSION; }
int pwpp_num_bins(const pwpp_ctx* ctx) { return ctx ? ctx->g.nbins : 0; }

int pwpp_create(const pwpp_params* params, int device, int num_streams, int64_t max_points_per_frame, pwpp_ctx** out) {
  if (!out) return fail(PWPP_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  int rc = validate_params(params);
  if (rc) return rc;
  if (num_streams < 1 || num_streams > 65535) return fail(PWPP_ERR_INVALID_ARG, "num_streams must be in [1,65535]");
  if (max_points_per_frame < 0) return fail(PWPP_ERR_INVALID_ARG, "max_points_per_frame < 0");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PWPP_ERR_NO_DEVICE, std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                        "); this library has no CPU path");
  if (device < 0 || device >= ndev) return fail(PWPP_ERR_INVALID_ARG, "device index out of range");
  pwpp_ctx* ctx = new pwpp_ctx();
  ctx->prm = *params;
  ctx->device = device;
  ctx->num_streams = num_streams;
  build_geometry(*params, ctx->g, ctx->ap, ctx->fast_bin);
  ctx->sw_serial_fit = env_int("PWPP_SERIAL_FIT", 0, 0, 1) != 0;       // diagnostic: the fit kernels one after another on the call's stream
  ctx->sw_front = env_int("PWPP_FRONT", PWPP_FRONT_DEFAULT, 0, 1);        // 1: cluster-per-frame front end, 0: k_bin_hist + k_bin_scan + k_scatter
  ctx->sw_patch = env_int("PWPP_FIT_PATCH", PWPP_FIT_PATCH_DEFAULT, 0, 1);   // 1: patches above 512 points on k_fit_patch
  ctx->sw_graph = env_int("PWPP_GRAPH", 1, 0, 1);
  ctx->small_call_frames = env_int("PWPP_SMALL_CALL", PWPP_SMALL_CALL_DEFAULT, 0, 64);   // calls of at most this many frames take the small-call kernels (0: never)
  ctx->nbp = ((ctx->g.nbins + PW_NUM_PSEUDO + 31) / 32) * 32;
  int max_sectors = 0;
  for (int k = 0; k < 4; ++k) max_sectors = std::max(max_sectors, ctx->g.num_sectors[k]);
  ctx->hcap = std::max(params->max_elevation_storage, params->max_flatness_storage) + 4 * max_sectors + 64;
  ctx->max_sectors = max_sectors;
#define CU_TRY_CTX(expr)                                                                                  \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e);                                 \
      pwpp_destroy(ctx);                                                                                  \
      return fail(PWPP_ERR_CUDA, m);                                                                      \
    }                                                                                                     \
  } while (0)
  CU_TRY_CTX(cudaSetDevice(device));
  CU_TRY_CTX(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CU_TRY_CTX(cudaStreamCreateWithFlags(&ctx->stream_h2d, cudaStreamNonBlocking));
  CU_TRY_CTX(cudaStreamCreateWithFlags(&ctx->stream_d2h, cudaStreamNonBlocking));
  CU_TRY_CTX(cudaEventCreate(&ctx->ev0));
  CU_TRY_CTX(cudaEventCreate(&ctx->ev1));
  CU_TRY_CTX(cudaEventCreate(&ctx->ev_begin));
  CU_TRY_CTX(cudaEventCreate(&ctx->ev_end));
  for (int i = 0; i <= PWPP_NUM_STAGES; ++i) CU_TRY_CTX(cudaEventCreate(&ctx->stage_ev[i]));
  for (int i = 0; i < 2; ++i) CU_TRY_CTX(cudaEventCreateWithFlags(&ctx->tab_ev[i], cudaEventDisableTiming));
  CU_TRY_CTX(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  for (int q = 0; q < NUM_SIDE; ++q) {
    CU_TRY_CTX(cudaStreamCreateWithFlags(&ctx->side[q], cudaStreamNonBlocking));
    CU_TRY_CTX(cudaEventCreateWithFlags(&ctx->ev_fit[q], cudaEventDisableTiming));
    CU_TRY_CTX(cudaEventCreateWithFlags(&ctx->ev_join[q], cudaEventDisableTiming));
  }
  CU_TRY_CTX(ctx->d_states.reserve(num_streams));
  CU_TRY_CTX(ctx->d_states_init.reserve(num_streams));
  {
    std::vector<StreamState> v(num_streams);
    for (auto& st : v) init_state(*params, st);
    CU_TRY_CTX(cudaMemcpy(ctx->d_states_init.p, v.data(), v.size() * sizeof(StreamState), cudaMemcpyHostToDevice));
  }
  CU_TRY_CTX(ctx->d_hist.reserve((size_t) num_streams * 2 * 4 * ctx->hcap));
  CU_TRY_CTX(ctx->d_counts.reserve((size_t) 3 * num_streams));
  CU_TRY_CTX(ctx->d_wq_ctr.reserve(2 * NUM_CLASSES + ORD_NUM_HEADS));
  {
    cudaDeviceProp prop;
    CU_TRY_CTX(cudaGetDeviceProperties(&prop, device));
    // Fit kernel of every patch-size class (launch shapes from the r01 / r02 measurements, profiles/):
    //   S  <= 64      k_fit_resident: 8 lanes x 8 register slots per patch
    //   M  <= 512     k_fit_warp, patch staged in shared memory, plane + moment sums in shared memory, 3 CTAs/SM (r02: 0.72 -> 0.68 ms)
    //   L1 <= 2048    k_fit_warp streaming from L2, the same  (0.83 -> 0.76 ms; 4 CTAs/SM at 64 registers: 0.81; a cp.async chunk ring: 0.77)
    //                 Both are instruction-fetch sensitive (125 KB of code each): the passes take TWO rows per batch, the LPR scans two
    //                 loads in flight — r02 ab17-19: rows per batch 1 / 2 / 4 / 8: L1 0.63 / 0.60 / 0.74 / 1.06 ms, M 0.69 / 0.69 / 0.76;
    //                 LPR scans 8 / 4 / 2 in flight: M 0.69 / 0.64 / 0.63, L1 0.60 / 0.58 / 0.57 ms.
    //                 PWPP_FIT_PATCH=1 (and calls of <= 4 frames): k_fit_patch, one patch per CTA of 4 / 8 / 16 warps held in registers
    //   L2 <= 4096    k_fit_cta, plane in shared memory, 3 CTAs/SM
    //   L3 <= 8192    k_fit_cta, 2 CTAs/SM
    //   X  >  8192    k_fit_big (dense sensors)
    const size_t sm_m = FITW_WARPS * CLS_M_MAX * sizeof(float4), sm_l2 = 3 * 4096 * sizeof(float), sm_l3 = 3 * 8192 * sizeof(float);
    ctx->fit[0] = {k_fit_resident<8, 8, 0, 2>, 0, FIT_THREADS, 0};
    ctx->fit[1] = {k_fit_warp<true, 1, 1, 2, 3, false, true>, 0, FITW_WARPS * 32, sm_m};
    ctx->fit[2] = {k_fit_warp<false, 2, 2, 2, 3, false, true>, 0, FITW_WARPS * 32, 0};
    ctx->fit[3] = {k_fit_cta<4096, 3, 3, 8, true, true>, 0, FIT_THREADS, sm_l2};
    ctx->fit[4] = {k_fit_cta<8192, 4, 2, 8, true>, 0, FIT_THREADS, sm_l3};
    ctx->fit[5] = {k_fit_big<16, 1, true>, 0, 512, 0};
    if (ctx->sw_patch) {
      ctx->fit[2] = {k_fit_patch<4, 4, 2>, 0, 4 * 32, (size_t) 4 * FP_STG * sizeof(float4)};
      ctx->fit[3] = {k_fit_patch<8, 2, 3>, 0, 8 * 32, (size_t) 8 * FP_STG * sizeof(float4)};
      ctx->fit[4] = {k_fit_patch<16, 1, 4>, 0, 16 * 32, (size_t) 16 * FP_STG * sizeof(float4)};
    }
    for (int c = 0; c < NUM_CLASSES; ++c) ctx->fit_small[c] = ctx->fit[c];
    ctx->fit_small[2] = {k_fit_patch<4, 4, 2>, 0, 4 * 32, (size_t) 4 * FP_STG * sizeof(float4)};
    ctx->fit_small[3] = {k_fit_patch<8, 2, 3>, 0, 8 * 32, (size_t) 8 * FP_STG * sizeof(float4)};
    ctx->fit_small[4] = {k_fit_patch<16, 1, 4>, 0, 16 * 32, (size_t) 16 * FP_STG * sizeof(float4)};
    for (int c = 0; c < 2 * NUM_CLASSES; ++c) {
      FitLaunch& k = c < NUM_CLASSES ? ctx->fit[c] : ctx->fit_small[c - NUM_CLASSES];
      if (k.smem > 0) CU_TRY_CTX(cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) k.smem));
      int per_sm = 1;
      CU_TRY_CTX(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k.fn, k.threads, k.smem));
      k.grid = std::max(1, per_sm) * prop.multiProcessorCount;
    }
    {
      typedef void (*OrdKernel)(const float4*, WorkQueues, int*, int*);
      const OrdKernel fn[ORD_NUM_HEADS] = {k_order_cta<512, 5>, k_order_cta<512, 4>, k_order_cta<256, 3>, k_order_cta<128, 2>, k_order_warp};
      const int th[ORD_NUM_HEADS] = {512, 512, 256, 128, ORD_WARP_THREADS};
      for (int q = 0; q < ORD_NUM_HEADS; ++q) {
        const size_t sm = q < 4 ? ord_cta_smem_bytes(th[q]) : 0;
        if (sm > 48 * 1024) CU_TRY_CTX(cudaFuncSetAttribute(fn[q], cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
        int per_sm = 1;
        CU_TRY_CTX(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn[q], th[q], sm));
        ctx->order_k[q].grid = std::max(1, per_sm) * prop.multiProcessorCount;
        ctx->order_k[q].threads = th[q];
        ctx->order_k[q].smem = sm;
      }
    }
    const size_t gle_smem = gle_smem_bytes(max_sectors);
    if (gle_smem > 48 * 1024) CU_TRY_CTX(cudaFuncSetAttribute(k_gle, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gle_smem));
  }
  {
    const size_t scat = (size_t) (CHUNK_THREADS / 32) * ctx->nbp * sizeof(unsigned int);
    if (scat > 48 * 1024) CU_TRY_CTX(cudaFuncSetAttribute(k_scatter<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) scat));
    const size_t sm_f = front_cluster_smem_bytes(ctx->nbp, FC_THREADS), sm_fd = front_cluster_smem_bytes(ctx->nbp, FC_THREADS_DENSE);
    if (sm_f > 220 * 1024) ctx->sw_front = 0;   // (thousands of bins: the per-warp count tables no longer fit next to the tiles)
    ctx->front_dense_ok = sm_fd <= 220 * 1024;
    if (ctx->sw_front) {
      CU_TRY_CTX(cudaFuncSetAttribute(k_front_cluster<true, CLS_L2_MAX, FC_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_f));
      CU_TRY_CTX(cudaFuncSetAttribute(k_front_cluster<false, CLS_L2_MAX, FC_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_f));
      if (ctx->front_dense_ok) {
        CU_TRY_CTX(cudaFuncSetAttribute(k_front_cluster<true, CLS_L2_MAX, FC_THREADS_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_fd));
        CU_TRY_CTX(cudaFuncSetAttribute(k_front_cluster<false, CLS_L2_MAX, FC_THREADS_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_fd));
      }
    }
  }
  *out = ctx;
  rc = pwpp_reset_all(ctx);
  if (rc) { pwpp_destroy(ctx); *out = nullptr; return rc; }
  CU_TRY_CTX(cudaStreamSynchronize(ctx->stream));   // the initial state is in place before any caller stream can read it
  if (max_points_per_frame > 0) {
    const size_t tot = (size_t) max_points_per_frame * num_streams;
    CU_TRY_CTX(ctx->d_bin_ids.reserve(tot));
    CU_TRY_CTX(ctx->d_sorted.reserve(tot));
    CU_TRY_CTX(ctx->d_part.reserve(tot));
    CU_TRY_CTX(ctx->d_out_idx.reserve(tot));
  }
#undef CU_TRY_CTX
  return PWPP_OK;
}

void pwpp_destroy(pwpp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i <= PWPP_NUM_STAGES; ++i) if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
  for (int i = 0; i < 2; ++i) if (ctx->gexec[i]) cudaGraphExecDestroy(ctx->gexec[i]);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (int i = 0; i < 2; ++i) if (ctx->tab_ev[i]) cudaEventDestroy(ctx->tab_ev[i]);
  for (int q = 0; q < NUM_SIDE; ++q) { if (ctx->ev_fit[q]) cudaEventDestroy(ctx->ev_fit[q]); if (ctx->ev_join[q]) cudaEventDestroy(ctx->ev_join[q]); if (ctx->side[q]) cudaStreamDestroy(ctx->side[q]); }
  ctx->d_states.release(); ctx->d_states_init.release(); ctx->d_hist.release(); ctx->d_in.release(); ctx->d_cm.release(); ctx->d_pt_off.release(); ctx->d_chunk_off.release();
  ctx->d_bin_ids.release(); ctx->d_chist.release(); ctx->d_cbase.release(); ctx->d_bin_off.release(); ctx->d_sorted.release();
  ctx->d_part.release(); ctx->d_labels.release(); ctx->d_fits.release(); ctx->d_segs.release(); ctx->d_wq_ctr.release();
  for (int c = 0; c < NUM_CLASSES; ++c) ctx->d_wq_items[c].release();
  ctx->d_out_idx.release(); ctx->d_counts.release();
  ctx->d_centers.release(); ctx->d_normals.release(); ctx->d_xyz.release();
  ctx->h_in.release(); for (int i = 0; i < 2; ++i) { ctx->h_pt_off_buf[i].release(); ctx->h_chunk_off_buf[i].release(); } ctx->h_out_idx.release(); ctx->h_counts.release();
  ctx->h_centers.release(); ctx->h_normals.release();
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
  if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->stream_h2d) cudaStreamDestroy(ctx->stream_h2d);
  if (ctx->stream_d2h) cudaStreamDestroy(ctx->stream_d2h);
  delete ctx;
}

int pwpp_reset_stream(pwpp_ctx* ctx, int f) {
  if (!ctx || f < 0 || f >= ctx->num_streams) return fail(PWPP_ERR_INVALID_ARG, "bad stream index");
  int rc = bind_device(ctx);
  if (rc) return rc;
  cudaStream_t s = ctx->last_stream ? ctx->last_stream : ctx->stream;
  CU_TRY(cudaMemcpyAsync(ctx->d_states.p + f, ctx->d_states_init.p + f, sizeof(StreamState), cudaMemcpyDeviceToDevice, s));
  return PWPP_OK;
}
int pwpp_reset_all(pwpp_ctx* ctx) {
  if (!ctx) return fail(PWPP_ERR_INVALID_ARG, "ctx is NULL");
  int rc = bind_device(ctx);
  if (rc) return rc;
  cudaStream_t s = ctx->last_stream ? ctx->last_stream : ctx->stream;
  CU_TRY(cudaMemcpyAsync(ctx->d_states.p, ctx->d_states_init.p, (size_t) ctx->num_streams * sizeof(StreamState), cudaMemcpyDeviceToDevice, s));
  return PWPP_OK;
}

void* pwpp_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { g_last_error = "cudaMallocHost failed"; return nullptr; }
  return p;
}
void pwpp_host_free(void* p) { if (p) cudaFreeHost(p); }

int pwpp_set_profiling(pwpp_ctx* ctx, int enabled) {
  if (!ctx) return fail(PWPP_ERR_INVALID_ARG, "ctx is NULL");
  ctx->profiling = enabled != 0;
  ctx->stage_valid = false;
  return PWPP_OK;
}
int pwpp_stage_times_ms(pwpp_ctx* ctx, float* ms) {
  if (!ctx || !ms) return fail(PWPP_ERR_INVALID_ARG, "NULL argument");
  if (!ctx->stage_valid) return fail(PWPP_ERR_INVALID_ARG, "no profiled estimate call yet (pwpp_set_profiling)");
  int rc = bind_device(ctx);
  if (rc) return rc;
  CU_TRY(cudaEventSynchronize(ctx->stage_ev[PWPP_NUM_STAGES]));
  for (int i = 0; i < PWPP_NUM_STAGES; ++i) CU_TRY(cudaEventElapsedTime(&ms[i], ctx->stage_ev[i], ctx->stage_ev[i + 1]));
  return PWPP_OK;
}
const char* pwpp_stage_name(int stage) {
  static const char* names[PWPP_NUM_STAGES] = {"k_bin_hist", "k_bin_scan", "k_scatter", "k_fit_S", "k_fit_L3", "k_fit_L2", "k_fit_L1", "k_fit_M", "k_fit_X", "k_gle", "k_emit"};
  return (stage >= 0 && stage < PWPP_NUM_STAGES) ? names[stage] : "";
}
int64_t pwpp_launch_count(const pwpp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int pwpp_estimate_host(pwpp_ctx* ctx, int nframes, const float* const* pts, const int64_t* n, int cols, int64_t row_stride, int64_t col_stride) {
  if (!ctx || !pts || !n) return fail(PWPP_ERR_INVALID_ARG, "NULL argument");
  if (nframes < 1 || nframes > ctx->num_streams) return fail(PWPP_ERR_INVALID_ARG, "nframes must be in [1, num_streams]");
  if (cols != 3 && cols != 4) return fail(PWPP_ERR_INVALID_ARG, "cols must be 3 or 4 (x,y,z[,intensity])");
  int rc = bind_device(ctx);
  if (rc) return rc;
  const auto t0 = std::chrono::steady_clock::now();
  ctx->pt_off.assign(nframes + 1, 0);
  for (int f = 0; f < nframes; ++f) {
    if (n[f] < 0 || (n[f] > 0 && !pts[f])) return fail(PWPP_ERR_INVALID_ARG, "bad frame pointer/size");
    ctx->pt_off[f + 1] = ctx->pt_off[f] + n[f];
  }
  const long long total = ctx->pt_off[nframes];
  CU_TRY(ctx->d_in.reserve((size_t) std::max<long long>(total, 1)));
  cudaStream_t s = ctx->stream, s_in = ctx->stream_h2d, s_out = ctx->stream_d2h;
  ctx->call_times_valid = false;
  // the staging buffers of the previous call must not be in flight any more (nor a device-input call on a caller's stream)
  if (ctx->last_stream && ctx->last_stream != s) CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  CU_TRY(cudaStreamSynchronize(s));
  CU_TRY(cudaStreamSynchronize(s_in));
  CU_TRY(cudaStreamSynchronize(s_out));
  rc = prepare_call(ctx, nframes, s);
  if (rc) return rc;
  CU_TRY(ctx->h_out_idx.reserve((size_t) std::max<long long>(total, 1)));
  CU_TRY(ctx->h_counts.reserve((size_t) 3 * ctx->num_streams));
  const bool packed = (cols == 4 && col_stride == 1 && row_stride == 4);
  bool staged_any = false;
  // Chunks of frames flow through three streams: H2D of chunk k+1 | kernels of chunk k | D2H of chunk k-1.
  // ~64 MB of points per chunk keeps every stage busy without delaying the first kernels.
  int chunk_frames = nframes;
  if (nframes > 1 && total > 0) {
    const long long per_frame = std::max<long long>(1, total / nframes);
    chunk_frames = (int) std::min<long long>(nframes, std::max<long long>(1, (4LL << 20) / per_frame));
  }
  const int nchunks = (nframes + chunk_frames - 1) / chunk_frames;
  // a call that is a single chunk has nothing to pipeline: copies and kernels share one stream (no cross-stream event hops,
  // ~10 us each on the critical path of a one-frame call) and four events bracket its three phases (pwpp_call_times_us)
  const bool one_stream = (nchunks == 1);
  if (one_stream) { s_in = s; s_out = s; CU_TRY(cudaEventRecord(ctx->ev_begin, s)); }
  for (int k = 0; k < nchunks; ++k) {
    const int f0 = k * chunk_frames, f1 = std::min(nframes, f0 + chunk_frames);
    for (int f = f0; f < f1; ++f) {
      const int64_t cnt = n[f];
      if (cnt == 0) continue;
      bool pinned = false;
      if (packed) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, pts[f]) == cudaSuccess) pinned = (attr.type == cudaMemoryTypeHost);
        else cudaGetLastError();
      }
      if (pinned) {  // page-locked caller buffer: DMA straight from it (the private copy of H:152 is the device buffer)
        // frames that follow each other in the caller's buffer travel as one copy
        int f_last = f;
        int64_t run = cnt;
        while (f_last + 1 < f1 && n[f_last + 1] > 0 && pts[f_last + 1] == pts[f_last] + n[f_last] * 4) { ++f_last; run += n[f_last]; }
        CU_TRY(cudaMemcpyAsync(ctx->d_in.p + ctx->pt_off[f], pts[f], (size_t) run * sizeof(float4), cudaMemcpyHostToDevice, s_in));
        f = f_last;
        continue;
      }
      if (row_stride == 1 && col_stride == cnt && cnt > 1) {
        // column-major frame (what the Eigen overload hands over): its cols x n floats are one contiguous block; upload it as
        // it is and repack on the device instead of gathering 4 strided columns per point on the host
        if ((size_t) total * 4 > ctx->d_cm.cap) { CU_TRY(cudaStreamSynchronize(s_in)); CU_TRY(ctx->d_cm.reserve((size_t) total * 4)); }
        float* cm = ctx->d_cm.p + (size_t) ctx->pt_off[f] * 4;
        CU_TRY(cudaMemcpyAsync(cm, pts[f], (size_t) cnt * cols * sizeof(float), cudaMemcpyHostToDevice, s_in));
        k_repack_colmajor<<<(unsigned) ((cnt + 255) / 256), 256, 0, s_in>>>(cm, cnt, cols, ctx->d_in.p + ctx->pt_off[f]);
        ++ctx->launches;
        continue;
      }
      if (!staged_any) { CU_TRY(ctx->h_in.reserve((size_t) std::max<long long>(total, 1))); staged_any = true; }
      float4* dst = ctx->h_in.p + ctx->pt_off[f];
      const float* src = pts[f];
      if (packed) {
        std::memcpy(dst, src, (size_t) cnt * sizeof(float4));
      } else {
        for (int64_t i = 0; i < cnt; ++i) {
          const float* r = src + i * row_stride;
          dst[i] = make_float4(r[0], r[col_stride], r[2 * col_stride], cols == 4 ? r[3 * col_stride] : 0.f);
        }
      }
      CU_TRY(cudaMemcpyAsync(ctx->d_in.p + ctx->pt_off[f], dst, (size_t) cnt * sizeof(float4), cudaMemcpyHostToDevice, s_in));
    }
    CU_TRY(cudaEventRecord(ctx->ev0, s_in));
    if (!one_stream) CU_TRY(cudaStreamWaitEvent(s, ctx->ev0, 0));
    rc = launch_range(ctx, f0, f1 - f0, ctx->d_in.p, cols == 4 ? 1 : 0, s, ctx->profiling && nchunks == 1);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(ctx->ev1, s));
    if (!one_stream) CU_TRY(cudaStreamWaitEvent(s_out, ctx->ev1, 0));
    const long long o0 = ctx->pt_off[f0], o1 = ctx->pt_off[f1];
    if (f0 == 0 && f1 == ctx->num_streams) {   // the three count rows are one contiguous block when the chunk covers every stream
      CU_TRY(cudaMemcpyAsync(ctx->h_counts.p, ctx->d_counts.p, (size_t) 3 * ctx->num_streams * sizeof(int), cudaMemcpyDeviceToHost, s_out));
    } else {
      for (int q = 0; q < 3; ++q)
        CU_TRY(cudaMemcpyAsync(ctx->h_counts.p + (size_t) q * ctx->num_streams + f0, ctx->d_counts.p + (size_t) q * ctx->num_streams + f0,
                               (size_t) (f1 - f0) * sizeof(int), cudaMemcpyDeviceToHost, s_out));
    }
    if (o1 > o0) CU_TRY(cudaMemcpyAsync(ctx->h_out_idx.p + o0, ctx->d_out_idx.p + o0, (size_t) (o1 - o0) * sizeof(int), cudaMemcpyDeviceToHost, s_out));
  }
  if (one_stream) CU_TRY(cudaEventRecord(ctx->ev_end, s));
  CU_TRY(cudaStreamSynchronize(s_out));
  CU_TRY(cudaStreamSynchronize(s));
  ctx->call_times_valid = one_stream;
  ctx->counts_fetched = true;
  ctx->idx_fetched = true;
  ctx->last_time_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  return PWPP_OK;
}

int pwpp_estimate_device(pwpp_ctx* ctx, int nframes, const void* d_pts, const int64_t* h_offsets, int has_intensity, void* cuda_stream) {
  if (!ctx || !h_offsets) return fail(PWPP_ERR_INVALID_ARG, "NULL argument");
  if (nframes < 1 || nframes > ctx->num_streams) return fail(PWPP_ERR_INVALID_ARG, "nframes must be in [1, num_streams]");
  int rc = bind_device(ctx);
  if (rc) return rc;
  const auto t0 = std::chrono::steady_clock::now();
  ctx->pt_off.assign(nframes + 1, 0);
  for (int f = 0; f <= nframes; ++f) {
    ctx->pt_off[f] = h_offsets[f] - h_offsets[0];
    if (f > 0 && ctx->pt_off[f] < ctx->pt_off[f - 1]) return fail(PWPP_ERR_INVALID_ARG, "offsets must be non-decreasing");
  }
  if (ctx->pt_off[nframes] > 0 && !d_pts) return fail(PWPP_ERR_INVALID_ARG, "d_pts is NULL");
  cudaStream_t s = cuda_stream ? (cudaStream_t) cuda_stream : ctx->stream;
  // work buffers are reused stream-ordered: a call on a different stream than the previous one waits for it
  if (ctx->last_stream && ctx->last_stream != s) CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  if (!ctx->last_stream && s != ctx->stream) CU_TRY(cudaStreamSynchronize(ctx->stream));   // resets issued before the first call ran on the ctx's own stream
  rc = run_path(ctx, nframes, (const float4*) d_pts + h_offsets[0], has_intensity, s);
  if (rc) return rc;
  ctx->last_time_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  return PWPP_OK;
}

int pwpp_estimate_device_xyz(pwpp_ctx* ctx, int nframes, const void* d_xyz, const int64_t* h_offsets, void* cuda_stream) {
  if (!ctx || !h_offsets) return fail(PWPP_ERR_INVALID_ARG, "NULL argument");
  if (nframes < 1 || nframes > ctx->num_streams) return fail(PWPP_ERR_INVALID_ARG, "nframes must be in [1, num_streams]");
  int rc = bind_device(ctx);
  if (rc) return rc;
  const long long total = h_offsets[nframes] - h_offsets[0];
  if (total < 0 || (total > 0 && !d_xyz)) return fail(PWPP_ERR_INVALID_ARG, "bad offsets / d_xyz is NULL");
  cudaStream_t s = cuda_stream ? (cudaStream_t) cuda_stream : ctx->stream;
  if (ctx->last_stream && ctx->last_stream != s) CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  if (!ctx->last_stream && s != ctx->stream) CU_TRY(cudaStreamSynchronize(ctx->stream));
  CU_TRY(ctx->d_in.reserve((size_t) std::max<long long>(total, 1)));
  if (total > 0) {
    k_pad_xyz<<<(unsigned) ((total + 255) / 256), 256, 0, s>>>((const float*) d_xyz + 3 * h_offsets[0], total, ctx->d_in.p);
    ++ctx->launches;
  }
  std::vector<int64_t> offs(h_offsets, h_offsets + nframes + 1);
  for (auto& o : offs) o -= h_offsets[0];
  return pwpp_estimate_device(ctx, nframes, ctx->d_in.p, offs.data(), 0, s);
}

int pwpp_device_synchronize(pwpp_ctx* ctx) {
  if (!ctx) return fail(PWPP_ERR_INVALID_ARG, "ctx is NULL");
  int rc = bind_device(ctx);
  if (rc) return rc;
  CU_TRY(cudaDeviceSynchronize());
  return PWPP_OK;
}

int pwpp_synchronize(pwpp_ctx* ctx) {
  if (!ctx) return fail(PWPP_ERR_INVALID_ARG, "ctx is NULL");
  int rc = bind_device(ctx);
  if (rc) return rc;
  CU_TRY(cudaStreamSynchronize(ctx->last_stream ? ctx->last_stream : ctx->stream));
  return PWPP_OK;
}

int64_t pwpp_num_ground(pwpp_ctx* ctx, int f) {
  if (check_frame(ctx, f) || fetch_counts(ctx)) return -1;
  return ctx->h_counts.p[f];
}
int64_t pwpp_num_nonground(pwpp_ctx* ctx, int f) {
  if (check_frame(ctx, f) || fetch_counts(ctx)) return -1;
  return (int64_t) frame_n(ctx, f) - ctx->h_counts.p[f] - ctx->h_counts.p[2 * ctx->num_streams + f];
}
int pwpp_copy_ground_indices(pwpp_ctx* ctx, int f, int32_t* dst) {
  int rc = check_frame(ctx, f);
  if (rc) return rc;
  rc = fetch_indices(ctx);
  if (rc) return rc;
  const int ng = ctx->h_counts.p[f];
  if (ng > 0) std::memcpy(dst, ctx->h_out_idx.p + ctx->pt_off[f], (size_t) ng * sizeof(int32_t));
  return PWPP_OK;
}
int pwpp_copy_nonground_indices(pwpp_ctx* ctx, int f, int32_t* dst) {
  int rc = check_frame(ctx, f);
  if (rc) return rc;
  rc = fetch_indices(ctx);
  if (rc) return rc;
  const int ng = ctx->h_counts.p[f];
  const int nn = frame_n(ctx, f) - ng - ctx->h_counts.p[2 * ctx->num_streams + f];
  if (nn > 0) std::memcpy(dst, ctx->h_out_idx.p + ctx->pt_off[f] + ng, (size_t) nn * sizeof(int32_t));
  return PWPP_OK;
}

static int copy_xyz(pwpp_ctx* ctx, int f, float* dst, bool ground) {
  int rc = check_frame(ctx, f);
  if (rc) return rc;
  rc = fetch_counts(ctx);
  if (rc) return rc;
  const int ng = ctx->h_counts.p[f];
  const int nn = frame_n(ctx, f) - ng - ctx->h_counts.p[2 * ctx->num_streams + f];
  const int cnt = ground ? ng : nn;
  if (cnt <= 0) return PWPP_OK;
  CU_TRY(ctx->d_xyz.reserve((size_t) cnt * 3));
  const int* idx = ctx->d_out_idx.p + ctx->pt_off[f] + (ground ? 0 : ng);
  k_gather_xyz<<<(cnt + 255) / 256, 256, 0, ctx->last_stream>>>(ctx->last_pts + ctx->pt_off[f], idx, cnt, ctx->d_xyz.p);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpyAsync(dst, ctx->d_xyz.p, (size_t) cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->last_stream));
  CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  return PWPP_OK;
}
int pwpp_copy_ground_xyz(pwpp_ctx* ctx, int f, float* dst) { return copy_xyz(ctx, f, dst, true); }
int pwpp_copy_nonground_xyz(pwpp_ctx* ctx, int f, float* dst) { return copy_xyz(ctx, f, dst, false); }

int pwpp_num_patches(pwpp_ctx* ctx, int f) {
  if (check_frame(ctx, f) || fetch_counts(ctx)) return -1;
  return ctx->h_counts.p[ctx->num_streams + f];
}
int pwpp_copy_centers(pwpp_ctx* ctx, int f, float* dst) {
  int rc = check_frame(ctx, f);
  if (rc) return rc;
  rc = fetch_patches(ctx);
  if (rc) return rc;
  const int k = ctx->h_counts.p[ctx->num_streams + f];
  if (k > 0) std::memcpy(dst, ctx->h_centers.p + (size_t) f * ctx->g.nbins * 3, (size_t) k * 3 * sizeof(float));
  return PWPP_OK;
}
int pwpp_copy_normals(pwpp_ctx* ctx, int f, float* dst) {
  int rc = check_frame(ctx, f);
  if (rc) return rc;
  rc = fetch_patches(ctx);
  if (rc) return rc;
  const int k = ctx->h_counts.p[ctx->num_streams + f];
  if (k > 0) std::memcpy(dst, ctx->h_normals.p + (size_t) f * ctx->g.nbins * 3, (size_t) k * 3 * sizeof(float));
  return PWPP_OK;
}

int pwpp_get_state(pwpp_ctx* ctx, int f, pwpp_state* out) {
  if (!ctx || !out || f < 0 || f >= ctx->num_streams) return fail(PWPP_ERR_INVALID_ARG, "bad argument");
  int rc = bind_device(ctx);
  if (rc) return rc;
  if (ctx->last_stream) CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  StreamState s;
  CU_TRY(cudaMemcpy(&s, ctx->d_states.p + f, sizeof(s), cudaMemcpyDeviceToHost));
  out->sensor_height = s.sensor_height;
  for (int i = 0; i < 4; ++i) {
    out->elevation_thr[i] = s.elevation_thr[i]; out->flatness_thr[i] = s.flatness_thr[i];
    out->n_elevation[i] = s.n_elev[i]; out->n_flatness[i] = s.n_flat[i];
  }
  return PWPP_OK;
}
double pwpp_height(pwpp_ctx* ctx, int f) {
  pwpp_state s;
  if (pwpp_get_state(ctx, f, &s)) return NAN;
  return s.sensor_height;
}
double pwpp_time_us(pwpp_ctx* ctx) { return ctx ? ctx->last_time_us : NAN; }

int pwpp_call_times_us(pwpp_ctx* ctx, float out[4]) {
  if (!ctx || !out) return fail(PWPP_ERR_INVALID_ARG, "NULL argument");
  if (!ctx->call_times_valid) return fail(PWPP_ERR_INVALID_ARG, "no single-chunk pwpp_estimate_host call to report on");
  int rc = bind_device(ctx);
  if (rc) return rc;
  float ms[4];
  CU_TRY(cudaEventElapsedTime(&ms[0], ctx->ev_begin, ctx->ev0));
  CU_TRY(cudaEventElapsedTime(&ms[1], ctx->ev0, ctx->ev1));
  CU_TRY(cudaEventElapsedTime(&ms[2], ctx->ev1, ctx->ev_end));
  CU_TRY(cudaEventElapsedTime(&ms[3], ctx->ev_begin, ctx->ev_end));
  for (int i = 0; i < 4; ++i) out[i] = ms[i] * 1000.f;
  return PWPP_OK;
}

int pwpp_copy_history(pwpp_ctx* ctx, int f, int ring, int which, double* dst) {
  if (!ctx || !dst || f < 0 || f >= ctx->num_streams || ring < 0 || ring > 3 || which < 0 || which > 1) return fail(PWPP_ERR_INVALID_ARG, "bad argument");
  pwpp_state s;
  int rc = pwpp_get_state(ctx, f, &s);
  if (rc) return rc;
  const int n = which ? s.n_flatness[ring] : s.n_elevation[ring];
  if (n > 0) CU_TRY(cudaMemcpy(dst, ctx->d_hist.p + (((size_t) f * 2 + which) * 4 + ring) * ctx->hcap, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost));
  return PWPP_OK;
}

namespace {
struct StateBlobHeader { uint32_t magic, version; int32_t hcap, state_bytes; };
constexpr uint32_t STATE_BLOB_MAGIC = 0x50575354u;  // "PWST"
}  // namespace
size_t pwpp_state_blob_size(const pwpp_ctx* ctx) {
  return ctx ? sizeof(StateBlobHeader) + sizeof(StreamState) + (size_t) 2 * 4 * ctx->hcap * sizeof(double) : 0;
}
int pwpp_export_state(pwpp_ctx* ctx, int f, void* blob) {
  if (!ctx || !blob || f < 0 || f >= ctx->num_streams) return fail(PWPP_ERR_INVALID_ARG, "bad argument");
  int rc = bind_device(ctx);
  if (rc) return rc;
  if (ctx->last_stream) CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  char* b = static_cast<char*>(blob);
  const StateBlobHeader h{STATE_BLOB_MAGIC, (uint32_t) PWPP_ABI_VERSION, ctx->hcap, (int32_t) sizeof(StreamState)};
  std::memcpy(b, &h, sizeof h);
  CU_TRY(cudaMemcpy(b + sizeof h, ctx->d_states.p + f, sizeof(StreamState), cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(b + sizeof h + sizeof(StreamState), ctx->d_hist.p + (size_t) f * 2 * 4 * ctx->hcap, (size_t) 2 * 4 * ctx->hcap * sizeof(double),
                    cudaMemcpyDeviceToHost));
  return PWPP_OK;
}
int pwpp_import_state(pwpp_ctx* ctx, int f, const void* blob, size_t bytes) {
  if (!ctx || !blob || f < 0 || f >= ctx->num_streams) return fail(PWPP_ERR_INVALID_ARG, "bad argument");
  if (bytes != pwpp_state_blob_size(ctx)) return fail(PWPP_ERR_INVALID_ARG, "state blob has the wrong size for this context (different storage parameters?)");
  const char* b = static_cast<const char*>(blob);
  StateBlobHeader h;
  std::memcpy(&h, b, sizeof h);
  if (h.magic != STATE_BLOB_MAGIC || h.version != (uint32_t) PWPP_ABI_VERSION || h.hcap != ctx->hcap || h.state_bytes != (int32_t) sizeof(StreamState))
    return fail(PWPP_ERR_INVALID_ARG, "not a state blob of this library version / parameter set");
  int rc = bind_device(ctx);
  if (rc) return rc;
  // after everything already enqueued (a synchronous copy from pageable memory: the blob may be freed on return)
  if (ctx->last_stream) CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  CU_TRY(cudaMemcpy(ctx->d_states.p + f, b + sizeof h, sizeof(StreamState), cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(ctx->d_hist.p + (size_t) f * 2 * 4 * ctx->hcap, b + sizeof h + sizeof(StreamState), (size_t) 2 * 4 * ctx->hcap * sizeof(double),
                    cudaMemcpyHostToDevice));
  return PWPP_OK;
}

int pwpp_device_results(pwpp_ctx* ctx, const int32_t** d_indices, const int32_t** d_num_ground) {
  if (!ctx) return fail(PWPP_ERR_INVALID_ARG, "ctx is NULL");
  if (d_indices) *d_indices = ctx->d_out_idx.p;
  if (d_num_ground) *d_num_ground = ctx->d_counts.p;
  return PWPP_OK;
}

int pwpp_set_output_order(pwpp_ctx* ctx, int order) {
  if (!ctx) return fail(PWPP_ERR_INVALID_ARG, "ctx is NULL");
  if (order != PWPP_ORDER_BIN && order != PWPP_ORDER_REFERENCE) return fail(PWPP_ERR_INVALID_ARG, "order must be PWPP_ORDER_BIN or PWPP_ORDER_REFERENCE");
  if (order != ctx->order_mode) ++g_alloc_gen;   // captured graphs were recorded with the other launch sequence
  ctx->order_mode = order;
  return PWPP_OK;
}

int pwpp_host_results(pwpp_ctx* ctx, const int32_t** h_indices, const int32_t** h_num_ground, const int64_t** h_offsets) {
  if (!ctx) return fail(PWPP_ERR_INVALID_ARG, "ctx is NULL");
  if (ctx->last_nframes <= 0) return fail(PWPP_ERR_INVALID_ARG, "no estimate call yet");
  int rc = fetch_indices(ctx);
  if (rc) return rc;
  static_assert(sizeof(long long) == sizeof(int64_t), "offset table type");
  if (h_indices) *h_indices = ctx->h_out_idx.p;
  if (h_num_ground) *h_num_ground = ctx->h_counts.p;
  if (h_offsets) *h_offsets = reinterpret_cast<const int64_t*>(ctx->pt_off.data());
  return PWPP_OK;
}

int pwpp_bind_host_to_device(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (char* c = bus; *c; ++c) *c = (char) std::tolower((unsigned char) *c);
  std::string base = std::string("/sys/bus/pci/devices/") + bus;
  int node = -1;
  if (FILE* f = std::fopen((base + "/numa_node").c_str(), "r")) { if (std::fscanf(f, "%d", &node) != 1) node = -1; std::fclose(f); }
  if (node < 0) return -1;
  std::string list;
  if (FILE* f = std::fopen(("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist").c_str(), "r")) {
    char buf[4096] = {0};
    if (std::fgets(buf, sizeof buf, f)) list = buf;
    std::fclose(f);
  }
  if (list.empty()) return -1;
  cpu_set_t set;
  CPU_ZERO(&set);
  int ncpu = 0;
  const char* p = list.c_str();
  while (*p) {   // "0-31,64-95"
    char* end = nullptr;
    long a = std::strtol(p, &end, 10);
    if (end == p) break;
    long b = a;
    if (*end == '-') { p = end + 1; b = std::strtol(p, &end, 10); }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int) c, &set); ++ncpu; }
    p = (*end == ',') ? end + 1 : end;
    if (*end != ',' ) break;
  }
  if (ncpu == 0 || sched_setaffinity(0, sizeof set, &set) != 0) return -1;
  return node;
}

int pwpp_copy_bin_results(pwpp_ctx* ctx, int f, pwpp_bin_result* dst) {
  int rc = check_frame(ctx, f);
  if (rc) return rc;
  rc = bind_device(ctx);
  if (rc) return rc;
  CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  CU_TRY(cudaMemcpy(dst, ctx->d_fits.p + (size_t) f * ctx->g.nbins, (size_t) ctx->g.nbins * sizeof(BinFit), cudaMemcpyDeviceToHost));
  return PWPP_OK;
}
int pwpp_copy_bin_ids(pwpp_ctx* ctx, int f, uint16_t* dst) {
  int rc = check_frame(ctx, f);
  if (rc) return rc;
  rc = bind_device(ctx);
  if (rc) return rc;
  CU_TRY(cudaStreamSynchronize(ctx->last_stream));
  const int n = frame_n(ctx, f);
  if (n > 0) CU_TRY(cudaMemcpy(dst, ctx->d_bin_ids.p + ctx->pt_off[f], (size_t) n * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  return PWPP_OK;
}

#if defined(PWPP_PHASE_CLOCKS)
// diagnostic builds only: cycles thread 0 of the k_fit_group CTAs spent per phase, [3 classes][16]; reset = 1 clears them
int pwpp_debug_phase_clocks(pwpp_ctx* ctx, unsigned long long* out, int reset) {
  if (!ctx) return PWPP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (out) cudaMemcpyFromSymbol(out, g_phase_clk, sizeof(unsigned long long) * 48);
  if (reset) { static unsigned long long z[48]; cudaMemcpyToSymbol(g_phase_clk, z, sizeof z); }
  return PWPP_OK;
}
int pwpp_debug_events(pwpp_ctx* ctx, unsigned* out, int max_events) {
  if (!ctx) return -1;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (out && max_events >= 16384) cudaMemcpyFromSymbol(out, g_ev, (size_t) 16384 * 16);
  static uint4 z[16384];
  cudaMemcpyToSymbol(g_ev, z, sizeof z);
  return 16384;
}
#endif

}  // extern "C"
