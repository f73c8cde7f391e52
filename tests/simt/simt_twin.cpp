// simt_twin.cpp — TEST-ONLY: the product's CUDA kernels (csrc/pwpp_kernels.cuh, csrc/pwpp_fit.cuh) executed on the CPU
// by the fiber-based SIMT stand-in of tests/simt/cuda_runtime.h, in the launch sequence of launch_range()
// (csrc/pwpp_capi.cu). tests/test_simt_kernels.py compares the result with the oracle: unlike tests/host_twin.cu,
// which restates the kernels' algorithm sequentially, this runs the kernel code itself — warp shuffles, ballots, block
// barriers, persistent work queues and all — so a kernel change can be checked in the GPU-less build container.
// Built with plain g++ (tests/conftest.py: build_simt). Not part of the product, never loaded by it.
#include <cuda_runtime.h>   // tests/simt/cuda_runtime.h (first on the include path)

#include <algorithm>
#include <string>
#include <vector>

#include "pwpp.h"
#include "pwpp_host.hpp"
#include "pwpp_kernels.cuh"

// ---------------------------------------------------------------------------------------------------------------
// runtime of the stand-in
uint3 threadIdx;
uint3 blockIdx;
dim3 blockDim, gridDim;

namespace simt {
static Cta g_cta_seq;                 // the one CTA of sequential launches
Cta* g_cta_p = &g_cta_seq;
uint3 g_tid[MAX_THREADS];
bool g_concurrent = false;
static std::vector<Cta*> g_ctas;      // live CTAs of the current launch (1 for sequential launches)
static int g_cur_cta = 0;
static void* g_main_sp = nullptr;
static unsigned long long g_spins = 0, g_last_progress = 0, g_progress_base = 0;
static unsigned long long g_switches = 0;

asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size simt_switch,.-simt_switch
)");

[[noreturn]] void deadlock(const char* what) {
  std::fprintf(stderr, "simt: DEADLOCK in kernel %s, block (%u,%u): %s (thread %d waits; %d of %d threads of its CTA alive, %zu CTA(s) live)\n", g_cta.kernel_name,
               blockIdx.x, blockIdx.y, what, g_cta.cur, g_cta.alive, g_cta.nthreads, g_ctas.size());
  std::abort();
}

static unsigned long long total_progress() {
  unsigned long long p = 0;
  for (Cta* c : g_ctas) p += c->progress;
  return p;
}

// With a non-zero seed a yielding fiber hands over to a RANDOM live fiber of a random live CTA instead of the next one in
// round-robin order: different seeds give different interleavings of concurrently live CTAs (simt_set_sched_seed).
static unsigned long long g_sched_state = 0;
static inline unsigned sched_rand() {
  g_sched_state = g_sched_state * 6364136223846793005ull + 1442695040888963407ull;
  return (unsigned) (g_sched_state >> 33);
}
// next live fiber after (cta ci, thread ti) in round-robin order over all live CTAs; false if there is none but itself
static bool next_fiber(int ci, int ti, int& nci, int& nti) {
  const int nc = (int) g_ctas.size();
  if (g_sched_state != 0 && g_concurrent) {
    for (int tries = 0; tries < 8; ++tries) {   // a few random probes, then fall back to the round-robin scan
      const int c = (int) (sched_rand() % (unsigned) nc);
      const int t = (int) (sched_rand() % (unsigned) g_ctas[c]->nthreads);
      if (!(c == ci && t == ti) && !g_ctas[c]->fib[t].done) { nci = c; nti = t; return true; }
    }
  }
  int c = ci, t = ti;
  for (int steps = 0; steps < nc * MAX_THREADS + MAX_THREADS; ++steps) {
    ++t;
    if (t >= g_ctas[c]->nthreads) { t = 0; c = (c + 1) % nc; }
    if (c == ci && t == ti) return false;
    if (!g_ctas[c]->fib[t].done) { nci = c; nti = t; return true; }
  }
  return false;
}

static void switch_from(Fiber& me, int nci, int nti) {
  Cta* n = g_ctas[nci];
  g_cur_cta = nci;
  g_cta_p = n;
  n->cur = nti;
  threadIdx = g_tid[nti];
  blockIdx = n->bid;
  ++g_switches;
  simt_switch(&me.sp, n->fib[nti].sp);
}

void yield_() {
  Cta& c = g_cta;
  const unsigned long long p = total_progress();
  if (p != g_last_progress) { g_last_progress = p; g_spins = 0; }
  else if (++g_spins > 64ull * (unsigned long long) c.nthreads * g_ctas.size() + 4096ull) deadlock("no thread can make progress");
  int nci, nti;
  if (!next_fiber(g_cur_cta, c.cur, nci, nti)) return;
  switch_from(c.fib[c.cur], nci, nti);
}

static void fiber_main() {
  {
    Cta& c = g_cta;
    c.body();
  }
  // the thread returned from the kernel
  Cta& c = g_cta;
  const int me = c.cur, myc = g_cur_cta;
  c.fib[me].done = true;
  --c.alive;
  ++c.progress;
  for (;;) {
    int nci, nti;
    if (next_fiber(myc, me, nci, nti)) switch_from(c.fib[me], nci, nti);
    else { threadIdx = g_tid[0]; simt_switch(&c.fib[me].sp, g_main_sp); }
  }
}

static void init_cta(Cta& c, const char* name, int nt, size_t smem, const std::function<void()>& body, uint3 bid) {
  c.kernel_name = name;
  c.body = body;
  c.dyn_smem.assign(smem + 64, 0);
  c.nthreads = nt; c.alive = nt; c.bar_arrived = 0; c.bid = bid; c.cur = 0;
  for (int w = 0; w < (nt + 31) / 32; ++w) { c.warps[w].nslots = 0; for (auto& s : c.warps[w].slots) { s.mask = 0; s.arrived = 0; s.departed = 0; s.draining = false; } }
  for (int t = 0; t < nt; ++t) {
    Fiber& f = c.fib[t];
    if (!f.stack) f.stack = (char*) std::malloc(STACK_BYTES);
    f.done = false;
    uintptr_t top = ((uintptr_t) (f.stack + STACK_BYTES)) & ~(uintptr_t) 15;
    void** sp = (void**) top;
    *--sp = nullptr;                    // fake return address of fiber_main (keeps the ABI stack alignment)
    *--sp = (void*) &fiber_main;        // popped by `ret` in simt_switch
    for (int k = 0; k < 6; ++k) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
    f.sp = sp;
  }
}

static void run_live_ctas() {
  g_spins = 0;
  g_last_progress = total_progress();
  g_cur_cta = 0;
  g_cta_p = g_ctas[0];
  g_cta_p->cur = 0;
  threadIdx = g_tid[0];
  blockIdx = g_cta_p->bid;
  simt_switch(&g_main_sp, g_cta_p->fib[0].sp);
  for (Cta* c : g_ctas) if (c->alive != 0) { g_cta_p = c; deadlock("returned to the launcher with live threads"); }
}

static void set_tids(dim3 block, int nt) {
  for (int t = 0; t < nt; ++t) { g_tid[t].x = t % block.x; g_tid[t].y = (t / block.x) % block.y; g_tid[t].z = t / (block.x * block.y); }
}

void launch_impl(const char* name, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nt = (int) (block.x * block.y * block.z);
  if (nt <= 0 || nt > MAX_THREADS) { std::fprintf(stderr, "simt: bad block size %d\n", nt); std::abort(); }
  blockDim = block; gridDim = grid;
  set_tids(block, nt);
  g_concurrent = false;
  g_ctas.assign(1, &g_cta_seq);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        uint3 bid; bid.x = bx; bid.y = by; bid.z = bz;
        init_cta(g_cta_seq, name, nt, smem, body, bid);
        run_live_ctas();
      }
  g_cta_p = &g_cta_seq;
}

// All CTAs of a 1-D grid live at once (a kernel whose CTAs wait for each other). The kernel must not use static
// __shared__ variables (they would be shared by all CTAs here).
// cluster barrier: every thread of every live CTA (threads that returned no longer count)
static int g_cl_arrived = 0;
static unsigned long long g_cl_gen = 0;
void cluster_sync_() {
  auto alive = [] { int a = 0; for (Cta* c : g_ctas) a += c->alive; return a; };
  ++g_cta.progress;
  const unsigned long long gen = g_cl_gen;
  if (++g_cl_arrived >= alive()) { g_cl_arrived = 0; ++g_cl_gen; return; }
  while (g_cl_gen == gen) {
    yield_();
    if (g_cl_gen == gen && g_cl_arrived >= alive()) { g_cl_arrived = 0; ++g_cl_gen; }
  }
}
void* cluster_map_(void* p, int rank) {
  char* base = g_cta.dyn_smem.data();
  const size_t off = (size_t) (static_cast<char*>(p) - base);
  if (off >= g_cta.dyn_smem.size() || rank < 0 || rank >= (int) g_ctas.size()) { std::fprintf(stderr, "simt: cluster_map of a pointer outside dynamic shared memory\n"); std::abort(); }
  return g_ctas[rank]->dyn_smem.data() + off;
}

void launch_concurrent_impl(const char* name, int nctas, dim3 block, size_t smem, const std::function<void()>& body, unsigned block_y) {
  const int nt = (int) (block.x * block.y * block.z);
  if (nt <= 0 || nt > MAX_THREADS || nctas < 1 || nctas > 16) { std::fprintf(stderr, "simt: bad concurrent launch (%d CTAs x %d threads)\n", nctas, nt); std::abort(); }
  blockDim = block; gridDim = dim3(nctas, 1, 1);
  set_tids(block, nt);
  static std::vector<Cta*> pool;
  while ((int) pool.size() < nctas) pool.push_back(new Cta());
  g_ctas.assign(pool.begin(), pool.begin() + nctas);
  for (int c = 0; c < nctas; ++c) { uint3 bid; bid.x = c; bid.y = block_y; bid.z = 0; init_cta(*g_ctas[c], name, nt, smem, body, bid); }
  g_cl_arrived = 0;
  g_concurrent = true;
  run_live_ctas();
  g_concurrent = false;
  g_ctas.assign(1, &g_cta_seq);
  g_cta_p = &g_cta_seq;
}
}  // namespace simt

// ---------------------------------------------------------------------------------------------------------------
using namespace pwpp;

namespace {
struct FrameOut {
  std::vector<int> ground, nonground;
  std::vector<uint16_t> bin_ids;
  std::vector<BinFit> fits;
  std::vector<float> centers, normals;
  int npatch = 0;
};

struct SimtTwin {
  pwpp_params prm;
  Geometry g;
  AlgoParams ap;
  bool fast = true;
  int nbp = 0, hcap = 0, max_sectors = 0, num_streams = 1;
  std::vector<StreamState> states;
  std::vector<double> hist;
  std::vector<FrameOut> out;
  int sel = 0;
  // switches (the PWPP_* environment switches of pwpp_create)
  int persistent_ctas = 2, front = 1, patch = 0, order = 0;
  std::string last_launches;
};

template <typename T>
T* ptr(std::vector<T>& v) { return v.empty() ? nullptr : v.data(); }
}  // namespace

extern "C" {

void* simt_create(const pwpp_params* p, int num_streams) {
  SimtTwin* t = new SimtTwin();
  t->prm = *p;
  t->num_streams = num_streams < 1 ? 1 : num_streams;
  build_geometry(*p, t->g, t->ap, t->fast);
  t->nbp = ((t->g.nbins + PW_NUM_PSEUDO + 31) / 32) * 32;
  for (int k = 0; k < 4; ++k) t->max_sectors = std::max(t->max_sectors, t->g.num_sectors[k]);
  t->hcap = std::max(p->max_elevation_storage, p->max_flatness_storage) + 4 * t->max_sectors + 64;
  t->states.resize(t->num_streams);
  for (auto& s : t->states) init_state(*p, s);
  t->hist.assign((size_t) t->num_streams * 2 * 4 * t->hcap, 0.0);
  t->out.resize(t->num_streams);
  return t;
}
void simt_destroy(void* h) { delete (SimtTwin*) h; }
int simt_num_bins(void* h) { return ((SimtTwin*) h)->g.nbins; }
void simt_select(void* h, int f) { ((SimtTwin*) h)->sel = f; }
int simt_set_option(void* h, const char* name, int v) {
  SimtTwin* t = (SimtTwin*) h;
  const std::string n(name);
  if (n == "persistent_ctas") t->persistent_ctas = v;
  else if (n == "front") t->front = v;
  else if (n == "patch") t->patch = v;
  else if (n == "order") t->order = v;
  else return -1;
  return 0;
}
unsigned long long simt_fiber_switches(void) { return simt::g_switches; }
void simt_set_sched_seed(unsigned long long seed) { simt::g_sched_state = seed; }

// One frame for each of the first nframes streams: the launch sequence of launch_range() (csrc/pwpp_capi.cu).
void simt_estimate_multi(void* h, int nframes, const float* const* pts_in, const int64_t* ns, int cols) {
  SimtTwin* t = (SimtTwin*) h;
  const Geometry g = t->g;
  const AlgoParams ap = t->ap;
  const int nb = g.nbins, nbp = t->nbp, nb_all = nb + PW_NUM_PSEUDO;
  std::vector<long long> pt_off(nframes + 1, 0);
  std::vector<int> chunk_off(nframes + 1, 0);
  int max_chunks = 0;
  for (int f = 0; f < nframes; ++f) {
    pt_off[f + 1] = pt_off[f] + ns[f];
    const int nc = (int) ((ns[f] + CHUNK_PTS - 1) / CHUNK_PTS);
    chunk_off[f + 1] = chunk_off[f] + nc;
    max_chunks = std::max(max_chunks, nc);
  }
  const long long total = pt_off[nframes];
  const int total_chunks = chunk_off[nframes];
  std::vector<float4> pts((size_t) total + 1);
  for (int f = 0; f < nframes; ++f)
    for (int64_t i = 0; i < ns[f]; ++i) {
      const float* p = pts_in[f] + i * cols;
      pts[(size_t) (pt_off[f] + i)] = make_float4(p[0], p[1], p[2], cols >= 4 ? p[3] : 0.f);
    }
  const int has_intensity = cols >= 4;
  std::vector<unsigned short> bin_ids((size_t) total + 1), chist((size_t) total_chunks * nbp + 1);
  std::vector<unsigned int> cbase((size_t) total_chunks * nbp + 1);
  std::vector<int> bin_off((size_t) nframes * (nbp + 1));
  std::vector<float4> sorted((size_t) total + 1);
  std::vector<int> part((size_t) total + 1, -7), out_idx((size_t) total + 1, -7);
  std::vector<BinFit> fits((size_t) nframes * nb);
  std::vector<BinSeg> segs((size_t) nframes * nb_all);
  std::vector<int4> items[NUM_CLASSES];
  for (auto& v : items) v.resize((size_t) nframes * nb + 1);
  std::vector<int> ctr(2 * NUM_CLASSES + ORD_NUM_HEADS, 0);
  std::vector<unsigned char> labels((size_t) total + 1, 0);
  std::vector<int> counts((size_t) 3 * nframes, 0);
  std::vector<float> centers((size_t) nframes * nb * 3), normals((size_t) nframes * nb * 3);

  FrameTable ft{pt_off.data(), chunk_off.data()};
  StreamState* states = t->states.data();
  const float4* d_pts = pts.data();
  WorkQueues wq;
  for (int c = 0; c < NUM_CLASSES; ++c) wq.items[c] = items[c].data();
  wq.count = ctr.data();
  wq.head = ctr.data() + NUM_CLASSES;
  wq.labels = t->order ? labels.data() : nullptr;
  if (t->front) {   // one cluster per frame (pwpp_front.cuh): the CTAs of a cluster run concurrently, frames one after another
    // front=1: 8 x 256 threads (KITTI-sized frames), front=3: 8 x 512 (what dense frames get)
    const int nt = t->front == 3 ? FC_THREADS_DENSE : FC_THREADS;
    const size_t sm_f = front_cluster_smem_bytes(nbp, nt);
    for (int f = 0; f < nframes; ++f) {
#define FC_ARGS d_pts, ft, states, g, ap, has_intensity, nbp, nb, bin_ids.data(), bin_off.data(), wq, fits.data(), sorted.data()
      if (nt == FC_THREADS_DENSE) {
        if (t->fast) simt::launch_concurrent("k_front_cluster<fast,512>", FC_CS, nt, sm_f, [&] { k_front_cluster<true, CLS_L2_MAX, FC_THREADS_DENSE>(FC_ARGS); }, (unsigned) f);
        else simt::launch_concurrent("k_front_cluster<exact,512>", FC_CS, nt, sm_f, [&] { k_front_cluster<false, CLS_L2_MAX, FC_THREADS_DENSE>(FC_ARGS); }, (unsigned) f);
      } else {
        if (t->fast) simt::launch_concurrent("k_front_cluster<fast>", FC_CS, nt, sm_f, [&] { k_front_cluster<true, CLS_L2_MAX, FC_THREADS>(FC_ARGS); }, (unsigned) f);
        else simt::launch_concurrent("k_front_cluster<exact>", FC_CS, nt, sm_f, [&] { k_front_cluster<false, CLS_L2_MAX, FC_THREADS>(FC_ARGS); }, (unsigned) f);
      }
#undef FC_ARGS
    }
  } else {
    if (max_chunks > 0) {
      dim3 grid(max_chunks, nframes);
      const size_t sm_h = nbp * sizeof(unsigned int);
#define HIST_ARGS d_pts, ft, states, g, ap, has_intensity, nbp, bin_ids.data(), chist.data()
      if (!t->fast) simt::launch("k_bin_hist<false,0>", grid, CHUNK_THREADS, sm_h, [&] { k_bin_hist<false, 0>(HIST_ARGS); });
      else simt::launch("k_bin_hist<true,2>", grid, CHUNK_THREADS, sm_h, [&] { k_bin_hist<true, 2>(HIST_ARGS); });
#undef HIST_ARGS
    }
    simt::launch("k_bin_scan", nframes, 512, (nbp + 1) * sizeof(int),
                 [&] { k_bin_scan<CLS_L2_MAX>(ft, nbp, nb, ap.num_min_pts, chist.data(), cbase.data(), bin_off.data(), wq, fits.data()); });
    if (max_chunks > 0) {
      dim3 grid(max_chunks, nframes);
      const size_t sm_sc = (size_t) (CHUNK_THREADS / 32) * nbp * sizeof(unsigned int);
      simt::launch("k_scatter<false,4>", grid, CHUNK_THREADS, sm_sc, [&] { k_scatter<false, 4>(d_pts, ft, nbp, bin_ids.data(), cbase.data(), sorted.data()); });
    }
  }
#define FIT_ARGS sorted.data(), ft, states, g, ap, nbp, bin_off.data(), wq, part.data(), fits.data()
  const int pg = t->persistent_ctas;
  const size_t sm_m = FITW_WARPS * CLS_M_MAX * sizeof(float4), sm_l2 = 3 * 4096 * sizeof(float), sm_l3 = 3 * 8192 * sizeof(float);
  if (t->patch) {   // PWPP_FIT_PATCH: patches above 512 points on k_fit_patch
    simt::launch("k_fit_patch<16>", pg, 16 * 32, (size_t) 16 * FP_STG * 16, [&] { k_fit_patch<16, 1, 4>(FIT_ARGS); });
    simt::launch("k_fit_patch<8>", pg, 8 * 32, (size_t) 8 * FP_STG * 16, [&] { k_fit_patch<8, 2, 3>(FIT_ARGS); });
    simt::launch("k_fit_patch<4>", pg, 4 * 32, (size_t) 4 * FP_STG * 16, [&] { k_fit_patch<4, 4, 2>(FIT_ARGS); });
  } else {
    simt::launch("k_fit_cta<8192,4,2,8,fuse>", pg, FIT_THREADS, sm_l3, [&] { k_fit_cta<8192, 4, 2, 8, true>(FIT_ARGS); });
    simt::launch("k_fit_cta<4096,3,3,8,fuse,pls>", pg, FIT_THREADS, sm_l2, [&] { k_fit_cta<4096, 3, 3, 8, true, true>(FIT_ARGS); });
    simt::launch("k_fit_warp<false,2,2,pls>", pg, FITW_WARPS * 32, 0, [&] { k_fit_warp<false, 2, 2, 2, 3, false, true>(FIT_ARGS); });
  }
  simt::launch("k_fit_warp<true,1,1,pls>", pg, FITW_WARPS * 32, sm_m, [&] { k_fit_warp<true, 1, 1, 2, 3, false, true>(FIT_ARGS); });
  simt::launch("k_fit_resident<8,8,0>", pg, FIT_THREADS, 0, [&] { k_fit_resident<8, 8, 0, 2>(FIT_ARGS); });
  simt::launch("k_fit_big<16,1,fuse>", pg, 512, 0, [&] { k_fit_big<16, 1, true>(FIT_ARGS); });
#undef FIT_ARGS
  for (int c = 0; c < NUM_CLASSES; ++c)
    if (ctr[NUM_CLASSES + c] < ctr[c]) { std::fprintf(stderr, "simt_twin: class %d queue not drained (%d of %d)\n", c, ctr[NUM_CLASSES + c], ctr[c]); std::abort(); }
  {
    char buf[256];
    std::snprintf(buf, sizeof buf, "S=%d M=%d L1=%d L2=%d L3=%d X=%d", ctr[0], ctr[1], ctr[2], ctr[3], ctr[4], ctr[5]);
    t->last_launches = buf;
  }
  if (t->order) {
    int* heads = ctr.data() + 2 * NUM_CLASSES;
    simt::launch("k_order_cta<X>", pg, 512, ord_cta_smem_bytes(512), [&] { k_order_cta<512, 5>(sorted.data(), wq, heads + 0, part.data()); });
    simt::launch("k_order_cta<L3>", pg, 512, ord_cta_smem_bytes(512), [&] { k_order_cta<512, 4>(sorted.data(), wq, heads + 1, part.data()); });
    simt::launch("k_order_cta<L2>", pg, 256, ord_cta_smem_bytes(256), [&] { k_order_cta<256, 3>(sorted.data(), wq, heads + 2, part.data()); });
    simt::launch("k_order_cta<L1>", pg, 128, ord_cta_smem_bytes(128), [&] { k_order_cta<128, 2>(sorted.data(), wq, heads + 3, part.data()); });
    simt::launch("k_order_warp", pg, ORD_WARP_THREADS, 0, [&] { k_order_warp(sorted.data(), wq, heads + 4, part.data()); });
  }
  int* d_ng = counts.data();
  int* d_np = counts.data() + nframes;
  int* d_nd = counts.data() + 2 * nframes;
  {
    const size_t gle_smem = gle_smem_bytes(t->max_sectors);
    simt::launch("k_gle", nframes, 32, gle_smem, [&] {
      k_gle(ft, states, t->hist.data(), t->hcap, g, ap, nbp, t->max_sectors, bin_off.data(), fits.data(), segs.data(), d_ng, d_np, centers.data(), normals.data(), d_nd);
    });
  }
  if (max_chunks > 0) {
    const long long max_pts = (long long) max_chunks * CHUNK_PTS;
    dim3 grid((unsigned) ((max_pts + (long long) EMIT_TILE * EMIT_WARPS - 1) / ((long long) EMIT_TILE * EMIT_WARPS)), nframes);
    simt::launch("k_emit", grid, EMIT_WARPS * 32, 0, [&] { k_emit(ft, g, nbp, bin_off.data(), fits.data(), segs.data(), part.data(), sorted.data(), out_idx.data()); });
  }
  for (int f = 0; f < nframes; ++f) {
    FrameOut& o = t->out[f];
    const long long p0 = pt_off[f];
    const int n = (int) ns[f], ng = d_ng[f], nd = d_nd[f];
    o.ground.assign(out_idx.begin() + p0, out_idx.begin() + p0 + ng);
    o.nonground.assign(out_idx.begin() + p0 + ng, out_idx.begin() + p0 + (n - nd));
    o.bin_ids.assign(bin_ids.begin() + p0, bin_ids.begin() + p0 + n);
    o.fits.assign(fits.begin() + (size_t) f * nb, fits.begin() + (size_t) (f + 1) * nb);
    o.npatch = d_np[f];
    o.centers.assign(centers.begin() + (size_t) f * nb * 3, centers.begin() + (size_t) f * nb * 3 + (size_t) o.npatch * 3);
    o.normals.assign(normals.begin() + (size_t) f * nb * 3, normals.begin() + (size_t) f * nb * 3 + (size_t) o.npatch * 3);
  }
}

void simt_estimate(void* h, const float* pts, int64_t n, int cols) {
  const float* p[1] = {pts};
  const int64_t ns[1] = {n};
  simt_estimate_multi(h, 1, p, ns, cols);
  ((SimtTwin*) h)->sel = 0;
}

static FrameOut& cur(void* h) { SimtTwin* t = (SimtTwin*) h; return t->out[t->sel]; }
int64_t simt_num_ground(void* h) { return (int64_t) cur(h).ground.size(); }
int64_t simt_num_nonground(void* h) { return (int64_t) cur(h).nonground.size(); }
void simt_ground_indices(void* h, int32_t* dst) { std::memcpy(dst, cur(h).ground.data(), cur(h).ground.size() * 4); }
void simt_nonground_indices(void* h, int32_t* dst) { std::memcpy(dst, cur(h).nonground.data(), cur(h).nonground.size() * 4); }
int simt_num_patches(void* h) { return cur(h).npatch; }
void simt_centers(void* h, float* dst) { std::memcpy(dst, cur(h).centers.data(), cur(h).centers.size() * 4); }
void simt_normals(void* h, float* dst) { std::memcpy(dst, cur(h).normals.data(), cur(h).normals.size() * 4); }
void simt_bin_ids(void* h, uint16_t* dst) { std::memcpy(dst, cur(h).bin_ids.data(), cur(h).bin_ids.size() * 2); }
void simt_bin_results(void* h, pwpp_bin_result* dst) { std::memcpy(dst, cur(h).fits.data(), cur(h).fits.size() * sizeof(BinFit)); }
const char* simt_queue_sizes(void* h) { return ((SimtTwin*) h)->last_launches.c_str(); }
void simt_get_state(void* h, pwpp_state* out) {
  SimtTwin* t = (SimtTwin*) h;
  const StreamState& st = t->states[t->sel];
  out->sensor_height = st.sensor_height;
  for (int i = 0; i < 4; ++i) {
    out->elevation_thr[i] = st.elevation_thr[i]; out->flatness_thr[i] = st.flatness_thr[i];
    out->n_elevation[i] = st.n_elev[i]; out->n_flatness[i] = st.n_flat[i];
  }
}
void simt_history(void* h, int ring, int which, double* dst) {
  SimtTwin* t = (SimtTwin*) h;
  const StreamState& st = t->states[t->sel];
  const int n = which ? st.n_flat[ring] : st.n_elev[ring];
  std::memcpy(dst, t->hist.data() + (((size_t) t->sel * 2 + which) * 4 + ring) * t->hcap, (size_t) n * 8);
}
}  // extern "C"
