/* TEST-ONLY stand-in for libpwpp_b200.so: just enough of include/pwpp.h for examples/pwpp_sequence.cpp to run its
 * reader / consumer pipeline in the GPU-less build container (tests/test_examples.py). "Ground" = points with z < -1.5. */
#include <stdlib.h>
#include <string.h>
#include "pwpp.h"
struct pwpp_ctx { int64_t n; int32_t* g; int32_t* ng; int64_t ngr, nng; double h; };
void pwpp_params_default(pwpp_params* p) { memset(p, 0, sizeof *p); p->num_zones = 4; }
int pwpp_create(const pwpp_params* p, int device, int ns, int64_t m, pwpp_ctx** out) { (void) p; (void) device; (void) ns; (void) m; *out = calloc(1, sizeof(pwpp_ctx)); (*out)->h = 1.723; return 0; }
void pwpp_destroy(pwpp_ctx* c) { if (c) { free(c->g); free(c->ng); free(c); } }
const char* pwpp_last_error(void) { return "stub"; }
void* pwpp_host_alloc(size_t b) { return malloc(b ? b : 1); }
void pwpp_host_free(void* p) { free(p); }
int pwpp_estimate_host(pwpp_ctx* c, int nf, const float* const* pts, const int64_t* n, int cols, int64_t rs, int64_t cs) {
  (void) nf; (void) cols;
  free(c->g); free(c->ng);
  c->g = malloc(sizeof(int32_t) * (size_t) (n[0] + 1)); c->ng = malloc(sizeof(int32_t) * (size_t) (n[0] + 1));
  c->ngr = c->nng = 0;
  for (int64_t i = 0; i < n[0]; ++i) { if (pts[0][i * rs + 2 * cs] < -1.5f) c->g[c->ngr++] = (int32_t) i; else c->ng[c->nng++] = (int32_t) i; }
  c->n = n[0];
  return 0;
}
int64_t pwpp_num_ground(pwpp_ctx* c, int f) { (void) f; return c->ngr; }
int64_t pwpp_num_nonground(pwpp_ctx* c, int f) { (void) f; return c->nng; }
int pwpp_copy_ground_indices(pwpp_ctx* c, int f, int32_t* d) { (void) f; memcpy(d, c->g, sizeof(int32_t) * (size_t) c->ngr); return 0; }
int pwpp_copy_nonground_indices(pwpp_ctx* c, int f, int32_t* d) { (void) f; memcpy(d, c->ng, sizeof(int32_t) * (size_t) c->nng); return 0; }
int pwpp_copy_ground_xyz(pwpp_ctx* c, int f, float* d) { (void) c; (void) f; (void) d; return 0; }
int pwpp_copy_nonground_xyz(pwpp_ctx* c, int f, float* d) { (void) c; (void) f; (void) d; return 0; }
int pwpp_num_patches(pwpp_ctx* c, int f) { (void) c; (void) f; return 2; }
int pwpp_copy_centers(pwpp_ctx* c, int f, float* d) { (void) c; (void) f; memset(d, 0, 24); return 0; }
int pwpp_copy_normals(pwpp_ctx* c, int f, float* d) { (void) c; (void) f; memset(d, 0, 24); return 0; }
double pwpp_height(pwpp_ctx* c, int f) { (void) f; return c->h; }
double pwpp_time_us(pwpp_ctx* c) { (void) c; return 1000.0; }
int pwpp_set_output_order(pwpp_ctx* c, int order) { (void) c; (void) order; return 0; }
int pwpp_device_synchronize(pwpp_ctx* c) { (void) c; return 0; }
int pwpp_estimate_device(pwpp_ctx* c, int nf, const void* d, const int64_t* o, int hi, void* s) { (void) c; (void) nf; (void) d; (void) o; (void) hi; (void) s; return -1; }
int pwpp_estimate_device_xyz(pwpp_ctx* c, int nf, const void* d, const int64_t* o, void* s) { (void) c; (void) nf; (void) d; (void) o; (void) s; return -1; }
int pwpp_device_results(pwpp_ctx* c, const int32_t** a, const int32_t** b) { (void) c; if (a) *a = 0; if (b) *b = 0; return 0; }
