// hostsim.cpp -- single-lane HOST emulation of the CUDA engine's per-env code (LANES = 1).
//
// TEST INFRASTRUCTURE ONLY.  It compiles the very same engine_core.h / engine_env.h the CUDA
// library is built from, with COSIM_HOST_EMU, so that kernel *logic* (layouts, indexing, solver
// control flow) can be checked against the fp64 oracle on machines without a GPU
// (`pytest -m "not gpu"`).  It is never loaded by the cosim_b200 package: the product path is
// libcosim_b200.so (engine.cu) and fails loudly without a GPU.
#define COSIM_HOST_EMU 1
#define COSIM_GENERAL 1      // the emulation carries the general constraint path too (tests/test_cones.py)
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include "../../cosim_b200/csrc/engine_env.h"
#include "../../cosim_b200/csrc/engine_setup.h"

// float inverse normal CDF for the emulation (Acklam's rational approximation + one Halley step)
static inline float fast_ndtri(float pf) {
  double p = pf;
  if (p <= 0) return -INFINITY;
  if (p >= 1) return INFINITY;
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02, 1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02, 6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00, -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
  double x;
  if (p < 0.02425) { double q = std::sqrt(-2 * std::log(p)); x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1); }
  else if (p > 1 - 0.02425) { double q = std::sqrt(-2 * std::log(1 - p)); x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1); }
  else { double q = p - 0.5, r = q * q; x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q / (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1); }
  double e = 0.5 * std::erfc(-x / std::sqrt(2.0)) - p, u = e * std::sqrt(2 * M_PI) * std::exp(x * x / 2);
  x = x - u / (1 + x * u / 2);
  return (float)x;
}

struct HostSim {
  ModelDev m; EnvArrays E; int N;
  std::vector<void*> allocs; std::vector<float> ws, overflow;
  ~HostSim() { for (void* p : allocs) free(p); }
};
static void* hs_up(void* ctx, const void* src, size_t bytes) { void* p = malloc(bytes); memcpy(p, src, bytes); ((HostSim*)ctx)->allocs.push_back(p); return p; }
static void* hs_za(void* ctx, size_t bytes) { void* p = calloc(bytes ? bytes : 4, 1); ((HostSim*)ctx)->allocs.push_back(p); return p; }

extern "C" {
void* hs_create(const void* blob, uint64_t nbytes, int num_envs, uint64_t seed, uint32_t env_offset) {
  HostSim* h = new HostSim;
  try {
    Uploader u = {hs_up, h};
    setup::build_model(blob, nbytes, seed, env_offset, u, h->m);
  } catch (std::exception& e) { fprintf(stderr, "hostsim: %s\n", e.what()); delete h; return nullptr; }
  h->N = num_envs;
  setup::alloc_env(h->m, num_envs, h->E, hs_za, h);
  setup::alloc_debug(h->m, num_envs, h->E, hs_za, h);
  h->ws.assign(h->m.ws_floats, 0.f);
  h->overflow.assign((size_t)h->m.gslot_floats + 4, 0.f);          // contact records beyond the shared-memory tier (the GPU's global slot)
  *(float**)(h->ws.data() + h->m.off[W_GPTR]) = h->overflow.data();
  for (int e = 0; e < num_envs; ++e) init_env(h->m, h->E, e, h->ws.data(), 0);
  return h;
}
void hs_destroy(void* hp) { delete (HostSim*)hp; }
// workspace layout (float offsets of every field, then ws_floats, shared_floats, arena_bytes): tools/ws_layout.py
int hs_layout(void* hp, int* out, int cap) { HostSim* h = (HostSim*)hp; int n = 0; for (int i = 0; i < W__COUNT && n < cap; ++i) out[n++] = h->m.off[i]; if (n + 3 <= cap) { out[n++] = h->m.ws_floats; out[n++] = h->m.shared_floats; out[n++] = h->m.arena_bytes; } return n; }
int hs_ws_floats(void* hp) { return ((HostSim*)hp)->m.ws_floats; }
void hs_reset(void* hp, const uint8_t* mask, const float* command, float* state) {
  HostSim* h = (HostSim*)hp; const int cd = h->m.dims[CD_command_dim], sd = h->m.dims[CD_state_dim];
  for (int e = 0; e < h->N; ++e) {
    if (mask && !mask[e]) continue;
    reset_env(h->m, h->E, e, h->ws.data(), command ? command + (size_t)e * cd : nullptr, state + (size_t)e * sd, 0);
  }
}
void hs_step(void* hp, const float* action, const float* command, float* state, uint8_t* term, uint8_t* trunc) {
  HostSim* h = (HostSim*)hp;
  StepArgs a = {action, command, nullptr, state, term, trunc, nullptr};
  for (int e = 0; e < h->N; ++e) step_env(h->m, h->E, e, h->ws.data(), a, 0);
}
void hs_substep(void* hp) {
  HostSim* h = (HostSim*)hp;
  for (int e = 0; e < h->N; ++e) substep_env(h->m, h->E, e, h->ws.data(), 0);
}
void hs_push(void* hp, const uint8_t* mask, const float* vel) {
  HostSim* h = (HostSim*)hp;
  for (int e = 0; e < h->N; ++e) if (!mask || mask[e]) push_env(h->m, h->E, e, vel + 3 * (size_t)e, 0);
}
int hs_field_dim(void* hp, const char* name) {
  HostSim* h = (HostSim*)hp;
  for (auto& f : setup::env_fields(h->m, h->E)) if (!strcmp(f.name, name)) return f.dim;
  return -1;
}
int hs_get(void* hp, const char* name, void* dst) {
  HostSim* h = (HostSim*)hp;
  for (auto& f : setup::env_fields(h->m, h->E)) if (!strcmp(f.name, name)) { memcpy(dst, f.ptr, (size_t)h->N * f.dim * 4); return f.is_int; }
  return -1;
}
int hs_set(void* hp, const char* name, const void* src) {
  HostSim* h = (HostSim*)hp;
  for (auto& f : setup::env_fields(h->m, h->E)) if (!strcmp(f.name, name)) { memcpy(f.ptr, src, (size_t)h->N * f.dim * 4); return 0; }
  return -1;
}
// unit access to the engine's Cholesky (engine_core.h chol_factor / chol_solve): A [nv x nv] row-major SPD, b [nv] -> x [nv]
int hs_chol(void* hp, const float* A_in, const float* b, int sparse, float* x) {
  HostSim* h = (HostSim*)hp; const ModelDev& m = h->m; float* ws = h->ws.data(); const int nv = MD(nv);
  for (int i = 0; i < nv * nv; ++i) WS(W_A)[i] = A_in[i];
  for (int i = 0; i < nv; ++i) WS(W_TMPV)[i] = b[i];
  chol_factor(m, WS(W_A), WS(W_INVD), nv, 0, sparse);
  chol_solve(WS(W_A), WS(W_INVD), WS(W_TMPV), WS(W_BUF), WS(W_GRAD), nv, 0);
  for (int i = 0; i < nv; ++i) x[i] = WS(W_GRAD)[i];
  return nv;
}
long hs_ls_evals() { return g_emu_ls_evals; }
uint32_t hs_philox(void* hp, uint32_t env, uint32_t stream, uint32_t step, uint32_t idx) { return philox_draw(((HostSim*)hp)->m, env, stream, step, idx); }
}
