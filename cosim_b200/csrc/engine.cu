// engine.cu -- kernels + C ABI of libcosim_b200.so (sm_100a).  See include/cosim_b200.h for the
// reference interface each entry point replaces.  One warp owns one environment; its working set
// lives in that warp's slice of dynamic shared memory (engine_core.h).  No CPU path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "engine_kernels.cuh"
#include "engine_setup.h"
#include "../../include/cosim_b200.h"

// ------------------------------------------------------------------------------------------ kernels
// per-env kernels: engine_kernels.cuh, instantiated here without and in engine_gen.cu with the general constraint path
KernelSet kernel_set_gen();
KernelSet kernel_set_fast24();
KernelSet kernel_set_fast12();
__global__ void k_push(const __grid_constant__ ModelDev m, const EnvArrays E, const uint8_t* mask, const float* vel) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= E.N) return;
  if (mask && !mask[env]) return;
  push_env(m, E, env, vel + 3 * (size_t)env, 0);
}
// counter-based RNG probe: out[env][i] = i-th 32-bit draw of (env, stream, step)
__global__ void k_rng_probe(const __grid_constant__ ModelDev m, int N, uint32_t stream, uint32_t step, int nidx, uint32_t* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * nidx) return;
  const int env = t / nidx, i = t - env * nidx;
  out[t] = philox_draw(m, (uint32_t)env, stream, step, (uint32_t)i);
}
// reporter statistics: out[k] = sum over envs (max for ST_MAX_TORQUE), fp64 accumulation
__global__ void k_stats(const float* stats, int N, double* out) {
  __shared__ double sh[8][ST__COUNT];
  const int k = threadIdx.x & (ST__COUNT - 1), part = threadIdx.x / ST__COUNT, nparts = blockDim.x / ST__COUNT;
  double acc = 0.0;
  for (int e = blockIdx.x * nparts + part; e < N; e += gridDim.x * nparts) {
    const double v = (double)stats[(size_t)e * ST__COUNT + k];
    if (k == ST_MAX_TORQUE) acc = v > acc ? v : acc; else acc += v;
  }
  sh[part][k] = acc;
  __syncthreads();
  if (part == 0) {
    for (int p = 1; p < nparts; ++p) { if (k == ST_MAX_TORQUE) acc = sh[p][k] > acc ? sh[p][k] : acc; else acc += sh[p][k]; }
    if (k == ST_MAX_TORQUE) {
      unsigned long long* addr = (unsigned long long*)(out + k); unsigned long long old = *addr, assumed;
      do { assumed = old; if (__longlong_as_double((long long)assumed) >= acc) break; old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(acc)); } while (assumed != old);
    } else atomicAdd(out + k, acc);
  }
}

// ------------------------------------------------------------------------------------------ handle
struct cosim_handle {
  ModelDev m; EnvArrays E; EnvArrays Edbg;   // Edbg keeps the debug pointers while dumps are switched off
  int N = 0, device = 0, wpb = 1, grid = 1, launches = 0, debug = 0;
  size_t smem = 0;
  int* sched = nullptr;                     // chunk counter of k_step (zeroed on the stream before every launch)
  PoolArgs pool = {nullptr, 0, 0, 0, 0, 0, nullptr};    // pooled step kernel (P = 0: off)
  KernelSet k = {nullptr, nullptr, nullptr, nullptr, nullptr};      // the kernel instance this model runs on (with / without the general constraint path)
  std::string err;
  std::vector<void*> allocs;
  cudaStream_t stream = nullptr;            // used by cosim_step_host
  float *d_action = nullptr, *d_command = nullptr, *d_state = nullptr; uint8_t *d_term = nullptr, *d_trunc = nullptr;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return COSIM_ERR_CUDA; } } while (0)

static void* dev_upload(void* ctx, const void* src, size_t bytes) {
  cosim_handle* h = (cosim_handle*)ctx; void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 4) != cudaSuccess) throw std::runtime_error("cudaMalloc failed (model tables)");
  if (cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) throw std::runtime_error("cudaMemcpy failed (model tables)");
  h->allocs.push_back(p); return p;
}
static void* dev_zalloc(void* ctx, size_t bytes) {
  cosim_handle* h = (cosim_handle*)ctx; void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 4) != cudaSuccess) throw std::runtime_error("cudaMalloc failed (env arrays)");
  cudaMemset(p, 0, bytes ? bytes : 4);
  h->allocs.push_back(p); return p;
}
static int grid_for(const cosim_handle* h) { return h->grid; }
// every entry point runs on the handle's device, whatever the caller's current device is
struct DeviceGuard { int prev = -1; explicit DeviceGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; } ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); } };
#define ON_DEVICE(h) DeviceGuard guard_((h)->device)

extern "C" {

int cosim_create(const void* blob, size_t nbytes, int num_envs, int device, uint64_t seed, uint32_t env_offset, cosim_handle** out) {
  if (!blob || !out || num_envs <= 0) return COSIM_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fprintf(stderr, "cosim_b200: no CUDA device -- this library has no CPU path\n"); return COSIM_ERR_CUDA; }
  if (device < 0 || device >= ndev) return COSIM_ERR_ARG;
  cosim_handle* h = new cosim_handle;
  h->N = num_envs; h->device = device;
  ON_DEVICE(h);
  try {
    Uploader u = {dev_upload, h};
    setup::build_model(blob, nbytes, seed, env_offset, u, h->m);
    setup::alloc_env(h->m, num_envs, h->E, dev_zalloc, h);
    h->m.phase = (unsigned long long*)dev_zalloc(h, PH__COUNT * sizeof(unsigned long long));
  } catch (std::exception& e) {
    fprintf(stderr, "cosim_create: %s\n", e.what());
    for (void* p : h->allocs) cudaFree(p);
    delete h; return COSIM_ERR_MODEL;
  }
  // launch geometry: as many env-warps per block as fit comfortably; blocks co-reside up to the 227 KB/SM limit
  const size_t per = (size_t)h->m.ws_floats * sizeof(float);
  // one CTA per SM with as many env-warps as its shared memory holds (<= 20: five warps of 96 registers fill the 16 K registers of each SM sub-partition): the warps of a CTA move through the
  // phases of a sub-step together (block barriers in k_step), which keeps the instruction working set per SM small
  int wpb = 20; size_t budget = 224 * 1024;         // tuning overrides (experiments): COSIM_MAX_WPB, COSIM_SMEM_KB
  wpb = h->m.wpb_cap > 20 ? 24 : 20;                // which kernel build the model was laid out for (engine_setup.h)
  { const char* e = getenv("COSIM_MAX_WPB"); if (e && atoi(e) >= 1 && atoi(e) <= wpb) wpb = atoi(e); }
  { const char* e = getenv("COSIM_SMEM_KB"); if (e && atoi(e) >= 16 && atoi(e) <= 227) budget = (size_t)atoi(e) * 1024; }
  while (wpb > 1 && per * wpb + h->m.shared_floats * sizeof(float) > budget) --wpb;
  if (num_envs < wpb * 148) { wpb = (num_envs + 147) / 148; if (wpb < 1) wpb = 1; }      // small batches: spread over the SMs
  if (per * wpb > 227 * 1024) { fprintf(stderr, "cosim_create: workspace %zu B/env exceeds shared memory\n", per); for (void* p : h->allocs) cudaFree(p); delete h; return COSIM_ERR_MODEL; }
  h->wpb = wpb; h->smem = per * wpb + h->m.shared_floats * sizeof(float);
  // The dynamic shared-memory limit is per-function state shared by every handle of the process: raise it to the device
  // opt-in maximum once instead of to this handle's size (a second, smaller handle must not lower it for the first).
  int optin = 0; cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if ((size_t)optin < h->smem + 16) { fprintf(stderr, "cosim_create: %zu B of shared memory per CTA exceed the device limit %d\n", h->smem, optin); for (void* p : h->allocs) cudaFree(p); delete h; return COSIM_ERR_MODEL; }
  // the kernel build whose launch bound fits the env-warps per CTA: 24 x 80 registers, <= 20 x 96, <= 12 x 168 (engine_w24.cu / engine.cu / engine_w12.cu)
  h->k = h->m.general ? kernel_set_gen() : (wpb > 20 ? kernel_set_fast24() : (wpb <= 12 ? kernel_set_fast12() : kernel_set_fast()));
  cudaError_t e1 = cudaFuncSetAttribute(h->k.init, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 16);
  cudaError_t e2 = cudaFuncSetAttribute(h->k.reset, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 16);
  cudaError_t e3 = cudaFuncSetAttribute(h->k.step, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 16);
  if (e3 == cudaSuccess) e3 = cudaFuncSetAttribute(h->k.substep, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 16);
  { // shared-memory carve-out of the unified 256 KB: what the CTA needs, rounded up by the driver to one of its configurations -- the
    // rest is L1, which is all the stack frames / spills of the out-of-line device functions have before L2 (COSIM_CARVEOUT = percent overrides)
    int carve = (int)((h->smem + 1024 + 2621) / 2622);          // percent of 256 KB, rounded up
    if (carve > 100) carve = 100;
    { const char* e = getenv("COSIM_CARVEOUT"); if (e && atoi(e) >= 0 && atoi(e) <= 100) carve = atoi(e); }
    cudaFuncSetAttribute(h->k.step, cudaFuncAttributePreferredSharedMemoryCarveout, carve); }
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { fprintf(stderr, "cosim_create: cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3))); for (void* p : h->allocs) cudaFree(p); delete h; return COSIM_ERR_CUDA; }
  // grid = the CTAs that are resident at once (persistent CTAs walk over chunks of wpb environments); every resident warp
  // owns one global-memory slot for the contact records that do not fit its shared-memory tier
  {
    int per_sm = 0, nsm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, h->k.step, 32 * h->wpb, h->smem);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    if (per_sm < 1) per_sm = 1;
    if (nsm < 1) nsm = 148;
    const int nchunks = (num_envs + h->wpb - 1) / h->wpb;
    h->grid = nchunks < per_sm * nsm ? nchunks : per_sm * nsm;
    if (getenv("COSIM_PRINT_OCC")) fprintf(stderr, "cosim_create: k_step %d threads, %zu B dynamic smem -> %d CTA(s) per SM, grid %d, %d contact records in shared memory (capacity %d), %llu B overflow slot per warp\n",
                                           32 * h->wpb, h->smem, per_sm, h->grid, h->m.cn_k, h->m.dims[CD_ncon_max], (unsigned long long)h->m.gslot_floats * 4);
    // pooled stepping (k_step_pool): P environments per CTA and stage queue, P ~ COSIM_POOL_R x wpb (default 3; 0 / 1 = off),
    // rounded so that every CTA gets the same number of pools
    {
      // Default: on for fine terrain rasters (cells under 5 cm: dozens of contacts per env, stage durations spread over an order
      // of magnitude -- w4_p_v2 on the stairs +43 %), off otherwise: with a handful of contacts the stages are short, and warps
      // that START a stage together fetch its code together; staggered starts cost more instruction-cache misses than
      // the queue saves in barrier waits (flamingo_p_v3 on rocky_hard 2.5 -> 1.6 M env-steps/s, profiles/r02_experiments.md).
      const double cell = h->m.dims[CD_ground_type] == 1 ? 2.0 * h->m.opts[CO_hf_sx] / (h->m.dims[CD_hf_ncol] > 1 ? h->m.dims[CD_hf_ncol] - 1 : 1) : 1.0;
      int R = cell < 0.05 ? 4 : 0; { const char* e = getenv("COSIM_POOL_R"); if (e) R = atoi(e); }
      const int grid_full = per_sm * nsm;
      if (R >= 2 && h->wpb >= 4 && num_envs >= 2 * h->wpb * grid_full) {
        int target = R * h->wpb; if (target > POOL_MAX) target = POOL_MAX;
        int k = (int)((double)num_envs / ((double)grid_full * target) + 0.5); if (k < 1) k = 1;
        int P = (num_envs + grid_full * k - 1) / (grid_full * k);
        if (P > POOL_MAX) P = POOL_MAX;
        if (P > h->wpb) {
          h->pool.P = P; h->pool.img_floats = h->m.ws_floats + LOC_FLOATS; h->pool.npools = (num_envs + P - 1) / P;
          h->pool.bounds = (1 << SG_KIN) | (1 << SG_COL) | (1 << SG_SMO) | (1 << SG_NEW);
          { const char* e = getenv("COSIM_POOL_BOUNDS"); if (e) h->pool.bounds = atoi(e); }
          h->grid = h->pool.npools < grid_full ? h->pool.npools : grid_full;
          cudaFuncSetAttribute(h->k.step_pool, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 16 - 2048);
          cudaFuncSetAttribute(h->k.step_pool, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
      }
      if (getenv("COSIM_PRINT_OCC")) fprintf(stderr, "cosim_create: pooled stepping %s (P = %d envs per CTA and stage queue, %d pools, boundaries 0x%x)\n", h->pool.P ? "on" : "off", h->pool.P, h->pool.npools, h->pool.bounds);
    }
    try {
      void* p = nullptr;
      const size_t slots_per_cta = (size_t)(h->pool.P > h->wpb ? h->pool.P : h->wpb);
      const size_t bytes = (size_t)(per_sm * nsm > h->grid ? per_sm * nsm : h->grid) * slots_per_cta * h->m.gslot_floats * sizeof(float);
      if (h->pool.P) {
        void* q = nullptr;
        if (cudaMalloc(&q, (size_t)h->grid * h->pool.P * h->pool.img_floats * sizeof(float)) != cudaSuccess) throw std::runtime_error("cudaMalloc failed (pool images)");
        h->allocs.push_back(q); h->pool.images = (float*)q;
        h->pool.cost_g = (unsigned short*)dev_zalloc(h, (size_t)num_envs * 4 * sizeof(unsigned short));
        { const char* e = getenv("COSIM_POOL_MODE"); h->pool.mode = e ? atoi(e) : 0; }
      }
      if (cudaMalloc(&p, bytes ? bytes : 4) != cudaSuccess) throw std::runtime_error("cudaMalloc failed (contact overflow slots)");
      h->allocs.push_back(p); h->m.gscratch = (float*)p;
      h->sched = (int*)dev_zalloc(h, 16);
    } catch (std::exception& e) { fprintf(stderr, "cosim_create: %s\n", e.what()); for (void* p : h->allocs) cudaFree(p); delete h; return COSIM_ERR_CUDA; }
  }
  cudaStreamCreate(&h->stream);     // blocking stream: ordered after work the caller queued on the legacy default stream (reset, set)
  { void* args[] = {&h->m, &h->E}; cudaLaunchKernel(h->k.init, dim3(grid_for(h)), dim3(32 * h->wpb), args, h->smem, h->stream); }
  h->launches++;
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { fprintf(stderr, "cosim_create: init kernel failed: %s\n", cudaGetErrorString(e)); for (void* p : h->allocs) cudaFree(p); delete h; return COSIM_ERR_CUDA; }
  *out = h;
  return COSIM_OK;
}

void cosim_destroy(cosim_handle* h) {
  if (!h) return;
  { ON_DEVICE(h);
  for (void* p : h->allocs) cudaFree(p);
  if (h->stream) cudaStreamDestroy(h->stream); }
  delete h;
}
const char* cosim_last_error(const cosim_handle* h) { return h ? h->err.c_str() : "null handle"; }

int cosim_reset(cosim_handle* h, const uint8_t* mask, const float* command, float* state_out, void* stream) {
  if (!h || !state_out) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  StepArgs a = {nullptr, command, nullptr, state_out, nullptr, nullptr, mask};
  { void* args[] = {&h->m, &h->E, &a}; cudaLaunchKernel(h->k.reset, dim3(grid_for(h)), dim3(32 * h->wpb), args, h->smem, (cudaStream_t)stream); }
  h->launches++;
  CK(cudaGetLastError());
  return COSIM_OK;
}

int cosim_step(cosim_handle* h, const float* action, const float* command, const float* user_command, float* state_out,
               uint8_t* terminated, uint8_t* truncated, void* stream) {
  if (!h || !action || !state_out || !terminated || !truncated) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  StepArgs a = {action, command, user_command, state_out, terminated, truncated, nullptr};
  CK(cudaMemsetAsync(h->sched, 0, sizeof(int), (cudaStream_t)stream));
  { void* args[] = {&h->m, &h->E, &a, &h->sched, &h->pool};       // k_step takes the first four
    cudaLaunchKernel(h->pool.P ? h->k.step_pool : h->k.step, dim3(grid_for(h)), dim3(32 * h->wpb), args, h->smem, (cudaStream_t)stream); }
  h->launches++;
  CK(cudaGetLastError());
  return COSIM_OK;
}

int cosim_step_host(cosim_handle* h, const float* action_host, const float* command_host, float* state_out_host,
                    uint8_t* terminated_host, uint8_t* truncated_host) {
  if (!h || !action_host || !state_out_host || !terminated_host || !truncated_host) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  const setup::EnvDims d = setup::env_dims(h->m);
  const size_t n = (size_t)h->N;
  if (!h->d_action) {
    try {
      h->d_action = (float*)dev_zalloc(h, n * d.nu * 4); h->d_command = (float*)dev_zalloc(h, n * (d.cd > 0 ? d.cd : 1) * 4);
      h->d_state = (float*)dev_zalloc(h, n * d.sd * 4); h->d_term = (uint8_t*)dev_zalloc(h, n); h->d_trunc = (uint8_t*)dev_zalloc(h, n);
    } catch (std::exception& e) { h->err = e.what(); return COSIM_ERR_CUDA; }
  }
  CK(cudaMemcpyAsync(h->d_action, action_host, n * d.nu * 4, cudaMemcpyHostToDevice, h->stream));
  if (command_host && d.cd > 0) CK(cudaMemcpyAsync(h->d_command, command_host, n * d.cd * 4, cudaMemcpyHostToDevice, h->stream));
  int rc = cosim_step(h, h->d_action, (command_host && d.cd > 0) ? h->d_command : nullptr, nullptr, h->d_state, h->d_term, h->d_trunc, h->stream);
  if (rc != COSIM_OK) return rc;
  CK(cudaMemcpyAsync(state_out_host, h->d_state, n * d.sd * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(terminated_host, h->d_term, n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(truncated_host, h->d_trunc, n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return COSIM_OK;
}

int cosim_substep(cosim_handle* h, void* stream) {
  if (!h) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  { void* args[] = {&h->m, &h->E}; cudaLaunchKernel(h->k.substep, dim3(grid_for(h)), dim3(32 * h->wpb), args, h->smem, (cudaStream_t)stream); }
  h->launches++;
  CK(cudaGetLastError());
  return COSIM_OK;
}

int cosim_push(cosim_handle* h, const uint8_t* mask, const float* vel_world, void* stream) {
  if (!h || !vel_world) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  k_push<<<(h->N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->m, h->E, mask, vel_world);
  h->launches++;
  CK(cudaGetLastError());
  return COSIM_OK;
}

int cosim_field_dim(const cosim_handle* h, const char* field) {
  if (!h || !field) return COSIM_ERR_ARG;
  for (auto& f : setup::env_fields(h->m, h->E)) if (!strcmp(f.name, field)) return f.ptr ? f.dim : COSIM_ERR_FIELD;
  return COSIM_ERR_FIELD;
}
int cosim_field_is_int(const cosim_handle* h, const char* field) {
  if (!h || !field) return COSIM_ERR_ARG;
  for (auto& f : setup::env_fields(h->m, h->E)) if (!strcmp(f.name, field)) return f.is_int;
  return COSIM_ERR_FIELD;
}
int cosim_get(cosim_handle* h, const char* field, void* dst, void* stream) {
  if (!h || !field || !dst) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  for (auto& f : setup::env_fields(h->m, h->E)) if (!strcmp(f.name, field)) {
    if (!f.ptr) { h->err = std::string("field '") + field + "' needs cosim_set_debug(h, 1)"; return COSIM_ERR_FIELD; }
    CK(cudaMemcpyAsync(dst, f.ptr, (size_t)h->N * f.dim * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return COSIM_OK;
  }
  h->err = std::string("unknown field '") + field + "'";
  return COSIM_ERR_FIELD;
}
int cosim_set(cosim_handle* h, const char* field, const void* src, void* stream) {
  if (!h || !field || !src) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  const std::string n(field);
  if (n != "qpos" && n != "qvel" && n != "qacc_warmstart" && n != "torque") { h->err = "cosim_set: only qpos, qvel, qacc_warmstart, torque are writable"; return COSIM_ERR_FIELD; }
  for (auto& f : setup::env_fields(h->m, h->E)) if (n == f.name) {
    CK(cudaMemcpyAsync(f.ptr, src, (size_t)h->N * f.dim * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return COSIM_OK;
  }
  return COSIM_ERR_FIELD;
}
int cosim_set_debug(cosim_handle* h, int enable) {
  if (!h) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  if (enable) {
    if (h->debug == 2) h->E = h->Edbg;      // re-enable: restore the saved pointers
    else if (!h->E.dbg_contacts) { try { setup::alloc_debug(h->m, h->N, h->E, dev_zalloc, h); } catch (std::exception& e) { h->err = e.what(); return COSIM_ERR_CUDA; } }
    h->debug = 1;
  } else if (h->debug == 1) {
    h->Edbg = h->E;
    h->E.dbg_contacts = h->E.dbg_heightmap = h->E.dbg_cfrc = h->E.dbg_sens = h->E.dbg_qacc = nullptr; h->E.dbg_hmcell = h->E.dbg_iters = nullptr;
    h->debug = 2;
  }
  return COSIM_OK;
}

int cosim_stats_reduce(cosim_handle* h, double* out, void* stream) {
  if (!h || !out) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  CK(cudaMemsetAsync(out, 0, COSIM_NSTAT * sizeof(double), (cudaStream_t)stream));
  int blocks = (h->N + 1023) / 1024; if (blocks > 148) blocks = 148; if (blocks < 1) blocks = 1;
  k_stats<<<blocks, 8 * ST__COUNT, 0, (cudaStream_t)stream>>>(h->E.stats, h->N, out);
  h->launches++;
  CK(cudaGetLastError());
  return COSIM_OK;
}
int cosim_stats_clear(cosim_handle* h, void* stream) {
  if (!h) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  CK(cudaMemsetAsync(h->E.stats, 0, (size_t)h->N * ST__COUNT * sizeof(float), (cudaStream_t)stream));
  return COSIM_OK;
}
int cosim_rng_probe(cosim_handle* h, uint32_t rng_stream, uint32_t step, int nidx, uint32_t* out, void* stream) {
  if (!h || !out || nidx <= 0) return COSIM_ERR_ARG;
  ON_DEVICE(h);
  const int n = h->N * nidx;
  k_rng_probe<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->m, h->N, rng_stream, step, nidx, out);
  h->launches++;
  CK(cudaGetLastError());
  return COSIM_OK;
}

/* profiling builds (-DCOSIM_PHASE_TIMING): cycle counters per phase, summed over env-warps; zeros otherwise */
int cosim_phase_cycles(cosim_handle* h, unsigned long long* out_host, int reset) {
  if (!h || !out_host) return COSIM_ERR_ARG;
  for (int i = 0; i < PH__COUNT; ++i) out_host[i] = 0;
#ifdef COSIM_PHASE_TIMING
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out_host, h->m.phase, PH__COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) CK(cudaMemset(h->m.phase, 0, PH__COUNT * sizeof(unsigned long long)));
#endif
  return COSIM_OK;
}
int cosim_num_envs(const cosim_handle* h) { return h ? h->N : COSIM_ERR_ARG; }
int cosim_dim(const cosim_handle* h, const char* name) { return (h && name) ? setup::dim_by_name(h->m, name) : COSIM_ERR_ARG; }
int cosim_launch_count(const cosim_handle* h) { return h ? h->launches : COSIM_ERR_ARG; }
int cosim_smem_bytes_per_env(const cosim_handle* h) { return h ? h->m.ws_floats * 4 : COSIM_ERR_ARG; }
int cosim_warps_per_block(const cosim_handle* h) { return h ? h->wpb : COSIM_ERR_ARG; }
int cosim_pool_size(const cosim_handle* h) { return h ? h->pool.P : COSIM_ERR_ARG; }
int cosim_general_path(const cosim_handle* h) { return h ? h->m.general : COSIM_ERR_ARG; }

}  // extern "C"
