// engine_kernels.cuh -- the per-env kernels of libcosim_b200.so (one warp owns one environment; its working set lives in that
// warp's slice of dynamic shared memory).  Included by engine.cu (COSIM_GENERAL 0: the reference's condim 3 / pyramidal / Newton
// case) and engine_gen.cu (COSIM_GENERAL 1: with the general constraint path); KN() gives the two sets distinct names and
// kernel_set_*() hands their addresses to the host code, which launches through cudaLaunchKernel.
#pragma once
#include <cuda_runtime.h>
#include "engine_env.h"
#if COSIM_GENERAL
#define KN(name) name##_gen
#elif defined(COSIM_W24)
#define KN(name) name##_fast24
#elif defined(COSIM_W12)
#define KN(name) name##_fast12
#else
#define KN(name) name##_fast
#endif
// threads per CTA the kernels are compiled for: 20 env-warps of 96 registers fill the register file of an SM; engine_w24.cu compiles
// the fast path for 768 threads (24 warps of 80 registers), engine_w12.cu for 384 (12 warps of up to 168 registers)
#ifndef COSIM_LB
#define COSIM_LB 640
#endif
struct KernelSet { const void *init, *reset, *step, *step_pool, *substep; };
enum { SG_PRO = 0, SG_KIN, SG_COL, SG_SMO, SG_NEW, SG_EPI, POOL_MAX = 128 };
struct PoolArgs { float* images; int P, img_floats, bounds, npools, mode; unsigned short* cost_g; };   // mode 0: stage queue, 1: sorted rounds (lock-step start); cost_g [N][4] stage durations carried from step to step
// image = [workspace (ws_floats) | locals of the control step (16 floats, read / written in place)]
enum { LOC_ACTIVE = 0, LOC_SIM_STEP, LOC_NSTEP, LOC_NOBS, LOC_RM, LOC_TABS, LOC_TSQ, LOC_TMAX, LOC_ITERS, LOC_FLOATS = 16 };

// ------------------------------------------------------------------------------------------ kernels
extern __shared__ __align__(16) float g_smem[];

// The model header (dims, opts, table pointers, workspace offsets: ~1.8 KB) is copied to shared memory once per
// CTA so that the out-of-line device functions read it with LDS instead of generic loads from the param bank.
#define MODEL_FLOATS ((int)((sizeof(ModelDev) + 15) / 16 * 4))
// CTA prologue: copy the model header and the table arena to shared memory, re-point the table pointers at the copy
#define CTA_PROLOGUE()                                                                                   \
  {                                                                                                      \
    const uint32_t* src_ = (const uint32_t*)&mp;                                                         \
    uint32_t* dst_ = (uint32_t*)g_smem;                                                                  \
    for (int i = threadIdx.x; i < (int)(sizeof(ModelDev) / 4); i += blockDim.x) dst_[i] = src_[i];       \
    const uint4* asrc_ = (const uint4*)mp.arena_g;                                                       \
    uint4* adst_ = (uint4*)(g_smem + MODEL_FLOATS);                                                      \
    for (int i = threadIdx.x; i < mp.arena_bytes / 16; i += blockDim.x) adst_[i] = __ldg(asrc_ + i);     \
    __syncthreads();                                                                                     \
    for (int i = threadIdx.x; i < mp.nslots; i += blockDim.x)                                            \
      *(const uint8_t**)((char*)g_smem + mp.slot_field[i]) = (const uint8_t*)adst_ + 16 * (size_t)mp.slot_off16[i]; \
    __syncthreads();                                                                                     \
  }                                                                                                      \
  const ModelDev& m = *(const ModelDev*)g_smem;                                                          \
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb_ = blockDim.x >> 5;                    \
  float* const wsf_ = g_smem + m.shared_floats + (size_t)warp * m.ws_floats;                             \
  const WSP ws = {(uint32_t)__cvta_generic_to_shared(wsf_)};                                             \
  /* global overflow slot of this resident warp (contact records beyond the shared-memory tier) */       \
  if (lane == 0) *(float**)(wsf_ + m.off[W_GPTR]) = m.gscratch + (size_t)(blockIdx.x * wpb_ + warp) * m.gslot_floats; \
  __syncwarp();
// The grid is sized to what is resident at once (engine.cu cosim_create); each CTA walks over chunks of wpb environments.
#define FOR_ENV_CHUNKS() for (int env = blockIdx.x * wpb_ + warp; env - warp < E.N; env += gridDim.x * wpb_)

__global__ void __launch_bounds__(COSIM_LB, 1) KN(k_init)(const __grid_constant__ ModelDev mp, const EnvArrays E) {
  CTA_PROLOGUE();
  FOR_ENV_CHUNKS() { if (env < E.N) init_env(m, E, env, ws, lane); __syncwarp(); }
}
__global__ void __launch_bounds__(COSIM_LB, 1) KN(k_reset)(const __grid_constant__ ModelDev mp, const EnvArrays E, const StepArgs a) {
  CTA_PROLOGUE();
  const int cd = MD(command_dim);
  FOR_ENV_CHUNKS() {
    if (env < E.N && !(a.mask && !a.mask[env]))
      reset_env(m, E, env, ws, a.command ? a.command + (size_t)env * cd : nullptr, a.state_out + (size_t)env * MD(state_dim), lane);
    __syncwarp();
  }
}
// KN(k_step): every warp of the CTA (also the padding warps of the last chunk) walks through the phase barriers.  Chunks of
// wpb environments are claimed from a counter (dynamic: the cost of a chunk depends on what its robots are doing).
__global__ void __launch_bounds__(COSIM_LB, 1) KN(k_step)(const __grid_constant__ ModelDev mp, const EnvArrays E, const StepArgs a, int* sched) {
  CTA_PROLOGUE();
  __shared__ int s_chunk;
  const int nchunks = (E.N + wpb_ - 1) / wpb_;
  for (;;) {
    if (threadIdx.x == 0) s_chunk = atomicAdd(sched, 1);
    __syncthreads();
    const int chunk = s_chunk;
    __syncthreads();
    if (chunk >= nchunks) break;
    const int env = chunk * wpb_ + warp, have_env = env < E.N;
    step_env(m, E, have_env ? env : 0, ws, a, lane, have_env, 1);
  }
}
// ---- pooled step kernel ----------------------------------------------------------------------------------------------
// KN(k_step) above moves the wpb environments of a chunk through the stages of a sub-step in lock step: a stage lasts as long as
// its slowest environment (a wheel touching down, a Newton solve that needs five iterations), and 35 - 65 % of the warp time
// is spent waiting at the stage barriers (profiles/).  Here a CTA owns a POOL of P > wpb environments whose workspace images
// live in global memory (L2-resident: 148 x P x ~10 KB), and a stage is a queue: every warp claims the next environment of the
// pool, copies its image into its shared-memory slot, runs the stage, copies it back.  All warps of the SM still execute ONE
// stage's code at a time (what the barriers were for: the instruction cache), but the barrier now ends a queue of P tasks
// instead of one task per warp, and the tasks are started longest-first (duration of the same stage in the previous
// sub-step).  Same arithmetic, same results as KN(k_step).
static __device__ __forceinline__ void ws_load(float* ws, const float* img, int nfl, int lane) {
  const float4* src = (const float4*)img; float4* dst = (float4*)ws; const int n4 = nfl >> 2;
  int i = lane;
  for (; i + 7 * 32 < n4; i += 8 * 32) {          // eight 16-byte loads in flight per lane
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(src + i + 32 * k);
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[i + 32 * k] = v[k];
  }
  for (; i < n4; i += 32) dst[i] = __ldcg(src + i);
  __syncwarp();
}
static __device__ __forceinline__ void ws_store(float* img, const float* ws, int nfl, int lane) {
  float4* dst = (float4*)img; const float4* src = (const float4*)ws; const int n4 = nfl >> 2;
  __syncwarp();
#pragma unroll 4
  for (int i = lane; i < n4; i += 32) __stcg(dst + i, src[i]);
}
static __device__ __forceinline__ int stage_kind(int pos, int nst) { return pos == 0 ? SG_PRO : (pos == nst - 1 ? SG_EPI : 1 + ((pos - 1) & 3)); }
__global__ void __launch_bounds__(COSIM_LB, 1) KN(k_step_pool)(const __grid_constant__ ModelDev mp, const EnvArrays E, const StepArgs a, int* sched, const PoolArgs pa) {
  CTA_PROLOGUE();
  __shared__ int s_pool, s_next;
  __shared__ unsigned short s_cost[4][POOL_MAX];      // cycles >> 8 of the last KIN / COL / SMO / NEW group of every pool env
  __shared__ unsigned char s_order[POOL_MAX];
  const int fs = MD(frame_skip), nst = 2 + 4 * fs, wsf = m.ws_floats;
  float* img0 = pa.images + (size_t)blockIdx.x * pa.P * pa.img_floats;
  for (;;) {
    if (threadIdx.x == 0) s_pool = atomicAdd(sched, 1);
    __syncthreads();
    const int pool = s_pool;
    if (pool >= pa.npools) break;
    const int base = pool * pa.P, n = min(pa.P, E.N - base);
    for (int i = threadIdx.x; i < 4 * n; i += blockDim.x) s_cost[i & 3][i >> 2] = pa.cost_g[(size_t)base * 4 + i];      // durations of the previous control step
    __syncthreads();
    int pos = 0;
    while (pos < nst) {
      int gend = pos;           // the group of stages [pos, gend] runs back to back on one warp; a CTA-wide boundary follows it
      while (gend < nst - 1 && !((pa.bounds >> stage_kind(gend, nst)) & 1)) ++gend;
      const int k0 = stage_kind(pos, nst);
      if ((int)threadIdx.x < n) {      // start order: longest first
        const int t = threadIdx.x; int rank = t;
        if (k0 >= SG_KIN && k0 <= SG_NEW) {
          const unsigned short c = s_cost[k0 - 1][t]; rank = 0;
          for (int j = 0; j < n; ++j) { const unsigned short cj = s_cost[k0 - 1][j]; rank += (cj > c) || (cj == c && j < t); }
        }
        s_order[rank] = (unsigned char)t;
      }
      if (threadIdx.x == 0) s_next = 0;
      __syncthreads();
#if defined(COSIM_PHASE_TIMING)
      long long tw0 = clock64();
#endif
      for (int round = 0;; ++round) {
        int i = 0;
        if (pa.mode == 1) {         // sorted rounds: the wpb tasks of a round (similar predicted duration) start together
          if (round > 0) __syncthreads();
          if (round * wpb_ >= n) break;
          i = round * wpb_ + warp;
          if (i >= n) continue;
        } else {
          if (lane == 0) i = atomicAdd(&s_next, 1);
          i = __shfl_sync(0xffffffffu, i, 0);
          if (i >= n) break;
        }
        i = s_order[i];
        const int env = base + i;
        float* img = img0 + (size_t)i * pa.img_floats;
        int* loc = (int*)(img + wsf); float* locf = img + wsf;
        if (pos > 0) {
          if (!__ldcg(loc + LOC_ACTIVE)) continue;        // the env was reset in the prologue: no step
          ws_load(wsf_, img, wsf, lane);
        }
        const long long t0 = clock64();
        int active = 1;
        for (int st = pos; st <= gend && active; ++st) {
          const int k = stage_kind(st, nst);
          PH_DECL;
          if (k == SG_PRO) {
            if (lane == 0) *(float**)(wsf_ + m.off[W_GPTR]) = m.gscratch + (size_t)(blockIdx.x * pa.P + i) * m.gslot_floats;
            __syncwarp();
            StepLocals L; step_prologue(m, E, env, ws, a, lane, L);
            if (lane == 0) {
              loc[LOC_ACTIVE] = L.active; loc[LOC_SIM_STEP] = L.sim_step; loc[LOC_NSTEP] = (int)L.nstep; loc[LOC_NOBS] = (int)L.nobs;
              locf[LOC_RM] = L.rm; locf[LOC_TABS] = L.tabs; locf[LOC_TSQ] = L.tsq; locf[LOC_TMAX] = L.tmax; loc[LOC_ITERS] = 0;
            }
            __syncwarp();
            active = L.active;
          } else if (k == SG_KIN) { substep_pre(m, ws, lane); stage_kin(m, ws, lane); PH_MARK(PH_KIN); }
          else if (k == SG_COL) { stage_collide(m, ws, lane); PH_MARK(PH_COLLIDE); }
          else if (k == SG_SMO) { stage_smooth(m, ws, lane); }
          else if (k == SG_NEW) {
            int it = stage_newton(m, ws, lane);
            PH_MARK(PH_NEWTON);
            it = substep_post(m, ws, lane, it);
            if (lane == 0) loc[LOC_ITERS] += it;
            __syncwarp();
          } else {
            StepLocals L; L.active = 1; L.sim_step = loc[LOC_SIM_STEP]; L.nstep = (uint32_t)loc[LOC_NSTEP]; L.nobs = (uint32_t)loc[LOC_NOBS];
            L.rm = locf[LOC_RM]; L.tabs = locf[LOC_TABS]; L.tsq = locf[LOC_TSQ]; L.tmax = locf[LOC_TMAX];
            step_epilogue(m, E, env, ws, a, lane, L, loc[LOC_ITERS]);
          }
        }
        if (k0 >= SG_KIN && k0 <= SG_NEW && lane == 0) { const long long d = (clock64() - t0) >> 8; s_cost[k0 - 1][i] = (unsigned short)(d > 65535 ? 65535 : d); }
        if (gend < nst - 1 && active) ws_store(img, wsf_, wsf, lane);
      }
#if defined(COSIM_PHASE_TIMING)
      { const long long tw1 = clock64(); __syncthreads();
        if (lane == 0) { const int slot = k0 == SG_COL ? PH_WAIT_COLLIDE : (k0 == SG_NEW ? PH_WAIT_NEWTON : (k0 == SG_SMO ? PH_WAIT_SMOOTH : PH_WAIT_KIN));
          atomicAdd(m.phase + slot, (unsigned long long)(clock64() - tw1)); atomicAdd(m.phase + PH_IO, (unsigned long long)(tw1 - tw0)); } }
#else
      __syncthreads();
#endif
      pos = gend + 1;
    }
    for (int i = threadIdx.x; i < 4 * n; i += blockDim.x) pa.cost_g[(size_t)base * 4 + i] = s_cost[i & 3][i >> 2];
  }
}
__global__ void __launch_bounds__(COSIM_LB, 1) KN(k_substep)(const __grid_constant__ ModelDev mp, const EnvArrays E) {
  CTA_PROLOGUE();
  FOR_ENV_CHUNKS() { if (env < E.N) substep_env(m, E, env, ws, lane); __syncwarp(); }
}
KernelSet KN(kernel_set)() {
  KernelSet k = {(const void*)KN(k_init), (const void*)KN(k_reset), (const void*)KN(k_step), (const void*)KN(k_step_pool), (const void*)KN(k_substep)};
  return k;
}
