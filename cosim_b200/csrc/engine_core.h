// engine_core.h -- warp-per-env rigid-body step for sm_100a (B200).
//
// One warp owns one environment: its whole working set (poses, spatial inertias, mass matrix,
// contact Jacobians, solver vectors) lives in that warp's slice of shared memory; lanes split the
// per-body / per-dof / per-row / per-vertex loops and meet at __syncwarp().  No tensor cores here:
// this is small dense fp32 linear algebra and branchy collision, bound by the FP32 pipe and
// shared-memory latency.
//
// The same source compiles in a single-lane host mode (COSIM_HOST_EMU, LANES = 1) used ONLY by
// tests/hostsim to debug kernel logic on machines without a GPU.  The product library is built
// from engine.cu with nvcc and has no CPU path.
//
// What each stage replaces (MuJoCo 3.2.7 is an un-vendored dependency of the reference; semantics
// per SURVEY.md Appendix B; reference call site: gymnasium do_simulation from
// /root/reference/envs/flamingo_p_v3/flamingo_p_v3.py:189):
//   kinematics/com_pos/crb  <- mj_kinematics, mj_comPos, mj_crb       (K1 fk_crb)
//   chol_factor/chol_solve  <- mj_factorM / mj_solveM (tree-sparse, leaf-first elimination; dense variant for coupled Hessians)
//   collide_*               <- mj_collision: hfield|plane x convex     (K3 collide)
//   make_constraint         <- mj_makeConstraint/Impedance/reference   (K4 constraints)
//   newton_solve            <- mj_fwdConstraint, Newton + exact LS     (K5 newton)
//   rne_bias/actuation/integrate <- mj_rne, mj_fwdActuation, mj_implicit(fast), mj_advance (K6)
//   pd_delay                <- <Robot>.step PD + ControlManager.delay_filter (K7)
//   sensors/observe         <- mj_sensor*, _get_obs, StateBuildWrapper (K8)
//   height_map              <- get_height_map + mj_rayHfield           (K9)
#pragma once
// COSIM_GENERAL = 1 compiles the general constraint path (engine_general.h) into the stage functions; the library builds its
// kernels twice (engine.cu: 0, engine_gen.cu: 1) and picks per model, so the reference's condim 3 / pyramidal / Newton case
// carries none of that code (it cost 4.7 % when it was a run-time branch of one kernel, profiles/r02_experiments.md).
#ifndef COSIM_GENERAL
#define COSIM_GENERAL 0
#endif
#if COSIM_GENERAL
#define IF_GENERAL(m) ((m).general)
#else
#define IF_GENERAL(m) 0
#endif
#include <stdint.h>
#include <math.h>
#include <float.h>
#include "../../include/cosim_blob.h"
#ifdef COSIM_HOST_EMU
struct float4 { float x, y, z, w; };
#endif

#ifdef COSIM_HOST_EMU
#define DEV static inline
#define DEVM inline                      /* member functions */
#define DEV_NOINLINE static
#define LANES 1
#define SYNC() ((void)0)
#define LDG(p) (*(p))
#define LDGB(p) (*(p))
#define SHP(p) (p)
typedef float* WSP;
#define WSF(w) (w)
#define LANE_REFRESH() ((void)0)
static inline float wsum(float v) { return v; }
static inline float wmaxf(float v) { return v; }
static inline int wor(int v) { return v; }
static inline void wargmax(float& v, int& i) {}
static inline unsigned wballot(int p) { return p ? 1u : 0u; }
template <class T> static inline T wshfl(T v, int) { return v; }
static inline float fast_ndtri(float p);
static inline int popc32(unsigned x) { return __builtin_popcount(x); }
static inline int __float_as_int_emu(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float_emu(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int ctz32(unsigned x) { return __builtin_ctz(x); }
#else
#define DEV static __device__ __forceinline__      /* internal linkage: engine.cu and engine_gen.cu compile different bodies */
#define DEVM __device__ __forceinline__
static __device__ __forceinline__ int popc32(unsigned x) { return __popc(x); }
static __device__ __forceinline__ int __float_as_int_emu(float f) { return __float_as_int(f); }
static __device__ __forceinline__ float __int_as_float_emu(int i) { return __int_as_float(i); }
static __device__ __forceinline__ int ctz32(unsigned x) { return __ffs(x) - 1; }
#define DEV_NOINLINE static __device__ __noinline__
#define LANES 32
#define SYNC() __syncwarp()
// Pointers into the table arena: the kernels' prologue re-points them at the shared-memory copy, but they travel through ModelDev as
// generic pointers, and a generic load costs LD.E + two R2UR (descriptor) where LDS would do.  SHP() wraps such a pointer into its
// 32-bit shared-window address with ld.shared accessors (__builtin_assume(__isShared(p)) made nvcc 12.9 drop the loops that used p).
// Never apply it to a field that may be NULL or that stays in global memory: hull_verts, hfield_data, sup_off, sup_cand, ctab, hf_max8.
template <class T> struct ShTab {      // read-only table in shared memory, addressed by its 32-bit shared-window address
  uint32_t a;
  static __device__ __forceinline__ T ld(uint32_t addr) {
    if constexpr (sizeof(T) == 2) { unsigned short v; asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr)); return (T)v; }
    else { static_assert(sizeof(T) == 4, "ShTab: 2- or 4-byte elements"); uint32_t v; asm("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr)); T r; memcpy(&r, &v, 4); return r; }
  }
  __device__ __forceinline__ T operator[](int i) const { return ld(a + (uint32_t)i * (uint32_t)sizeof(T)); }
  __device__ __forceinline__ T operator*() const { return ld(a); }
  __device__ __forceinline__ ShTab operator+(int k) const { return ShTab{a + (uint32_t)k * (uint32_t)sizeof(T)}; }
  __device__ __forceinline__ operator const T*() const { return (const T*)__cvta_shared_to_generic(a); }      // sites that hand the table to a pointer interface
};
template <class T> static __device__ __forceinline__ ShTab<T> shp_(const T* p) { return ShTab<T>{(uint32_t)__cvta_generic_to_shared(p)}; }
#define SHP(p) shp_(p)
// The env's workspace travels through the out-of-line functions as its 32-bit shared-window address, not as a 64-bit generic pointer
// (one register / stack slot instead of two on every call; the loads stay LDS because the address space is known at the conversion).
struct WSP { uint32_t a; };
// `lane` is re-read from the special register at the top of every out-of-line function: the parameter then dies (the functions are
// static, so it is not passed at all) instead of travelling through local memory next to ws: +3 % on flamingo_p_v3 (same-box A / B,
// profiles/r02_experiments.md).  Same values, but another instruction order: results differ from the previous build in the last bits
// (tools/lib_check.py), as they do between any two builds.
static __device__ __forceinline__ int lane_id_() { int l; asm("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }
#define LANE_REFRESH() lane = lane_id_()
#define WSF(w) ((float*)__cvta_shared_to_generic((size_t)(w).a))
#define LDG(p) (*(p))          // small model tables: shared-memory copy of the arena (plain load; __ldg would fault)
#define LDGB(p) __ldg(p)       // big read-only tables in global memory: hull vertices, support maps, height field
DEV float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DEV float wmaxf(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
DEV int wor(int v) { return __any_sync(0xffffffffu, v); }
DEV unsigned wballot(int p) { return __ballot_sync(0xffffffffu, p); }
template <class T> DEV T wshfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// arg-max with lowest-index tie break; result broadcast to all lanes
DEV void wargmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}
#endif

// optional per-phase cycle counters (profiling builds only: -DCOSIM_PHASE_TIMING)
enum { PH_KIN = 0, PH_COLLIDE, PH_CONSTRAINT, PH_SMOOTH, PH_NEWTON, PH_INTEGRATE, PH_OBS, PH_IO, PH_NEWTON_ITERS, PH_LS_EVALS, PH_SUPPORT_CALLS, PH_MPR_CALLS,
       PH_WAIT_KIN, PH_WAIT_COLLIDE, PH_WAIT_SMOOTH, PH_WAIT_NEWTON, PH_HIST = 16, PH_HIST_BINS = 64, PH__COUNT = PH_HIST + 2 * PH_HIST_BINS };     // PH_WAIT_*: cycles spent at the barrier that ends the phase
#if defined(COSIM_PHASE_TIMING) && !defined(COSIM_HOST_EMU)
#define PH_DECL long long ph_t0_ = clock64()
#define PH_MARK(k) do { long long t_ = clock64(); if (lane == 0) { atomicAdd((unsigned long long*)m.phase + (k), (unsigned long long)(t_ - ph_t0_)); \
    /* histograms (8 K-cycle bins) of the per-warp collision and Newton phase times */ \
    if ((k) == PH_COLLIDE || (k) == PH_NEWTON) { long long b_ = (t_ - ph_t0_) >> 13; if (b_ >= PH_HIST_BINS) b_ = PH_HIST_BINS - 1; \
      atomicAdd((unsigned long long*)m.phase + PH_HIST + ((k) == PH_NEWTON ? PH_HIST_BINS : 0) + b_, 1ull); } } ph_t0_ = clock64(); } while (0)
#define PH_COUNT(k, n) do { if (lane == 0) atomicAdd((unsigned long long*)m.phase + (k), (unsigned long long)(n)); } while (0)
#else
#define PH_DECL ((void)0)
#define PH_MARK(k) ((void)0)
#define PH_COUNT(k, n) ((void)0)
#endif

// block-wide phase barrier (k_step only): all env-warps of a CTA walk through the phases of a sub-step together, so the
// SM's instruction caches hold ONE phase's code at a time instead of a dozen unrelated program counters (profiles/)
#ifdef COSIM_HOST_EMU
#define BSYNC(on) ((void)0)
#define BSYNC_IF(on, bit) ((void)0)
#define CTA_SYNC() ((void)0)
#define CTA_SYNC_OR(x) ((x) != 0)
#else
#define BSYNC(on) do { if (on) __syncthreads(); } while (0)
#define BSYNC_IF(on, bit) do { if ((on) && ((m.bsync_mask >> (bit)) & 1)) __syncthreads(); } while (0)
#define CTA_SYNC() __syncthreads()
#define CTA_SYNC_OR(x) __syncthreads_or(x)
#endif

DEV int imax(int a, int b) { return a > b ? a : b; }
DEV int imin(int a, int b) { return a < b ? a : b; }
// Loops are kept rolled (unroll 1) unless marked otherwise: the step kernel is instruction-fetch bound (hot code
// must fit the 32 KB L1.5 I-cache of an SM that runs a dozen env-warps at different program counters), see profiles/.
#ifdef COSIM_HOST_EMU
#define NOUNROLL
#else
#define NOUNROLL _Pragma("unroll 1")
#endif
#define FOR_LANE(i, n) NOUNROLL for (int i = lane; i < (n); i += LANES)
#define MINVALF 1e-15f

// ------------------------------------------------------------------------------------------ model
// Read-only model shared by every env (device pointers; f64 blob sections are converted to f32 once
// at create time).  Passed to kernels by value (constant bank).
struct ModelDev {
  int dims[48];
  float opts[48];
  // bodies (index 0 = world)
  const int *body_parent, *body_jnt, *body_dofadr, *body_dofnum, *body_subsize, *body_dofmask;
  const float *body_pos, *body_quat, *body_ipos, *body_inertia, *body_mass;
  const int *level_start, *level_body; int nlevels;
  // joints / dofs
  const int *jnt_type, *jnt_body, *jnt_qposadr, *jnt_dofadr, *jnt_limited, *jnt_actfrclimited;
  const float *jnt_pos, *jnt_axis, *jnt_range, *jnt_actfrcrange;
  const int *dof_body, *dof_jnt, *dof_parent, *dof_fl_random;
  const float *dof_armature, *dof_damping, *dof_frictionloss;
  const int *mpair_i, *mpair_j; int nmpair;
  const int* pair_geom;                        // [npair][2] geom-geom candidates (cosim_b200/model.py self_collision_pairs)
  const float* geom_aabb;                      // [ngeom][6] geom-frame box: centre, half extents
  const float *qpos0;
  // actuators + PD tables
  const int *act_dof, *act_qadr, *act_mode, *act_ctrllimited;
  const float *act_gear, *act_ctrlrange, *act_kp, *act_kd, *act_scale, *act_posfac, *act_gamma, *act_clip;
  // geoms
  const int *geom_type, *geom_body, *geom_vadr, *geom_vnum, *geom_fr_random;
  const float *geom_size, *geom_pos, *geom_quat, *geom_friction, *geom_center, *geom_rbound;
  const float *hull_verts, *hfield_data;
  // hull support maps (cosim_b200/model.py:build_support_map): per mesh geom a table of 6*8*8 direction buckets;
  // sup_cand[k] = (x, y, z, vertex index) of the k-th candidate, so one 16-byte load per candidate
  const int *geom_supadr, *sup_off; const float4* sup_cand;
  const uint16_t* sup_off16; const int* geom_supbase;   // bucket offsets relative to the mesh's first candidate, in the shared-memory arena when they fit (else NULL)
  // lower-triangle (i, k) pairs of an nv x nv matrix sorted by k descending, packed (i << 8) | k: the trailing
  // sub-matrix update of Cholesky column j is the prefix of length (nv-j-1)(nv-j)/2
  const int* tri; int shared_floats;          // tri: pairs (i, k <= i) of the lower triangle, rows in increasing order
  const int *ctab, *coff;                     // tree-sparse elimination: pair list of column j = ctab[coff[j] .. coff[j + 1]) (ctab in global memory)
  const uint16_t* ctab16;                     // 16-bit copy of ctab in the shared-memory arena when it fits without costing an env-warp (else NULL)
  // table arena (engine_setup.h): global copy + which pointer fields of this struct point into it
  const uint8_t* arena_g; int arena_bytes, nslots; uint16_t slot_field[112], slot_off16[112];
  float ground_friction[4];
  // equality
  const int *eq_body1, *eq_body2;
  const float *eq_anchor1, *eq_anchor2, *eq_solref, *eq_solimp;
  float imu_pos[3], imu_quat[4];
  // env layer tables
  const int *dofpos_qadr, *dofvel_dadr, *initnoise_qadr, *term_body, *massnoise_body;
  const float *dofpos_fac, *dofvel_fac;
  const int *sobs_kind, *sobs_dim, *sobs_interval, *sobs_off, *nobs_kind, *nobs_dim, *nobs_interval, *nobs_off;
  const float *sobs_scale, *nobs_scale, *noise;
  // randomisation ranges pre-reduced on the host in double: lo and (hi - lo)
  float rnd_lo[8], rnd_span[8];
  // workspace layout (float offsets into the per-warp shared slice)
  int off[80]; int ws_floats;
  uint32_t seed_lo, seed_hi, env_offset;
  unsigned long long* phase;      // [PH__COUNT] cycle counters, profiling builds only (else NULL)
  int bsync_mask;                 // which of the CTA-wide phase barriers are enabled (bits 0 .. 3: after the four phases of forward(), 4: end of a sub-step, 5: after the step prologue, 6 .. 8: inside the kinematics / collision / Newton phase)
  // contact store: one record of cr_stride floats per contact (layout CR_*).  The first cn_k records of an env live in
  // its shared-memory workspace (W_CN_REC), the rest in the global-memory overflow slot of the resident warp
  // (gscratch + slot * gslot_floats, L2-resident: only ~3000 slots exist per GPU).  Capacity = dims[CD_ncon_max].
  int cn_k, cr_stride; float* gscratch; unsigned long long gslot_floats;
  int wpb_cap;                    // env-warps per CTA the model was laid out for: 20 (640-thread kernels) or 24 (768-thread kernels, engine_w24.cu)
  // conservative culls ahead of the terrain narrow phase: max height per 8 x 8 block of cells, and per-geom bounding
  // cylinders (geom frame: centre, unit axis, radius, half length; radius 0 = none)
  const float* hf_max8; int hf_mrow, hf_mcol; const float* geom_bcyl;
  int hf_fine;                    // raster cells under 0.2 m: the terrain pass with culls and block-wise sweep (collide_hfield_all<true>)
  // general constraint path (engine_general.h: condim 1 / 4 / 6, elliptic cone, PGS): on / off, row capacity, float offset of its
  // region inside the env's global-memory slot (behind the contact-record overflow)
  int general, gen_rows; unsigned long long gen_off;
  // box-box pairs (mjc_BoxBox: up to 8 contacts): scratch of 32 x 56 floats in the env's global-memory slot, offset bb_off (0 pairs: unused)
  int n_boxbox; unsigned long long bb_off;
};
// general path: contacts that get rows, and the floats of its region for NR rows of nv columns and NC contacts (layout: engine_general.h gen_view)
enum { GEN_MAX_CON = 64, GEN_CON_STRIDE = 44 };
static inline size_t gen_region_floats(int NR, int nv, int NC) { return (size_t)2 * NR * nv + (size_t)13 * NR + (size_t)NC * GEN_CON_STRIDE + 32; }
#define TB(field) SHP(m.field)      /* arena table `field` (shared memory inside the kernels) */
#define MD(name) (m.dims[CD_##name])
#define MO(name) (m.opts[CO_##name])

enum WsField {
  W_QPOS, W_QVEL, W_CTRL, W_WARM, W_QACC, W_XPOS, W_XQUAT, W_XMAT, W_XIPOS, W_XANCHOR, W_XAXIS, W_GXPOS, W_GXMAT,
  W_SCOM, W_CINERT, W_CRB, W_CDOF, W_CDOFDOT, W_CVEL, W_CACC, W_CFRC, W_BUF, W_M, W_A, W_INVD,
  W_FSMOOTH, W_ASMOOTH, W_FCON, W_GRAD, W_SEARCH, W_MV, W_MA, W_TMPV, W_TMPW,
  W_BMASS, W_INVWD, W_INVWB, W_FLOSS, W_GMU, W_SCAL,
  W_FR_D, W_FR_AREF, W_LM_SIGN, W_LM_D, W_LM_AREF,
  W_CN_REC, W_GPTR, W_RING, W_BV, W_CSTART, W_BS, W_CN_J,
  W_EQ_J, W_EQ_D, W_EQ_AREF, W_EQ_X, W_EQ_V, W_EQ_F, W_SENS, W_RAW, W_ACT, W_FILT, W_KP, W_KD, W_GTASK, W_CNT, W_PAXIS, W__COUNT   // keep <= 80 (ModelDev::off)
};
static_assert(W__COUNT <= 80, "ModelDev::off too small");
#define WS(f) (WSF(ws) + m.off[f])
#define WSI(f) ((int*)(WSF(ws) + m.off[f]))
// record of contact c: shared-memory tier below cn_k, else the warp's global overflow slot (pointer kept in W_GPTR)
#define CREC(c) cn_rec(m, ws, (c))
#define CRECS(c) (WSF(ws) + m.off[W_CN_REC] + (c) * CR_STRIDE)      /* few-contact paths: every record is in the shared-memory tier (cn_k >= FEW_CONTACTS) */
#define CRECI(c) ((int*)cn_rec(m, ws, (c)))
// contact record (floats): position, frame (normal, t1, t2 rows), distance, friction, body / geom / cell (ints), the
// pyramid's shared D, then per edge: reference acceleration, residual X = J a - aref, V = J search (also: the Hessian weights
// of the contact); frame force; world wrench about the subtree COM (torque, force); world 3 x 3 Hessian weight (xx yy zz xy xz yz).
// There is NO stored contact Jacobian: J v, J^T f and J^T W J are evaluated through the bodies (see "Jacobian-free rows").
enum { CR_POS = 0, CR_FRAME = 3, CR_DIST = 12, CR_MU = 13, CR_BODY = 14, CR_GEOM = 15, CR_CELL = 16, CR_D = 17, CR_AREF = 18, CR_X = 22, CR_V = 26, CR_F = 30, CR_WF = 33, CR_WW = 39, CR_STRIDE = 45 };
enum { RING_SIZE = 64 };
// W_SCAL: [0] ground mu, [1] meaninertia, [2] delay_prob, [3] unused, [4] torsional / [5] rolling friction draw of the env (general path only)
// W_CNT (ints, warp-uniform counters kept in shared memory so they need not travel by reference through the out-of-line
// calls): [0] contacts of the last forward pass, [1] contacts dropped in it, [2] NaN resets, [3] dropped in this control step
// [4] ground contacts of the last pass (they come first; geom-geom contacts follow), [5] bit mask of the bodies that carry a contact
// [6] bit mask of the bodies with ground contacts of their own
enum { CNT_NCON = 0, CNT_DROPPED = 1, CNT_NAN = 2, CNT_DROPPED_STEP = 3, CNT_NCG = 4, CNT_CBMASK = 5, CNT_CBGMASK = 6, CNT_ROWS = 7 };      // [7]: "any constraint row" flag (general path: the row count), handed from stage_smooth to stage_newton

// Per-env arrays in HBM: one row per env, rows contiguous (a warp reads its env's row coalesced).
struct EnvArrays {
  int N;
  float *qpos, *qvel, *warm;
  float *paxis;      // [N][4 * PAXIS_SLOTS] cached separating axes of geom-geom pairs (pair id + 1, axis in the first geom's frame)
  float *body_mass, *invw_dof, *invw_body, *floss, *gmu, *scal, *kp, *kd;
  float *prev_action, *delay_prev, *obs_buffer, *freq_cache, *torque, *info, *last_action;
  int *counters;     // [N][8]: sim_step, has_delay_prev, n_reset, n_obs, n_step, nan_count, need_reset, ncon
  // episode statistics (reporter): [N][NSTAT]
  float *stats;
  // optional debug dumps (NULL when disabled)
  float *dbg_contacts, *dbg_heightmap, *dbg_cfrc, *dbg_sens, *dbg_qacc; int *dbg_hmcell, *dbg_iters;
};
enum { PAXIS_SLOTS = 8 };
enum { CT_SIM_STEP = 0, CT_HAS_DELAY = 1, CT_NRESET = 2, CT_NOBS = 3, CT_NSTEP = 4, CT_NAN = 5, CT_NEED_RESET = 6, CT_NCON = 7 };
// per-env episode statistics accumulated on device (reporter semantics)
enum { ST_STEPS = 0, ST_EPISODES, ST_SUCCESS, ST_TERMINATED, ST_ERR_VX, ST_ERR_VY, ST_ERR_WZ, ST_RMSE, ST_ABS_TORQUE, ST_SQ_TORQUE, ST_MAX_TORQUE, ST_NCON, ST_ITERS, ST_DROPPED, ST_NAN, ST__COUNT = 16 };

// ------------------------------------------------------------------------------------------ small math
DEV void v3copy(float* r, const float* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
DEV void v3add(float* r, const float* a, const float* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
DEV void v3sub(float* r, const float* a, const float* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
DEV void v3scl(float* r, const float* a, float s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
DEV void v3addscl(float* r, const float* a, const float* b, float s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
DEV float v3dot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
DEV void v3cross(float* r, const float* a, const float* b) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
DEV float v3norm(const float* a) { return sqrtf(v3dot(a, a)); }
DEV float v3normalize(float* a) {
  float n = v3norm(a);
  if (n < MINVALF) { a[0] = 1.f; a[1] = a[2] = 0.f; return n; }
  float s = 1.f / n; a[0] *= s; a[1] *= s; a[2] *= s; return n;
}
DEV void m3mulv(float* r, const float* M, const float* v) {
  float x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2], z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
DEV void m3tmulv(float* r, const float* M, const float* v) {
  float x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2], z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
DEV void quat_mul(float* r, const float* a, const float* b) {
  float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  float y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  float z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
DEV void quat_normalize(float* q) {
  float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVALF) { q[0] = 1.f; q[1] = q[2] = q[3] = 0.f; return; }
  float s = 1.f / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
DEV void quat_to_mat(float* M, const float* q) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  M[0] = 1.f - 2.f * (y * y + z * z); M[1] = 2.f * (x * y - z * w); M[2] = 2.f * (x * z + y * w);
  M[3] = 2.f * (x * y + z * w); M[4] = 1.f - 2.f * (x * x + z * z); M[5] = 2.f * (y * z - x * w);
  M[6] = 2.f * (x * z - y * w); M[7] = 2.f * (y * z + x * w); M[8] = 1.f - 2.f * (x * x + y * y);
}
DEV void quat_rot(float* r, const float* q, const float* v) { float M[9]; quat_to_mat(M, q); m3mulv(r, M, v); }
DEV void inert_mul(float* r, const float* I, const float* v) {
  const float* w = v; const float* l = v + 3; const float* mh = I + 6; float ms = I[9];
  float c1[3], c2[3]; v3cross(c1, mh, l); v3cross(c2, mh, w);
  r[0] = I[0] * w[0] + I[3] * w[1] + I[4] * w[2] + c1[0];
  r[1] = I[3] * w[0] + I[1] * w[1] + I[5] * w[2] + c1[1];
  r[2] = I[4] * w[0] + I[5] * w[1] + I[2] * w[2] + c1[2];
  r[3] = ms * l[0] - c2[0]; r[4] = ms * l[1] - c2[1]; r[5] = ms * l[2] - c2[2];
}
DEV void cross_motion(float* r, const float* v, const float* s) {
  float a[3], b[3], c[3]; v3cross(a, v, s); v3cross(b, v, s + 3); v3cross(c, v + 3, s);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
DEV void cross_force(float* r, const float* v, const float* f) {
  float a[3], b[3], c[3]; v3cross(a, v, f); v3cross(b, v + 3, f + 3); v3cross(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

DEV float* cn_rec(const ModelDev& m, WSP ws, int c) {
  if (c < m.cn_k) return WSF(ws) + m.off[W_CN_REC] + c * CR_STRIDE;
  float* g = *(float* const*)(WSF(ws) + m.off[W_GPTR]);
  return g + (size_t)(c - m.cn_k) * CR_STRIDE;
}

// ------------------------------------------------------------------------------------------ RNG
// Philox4x32-10, counter = (global env id, stream, step, block), key = seed
DEV uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef COSIM_HOST_EMU
  return (uint32_t)(((uint64_t)a * b) >> 32);
#else
  return __umulhi(a, b);
#endif
}
DEV_NOINLINE uint32_t philox_draw(const ModelDev& m, uint32_t env, uint32_t stream, uint32_t step, uint32_t idx) {
  uint32_t c0 = env + m.env_offset, c1 = stream, c2 = step, c3 = idx >> 2, k0 = m.seed_lo, k1 = m.seed_hi;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0, h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  uint32_t o[4] = {c0, c1, c2, c3};
  return o[idx & 3];
}
DEV float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
DEV float uni(const ModelDev& m, int env, uint32_t stream, uint32_t step, uint32_t idx) { return u01(philox_draw(m, (uint32_t)env, stream, step, idx)); }

DEV float ndtri(float p) {
#ifdef COSIM_HOST_EMU
  return fast_ndtri(p);
#else
  return normcdfinvf(p);
#endif
}

// ------------------------------------------------------------------------------------------ kinematics
DEV_NOINLINE void kinematics(const ModelDev& m, WSP ws, int lane) { LANE_REFRESH();
  float* qpos = WS(W_QPOS); float* xpos = WS(W_XPOS); float* xquat = WS(W_XQUAT); float* xmat = WS(W_XMAT);
  float* xipos = WS(W_XIPOS); float* xanchor = WS(W_XANCHOR); float* xaxis = WS(W_XAXIS);
  if (lane == 0) {
    xpos[0] = xpos[1] = xpos[2] = 0.f; xquat[0] = 1.f; xquat[1] = xquat[2] = xquat[3] = 0.f;
    xmat[0] = xmat[4] = xmat[8] = 1.f; xmat[1] = xmat[2] = xmat[3] = xmat[5] = xmat[6] = xmat[7] = 0.f;
    xipos[0] = xipos[1] = xipos[2] = 0.f;
  }
  SYNC();
  NOUNROLL for (int l = 1; l < m.nlevels; ++l) {
    NOUNROLL for (int idx = TB(level_start)[l] + lane; idx < TB(level_start)[l + 1]; idx += LANES) {
      const int b = TB(level_body)[idx], j = TB(body_jnt)[b];
      float xp[3], xq[4];
      if (j >= 0 && TB(jnt_type)[j] == 0) {
        const int qa = TB(jnt_qposadr)[j];
        quat_normalize(qpos + qa + 3);
        v3copy(xp, qpos + qa);
        xq[0] = qpos[qa + 3]; xq[1] = qpos[qa + 4]; xq[2] = qpos[qa + 5]; xq[3] = qpos[qa + 6];
        v3copy(xanchor + 3 * j, xp);
        float ax[3] = {LDG(TB(jnt_axis) + 3 * j), LDG(TB(jnt_axis) + 3 * j + 1), LDG(TB(jnt_axis) + 3 * j + 2)};
        quat_rot(xaxis + 3 * j, xq, ax);
      } else {
        const int p = TB(body_parent)[b];
        float bp[3] = {LDG(TB(body_pos) + 3 * b), LDG(TB(body_pos) + 3 * b + 1), LDG(TB(body_pos) + 3 * b + 2)};
        float bq[4] = {LDG(TB(body_quat) + 4 * b), LDG(TB(body_quat) + 4 * b + 1), LDG(TB(body_quat) + 4 * b + 2), LDG(TB(body_quat) + 4 * b + 3)};
        float r[3]; m3mulv(r, xmat + 9 * p, bp); v3add(xp, xpos + 3 * p, r);
        quat_mul(xq, xquat + 4 * p, bq);
        if (j >= 0) {
          float jp[3] = {LDG(TB(jnt_pos) + 3 * j), LDG(TB(jnt_pos) + 3 * j + 1), LDG(TB(jnt_pos) + 3 * j + 2)};
          float ax[3] = {LDG(TB(jnt_axis) + 3 * j), LDG(TB(jnt_axis) + 3 * j + 1), LDG(TB(jnt_axis) + 3 * j + 2)};
          float v[3]; quat_rot(v, xq, jp); v3add(xanchor + 3 * j, xp, v);
          quat_rot(xaxis + 3 * j, xq, ax);
          const int qa = TB(jnt_qposadr)[j];
          float ang = (qpos[qa] - LDG(TB(qpos0) + qa)) * 0.5f, s, c;
          sincosf(ang, &s, &c);
          float ql[4] = {c, ax[0] * s, ax[1] * s, ax[2] * s};
          quat_mul(xq, xq, ql);
          quat_rot(v, xq, jp); v3sub(xp, xanchor + 3 * j, v);
        }
      }
      quat_normalize(xq);
      v3copy(xpos + 3 * b, xp);
      xquat[4 * b] = xq[0]; xquat[4 * b + 1] = xq[1]; xquat[4 * b + 2] = xq[2]; xquat[4 * b + 3] = xq[3];
      quat_to_mat(xmat + 9 * b, xq);
      float ip[3] = {LDG(TB(body_ipos) + 3 * b), LDG(TB(body_ipos) + 3 * b + 1), LDG(TB(body_ipos) + 3 * b + 2)}, r[3];
      m3mulv(r, xmat + 9 * b, ip); v3add(xipos + 3 * b, xp, r);
    }
    SYNC();
  }
  float* gxpos = WS(W_GXPOS); float* gxmat = WS(W_GXMAT);
  FOR_LANE(g, MD(ngeom)) {
    const int b = TB(geom_body)[g];
    float gp[3] = {LDG(TB(geom_pos) + 3 * g), LDG(TB(geom_pos) + 3 * g + 1), LDG(TB(geom_pos) + 3 * g + 2)};
    float gq[4] = {LDG(TB(geom_quat) + 4 * g), LDG(TB(geom_quat) + 4 * g + 1), LDG(TB(geom_quat) + 4 * g + 2), LDG(TB(geom_quat) + 4 * g + 3)};
    float r[3], q[4]; m3mulv(r, xmat + 9 * b, gp); v3add(gxpos + 3 * g, xpos + 3 * b, r);
    quat_mul(q, xquat + 4 * b, gq); quat_to_mat(gxmat + 9 * g, q);
  }
  SYNC();
}

// subtree COM of the (single) tree, spatial inertias about it, motion axes
DEV_NOINLINE void com_pos(const ModelDev& m, WSP ws, int lane) { LANE_REFRESH();
  const int nb = MD(nbody), nv = MD(nv);
  const float* xipos = WS(W_XIPOS); const float* xmat = WS(W_XMAT); const float* bmass = WS(W_BMASS);
  float sx = 0.f, sy = 0.f, sz = 0.f, sm = 0.f;
  NOUNROLL for (int b = 1 + lane; b < nb; b += LANES) { float ms = bmass[b]; sx += ms * xipos[3 * b]; sy += ms * xipos[3 * b + 1]; sz += ms * xipos[3 * b + 2]; sm += ms; }
  sx = wsum(sx); sy = wsum(sy); sz = wsum(sz); sm = wsum(sm);
  const float inv = 1.f / sm;
  const float com[3] = {sx * inv, sy * inv, sz * inv};
  float* scom = WS(W_SCOM);
  if (lane == 0) { scom[0] = com[0]; scom[1] = com[1]; scom[2] = com[2]; scom[3] = sm; }
  float* cinert = WS(W_CINERT);
  NOUNROLL for (int b = 1 + lane; b < nb; b += LANES) {
    const float* Ib = TB(body_inertia) + 6 * b;
    const float I[9] = {LDG(Ib), LDG(Ib + 3), LDG(Ib + 4), LDG(Ib + 3), LDG(Ib + 1), LDG(Ib + 5), LDG(Ib + 4), LDG(Ib + 5), LDG(Ib + 2)};
    const float* R = xmat + 9 * b;
    float RI[9], Iw[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) RI[3 * i + j] = R[3 * i] * I[j] + R[3 * i + 1] * I[3 + j] + R[3 * i + 2] * I[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Iw[3 * i + j] = RI[3 * i] * R[3 * j] + RI[3 * i + 1] * R[3 * j + 1] + RI[3 * i + 2] * R[3 * j + 2];
    float h[3]; v3sub(h, xipos + 3 * b, com);
    const float ms = bmass[b], hh = v3dot(h, h);
    float* ci = cinert + 10 * b;
    ci[0] = Iw[0] + ms * (hh - h[0] * h[0]); ci[1] = Iw[4] + ms * (hh - h[1] * h[1]); ci[2] = Iw[8] + ms * (hh - h[2] * h[2]);
    ci[3] = Iw[1] - ms * h[0] * h[1]; ci[4] = Iw[2] - ms * h[0] * h[2]; ci[5] = Iw[5] - ms * h[1] * h[2];
    ci[6] = ms * h[0]; ci[7] = ms * h[1]; ci[8] = ms * h[2]; ci[9] = ms;
  }
  float* cdof = WS(W_CDOF); const float* xanchor = WS(W_XANCHOR); const float* xaxis = WS(W_XAXIS);
  FOR_LANE(k, nv) {
    const int j = TB(dof_jnt)[k], b = TB(dof_body)[k];
    float off[3]; v3sub(off, com, xanchor + 3 * j);
    float* cd = cdof + 6 * k;
    if (TB(jnt_type)[j] == 0) {
      const int r = k - TB(jnt_dofadr)[j];
      if (r < 3) { cd[0] = cd[1] = cd[2] = 0.f; cd[3] = (r == 0); cd[4] = (r == 1); cd[5] = (r == 2); }
      else { float ax[3] = {xmat[9 * b + r - 3], xmat[9 * b + 3 + r - 3], xmat[9 * b + 6 + r - 3]}; v3copy(cd, ax); v3cross(cd + 3, ax, off); }
    } else { v3copy(cd, xaxis + 3 * j); v3cross(cd + 3, xaxis + 3 * j, off); }
  }
  SYNC();
}

DEV_NOINLINE void crb(const ModelDev& m, WSP ws, int lane) { LANE_REFRESH();
  const int nb = MD(nbody), nv = MD(nv);
  const float* cinert = WS(W_CINERT); float* crbI = WS(W_CRB); const float* cdof = WS(W_CDOF);
  float* buf = WS(W_BUF); float* M = WS(W_M);
  NOUNROLL for (int b = 1 + lane; b < nb; b += LANES) {   // composite = sum over the contiguous DFS subtree, fixed order
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = cinert[10 * b + i];
    const int end = b + TB(body_subsize)[b];
    NOUNROLL for (int c = b + 1; c < end; ++c)
#pragma unroll
      for (int i = 0; i < 10; ++i) acc[i] += cinert[10 * c + i];
#pragma unroll
    for (int i = 0; i < 10; ++i) crbI[10 * b + i] = acc[i];
  }
  FOR_LANE(i, nv * nv) M[i] = 0.f;
  SYNC();
  FOR_LANE(i, nv) inert_mul(buf + 6 * i, crbI + 10 * TB(dof_body)[i], cdof + 6 * i);
  SYNC();
  FOR_LANE(p, m.nmpair) {
    const int i = TB(mpair_i)[p], j = TB(mpair_j)[p];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += cdof[6 * j + k] * buf[6 * i + k];
    if (i == j) s += LDG(TB(dof_armature) + i);
    M[i * nv + j] = s; M[j * nv + i] = s;
  }
  SYNC();
}

// Cholesky-type factorization A = U U^T of the n x n SPD matrix in A (lower triangle used), eliminating the LAST dof first
// [upstream mj_factorM: L^T D L from the leaves of the kinematic tree towards the root].  With dofs numbered parent-before-child
// the mass matrix -- and every Hessian M + J^T D J whose rows act on single bodies (ground contacts, limits, friction loss) --
// couples a dof only with its ancestors, and in this elimination order no fill-in appears: column j updates only the pairs
// (i, k) of ancestors of j (`sparse` = 1: per-column pair lists m.ctab / TB(coff); humanoid 1018 pair updates in 48 passes
// instead of 4060 in 141, w4 555 / 29 instead of 1771 / 67, flamingo_p_v3 295 / 17 instead of 455 / 22).  Rows that couple two
// branches (geom-geom contacts, connect constraints) need the dense variant (`sparse` = 0): all pairs i, k < j, which are the
// first j (j + 1) / 2 entries of TB(tri).  The factor is kept UNSCALED: after the call A[j][i] (i < j) holds U_ij * U_jj and
// invd[j] = 1 / U_jj; chol_solve folds the scaling into its pivots (no column-scaling pass, one warp barrier per column).
// Entries outside the tree pattern are never touched by the sparse variant and must be zero (chol_solve reads whole rows).
DEV_NOINLINE void chol_factor(const ModelDev& m, float* A, float* invd, int n, int lane, int sparse) { LANE_REFRESH();
#ifdef COSIM_CHOL_ALLDENSE
  sparse = 0;        // A / B builds: dense elimination everywhere
#endif
  // table pointers and the column offsets' base in registers: the stores to A below would otherwise force their re-load per column
  const int* const tri = TB(tri); const int* const coff = TB(coff); const int* const ctab = m.ctab; const uint16_t* const t16 = m.ctab16 ? SHP(m.ctab16) : nullptr;
  NOUNROLL for (int j = n - 1; j >= 0; --j) {
    const float* row = A + j * n;
    const float d = fmaxf(row[j], 1e-30f);
#ifdef COSIM_HOST_EMU
    const float inv = 1.f / sqrtf(d);
#else
    const float inv = rsqrtf(d);
#endif
    const float inv2 = inv * inv;
    if (lane == 0) invd[j] = inv;
    if (sparse) {
      const int o0 = coff[j], o1 = coff[j + 1];
      if (t16) {
        NOUNROLL for (int idx = o0 + lane; idx < o1; idx += LANES) {
          const int t = (int)t16[idx], i = t >> 8, k = t & 255;
          A[i * n + k] -= row[i] * row[k] * inv2;
        }
      } else {
        NOUNROLL for (int idx = o0 + lane; idx < o1; idx += LANES) {
          const int t = LDGB(ctab + idx), i = t >> 8, k = t & 255;
          A[i * n + k] -= row[i] * row[k] * inv2;
        }
      }
    } else {
      const int T = (j * (j + 1)) >> 1;
      NOUNROLL for (int idx = lane; idx < T; idx += LANES) {
        const int t = tri[idx], i = t >> 8, k = t & 255;
        A[i * n + k] -= row[i] * row[k] * inv2;
      }
    }
    SYNC();
  }
}
// solves U U^T x = b with the unscaled factor of chol_factor (either variant).  b is destroyed, tmp is scratch, result in out
DEV_NOINLINE void chol_solve(const float* A, const float* invd, float* b, float* tmp, float* out, int n, int lane) { LANE_REFRESH();
#ifdef COSIM_HOST_EMU
  for (int j = n - 1; j >= 0; --j) {            // U z = b, column-oriented from the last dof
    const float zj = b[j] * invd[j], s = zj * invd[j];
    for (int i = lane; i < j; i += LANES) b[i] -= A[j * n + i] * s;
    if (lane == 0) tmp[j] = zj;
    SYNC();
  }
  for (int i = 0; i < n; ++i) {                 // U^T x = z, column-oriented from the root
    const float xi = tmp[i] * invd[i];
    for (int j = i + 1 + lane; j < n; j += LANES) tmp[j] -= A[j * n + i] * invd[j] * xi;
    if (lane == 0) out[i] = xi;
    SYNC();
  }
#else
  // n <= 32: lane k keeps element k in a register, the pivot travels by shuffle (same arithmetic as above)
  float bk = lane < n ? b[lane] : 0.f;
  const float dk = lane < n ? invd[lane] : 0.f;
  NOUNROLL for (int j = n - 1; j >= 0; --j) {
    const float dj = __shfl_sync(0xffffffffu, dk, j);
    const float zj = __shfl_sync(0xffffffffu, bk, j) * dj;
    if (lane < j) bk -= A[j * n + lane] * (zj * dj);
    if (lane == j) bk = zj;
  }
  NOUNROLL for (int i = 0; i < n; ++i) {
    const float xi = __shfl_sync(0xffffffffu, bk, i) * __shfl_sync(0xffffffffu, dk, i);
    if (lane > i && lane < n) bk -= A[lane * n + i] * dk * xi;
    if (lane == i) bk = xi;
  }
  if (lane < n) out[lane] = bk;
  SYNC();
  (void)tmp;
#endif
}

// ------------------------------------------------------------------------------------------ velocity / bias
DEV_NOINLINE void com_vel(const ModelDev& m, WSP ws, int lane) { LANE_REFRESH();
  const int nv = MD(nv);
  const float* cdof = WS(W_CDOF); float* cdd = WS(W_CDOFDOT); float* cvel = WS(W_CVEL); const float* qvel = WS(W_QVEL);
  FOR_LANE(i, 6) cvel[i] = 0.f;
  SYNC();
  NOUNROLL for (int l = 1; l < m.nlevels; ++l) {
    NOUNROLL for (int idx = TB(level_start)[l] + lane; idx < TB(level_start)[l + 1]; idx += LANES) {
      const int b = TB(level_body)[idx], p = TB(body_parent)[b], j = TB(body_jnt)[b];
      float v[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) v[i] = cvel[6 * p + i];
      if (j >= 0) {
        const int da = TB(jnt_dofadr)[j];
        if (TB(jnt_type)[j] == 0) {
          for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int i = 0; i < 6; ++i) { cdd[6 * (da + k) + i] = 0.f; v[i] += cdof[6 * (da + k) + i] * qvel[da + k]; }
          }
          for (int k = 3; k < 6; ++k) cross_motion(cdd + 6 * (da + k), v, cdof + 6 * (da + k));
          for (int k = 3; k < 6; ++k)
#pragma unroll
            for (int i = 0; i < 6; ++i) v[i] += cdof[6 * (da + k) + i] * qvel[da + k];
        } else {
          cross_motion(cdd + 6 * da, v, cdof + 6 * da);
#pragma unroll
          for (int i = 0; i < 6; ++i) v[i] += cdof[6 * da + i] * qvel[da];
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) cvel[6 * b + i] = v[i];
    }
    SYNC();
  }
  (void)nv;
}

// qfrc_bias into `out` (length nv)
DEV_NOINLINE void rne_bias(const ModelDev& m, WSP ws, float* out, int lane) { LANE_REFRESH();
  const int nb = MD(nbody), nv = MD(nv);
  const float* cdof = WS(W_CDOF); const float* cdd = WS(W_CDOFDOT); const float* cvel = WS(W_CVEL);
  float* cacc = WS(W_CACC); float* cfrc = WS(W_CFRC); const float* cinert = WS(W_CINERT); const float* qvel = WS(W_QVEL);
  if (lane == 0) { cacc[0] = cacc[1] = cacc[2] = 0.f; cacc[3] = -MO(gx); cacc[4] = -MO(gy); cacc[5] = -MO(gz); }
  SYNC();
  NOUNROLL for (int l = 1; l < m.nlevels; ++l) {
    NOUNROLL for (int idx = TB(level_start)[l] + lane; idx < TB(level_start)[l + 1]; idx += LANES) {
      const int b = TB(level_body)[idx], p = TB(body_parent)[b];
      float a[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) a[i] = cacc[6 * p + i];
      const int d0 = TB(body_dofadr)[b], d1 = d0 + TB(body_dofnum)[b];
      NOUNROLL for (int k = d0; k < d1; ++k)
#pragma unroll
        for (int i = 0; i < 6; ++i) a[i] += cdd[6 * k + i] * qvel[k];
#pragma unroll
      for (int i = 0; i < 6; ++i) cacc[6 * b + i] = a[i];
      float Iv[6], t1[6], Ia[6];
      inert_mul(Iv, cinert + 10 * b, cvel + 6 * b);
      cross_force(t1, cvel + 6 * b, Iv);
      inert_mul(Ia, cinert + 10 * b, a);
#pragma unroll
      for (int i = 0; i < 6; ++i) cfrc[6 * b + i] = Ia[i] + t1[i];
    }
    SYNC();
  }
  // subtree sums (deterministic order) folded directly into the dof projection
  FOR_LANE(k, nv) {
    const int b = TB(dof_body)[k], end = b + TB(body_subsize)[b];
    float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    NOUNROLL for (int c = b; c < end; ++c)
#pragma unroll
      for (int i = 0; i < 6; ++i) f[i] += cfrc[6 * c + i];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += cdof[6 * k + i] * f[i];
    out[k] = s;
  }
  SYNC();
  (void)nb;
}

// ------------------------------------------------------------------------------------------ collision
#ifdef COSIM_HOST_EMU
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
#endif
// direction -> cube-map bucket; must match cosim_b200/model.py:support_bucket (float32 arithmetic)
DEV int support_bucket(const float* d) {
  const int G = 8;
  const float a0 = fabsf(d[0]), a1 = fabsf(d[1]), a2 = fabsf(d[2]);
  const int ax = (a0 >= a1 && a0 >= a2) ? 0 : (a1 >= a2 ? 1 : 2);
  const float dm = ax == 0 ? d[0] : (ax == 1 ? d[1] : d[2]);
  const float du = ax == 0 ? d[1] : (ax == 1 ? d[2] : d[0]);
  const float dv = ax == 0 ? d[2] : (ax == 1 ? d[0] : d[1]);
  const int face = 2 * ax + (dm < 0.f ? 1 : 0);
  const float inv = 1.0f / fmaxf(fabsf(dm), 1e-30f);
  const float u = du * inv, v = dv * inv;
  const int iu = imin(G - 1, imax(0, (int)floorf((u + 1.0f) * (0.5f * G))));
  const int iv = imin(G - 1, imax(0, (int)floorf((v + 1.0f) * (0.5f * G))));
  return (face * G + iu) * G + iv;
}
struct GeomW { int type; const float* pos; const float* mat; float size[3]; const float* verts; int nvert; float center[3]; const int* sup_off; const float4* sup_cand; };

DEV GeomW make_geom(const ModelDev& m, WSP ws, int g) {
  GeomW G; G.type = TB(geom_type)[g]; G.pos = WS(W_GXPOS) + 3 * g; G.mat = WS(W_GXMAT) + 9 * g;
  G.size[0] = LDG(TB(geom_size) + 3 * g); G.size[1] = LDG(TB(geom_size) + 3 * g + 1); G.size[2] = LDG(TB(geom_size) + 3 * g + 2);
  G.verts = m.hull_verts + 3 * TB(geom_vadr)[g]; G.nvert = TB(geom_vnum)[g];
  { const int sa = TB(geom_supadr)[g]; G.sup_off = sa >= 0 ? m.sup_off + sa : nullptr; G.sup_cand = m.sup_cand; }
  float c[3] = {LDG(TB(geom_center) + 3 * g), LDG(TB(geom_center) + 3 * g + 1), LDG(TB(geom_center) + 3 * g + 2)}, r[3];
  m3mulv(r, G.mat, c); v3add(G.center, G.pos, r);
  return G;
}
// libccd's degenerate-case guards use CCD_EPS = DBL_EPSILON (MuJoCo builds it in double).  They are
// ABSOLUTE thresholds, so the fp32 engine keeps the double value: with FLT_EPSILON the guards fire on
// ordinary millimetre-scale portal triangles and the contact normal degenerates.
#define CCD_EPS 2.220446e-16f
DEV bool f_is_zero(float x) { return fabsf(x) < CCD_EPS; }
DEV bool f_eq(float a, float b) {
  float ab = fabsf(a - b);
  if (ab < CCD_EPS) return true;
  float aa = fabsf(a), bb = fabsf(b);
  return ab < CCD_EPS * (bb > aa ? bb : aa);
}
DEV bool v3eq0(const float* a) { return f_eq(a[0], 0.f) && f_eq(a[1], 0.f) && f_eq(a[2], 0.f); }
DEV float seg_dist2(const float* A, const float* B, float* wit) {
  float dd[3]; v3sub(dd, B, A);
  float tt = -v3dot(A, dd) / v3dot(dd, dd);
  if (tt < 0.f || f_is_zero(tt)) { v3copy(wit, A); return v3dot(A, A); }
  if (tt > 1.f || f_eq(tt, 1.f)) { v3copy(wit, B); return v3dot(B, B); }
  float w3[3]; v3addscl(w3, A, dd, tt); v3copy(wit, w3); return v3dot(w3, w3);
}
DEV_NOINLINE float point_tri_dist2(const float* x0, const float* B, const float* C, float* witness) {
  float d1[3], d2[3]; v3sub(d1, B, x0); v3sub(d2, C, x0);
  float v = v3dot(d1, d1), w = v3dot(d2, d2), p = v3dot(x0, d1), q = v3dot(x0, d2), r = v3dot(d1, d2);
  float div = w * v - r * r, s, t = 0.f, dist;
  if (f_is_zero(div)) s = -1.f; else { s = (q * r - w * p) / div; t = (-s * r - q) / w; }
  if ((f_is_zero(s) || s > 0.f) && (f_eq(s, 1.f) || s < 1.f) && (f_is_zero(t) || t > 0.f) && (f_eq(t, 1.f) || t < 1.f) && (f_eq(t + s, 1.f) || t + s < 1.f)) {
    float wv[3]; v3addscl(wv, x0, d1, s); v3addscl(wv, wv, d2, t); v3copy(witness, wv); dist = v3dot(wv, wv);
  } else {
    float w2[3]; dist = seg_dist2(x0, B, witness);
    float d2v = seg_dist2(x0, C, w2); if (d2v < dist) { dist = d2v; v3copy(witness, w2); }
    d2v = seg_dist2(B, C, w2); if (d2v < dist) { dist = d2v; v3copy(witness, w2); }
  }
  return dist;
}
DEV void make_frame(float* frame) {
  float* n = frame; float* t1 = frame + 3; float* t2 = frame + 6;
  v3normalize(n);
  float tmp[3] = {0.f, 0.f, 0.f};
  if (n[1] < 0.5f && n[1] > -0.5f) tmp[1] = 1.f; else tmp[2] = 1.f;
  float dn = v3dot(n, tmp);
  v3addscl(t1, tmp, n, -dn); v3normalize(t1);
  v3cross(t2, n, t1);
}
// write the geometric part of contact record `slot` (one lane)
DEV void write_contact(const ModelDev& m, WSP ws, int slot, const float* pos, const float* normal, float dist, float mu, int body, int g, int cell) {
  float* r = CREC(slot);
  v3copy(r + CR_POS, pos); v3copy(r + CR_FRAME, normal); make_frame(r + CR_FRAME);
  r[CR_DIST] = dist; r[CR_MU] = mu;
  ((int*)r)[CR_BODY] = body; ((int*)r)[CR_GEOM] = g; ((int*)r)[CR_CELL] = cell;
}
// all lanes call with identical arguments; lane 0 writes.  ncon is warp-uniform (register) state
DEV_NOINLINE void add_contact(const ModelDev& m, WSP ws, int& ncon, int& dropped, const float* pos, const float* normal, float dist, int g, int cell, int lane) { LANE_REFRESH();
  if (ncon >= MD(ncon_max)) { ++dropped; return; }
  if (lane == 0) write_contact(m, ws, ncon, pos, normal, dist, fmaxf(WS(W_SCAL)[0], WS(W_GMU)[g]), TB(geom_body)[g], g, cell);
  ++ncon;
}

// ------------------------------------------------------------------------------------------ lane-parallel hfield collision
// mjc_ConvexHField restated (oracle/oracle.hpp collide_hfield is the serial version), different schedule: stage 1 gives every geom a lane
// (bounding tests, AABB from 6 support queries, sub-grid), stage 2 enumerates the (geom, row, col, triangle) prisms of
// the whole env in the reference order and gives every prism a lane; each lane runs its own scalar MPR query.  Contacts
// are appended in task order (ballot prefix), so order, the 50-per-geom cap and the ncon_max cap match the serial loop.
struct PV { float x, y, z; int pi; };     // Minkowski-difference vertex + index of the prism vertex it came from
struct F3 { float x, y, z; };
// Hull support query.  Everything arrives in registers (scalars) and the geom's data is read from shared memory
// (pose in the workspace, tables in the arena): no by-reference structs, hence no local-memory traffic on this hot path.
// grp = sub | (gsize << 8): the lanes [sub = 0 .. gsize) of one group (mask gmask) work on the same query and split the
// candidate scan.  (ox, oy) = origin of the local frame the query runs in.  Same arithmetic / tie-breaks as the oracle's
// serial scan over all hull vertices.
DEV_NOINLINE F3 support_lane(const ModelDev& m, WSP ws, int g, int grp, unsigned gmask, float ox, float oy, float dx, float dy, float dz) {
  const float* M = WS(W_GXMAT) + 9 * g; const float* pos = WS(W_GXPOS) + 3 * g;
  const int type = TB(geom_type)[g];
  const float l0 = M[0] * dx + M[3] * dy + M[6] * dz, l1 = M[1] * dx + M[4] * dy + M[7] * dz, l2 = M[2] * dx + M[5] * dy + M[8] * dz;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (type == GEOM_SPHERE) { const float r = LDG(TB(geom_size) + 3 * g); px = l0 * r; py = l1 * r; pz = l2 * r; }
  else if (type == GEOM_CYLINDER) {
    const float r = LDG(TB(geom_size) + 3 * g), hh = LDG(TB(geom_size) + 3 * g + 1);
    const float n = sqrtf(l0 * l0 + l1 * l1);
    if (n > MINVALF) { px = l0 / n * r; py = l1 / n * r; }
    pz = (l2 > 0.f ? 1.f : (l2 < 0.f ? -1.f : 0.f)) * hh;
  } else if (type == GEOM_BOX) {
    px = (l0 > 0.f ? 1.f : -1.f) * LDG(TB(geom_size) + 3 * g); py = (l1 > 0.f ? 1.f : -1.f) * LDG(TB(geom_size) + 3 * g + 1); pz = (l2 > 0.f ? 1.f : -1.f) * LDG(TB(geom_size) + 3 * g + 2);
  } else if (TB(geom_supadr)[g] >= 0) {
    const int* sup_off = m.sup_off + TB(geom_supadr)[g];
    const float ld[3] = {l0, l1, l2};
    const int bucket = support_bucket(ld);
    int o0, o1;
    if (m.sup_off16) { const uint16_t* t16 = SHP(m.sup_off16) + TB(geom_supadr)[g]; const int b0 = TB(geom_supbase)[g]; o0 = b0 + t16[bucket]; o1 = b0 + t16[bucket + 1]; }
    else { o0 = LDGB(sup_off + bucket); o1 = LDGB(sup_off + bucket + 1); }
    float bv = -INFINITY; int bk = 0x7fffffff;
    const int gs = grp >> 8, sub = grp & 255;
    NOUNROLL for (int k = o0 + sub; k < o1; k += 4 * gs) {    // four independent 16-byte loads in flight per lane
      const int k1 = imin(k + gs, o1 - 1), k2 = imin(k + 2 * gs, o1 - 1), k3 = imin(k + 3 * gs, o1 - 1);
#ifdef COSIM_HOST_EMU
      const float4 c0 = m.sup_cand[k], c1 = m.sup_cand[k1], c2 = m.sup_cand[k2], c3 = m.sup_cand[k3];
#else
      const float4 c0 = __ldg(m.sup_cand + k), c1 = __ldg(m.sup_cand + k1), c2 = __ldg(m.sup_cand + k2), c3 = __ldg(m.sup_cand + k3);
#endif
      const float v0 = c0.x * l0 + c0.y * l1 + c0.z * l2, v1 = c1.x * l0 + c1.y * l1 + c1.z * l2;
      const float v2 = c2.x * l0 + c2.y * l1 + c2.z * l2, v3 = c3.x * l0 + c3.y * l1 + c3.z * l2;
      // candidates are sorted by vertex index: on equal support the earliest position wins, as in a serial scan
      if (v0 > bv || (v0 == bv && k < bk)) { bv = v0; bk = k; px = c0.x; py = c0.y; pz = c0.z; }
      if (v1 > bv || (v1 == bv && k1 < bk)) { bv = v1; bk = k1; px = c1.x; py = c1.y; pz = c1.z; }
      if (v2 > bv || (v2 == bv && k2 < bk)) { bv = v2; bk = k2; px = c2.x; py = c2.y; pz = c2.z; }
      if (v3 > bv || (v3 == bv && k3 < bk)) { bv = v3; bk = k3; px = c3.x; py = c3.y; pz = c3.z; }
    }
#ifndef COSIM_HOST_EMU
    if (gs > 1) {        // reduce over the lanes of the group (largest support, ties: earliest candidate), then fetch the winner's coordinates
      const float bvc = bv == 0.f ? 0.f : bv;                                   // -0 and +0 compare equal, as in the serial scan
      unsigned key = __float_as_uint(bvc); key = (key & 0x80000000u) ? ~key : (key | 0x80000000u);      // order-preserving map float -> uint
      const unsigned kmax = __reduce_max_sync(gmask, key);
      const int kmin = __reduce_min_sync(gmask, key == kmax ? bk : 0x7fffffff);
      const int src = __ffs(__ballot_sync(gmask, key == kmax && bk == kmin)) - 1;
      px = __shfl_sync(gmask, px, src); py = __shfl_sync(gmask, py, src); pz = __shfl_sync(gmask, pz, src);
    }
#endif
  } else {
    const float* verts = m.hull_verts + 3 * TB(geom_vadr)[g]; const int nvert = TB(geom_vnum)[g];
    float bv = -INFINITY;
    NOUNROLL for (int i = 0; i < nvert; ++i) {
      const float vx = LDGB(verts + 3 * i), vy = LDGB(verts + 3 * i + 1), vz = LDGB(verts + 3 * i + 2);
      const float v = vx * l0 + vy * l1 + vz * l2;
      if (v > bv) { bv = v; px = vx; py = vy; pz = vz; }
    }
  }
  F3 out;      // (pos - origin): exact subtraction of nearby numbers
  out.x = (pos[0] - ox) + (M[0] * px + M[1] * py + M[2] * pz);
  out.y = (pos[1] - oy) + (M[3] * px + M[4] * py + M[5] * pz);
  out.z = pos[2] + (M[6] * px + M[7] * py + M[8] * pz);
  return out;
}
// prism in registers: three columns (x, y, top z), bottoms at -base.  Vertex order as in the strip walk: 0..2 bottoms, 3..5 tops
struct PrismL { float x[3], y[3], z[3], base; };
DEV void prism_vertex(const PrismL& P, int i, float* v) {
  const int c = i >= 3 ? i - 3 : i;
  v[0] = c == 0 ? P.x[0] : (c == 1 ? P.x[1] : P.x[2]);
  v[1] = c == 0 ? P.y[0] : (c == 1 ? P.y[1] : P.y[2]);
  v[2] = i >= 3 ? (c == 0 ? P.z[0] : (c == 1 ? P.z[1] : P.z[2])) : -P.base;
}
#define GQ_PARAMS const ModelDev& m, WSP ws, int g, int grp, unsigned gmask, float ox, float oy
#define GQ_ARGS m, ws, g, grp, gmask, ox, oy
// First MPR object ("shape A"): a terrain prism (mjc_ConvexHField) or a convex geom (mjc_Convex).  A shape provides its
// portal-vertex type PV (Minkowski-difference point x, y, z + what it needs to recover the witness on A), the Minkowski
// support, the interior-point vertex p0 and the witness lookup.  The second object is always geom g (GQ_PARAMS).
struct PrismA {
  typedef ::PV PV;
  PrismL P;
  DEVM PV mink(GQ_PARAMS, float dx, float dy, float dz) const {
    int best = 0; float bv = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = i >= 3 ? i - 3 : i;
      const float vx = P.x[c], vy = P.y[c], vz = i >= 3 ? P.z[c] : -P.base;
      const float v = vx * dx + vy * dy + vz * dz;
      if (i == 0 || v > bv) { bv = v; best = i; bx = vx; by = vy; bz = vz; }
    }
    const F3 s2 = support_lane(GQ_ARGS, -dx, -dy, -dz);
    PV r; r.x = bx - s2.x; r.y = by - s2.y; r.z = bz - s2.z; r.pi = best;
    return r;
  }
  DEVM void center(float* c1) const {
    c1[0] = c1[1] = c1[2] = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) { float v[3]; prism_vertex(P, i, v); v3add(c1, c1, v); }
    v3scl(c1, c1, 1.f / 6.f);
  }
  DEVM PV make_p0(const float* c1, const float* gcenter) const { PV p; p.x = c1[0] - gcenter[0]; p.y = c1[1] - gcenter[1]; p.z = c1[2] - gcenter[2]; p.pi = -1; return p; }
  // witness on the prism: exact vertex (or the centre for p0)
  DEVM void witness(const PV& p, const float* c1, float* v1) const { if (p.pi < 0) v3copy(v1, c1); else prism_vertex(P, p.pi, v1); }
};
struct PVG { float x, y, z, wx, wy, wz; };     // Minkowski-difference vertex + its witness point on the first geom
struct GeomA {
  typedef PVG PV;
  int g1; float c[3];                          // first geom and its centre in the local frame
  DEVM PV mink(GQ_PARAMS, float dx, float dy, float dz) const {
    const F3 s1 = support_lane(m, ws, g1, grp, gmask, ox, oy, dx, dy, dz);
    const F3 s2 = support_lane(GQ_ARGS, -dx, -dy, -dz);
    PV r; r.x = s1.x - s2.x; r.y = s1.y - s2.y; r.z = s1.z - s2.z; r.wx = s1.x; r.wy = s1.y; r.wz = s1.z;
    return r;
  }
  DEVM void center(float* c1) const { v3copy(c1, c); }
  DEVM PV make_p0(const float* c1, const float* gcenter) const { PV p; p.x = c1[0] - gcenter[0]; p.y = c1[1] - gcenter[1]; p.z = c1[2] - gcenter[2]; p.wx = c1[0]; p.wy = c1[1]; p.wz = c1[2]; return p; }
  DEVM void witness(const PV& p, const float*, float* v1) const { v1[0] = p.wx; v1[1] = p.wy; v1[2] = p.wz; }
};
template <class V> DEV float pv_dot(const V& a, float x, float y, float z) { return a.x * x + a.y * y + a.z * z; }
template <class V> DEV void pv_portal_dir(const V& p1, const V& p2, const V& p3, float* dir) {
  float a[3] = {p2.x - p1.x, p2.y - p1.y, p2.z - p1.z}, b[3] = {p3.x - p1.x, p3.y - p1.y, p3.z - p1.z};
  v3cross(dir, a, b); v3normalize(dir);
}
template <class V> DEV bool pv_reach_tol(const V& p1, const V& p2, const V& p3, const V& v4, const float* dir, float tol) {
  const float dv1 = pv_dot(p1, dir[0], dir[1], dir[2]), dv2 = pv_dot(p2, dir[0], dir[1], dir[2]), dv3 = pv_dot(p3, dir[0], dir[1], dir[2]), dv4 = pv_dot(v4, dir[0], dir[1], dir[2]);
  const float dm = fminf(dv4 - dv1, fminf(dv4 - dv2, dv4 - dv3));
  return f_eq(dm, tol) || dm < tol;
}
template <class V> DEV void pv_expand(const V& p0, V& p1, V& p2, V& p3, const V& v4) {
  const float a[3] = {v4.x, v4.y, v4.z}, b[3] = {p0.x, p0.y, p0.z};
  float c[3]; v3cross(c, a, b);
  float dot = pv_dot(p1, c[0], c[1], c[2]);
  if (dot > 0.f) { dot = pv_dot(p2, c[0], c[1], c[2]); if (dot > 0.f) p1 = v4; else p3 = v4; }
  else { dot = pv_dot(p3, c[0], c[1], c[2]); if (dot > 0.f) p2 = v4; else p1 = v4; }
}
// MPR (XenoCollide) penetration query, restating libccd ccdMPRPenetration as driven by mjc_ConvexHField / mjc_Convex
// (oracle/oracle.hpp mpr_core is the readable fp64 version); 0 = hit.  Run by one lane, or by a group of lanes in lock
// step that share the hull support scans (grp = sub | gsize << 8).
// The control flow is a state machine with ONE Minkowski-support call per trip: the lanes of a warp work on different
// queries that sit in different stages of the algorithm (first / second vertex, portal discovery, portal refinement,
// penetration depth), and with a call site per stage every stage's hull scan ran on its own handful of lanes (7 of 32
// active in support_lane, profiles/r02_w4_by_function_before.txt).  Same steps in the same order with the same arithmetic
// as the staged version; the result of a hit is computed after the loop, once for all lanes.
template <class A> DEV int mpr_lane_sm(const A& P, GQ_PARAMS, const float* gcenter, float* depth, float* dir_out, float* pos) {
  typedef typename A::PV PV;
  enum { S_P1 = 0, S_P2, S_DISCOVER, S_REFINE, S_PENETRATE, S_DONE, S_TOUCH, S_HIT };
  const float tol = MO(ccd_tolerance); const int maxit = MD(ccd_iterations);
  float c1[3]; P.center(c1);
  PV p0 = P.make_p0(c1, gcenter), p1 = p0, p2 = p0, p3 = p0;
  if (f_eq(p0.x, 0.f) && f_eq(p0.y, 0.f) && f_eq(p0.z, 0.f)) p0.x += CCD_EPS * 10.f;
  float dir[3] = {-p0.x, -p0.y, -p0.z}; v3normalize(dir);
  int state = S_P1, ret = -2, guard = 0, it = 0;
  NOUNROLL while (state < S_DONE) {
    // ---- what the stage does before its support query
    if (state == S_DISCOVER) { if (++guard > 100) { ret = -2; state = S_DONE; } }
    else if (state == S_REFINE) {
      if (++guard > 1000) { ret = -2; state = S_DONE; }
      else {
        pv_portal_dir(p1, p2, p3, dir);
        const float d1 = pv_dot(p1, dir[0], dir[1], dir[2]);
        if (f_is_zero(d1) || d1 > 0.f) state = S_PENETRATE;      // the origin is inside the portal; the depth stage starts from the same direction
      }
    } else if (state == S_PENETRATE) pv_portal_dir(p1, p2, p3, dir);
    if (state >= S_DONE) break;
    const PV v = P.mink(GQ_ARGS, dir[0], dir[1], dir[2]);
    const float dot = pv_dot(v, dir[0], dir[1], dir[2]);
    if (state == S_P1) {
      p1 = v;
      if (f_is_zero(dot) || dot < 0.f) { v3copy(dir_out, dir); ret = -1; state = S_DONE; }      // -1: `dir_out` separates the two objects
      else {
        { const float a[3] = {p0.x, p0.y, p0.z}, b[3] = {p1.x, p1.y, p1.z}; v3cross(dir, a, b); }
        if (f_is_zero(v3dot(dir, dir))) state = S_TOUCH;
        else { v3normalize(dir); state = S_P2; }
      }
    } else if (state == S_P2) {
      p2 = v;
      if (f_is_zero(dot) || dot < 0.f) { v3copy(dir_out, dir); ret = -1; state = S_DONE; }
      else {
        { float va[3] = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, vb[3] = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z}; v3cross(dir, va, vb); v3normalize(dir); }
        if (pv_dot(p0, dir[0], dir[1], dir[2]) > 0.f) { const PV t = p1; p1 = p2; p2 = t; dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2]; }
        state = S_DISCOVER; guard = 0;
      }
    } else if (state == S_DISCOVER) {
      p3 = v;
      if (f_is_zero(dot) || dot < 0.f) { v3copy(dir_out, dir); ret = -1; state = S_DONE; }
      else {
        int cont = 0; float d2;
        { const float a[3] = {p1.x, p1.y, p1.z}, b[3] = {p3.x, p3.y, p3.z}; float va[3]; v3cross(va, a, b); d2 = pv_dot(p0, va[0], va[1], va[2]); }
        if (d2 < 0.f && !f_is_zero(d2)) { p2 = p3; cont = 1; }
        if (!cont) {
          const float a[3] = {p3.x, p3.y, p3.z}, b[3] = {p2.x, p2.y, p2.z}; float va[3]; v3cross(va, a, b); d2 = pv_dot(p0, va[0], va[1], va[2]);
          if (d2 < 0.f && !f_is_zero(d2)) { p1 = p3; cont = 1; }
        }
        if (cont) { float va[3] = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, vb[3] = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z}; v3cross(dir, va, vb); v3normalize(dir); }
        else { state = S_REFINE; guard = 0; }
      }
    } else if (state == S_REFINE) {
      if (!(f_is_zero(dot) || dot > 0.f)) { v3copy(dir_out, dir); ret = -1; state = S_DONE; }
      else if (pv_reach_tol(p1, p2, p3, v, dir, tol)) { ret = -2; state = S_DONE; }            // no intersection, but `dir` is not a separating axis
      else pv_expand(p0, p1, p2, p3, v);
    } else {        // S_PENETRATE
      if (pv_reach_tol(p1, p2, p3, v, dir, tol) || it > maxit) state = S_HIT;
      else { pv_expand(p0, p1, p2, p3, v); ++it; }
    }
  }
  if (state == S_TOUCH) {
    float v1[3]; P.witness(p1, c1, v1);
    for (int k = 0; k < 3; ++k) { const float pk = k == 0 ? p1.x : (k == 1 ? p1.y : p1.z); pos[k] = (v1[k] + (v1[k] - pk)) * 0.5f; }
    if (f_eq(p1.x, 0.f) && f_eq(p1.y, 0.f) && f_eq(p1.z, 0.f)) { *depth = 0.f; dir_out[0] = dir_out[1] = dir_out[2] = 0.f; return 0; }
    float pv[3] = {p1.x, p1.y, p1.z};
    *depth = v3norm(pv); v3copy(dir_out, pv); v3normalize(dir_out); return 0;
  }
  if (state == S_HIT) {
    const float a[3] = {p1.x, p1.y, p1.z}, b[3] = {p2.x, p2.y, p2.z}, c[3] = {p3.x, p3.y, p3.z};
    float wit[3];
    const float d2 = point_tri_dist2(a, b, c, wit);
    *depth = sqrtf(d2);
    if (f_is_zero(*depth)) { dir_out[0] = dir_out[1] = dir_out[2] = 0.f; } else { v3copy(dir_out, wit); v3normalize(dir_out); }
    // find_pos: barycentric blend of the witness points
    const float z0[3] = {p0.x, p0.y, p0.z};
    float bw[4], vec[3];
    v3cross(vec, a, b); bw[0] = v3dot(vec, c);
    v3cross(vec, c, b); bw[1] = v3dot(vec, z0);
    v3cross(vec, z0, a); bw[2] = v3dot(vec, c);
    v3cross(vec, b, a); bw[3] = v3dot(vec, z0);
    float sum = bw[0] + bw[1] + bw[2] + bw[3];
    if (f_is_zero(sum) || sum < 0.f) {
      bw[0] = 0.f;
      v3cross(vec, b, c); bw[1] = v3dot(vec, dir);
      v3cross(vec, c, a); bw[2] = v3dot(vec, dir);
      v3cross(vec, a, b); bw[3] = v3dot(vec, dir);
      sum = bw[1] + bw[2] + bw[3];
    }
    const float inv = 1.f / sum;
    float s1[3] = {0.f, 0.f, 0.f}, s2[3] = {0.f, 0.f, 0.f}, v1[3];
    P.witness(p0, c1, v1); v3addscl(s1, s1, v1, bw[0]); { const float v2[3] = {v1[0] - p0.x, v1[1] - p0.y, v1[2] - p0.z}; v3addscl(s2, s2, v2, bw[0]); }
    P.witness(p1, c1, v1); v3addscl(s1, s1, v1, bw[1]); { const float v2[3] = {v1[0] - p1.x, v1[1] - p1.y, v1[2] - p1.z}; v3addscl(s2, s2, v2, bw[1]); }
    P.witness(p2, c1, v1); v3addscl(s1, s1, v1, bw[2]); { const float v2[3] = {v1[0] - p2.x, v1[1] - p2.y, v1[2] - p2.z}; v3addscl(s2, s2, v2, bw[2]); }
    P.witness(p3, c1, v1); v3addscl(s1, s1, v1, bw[3]); { const float v2[3] = {v1[0] - p3.x, v1[1] - p3.y, v1[2] - p3.z}; v3addscl(s2, s2, v2, bw[3]); }
    for (int k = 0; k < 3; ++k) pos[k] = (s1[k] * inv + s2[k] * inv) * 0.5f;
    return 0;
  }
  return ret;
}

// The same query with one support call site per stage (less bookkeeping per trip, five inlined copies of the Minkowski support).
// Measured on B200 (round 2): with every query on the state machine above the executed code of the collision phase shrinks by
// ~15 KB (the phase's 54 KB exceed the 32 KB L1.5 instruction cache) and the step runs 1.4 - 2.2 % faster on flamingo_p_v3 /
// rocky_hard and 2 % faster on the humanoid, unchanged on w4 / stairs: the staged version is kept for reference behind
// COSIM_MPR_STAGED and is not compiled by default.
#ifdef COSIM_MPR_STAGED
template <class A> DEV int mpr_lane_staged(const A& P, GQ_PARAMS, const float* gcenter, float* depth, float* dir_out, float* pos) {
  typedef typename A::PV PV;
  const float tol = MO(ccd_tolerance); const int maxit = MD(ccd_iterations);
  float c1[3]; P.center(c1);
  PV p0 = P.make_p0(c1, gcenter), p1, p2, p3;
  if (f_eq(p0.x, 0.f) && f_eq(p0.y, 0.f) && f_eq(p0.z, 0.f)) p0.x += CCD_EPS * 10.f;
  float dir[3] = {-p0.x, -p0.y, -p0.z}; v3normalize(dir);
  p1 = P.mink(GQ_ARGS, dir[0], dir[1], dir[2]);
  float dot = pv_dot(p1, dir[0], dir[1], dir[2]);
  if (f_is_zero(dot) || dot < 0.f) { v3copy(dir_out, dir); return -1; }      // -1: `dir_out` separates the two objects
  { const float a[3] = {p0.x, p0.y, p0.z}, b[3] = {p1.x, p1.y, p1.z}; v3cross(dir, a, b); }
  if (f_is_zero(v3dot(dir, dir))) {
    float v1[3]; P.witness(p1, c1, v1);
    for (int k = 0; k < 3; ++k) { const float pk = k == 0 ? p1.x : (k == 1 ? p1.y : p1.z); pos[k] = (v1[k] + (v1[k] - pk)) * 0.5f; }
    if (f_eq(p1.x, 0.f) && f_eq(p1.y, 0.f) && f_eq(p1.z, 0.f)) { *depth = 0.f; dir_out[0] = dir_out[1] = dir_out[2] = 0.f; return 0; }
    float pv[3] = {p1.x, p1.y, p1.z};
    *depth = v3norm(pv); v3copy(dir_out, pv); v3normalize(dir_out); return 0;
  }
  v3normalize(dir);
  p2 = P.mink(GQ_ARGS, dir[0], dir[1], dir[2]);
  dot = pv_dot(p2, dir[0], dir[1], dir[2]);
  if (f_is_zero(dot) || dot < 0.f) { v3copy(dir_out, dir); return -1; }
  { float va[3] = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, vb[3] = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z}; v3cross(dir, va, vb); v3normalize(dir); }
  dot = pv_dot(p0, dir[0], dir[1], dir[2]);
  if (dot > 0.f) { const PV t = p1; p1 = p2; p2 = t; dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2]; }
  int guard = 0;
  while (true) {       // portal discovery
    if (++guard > 100) return -2;
    p3 = P.mink(GQ_ARGS, dir[0], dir[1], dir[2]);
    dot = pv_dot(p3, dir[0], dir[1], dir[2]);
    if (f_is_zero(dot) || dot < 0.f) { v3copy(dir_out, dir); return -1; }
    int cont = 0;
    { const float a[3] = {p1.x, p1.y, p1.z}, b[3] = {p3.x, p3.y, p3.z}; float va[3]; v3cross(va, a, b); dot = pv_dot(p0, va[0], va[1], va[2]); }
    if (dot < 0.f && !f_is_zero(dot)) { p2 = p3; cont = 1; }
    if (!cont) {
      const float a[3] = {p3.x, p3.y, p3.z}, b[3] = {p2.x, p2.y, p2.z}; float va[3]; v3cross(va, a, b); dot = pv_dot(p0, va[0], va[1], va[2]);
      if (dot < 0.f && !f_is_zero(dot)) { p1 = p3; cont = 1; }
    }
    if (!cont) break;
    float va[3] = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, vb[3] = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z}; v3cross(dir, va, vb); v3normalize(dir);
  }
  guard = 0;
  while (true) {       // portal refinement
    if (++guard > 1000) return -2;
    pv_portal_dir(p1, p2, p3, dir);
    dot = pv_dot(p1, dir[0], dir[1], dir[2]);
    if (f_is_zero(dot) || dot > 0.f) break;
    const PV v4 = P.mink(GQ_ARGS, dir[0], dir[1], dir[2]);
    dot = pv_dot(v4, dir[0], dir[1], dir[2]);
    if (!(f_is_zero(dot) || dot > 0.f)) { v3copy(dir_out, dir); return -1; }
    if (pv_reach_tol(p1, p2, p3, v4, dir, tol)) return -2;                      // no intersection, but `dir` is not a separating axis
    pv_expand(p0, p1, p2, p3, v4);
  }
  int it = 0;
  while (true) {       // penetration depth
    pv_portal_dir(p1, p2, p3, dir);
    const PV v4 = P.mink(GQ_ARGS, dir[0], dir[1], dir[2]);
    if (pv_reach_tol(p1, p2, p3, v4, dir, tol) || it > maxit) {
      const float a[3] = {p1.x, p1.y, p1.z}, b[3] = {p2.x, p2.y, p2.z}, c[3] = {p3.x, p3.y, p3.z};
      float wit[3];
      const float d2 = point_tri_dist2(a, b, c, wit);
      *depth = sqrtf(d2);
      if (f_is_zero(*depth)) { dir_out[0] = dir_out[1] = dir_out[2] = 0.f; } else { v3copy(dir_out, wit); v3normalize(dir_out); }
      // find_pos: barycentric blend of the witness points
      const float z0[3] = {p0.x, p0.y, p0.z};
      float bw[4], vec[3];
      v3cross(vec, a, b); bw[0] = v3dot(vec, c);
      v3cross(vec, c, b); bw[1] = v3dot(vec, z0);
      v3cross(vec, z0, a); bw[2] = v3dot(vec, c);
      v3cross(vec, b, a); bw[3] = v3dot(vec, z0);
      float sum = bw[0] + bw[1] + bw[2] + bw[3];
      if (f_is_zero(sum) || sum < 0.f) {
        bw[0] = 0.f;
        v3cross(vec, b, c); bw[1] = v3dot(vec, dir);
        v3cross(vec, c, a); bw[2] = v3dot(vec, dir);
        v3cross(vec, a, b); bw[3] = v3dot(vec, dir);
        sum = bw[1] + bw[2] + bw[3];
      }
      const float inv = 1.f / sum;
      float s1[3] = {0.f, 0.f, 0.f}, s2[3] = {0.f, 0.f, 0.f}, v1[3];
      P.witness(p0, c1, v1); v3addscl(s1, s1, v1, bw[0]); { const float v2[3] = {v1[0] - p0.x, v1[1] - p0.y, v1[2] - p0.z}; v3addscl(s2, s2, v2, bw[0]); }
      P.witness(p1, c1, v1); v3addscl(s1, s1, v1, bw[1]); { const float v2[3] = {v1[0] - p1.x, v1[1] - p1.y, v1[2] - p1.z}; v3addscl(s2, s2, v2, bw[1]); }
      P.witness(p2, c1, v1); v3addscl(s1, s1, v1, bw[2]); { const float v2[3] = {v1[0] - p2.x, v1[1] - p2.y, v1[2] - p2.z}; v3addscl(s2, s2, v2, bw[2]); }
      P.witness(p3, c1, v1); v3addscl(s1, s1, v1, bw[3]); { const float v2[3] = {v1[0] - p3.x, v1[1] - p3.y, v1[2] - p3.z}; v3addscl(s2, s2, v2, bw[3]); }
      for (int k = 0; k < 3; ++k) pos[k] = (s1[k] * inv + s2[k] * inv) * 0.5f;
      return 0;
    }
    pv_expand(p0, p1, p2, p3, v4);
    ++it;
  }
}
#endif

template <class A> DEV int mpr_lane(const A& P, GQ_PARAMS, const float* gcenter, float* depth, float* dir_out, float* pos) {
#ifdef COSIM_MPR_STAGED
  if ((grp >> 8) > 1) return mpr_lane_staged(P, GQ_ARGS, gcenter, depth, dir_out, pos);
#endif
  return mpr_lane_sm(P, GQ_ARGS, gcenter, depth, dir_out, pos);
}

// Conservative separation tests between a terrain prism and the BOUNDING shapes of a geom (oriented box from geom_aabb,
// bounding sphere, bounding cylinder of wheel-like hulls), in the local frame of the query.  A prism they remove is separated
// from the geom by more than CULL_MARGIN, so the MPR query of the serial reference loop would report "no contact" for it:
// results do not change, only the number of MPR queries (w4_p_v2 on the 9.8 mm stairs raster: thousands of prisms per sub-step
// under the AABBs of its 17 hulls, a few dozen of them near a wheel).
#define CULL_MARGIN 2e-5f
struct GeomBound { float cb[3], h[3], cs[3], rb, cc[3], ca[3], cr, chl; const float* R; };
// largest d . y over the bounding shapes (d unit length)
DEV float bound_support(const GeomBound& B, const float* d) {
  const float* R = B.R;
  const float a0 = d[0] * R[0] + d[1] * R[3] + d[2] * R[6], a1 = d[0] * R[1] + d[1] * R[4] + d[2] * R[7], a2 = d[0] * R[2] + d[1] * R[5] + d[2] * R[8];
  float s = v3dot(d, B.cb) + B.h[0] * fabsf(a0) + B.h[1] * fabsf(a1) + B.h[2] * fabsf(a2);
  s = fminf(s, v3dot(d, B.cs) + B.rb);
  if (B.cr > 0.f) { const float ad = v3dot(d, B.ca); s = fminf(s, v3dot(d, B.cc) + B.chl * fabsf(ad) + B.cr * sqrtf(fmaxf(0.f, 1.f - ad * ad))); }
  return s;
}
DEV void make_bound(const ModelDev& m, WSP ws, int g, float ox, float oy, GeomBound& B) {
  const float* R = WS(W_GXMAT) + 9 * g; const float* pos = WS(W_GXPOS) + 3 * g; const float* ab = TB(geom_aabb) + 6 * g; const float* bc = TB(geom_bcyl) + 8 * g;
  B.R = R;
  const float o[3] = {pos[0] - ox, pos[1] - oy, pos[2]};
  { const float c[3] = {LDG(ab), LDG(ab + 1), LDG(ab + 2)}; float r[3]; m3mulv(r, R, c); v3add(B.cb, o, r); B.h[0] = LDG(ab + 3); B.h[1] = LDG(ab + 4); B.h[2] = LDG(ab + 5); }
  { const float c[3] = {LDG(TB(geom_center) + 3 * g), LDG(TB(geom_center) + 3 * g + 1), LDG(TB(geom_center) + 3 * g + 2)}; float r[3]; m3mulv(r, R, c); v3add(B.cs, o, r); B.rb = LDG(TB(geom_rbound) + g) * 1.0001f + 1e-6f; }
  B.cr = LDG(bc + 6); B.chl = LDG(bc + 7);
  { const float c[3] = {LDG(bc), LDG(bc + 1), LDG(bc + 2)}, a[3] = {LDG(bc + 3), LDG(bc + 4), LDG(bc + 5)}; float r[3]; m3mulv(r, R, c); v3add(B.cc, o, r); m3mulv(B.ca, R, a); }
}
// true = the prism with top vertices (x, y, z)[0..2] (bottoms at -base) cannot touch the geom
DEV bool prism_culled(const GeomBound& B, const float* x, const float* y, const float* z) {
  // (1) the prism lies below the plane of its top face
  float n[3]; { const float e1[3] = {x[1] - x[0], y[1] - y[0], z[1] - z[0]}, e2[3] = {x[2] - x[0], y[2] - y[0], z[2] - z[0]}; v3cross(n, e1, e2); }
  const float nn = v3norm(n);
  if (nn > 1e-12f) {
    const float sgn = n[2] < 0.f ? 1.f / nn : -1.f / nn;        // d = downward unit normal of the top face
    const float d[3] = {n[0] * sgn, n[1] * sgn, n[2] * sgn};
    if (bound_support(B, d) < d[0] * x[0] + d[1] * y[0] + d[2] * z[0] - CULL_MARGIN) return true;
  }
  // (2) direction from the bounding shape's core (cylinder axis segment, else box centre) to the top face
  const float p[3] = {(x[0] + x[1] + x[2]) * (1.f / 3.f), (y[0] + y[1] + y[2]) * (1.f / 3.f), (z[0] + z[1] + z[2]) * (1.f / 3.f)};
  float q[3];
  if (B.cr > 0.f) { float w[3]; v3sub(w, p, B.cc); const float t = fminf(B.chl, fmaxf(-B.chl, v3dot(w, B.ca))); v3addscl(q, B.cc, B.ca, t); }
  else v3copy(q, B.cb);
  float d[3]; v3sub(d, p, q);
  const float dn = v3norm(d);
  if (dn > 1e-6f && d[2] < 0.f) {       // pointing down: the prism's smallest d . x is at a top vertex
    v3scl(d, d, 1.f / dn);
    const float pm = fminf(d[0] * x[0] + d[1] * y[0] + d[2] * z[0], fminf(d[0] * x[1] + d[1] * y[1] + d[2] * z[1], d[0] * x[2] + d[1] * y[2] + d[2] * z[2]));
    if (pm - bound_support(B, d) > CULL_MARGIN) return true;
  }
  return false;
}

// per-geom task record in W_GTASK: [cmin, rmin, ncols, nrows, unused, contacts so far, zmin (float), unused]
// one batch of `n` (<= LANES / gs) narrow-phase queries from the ring (entry = geom << 24 | prism index inside the geom's sub-grid),
// gs lanes per query; hits are appended in ring (= task) order, honouring the 50-per-geom cap and the store's capacity
DEV_NOINLINE void mpr_batch(const ModelDev& m, WSP ws, int head, int n, int gs, int lane) { LANE_REFRESH();
  const int ncol = MD(hf_ncol), nrow = MD(hf_nrow);
  const float sx = MO(hf_sx), sy = MO(hf_sy), sz = MO(hf_sz), base = MO(hf_base);
  const float dx = 2.f * sx / (float)(ncol - 1), dy = 2.f * sy / (float)(nrow - 1);
  int* task = WSI(W_GTASK); const int* ring = WSI(W_RING);
  const int sub = lane & (gs - 1), k = lane / gs;
  const unsigned gmask = ((gs >= 32 ? 0u : (1u << gs)) - 1u) << (lane & ~(gs - 1));
  int hit = 0, g = 0, cell = -1; float depth = 0.f, nrm[3] = {0.f, 0.f, 0.f}, cp[3] = {0.f, 0.f, 0.f};
  if (k < n) {
    const int ent = ring[(head + k) & (RING_SIZE - 1)];
    g = (int)((unsigned)ent >> 24);
    const int local = ent & 0xffffff;
    const int* tk = task + 8 * g;
    if (tk[5] < 50) {
      const int per_row = 2 * tk[2];
      const int r = tk[1] + local / per_row, rem = local % per_row, c = tk[0] + 1 + (rem >> 1), i = rem & 1;
      // strip walk of the reference: triangle i of cell (r, c-1): i = 0 -> (c-1,r) (c-1,r+1) (c,r); i = 1 -> (c-1,r+1) (c,r) (c,r+1)
      const int ca = c - 1, ra = r + i, cb = i ? c : c - 1, rbb = i ? r : r + 1, cc = c, rc = r + i;
      // The query runs in a frame whose xy origin is the first grid vertex of this geom's sub-grid: far from the world
      // origin (terrains span +-140 m) fp32 world coordinates resolve ~1e-5 m, ten times the MPR tolerance, while the
      // local coordinates of prism and geom stay below a few metres.  Only the origin shift is rounded, and it is the
      // same for both shapes.
      const float ox = dx * (float)tk[0] - sx, oy = dy * (float)tk[1] - sy;
      PrismA PA; PrismL& P = PA.P; P.base = base;
      P.x[0] = dx * (float)(ca - tk[0]); P.y[0] = dy * (float)(ra - tk[1]); P.z[0] = LDGB(m.hfield_data + (size_t)ra * ncol + ca) * sz;
      P.x[1] = dx * (float)(cb - tk[0]); P.y[1] = dy * (float)(rbb - tk[1]); P.z[1] = LDGB(m.hfield_data + (size_t)rbb * ncol + cb) * sz;
      P.x[2] = dx * (float)(cc - tk[0]); P.y[2] = dy * (float)(rc - tk[1]); P.z[2] = LDGB(m.hfield_data + (size_t)rc * ncol + cc) * sz;
#if defined(COSIM_PHASE_TIMING) && !defined(COSIM_HOST_EMU)
      if (sub == 0) atomicAdd(m.phase + PH_MPR_CALLS, 1ull);
#endif
      const float cl[3] = {LDG(TB(geom_center) + 3 * g), LDG(TB(geom_center) + 3 * g + 1), LDG(TB(geom_center) + 3 * g + 2)};
      const float* gpos = WS(W_GXPOS) + 3 * g;
      float gc[3]; m3mulv(gc, WS(W_GXMAT) + 9 * g, cl);
      gc[0] += gpos[0] - ox; gc[1] += gpos[1] - oy; gc[2] += gpos[2];
      const int cell_ = ((r * ncol + (c - 1)) << 1) | i;               // before the query: one live value instead of four
      if (mpr_lane(PA, m, ws, g, sub | (gs << 8), gmask, ox, oy, gc, &depth, nrm, cp) == 0 && !(nrm[0] == 0.f && nrm[1] == 0.f && nrm[2] == 0.f) && depth == depth) {
        hit = (sub == 0); cell = cell_;       // one lane per group reports the contact
        cp[0] += ox; cp[1] += oy;
      }
    }
  }
  // ---- append the hits of this batch in task order, honouring the per-geom cap (50) and the capacity of the store
  task = WSI(W_GTASK);          // re-derived: nothing but the outcome of the query needs to survive the narrow phase in registers
  int ncon = WSI(W_CNT)[CNT_NCON], dropped = WSI(W_CNT)[CNT_DROPPED];
  const int cap = MD(ncon_max);
  unsigned hits = wballot(hit);
  while (hits) {
    const int src = ctz32(hits);
    const int gsrc = wshfl(g, src);
    const unsigned same = wballot(hit && g == gsrc);      // hits of this geom, in task order
    const int have = task[8 * gsrc + 5];
    const int rank = popc32(same & ((1u << lane) - 1u));
    const int keep = hit && g == gsrc && (have + rank) < 50;
    const unsigned keepm = wballot(keep);
    const int slot = ncon + popc32(keepm & ((1u << lane) - 1u));
    if (keep && slot < cap) write_contact(m, ws, slot, cp, nrm, -depth, fmaxf(WS(W_SCAL)[0], WS(W_GMU)[g]), TB(geom_body)[g], g, cell);
    const int nkeep = popc32(keepm), room = imax(0, cap - ncon);
    dropped += imax(0, nkeep - room); ncon += imin(nkeep, room);
    SYNC();
    if (lane == 0) task[8 * gsrc + 5] = have + nkeep;
    SYNC();
    hits &= ~same;
  }
  if (lane == 0) { WSI(W_CNT)[CNT_NCON] = ncon; WSI(W_CNT)[CNT_DROPPED] = dropped; }
  SYNC();
}

// part 0: the whole pass; 1: stage 1 only (per-geom bounds and sub-grids -> W_GTASK); 2: stage 2 only (prisms, narrow phase)
// FINE = false: the instance for coarse rasters (cells of 0.2 m and more: a geom covers a handful of prisms) carries neither the
// bounding-shape culls nor the block-wise sweep (both only skip work; same results), which frees ~25 registers in the prism loop
template <bool FINE> DEV_NOINLINE void collide_hfield_all(const ModelDev& m, WSP ws, int lane, int part = 0) { LANE_REFRESH();
  const int ng = MD(ngeom), nrow = MD(hf_nrow), ncol = MD(hf_ncol);
  const float sx = MO(hf_sx), sy = MO(hf_sy), sz = MO(hf_sz), base = MO(hf_base);
  const float dx = 2.f * sx / (float)(ncol - 1), dy = 2.f * sy / (float)(nrow - 1);
  int* task = WSI(W_GTASK);
  if (part != 2) {
  if (lane == 0) { WSI(W_CNT)[CNT_NCON] = 0; WSI(W_CNT)[CNT_DROPPED] = 0; }
  // ---- stage 1: one group of gs1 lanes per geom (gs1 = largest power of two with ng * gs1 <= LANES)
  int gs1 = 1;
#ifndef COSIM_HOST_EMU
  while (gs1 < 8 && ng * gs1 * 2 <= LANES) gs1 <<= 1;
#endif
  const int sub1 = lane & (gs1 - 1);
  const unsigned gmask1 = gs1 >= LANES ? 0xffffffffu : (((1u << gs1) - 1u) << (lane & ~(gs1 - 1)));
  NOUNROLL for (int g = lane / gs1; g < ng; g += LANES / gs1) {
    int* tk = task + 8 * g;
    if (sub1 == 0) { tk[0] = tk[1] = tk[2] = tk[3] = tk[4] = tk[5] = 0; }
    const int grp1 = sub1 | (gs1 << 8);
    const float cl[3] = {LDG(TB(geom_center) + 3 * g), LDG(TB(geom_center) + 3 * g + 1), LDG(TB(geom_center) + 3 * g + 2)};
    float pos[3]; m3mulv(pos, WS(W_GXMAT) + 9 * g, cl); v3add(pos, WS(W_GXPOS) + 3 * g, pos);
    const float rb = LDG(TB(geom_rbound) + g);
    if (pos[0] - rb > sx || pos[0] + rb < -sx || pos[1] - rb > sy || pos[1] + rb < -sy) continue;
    if (pos[2] - rb > sz || pos[2] + rb < -base) continue;
    float hmax = INFINITY;
    {  // conservative early-outs (do not change results): highest terrain vertex under the bounding sphere
      const int c0 = imax(0, (int)floorf((pos[0] - rb + sx) / dx)), c1 = imin(ncol - 1, (int)ceilf((pos[0] + rb + sx) / dx));
      const int r0 = imax(0, (int)floorf((pos[1] - rb + sy) / dy)), r1 = imin(nrow - 1, (int)ceilf((pos[1] + rb + sy) / dy));
      const int w = c1 - c0 + 1, cnt = w * (r1 - r0 + 1);
      if (cnt <= 64) {
        hmax = -INFINITY;
        NOUNROLL for (int t = 0; t < cnt; t += 4) {          // four independent loads in flight
          const int t1 = imin(t + 1, cnt - 1), t2 = imin(t + 2, cnt - 1), t3 = imin(t + 3, cnt - 1);
          const float h0 = LDGB(m.hfield_data + (size_t)(r0 + t / w) * ncol + c0 + t % w), h1 = LDGB(m.hfield_data + (size_t)(r0 + t1 / w) * ncol + c0 + t1 % w);
          const float h2 = LDGB(m.hfield_data + (size_t)(r0 + t2 / w) * ncol + c0 + t2 % w), h3 = LDGB(m.hfield_data + (size_t)(r0 + t3 / w) * ncol + c0 + t3 % w);
          hmax = fmaxf(fmaxf(hmax, fmaxf(h0, h1)), fmaxf(h2, h3));
        }
        hmax *= sz;
      } else if (FINE && m.hf_max8 && c1 >= c0 && r1 >= r0) {        // fine rasters: block maxima (8 x 8 cells per block)
        hmax = -INFINITY;
        const int bw = (c1 >> 3) - (c0 >> 3) + 1, bcnt = bw * ((r1 >> 3) - (r0 >> 3) + 1);
        NOUNROLL for (int t = 0; t < bcnt; ++t) hmax = fmaxf(hmax, LDGB(m.hf_max8 + (size_t)((r0 >> 3) + t / bw) * m.hf_mcol + (c0 >> 3) + t % bw));
        hmax *= sz;
      }
      if (pos[2] - rb > hmax) continue;
    }
    float xmin[3], xmax[3];
    xmin[2] = support_lane(m, ws, g, grp1, gmask1, 0.f, 0.f, 0.f, 0.f, -1.f).z;
    if (xmin[2] > hmax) continue;         // the geom's lowest point clears every terrain vertex it could reach: no contact
    xmax[0] = support_lane(m, ws, g, grp1, gmask1, 0.f, 0.f, 1.f, 0.f, 0.f).x; xmin[0] = support_lane(m, ws, g, grp1, gmask1, 0.f, 0.f, -1.f, 0.f, 0.f).x;
    xmax[1] = support_lane(m, ws, g, grp1, gmask1, 0.f, 0.f, 0.f, 1.f, 0.f).y; xmin[1] = support_lane(m, ws, g, grp1, gmask1, 0.f, 0.f, 0.f, -1.f, 0.f).y;
    xmax[2] = support_lane(m, ws, g, grp1, gmask1, 0.f, 0.f, 0.f, 0.f, 1.f).z;
    if (xmin[0] > sx || xmax[0] < -sx || xmin[1] > sy || xmax[1] < -sy || xmin[2] > sz || xmax[2] < -base) continue;
    int cmin = (int)floorf((xmin[0] + sx) / (2.f * sx) * (float)(ncol - 1));
    int cmax = (int)ceilf((xmax[0] + sx) / (2.f * sx) * (float)(ncol - 1));
    int rmin = (int)floorf((xmin[1] + sy) / (2.f * sy) * (float)(nrow - 1));
    int rmax = (int)ceilf((xmax[1] + sy) / (2.f * sy) * (float)(nrow - 1));
    cmin = imax(0, cmin); rmin = imax(0, rmin); cmax = imin(ncol - 1, cmax); rmax = imin(nrow - 1, rmax);
    if (cmax <= cmin || rmax <= rmin) continue;
    if (sub1 == 0) { tk[0] = cmin; tk[1] = rmin; tk[2] = cmax - cmin; tk[3] = rmax - rmin; ((float*)tk)[6] = xmin[2]; }
  }
  SYNC();
  }
  if (part == 1) return;
  // ---- stage 2: a lane per prism runs the reference's height test and the bounding-shape culls; the survivors queue up in
  //      a ring (task order), and whenever 32 of them wait one MPR batch runs with a lane per query.  What is left at the
  //      end runs with wider lane groups that share the hull scans (the common case on coarse rasters: a handful of prisms).
  //      Large sub-grids (fine rasters: thousands of prisms under one hull) are walked by blocks of 8 x 8 cells: of every
  //      band of 8 rows only the column blocks whose highest vertex reaches the geom's lowest point are enumerated, as
  //      `segments` of each row; rows and cells keep the reference order.
  int* ring = WSI(W_RING);
  int* seg = ring + RING_SIZE;              // [0..7] first cell column of a segment, [8..16] prisms of a row before it
  int rhead = 0, rcount = 0;
  NOUNROLL for (int g = 0; g < ng; ++g) {
    const int* tk = task + 8 * g;
    const int cmin = tk[0], rmin = tk[1], ncols = tk[2], nrows = tk[3], Tg = 2 * ncols * nrows;
    if (Tg == 0) continue;
    if (Tg >= (1 << 24)) { if (lane == 0) WSI(W_CNT)[CNT_DROPPED] += 1; continue; }      // sub-grid beyond the task encoding (a geom spanning > 2000 x 2000 cells)
    const float zmin = ((const float*)tk)[6];
    const float ox = dx * (float)cmin - sx, oy = dy * (float)rmin - sy;
    // the bounding-shape culls pay off on fine rasters (dozens to thousands of prisms under a hull); a handful of prisms goes
    // straight to the narrow phase after the reference's height test
    const bool cull = FINE && Tg > 16;
    GeomBound B;
    if (cull) make_bound(m, ws, g, ox, oy, B);
    const int per_row = 2 * ncols;
    const int cb0 = cmin >> 3, ncb = ((cmin + ncols - 1) >> 3) - cb0 + 1;
    const bool blocked = FINE && m.hf_max8 != nullptr && Tg > 256 && ncb <= 8;
    const int band0 = blocked ? (rmin >> 3) : 0, band1 = blocked ? ((rmin + nrows - 1) >> 3) : 0;
    bool full = false;
    NOUNROLL for (int band = band0; band <= band1 && !full; ++band) {
      int ra = rmin, re = rmin + nrows, nseg = 1, row_prisms = per_row;
      SYNC();
      if (!blocked) { if (lane == 0) { seg[0] = cmin; seg[8] = 0; } }
      else {
        ra = imax(rmin, 8 * band); re = imin(rmin + nrows, 8 * band + 8);
        unsigned live = 0u;
        NOUNROLL for (int j0 = 0; j0 < ncb; j0 += LANES) {      // one pass on the GPU (ncb <= 8); the single-lane host emulation loops
          const int j = j0 + lane;
          const int isl = j < ncb && LDGB(m.hf_max8 + (size_t)band * m.hf_mcol + cb0 + j) * sz >= zmin;
          live |= wballot(isl) << j0;
        }
        if (!live) continue;
        nseg = 0; row_prisms = 0;
        NOUNROLL while (live) {
          const int j = ctz32(live); live &= live - 1u;
          const int c0 = imax(cmin, 8 * (cb0 + j)), c1 = imin(cmin + ncols, 8 * (cb0 + j) + 8);
          if (lane == 0) { seg[nseg] = c0; seg[8 + nseg] = row_prisms; }
          row_prisms += 2 * (c1 - c0); ++nseg;
        }
      }
      SYNC();
      const int total = (re - ra) * row_prisms;
      NOUNROLL for (int s0 = 0; s0 < total; s0 += LANES) {
        if (tk[5] >= 50) { full = true; break; }                 // this geom already has its 50 contacts (mjMAXCONPAIR)
        const int slot = s0 + lane;
        int pass = 0, t = 0;
        if (slot < total) {
          const int rr = slot / row_prisms, rem = slot - rr * row_prisms, r = ra + rr;
          int sg = 0;
          NOUNROLL while (sg + 1 < nseg && seg[8 + sg + 1] <= rem) ++sg;
          const int rem2 = rem - seg[8 + sg], cellc = seg[sg] + (rem2 >> 1), i = rem2 & 1, c = cellc + 1;
          t = (r - rmin) * per_row + 2 * (cellc - cmin) + i;
          const int ca = c - 1, rra = r + i, cb = i ? c : c - 1, rbb = i ? r : r + 1, cc = c, rc = r + i;
          float x[3], y[3], z[3];
          x[0] = dx * (float)(ca - cmin); y[0] = dy * (float)(rra - rmin); z[0] = LDGB(m.hfield_data + (size_t)rra * ncol + ca) * sz;
          x[1] = dx * (float)(cb - cmin); y[1] = dy * (float)(rbb - rmin); z[1] = LDGB(m.hfield_data + (size_t)rbb * ncol + cb) * sz;
          x[2] = dx * (float)(cc - cmin); y[2] = dy * (float)(rc - rmin); z[2] = LDGB(m.hfield_data + (size_t)rc * ncol + cc) * sz;
          pass = !(z[0] < zmin && z[1] < zmin && z[2] < zmin) && !(cull && prism_culled(B, x, y, z));
        }
        const unsigned pm = wballot(pass);
        if (pass) ring[(rhead + rcount + popc32(pm & ((1u << lane) - 1u))) & (RING_SIZE - 1)] = (g << 24) | t;
        rcount += popc32(pm);
        SYNC();
        NOUNROLL while (rcount >= LANES) { mpr_batch(m, ws, rhead, LANES, 1, lane); rhead += LANES; rcount -= LANES; }
      }
    }
  }
  NOUNROLL while (rcount > 0) {
    int gs2 = 1;
#ifndef COSIM_HOST_EMU
    gs2 = rcount <= 4 ? 8 : (rcount <= 8 ? 4 : (rcount <= 16 ? 2 : 1));
#endif
    const int n = imin(rcount, LANES / gs2);
    mpr_batch(m, ws, rhead, n, gs2, lane); rhead += n; rcount -= n;
  }
  SYNC();
}

// body of the first geom of a contact: 0 (world) for ground contacts; geom-geom contacts carry -2 - geom1 in the record's cell slot
DEV int contact_body1(const ModelDev& m, WSP ws, int c) { const int cell = CRECI(c)[CR_CELL]; return cell <= -2 ? TB(geom_body)[-2 - cell] : 0; }

// mjc_fixNormal restated (oracle/oracle.hpp fix_normal): for smooth primitives the contact normal is rebuilt from the contact
// point -- sphere: centre to point; cylinder: radial direction of the wall unless the point sits on / near a cap.  A normal
// from geom 2 is flipped, two normals are averaged.  pos is in world coordinates.
DEV void fix_normal(const ModelDev& m, WSP ws, int g1, int g2, const float* pos, float* normal) {
  float acc[3] = {0.f, 0.f, 0.f}; int n = 0;
  for (int i = 0; i < 2; ++i) {
    const int g = i ? g2 : g1, type = TB(geom_type)[g];
    if (type != GEOM_SPHERE && type != GEOM_CYLINDER) continue;
    const float* R = WS(W_GXMAT) + 9 * g; const float* x = WS(W_GXPOS) + 3 * g;
    float rel[3], loc[3]; v3sub(rel, pos, x); m3tmulv(loc, R, rel);
    if (type == GEOM_CYLINDER) {
      const float rad = LDG(TB(geom_size) + 3 * g), hl = LDG(TB(geom_size) + 3 * g + 1);
      if (fabsf(loc[2]) > 0.95f * hl) continue;
      const float dflat = fabsf(hl - fabsf(loc[2])), dround = fabsf(rad - sqrtf(loc[0] * loc[0] + loc[1] * loc[1]));
      if (!(dround < dflat)) continue;
      loc[2] = 0.f;
    }
    const float nn = v3norm(loc);
    if (nn < MINVALF) continue;
    v3scl(loc, loc, 1.f / nn);
    float w[3]; m3mulv(w, R, loc);
    if (i == 1) v3scl(w, w, -1.f);
    v3add(acc, acc, w); ++n;
  }
  if (!n) return;
  const float nn = v3norm(acc);
  if (nn < MINVALF) return;
  v3scl(normal, acc, 1.f / nn);
}

// ------------------------------------------------------------------------------------------ box-box (mjc_BoxBox)
// Restates oracle/oracle.hpp box_box (separating axes: 3 + 3 face normals, 9 edge-edge directions with the faces preferred by a
// factor 1.05; face contact = incident face clipped against the reference rectangle, up to 8 points; edge-edge = one point).
// out[k] = pos(3), normal(3), dist; positions relative to the caller's origin (p1 / p2 are given in that frame).
DEV int clip_rect_quad(const float* h, const float* quad, float* ret) {
  int nq = 4, nr = 0; float buf[16]; const float* q = quad; float* r = ret;
  for (int dir = 0; dir <= 1; ++dir) for (int sign = -1; sign <= 1; sign += 2) {
    const float* pq = q; float* pr = r; nr = 0; bool full = false;
    for (int i = nq; i > 0 && !full; --i) {
      if (sign * pq[dir] < h[dir]) { pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; ++nr; if (nr & 8) { full = true; break; } }
      const float* nextq = (i > 1) ? pq + 2 : q;
      if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
        pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
        pr[dir] = sign * h[dir]; pr += 2; ++nr; if (nr & 8) { full = true; break; }
      }
      pq += 2;
    }
    q = r;
    if (full) { dir = 2; break; }
    r = (q == ret) ? buf : ret; nq = nr;
  }
  if (q != ret) for (int i = 0; i < 2 * nr; ++i) ret[i] = q[i];
  return nr;
}
DEV_NOINLINE int box_box(const float* p1, const float* R1, const float* A, const float* p2, const float* R2, const float* B, float* out) {
  const float fudge = 1.05f, eps = 1e-6f;
  float p[3], pp[3], Rm[3][3], Q[3][3], a1[3][3], a2[3][3];
  v3sub(p, p2, p1);
  for (int j = 0; j < 3; ++j) { a1[j][0] = R1[j]; a1[j][1] = R1[3 + j]; a1[j][2] = R1[6 + j]; a2[j][0] = R2[j]; a2[j][1] = R2[3 + j]; a2[j][2] = R2[6 + j]; }
  for (int i = 0; i < 3; ++i) { pp[i] = v3dot(a1[i], p); for (int j = 0; j < 3; ++j) { Rm[i][j] = v3dot(a1[i], a2[j]); Q[i][j] = fabsf(Rm[i][j]); } }
  float s = -INFINITY, normal[3] = {0.f, 0.f, 0.f}; int code = 0; bool invert = false;
  for (int i = 0; i < 3; ++i) {
    const float e = pp[i], s2 = fabsf(e) - (A[i] + B[0] * Q[i][0] + B[1] * Q[i][1] + B[2] * Q[i][2]);
    if (s2 > 0.f) return 0;
    if (s2 > s) { s = s2; v3copy(normal, a1[i]); invert = e < 0.f; code = 1 + i; }
  }
  for (int j = 0; j < 3; ++j) {
    const float e = v3dot(a2[j], p), s2 = fabsf(e) - (A[0] * Q[0][j] + A[1] * Q[1][j] + A[2] * Q[2][j] + B[j]);
    if (s2 > 0.f) return 0;
    if (s2 > s) { s = s2; v3copy(normal, a2[j]); invert = e < 0.f; code = 4 + j; }
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    const float r[3] = {Rm[0][j], Rm[1][j], Rm[2][j]}; float n[3] = {0.f, 0.f, 0.f};
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    n[i1] = -r[i2]; n[i2] = r[i1];
    const float e = pp[0] * n[0] + pp[1] * n[1] + pp[2] * n[2];
    float s2 = fabsf(e) - (A[i1] * (Q[i2][j] + eps) + A[i2] * (Q[i1][j] + eps) + B[j1] * (Q[i][j2] + eps) + B[j2] * (Q[i][j1] + eps));
    if (s2 > eps) return 0;
    const float l = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (l > eps) {
      s2 /= l;
      if (s2 * fudge > s) { s = s2; for (int k = 0; k < 3; ++k) normal[k] = (a1[0][k] * n[0] + a1[1][k] * n[1] + a1[2][k] * n[2]) / l; invert = e < 0.f; code = 7 + 3 * i + j; }
    }
  }
  if (!code) return 0;
  if (invert) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
  const float depth = -s;
  if (code > 6) {
    float pa[3], pb[3]; v3copy(pa, p1); v3copy(pb, p2);
    for (int j = 0; j < 3; ++j) { const float sg = v3dot(normal, a1[j]) > 0.f ? 1.f : -1.f; v3addscl(pa, pa, a1[j], sg * A[j]); }
    for (int j = 0; j < 3; ++j) { const float sg = v3dot(normal, a2[j]) > 0.f ? -1.f : 1.f; v3addscl(pb, pb, a2[j], sg * B[j]); }
    const float* ua = a1[(code - 7) / 3]; const float* ub = a2[(code - 7) % 3];
    float dp[3]; v3sub(dp, pb, pa);
    const float uaub = v3dot(ua, ub), q1 = v3dot(ua, dp), q2 = -v3dot(ub, dp), dd = 1.f - uaub * uaub;
    float alpha = 0.f, beta = 0.f;
    if (dd > 1e-4f) { alpha = (q1 + uaub * q2) / dd; beta = (uaub * q1 + q2) / dd; }
    v3addscl(pa, pa, ua, alpha); v3addscl(pb, pb, ub, beta);
    for (int k = 0; k < 3; ++k) { out[k] = 0.5f * (pa[k] + pb[k]); out[3 + k] = normal[k]; }
    out[6] = -depth;
    return 1;
  }
  const bool first = code <= 3;
  const float (*Aa)[3] = first ? a1 : a2; const float (*Ab)[3] = first ? a2 : a1;
  const float* pa = first ? p1 : p2; const float* pb = first ? p2 : p1; const float* Sa = first ? A : B; const float* Sb = first ? B : A;
  float n2[3]; for (int k = 0; k < 3; ++k) n2[k] = first ? normal[k] : -normal[k];
  float nr[3], anr[3]; for (int j = 0; j < 3; ++j) { nr[j] = v3dot(Ab[j], n2); anr[j] = fabsf(nr[j]); }
  int lanr, b1, b2;
  if (anr[1] > anr[0]) { if (anr[1] > anr[2]) { b1 = 0; lanr = 1; b2 = 2; } else { b1 = 0; b2 = 1; lanr = 2; } }
  else { if (anr[0] > anr[2]) { lanr = 0; b1 = 1; b2 = 2; } else { b1 = 0; b2 = 1; lanr = 2; } }
  float center[3];
  for (int k = 0; k < 3; ++k) center[k] = pb[k] - pa[k] + (nr[lanr] < 0.f ? 1.f : -1.f) * Sb[lanr] * Ab[lanr][k];
  const int codeN = first ? code - 1 : code - 4, c1i = codeN == 0 ? 1 : 0, c2i = codeN == 2 ? 1 : 2;
  const float c1 = v3dot(center, Aa[c1i]), c2 = v3dot(center, Aa[c2i]);
  const float m11 = v3dot(Aa[c1i], Ab[b1]), m12 = v3dot(Aa[c1i], Ab[b2]), m21 = v3dot(Aa[c2i], Ab[b1]), m22 = v3dot(Aa[c2i], Ab[b2]);
  const float k1 = m11 * Sb[b1], k2 = m21 * Sb[b1], k3 = m12 * Sb[b2], k4 = m22 * Sb[b2];
  const float quad[8] = {c1 - k1 - k3, c2 - k2 - k4, c1 - k1 + k3, c2 - k2 + k4, c1 + k1 + k3, c2 + k2 + k4, c1 + k1 - k3, c2 + k2 - k4};
  const float rect[2] = {Sa[c1i], Sa[c2i]};
  float ret[16];
  const int n = clip_rect_quad(rect, quad, ret);
  if (n < 1) return 0;
  const float det1 = 1.f / (m11 * m22 - m12 * m21);
  int cnum = 0;
  for (int j = 0; j < n && cnum < 8; ++j) {
    const float kk1 = (m22 * (ret[2 * j] - c1) - m12 * (ret[2 * j + 1] - c2)) * det1, kk2 = (-m21 * (ret[2 * j] - c1) + m11 * (ret[2 * j + 1] - c2)) * det1;
    float pt[3]; for (int k = 0; k < 3; ++k) pt[k] = center[k] + kk1 * Ab[b1][k] + kk2 * Ab[b2][k];
    const float dep = Sa[codeN] - v3dot(n2, pt);
    if (dep >= 0.f) {
      float* o = out + 7 * cnum;
      for (int k = 0; k < 3; ++k) { o[k] = pt[k] + pa[k] + n2[k] * dep * 0.5f; o[3 + k] = normal[k]; }
      o[6] = -dep; ++cnum;
    }
  }
  return cnum;
}

// ------------------------------------------------------------------------------------------ geom-geom (self) collision
// mj_collideGeoms -> mj_filterSphere -> mjc_Convex restated (oracle/oracle.hpp collide_pairs is the serial version), different
// schedule: (1) a lane per candidate pair runs a separating-axis cull on the two geom-frame bounding boxes (conservative: it
// only removes pairs MPR would report as separated; MuJoCo's mid-phase culls oriented boxes the same way) ahead of the
// reference's bounding-sphere filter; (2) the survivors of each 32-pair chunk are handed to lane groups (8 / 4
// / 2 / 1 lanes per pair, sharing the hull support scans) that run the MPR query in a frame centred on the first geom.
// Hits are appended after the ground contacts in pair order.
// BB: the model has box-box pairs (mjc_BoxBox: several contacts per pair, appended by prefix sum); models without them run the
// instance that carries none of that code (the headline workload lost 2.5 % to it when there was one instance)
template <bool BB> DEV_NOINLINE void collide_pairs(const ModelDev& m, WSP ws, int lane) { LANE_REFRESH();
  const int npair = MD(npair);
  int ncon = WSI(W_CNT)[CNT_NCON], dropped = WSI(W_CNT)[CNT_DROPPED];
  SYNC();
  NOUNROLL for (int p0 = 0; p0 < npair; p0 += LANES) {
    const int p = p0 + lane;
    int cand = 0;
    if (p < npair) {      // separating-axis test on the six face normals of the two geom-frame boxes
      const int g1 = TB(pair_geom)[2 * p], g2 = TB(pair_geom)[2 * p + 1];
      const float* R1 = WS(W_GXMAT) + 9 * g1; const float* R2 = WS(W_GXMAT) + 9 * g2;
      const float* x1 = WS(W_GXPOS) + 3 * g1; const float* x2 = WS(W_GXPOS) + 3 * g2;
      const float* a1 = TB(geom_aabb) + 6 * g1; const float* a2 = TB(geom_aabb) + 6 * g2;
      const float c1[3] = {LDG(a1), LDG(a1 + 1), LDG(a1 + 2)}, c2[3] = {LDG(a2), LDG(a2 + 1), LDG(a2 + 2)};
      const float h1[3] = {LDG(a1 + 3), LDG(a1 + 4), LDG(a1 + 5)}, h2[3] = {LDG(a2 + 3), LDG(a2 + 4), LDG(a2 + 5)};
      float w1[3], w2[3], t[3]; m3mulv(w1, R1, c1); m3mulv(w2, R2, c2);
      for (int k = 0; k < 3; ++k) t[k] = (x2[k] - x1[k]) + (w2[k] - w1[k]);
      float C[3][3];        // |axis i of frame 1 . axis j of frame 2|
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i][j] = fabsf(R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j]);
      cand = 1;
      for (int i = 0; i < 3; ++i) {
        const float proj = fabsf(t[0] * R1[i] + t[1] * R1[3 + i] + t[2] * R1[6 + i]);
        const float rad = h1[i] + C[i][0] * h2[0] + C[i][1] * h2[1] + C[i][2] * h2[2];
        if (proj > rad * 1.0001f + 1e-5f) cand = 0;
      }
      for (int j = 0; j < 3; ++j) {
        const float proj = fabsf(t[0] * R2[j] + t[1] * R2[3 + j] + t[2] * R2[6 + j]);
        const float rad = h2[j] + C[0][j] * h1[0] + C[1][j] * h1[1] + C[2][j] * h1[2];
        if (proj > rad * 1.0001f + 1e-5f) cand = 0;
      }
    }
#ifdef COSIM_HOST_EMU
    unsigned cmask = cand ? 1u : 0u;
#else
    unsigned cmask = __ballot_sync(0xffffffffu, cand);
#endif
    if (!cmask) continue;
    const int nc = popc32(cmask);
    int gs = 1;
#ifndef COSIM_HOST_EMU
    gs = nc <= 4 ? 8 : (nc <= 8 ? 4 : (nc <= 16 ? 2 : 1));
#endif
    const int sub = lane & (gs - 1);
    const unsigned gmask = ((gs >= 32 ? 0u : (1u << gs)) - 1u) << (lane & ~(gs - 1));
    NOUNROLL for (int k0 = 0; k0 < nc; k0 += LANES / gs) {
      const int k = k0 + lane / gs;           // k-th surviving pair of this chunk (pair order)
      int hit = 0, g1 = 0, g2 = 0; float depth = 0.f, nrm[3] = {0.f, 0.f, 0.f}, cp[3] = {0.f, 0.f, 0.f};
      float* bbout = nullptr; float bbx = 0.f, bby = 0.f;          // box-box pair: `hit` = number of contacts waiting in the lane's scratch block
      if (k < nc) {
        unsigned mm = cmask; for (int i = 0; i < k; ++i) mm &= mm - 1;
        const int pp = p0 + ctz32(mm);
        g1 = TB(pair_geom)[2 * pp]; g2 = TB(pair_geom)[2 * pp + 1];
        const float cl1[3] = {LDG(TB(geom_center) + 3 * g1), LDG(TB(geom_center) + 3 * g1 + 1), LDG(TB(geom_center) + 3 * g1 + 2)};
        const float cl2[3] = {LDG(TB(geom_center) + 3 * g2), LDG(TB(geom_center) + 3 * g2 + 1), LDG(TB(geom_center) + 3 * g2 + 2)};
        const float* x1 = WS(W_GXPOS) + 3 * g1; const float* x2 = WS(W_GXPOS) + 3 * g2;
        float r1[3], r2[3]; m3mulv(r1, WS(W_GXMAT) + 9 * g1, cl1); m3mulv(r2, WS(W_GXMAT) + 9 * g2, cl2);
        const float ox = x1[0], oy = x1[1];
        GeomA A; A.g1 = g1; A.c[0] = r1[0]; A.c[1] = r1[1]; A.c[2] = x1[2] + r1[2];
        const float gc[3] = {(x2[0] - ox) + r2[0], (x2[1] - oy) + r2[1], x2[2] + r2[2]};
        const float dv[3] = {gc[0] - A.c[0], gc[1] - A.c[1], gc[2] - A.c[2]};
        const float bound = LDG(TB(geom_rbound) + g1) + LDG(TB(geom_rbound) + g2);
        if (BB && v3dot(dv, dv) <= bound * bound && TB(geom_type)[g1] == GEOM_BOX && TB(geom_type)[g2] == GEOM_BOX) {      // mjc_BoxBox: up to 8 contacts, one lane
          if (sub == 0) {
            bbout = *(float* const*)(WSF(ws) + m.off[W_GPTR]) + m.bb_off + 56 * lane;
            const float q1[3] = {0.f, 0.f, x1[2]}, q2[3] = {x2[0] - ox, x2[1] - oy, x2[2]};      // frame centred on geom 1 in x, y
            const float sA[3] = {LDG(TB(geom_size) + 3 * g1), LDG(TB(geom_size) + 3 * g1 + 1), LDG(TB(geom_size) + 3 * g1 + 2)};
            const float sB[3] = {LDG(TB(geom_size) + 3 * g2), LDG(TB(geom_size) + 3 * g2 + 1), LDG(TB(geom_size) + 3 * g2 + 2)};
            hit = box_box(q1, WS(W_GXMAT) + 9 * g1, sA, q2, WS(W_GXMAT) + 9 * g2, sB, bbout);
            bbx = ox; bby = oy;
          }
        } else if (v3dot(dv, dv) <= bound * bound) {      // mj_filterSphere, margin 0
          // temporal coherence: the direction along which MPR last found this pair separated (kept in geom 1's frame, so
          // it follows the robot).  If the two supports still leave a gap the pair cannot intersect and the query is
          // skipped; a stale or useless axis only fails this test, so the cull is conservative like the ones above.
          int res = 1;
          {
            const float* ax = WS(W_PAXIS) + 4 * (pp % PAXIS_SLOTS);
            if (__float_as_int_emu(ax[0]) == pp + 1) {
              const float* R1 = WS(W_GXMAT) + 9 * g1;
              const float d[3] = {R1[0] * ax[1] + R1[1] * ax[2] + R1[2] * ax[3], R1[3] * ax[1] + R1[4] * ax[2] + R1[5] * ax[3], R1[6] * ax[1] + R1[7] * ax[2] + R1[8] * ax[3]};
              const F3 s1 = support_lane(m, ws, g1, sub | (gs << 8), gmask, ox, oy, d[0], d[1], d[2]);
              const F3 s2 = support_lane(m, ws, g2, sub | (gs << 8), gmask, ox, oy, -d[0], -d[1], -d[2]);
              if ((s2.x - s1.x) * d[0] + (s2.y - s1.y) * d[1] + (s2.z - s1.z) * d[2] > 1e-5f) res = -2;
            }
          }
#if defined(COSIM_PHASE_TIMING) && !defined(COSIM_HOST_EMU)
          if (res == 1 && sub == 0) atomicAdd(m.phase + PH_SUPPORT_CALLS, 1ull);       // counter slot reused: geom-geom MPR queries
#endif
          if (res == 1) res = mpr_lane(A, m, ws, g2, sub | (gs << 8), gmask, ox, oy, gc, &depth, nrm, cp);
          if (res == 0 && !(nrm[0] == 0.f && nrm[1] == 0.f && nrm[2] == 0.f) && depth == depth) {
            hit = (sub == 0); cp[0] += ox; cp[1] += oy;
            fix_normal(m, ws, g1, g2, cp, nrm);
          } else if (res == -1 && sub == 0) {          // remember the separating direction (in geom 1's frame)
            const float* R1 = WS(W_GXMAT) + 9 * g1; float* ax = WS(W_PAXIS) + 4 * (pp % PAXIS_SLOTS);
            ax[0] = __int_as_float_emu(pp + 1);
            ax[1] = R1[0] * nrm[0] + R1[3] * nrm[1] + R1[6] * nrm[2]; ax[2] = R1[1] * nrm[0] + R1[4] * nrm[1] + R1[7] * nrm[2]; ax[3] = R1[2] * nrm[0] + R1[5] * nrm[1] + R1[8] * nrm[2];
          }
        }
      }
#ifdef COSIM_HOST_EMU
      const int nhit = hit, slot = ncon;
#else
      int nhit, slot;
      if (BB) {
        int incl = hit;          // contacts are appended in pair order: exclusive prefix sum of the per-lane counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        nhit = __shfl_sync(0xffffffffu, incl, 31); slot = ncon + incl - hit;
      } else {                   // at most one contact per pair
        const unsigned hits = __ballot_sync(0xffffffffu, hit);
        nhit = __popc(hits); slot = ncon + __popc(hits & ((1u << lane) - 1u));
      }
#endif
      if (!BB || !bbout) { if (hit && slot < MD(ncon_max)) write_contact(m, ws, slot, cp, nrm, -depth, fmaxf(WS(W_GMU)[g1], WS(W_GMU)[g2]), TB(geom_body)[g2], g2, -2 - g1); }
      else NOUNROLL for (int c = 0; c < hit; ++c) if (slot + c < MD(ncon_max)) {
        const float* o = bbout + 7 * c; const float wp[3] = {o[0] + bbx, o[1] + bby, o[2]};
        write_contact(m, ws, slot + c, wp, o + 3, o[6], fmaxf(WS(W_GMU)[g1], WS(W_GMU)[g2]), TB(geom_body)[g2], g2, -2 - g1);
      }
      const int room = imax(0, MD(ncon_max) - ncon);
      dropped += imax(0, nhit - room); ncon += imin(nhit, room);
      SYNC();       // cached axes written above are read by the next chunk's candidate test
    }
  }
  SYNC();
  if (lane == 0) { WSI(W_CNT)[CNT_NCON] = ncon; WSI(W_CNT)[CNT_DROPPED] = dropped; }
  SYNC();
}

DEV_NOINLINE void collide_plane(const ModelDev& m, WSP ws, int g, int& ncon, int& dropped, int lane) { LANE_REFRESH();
  const GeomW G = make_geom(m, ws, g);
  const float n[3] = {0.f, 0.f, 1.f};
  if (G.type == GEOM_SPHERE) {
    float dist = G.pos[2] - G.size[0];
    if (dist > 0.f) return;
    float p[3]; v3addscl(p, G.pos, n, -(G.size[0] + dist * 0.5f));
    add_contact(m, ws, ncon, dropped, p, n, dist, g, -1, lane);
  } else if (G.type == GEOM_CYLINDER) {
    float axis[3] = {G.mat[2], G.mat[5], G.mat[8]};
    float prjaxis = v3dot(n, axis);
    if (prjaxis > 0.f) { v3scl(axis, axis, -1.f); prjaxis = -prjaxis; }
    float dist0 = G.pos[2];
    float vec[3]; v3scl(vec, axis, prjaxis); v3sub(vec, vec, n);
    float len2 = v3dot(vec, vec);
    if (len2 >= MINVALF * MINVALF) v3scl(vec, vec, G.size[0] / sqrtf(len2));
    else { vec[0] = G.mat[0] * G.size[0]; vec[1] = G.mat[3] * G.size[0]; vec[2] = G.mat[6] * G.size[0]; }
    float prjvec = v3dot(vec, n);
    v3scl(axis, axis, G.size[1]); prjaxis *= G.size[1];
    if (dist0 + prjaxis + prjvec > 0.f) return;
    float dist = dist0 + prjaxis + prjvec, p[3];
    for (int k = 0; k < 3; ++k) p[k] = G.pos[k] + vec[k] + axis[k] - n[k] * dist * 0.5f;
    add_contact(m, ws, ncon, dropped, p, n, dist, g, -1, lane);
    if (dist0 - prjaxis + prjvec <= 0.f) {
      dist = dist0 - prjaxis + prjvec;
      for (int k = 0; k < 3; ++k) p[k] = G.pos[k] + vec[k] - axis[k] - n[k] * dist * 0.5f;
      add_contact(m, ws, ncon, dropped, p, n, dist, g, -1, lane);
    }
    float prjvec1 = -prjvec * 0.5f;
    if (dist0 + prjaxis + prjvec1 <= 0.f) {
      float vec1[3]; v3cross(vec1, vec, axis); v3normalize(vec1); v3scl(vec1, vec1, G.size[0] * sqrtf(3.f) * 0.5f);
      dist = dist0 + prjaxis + prjvec1;
      for (int k = 0; k < 3; ++k) p[k] = G.pos[k] + vec1[k] + axis[k] - vec[k] * 0.5f - n[k] * dist * 0.5f;
      add_contact(m, ws, ncon, dropped, p, n, dist, g, -1, lane);
      for (int k = 0; k < 3; ++k) p[k] = G.pos[k] - vec1[k] + axis[k] - vec[k] * 0.5f - n[k] * dist * 0.5f;
      add_contact(m, ws, ncon, dropped, p, n, dist, g, -1, lane);
    }
  } else if (G.type == GEOM_BOX) {
    int cnt = 0;
    NOUNROLL for (int i = 0; i < 8 && cnt < 4; ++i) {
      float lc[3] = {(i & 1 ? G.size[0] : -G.size[0]), (i & 2 ? G.size[1] : -G.size[1]), (i & 4 ? G.size[2] : -G.size[2])}, w[3], p[3];
      m3mulv(w, G.mat, lc);
      float dist = G.pos[2] + w[2];
      if (dist > 0.f) continue;
      for (int k = 0; k < 3; ++k) p[k] = G.pos[k] + w[k] - n[k] * dist * 0.5f;
      add_contact(m, ws, ncon, dropped, p, n, dist, g, -1, lane); ++cnt;
    }
  } else if (G.type == GEOM_MESH) {
    // up to 4 deepest hull vertices below the plane (ties: lowest index): 4 rounds of warp arg-min with exclusion
    int taken[4] = {-1, -1, -1, -1};
    NOUNROLL for (int c = 0; c < 4; ++c) {
      float bv = -INFINITY; int best = 0x7fffffff;
      NOUNROLL for (int i = lane; i < G.nvert; i += LANES) {
        if (i == taken[0] || i == taken[1] || i == taken[2]) continue;
        float z = G.mat[6] * LDGB(G.verts + 3 * i) + G.mat[7] * LDGB(G.verts + 3 * i + 1) + G.mat[8] * LDGB(G.verts + 3 * i + 2);
        if (-z > bv) { bv = -z; best = i; }
      }
      wargmax(bv, best);
      if (best == 0x7fffffff) break;
      float lv[3] = {LDGB(G.verts + 3 * best), LDGB(G.verts + 3 * best + 1), LDGB(G.verts + 3 * best + 2)}, w[3], p[3];
      m3mulv(w, G.mat, lv);
      float dist = G.pos[2] + w[2];
      if (dist > 0.f) break;
      taken[c] = best;
      for (int k = 0; k < 3; ++k) p[k] = G.pos[k] + w[k] - n[k] * dist * 0.5f;
      add_contact(m, ws, ncon, dropped, p, n, dist, g, -1, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------ constraints
DEV_NOINLINE float impedance(const float* s_in, float pos) {
  float s0 = fminf(0.9999f, fmaxf(1e-4f, s_in[0])), s1 = fminf(0.9999f, fmaxf(1e-4f, s_in[1]));
  float s2 = fmaxf(0.f, s_in[2]), s3 = fminf(0.9999f, fmaxf(1e-4f, s_in[3])), s4 = fmaxf(1.f, s_in[4]);
  if (s0 == s1 || s2 <= MINVALF) return 0.5f * (s0 + s1);
  float x = fabsf(pos) / s2;
  if (x >= 1.f) return s1;
  if (x <= 0.f) return s0;
  float y;
  if (s4 == 1.f) y = x;
  else if (s4 == 2.f) y = (x <= s3) ? x * x / s3 : 1.f - (1.f - x) * (1.f - x) / (1.f - s3);      // default solimp power
  else if (x <= s3) { float a = 1.f / powf(s3, s4 - 1.f); y = a * powf(x, s4); }
  else { float b = 1.f / powf(1.f - s3, s4 - 1.f); y = 1.f - b * powf(1.f - x, s4); }
  return s0 + y * (s1 - s0);
}
// K, B of the reference acceleration for given solref/solimp (refsafe)
DEV void kb_params(const ModelDev& m, const float* solref, const float* solimp, float* K, float* B) {
  float dmax = fminf(0.9999f, fmaxf(1e-4f, solimp[1]));
  float tc = fmaxf(solref[0], 2.f * MO(timestep)), dr = solref[1];
  *K = 1.f / fmaxf(MINVALF, dmax * dmax * tc * tc * dr * dr);
  *B = 2.f / fmaxf(MINVALF, dmax * tc);
}
// translational point Jacobian column for dof k at world point (offset from subtree COM), if dof k moves `body`
DEV void jac_col(const ModelDev& m, WSP ws, int body, int k, const float* off, float* jp) {
  if ((TB(body_dofmask)[body] >> k) & 1) {
    const float* cd = WS(W_CDOF) + 6 * k; float c[3]; v3cross(c, cd, off);
    jp[0] = cd[3] + c[0]; jp[1] = cd[4] + c[1]; jp[2] = cd[5] + c[2];
  } else { jp[0] = jp[1] = jp[2] = 0.f; }
}

// ------------------------------------------------------------------------------------------ Jacobian-free contact rows
// The contact Jacobian is never stored (3 x nv floats per contact would push the records of a contact-rich env out of shared
// memory: w4_p_v2 on the 9.8 mm stairs raster carries 30 - 150 contacts).  Instead
//   J v      = velocity of the contact point under the generalized velocity v  -> body spatial velocities (body_vel), then
//              3 dot products per contact (contact_edge_rows);
//   J^T f    = the contact wrenches summed per body, then projected on the motion axes of the ancestor dofs (update_forces);
//   J^T W J  = for the dof pair (i, j), j an ancestor of i: sum over the contacts of the subtree of body(i) of
//              jac_i^T W jac_j with jac_k = cdof_lin[k] + cdof_ang[k] x offset (newton_direction).
// Ground contacts are generated in geom order and geoms are numbered in body (depth-first) order, so the contacts of a body --
// and of a whole subtree -- form one contiguous range of the list: W_CSTART[b] = first contact of body b.  Geom-geom contacts
// (two bodies, appended after the ground contacts) are handled by explicit loops; they are few.
DEV void point_vel(const float* v6, const float* off, float* vp) { float c[3]; v3cross(c, v6, off); vp[0] = v6[3] + c[0]; vp[1] = v6[4] + c[1]; vp[2] = v6[5] + c[2]; }
// Up to FEW_CONTACTS contacts (the common case on coarse terrain: one or two wheel contacts) the solver keeps explicit 3 x nv
// frame Jacobians in shared memory (W_CN_J) -- cheaper than the per-body passes when there is next to nothing to sum; above
// that the Jacobian-free formulation takes over.  Both evaluate the same rows.
enum { FEW_CONTACTS = 8 };
// the four pyramid-edge projections n +- mu t1, n +- mu t2 of the contact-point velocity given the body velocities `bv` [nbody][6]
DEV void contact_edge_rows(const ModelDev& m, WSP ws, const float* rec, const float* bv, float* r) {
  float off[3], vp[3]; v3sub(off, rec + CR_POS, WS(W_SCOM));
  point_vel(bv + 6 * ((const int*)rec)[CR_BODY], off, vp);
  const int cell = ((const int*)rec)[CR_CELL];
  if (cell <= -2) { const int b1 = TB(geom_body)[-2 - cell]; if (b1 > 0) { float v1[3]; point_vel(bv + 6 * b1, off, v1); v3sub(vp, vp, v1); } }
  const float* fr = rec + CR_FRAME; const float mu = rec[CR_MU];
  const float vn = v3dot(fr, vp), v1 = mu * v3dot(fr + 3, vp), v2 = mu * v3dot(fr + 6, vp);
  r[0] = vn + v1; r[1] = vn - v1; r[2] = vn + v2; r[3] = vn - v2;
}
// per-body contact ranges, number of ground contacts, mask of the bodies in contact
DEV_NOINLINE void contact_index(const ModelDev& m, WSP ws, int ncon, int lane) { LANE_REFRESH();
  const int nb = MD(nbody);
  int ng_ = 0; unsigned mask = 0;
  NOUNROLL for (int c = lane; c < ncon; c += LANES) {
    const int* rec = CRECI(c); const int cell = rec[CR_CELL];
    ng_ += cell > -2; mask |= 1u << rec[CR_BODY];
    if (cell <= -2) mask |= 1u << TB(geom_body)[-2 - cell];
  }
#ifndef COSIM_HOST_EMU
  ng_ = __reduce_add_sync(0xffffffffu, ng_); mask = __reduce_or_sync(0xffffffffu, mask);
#endif
  const int ncg = ng_;
  int* cs = WSI(W_CSTART);
  unsigned gm = 0;
  NOUNROLL for (int b0 = 0; b0 <= nb; b0 += LANES) {
    const int b = b0 + lane; int n = 0, own = 0;
    if (b <= nb) { NOUNROLL for (int c = 0; c < ncg; ++c) { const int cb_ = CRECI(c)[CR_BODY]; n += cb_ < b; own |= cb_ == b; } cs[b] = n; }
    gm |= wballot(own) << b0;
  }
  if (lane == 0) { WSI(W_CNT)[CNT_NCG] = ncg; WSI(W_CNT)[CNT_CBMASK] = (int)(mask & ~1u); WSI(W_CNT)[CNT_CBGMASK] = (int)(gm & ~1u); }
  SYNC();
}
// W_BV[b] = spatial velocity (about the subtree COM) of body b under the generalized velocity `vec`, for the bodies in contact
DEV_NOINLINE void body_vel(const ModelDev& m, WSP ws, const float* vec, int lane) { LANE_REFRESH();
  const int nb = MD(nbody); const unsigned cb = (unsigned)WSI(W_CNT)[CNT_CBMASK];
  float* bv = WS(W_BV); const float* cdof = WS(W_CDOF);
  NOUNROLL for (int idx = lane; idx < 6 * nb; idx += LANES) {
    const int b = idx / 6, i = idx - 6 * b;
    if (!((cb >> b) & 1u)) continue;
    unsigned mask = (unsigned)TB(body_dofmask)[b]; float acc = 0.f;
    NOUNROLL while (mask) { const int k = ctz32(mask); mask &= mask - 1u; acc += cdof[6 * k + i] * vec[k]; }
    bv[idx] = acc;
  }
  SYNC();
}

DEV_NOINLINE void make_constraint(const ModelDev& m, WSP ws, int ncon, int lane) { LANE_REFRESH();
  const int nv = MD(nv), njnt = MD(njnt), neq = MD(neq);
  const float solref[2] = {MO(solref0), MO(solref1)};
  const float solimp[5] = {MO(solimp0), MO(solimp1), MO(solimp2), MO(solimp3), MO(solimp4)};
  float K, B; kb_params(m, solref, solimp, &K, &B);
  const float* qvel = WS(W_QVEL); const float* qpos = WS(W_QPOS); const float* scom = WS(W_SCOM);
  // equality: connect
  NOUNROLL for (int e = 0; e < neq; ++e) {
    const int b1 = TB(eq_body1)[e], b2 = TB(eq_body2)[e];
    float a1[3] = {LDG(TB(eq_anchor1) + 3 * e), LDG(TB(eq_anchor1) + 3 * e + 1), LDG(TB(eq_anchor1) + 3 * e + 2)};
    float a2[3] = {LDG(TB(eq_anchor2) + 3 * e), LDG(TB(eq_anchor2) + 3 * e + 1), LDG(TB(eq_anchor2) + 3 * e + 2)};
    float p1[3], p2[3], r[3], o1[3], o2[3];
    m3mulv(r, WS(W_XMAT) + 9 * b1, a1); v3add(p1, WS(W_XPOS) + 3 * b1, r);
    m3mulv(r, WS(W_XMAT) + 9 * b2, a2); v3add(p2, WS(W_XPOS) + 3 * b2, r);
    v3sub(o1, p1, scom); v3sub(o2, p2, scom);
    float* J = WS(W_EQ_J) + (size_t)3 * e * nv;
    float vel[3] = {0.f, 0.f, 0.f};
    FOR_LANE(k, nv) {
      float j1[3], j2[3]; jac_col(m, ws, b1, k, o1, j1); jac_col(m, ws, b2, k, o2, j2);
      for (int r_ = 0; r_ < 3; ++r_) { float v = j1[r_] - j2[r_]; J[r_ * nv + k] = v; vel[r_] += v * qvel[k]; }
    }
    vel[0] = wsum(vel[0]); vel[1] = wsum(vel[1]); vel[2] = wsum(vel[2]);
    float pe[3]; v3sub(pe, p1, p2);
    float si[5]; for (int i = 0; i < 5; ++i) si[i] = LDG(TB(eq_solimp) + 5 * e + i);
    float sr[2] = {LDG(TB(eq_solref) + 2 * e), LDG(TB(eq_solref) + 2 * e + 1)};
    float imp = impedance(si, v3norm(pe)), Ke, Be; kb_params(m, sr, si, &Ke, &Be);
    float diag = WS(W_INVWB)[b1] + WS(W_INVWB)[b2];
    float R = fmaxf(MINVALF, (1.f - imp) * diag / imp);
    FOR_LANE(i, 3) { WS(W_EQ_D)[3 * e + i] = 1.f / R; WS(W_EQ_AREF)[3 * e + i] = -Be * vel[i] - Ke * imp * pe[i]; }
  }
  // dof friction loss: D = 0 marks "no row"
  FOR_LANE(k, nv) {
    float fl = WS(W_FLOSS)[k];
    if (fl > 0.f) {
      float imp = impedance(solimp, 0.f);
      float R = fmaxf(MINVALF, (1.f - imp) * WS(W_INVWD)[k] / imp);
      WS(W_FR_D)[k] = 1.f / R; WS(W_FR_AREF)[k] = -B * qvel[k];
    } else { WS(W_FR_D)[k] = 0.f; WS(W_FR_AREF)[k] = 0.f; }
  }
  // joint limits: at most one side active per joint (lo < hi); sign 0 marks "no row"
  FOR_LANE(j, njnt) {
    float sign = 0.f, D = 0.f, aref = 0.f;
    if (TB(jnt_limited)[j]) {
      const float q = qpos[TB(jnt_qposadr)[j]]; const int k = TB(jnt_dofadr)[j];
      float dlo = q - LDG(TB(jnt_range) + 2 * j), dhi = LDG(TB(jnt_range) + 2 * j + 1) - q;
      float dist = 0.f;
      if (dlo < 0.f) { sign = 1.f; dist = dlo; } else if (dhi < 0.f) { sign = -1.f; dist = dhi; }
      if (sign != 0.f) {
        float imp = impedance(solimp, dist);
        float R = fmaxf(MINVALF, (1.f - imp) * WS(W_INVWD)[k] / imp);
        D = 1.f / R; aref = -B * (sign * qvel[k]) - K * imp * dist;
      }
    }
    WS(W_LM_SIGN)[j] = sign; WS(W_LM_D)[j] = D; WS(W_LM_AREF)[j] = aref;
  }
  if (IF_GENERAL(m)) { SYNC(); return; }      // contact rows of the general path: gen_make_rows()
  // contacts: 4 pyramid edges share D = 1/(2 mu^2 R_first)
  if (ncon <= FEW_CONTACTS) {        // few contacts: explicit 3 x nv frame Jacobians in W_CN_J
    NOUNROLL for (int idx = lane; idx < ncon * nv; idx += LANES) {
      const int c = idx / nv, k = idx - c * nv;
      const float* rec = CRECS(c);
      const float* fr = rec + CR_FRAME; float off[3], jp[3];
      v3sub(off, rec + CR_POS, scom);
      jac_col(m, ws, ((const int*)rec)[CR_BODY], k, off, jp);
      { const int cell = ((const int*)rec)[CR_CELL]; if (cell <= -2) { const int b1 = TB(geom_body)[-2 - cell]; if (b1 > 0) { float j1[3]; jac_col(m, ws, b1, k, off, j1); v3sub(jp, jp, j1); } } }
      float* J = WS(W_CN_J) + (size_t)3 * c * nv;
      J[k] = v3dot(fr, jp); J[nv + k] = v3dot(fr + 3, jp); J[2 * nv + k] = v3dot(fr + 6, jp);
    }
    SYNC();
    NOUNROLL for (int idx = lane; idx < ncon * 4; idx += LANES) {
      const int c = idx >> 2, e = idx & 3;
      float* rec = CRECS(c);
      const float* J = WS(W_CN_J) + (size_t)3 * c * nv; const float* Jt = J + (1 + (e >> 1)) * nv;
      const float mu = rec[CR_MU], sg = (e & 1) ? -mu : mu, dist = rec[CR_DIST];
      float vel = 0.f;
      NOUNROLL for (int k = 0; k < nv; ++k) vel += (J[k] + sg * Jt[k]) * qvel[k];
      const float imp = impedance(solimp, dist);
      rec[CR_AREF + e] = -B * vel - K * imp * dist;
      if (e == 0) {
        float tran = WS(W_INVWB)[((const int*)rec)[CR_BODY]];
        { const int b1 = contact_body1(m, ws, c); if (b1 > 0) tran += WS(W_INVWB)[b1]; }
        const float R = fmaxf(MINVALF, (1.f - imp) * (tran + mu * mu * tran) / imp);
        rec[CR_D] = 1.f / (2.f * mu * mu * R);
      }
    }
    SYNC();
    return;
  }
  // many contacts: no Jacobian; the edge velocities come from the body velocities
  contact_index(m, ws, ncon, lane);
  const float* cvel = WS(W_CVEL);
  NOUNROLL for (int c = lane; c < ncon; c += LANES) {
    float* rec = CREC(c);
    float r[4]; contact_edge_rows(m, ws, rec, cvel, r);
    const float mu = rec[CR_MU], dist = rec[CR_DIST];
    const float imp = impedance(solimp, dist);
#pragma unroll
    for (int e = 0; e < 4; ++e) rec[CR_AREF + e] = -B * r[e] - K * imp * dist;
    float tran = WS(W_INVWB)[((const int*)rec)[CR_BODY]];
    { const int b1 = contact_body1(m, ws, c); if (b1 > 0) tran += WS(W_INVWB)[b1]; }
    const float R = fmaxf(MINVALF, (1.f - imp) * (tran + mu * mu * tran) / imp);
    rec[CR_D] = 1.f / (2.f * mu * mu * R);
  }
  SYNC();
}

// ------------------------------------------------------------------------------------------ Newton solver
// evaluates rows at x0 + alpha*dx; returns per-lane partial sums of (cost, d0, d1); writes nothing.
// X arrays hold J*qacc - aref ("Jaref") per row, V arrays hold J*search.
struct RowSum { float cost, d0, d1; };
DEV void row_acc(RowSum& s, float x0, float jv, float a, float D, int kind, float Rf, float f) {
  const float x = x0 + a * jv;
  if (kind == 1) {           // friction loss (Huber)
    if (x <= -Rf) { s.cost += f * (-0.5f * Rf - x); s.d0 -= f * jv; return; }
    if (x >= Rf) { s.cost += f * (-0.5f * Rf + x); s.d0 += f * jv; return; }
  } else if (kind == 2) {    // inequality (limit, pyramid edge)
    if (x >= 0.f) return;
  }
  s.cost += 0.5f * D * x * x; s.d0 += D * x * jv; s.d1 += D * jv * jv;
}
DEV_NOINLINE RowSum eval_rows(const ModelDev& m, WSP ws, int ncon, float a, bool use_v, int lane) { LANE_REFRESH();
  const int nv = MD(nv), njnt = MD(njnt), neq3 = 3 * MD(neq);
  RowSum s = {0.f, 0.f, 0.f};
  FOR_LANE(i, neq3) row_acc(s, WS(W_EQ_X)[i], use_v ? WS(W_EQ_V)[i] : 0.f, a, WS(W_EQ_D)[i], 0, 0.f, 0.f);
  FOR_LANE(k, nv) {
    const float D = WS(W_FR_D)[k];
    if (D > 0.f) { const float f = WS(W_FLOSS)[k]; row_acc(s, WS(W_TMPW)[k], use_v ? WS(W_SEARCH)[k] : 0.f, a, D, 1, f / D, f); }
  }
  FOR_LANE(j, njnt) {
    const float sg = WS(W_LM_SIGN)[j];
    if (sg != 0.f) { const int k = TB(jnt_dofadr)[j]; row_acc(s, sg * WS(W_QACC)[k] - WS(W_LM_AREF)[j], use_v ? sg * WS(W_SEARCH)[k] : 0.f, a, WS(W_LM_D)[j], 2, 0.f, 0.f); }
  }
  if (ncon <= FEW_CONTACTS) { NOUNROLL for (int idx = lane; idx < 4 * ncon; idx += LANES) { const float* rec = CRECS(idx >> 2); const int e = idx & 3; row_acc(s, rec[CR_X + e], use_v ? rec[CR_V + e] : 0.f, a, rec[CR_D], 2, 0.f, 0.f); } }
  else NOUNROLL for (int idx = lane; idx < 4 * ncon; idx += LANES) { const float* rec = CREC(idx >> 2); const int e = idx & 3; row_acc(s, rec[CR_X + e], use_v ? rec[CR_V + e] : 0.f, a, rec[CR_D], 2, 0.f, 0.f); }
  return s;
}
// X = J*q - aref for every row class, given q (length nv).  friction rows' X go to W_TMPW.
DEV_NOINLINE void compute_jaref(const ModelDev& m, WSP ws, int ncon, const float* q, int lane) { LANE_REFRESH();
  const int nv = MD(nv), neq3 = 3 * MD(neq);
  FOR_LANE(i, neq3) { const float* J = WS(W_EQ_J) + (size_t)i * nv; float s = 0.f; NOUNROLL for (int k = 0; k < nv; ++k) s += J[k] * q[k]; WS(W_EQ_X)[i] = s - WS(W_EQ_AREF)[i]; }
  FOR_LANE(k, nv) WS(W_TMPW)[k] = q[k] - WS(W_FR_AREF)[k];
  if (ncon > FEW_CONTACTS) {
    body_vel(m, ws, q, lane);
    NOUNROLL for (int c = lane; c < ncon; c += LANES) {
      float* rec = CREC(c); float r[4]; contact_edge_rows(m, ws, rec, WS(W_BV), r);
#pragma unroll
      for (int e = 0; e < 4; ++e) rec[CR_X + e] = r[e] - rec[CR_AREF + e];
    }
  } else {
    NOUNROLL for (int idx = lane; idx < 4 * ncon; idx += LANES) {
      const int c = idx >> 2, e = idx & 3;
      float* rec = CRECS(c);
      const float* J = WS(W_CN_J) + (size_t)3 * c * nv; const float* Jt = J + (1 + (e >> 1)) * nv;
      const float mu = rec[CR_MU], sg = (e & 1) ? -mu : mu;
      float s = 0.f; NOUNROLL for (int k = 0; k < nv; ++k) s += (J[k] + sg * Jt[k]) * q[k];
      rec[CR_X + e] = s - rec[CR_AREF + e];
    }
  }
  SYNC();
}
DEV_NOINLINE void compute_jv(const ModelDev& m, WSP ws, int ncon, const float* v, int lane) { LANE_REFRESH();
  const int nv = MD(nv), neq3 = 3 * MD(neq);
  FOR_LANE(i, neq3) { const float* J = WS(W_EQ_J) + (size_t)i * nv; float s = 0.f; NOUNROLL for (int k = 0; k < nv; ++k) s += J[k] * v[k]; WS(W_EQ_V)[i] = s; }
  if (ncon > FEW_CONTACTS) {
    body_vel(m, ws, v, lane);
    NOUNROLL for (int c = lane; c < ncon; c += LANES) {
      float* rec = CREC(c); float r[4]; contact_edge_rows(m, ws, rec, WS(W_BV), r);
#pragma unroll
      for (int e = 0; e < 4; ++e) rec[CR_V + e] = r[e];
    }
  } else {
    NOUNROLL for (int idx = lane; idx < 4 * ncon; idx += LANES) {
      const int c = idx >> 2, e = idx & 3;
      float* rec = CRECS(c);
      const float* J = WS(W_CN_J) + (size_t)3 * c * nv; const float* Jt = J + (1 + (e >> 1)) * nv;
      const float mu = rec[CR_MU], sg = (e & 1) ? -mu : mu;
      float s = 0.f; NOUNROLL for (int k = 0; k < nv; ++k) s += (J[k] + sg * Jt[k]) * v[k];
      rec[CR_V + e] = s;
    }
  }
  SYNC();
}
DEV_NOINLINE void mat_vec(const float* M, const float* x, float* y, int n, int lane) { LANE_REFRESH();
  FOR_LANE(i, n) { float s = 0.f; NOUNROLL for (int k = 0; k < n; ++k) s += M[i * n + k] * x[k]; y[i] = s; }
  SYNC();
}
// constraint forces from current X; qfrc_constraint -> W_FCON; contact frame forces / world wrenches -> the records; returns constraint cost
DEV_NOINLINE float update_forces(const ModelDev& m, WSP ws, int ncon, int lane) { LANE_REFRESH();
  const int nv = MD(nv), njnt = MD(njnt), neq = MD(neq), nb = MD(nbody);
  float cost = 0.f;
  FOR_LANE(i, 3 * neq) { const float x = WS(W_EQ_X)[i], D = WS(W_EQ_D)[i]; WS(W_EQ_F)[i] = -D * x; cost += 0.5f * D * x * x; }
  NOUNROLL for (int c = lane; c < ncon; c += LANES) {
    float* rec = CREC(c);
    const float D = rec[CR_D], mu = rec[CR_MU]; float f[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float x = rec[CR_X + e]; f[e] = x < 0.f ? -D * x : 0.f; if (x < 0.f) cost += 0.5f * D * x * x; }
    const float F[3] = {f[0] + f[1] + f[2] + f[3], mu * (f[0] - f[1]), mu * (f[2] - f[3])};
    rec[CR_F] = F[0]; rec[CR_F + 1] = F[1]; rec[CR_F + 2] = F[2];
    if (ncon > FEW_CONTACTS) {
      float wf[3], off[3], tq[3];
      m3tmulv(wf, rec + CR_FRAME, F); v3sub(off, rec + CR_POS, WS(W_SCOM)); v3cross(tq, off, wf);
      rec[CR_WF] = tq[0]; rec[CR_WF + 1] = tq[1]; rec[CR_WF + 2] = tq[2]; rec[CR_WF + 3] = wf[0]; rec[CR_WF + 4] = wf[1]; rec[CR_WF + 5] = wf[2];
    }
  }
  SYNC();
  const unsigned cb = ncon > FEW_CONTACTS ? (unsigned)WSI(W_CNT)[CNT_CBMASK] : 0u;
  if (cb) {      // wrench per body (fixed summation order: contact order)
    const int* cs = WSI(W_CSTART); const int ncg = WSI(W_CNT)[CNT_NCG]; float* bf = WS(W_BV);
    NOUNROLL for (int idx = lane; idx < 6 * nb; idx += LANES) {
      const int b = idx / 6, i = idx - 6 * b;
      if (!((cb >> b) & 1u)) continue;
      float acc = 0.f;
      NOUNROLL for (int c = cs[b]; c < cs[b + 1]; ++c) acc += CREC(c)[CR_WF + i];
      NOUNROLL for (int c = ncg; c < ncon; ++c) {
        const float* rec = CREC(c);
        if (((const int*)rec)[CR_BODY] == b) acc += rec[CR_WF + i];
        if (TB(geom_body)[-2 - ((const int*)rec)[CR_CELL]] == b) acc -= rec[CR_WF + i];
      }
      bf[idx] = acc;
    }
    SYNC();
  }
  FOR_LANE(k, nv) {
    float q = 0.f;
    const float D = WS(W_FR_D)[k];
    if (D > 0.f) {
      const float x = WS(W_TMPW)[k], f = WS(W_FLOSS)[k], Rf = f / D;
      if (x <= -Rf) { q += f; cost += f * (-0.5f * Rf - x); }
      else if (x >= Rf) { q -= f; cost += f * (-0.5f * Rf + x); }
      else { q -= D * x; cost += 0.5f * D * x * x; }
    }
    const int j = TB(dof_jnt)[k];
    const float sg = WS(W_LM_SIGN)[j];
    if (sg != 0.f && TB(jnt_dofadr)[j] == k) {
      const float x = sg * WS(W_QACC)[k] - WS(W_LM_AREF)[j], Dl = WS(W_LM_D)[j];
      if (x < 0.f) { q += sg * (-Dl * x); cost += 0.5f * Dl * x * x; }
    }
    if (cb) {      // J^T f: the wrenches of the bodies below dof k, projected on its motion axis
      const int b0 = TB(dof_body)[k], b1 = b0 + TB(body_subsize)[b0];
      const float* cd = WS(W_CDOF) + 6 * k; const float* bf = WS(W_BV);
      NOUNROLL for (int b = b0; b < b1; ++b) if ((cb >> b) & 1u) {
        const float* w = bf + 6 * b;
        q += cd[0] * w[0] + cd[1] * w[1] + cd[2] * w[2] + cd[3] * w[3] + cd[4] * w[4] + cd[5] * w[5];
      }
    }
    if (!cb) {      // few contacts: J^T f from the explicit Jacobians
      NOUNROLL for (int c = 0; c < ncon; ++c) {
        const float* J = WS(W_CN_J) + (size_t)3 * c * nv; const float* F = CRECS(c) + CR_F;
        q += J[k] * F[0] + J[nv + k] * F[1] + J[2 * nv + k] * F[2];
      }
    }
    NOUNROLL for (int i = 0; i < 3 * neq; ++i) q += WS(W_EQ_J)[(size_t)i * nv + k] * WS(W_EQ_F)[i];
    WS(W_FCON)[k] = q;
  }
  SYNC();
  (void)njnt;
  return wsum(cost);
}
// total cost at q (Gauss + constraints); leaves X / forces for q
DEV_NOINLINE float total_cost(const ModelDev& m, WSP ws, int ncon, const float* q, float* Mq, int lane) { LANE_REFRESH();
  const int nv = MD(nv);
  mat_vec(WS(W_M), q, Mq, nv, lane);
  if (q != WS(W_QACC)) { FOR_LANE(k, nv) WS(W_QACC)[k] = q[k]; SYNC(); }
  compute_jaref(m, ws, ncon, q, lane);
  float c = update_forces(m, ws, ncon, lane);
  float g = 0.f;
  FOR_LANE(k, nv) g += (Mq[k] - WS(W_FSMOOTH)[k]) * (q[k] - WS(W_ASMOOTH)[k]);
  return c + 0.5f * wsum(g);
}

// Newton with exact line search on the primal cost [SURVEY.md B.8]; same control flow as the
// oracle's solve()/linesearch() (oracle/oracle.hpp), lanes split the rows inside every evaluation.
// fp32 note: the reference's line-search gradient tolerance (tolerance * ls_tolerance * |search| *
// meaninertia * nv ~ 1e-10) sits far below fp32 round-off of the derivative, so it is floored at a
// few ulps of the magnitude of the terms the derivative is summed from.
#ifndef COST_EPS
#define COST_EPS 5e-7f
#endif
#ifndef GRAD_EPS
#define GRAD_EPS 1e-6f
#endif
#ifdef COSIM_HOST_EMU
static long g_emu_ls_evals = 0;      // analysis aid of the host emulation (tests/hostsim)
#endif
struct LSPoint { float alpha, cost, d0, d1; };
#if COSIM_GENERAL
DEV_NOINLINE RowSum gen_eval_rows(const ModelDev& m, WSP ws, int nefc, float a, int lane);      // engine_general.h
#endif
DEV_NOINLINE LSPoint ls_eval(const ModelDev& m, WSP ws, int ncon, float a, float q0, float q1, float q2, int lane) { LANE_REFRESH();
#if COSIM_GENERAL
  RowSum s = m.general ? gen_eval_rows(m, ws, ncon, a, lane) : eval_rows(m, ws, ncon, a, true, lane);       // general path: `ncon` carries the row count
#else
  RowSum s = eval_rows(m, ws, ncon, a, true, lane);
#endif
  PH_COUNT(PH_LS_EVALS, 1);
#ifdef COSIM_HOST_EMU
  ++g_emu_ls_evals;
#endif
  LSPoint p; p.alpha = a;
  p.cost = a * a * q2 + a * q1 + q0 + wsum(s.cost);
  p.d0 = 2.f * a * q2 + q1 + wsum(s.d0);
  p.d1 = 2.f * q2 + wsum(s.d1);
  if (p.d1 <= 0.f) p.d1 = MINVALF;
  return p;
}
DEV int ls_update(LSPoint& p, const LSPoint* cand) {
  int flag = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (p.d0 < 0.f && cand[i].d0 < 0.f && p.d0 < cand[i].d0) { p = cand[i]; flag = 1; }
    else if (p.d0 > 0.f && cand[i].d0 > 0.f && p.d0 > cand[i].d0) { p = cand[i]; flag = 2; }
  }
  return flag;
}
DEV_NOINLINE float linesearch(const ModelDev& m, WSP ws, int ncon, float gauss, float q1, float q2, float snorm, float absterms, int lane) { LANE_REFRESH();
  if (snorm < MINVALF) return 0.f;
  const float gtol = fmaxf(MO(tolerance) * MO(ls_tolerance) * snorm * WS(W_SCAL)[1] * (float)imax(1, MD(nv)), 1e-6f * absterms);
  const int maxit = MD(ls_iterations);
  int iter = 0;
#define LS_EVAL(a) (++iter, ls_eval(m, ws, ncon, (a), gauss, q1, q2, lane))
  const LSPoint p0 = LS_EVAL(0.f);
  LSPoint p1 = LS_EVAL(p0.alpha - p0.d0 / p0.d1);
  if (p0.cost < p1.cost) p1 = p0;
  if (fabsf(p1.d0) < gtol) return p1.alpha;
  const float dir = p1.d0 < 0.f ? 1.f : -1.f;
  LSPoint p2 = p1; bool p2upd = false;
  while (p1.d0 * dir <= -gtol && iter < maxit) {
    p2 = p1; p2upd = true;
    p1 = LS_EVAL(p1.alpha - p1.d0 / p1.d1);
    if (fabsf(p1.d0) < gtol) return p1.alpha;
  }
  if (iter >= maxit || !p2upd) return p1.alpha;
  LSPoint p2next = p1, p1next = LS_EVAL(p1.alpha - p1.d0 / p1.d1);
  while (iter < maxit) {
    const LSPoint pmid = LS_EVAL(0.5f * (p1.alpha + p2.alpha));
    const LSPoint cand[3] = {p1next, p2next, pmid};
    int best = -1; float bc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) if (fabsf(cand[i].d0) < gtol && (best == -1 || cand[i].cost < bc)) { bc = cand[i].cost; best = i; }
    if (best >= 0) return cand[best].alpha;
    const int b1 = ls_update(p1, cand); if (b1) p1next = LS_EVAL(p1.alpha - p1.d0 / p1.d1);
    const int b2 = ls_update(p2, cand); if (b2) p2next = LS_EVAL(p2.alpha - p2.d0 / p2.d1);
    if (!b1 && !b2) return pmid.alpha;
  }
#undef LS_EVAL
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0.f;
}

// grad = Ma - qfrc_smooth - qfrc_constraint; H = M + J' D_quad J; search = -H^-1 grad.  Returns |grad|.
// signature of the set of rows in their quadratic zone (the only thing besides M the Hessian depends on)
DEV uint32_t wxor(uint32_t v) {
#ifndef COSIM_HOST_EMU
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
DEV uint32_t active_set_signature(const ModelDev& m, WSP ws, int ncon, int lane) {
  const int nv = MD(nv), njnt = MD(njnt);
  const float* qacc = WS(W_QACC);
  uint32_t h = 0;
  NOUNROLL for (int idx = lane; idx < 4 * ncon; idx += LANES) if (CREC(idx >> 2)[CR_X + (idx & 3)] < 0.f) h ^= (uint32_t)(idx + 1) * 2654435761u;
  FOR_LANE(k, nv) {
    const float D = WS(W_FR_D)[k];
    if (D > 0.f) { const float x = WS(W_TMPW)[k], Rf = WS(W_FLOSS)[k] / D; if (x > -Rf && x < Rf) h ^= (uint32_t)(k + 1001) * 2246822519u; }
  }
  FOR_LANE(j, njnt) {
    const float sg = WS(W_LM_SIGN)[j];
    if (sg != 0.f && (sg * qacc[TB(jnt_dofadr)[j]] - WS(W_LM_AREF)[j]) < 0.f) h ^= (uint32_t)(j + 2001) * 3266489917u;
  }
  return wxor(h) | 1u;       // never 0: 0 means "no factor yet"
}
// `sig` carries the signature of the factor currently held in W_A: when the active set did not change since the previous
// iteration the Hessian is the same matrix and only the triangular solves are repeated
DEV_NOINLINE float newton_direction(const ModelDev& m, WSP ws, int ncon, int lane, uint32_t& sig, float* termnorm = nullptr) { LANE_REFRESH();
  const int nv = MD(nv), neq = MD(neq);
  float* grad = WS(W_GRAD); float* H = WS(W_A); const float* M = WS(W_M); const float* qacc = WS(W_QACC); const float* Ma = WS(W_MA);
  float gn = 0.f, fn = 0.f;
  FOR_LANE(k, nv) {
    const float a = Ma[k], b = WS(W_FSMOOTH)[k], c = WS(W_FCON)[k], g = a - b - c;
    grad[k] = g; gn += g * g; fn += a * a + b * b + c * c; WS(W_TMPV)[k] = -g;
  }
  gn = sqrtf(wsum(gn));
  if (termnorm) *termnorm = sqrtf(wsum(fn));
  const uint32_t now = active_set_signature(m, ws, ncon, lane);
  if (now == sig) {
    SYNC();
    chol_solve(H, WS(W_INVD), WS(W_TMPV), WS(W_BUF), WS(W_SEARCH), nv, lane);
    return gn;
  }
  sig = now;
  const int npair = (nv * (nv + 1)) >> 1;
  const bool few = ncon <= FEW_CONTACTS;
  // a geom-geom contact couples the dofs of two branches: the Hessian then has entries outside the tree pattern
  int coupled;
  if (few) { int cc_ = 0; NOUNROLL for (int c = lane; c < ncon; c += LANES) cc_ |= (((const int*)CRECS(c))[CR_CELL] <= -2); coupled = wor(cc_); }
  else coupled = WSI(W_CNT)[CNT_NCG] < ncon;
  const int* cs = WSI(W_CSTART); const int ncg = few ? 0 : WSI(W_CNT)[CNT_NCG];
  const unsigned cbg = few ? 0u : (unsigned)WSI(W_CNT)[CNT_CBGMASK];
  const float* scom = WS(W_SCOM); const float* cdof = WS(W_CDOF);
  float* BS = WS(W_BS);
  // per-contact 3x3 weight in the contact frame (sum over the active pyramid edges), once per contact.  Few contacts: kept
  // as (gnn, gn1, gn2, g11) in the V slots of the record (free here).  Many: its world-frame image WW = R^T Wf R (rows of
  // R = normal, t1, t2); WW[0] < 0 marks a contact without an active edge
  NOUNROLL for (int c = lane; c < ncon; c += LANES) {
    float* rec = CREC(c);
    const float* X = rec + CR_X;
    const float s0 = X[0] < 0.f, s1 = X[1] < 0.f, s2 = X[2] < 0.f, s3 = X[3] < 0.f;
    const float D = rec[CR_D], mu = rec[CR_MU];
    const float gnn = D * (s0 + s1 + s2 + s3), gn1 = D * mu * (s0 - s1), gn2 = D * mu * (s2 - s3), g11 = D * mu * mu * (s0 + s1), g22 = mu * mu * gnn - g11;
    if (few) { float* cw = rec + CR_V; cw[0] = gnn; cw[1] = gn1; cw[2] = gn2; cw[3] = g11; continue; }
    float* W = rec + CR_WW;
    if (gnn == 0.f) { W[0] = -1.f; continue; }
    const float* n = rec + CR_FRAME; const float* t1 = n + 3; const float* t2 = n + 6;
    float a[3], b[3], cc[3];       // Wf R, row by row: a = gnn n + gn1 t1 + gn2 t2, b = gn1 n + g11 t1, cc = gn2 n + g22 t2
#pragma unroll
    for (int k = 0; k < 3; ++k) { a[k] = gnn * n[k] + gn1 * t1[k] + gn2 * t2[k]; b[k] = gn1 * n[k] + g11 * t1[k]; cc[k] = gn2 * n[k] + g22 * t2[k]; }
    W[0] = fmaxf(0.f, n[0] * a[0] + t1[0] * b[0] + t2[0] * cc[0]); W[1] = n[1] * a[1] + t1[1] * b[1] + t2[1] * cc[1]; W[2] = n[2] * a[2] + t1[2] * b[2] + t2[2] * cc[2];
    W[3] = n[0] * a[1] + t1[0] * b[1] + t2[0] * cc[1]; W[4] = n[0] * a[2] + t1[0] * b[2] + t2[0] * cc[2]; W[5] = n[1] * a[2] + t1[1] * b[2] + t2[1] * cc[2];
  }
  SYNC();
  // Many contacts: the ground contacts of one body enter the Hessian only through S_b = sum_c X_c^T WW_c X_c, a symmetric 6 x 6
  // "contact inertia" about the subtree COM (X_c = [A_c | 1] maps the body's spatial velocity to the velocity of the contact
  // point, A_c w = w x offset): J^T W J for the dof pair (i, j) is cdof_i^T (sum of S_b over the bodies below dof i) cdof_j.  The
  // cost of the assembly no longer grows with (contacts x dof pairs): 21 sums per body in contact, then 21 terms per pair and body.
  if (cbg) {
    const int nb = MD(nbody);
    NOUNROLL for (int idx = lane; idx < 21 * nb; idx += LANES) {
      const int b = idx / 21, e = idx - 21 * b;
      if (!((cbg >> b) & 1u)) continue;
      int p = 0; NOUNROLL while (((p + 1) * (p + 2)) >> 1 <= e) ++p;
      const int q = e - ((p * (p + 1)) >> 1);          // packed lower triangle: e = p (p + 1) / 2 + q, q <= p
      float acc = 0.f;
      NOUNROLL for (int c = cs[b]; c < cs[b + 1]; ++c) {
        const float* rec = CREC(c); const float* W = rec + CR_WW;
        if (W[0] < 0.f) continue;
        const float ox_ = rec[CR_POS] - scom[0], oy_ = rec[CR_POS + 1] - scom[1], oz_ = rec[CR_POS + 2] - scom[2];
        // columns of X: 0..2 = columns of A (w x offset), 3..5 = unit vectors
        const float xq0 = q == 0 ? 0.f : (q == 1 ? oz_ : (q == 2 ? -oy_ : (q == 3 ? 1.f : 0.f)));
        const float xq1 = q == 0 ? -oz_ : (q == 1 ? 0.f : (q == 2 ? ox_ : (q == 4 ? 1.f : 0.f)));
        const float xq2 = q == 0 ? oy_ : (q == 1 ? -ox_ : (q == 2 ? 0.f : (q == 5 ? 1.f : 0.f)));
        const float xp0 = p == 0 ? 0.f : (p == 1 ? oz_ : (p == 2 ? -oy_ : (p == 3 ? 1.f : 0.f)));
        const float xp1 = p == 0 ? -oz_ : (p == 1 ? 0.f : (p == 2 ? ox_ : (p == 4 ? 1.f : 0.f)));
        const float xp2 = p == 0 ? oy_ : (p == 1 ? -ox_ : (p == 2 ? 0.f : (p == 5 ? 1.f : 0.f)));
        acc += xp0 * (W[0] * xq0 + W[3] * xq1 + W[4] * xq2) + xp1 * (W[3] * xq0 + W[1] * xq1 + W[5] * xq2) + xp2 * (W[4] * xq0 + W[5] * xq1 + W[2] * xq2);
      }
      BS[idx] = acc;
    }
    SYNC();
  }
  NOUNROLL for (int idx = lane; idx < npair; idx += LANES) {          // lower triangle only, one (i, j <= i) pair per lane
    const int t = TB(tri)[idx], i = t >> 8, j = t & 255;
    float h = M[i * nv + j];
    if (few) {
      NOUNROLL for (int c = 0; c < ncon; ++c) {
        const float* rec = CRECS(c); const float* cw = rec + CR_V;
        const float gnn = cw[0];
        if (gnn == 0.f) continue;
        const float gn1 = cw[1], gn2 = cw[2], g11 = cw[3], mu = rec[CR_MU], g22 = mu * mu * gnn - g11;
        const float* J = WS(W_CN_J) + (size_t)3 * c * nv;
        const float jn = J[j], j1 = J[nv + j], j2 = J[2 * nv + j];
        h += J[i] * (gnn * jn + gn1 * j1 + gn2 * j2) + J[nv + i] * (gn1 * jn + g11 * j1) + J[2 * nv + i] * (gn2 * jn + g22 * j2);
      }
    } else {
      const int bi = TB(dof_body)[i];
      if (cbg && ((TB(body_dofmask)[bi] >> j) & 1)) {        // j moves body(i): contact inertias of the subtree of body(i)
        const int bend = bi + TB(body_subsize)[bi];
        const float* di = cdof + 6 * i; const float* dj = cdof + 6 * j;
        NOUNROLL for (int b = bi; b < bend; ++b) if ((cbg >> b) & 1u) {
          const float* S = BS + 21 * b;
          float acc = 0.f; int e = 0;
#pragma unroll
          for (int p = 0; p < 6; ++p) {
#pragma unroll
            for (int q = 0; q <= p; ++q, ++e) acc += S[e] * (p == q ? di[p] * dj[p] : di[p] * dj[q] + di[q] * dj[p]);
          }
          h += acc;
        }
      }
      NOUNROLL for (int c = ncg; c < ncon; ++c) {        // geom-geom contacts: J = J(body 2) - J(body 1)
        const float* rec = CREC(c); const float* W = rec + CR_WW;
        if (W[0] < 0.f) continue;
        const int b2 = ((const int*)rec)[CR_BODY], b1 = TB(geom_body)[-2 - ((const int*)rec)[CR_CELL]];
        float off[3], ji[3], jj[3], t3[3]; v3sub(off, rec + CR_POS, scom);
        jac_col(m, ws, b2, i, off, ji); jac_col(m, ws, b1, i, off, t3); v3sub(ji, ji, t3);
        jac_col(m, ws, b2, j, off, jj); jac_col(m, ws, b1, j, off, t3); v3sub(jj, jj, t3);
        h += ji[0] * (W[0] * jj[0] + W[3] * jj[1] + W[4] * jj[2]) + ji[1] * (W[3] * jj[0] + W[1] * jj[1] + W[5] * jj[2]) + ji[2] * (W[4] * jj[0] + W[5] * jj[1] + W[2] * jj[2]);
      }
    }
    NOUNROLL for (int e = 0; e < 3 * neq; ++e) h += WS(W_EQ_D)[e] * WS(W_EQ_J)[(size_t)e * nv + i] * WS(W_EQ_J)[(size_t)e * nv + j];
    if (i == j) {
      const float D = WS(W_FR_D)[i];
      if (D > 0.f) { const float x = WS(W_TMPW)[i], Rf = WS(W_FLOSS)[i] / D; if (x > -Rf && x < Rf) h += D; }
      const int jn_ = TB(dof_jnt)[i]; const float sg = WS(W_LM_SIGN)[jn_];
      if (sg != 0.f && TB(jnt_dofadr)[jn_] == i && (sg * qacc[i] - WS(W_LM_AREF)[jn_]) < 0.f) h += WS(W_LM_D)[jn_];
    }
    H[i * nv + j] = h;
  }
  SYNC();
  // rows that couple two branches of the tree (connect constraints, geom-geom contacts) fill the Hessian outside the tree pattern
  chol_factor(m, H, WS(W_INVD), nv, lane, neq == 0 && !coupled);
  chol_solve(H, WS(W_INVD), WS(W_TMPV), WS(W_BUF), WS(W_SEARCH), nv, lane);
  return gn;
}

// part 0: the whole solve; 1: up to the first search direction (cost and factor signature are handed back in nc);
// 2: the iterations, continuing from part 1 (forward() may put a CTA-wide barrier between the two)
// 3: the whole solve with the ITERATIONS of all env-warps of the CTA in lock step (a CTA-wide barrier in front of every
//    iteration; converged warps keep answering the barrier until the last one is done): every warp of the CTA must call,
//    the ones without an env with part 4 (barriers only).  The state of the solve stays in registers (no carry).
struct NewtonCarry { float cost; uint32_t sig; };
// a second barrier per lock-step iteration, after the line search (its ~4 cost evaluations vary per env): same-box A / B of the
// 768-thread build on flamingo_p_v3 / rocky_hard 2.765 -> 2.81 M env-steps/s at steady state, 3.306 -> 3.338 M at steps 3 - 23
#ifndef COSIM_NEWTON_LS_BARRIER
#define COSIM_NEWTON_LS_BARRIER 1
#endif
#if COSIM_NEWTON_LS_BARRIER
#define NEWTON_LS_SYNC() CTA_SYNC()
#else
#define NEWTON_LS_SYNC() ((void)0)
#endif
template <bool LOCK> DEV_NOINLINE int newton_solve(const ModelDev& m, WSP ws, int ncon, int has_rows, int lane, int part, NewtonCarry& nc) { LANE_REFRESH();
  const int nv = MD(nv), neq = MD(neq);
  // the field pointers are re-derived at their uses (one LDS + add) instead of being kept alive across the out-of-line calls, where
  // the 80-register build spilled five of them to local memory (L1 holds < 10 stack lines per warp): +1.1 % on flamingo_p_v3
#define qacc WS(W_QACC)
#define Ma WS(W_MA)
#define Mv WS(W_MV)
#define search WS(W_SEARCH)
#define M WS(W_M)
  constexpr bool lock = LOCK;      // two instances: the lock-step one costs the plain one nothing
  if (!has_rows) {
    if (part != 2 && part != 4) { FOR_LANE(k, nv) { qacc[k] = WS(W_ASMOOTH)[k]; WS(W_WARM)[k] = WS(W_ASMOOTH)[k]; WS(W_FCON)[k] = 0.f; } SYNC(); }
    if (!lock) return 0;
  }
  // warm start: cheaper of qacc_warmstart and qacc_smooth [upstream mj_fwdConstraint].  The warm start usually wins, so it
  // is evaluated last: its row residuals / forces are then already in place and only a smooth-start win pays a third pass.
  float cost = nc.cost; uint32_t sig = nc.sig;
  if (part != 2 && has_rows) {
    const float cs = total_cost(m, ws, ncon, WS(W_ASMOOTH), Ma, lane);
    cost = total_cost(m, ws, ncon, WS(W_WARM), Ma, lane);
    if (cost > cs) cost = total_cost(m, ws, ncon, WS(W_ASMOOTH), Ma, lane);
    sig = 0;
    (void)newton_direction(m, ws, ncon, lane, sig);
    if (part == 1) { nc.cost = cost; nc.sig = sig; return 0; }
  }
  const float scale = 1.f / (WS(W_SCAL)[1] * (float)imax(1, nv));
  const float tol = MO(tolerance);
  const int maxiter = MD(iterations);
  int iter = 0;
  bool go = has_rows != 0;
  if (lock) CTA_SYNC();
  for (;;) {
    const int cont = go && iter < maxiter;
    if (lock) { if (!CTA_SYNC_OR(cont)) break; }
    if (!cont) { if (lock) { NEWTON_LS_SYNC(); continue; } else break; }
    mat_vec(M, search, Mv, nv, lane);
    compute_jv(m, ws, ncon, search, lane);
    float q1 = 0.f, q2 = 0.f, sn = 0.f, gauss = 0.f, absterms = 0.f;
    FOR_LANE(k, nv) {
      const float r = Ma[k] - WS(W_FSMOOTH)[k];
      q1 += search[k] * r; q2 += 0.5f * search[k] * Mv[k]; sn += search[k] * search[k];
      gauss += r * (qacc[k] - WS(W_ASMOOTH)[k]); absterms += fabsf(search[k] * r);
      const float D = WS(W_FR_D)[k]; if (D > 0.f) absterms += fminf(fabsf(D * WS(W_TMPW)[k]), WS(W_FLOSS)[k]) * fabsf(search[k]);
    }
    FOR_LANE(i, 3 * neq) absterms += fabsf(WS(W_EQ_D)[i] * WS(W_EQ_X)[i] * WS(W_EQ_V)[i]);
    NOUNROLL for (int idx = lane; idx < 4 * ncon; idx += LANES) { const float* rec = CREC(idx >> 2); const int e = idx & 3; absterms += fabsf(rec[CR_D] * rec[CR_X + e] * rec[CR_V + e]); }
    q1 = wsum(q1); q2 = wsum(q2); sn = sqrtf(wsum(sn)); gauss = 0.5f * wsum(gauss); absterms = wsum(absterms);
    const float alpha = linesearch(m, ws, ncon, gauss, q1, q2, sn, absterms, lane);
    if (lock) NEWTON_LS_SYNC();
    if (alpha == 0.f) { if (lock) { go = false; continue; } else break; }
    FOR_LANE(k, nv) { qacc[k] += alpha * search[k]; Ma[k] += alpha * Mv[k]; }
    FOR_LANE(i, 3 * neq) WS(W_EQ_X)[i] += alpha * WS(W_EQ_V)[i];
    NOUNROLL for (int idx = lane; idx < 4 * ncon; idx += LANES) { float* rec = CREC(idx >> 2); const int e = idx & 3; rec[CR_X + e] += alpha * rec[CR_V + e]; }
    SYNC();
    FOR_LANE(k, nv) WS(W_TMPW)[k] = qacc[k] - WS(W_FR_AREF)[k];
    SYNC();
    const float old = cost;
    float g = 0.f;
    FOR_LANE(k, nv) g += (Ma[k] - WS(W_FSMOOTH)[k]) * (qacc[k] - WS(W_ASMOOTH)[k]);
    cost = update_forces(m, ws, ncon, lane) + 0.5f * wsum(g);
    float fn;
    const float gn = newton_direction(m, ws, ncon, lane, sig, &fn);
    ++iter;
    // reference criteria (scaled improvement / gradient below `tolerance` = 1e-8) with fp32 round-off floors: the cost is a
    // sum of O(|cost|) terms, so an improvement below a few ulps of it is noise (without the floor the fp32 engine spends
    // one more iteration per sub-step than the fp64 oracle just to see the improvement turn negative)
    // ... and the gradient is a difference of three force vectors of norm fn, so it cannot be resolved below a few ulps of fn
    if ((old - cost) < fmaxf(tol / scale, COST_EPS * fabsf(old)) || gn < fmaxf(tol / scale, GRAD_EPS * fn)) { if (lock) go = false; else break; }
  }
  if (!has_rows) return 0;
  FOR_LANE(k, nv) WS(W_WARM)[k] = qacc[k];
  SYNC();
  return iter;
}
#undef qacc
#undef Ma
#undef Mv
#undef search
#undef M

// ------------------------------------------------------------------------------------------ sensors
DEV_NOINLINE void sensors(const ModelDev& m, WSP ws, int lane) { LANE_REFRESH();
  if (lane == 0) {
    const int b = MD(imu_body);
    float sq[4] = {m.imu_quat[0], m.imu_quat[1], m.imu_quat[2], m.imu_quat[3]};
    quat_normalize(sq);
    float* S = WS(W_SENS);
    quat_mul(S, WS(W_XQUAT) + 4 * b, sq);
    float smat[9]; quat_to_mat(smat, S);
    float r[3], spos[3], off[3], c[3], lin[3];
    m3mulv(r, WS(W_XMAT) + 9 * b, m.imu_pos); v3add(spos, WS(W_XPOS) + 3 * b, r);
    const float* cv = WS(W_CVEL) + 6 * b;
    v3sub(off, spos, WS(W_SCOM)); v3cross(c, cv, off); v3add(lin, cv + 3, c);
    m3tmulv(S + 4, smat, cv); m3tmulv(S + 7, smat, lin);
    for (int k = 0; k < 3; ++k) { S[4 + k] = fminf(34.9f, fmaxf(-34.9f, S[4 + k])); S[7 + k] = fminf(30.f, fmaxf(-30.f, S[7 + k])); }
  }
}

#if COSIM_GENERAL
#include "engine_general.h"
#endif

// ------------------------------------------------------------------------------------------ forward + one sub-step
// The forward pass in four stages.  They are separate functions because the pooled step kernel (engine.cu k_step_pool) runs a
// stage for every environment of a CTA's pool before any environment enters the next one; forward() below strings them
// together for one environment (with optional CTA-wide barriers in between).
// stage 1: kinematics, spatial inertias, mass matrix and its Cholesky factor
DEV void stage_kin(const ModelDev& m, WSP ws, int lane, int part = 0) {
  const int nv = MD(nv);
  float* A = WS(W_A);
  if (part != 2) {
    kinematics(m, ws, lane);
    com_pos(m, ws, lane);
  }
  if (part == 1) return;
  crb(m, ws, lane);
  const float* M = WS(W_M);
  FOR_LANE(i, nv * nv) A[i] = M[i];
  SYNC();
  chol_factor(m, A, WS(W_INVD), nv, lane, 1);
}
// stage 2: collision
DEV void stage_collide(const ModelDev& m, WSP ws, int lane, int part = 0) {
  if (MD(ground_type) == 1) { if (m.hf_fine) collide_hfield_all<true>(m, ws, lane, part); else collide_hfield_all<false>(m, ws, lane, part); if (part == 1) return; }
  else if (part == 1) return;
  else {
    int ncon = 0, dropped = 0;
    NOUNROLL for (int g = 0; g < MD(ngeom); ++g) collide_plane(m, ws, g, ncon, dropped, lane);
    if (lane == 0) { WSI(W_CNT)[CNT_NCON] = ncon; WSI(W_CNT)[CNT_DROPPED] = dropped; }
  }
  SYNC();
  if (MD(npair) > 0) { if (m.n_boxbox) collide_pairs<true>(m, ws, lane); else collide_pairs<false>(m, ws, lane); }
}
// stage 3: sensors, smooth forces and acceleration, constraint rows
DEV void stage_smooth(const ModelDev& m, WSP ws, int lane) {
  const int nv = MD(nv), nu = MD(nu), njnt = MD(njnt);
  PH_DECL;
  float* A = WS(W_A);
  const int ncon = WSI(W_CNT)[CNT_NCON];
  com_vel(m, ws, lane);
  sensors(m, ws, lane);
  // smooth forces: passive damping - bias + actuation (ctrl clamp, gear, actuatorfrcrange clamp)
  rne_bias(m, ws, WS(W_TMPV), lane);
  float* fs = WS(W_FSMOOTH);
  FOR_LANE(k, nv) fs[k] = -LDG(TB(dof_damping) + k) * WS(W_QVEL)[k] - WS(W_TMPV)[k];
  SYNC();
  FOR_LANE(a, nu) {
    float c = WS(W_CTRL)[a];
    if (TB(act_ctrllimited)[a]) c = fminf(LDG(TB(act_ctrlrange) + 2 * a + 1), fmaxf(LDG(TB(act_ctrlrange) + 2 * a), c));
    float f = LDG(TB(act_gear) + a) * c;
    const int k = TB(act_dof)[a], j = TB(dof_jnt)[k];
    if (TB(jnt_actfrclimited)[j]) f = fminf(LDG(TB(jnt_actfrcrange) + 2 * j + 1), fmaxf(LDG(TB(jnt_actfrcrange) + 2 * j), f));
    fs[k] += f;     // one actuator per joint in all four robots
  }
  SYNC();
  FOR_LANE(k, nv) WS(W_TMPV)[k] = fs[k];
  SYNC();
  chol_solve(A, WS(W_INVD), WS(W_TMPV), WS(W_BUF), WS(W_ASMOOTH), nv, lane);
  PH_MARK(PH_SMOOTH);
  // constraint rows last: the few-contact Jacobians (W_CN_J) take the place of the body velocities / accelerations,
  // which sensors() and rne_bias() have consumed by now
  make_constraint(m, ws, ncon, lane);
  PH_MARK(PH_CONSTRAINT);
#if COSIM_GENERAL
  if (m.general) { gen_make_rows(m, ws, ncon, lane); return; }      // leaves the row count in W_CNT[CNT_ROWS]
#endif
  // any constraint row?
  int rows = (ncon > 0) || (MD(neq) > 0);
  { int r = 0; FOR_LANE(k, nv) r |= (WS(W_FR_D)[k] > 0.f); FOR_LANE(j, njnt) r |= (WS(W_LM_SIGN)[j] != 0.f); rows |= wor(r); }
  if (lane == 0) WSI(W_CNT)[CNT_ROWS] = rows;
  SYNC();
}
// stage 4: constraint solve; returns the solver iterations
DEV int stage_newton(const ModelDev& m, WSP ws, int lane, int part, NewtonCarry& nc) {
#if COSIM_GENERAL
  if (m.general) { if (part == 2 || part == 4) return 0; const int it_ = gen_solve(m, ws, WSI(W_CNT)[CNT_NCON], lane); PH_COUNT(PH_NEWTON_ITERS, it_); return it_; }
#endif
  const int iters = part >= 3 ? newton_solve<true>(m, ws, part == 4 ? 0 : WSI(W_CNT)[CNT_NCON], part == 4 ? 0 : WSI(W_CNT)[CNT_ROWS], lane, part, nc)
                              : newton_solve<false>(m, ws, WSI(W_CNT)[CNT_NCON], WSI(W_CNT)[CNT_ROWS], lane, part, nc);
  if (part != 1) PH_COUNT(PH_NEWTON_ITERS, iters);
  return iters;
}
DEV int stage_newton(const ModelDev& m, WSP ws, int lane) { NewtonCarry nc = {0.f, 0u}; return stage_newton(m, ws, lane, 0, nc); }
// returns solver iterations; the contact count of this pass is left in W_CNT
// `active` = this warp has an env to advance; `bsync` = CTA-wide phase barriers (must then be called by every warp)
DEV_NOINLINE int forward(const ModelDev& m, WSP ws, int lane, int active = 1, int bsync = 0) { LANE_REFRESH();
  PH_DECL;
  int iters = 0;
  // extra barriers INSIDE the three big phases (bits 6, 7, 8 of the mask): the narrower the window of code the warps of an SM
  // execute together, the fewer instruction-cache misses each of them takes (profiles/r02_experiments.md)
  const int sp = bsync ? (m.bsync_mask >> 6) : 0;
  if (active) stage_kin(m, ws, lane, (sp & 1) ? 1 : 0);
  if (sp & 1) { CTA_SYNC(); if (active) stage_kin(m, ws, lane, 2); }
  PH_MARK(PH_KIN);
  BSYNC_IF(bsync, 0);
  PH_MARK(PH_WAIT_KIN);
  if (active) stage_collide(m, ws, lane, (sp & 2) ? 1 : 0);
  if (sp & 2) { CTA_SYNC(); if (active) stage_collide(m, ws, lane, 2); }
  PH_MARK(PH_COLLIDE);
  BSYNC_IF(bsync, 1);
  PH_MARK(PH_WAIT_COLLIDE);
  if (active) stage_smooth(m, ws, lane);
  PH_MARK(PH_SMOOTH);
  BSYNC_IF(bsync, 2);
  PH_MARK(PH_WAIT_SMOOTH);
  { NewtonCarry nc = {0.f, 0u};
    if (sp & 8) iters = stage_newton(m, ws, lane, active ? 3 : 4, nc);      // iterations of all warps in lock step (fast path only)
    else {
      if (active) iters = stage_newton(m, ws, lane, (sp & 4) ? 1 : 0, nc);
      if (sp & 4) { CTA_SYNC(); if (active) iters = stage_newton(m, ws, lane, 2, nc); } } }
  PH_MARK(PH_NEWTON);
  BSYNC_IF(bsync, 3);
  PH_MARK(PH_WAIT_NEWTON);
  return iters;
}

DEV_NOINLINE int bad_state(const ModelDev& m, WSP ws, int lane) { LANE_REFRESH();
  int bad = 0;
  FOR_LANE(i, MD(nq)) { float x = WS(W_QPOS)[i]; bad |= !(x == x) || fabsf(x) > 1e10f; }
  FOR_LANE(i, MD(nv)) { float x = WS(W_QVEL)[i]; bad |= !(x == x) || fabsf(x) > 1e10f; }
  return wor(bad);
}
DEV_NOINLINE void reset_data(const ModelDev& m, WSP ws, int lane) { LANE_REFRESH();
  FOR_LANE(i, MD(nq)) WS(W_QPOS)[i] = LDG(TB(qpos0) + i);
  FOR_LANE(i, MD(nv)) { WS(W_QVEL)[i] = 0.f; WS(W_WARM)[i] = 0.f; }
  SYNC();
}

// what mj_step does around the forward pass: the state check in front of it ...
DEV void substep_pre(const ModelDev& m, WSP ws, int lane) {
  if (bad_state(m, ws, lane)) { reset_data(m, ws, lane); if (lane == 0) WSI(W_CNT)[CNT_NAN]++; }
}
// ... and the acceleration check + implicitfast integration after it (returns the solver iterations that count)
DEV_NOINLINE int substep_post(const ModelDev& m, WSP ws, int lane, int iters) { LANE_REFRESH();
  const int nv = MD(nv), njnt = MD(njnt);
  { int bad = 0; FOR_LANE(i, nv) { float x = WS(W_QACC)[i]; bad |= !(x == x) || fabsf(x) > 1e10f; }
    if (wor(bad)) { reset_data(m, ws, lane); if (lane == 0) WSI(W_CNT)[CNT_NAN]++; iters = forward(m, ws, lane, 1, 0); } }   // rare: no barriers inside
  // implicitfast: (M + dt diag(damping)) a = qfrc_smooth + qfrc_constraint
  PH_DECL;
  const float dt = MO(timestep);
  float* A = WS(W_A); const float* M = WS(W_M);
  FOR_LANE(i, nv * nv) { const int r = i / nv, c = i - r * nv; A[i] = M[i] + (r == c ? dt * LDG(TB(dof_damping) + r) : 0.f); }
  FOR_LANE(k, nv) WS(W_TMPV)[k] = WS(W_FSMOOTH)[k] + WS(W_FCON)[k];
  SYNC();
  chol_factor(m, A, WS(W_INVD), nv, lane, 1);
  chol_solve(A, WS(W_INVD), WS(W_TMPV), WS(W_BUF), WS(W_GRAD), nv, lane);
  float* qvel = WS(W_QVEL); float* qpos = WS(W_QPOS);
  FOR_LANE(k, nv) qvel[k] += dt * WS(W_GRAD)[k];
  SYNC();
  FOR_LANE(j, njnt) {
    const int qa = TB(jnt_qposadr)[j], da = TB(jnt_dofadr)[j];
    if (TB(jnt_type)[j] == 0) {
      for (int k = 0; k < 3; ++k) qpos[qa + k] += dt * qvel[da + k];
      float w[3] = {qvel[da + 3], qvel[da + 4], qvel[da + 5]};
      float ang = v3norm(w) * dt;
      if (ang > 0.f) {
        v3normalize(w);
        float s, c; sincosf(ang * 0.5f, &s, &c);
        float dq[4] = {c, w[0] * s, w[1] * s, w[2] * s}, q[4];
        quat_mul(q, qpos + qa + 3, dq); quat_normalize(q);
        qpos[qa + 3] = q[0]; qpos[qa + 4] = q[1]; qpos[qa + 5] = q[2]; qpos[qa + 6] = q[3];
      }
    } else qpos[qa] += dt * qvel[da];
  }
  if (lane == 0) WSI(W_CNT)[CNT_DROPPED_STEP] += WSI(W_CNT)[CNT_DROPPED];
  SYNC();
  PH_MARK(PH_INTEGRATE);
  return iters;
}
DEV_NOINLINE int substep(const ModelDev& m, WSP ws, int lane, int active = 1, int bsync = 0) { LANE_REFRESH();
  if (active) substep_pre(m, ws, lane);
  int iters = forward(m, ws, lane, active, bsync);
  if (active) iters = substep_post(m, ws, lane, iters);
  BSYNC_IF(bsync, 4);
  return iters;
}

// cfrc_ext of the bodies in contact / connected -> W_CACC region reused as [nbody][6]
DEV_NOINLINE void cfrc_ext(const ModelDev& m, WSP ws, int ncon, int lane) { LANE_REFRESH();
  const int nb = MD(nbody), neq = MD(neq);
  float* out = WS(W_CACC); const float* scom = WS(W_SCOM);
  FOR_LANE(i, 6 * nb) out[i] = 0.f;
  SYNC();
  if (lane == 0) {   // few contacts; sequential keeps the summation order fixed
    NOUNROLL for (int c = 0; c < ncon; ++c) {
      const float* rec = CREC(c); const float* F = rec + CR_F; float wf[3], arm[3], tq[3];
      m3tmulv(wf, rec + CR_FRAME, F);
      v3sub(arm, rec + CR_POS, scom); v3cross(tq, arm, wf);
      if (IF_GENERAL(m)) { float wt[3]; m3tmulv(wt, rec + CR_FRAME, rec + CR_WF); v3add(tq, tq, wt); }      // torsional / rolling moments (contact frame -> world)
      float* o = out + 6 * ((const int*)rec)[CR_BODY];
      for (int k = 0; k < 3; ++k) { o[k] += tq[k]; o[3 + k] += wf[k]; }
      const int b1 = contact_body1(m, ws, c);
      if (b1 > 0) { float* o1 = out + 6 * b1; for (int k = 0; k < 3; ++k) { o1[k] -= tq[k]; o1[3 + k] -= wf[k]; } }
    }
    NOUNROLL for (int e = 0; e < neq; ++e) {
      const int b1 = TB(eq_body1)[e], b2 = TB(eq_body2)[e];
      float a1[3] = {TB(eq_anchor1)[3 * e], TB(eq_anchor1)[3 * e + 1], TB(eq_anchor1)[3 * e + 2]}, p1[3], r[3], arm[3], tq[3];
      m3mulv(r, WS(W_XMAT) + 9 * b1, a1); v3add(p1, WS(W_XPOS) + 3 * b1, r);
      const float* wf = WS(W_EQ_F) + 3 * e;
      v3sub(arm, p1, scom); v3cross(tq, arm, wf);
      for (int k = 0; k < 3; ++k) { out[6 * b1 + k] += tq[k]; out[6 * b1 + 3 + k] += wf[k]; out[6 * b2 + k] -= tq[k]; out[6 * b2 + 3 + k] -= wf[k]; }
    }
  }
  SYNC();
}

// ------------------------------------------------------------------------------------------ height map (K9)
DEV_NOINLINE float hfield_height(const ModelDev& m, float x, float y, int* cell) {
  const int nrow = MD(hf_nrow), ncol = MD(hf_ncol);
  const float sx = MO(hf_sx), sy = MO(hf_sy), sz = MO(hf_sz);
  *cell = -1;
  if (x < -sx || x > sx || y < -sy || y > sy) return NAN;
  const float dx = 2.f * sx / (float)(ncol - 1), dy = 2.f * sy / (float)(nrow - 1);
  int c = (int)floorf((x + sx) / dx), r = (int)floorf((y + sy) / dy);
  c = imin(imax(c, 0), ncol - 2); r = imin(imax(r, 0), nrow - 2);
  const float u = (x - (dx * (float)c - sx)) / dx, v = (y - (dy * (float)r - sy)) / dy;
  const float* h = m.hfield_data + (size_t)r * ncol + c;
  const float z00 = LDGB(h) * sz, z10 = LDGB(h + 1) * sz, z01 = LDGB(h + ncol) * sz, z11 = LDGB(h + ncol + 1) * sz;
  float z; int tri;
  if (u >= v) { tri = 0; z = z00 + u * (z10 - z00) + v * (z11 - z10); }
  else { tri = 1; z = z00 + v * (z01 - z00) + u * (z11 - z01); }
  *cell = ((r * ncol + c) << 1) | tri;
  return z;
}
