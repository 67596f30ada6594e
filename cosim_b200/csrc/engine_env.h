// engine_env.h -- env layer of the warp-per-env engine: domain randomisation + mj_setConst,
// reset, PD/delay, observation build, statistics.  Each *_env function is the body one warp runs
// for one environment; engine.cu wraps them in __global__ kernels.
//
// Reference code restated here (all under /root/reference/envs/):
//   init_env     <- manager/xml_manager.py:43-87 (per-env draws) + MuJoCo mj_setConst [upstream]
//   reset_env    <- flamingo_p_v3/flamingo_p_v3.py:235-255, wrappers.py:245-256,385-389
//   step_env     <- flamingo_p_v3.py:150-199 (delay, PD, clip, do_simulation, obs, info, done),
//                   manager/control_manager.py:13-26, wrappers.py:258-269,309-320,391-405
//   observe      <- flamingo_p_v3.py:115-148, utils/mujoco_utils.py:99-189, wrappers.py:160-243
#pragma once
#include "engine_core.h"

struct StepArgs {
  const float* action;         // [N][nu]
  const float* command;        // [N][command_dim] applied (scaled) command, may be NULL (zeros)
  const float* user_command;   // [N][command_dim] raw user command for statistics, may be NULL (= command)
  float* state_out;            // [N][state_dim]
  uint8_t* terminated;         // [N]
  uint8_t* truncated;          // [N]
  const uint8_t* mask;         // reset: [N] or NULL (all)
};

// Per-env rows in HBM are touched once per control step: streaming loads / stores (evict-first) keep them from pushing the
// warps' local-memory lines (callee-saved registers of the out-of-line stage functions) out of L2 -- those evictions, not the
// env rows, were 80 % of the DRAM writes of a step (profiles/r02_dram_traffic.md).
#ifdef COSIM_HOST_EMU
#define ELD(p) (*(p))
#define EST(p, v) (*(p) = (v))
#else
#define ELD(p) __ldcs(p)
#define EST(p, v) __stcs((p), (v))
#endif

DEV_NOINLINE void load_params(const ModelDev& m, const EnvArrays& E, int env, WSP ws, int lane) { LANE_REFRESH();
  const int nb = MD(nbody), nv = MD(nv), ng = MD(ngeom), nu = MD(nu);
  FOR_LANE(i, nb) { WS(W_BMASS)[i] = ELD(E.body_mass + (size_t)env * nb + i); WS(W_INVWB)[i] = ELD(E.invw_body + (size_t)env * nb + i); }
  FOR_LANE(i, nv) { WS(W_INVWD)[i] = ELD(E.invw_dof + (size_t)env * nv + i); WS(W_FLOSS)[i] = ELD(E.floss + (size_t)env * nv + i); }
  FOR_LANE(i, ng) WS(W_GMU)[i] = ELD(E.gmu + (size_t)env * ng + i);
  FOR_LANE(i, nu) { WS(W_KP)[i] = ELD(E.kp + (size_t)env * nu + i); WS(W_KD)[i] = ELD(E.kd + (size_t)env * nu + i); }
  FOR_LANE(i, 4) WS(W_SCAL)[i] = ELD(E.scal + (size_t)env * 4 + i);
  // torsional / rolling friction of the env (xml_manager.py:57-75 writes them next to the sliding coefficient): the same draws
  // init_env() makes, recomputed from the counter-based RNG -- only the general constraint path reads them
  if (IF_GENERAL(m)) FOR_LANE(i, 2) WS(W_SCAL)[4 + i] = fmaf(uni(m, env, RNG_MODEL, 0, 1 + i), m.rnd_span[1 + i], m.rnd_lo[1 + i]);
}
DEV_NOINLINE void load_state(const ModelDev& m, const EnvArrays& E, int env, WSP ws, int lane) { LANE_REFRESH();
  const int nq = MD(nq), nv = MD(nv);
  FOR_LANE(i, nq) WS(W_QPOS)[i] = ELD(E.qpos + (size_t)env * nq + i);
  FOR_LANE(i, nv) { WS(W_QVEL)[i] = ELD(E.qvel + (size_t)env * nv + i); WS(W_WARM)[i] = ELD(E.warm + (size_t)env * nv + i); }
  if (MD(npair) > 0) FOR_LANE(i, 4 * PAXIS_SLOTS) WS(W_PAXIS)[i] = ELD(E.paxis + (size_t)env * 4 * PAXIS_SLOTS + i);
}
DEV_NOINLINE void store_state(const ModelDev& m, const EnvArrays& E, int env, WSP ws, int lane) { LANE_REFRESH();
  const int nq = MD(nq), nv = MD(nv);
  FOR_LANE(i, nq) EST(E.qpos + (size_t)env * nq + i, WS(W_QPOS)[i]);
  FOR_LANE(i, nv) { EST(E.qvel + (size_t)env * nv + i, WS(W_QVEL)[i]); EST(E.warm + (size_t)env * nv + i, WS(W_WARM)[i]); }
  if (MD(npair) > 0) FOR_LANE(i, 4 * PAXIS_SLOTS) EST(E.paxis + (size_t)env * 4 * PAXIS_SLOTS + i, WS(W_PAXIS)[i]);
}

// ------------------------------------------------------------------------------------------ init: randomise + setConst
DEV_NOINLINE void init_env(const ModelDev& m, const EnvArrays& E, int env, WSP ws, int lane) { LANE_REFRESH();
  const int nb = MD(nbody), nv = MD(nv), ng = MD(ngeom), nu = MD(nu), njnt = MD(njnt);
  float draw[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) draw[i] = fmaf(uni(m, env, RNG_MODEL, 0, i), m.rnd_span[i], m.rnd_lo[i]);
  const float slide = draw[0], floss = draw[3], delay = draw[4], load = draw[5], kps = draw[6], kds = draw[7];
  FOR_LANE(b, nb) WS(W_BMASS)[b] = LDG(TB(body_mass) + b);
  SYNC();
  FOR_LANE(i, MD(n_massnoise)) {
    const int b = TB(massnoise_body)[i];
    const float m0 = LDG(TB(body_mass) + b), mk = m0 * MO(mass_noise);
    float mass = m0 + fmaf(uni(m, env, RNG_MODEL, 0, 8 + i), 2.f * mk, -mk);
    if (b == MD(base_body)) mass += load;
    WS(W_BMASS)[b] = mass;
  }
  FOR_LANE(g, ng) WS(W_GMU)[g] = TB(geom_fr_random)[g] ? slide : LDG(TB(geom_friction) + 3 * g);
  FOR_LANE(k, nv) WS(W_FLOSS)[k] = TB(dof_fl_random)[k] ? floss : LDG(TB(dof_frictionloss) + k);
  FOR_LANE(a, nu) { E.kp[(size_t)env * nu + a] = LDG(TB(act_kp) + a) * kps; E.kd[(size_t)env * nu + a] = LDG(TB(act_kd) + a) * kds; }
  FOR_LANE(i, MD(nq)) WS(W_QPOS)[i] = LDG(TB(qpos0) + i);
  SYNC();
  // mj_setConst at qpos0: M, M^-1, dof / body inverse weights, meaninertia
  kinematics(m, ws, lane); com_pos(m, ws, lane); crb(m, ws, lane);
  float* A = WS(W_A); float* M = WS(W_M);
  float tr = 0.f;
  FOR_LANE(i, nv) tr += M[i * nv + i];
  const float meaninertia = wsum(tr) / (float)nv;
  FOR_LANE(i, nv * nv) A[i] = M[i];
  SYNC();
  chol_factor(m, A, WS(W_INVD), nv, lane, 1);
  NOUNROLL for (int c = 0; c < nv; ++c) {           // column c of M^-1 overwrites M
    FOR_LANE(k, nv) WS(W_TMPV)[k] = (k == c) ? 1.f : 0.f;
    SYNC();
    chol_solve(A, WS(W_INVD), WS(W_TMPV), WS(W_BUF), WS(W_GRAD), nv, lane);
    FOR_LANE(k, nv) M[k * nv + c] = WS(W_GRAD)[k];
    SYNC();
  }
  FOR_LANE(j, njnt) {
    const int a = TB(jnt_dofadr)[j];
    if (TB(jnt_type)[j] == 0) {
      const float t = (M[a * nv + a] + M[(a + 1) * nv + a + 1] + M[(a + 2) * nv + a + 2]) / 3.f;
      const float r = (M[(a + 3) * nv + a + 3] + M[(a + 4) * nv + a + 4] + M[(a + 5) * nv + a + 5]) / 3.f;
      NOUNROLL for (int k = 0; k < 3; ++k) { WS(W_INVWD)[a + k] = t; WS(W_INVWD)[a + 3 + k] = r; }
    } else WS(W_INVWD)[a] = M[a * nv + a];
  }
  if (lane == 0) WS(W_INVWB)[0] = 0.f;
  NOUNROLL for (int b = 1; b < nb; ++b) {
    float off[3]; v3sub(off, WS(W_XIPOS) + 3 * b, WS(W_SCOM));
    float tran = 0.f;
    NOUNROLL for (int r = 0; r < 3; ++r) {
      FOR_LANE(k, nv) { float jp[3]; jac_col(m, ws, b, k, off, jp); WS(W_TMPV)[k] = jp[r]; }
      SYNC();
      float acc = 0.f;
      FOR_LANE(i, nv) { float s = 0.f; NOUNROLL for (int k = 0; k < nv; ++k) s += M[i * nv + k] * WS(W_TMPV)[k]; acc += WS(W_TMPV)[i] * s; }
      tran += wsum(acc);
      SYNC();
    }
    if (lane == 0) WS(W_INVWB)[b] = tran / 3.f;
  }
  SYNC();
  FOR_LANE(i, nb) { E.body_mass[(size_t)env * nb + i] = WS(W_BMASS)[i]; E.invw_body[(size_t)env * nb + i] = WS(W_INVWB)[i]; }
  FOR_LANE(i, nv) { E.invw_dof[(size_t)env * nv + i] = WS(W_INVWD)[i]; E.floss[(size_t)env * nv + i] = WS(W_FLOSS)[i]; }
  FOR_LANE(i, ng) E.gmu[(size_t)env * ng + i] = WS(W_GMU)[i];
  if (lane == 0) {
    float* sc = E.scal + (size_t)env * 4;
    sc[0] = m.ground_friction[3] != 0.f ? slide : m.ground_friction[0]; sc[1] = meaninertia; sc[2] = delay; sc[3] = 0.f;
    int* ct = E.counters + (size_t)env * 8;
    NOUNROLL for (int i = 0; i < 8; ++i) ct[i] = 0;
    ct[CT_NEED_RESET] = 1;
  }
  FOR_LANE(i, ST__COUNT) E.stats[(size_t)env * ST__COUNT + i] = 0.f;
}

// ------------------------------------------------------------------------------------------ observation build
DEV_NOINLINE float trunc_noise(const ModelDev& m, int env, uint32_t nobs, int which, uint32_t idx) {
  if (MD(zero_noise)) return 0.f;
  const float* nz = TB(noise) + 6 * which;
  const float u = uni(m, env, RNG_NOISE, nobs, idx);
  const float p = LDG(nz + 4) + u * (LDG(nz + 5) - LDG(nz + 4));
  float x = LDG(nz) + LDG(nz + 1) * ndtri(p);
  return fminf(LDG(nz + 3), fmaxf(LDG(nz + 2), x));
}
DEV int raw_offset(const ModelDev& m, int kind) {
  const int np = MD(n_dofpos), nvl = MD(n_dofvel), nu = MD(nu);
  switch (kind) { case 0: return 0; case 1: return np; case 2: return np + nvl; case 3: return np + nvl + 3; case 4: return np + nvl + 6; case 5: return np + nvl + 9; case 6: return np + nvl + 9 + nu; }
  return 0;
}
// raw noisy observations -> W_RAW: [dof_pos | dof_vel | ang_vel | lin_vel | projected_gravity | last_action | height_map]
DEV_NOINLINE void get_obs(const ModelDev& m, const EnvArrays& E, int env, WSP ws, uint32_t nobs, int lane) { LANE_REFRESH();
  const int np = MD(n_dofpos), nvl = MD(n_dofvel), nu = MD(nu), rx = MD(hm_res_x), ry = MD(hm_res_y), nh = rx * ry;
  float* raw = WS(W_RAW); const float* qpos = WS(W_QPOS); const float* qvel = WS(W_QVEL); const float* S = WS(W_SENS);
  FOR_LANE(i, np) raw[i] = qpos[TB(dofpos_qadr)[i]] * LDG(TB(dofpos_fac) + i) + trunc_noise(m, env, nobs, 0, i);
  FOR_LANE(i, nvl) raw[np + i] = qvel[TB(dofvel_dadr)[i]] * LDG(TB(dofvel_fac) + i) + trunc_noise(m, env, nobs, 1, np + i);
  FOR_LANE(i, 3) {
    const int o = np + nvl;
    raw[o + i] = S[4 + i] + trunc_noise(m, env, nobs, 2, o + i);
    raw[o + 3 + i] = S[7 + i] + trunc_noise(m, env, nobs, 3, o + 3 + i);
    float q[4] = {S[0], S[1], S[2], S[3]};
    if (q[0] == 0.f && q[1] == 0.f && q[2] == 0.f && q[3] == 0.f) q[0] = 1.f;
    quat_normalize(q);
    float R[9]; quat_to_mat(R, q);
    raw[o + 6 + i] = -R[6 + i] + trunc_noise(m, env, nobs, 4, o + 6 + i);    // R^T (0,0,-1)
  }
  FOR_LANE(i, nu) raw[np + nvl + 9 + i] = WS(W_ACT)[i];
  if (nh) {
    const float sxm = MO(hm_size_x), sym = MO(hm_size_y), zmin = MO(hm_zmin);
    float R[9]; quat_to_mat(R, qpos + 3);
    const int o = np + nvl + 9 + nu;
    FOR_LANE(t, nh) {
      const int i = t / rx, j = t - i * rx;
      const float xr = rx > 1 ? -sxm * 0.5f + sxm * (float)j / (float)(rx - 1) : -sxm * 0.5f;
      const float yr = ry > 1 ? -sym * 0.5f + sym * (float)i / (float)(ry - 1) : -sym * 0.5f;
      const float pw[3] = {qpos[0] + R[0] * xr + R[1] * yr, qpos[1] + R[3] * xr + R[4] * yr, qpos[2] + R[6] * xr + R[7] * yr};
      int cell; const float z = hfield_height(m, pw[0], pw[1], &cell);
      float h;
      if (z == z) { const float origin = pw[2] + 10.f, dist = origin - z; h = dist >= 0.f ? qpos[2] - (origin - dist) : qpos[2] - zmin; if (dist < 0.f) cell = -1; }
      else h = qpos[2] - zmin;
      if (E.dbg_heightmap) { E.dbg_heightmap[(size_t)env * nh + t] = h; E.dbg_hmcell[(size_t)env * nh + t] = cell; }
      raw[o + t] = h + trunc_noise(m, env, nobs, 5, o - nu + t);
    }
  }
  SYNC();
}
// noise draw index convention: position in [dof_pos | dof_vel | ang_vel | lin_vel | proj_grav | height_map]

DEV_NOINLINE void concat_obs(const ModelDev& m, const EnvArrays& E, int env, WSP ws, int sim_step, bool stacked, const float* cmd, float* out, int lane) { LANE_REFRESH();
  const int* kind = stacked ? TB(sobs_kind) : TB(nobs_kind); const int* dim = stacked ? TB(sobs_dim) : TB(nobs_dim);
  const float* scale = stacked ? TB(sobs_scale) : TB(nobs_scale); const int* itv = stacked ? TB(sobs_interval) : TB(nobs_interval);
  const int* off = stacked ? TB(sobs_off) : TB(nobs_off); const int n = stacked ? MD(n_sobs) : MD(n_nobs);
  float* cache = E.freq_cache + (size_t)env * imax(1, MD(cache_dim));
  int o = 0;
  NOUNROLL for (int k = 0; k < n; ++k) {
    const int d = dim[k];
    if (kind[k] == OBS_COMMAND) { FOR_LANE(i, d) EST(out + o + i, cmd ? cmd[i] : 0.f); o += d; continue; }
    const bool upd = (sim_step == 0) || (sim_step % itv[k] == 0);
    const int ro = raw_offset(m, kind[k]);
    FOR_LANE(i, d) {
      float v;
      if (upd) { v = WS(W_RAW)[ro + i] * LDG(scale + k); EST(cache + off[k] + i, v); } else v = ELD(cache + off[k] + i);
      EST(out + o + i, v);
    }
    o += d;
  }
}
// StateBuildWrapper._build_state + CommandWrapper._apply_command_inplace
DEV_NOINLINE void build_state(const ModelDev& m, const EnvArrays& E, int env, WSP ws, int sim_step, bool reset, const float* cmd, float* state, int lane) { LANE_REFRESH();
  const int ss = MD(stack_size), sd = MD(stacked_dim);
  float* buf = E.obs_buffer + (size_t)env * ss * sd;
  // newest frame straight into state[0:sd]; older frames shift by one (or copy on reset)
  concat_obs(m, E, env, ws, sim_step, true, cmd, state, lane);
  SYNC();
  FOR_LANE(i, sd) {
    const float v = state[i];
    if (reset) { for (int k = 0; k < ss; ++k) { EST(buf + k * sd + i, v); EST(state + k * sd + i, v); } }
    else {
      NOUNROLL for (int k = ss - 1; k > 0; --k) { const float o = ELD(buf + (k - 1) * sd + i); EST(buf + k * sd + i, o); EST(state + k * sd + i, o); }
      EST(buf + i, v);
    }
  }
  concat_obs(m, E, env, ws, sim_step, false, cmd, state + ss * sd, lane);
  // command slots inside stacked frames carry the current command in every frame
  int o = 0;
  NOUNROLL for (int k = 0; k < MD(n_sobs); ++k) {
    if (TB(sobs_kind)[k] == OBS_COMMAND) for (int f = 1; f < ss; ++f) FOR_LANE(i, TB(sobs_dim)[k]) EST(state + f * sd + o + i, cmd ? cmd[i] : 0.f);
    o += TB(sobs_dim)[k];
  }
}

// ------------------------------------------------------------------------------------------ reset
DEV_NOINLINE void reset_env(const ModelDev& m, const EnvArrays& E, int env, WSP ws, const float* cmd, float* state, int lane) { LANE_REFRESH();
  const int nq = MD(nq), nv = MD(nv), nu = MD(nu);
  int* ct = E.counters + (size_t)env * 8;
  const uint32_t nreset = (uint32_t)ct[CT_NRESET], nobs = (uint32_t)ct[CT_NOBS];
  load_params(m, E, env, ws, lane);
  FOR_LANE(i, nq) WS(W_QPOS)[i] = (i == 2) ? MO(z0) : (i == 3 ? 1.f : 0.f);
  SYNC();
  const float lo = -MO(init_noise), hi = MO(init_noise);
  FOR_LANE(i, MD(n_initnoise)) WS(W_QPOS)[TB(initnoise_qadr)[i]] += fmaf(uni(m, env, RNG_RESET, nreset, i), hi - lo, lo);
  // optional spawn offset (engine.spawn_spread, not in the reference: oracle.hpp reset has the same lines): uniform in
  // [-spread, spread]^2, lifted by the highest terrain vertex within spawn_radius of the spot
  const float spread = MO(spawn_spread);
  if (spread > 0.f) {
    const float ox = (2.f * uni(m, env, RNG_RESET, nreset, 1000u) - 1.f) * spread, oy = (2.f * uni(m, env, RNG_RESET, nreset, 1001u) - 1.f) * spread;
    float hmax = 0.f;
    if (MD(ground_type) == 1) {
      const int ncol = MD(hf_ncol), nrow = MD(hf_nrow);
      const float sx = MO(hf_sx), sy = MO(hf_sy), rad = MO(spawn_radius);
      const float dx = 2.f * sx / (float)(ncol - 1), dy = 2.f * sy / (float)(nrow - 1);
      const int c0 = imax(0, (int)floorf((ox - rad + sx) / dx)), c1 = imin(ncol - 1, (int)ceilf((ox + rad + sx) / dx));
      const int r0 = imax(0, (int)floorf((oy - rad + sy) / dy)), r1 = imin(nrow - 1, (int)ceilf((oy + rad + sy) / dy));
      const int w = c1 - c0 + 1, cnt = w * (r1 - r0 + 1);
      float h = -INFINITY;
      FOR_LANE(t, cnt) h = fmaxf(h, LDGB(m.hfield_data + (size_t)(r0 + t / w) * ncol + c0 + t % w));
      hmax = wmaxf(h) * MO(hf_sz);
    }
    SYNC();
    if (lane == 0) { WS(W_QPOS)[0] += ox; WS(W_QPOS)[1] += oy; WS(W_QPOS)[2] += hmax; }
  }
  FOR_LANE(i, nv) { WS(W_QVEL)[i] = 0.f; WS(W_WARM)[i] = 0.f; }
  FOR_LANE(i, nu) { WS(W_CTRL)[i] = 0.f; WS(W_ACT)[i] = 0.f; E.prev_action[(size_t)env * nu + i] = 0.f; E.delay_prev[(size_t)env * nu + i] = 0.f; E.torque[(size_t)env * nu + i] = 0.f; E.last_action[(size_t)env * nu + i] = 0.f; }
  SYNC();
  int iters = forward(m, ws, lane);
  const int ncon = WSI(W_CNT)[CNT_NCON];
  get_obs(m, E, env, ws, nobs, lane);
  build_state(m, E, env, ws, 0, true, cmd, state, lane);
  store_state(m, E, env, ws, lane);
  if (lane == 0) {
    ct[CT_SIM_STEP] = 0; ct[CT_HAS_DELAY] = 0; ct[CT_NRESET] = (int)(nreset + 1); ct[CT_NOBS] = (int)(nobs + 1); ct[CT_NEED_RESET] = 0; ct[CT_NCON] = ncon;
    float* info = E.info + (size_t)env * 4;
    info[0] = 0.f; info[1] = WS(W_SENS)[7]; info[2] = WS(W_SENS)[8]; info[3] = WS(W_SENS)[6];
    if (E.dbg_iters) E.dbg_iters[env] = iters;
  }
  if (E.dbg_sens) FOR_LANE(i, 10) E.dbg_sens[(size_t)env * 10 + i] = WS(W_SENS)[i];
  if (E.dbg_qacc) FOR_LANE(i, nv) E.dbg_qacc[(size_t)env * nv + i] = WS(W_QACC)[i];
  SYNC();
}

DEV_NOINLINE void dump_contacts(const ModelDev& m, const EnvArrays& E, int env, WSP ws, int ncon, int lane) { LANE_REFRESH();
  if (!E.dbg_contacts) return;
  const int cap = MD(ncon_max);
  float* out = E.dbg_contacts + (size_t)env * cap * 10;
  FOR_LANE(c, ncon) {
    float* o = out + 10 * c;
    const float* rec = CREC(c);
    o[0] = rec[CR_DIST];
    NOUNROLL for (int k = 0; k < 3; ++k) { o[1 + k] = rec[CR_POS + k]; o[4 + k] = rec[CR_FRAME + k]; }
    o[7] = (float)((const int*)rec)[CR_GEOM]; o[8] = (float)((const int*)rec)[CR_CELL]; o[9] = rec[CR_MU];
  }
}

// ------------------------------------------------------------------------------------------ step
// What a control step does around its sub-steps, in two parts so that the pooled kernel (engine.cu k_step_pool) can run them as
// stages of their own.  StepLocals = what the first part hands to the second.
struct StepLocals { int active, sim_step; uint32_t nstep, nobs; float rm, tabs, tsq, tmax; };
// part 1: pending reset (the reference asserts reset-before-step; the batched engine resets the env instead and skips the
// step), parameters and state -> workspace, delay filter, PD law, clip.  L.active = 0: nothing more to do for this env.
DEV void step_prologue(const ModelDev& m, const EnvArrays& E, int env, WSP ws, const StepArgs& a, int lane, StepLocals& L) {
  const int nu = MD(nu), cd = MD(command_dim), sdim = MD(state_dim);
  int* ct = E.counters + (size_t)env * 8;
  const float* cmd = a.command ? a.command + (size_t)env * cd : nullptr;
  float* state = a.state_out + (size_t)env * sdim;
  L.active = 1; L.sim_step = 0; L.nstep = 0; L.nobs = 0; L.rm = L.tabs = L.tsq = L.tmax = 0.f;
  if (ct[CT_NEED_RESET]) {
    if (MD(auto_reset)) { reset_env(m, E, env, ws, cmd, state, lane); if (lane == 0) { a.terminated[env] = 0; a.truncated[env] = 0; } }
    L.active = 0;
    return;
  }
  L.sim_step = ct[CT_SIM_STEP] + 1;
  L.nstep = (uint32_t)ct[CT_NSTEP]; L.nobs = (uint32_t)ct[CT_NOBS];
  const int has_prev = ct[CT_HAS_DELAY];
  load_params(m, E, env, ws, lane);
  load_state(m, E, env, ws, lane);
  const float* act = a.action + (size_t)env * nu;
  // delay filter: Bernoulli(delay_prob) one-step hold
  const float v = uni(m, env, RNG_DELAY, L.nstep, 0);
  SYNC();
  const bool delay = (WS(W_SCAL)[2] > v) && has_prev;
  float rm = 0.f, tabs = 0.f, tsq = 0.f, tmax = 0.f;
  FOR_LANE(k, nu) {
    const float ak = act[k];
    const float f = delay ? E.delay_prev[(size_t)env * nu + k] : ak;
    E.delay_prev[(size_t)env * nu + k] = ak;
    WS(W_ACT)[k] = ak;
    const float target = f * LDG(TB(act_scale) + k);
    float tq;
    if (TB(act_mode)[k] == 0) {
      const float pf = LDG(TB(act_posfac) + k);
      const float q = WS(W_QPOS)[TB(act_qadr)[k]] * pf, qd = WS(W_QVEL)[TB(act_dof)[k]] * pf;
      tq = WS(W_KP)[k] * (target - q) + WS(W_KD)[k] * (0.f - qd);
      tq = tq * LDG(TB(act_gamma) + k);
    } else {
      tq = WS(W_KD)[k] * (target - WS(W_QVEL)[TB(act_dof)[k]]);
    }
    const float cl = LDG(TB(act_clip) + k);
    tq = fminf(cl, fmaxf(-cl, tq));
    WS(W_CTRL)[k] = tq;
    E.torque[(size_t)env * nu + k] = tq;
    const float df = ak - E.prev_action[(size_t)env * nu + k];
    rm += df * df;
    E.prev_action[(size_t)env * nu + k] = ak;
    E.last_action[(size_t)env * nu + k] = ak;
    tabs += fabsf(tq); tsq += tq * tq; tmax = fmaxf(tmax, fabsf(tq));
  }
  L.rm = sqrtf(wsum(rm) / (float)nu); L.tabs = wsum(tabs); L.tsq = wsum(tsq); L.tmax = wmaxf(tmax);
  if (lane == 0) { WSI(W_CNT)[CNT_NCON] = 0; WSI(W_CNT)[CNT_DROPPED] = 0; WSI(W_CNT)[CNT_NAN] = 0; WSI(W_CNT)[CNT_DROPPED_STEP] = 0; }
  SYNC();
}
// part 2 (after the sub-steps): cfrc_ext, termination, observations, state build, write-back, bookkeeping and statistics
DEV void step_epilogue(const ModelDev& m, const EnvArrays& E, int env, WSP ws, const StepArgs& a, int lane, const StepLocals& L, int iters) {
  const int nv = MD(nv), cd = MD(command_dim), sdim = MD(state_dim), nb = MD(nbody);
  int* ct = E.counters + (size_t)env * 8;
  const float* cmd = a.command ? a.command + (size_t)env * cd : nullptr;
  float* state = a.state_out + (size_t)env * sdim;
  const int sim_step = L.sim_step; const uint32_t nstep = L.nstep, nobs = L.nobs;
  const float rm = L.rm, tabs = L.tabs, tsq = L.tsq, tmax = L.tmax;
  const int ncon = WSI(W_CNT)[CNT_NCON], nan_count = WSI(W_CNT)[CNT_NAN], dropped_total = WSI(W_CNT)[CNT_DROPPED_STEP];
  PH_DECL;
  cfrc_ext(m, ws, ncon, lane);
  // termination: signed cfrc_ext component above threshold on the listed bodies
  int term = 0;
  { const float* cf = WS(W_CACC); const float thr = MO(term_threshold);
    FOR_LANE(i, 6 * MD(n_term_body)) term |= (cf[6 * TB(term_body)[i / 6] + i % 6] > thr);
    term = wor(term); }
  if (E.dbg_cfrc) FOR_LANE(i, 6 * nb) E.dbg_cfrc[(size_t)env * 6 * nb + i] = WS(W_CACC)[i];
  dump_contacts(m, E, env, ws, ncon, lane);
  get_obs(m, E, env, ws, nobs, lane);
  build_state(m, E, env, ws, sim_step, false, cmd, state, lane);
  store_state(m, E, env, ws, lane);
  const int trunc = (sim_step == MD(max_episode_steps));
  PH_MARK(PH_OBS);
  if (E.dbg_sens) FOR_LANE(i, 10) E.dbg_sens[(size_t)env * 10 + i] = WS(W_SENS)[i];
  if (E.dbg_qacc) FOR_LANE(i, nv) E.dbg_qacc[(size_t)env * nv + i] = WS(W_QACC)[i];
  if (lane == 0) {
    ct[CT_SIM_STEP] = sim_step; ct[CT_HAS_DELAY] = 1; ct[CT_NSTEP] = (int)(nstep + 1); ct[CT_NOBS] = (int)(nobs + 1);
    ct[CT_NAN] += nan_count; ct[CT_NCON] = ncon; ct[CT_NEED_RESET] = (term || trunc);
    a.terminated[env] = (uint8_t)term; a.truncated[env] = (uint8_t)trunc;
    const float* S = WS(W_SENS);
    float* info = E.info + (size_t)env * 4;
    info[0] = rm; info[1] = S[7]; info[2] = S[8]; info[3] = S[6];
    if (E.dbg_iters) E.dbg_iters[env] = iters;
    // reporter statistics: tracking error of lin_vel_x / lin_vel_y / ang_vel_yaw vs user command
    const float* uc = a.user_command ? a.user_command + (size_t)env * cd : cmd;
    float* st = E.stats + (size_t)env * ST__COUNT;
    st[ST_STEPS] += 1.f;
    if (uc) { if (cd > 0) st[ST_ERR_VX] += fabsf(S[7] - uc[0]); if (cd > 1) st[ST_ERR_VY] += fabsf(S[8] - uc[1]); if (cd > 2) st[ST_ERR_WZ] += fabsf(S[6] - uc[2]); }
    st[ST_RMSE] += rm; st[ST_ABS_TORQUE] += tabs; st[ST_SQ_TORQUE] += tsq; st[ST_MAX_TORQUE] = fmaxf(st[ST_MAX_TORQUE], tmax);
    st[ST_NCON] += (float)ncon; st[ST_ITERS] += (float)iters; st[ST_DROPPED] += (float)dropped_total; st[ST_NAN] += (float)nan_count;
    if (term || trunc) { st[ST_EPISODES] += 1.f; st[ST_TERMINATED] += (float)term; st[ST_SUCCESS] += (float)(trunc && !term); }
  }
  SYNC();
}
// One control step of one env by one warp, all stages back to back.
// `have_env` = this warp owns an env (false for the padding warps of the last CTA); `bsync` = CTA-wide phase barriers
DEV_NOINLINE void step_env(const ModelDev& m, const EnvArrays& E, int env, WSP ws, const StepArgs& a, int lane, int have_env = 1, int bsync = 0) { LANE_REFRESH();
  StepLocals L; L.active = 0;
  if (have_env) step_prologue(m, E, env, ws, a, lane, L);
  const int active = L.active;
  BSYNC_IF(bsync, 5);
  int iters = 0;
  const int fs = MD(frame_skip);
  NOUNROLL for (int s = 0; s < fs; ++s) iters += substep(m, ws, lane, active, bsync);
  if (!active) return;
  step_epilogue(m, E, env, ws, a, lane, L, iters);
}

// one raw physics sub-step (mj_step) from the stored state with ctrl = the last applied torque; debug / parity aid
// mirroring the oracle's orc_substep
DEV_NOINLINE void substep_env(const ModelDev& m, const EnvArrays& E, int env, WSP ws, int lane) { LANE_REFRESH();
  const int nv = MD(nv), nu = MD(nu), nb = MD(nbody);
  load_params(m, E, env, ws, lane);
  load_state(m, E, env, ws, lane);
  FOR_LANE(k, nu) WS(W_CTRL)[k] = E.torque[(size_t)env * nu + k];
  SYNC();
  if (lane == 0) { WSI(W_CNT)[CNT_NAN] = 0; WSI(W_CNT)[CNT_DROPPED_STEP] = 0; }
  SYNC();
  const int iters = substep(m, ws, lane);
  const int ncon = WSI(W_CNT)[CNT_NCON];
  cfrc_ext(m, ws, ncon, lane);
  if (E.dbg_cfrc) FOR_LANE(i, 6 * nb) E.dbg_cfrc[(size_t)env * 6 * nb + i] = WS(W_CACC)[i];
  dump_contacts(m, E, env, ws, ncon, lane);
  store_state(m, E, env, ws, lane);
  if (E.dbg_sens) FOR_LANE(i, 10) E.dbg_sens[(size_t)env * 10 + i] = WS(W_SENS)[i];
  if (E.dbg_qacc) FOR_LANE(i, nv) E.dbg_qacc[(size_t)env * nv + i] = WS(W_QACC)[i];
  if (lane == 0) { E.counters[(size_t)env * 8 + CT_NCON] = ncon; if (E.dbg_iters) E.dbg_iters[env] = iters; }
  SYNC();
}

// push event: qvel[0:2] = (R(q)^T v)[0:2], qvel[2] = v[2]  (flamingo_p_v3.py:257-266, quirk C-12)
DEV void push_env(const ModelDev& m, const EnvArrays& E, int env, const float* vel, int lane) {
  if (lane == 0) {
    const int nq = MD(nq), nv = MD(nv);
    const float* q = E.qpos + (size_t)env * nq + 3; float R[9]; quat_to_mat(R, q);
    float r[3]; m3tmulv(r, R, vel);
    float* qv = E.qvel + (size_t)env * nv;
    qv[0] = r[0]; qv[1] = r[1]; qv[2] = vel[2];
  }
}
