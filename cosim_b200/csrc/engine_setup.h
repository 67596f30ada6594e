// engine_setup.h -- host side of the engine: model blob -> ModelDev (read-only tables in device
// memory) + the per-warp shared-memory workspace layout.  Plain C++; memory comes from the
// `Uploader` the caller supplies (cudaMalloc+cudaMemcpy in engine.cu; malloc in tests/hostsim).
//
// Replaces, together with cosim_b200/model.py, what MjModel.from_xml_path hands to mj_step in the
// reference (gymnasium MujocoEnv.__init__, /root/reference/envs/flamingo_p_v3/flamingo_p_v3.py:94-100).
#pragma once
#include <string>
#include <vector>
#include <stdexcept>
#include <stdlib.h>
#include <algorithm>
#include "engine_core.h"

struct Uploader {
  void* (*up)(void* ctx, const void* host, size_t bytes);   // returns a device copy (never NULL for bytes > 0)
  void* ctx;
};

namespace setup {

template <class V> static std::vector<V> section(const void* blob, const char* name) {
  const cosim_blob_entry* e = cosim_blob_find(blob, name);
  if (!e) throw std::runtime_error(std::string("model blob: section '") + name + "' missing");
  const uint8_t* p = (const uint8_t*)blob + e->offset;
  std::vector<V> out(e->count);
  for (uint64_t i = 0; i < e->count; ++i) {
    if (e->dtype == 0) out[i] = (V)((const int32_t*)p)[i];
    else if (e->dtype == 1) out[i] = (V)((const float*)p)[i];
    else out[i] = (V)((const double*)p)[i];
  }
  return out;
}
template <class V> static const V* push(const Uploader& u, const std::vector<V>& v) {
  static const V zero[4] = {0, 0, 0, 0};
  if (v.empty()) return (const V*)u.up(u.ctx, zero, sizeof(zero));      // keep pointers valid
  return (const V*)u.up(u.ctx, v.data(), v.size() * sizeof(V));
}

// Small read-only tables (body tree, joints, dofs, actuators, geoms, observation layout ...) are packed into ONE arena:
// a CUDA kernel copies it to shared memory in its prologue and re-points the ModelDev fields listed in `slot_*`
// at the copy, so table look-ups are shared-memory reads instead of dependent L2 round trips.
struct Arena {
  std::vector<uint8_t> bytes;
  std::vector<std::pair<size_t, size_t>> slots;       // (byte offset of the pointer field in ModelDev, byte offset in the arena)
  template <class V> void add(ModelDev& m, const V*& field, const std::vector<V>& v) {
    const size_t off = (bytes.size() + 15) & ~(size_t)15, n = v.size() * sizeof(V);
    bytes.resize(off + (n < 16 ? 16 : n), 0);
    if (n) memcpy(bytes.data() + off, v.data(), n);
    slots.push_back({(size_t)((const char*)&field - (const char*)&m), off});
  }
};

static inline void build_model(const void* blob, size_t nbytes, uint64_t seed, uint32_t env_offset, const Uploader& u, ModelDev& m) {
  Arena arena;
  std::vector<uint16_t> sup_off16, ctab16;
  if (nbytes < 8 || memcmp(blob, "CSB1", 4) != 0) throw std::runtime_error("model blob: bad magic");
  memset(&m, 0, sizeof(m));
  std::vector<int> dims = section<int>(blob, "dims");
  std::vector<double> opts = section<double>(blob, "opts");
  if ((int)dims.size() < CD__count || (int)opts.size() < CO__count) throw std::runtime_error("model blob: dims/opts too short");
  for (int i = 0; i < 48; ++i) { m.dims[i] = i < (int)dims.size() ? dims[i] : 0; m.opts[i] = i < (int)opts.size() ? (float)opts[i] : 0.f; }
  const int nb = m.dims[CD_nbody], nv = m.dims[CD_nv], njnt = m.dims[CD_njnt];
  if (nv > 31) throw std::runtime_error("engine supports nv <= 31 (dof masks are 32-bit)");
  if (nb > 32) throw std::runtime_error("engine supports nbody <= 32 (body masks are 32-bit)");
  { std::vector<int> gb = section<int>(blob, "geom_body");      // contact ranges per body rely on geoms numbered in body order
    for (size_t g = 1; g < gb.size(); ++g) if (gb[g] < gb[g - 1]) throw std::runtime_error("geoms must be numbered in body order"); }

#define FSEC(field, name) arena.add(m, m.field, section<float>(blob, name))
#define ISEC(field, name) arena.add(m, m.field, section<int>(blob, name))
  std::vector<int> body_parent = section<int>(blob, "body_parent"), body_jntadr = section<int>(blob, "body_jntadr"),
                   body_jntnum = section<int>(blob, "body_jntnum"), body_dofadr = section<int>(blob, "body_dofadr"),
                   body_dofnum = section<int>(blob, "body_dofnum"), dof_parent = section<int>(blob, "dof_parent"),
                   dof_body = section<int>(blob, "dof_body");
  std::vector<int> body_jnt(nb, -1), subsize(nb, 1), dofmask(nb, 0), depth(nb, 0);
  for (int b = 1; b < nb; ++b) {
    if (body_jntnum[b] > 1) throw std::runtime_error("engine supports at most one joint per body");
    if (body_jntnum[b] == 1) body_jnt[b] = body_jntadr[b];
    if (body_parent[b] >= b) throw std::runtime_error("bodies must be in depth-first order");
    depth[b] = depth[body_parent[b]] + 1;
  }
  for (int b = nb - 1; b >= 1; --b) subsize[body_parent[b]] += subsize[b];
  // contiguous-subtree check (depth-first numbering): every body in (b, b+subsize) descends from b
  for (int b = 1; b < nb; ++b) for (int c = b + 1; c < b + subsize[b]; ++c) {
    int a = c; while (a > b) a = body_parent[a];
    if (a != b) throw std::runtime_error("body numbering is not depth-first");
  }
  for (int b = 1; b < nb; ++b) {
    int mask = dofmask[body_parent[b]];
    for (int k = body_dofadr[b]; k < body_dofadr[b] + body_dofnum[b]; ++k) mask |= (1 << k);
    dofmask[b] = mask;
  }
  int maxd = 0; for (int b = 0; b < nb; ++b) maxd = depth[b] > maxd ? depth[b] : maxd;
  std::vector<int> level_start(maxd + 2, 0), level_body;
  for (int l = 0; l <= maxd; ++l) { level_start[l] = (int)level_body.size(); for (int b = 0; b < nb; ++b) if (depth[b] == l) level_body.push_back(b); }
  level_start[maxd + 1] = (int)level_body.size();
  m.nlevels = maxd + 1;
  std::vector<int> mi, mj;
  for (int i = 0; i < nv; ++i) for (int j = i; j >= 0; j = dof_parent[j]) { mi.push_back(i); mj.push_back(j); }
  m.nmpair = (int)mi.size();

  arena.add(m, m.body_parent, body_parent); arena.add(m, m.body_jnt, body_jnt); arena.add(m, m.body_dofadr, body_dofadr); arena.add(m, m.body_dofnum, body_dofnum);
  arena.add(m, m.body_subsize, subsize); arena.add(m, m.body_dofmask, dofmask);
  arena.add(m, m.level_start, level_start); arena.add(m, m.level_body, level_body);
  arena.add(m, m.mpair_i, mi); arena.add(m, m.mpair_j, mj);
  { std::vector<int> pg = section<int>(blob, "pair_geom"); arena.add(m, m.pair_geom, pg);
    if (m.dims[CD_npair] > 32 * m.dims[CD_ngeom]) throw std::runtime_error("too many geom-geom candidate pairs for the task records (npair > 32 ngeom)"); }
  FSEC(body_pos, "body_pos"); FSEC(body_quat, "body_quat"); FSEC(body_ipos, "body_ipos"); FSEC(body_inertia, "body_inertia"); FSEC(body_mass, "body_mass");
  ISEC(jnt_type, "jnt_type"); ISEC(jnt_body, "jnt_body"); ISEC(jnt_qposadr, "jnt_qposadr"); ISEC(jnt_dofadr, "jnt_dofadr");
  ISEC(jnt_limited, "jnt_limited"); ISEC(jnt_actfrclimited, "jnt_actfrclimited");
  FSEC(jnt_pos, "jnt_pos"); FSEC(jnt_axis, "jnt_axis"); FSEC(jnt_range, "jnt_range"); FSEC(jnt_actfrcrange, "jnt_actfrcrange");
  arena.add(m, m.dof_body, dof_body); ISEC(dof_jnt, "dof_jnt"); arena.add(m, m.dof_parent, dof_parent); ISEC(dof_fl_random, "dof_fl_random");
  FSEC(dof_armature, "dof_armature"); FSEC(dof_damping, "dof_damping"); FSEC(dof_frictionloss, "dof_frictionloss");
  FSEC(qpos0, "qpos0");
  ISEC(act_dof, "act_dof"); ISEC(act_qadr, "act_qadr"); ISEC(act_mode, "act_mode"); ISEC(act_ctrllimited, "act_ctrllimited");
  FSEC(act_gear, "act_gear"); FSEC(act_ctrlrange, "act_ctrlrange"); FSEC(act_kp, "act_kp"); FSEC(act_kd, "act_kd");
  FSEC(act_scale, "act_scale"); FSEC(act_posfac, "act_posfac"); FSEC(act_gamma, "act_gamma"); FSEC(act_clip, "act_clip");
  ISEC(geom_type, "geom_type"); ISEC(geom_body, "geom_body"); ISEC(geom_vadr, "geom_vadr"); ISEC(geom_vnum, "geom_vnum"); ISEC(geom_fr_random, "geom_fr_random");
  FSEC(geom_size, "geom_size"); FSEC(geom_pos, "geom_pos"); FSEC(geom_quat, "geom_quat"); FSEC(geom_friction, "geom_friction");
  FSEC(geom_center, "geom_center"); FSEC(geom_rbound, "geom_rbound");
  m.hull_verts = push(u, section<float>(blob, "hull_verts")); m.hfield_data = push(u, section<float>(blob, "hfield_data"));     // big: stay in global memory
  { // support maps: candidate (x, y, z, index) quadruples per direction bucket
    std::vector<int> supadr = section<int>(blob, "geom_supadr"), sup_off = section<int>(blob, "sup_off"), sup_idx = section<int>(blob, "sup_idx");
    std::vector<int> gtype = section<int>(blob, "geom_type"), vadr = section<int>(blob, "geom_vadr");
    std::vector<float> hv = section<float>(blob, "hull_verts");
    std::vector<float> cand(4 * (sup_idx.size() > 0 ? sup_idx.size() : 1), 0.f);
    for (size_t g = 0; g < supadr.size(); ++g) {
      if (supadr[g] < 0) continue;
      for (int k = sup_off[supadr[g]]; k < sup_off[supadr[g] + 6 * 8 * 8]; ++k) {
        const int vi = sup_idx[k];
        for (int c = 0; c < 3; ++c) cand[4 * (size_t)k + c] = hv[3 * (size_t)(vadr[g] + vi) + c];
        memcpy(&cand[4 * (size_t)k + 3], &vi, 4);
      }
    }
    arena.add(m, m.geom_supadr, supadr); m.sup_off = push(u, sup_off); m.sup_cand = (const float4*)push(u, cand);
    // 16-bit copy of the bucket offsets, relative to each mesh's first candidate: goes into the shared-memory arena below
    // if that does not cost an env-warp (saves one dependent L2 round trip per hull support query)
    std::vector<int> supbase(supadr.size(), 0);
    sup_off16.assign(sup_off.size(), 0);
    bool ok16 = true;
    for (size_t g = 0; g < supadr.size(); ++g) {
      if (supadr[g] < 0) continue;
      const int b0 = sup_off[supadr[g]];
      supbase[g] = b0;
      for (int k = 0; k <= 6 * 8 * 8; ++k) { const int rel = sup_off[supadr[g] + k] - b0; if (rel < 0 || rel > 65535) ok16 = false; else sup_off16[supadr[g] + k] = (uint16_t)rel; }
    }
    if (!ok16) sup_off16.clear();
    arena.add(m, m.geom_supbase, supbase);
  }
  { // geom-frame bounding boxes [centre(3), half extents(3)]: conservative cull ahead of the geom-geom narrow phase
    std::vector<int> gtype = section<int>(blob, "geom_type"), vadr = section<int>(blob, "geom_vadr"), vnum = section<int>(blob, "geom_vnum");
    std::vector<float> hv = section<float>(blob, "hull_verts"), gsize = section<float>(blob, "geom_size");
    std::vector<float> aabb(6 * gtype.size(), 0.f);
    for (size_t g = 0; g < gtype.size(); ++g) {
      float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
      if (gtype[g] == 7) {
        for (int c = 0; c < 3; ++c) { lo[c] = 1e30f; hi[c] = -1e30f; }
        for (int i = 0; i < vnum[g]; ++i) for (int c = 0; c < 3; ++c) { const float v = hv[3 * (size_t)(vadr[g] + i) + c]; lo[c] = std::min(lo[c], v); hi[c] = std::max(hi[c], v); }
      } else if (gtype[g] == 2) { for (int c = 0; c < 3; ++c) { lo[c] = -gsize[3 * g]; hi[c] = gsize[3 * g]; } }
      else if (gtype[g] == 5) { lo[0] = lo[1] = -gsize[3 * g]; hi[0] = hi[1] = gsize[3 * g]; lo[2] = -gsize[3 * g + 1]; hi[2] = gsize[3 * g + 1]; }
      else { for (int c = 0; c < 3; ++c) { lo[c] = -gsize[3 * g + c]; hi[c] = gsize[3 * g + c]; } }
      for (int c = 0; c < 3; ++c) { aabb[6 * g + c] = 0.5f * (lo[c] + hi[c]); aabb[6 * g + 3 + c] = 0.5f * (hi[c] - lo[c]) * 1.0001f + 1e-6f; }
    }
    arena.add(m, m.geom_aabb, aabb);
    // bounding cylinders (geom frame) for the terrain culls: [centre(3), unit axis(3), radius, half length]; radius 0 = none.
    // Mesh hulls get the smallest of the three axis-aligned cylinders around the box centre when it is clearly smaller
    // than the box (wheels); cylinder geoms get themselves.
    std::vector<float> bcyl(8 * gtype.size(), 0.f);
    for (size_t g = 0; g < gtype.size(); ++g) {
      float* o = &bcyl[8 * g];
      if (gtype[g] == 5) { o[5] = 1.f; o[6] = gsize[3 * g] * 1.0001f + 1e-6f; o[7] = gsize[3 * g + 1] * 1.0001f + 1e-6f; continue; }
      if (gtype[g] != 7 || vnum[g] < 4) continue;
      const float* c = &aabb[6 * g]; const float* h = c + 3;
      double bestvol = 0.95 * 8.0 * h[0] * h[1] * h[2]; int best = -1; float bestr = 0.f;
      for (int a = 0; a < 3; ++a) {
        const int u = (a + 1) % 3, v = (a + 2) % 3;
        float r2 = 0.f;
        for (int i = 0; i < vnum[g]; ++i) { const float du = hv[3 * (size_t)(vadr[g] + i) + u] - c[u], dv = hv[3 * (size_t)(vadr[g] + i) + v] - c[v]; r2 = std::max(r2, du * du + dv * dv); }
        const double vol = 3.14159265358979 * r2 * 2.0 * h[a];
        if (vol < bestvol) { bestvol = vol; best = a; bestr = std::sqrt(r2); }
      }
      if (best >= 0) { o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3 + best] = 1.f; o[6] = bestr * 1.0001f + 1e-6f; o[7] = h[best]; }
    }
    arena.add(m, m.geom_bcyl, bcyl);
  }
  { // max height per 8 x 8 block of cells (normalised data, like hfield_data): early-out for geoms above a fine raster
    std::vector<float> hf = section<float>(blob, "hfield_data");
    const int nrow = m.dims[CD_hf_nrow], ncol = m.dims[CD_hf_ncol];
    m.hf_max8 = nullptr; m.hf_mrow = m.hf_mcol = 0;
    if (m.dims[CD_ground_type] == 1 && nrow > 1 && ncol > 1 && hf.size() >= (size_t)nrow * ncol) {
      const int mr = ((nrow - 1) >> 3) + 1, mc = ((ncol - 1) >> 3) + 1;
      std::vector<float> mx((size_t)mr * mc, 0.f);
      for (int br = 0; br < mr; ++br) for (int bc = 0; bc < mc; ++bc) {
        float v = 0.f;
        for (int r = 8 * br; r <= std::min(8 * br + 8, nrow - 1); ++r) for (int c = 8 * bc; c <= std::min(8 * bc + 8, ncol - 1); ++c) v = std::max(v, hf[(size_t)r * ncol + c]);
        mx[(size_t)br * mc + bc] = v;
      }
      m.hf_max8 = push(u, mx); m.hf_mrow = mr; m.hf_mcol = mc;
      const double cellx = 2.0 * opts[CO_hf_sx] / (ncol - 1), celly = 2.0 * opts[CO_hf_sy] / (nrow - 1);
      m.hf_fine = std::min(cellx, celly) < 0.2 ? 1 : 0;
      { const char* e = getenv("COSIM_HF_FINE"); if (e) m.hf_fine = atoi(e) ? 1 : 0; }
    }
  }
  { // Cholesky pair tables (engine_core.h chol_factor).  tri: all pairs (i, k <= i), rows in increasing order, so that the first
    // j (j + 1) / 2 entries are the pairs below row j (dense elimination of column j); ctab / coff: per column j the pairs of
    // its ancestors (tree-sparse elimination; kept in global memory, it is read once per column by all warps alike)
    std::vector<int> tri, ctab, coff(nv + 1, 0);
    for (int i = 0; i < nv; ++i) for (int k = 0; k <= i; ++k) tri.push_back((i << 8) | k);
    for (int j = 0; j < nv; ++j) {
      std::vector<int> anc;
      for (int a = dof_parent[j]; a >= 0; a = dof_parent[a]) anc.push_back(a);
      coff[j] = (int)ctab.size();
      for (size_t x = anc.size(); x-- > 0;) for (size_t y = anc.size(); y-- > x;) ctab.push_back((anc[x] << 8) | anc[y]);     // anc is descending: anc[x] >= anc[y] for y >= x
    }
    coff[nv] = (int)ctab.size();
    arena.add(m, m.tri, tri); arena.add(m, m.coff, coff); m.ctab = push(u, ctab); m.ctab16 = nullptr;
    ctab16.assign(ctab.begin(), ctab.end());
  }
  { std::vector<float> gf = section<float>(blob, "ground_friction"); for (int i = 0; i < 4; ++i) m.ground_friction[i] = i < (int)gf.size() ? gf[i] : 0.f; }
  ISEC(eq_body1, "eq_body1"); ISEC(eq_body2, "eq_body2");
  FSEC(eq_anchor1, "eq_anchor1"); FSEC(eq_anchor2, "eq_anchor2"); FSEC(eq_solref, "eq_solref"); FSEC(eq_solimp, "eq_solimp");
  { std::vector<float> p = section<float>(blob, "imu_pos"), q = section<float>(blob, "imu_quat");
    for (int i = 0; i < 3; ++i) m.imu_pos[i] = p[i];
    for (int i = 0; i < 4; ++i) m.imu_quat[i] = q[i]; }
  ISEC(dofpos_qadr, "dofpos_qadr"); ISEC(dofvel_dadr, "dofvel_dadr"); ISEC(initnoise_qadr, "initnoise_qadr"); ISEC(term_body, "term_body");
  ISEC(massnoise_body, "massnoise_body"); FSEC(dofpos_fac, "dofpos_fac"); FSEC(dofvel_fac, "dofvel_fac");
  ISEC(sobs_kind, "sobs_kind"); ISEC(sobs_dim, "sobs_dim"); ISEC(sobs_interval, "sobs_interval"); ISEC(sobs_off, "sobs_off");
  ISEC(nobs_kind, "nobs_kind"); ISEC(nobs_dim, "nobs_dim"); ISEC(nobs_interval, "nobs_interval"); ISEC(nobs_off, "nobs_off");
  FSEC(sobs_scale, "sobs_scale"); FSEC(nobs_scale, "nobs_scale"); FSEC(noise, "noise");
#undef FSEC
#undef ISEC
  // randomisation ranges: [slide, torsional, rolling, frictionloss, delay, load, kp scale, kd scale]
  const int lo_idx[8] = {CO_slide_lo, CO_tors_lo, CO_roll_lo, CO_floss_lo, CO_delay_lo, CO_load_lo, CO_kp_lo, CO_kd_lo};
  for (int i = 0; i < 8; ++i) { m.rnd_lo[i] = (float)opts[lo_idx[i]]; m.rnd_span[i] = (float)(opts[lo_idx[i] + 1] - opts[lo_idx[i]]); }
  { const char* e = getenv("COSIM_BSYNC_MASK"); m.bsync_mask = e ? atoi(e) : (m.dims[CD_ground_type] == 1 ? 447 : 63); /* all phase barriers; on height fields also the ones inside the collision and Newton phases (+2.4 % on flamingo / rocky, -1.3 % on the plane) */ }
  m.seed_lo = (uint32_t)seed; m.seed_hi = (uint32_t)(seed >> 32); m.env_offset = env_offset;

  // ---- workspace layout (floats), every field padded to 4 floats
  const int nq = m.dims[CD_nq], nu = m.dims[CD_nu], ng = m.dims[CD_ngeom], neq = m.dims[CD_neq], ncap = m.dims[CD_ncon_max];
  const int nh = m.dims[CD_hm_res_x] * m.dims[CD_hm_res_y];
  const int nraw = m.dims[CD_n_dofpos] + m.dims[CD_n_dofvel] + 9 + nu + nh;
  m.cr_stride = CR_STRIDE;                      // odd record stride: records of neighbouring contacts fall on different shared-memory banks
  auto pad4 = [](int n) { return (n + 3) & ~3; };
  // Lays the fields out back to back, overlaying fields whose lifetimes never overlap (shared memory per env bounds how many
  // env-warps an SM holds, and the step is latency-bound, so every KB counts).  K = contact records kept in shared memory.
  auto layout = [&](int K, int* off) -> int {
    int size[W__COUNT];
    for (int i = 0; i < W__COUNT; ++i) size[i] = 0;
    size[W_QPOS] = nq; size[W_QVEL] = nv; size[W_CTRL] = nu; size[W_WARM] = nv; size[W_QACC] = nv;
    size[W_XPOS] = 3 * nb; size[W_XQUAT] = 4 * nb; size[W_XMAT] = 9 * nb; size[W_XIPOS] = 3 * nb;
    size[W_XANCHOR] = 3 * njnt; size[W_XAXIS] = 3 * njnt; size[W_GXPOS] = 3 * ng; size[W_GXMAT] = 9 * ng; size[W_SCOM] = 4;
    size[W_CINERT] = 10 * nb; size[W_CRB] = 10 * nb; size[W_CDOF] = 6 * nv; size[W_CDOFDOT] = 6 * nv;
    size[W_CVEL] = 6 * nb; size[W_CACC] = 6 * nb; size[W_CFRC] = 6 * nb; size[W_BUF] = 6 * nv;
    size[W_M] = nv * nv; size[W_A] = nv * nv; size[W_INVD] = nv;
    size[W_FSMOOTH] = size[W_ASMOOTH] = size[W_FCON] = size[W_GRAD] = size[W_SEARCH] = size[W_MV] = size[W_MA] = size[W_TMPV] = size[W_TMPW] = nv;
    size[W_BMASS] = nb; size[W_INVWD] = nv; size[W_INVWB] = nb; size[W_FLOSS] = nv; size[W_GMU] = ng; size[W_SCAL] = 8;
    size[W_FR_D] = nv; size[W_FR_AREF] = nv; size[W_LM_SIGN] = njnt; size[W_LM_D] = njnt; size[W_LM_AREF] = njnt;
    size[W_CN_REC] = K * m.cr_stride; size[W_GPTR] = 4; size[W_RING] = RING_SIZE + 20; size[W_BV] = 6 * nb; size[W_CSTART] = nb + 1; size[W_BS] = 21 * nb; size[W_CN_J] = FEW_CONTACTS * 3 * nv;
    size[W_EQ_J] = 3 * neq * nv; size[W_EQ_D] = size[W_EQ_AREF] = size[W_EQ_X] = size[W_EQ_V] = size[W_EQ_F] = 3 * neq;
    size[W_SENS] = 12; size[W_RAW] = nraw; size[W_ACT] = nu; size[W_FILT] = 0; size[W_KP] = nu; size[W_KD] = nu; size[W_GTASK] = 8 * ng; size[W_CNT] = 8; size[W_PAXIS] = m.dims[CD_npair] > 0 ? 4 * PAXIS_SLOTS : 0;
    int o = 0;
    bool placed[W__COUNT];
    for (int i = 0; i < W__COUNT; ++i) placed[i] = false;
    auto place = [&](int f, int at) { off[f] = at; placed[f] = true; };
    { // composite-inertia sums (crb(), phase 1) | collision task table (phase 2) | body velocities / accelerations (from com_vel
      // on, phase 3; cfrc_ext output after the last solve) | per-body contact inertias of the Hessian assembly (phase 4)
      // | explicit Jacobians of the few-contact solver path (end of phase 3 .. phase 4; never together with the contact inertias)
      const int n = std::max(std::max(std::max(std::max(pad4(size[W_CRB]), pad4(size[W_GTASK])), pad4(size[W_CVEL]) + pad4(size[W_CACC])), pad4(size[W_BS])), pad4(size[W_CN_J]));
      place(W_CRB, o); place(W_GTASK, o); place(W_CVEL, o); place(W_CACC, o + pad4(size[W_CVEL])); place(W_BS, o); place(W_CN_J, o); o += n; }
    { // crb() scratch (phase 1; the host emulation also uses it inside solves) | collision task ring (phase 2) | cdof_dot (com_vel .. rne_bias, phase 3)
      const int n = std::max(std::max(pad4(size[W_BUF]), pad4(size[W_RING])), pad4(size[W_CDOFDOT]));
      place(W_BUF, o); place(W_RING, o); place(W_CDOFDOT, o); o += n; }
    { // body forces of rne_bias (phase 3) | body velocities / wrenches of the Jacobian-free contact rows (phase 4)
      const int n = std::max(pad4(size[W_CFRC]), pad4(size[W_BV])); place(W_CFRC, o); place(W_BV, o); o += n; }
    { // observation staging (after the sub-steps) over the shared-memory contact records (dead after cfrc_ext / the contact dump)
      const int n = std::max(pad4(size[W_RAW]), pad4(size[W_CN_REC])); place(W_RAW, o); place(W_CN_REC, o); o += n; }
    { // composite inertias live from com_pos to rne_bias (phases 1-3); gradient, search direction and M * search exist only
      // from the Newton solve on (phase 4) and are rewritten before every use
      const int n = std::max(pad4(size[W_CINERT]), pad4(size[W_GRAD]) + pad4(size[W_SEARCH]) + pad4(size[W_MV]));
      place(W_CINERT, o); place(W_GRAD, o); place(W_SEARCH, o + pad4(size[W_GRAD])); place(W_MV, o + pad4(size[W_GRAD]) + pad4(size[W_SEARCH])); o += n; }
    { // world joint anchors / axes are consumed by com_pos (phase 1); friction-loss and limit rows are built in phase 3
      const int a = pad4(size[W_XANCHOR]) + pad4(size[W_XAXIS]);
      const int b = pad4(size[W_FR_D]) + pad4(size[W_FR_AREF]) + pad4(size[W_LM_SIGN]) + pad4(size[W_LM_D]) + pad4(size[W_LM_AREF]);
      place(W_XANCHOR, o); place(W_XAXIS, o + pad4(size[W_XANCHOR]));
      int q = o; place(W_FR_D, q); q += pad4(size[W_FR_D]); place(W_FR_AREF, q); q += pad4(size[W_FR_AREF]);
      place(W_LM_SIGN, q); q += pad4(size[W_LM_SIGN]); place(W_LM_D, q); q += pad4(size[W_LM_D]); place(W_LM_AREF, q);
      o += std::max(a, b); }
    for (int i = 0; i < W__COUNT; ++i) if (!placed[i]) { off[i] = o; o += pad4(size[i]); }
    return o;
  };
  // How many contact records stay in shared memory: the largest K (4 .. 24, at most the capacity) that still gives the
  // largest number of env-warps per SM (<= 20: five warps of 96 registers per SM sub-partition); the rest of an env's
  // contacts spills into the warp's global-memory slot.  COSIM_CN_K overrides (experiments).
  {
    const size_t budget = 224 * 1024;
    int tmp[80];
    // Two kernel builds exist for the fast path (engine.cu: 640 threads = 20 env-warps of 96 registers; engine_w24.cu: 768 threads
    // = 24 env-warps of 80 registers).  A model whose workspace (with 8 contact records in shared memory) lets 24 env-warps
    // share an SM runs the second one: +6 % on flamingo_p_v3 (B200, round 2) -- more environments in flight beat the extra
    // spills; models that stay at <= 20 warps for shared memory (flamingo_light 18, w4 12, humanoid 11) lose 2 - 5 % on the
    // 80-register build and keep the first.  COSIM_WPB_CAP = 20 forces the first build (A / B runs).
    size_t wcap = 24;
    { const int cdm = m.dims[CD_condim];
      const bool general_path = (cdm == 1 || cdm == 4 || cdm == 6) || m.dims[CD_cone] == 1 || m.dims[CD_solver] == 1 || (opts[CO_impratio] > 0 && opts[CO_impratio] != 1.0);
      if (general_path) wcap = 20; }
    { const char* e = getenv("COSIM_WPB_CAP"); if (e && (atoi(e) == 20 || atoi(e) == 24)) wcap = (size_t)atoi(e); }
    auto warps = [&](int K, size_t extra) {
      const size_t head = (sizeof(ModelDev) + 15) / 16 * 16 + ((arena.bytes.size() + extra + 31) & ~(size_t)15) + 64, per = (size_t)layout(K, tmp) * 4;
      return (int)std::min<size_t>(wcap, (budget - head) / per); };
    const int kmin = std::min(4, ncap), kmax = std::min(64, ncap);
    if (wcap > 20 && warps(std::min(8, ncap), 0) <= 20) wcap = 20;      // shared memory holds no more than 20 env-warps anyway
    // 16-bit copy of the support-map bucket offsets in the arena (one dependent L2 round trip less per hull support query)
    // if that does not cost an env-warp
    if (!sup_off16.empty() && warps(kmin, sup_off16.size() * 2) == warps(kmin, 0) && warps(kmin, 0) >= 1) arena.add(m, m.sup_off16, sup_off16);
    // likewise the pair lists of the tree-sparse Cholesky (read once per column of every factorization)
    if (!ctab16.empty() && warps(kmin, ctab16.size() * 2) == warps(kmin, 0) && warps(kmin, 0) >= 1) arena.add(m, m.ctab16, ctab16);
    // Coarse rasters / the plane: a handful of contacts per env, keep every env-warp the SM can hold.  Fine rasters (cells
    // under 5 cm: a wheel alone touches a dozen prisms) give up env-warps, down to 12 per SM, for records in shared memory.
    const double cell = m.dims[CD_ground_type] == 1 ? 2.0 * opts[CO_hf_sx] / std::max(1, m.dims[CD_hf_ncol] - 1) : 1.0;
    const int best = warps(kmin, 0), floor_w = cell < 0.05 ? std::min(best, 12) : best;
    int K = kmin;
    for (int k = kmin; k <= kmax; ++k) if (warps(k, 0) >= floor_w) K = k;
    K = std::max(K, std::min(8, ncap));          // at least 8 records in shared memory, even if that costs an env-warp
    { const char* e = getenv("COSIM_CN_K"); if (e && atoi(e) >= 1) K = std::min(atoi(e), ncap); }
    m.cn_k = K;
    m.wpb_cap = (int)wcap;
    // 24 env-warps on a height field (flamingo_p_v3): Newton iterations of the whole CTA in lock step (bit 9; +3 % there, a loss for
    // the humanoid's 11 and w4's 12 - 18 warps and on the plane, profiles/r02_experiments.md)
    if (!getenv("COSIM_BSYNC_MASK") && wcap > 20 && m.dims[CD_ground_type] == 1) m.bsync_mask |= 512;
    m.ws_floats = layout(K, m.off);
    m.gslot_floats = (unsigned long long)(((size_t)std::max(0, ncap - K) * m.cr_stride + 31) & ~(size_t)31);
    m.gscratch = nullptr;
    // general constraint path (engine_general.h): anything but condim 3 / pyramidal / Newton / impratio 1
    const int condim = m.dims[CD_condim];
    { // box-box pairs get a scratch block per lane for their up to 8 contacts
      std::vector<int> pg = section<int>(blob, "pair_geom"), gt = section<int>(blob, "geom_type");
      m.n_boxbox = 0; for (size_t i = 0; i + 1 < pg.size(); i += 2) m.n_boxbox += (gt[pg[i]] == 6 && gt[pg[i + 1]] == 6);
      m.bb_off = m.gslot_floats;
      if (m.n_boxbox) m.gslot_floats += 32 * 56;
    }
    m.general = ((condim == 1 || condim == 4 || condim == 6) || m.dims[CD_cone] == 1 || m.dims[CD_solver] == 1 || (opts[CO_impratio] > 0 && opts[CO_impratio] != 1.0)) ? 1 : 0;
    if (m.general) m.bsync_mask &= ~(512 | 1024);      // the lock-step Newton iterations / narrow-phase batches exist on the fast path only
    m.gen_rows = 0; m.gen_off = m.gslot_floats;
    if (m.general) {
      const int cd = (condim == 1 || condim == 4 || condim == 6) ? condim : 3, rpc = cd == 1 ? 1 : (m.dims[CD_cone] == 1 ? cd : 2 * (cd - 1));
      const int NC = std::min(ncap, (int)GEN_MAX_CON);
      m.gen_rows = 3 * neq + nv + njnt + rpc * NC;
      m.gslot_floats += (unsigned long long)((gen_region_floats(m.gen_rows, nv, NC) + 31) & ~(size_t)31);
    }
  }
  // upload the arena and point the fields at it; CTA-shared area in front of the per-warp workspaces: [ModelDev copy | arena copy]
  if (arena.slots.size() > sizeof(m.slot_field) / sizeof(m.slot_field[0])) throw std::runtime_error("too many model tables for ModelDev::slot_field");
  arena.bytes.resize((arena.bytes.size() + 15) & ~(size_t)15, 0);
  if (arena.bytes.size() / 16 > 65535) throw std::runtime_error("model table arena too large");
  const uint8_t* dev = (const uint8_t*)u.up(u.ctx, arena.bytes.data(), arena.bytes.size());
  m.arena_g = dev; m.arena_bytes = (int)arena.bytes.size(); m.nslots = (int)arena.slots.size();
  for (size_t i = 0; i < arena.slots.size(); ++i) {
    *(const uint8_t**)((char*)&m + arena.slots[i].first) = dev + arena.slots[i].second;
    m.slot_field[i] = (uint16_t)arena.slots[i].first; m.slot_off16[i] = (uint16_t)(arena.slots[i].second / 16);
  }
  m.shared_floats = (int)((sizeof(ModelDev) + 15) / 16 * 4) + m.arena_bytes / 4;
}

// per-env arrays: sizes in elements per env
struct EnvDims { int nq, nv, nu, nb, ng, obsbuf, cache, nh, ncon, cd, sd; };
static inline EnvDims env_dims(const ModelDev& m) {
  EnvDims d;
  d.nq = m.dims[CD_nq]; d.nv = m.dims[CD_nv]; d.nu = m.dims[CD_nu]; d.nb = m.dims[CD_nbody]; d.ng = m.dims[CD_ngeom];
  d.obsbuf = m.dims[CD_stack_size] * m.dims[CD_stacked_dim]; if (d.obsbuf < 1) d.obsbuf = 1;
  d.cache = m.dims[CD_cache_dim] < 1 ? 1 : m.dims[CD_cache_dim];
  d.nh = m.dims[CD_hm_res_x] * m.dims[CD_hm_res_y]; d.ncon = m.dims[CD_ncon_max];
  d.cd = m.dims[CD_command_dim]; d.sd = m.dims[CD_state_dim];
  return d;
}


// ---- per-env arrays: allocation + the name table behind cosim_get / cosim_set
struct Field { const char* name; void* ptr; int dim; int is_int; };
typedef void* (*ZAlloc)(void* ctx, size_t bytes);   // zero-initialised device memory

static inline void alloc_env(const ModelDev& m, int N, EnvArrays& E, ZAlloc za, void* ctx) {
  const EnvDims d = env_dims(m);
  memset(&E, 0, sizeof(E));
  E.N = N;
  const size_t n = (size_t)N;
#define FA(field, dim) E.field = (float*)za(ctx, n * (size_t)(dim) * sizeof(float))
  FA(qpos, d.nq); FA(qvel, d.nv); FA(warm, d.nv); FA(paxis, 4 * PAXIS_SLOTS);
  FA(body_mass, d.nb); FA(invw_dof, d.nv); FA(invw_body, d.nb); FA(floss, d.nv); FA(gmu, d.ng); FA(scal, 4); FA(kp, d.nu); FA(kd, d.nu);
  FA(prev_action, d.nu); FA(delay_prev, d.nu); FA(obs_buffer, d.obsbuf); FA(freq_cache, d.cache); FA(torque, d.nu); FA(info, 4); FA(last_action, d.nu);
  FA(stats, ST__COUNT);
#undef FA
  E.counters = (int*)za(ctx, n * 8 * sizeof(int));
}
static inline void alloc_debug(const ModelDev& m, int N, EnvArrays& E, ZAlloc za, void* ctx) {
  const EnvDims d = env_dims(m);
  const size_t n = (size_t)N;
  if (E.dbg_contacts) return;
  E.dbg_contacts = (float*)za(ctx, n * (size_t)d.ncon * 10 * sizeof(float));
  E.dbg_heightmap = (float*)za(ctx, n * (size_t)(d.nh < 1 ? 1 : d.nh) * sizeof(float));
  E.dbg_hmcell = (int*)za(ctx, n * (size_t)(d.nh < 1 ? 1 : d.nh) * sizeof(int));
  E.dbg_cfrc = (float*)za(ctx, n * (size_t)d.nb * 6 * sizeof(float));
  E.dbg_sens = (float*)za(ctx, n * 10 * sizeof(float));
  E.dbg_qacc = (float*)za(ctx, n * (size_t)d.nv * sizeof(float));
  E.dbg_iters = (int*)za(ctx, n * sizeof(int));
}
static inline std::vector<Field> env_fields(const ModelDev& m, const EnvArrays& E) {
  const EnvDims d = env_dims(m);
  std::vector<Field> f = {
    {"qpos", E.qpos, d.nq, 0}, {"qvel", E.qvel, d.nv, 0}, {"qacc_warmstart", E.warm, d.nv, 0},
    {"body_mass", E.body_mass, d.nb, 0}, {"invweight_dof", E.invw_dof, d.nv, 0}, {"invweight_body", E.invw_body, d.nb, 0},
    {"frictionloss", E.floss, d.nv, 0}, {"geom_mu", E.gmu, d.ng, 0}, {"scal", E.scal, 4, 0}, {"kp", E.kp, d.nu, 0}, {"kd", E.kd, d.nu, 0},
    {"torque", E.torque, d.nu, 0}, {"info", E.info, 4, 0}, {"last_action", E.last_action, d.nu, 0}, {"prev_action", E.prev_action, d.nu, 0},
    {"obs_buffer", E.obs_buffer, d.obsbuf, 0}, {"counters", E.counters, 8, 1}, {"stats", E.stats, ST__COUNT, 0},
    {"contacts", E.dbg_contacts, d.ncon * 10, 0}, {"heightmap", E.dbg_heightmap, d.nh < 1 ? 1 : d.nh, 0},
    {"hm_cell", E.dbg_hmcell, d.nh < 1 ? 1 : d.nh, 1}, {"cfrc_ext", E.dbg_cfrc, d.nb * 6, 0}, {"sens", E.dbg_sens, 10, 0},
    {"qacc", E.dbg_qacc, d.nv, 0}, {"iters", E.dbg_iters, 1, 1},
  };
  return f;
}
static inline int dim_by_name(const ModelDev& m, const std::string& n) {
  static const char* names[] = {"nq", "nv", "nu", "nbody", "njnt", "ngeom", "nhullvert", "neq", "ground_type", "hf_nrow", "hf_ncol",
    "frame_skip", "iterations", "ls_iterations", "ccd_iterations", "ncon_max", "hm_res_x", "hm_res_y", "state_dim", "stack_size",
    "stacked_dim", "nonstacked_dim", "command_dim", "n_term_body", "n_dofpos", "n_dofvel", "n_initnoise", "max_episode_steps",
    "lin_vel_f32", "n_sobs", "n_nobs", "cache_dim", "n_state_pos", "n_state_vel", "position_command", "nefc_max", "imu_body",
    "n_massnoise", "base_body", "zero_noise", "auto_reset", "nfl", "nlimit_max", "npair", "condim", "cone", "solver"};
  if (n == "action_dim") return m.dims[CD_nu];
  for (int i = 0; i < CD__count; ++i) if (n == names[i]) return m.dims[i];
  return -1;
}

}  // namespace setup
