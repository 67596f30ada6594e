// oracle_capi.cpp -- C API over the templated oracle (ctypes binding in oracle/oracle.py).
// TEST INFRASTRUCTURE ONLY (see oracle.hpp).
#include <cstring>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle.hpp"

namespace {
struct IOracle {
  virtual ~IOracle() {}
  virtual int num_envs() const = 0;
  virtual void reset(const uint8_t* mask, const double* command, float* state) = 0;
  virtual void step(const double* action, const double* command, float* state, uint8_t* term, uint8_t* trunc, int nthreads) = 0;
  virtual int get(const char* name, double* out) = 0;
  virtual int set(const char* name, const double* in) = 0;
  virtual void forward_all() = 0;
  virtual void substep_all() = 0;
  virtual int contacts(int env, double* out, int cap) = 0;
  virtual int contact_forces(int env, double* out, int cap) = 0;
  virtual int box_box(const double* geoms, double* out) = 0;
  virtual void rne_post_all() = 0;
  virtual double ray_hfield(double x, double y) = 0;
  virtual int convex_pair(const double* geoms, double* out) = 0;
};

template <class T> struct OracleT : IOracle {
  orc::Engine<T> E;
  OracleT(const void* blob, int n, uint64_t seed, uint32_t off) : E(blob, n, seed, off) {}
  int num_envs() const override { return E.N; }
  void reset(const uint8_t* mask, const double* command, float* state) override {
    const int cd = E.m.D(CD_command_dim), sd = E.m.D(CD_state_dim);
    for (int e = 0; e < E.N; ++e) {
      if (mask && !mask[e]) continue;
      std::vector<T> c(std::max(1, cd));
      for (int i = 0; i < cd; ++i) c[i] = command ? (T)command[(size_t)e * cd + i] : T(0);
      E.reset(e, c.data(), state + (size_t)e * sd);
    }
  }
  void step(const double* action, const double* command, float* state, uint8_t* term, uint8_t* trunc, int nthreads) override {
    const int cd = E.m.D(CD_command_dim), sd = E.m.D(CD_state_dim), nu = E.nu;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int e = 0; e < E.N; ++e) {
      std::vector<T> a(nu), c(std::max(1, cd));
      for (int i = 0; i < nu; ++i) a[i] = (T)action[(size_t)e * nu + i];
      for (int i = 0; i < cd; ++i) c[i] = command ? (T)command[(size_t)e * cd + i] : T(0);
      E.step(e, a.data(), c.data(), state + (size_t)e * sd, term + e, trunc + e);
    }
  }
  void forward_all() override { for (auto& d : E.envs) E.forward(d); }
  void substep_all() override { for (auto& d : E.envs) E.substep(d); }
  void rne_post_all() override { for (auto& d : E.envs) E.cfrc_ext(d); }
  double ray_hfield(double x, double y) override { int cell; return (double)E.hfield_height((T)x, (T)y, &cell); }
  // known-answer hook for the geom-geom narrow phase: geoms = 2 x [type, size(3), pos(3), mat(9)] of primitives;
  // out = [depth, normal(3), pos(3)]; returns 0 on penetration, -1 otherwise
  int convex_pair(const double* geoms, double* out) override {
    typename orc::Engine<T>::Geom G[2]; T pos[2][3], mat[2][9];
    for (int i = 0; i < 2; ++i) {
      const double* g = geoms + 16 * i;
      G[i].type = (int)g[0]; for (int k = 0; k < 3; ++k) { G[i].size[k] = (T)g[1 + k]; pos[i][k] = (T)g[4 + k]; G[i].center[k] = (T)g[4 + k]; }
      for (int k = 0; k < 9; ++k) mat[i][k] = (T)g[7 + k];
      G[i].pos = pos[i]; G[i].mat = mat[i]; G[i].verts = nullptr; G[i].nvert = 0;
    }
    typename orc::Engine<T>::ShapeA A{nullptr, &G[0]};
    T depth = 0, dir[3] = {0, 0, 0}, cp[3] = {0, 0, 0};
    int r = E.mpr_core(A, G[0].center, G[1], &depth, dir, cp);
    if (r == 0 && !(dir[0] == 0 && dir[1] == 0 && dir[2] == 0)) orc::Engine<T>::fix_normal(G[0], G[1], cp, dir);      // as mjc_Convex does
    out[0] = depth; for (int k = 0; k < 3; ++k) { out[1 + k] = dir[k]; out[4 + k] = cp[k]; }
    return r;
  }
  // known-answer hook for mjc_BoxBox: geoms = 2 x [type, size(3), pos(3), mat(9)]; out = up to 8 x [pos(3), normal(3), dist]
  int box_box(const double* geoms, double* out) override {
    T sz[2][3], pos[2][3], mat[2][9], bb[8][7];
    for (int i = 0; i < 2; ++i) { const double* g = geoms + 16 * i; for (int k = 0; k < 3; ++k) { sz[i][k] = (T)g[1 + k]; pos[i][k] = (T)g[4 + k]; } for (int k = 0; k < 9; ++k) mat[i][k] = (T)g[7 + k]; }
    const int n = orc::Engine<T>::box_box(pos[0], mat[0], sz[0], pos[1], mat[1], sz[1], bb);
    for (int c = 0; c < n; ++c) for (int k = 0; k < 7; ++k) out[7 * c + k] = (double)bb[c][k];
    return n;
  }
  template <class V> static void put(double* out, const V& v, size_t stride, int e) { for (size_t i = 0; i < stride; ++i) out[(size_t)e * stride + i] = (double)v[i]; }
  int get(const char* name_, double* out) override {
    std::string n(name_);
    for (int e = 0; e < E.N; ++e) {
      auto& d = E.envs[e];
#define G(key, vec) if (n == key) { put(out, vec, vec.size(), e); continue; }
      G("qpos", d.qpos) G("qvel", d.qvel) G("ctrl", d.ctrl) G("qacc_warmstart", d.warm) G("qacc", d.qacc) G("qacc_smooth", d.qacc_smooth)
      G("qfrc_bias", d.qfrc_bias) G("qfrc_smooth", d.qfrc_smooth) G("qfrc_constraint", d.qfrc_constraint) G("M", d.M)
      G("xpos", d.xpos) G("xquat", d.xquat) G("xipos", d.xipos) G("geom_xpos", d.geom_xpos) G("cfrc_ext", d.cfrc_ext)
      G("body_mass", d.body_mass) G("geom_friction", d.geom_friction) G("dof_frictionloss", d.dof_frictionloss)
      G("kp", d.kp) G("kd", d.kd) G("dof_invweight0", d.dof_invweight0) G("body_invweight0", d.body_invweight0)
      G("torque", d.torque) G("action", d.action) G("heightmap", d.heightmap) G("obs_buffer", d.obs_buffer)
#undef G
      if (n == "subtree_com") { for (int i = 0; i < 3; ++i) out[3 * e + i] = d.subtree_com[i]; continue; }
      if (n == "ground_friction") { for (int i = 0; i < 3; ++i) out[3 * e + i] = d.ground_friction[i]; continue; }
      if (n == "delay_prob") { out[e] = d.delay_prob; continue; }
      if (n == "meaninertia") { out[e] = d.meaninertia; continue; }
      if (n == "ncon") { out[e] = (double)d.con.size(); continue; }
      if (n == "nefc") { out[e] = d.nefc; continue; }
      if (n == "solver_iter") { out[e] = d.solver_iter; continue; }
      if (n == "ls_evals") { out[e] = (double)d.ls_evals; continue; }
      if (n == "ncon_dropped") { out[e] = d.ncon_dropped; continue; }
      if (n == "sim_step") { out[e] = d.sim_step; continue; }
      if (n == "nan_count") { out[e] = d.nan_count; continue; }
      if (n == "sens_gyro") { for (int i = 0; i < 3; ++i) out[3 * e + i] = d.sens_gyro[i]; continue; }
      if (n == "sens_vel") { for (int i = 0; i < 3; ++i) out[3 * e + i] = d.sens_vel[i]; continue; }
      if (n == "sens_quat") { for (int i = 0; i < 4; ++i) out[4 * e + i] = d.sens_quat[i]; continue; }
      if (n == "hm_cell") { for (size_t i = 0; i < d.hm_cell.size(); ++i) out[e * d.hm_cell.size() + i] = d.hm_cell[i]; continue; }
      if (n == "info") { out[4 * e] = d.info_rmse; out[4 * e + 1] = d.info_linx; out[4 * e + 2] = d.info_liny; out[4 * e + 3] = d.info_yaw; continue; }
      return -1;
    }
    return 0;
  }
  int set(const char* name_, const double* in) override {
    std::string n(name_);
    for (int e = 0; e < E.N; ++e) {
      auto& d = E.envs[e];
#define S(key, vec) if (n == key) { for (size_t i = 0; i < vec.size(); ++i) vec[i] = (T)in[(size_t)e * vec.size() + i]; continue; }
      S("qpos", d.qpos) S("qvel", d.qvel) S("ctrl", d.ctrl) S("qacc_warmstart", d.warm)
#undef S
      return -1;
    }
    return 0;
  }
  // per contact: dist, pos(3), normal(3), geom, cell, mu  -> 10 doubles
  int contacts(int env, double* out, int cap) override {
    auto& d = E.envs[env]; int n = 0;
    for (auto& c : d.con) {
      if (n >= cap) break;
      double* o = out + 10 * n++;
      o[0] = c.dist; for (int k = 0; k < 3; ++k) { o[1 + k] = c.pos[k]; o[4 + k] = c.frame[k]; }
      o[7] = c.geom; o[8] = c.cell; o[9] = c.mu;
    }
    return (int)d.con.size();
  }
  // per contact: force in the contact frame (normal, t1, t2, torsion, roll1, roll2), friction[5], condim -> 12 doubles [upstream mj_contactForce]
  int contact_forces(int env, double* out, int cap) override {
    auto& d = E.envs[env]; int n = 0;
    for (auto& c : d.con) {
      if (n >= cap) break;
      double* o = out + 12 * n++;
      for (int k = 0; k < 12; ++k) o[k] = 0;
      if (c.efc < 0) continue;
      const T* f = &d.efc_force[c.efc];
      if (c.dim == 1) o[0] = f[0];
      else if (!E.elliptic()) { for (int k = 0; k < c.dim - 1; ++k) { o[0] += f[2 * k] + f[2 * k + 1]; o[1 + k] = (f[2 * k] - f[2 * k + 1]) * c.fric[k]; } }
      else for (int k = 0; k < c.dim; ++k) o[k] = f[k];
      for (int k = 0; k < 5; ++k) o[6 + k] = c.fric[k];
      o[11] = c.dim;
    }
    return (int)d.con.size();
  }
};
}  // namespace

extern "C" {
void* orc_create(const void* blob, uint64_t nbytes, int num_envs, uint64_t seed, uint32_t env_offset, int use_float) {
  (void)nbytes;
  if (use_float) return new OracleT<float>(blob, num_envs, seed, env_offset);
  return new OracleT<double>(blob, num_envs, seed, env_offset);
}
void orc_destroy(void* h) { delete (IOracle*)h; }
void orc_reset(void* h, const uint8_t* mask, const double* command, float* state) { ((IOracle*)h)->reset(mask, command, state); }
void orc_step(void* h, const double* action, const double* command, float* state, uint8_t* term, uint8_t* trunc, int nthreads) {
  ((IOracle*)h)->step(action, command, state, term, trunc, nthreads);
}
int orc_get(void* h, const char* name, double* out) { return ((IOracle*)h)->get(name, out); }
int orc_set(void* h, const char* name, const double* in) { return ((IOracle*)h)->set(name, in); }
void orc_forward(void* h) { ((IOracle*)h)->forward_all(); }
void orc_substep(void* h) { ((IOracle*)h)->substep_all(); }
int orc_contacts(void* h, int env, double* out, int cap) { return ((IOracle*)h)->contacts(env, out, cap); }
int orc_contact_forces(void* h, int env, double* out, int cap) { return ((IOracle*)h)->contact_forces(env, out, cap); }
uint32_t orc_philox(uint64_t seed, uint32_t env, uint32_t stream, uint32_t step, uint32_t idx) { return orc::Philox::draw(seed, env, stream, step, idx); }
void orc_philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  orc::Philox::gen(out, c0, c1, c2, c3, (uint64_t)k0 | ((uint64_t)k1 << 32));
}
void orc_rne_post(void* h);
double orc_norm_ppf(double p) { return orc::norm_ppf(p); }
void orc_rne_post(void* h) { ((IOracle*)h)->rne_post_all(); }
double orc_ray_hfield(void* h, double x, double y) { return ((IOracle*)h)->ray_hfield(x, y); }
int orc_convex_pair(void* h, const double* geoms, double* out) { return ((IOracle*)h)->convex_pair(geoms, out); }
int orc_box_box(void* h, const double* geoms, double* out) { return ((IOracle*)h)->box_box(geoms, out); }
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}
