// engine_general.h -- the GENERAL constraint path of the engine: contact dimensions 1 / 4 / 6 (torsional and rolling friction),
// the elliptic friction cone, the PGS (dual) solver, impratio.
//
// The reference's four MJCF files all say condim 3, pyramidal cone, Newton, impratio 1
// (/root/reference/envs/flamingo_p_v3/assets/xml/flamingo_p_v3.xml:3,32 and the three siblings), and that case has its own
// specialised code in engine_core.h (four pyramid edges per contact, Jacobian-free rows, shared-memory records).  Everything else
// MuJoCo's <option cone= solver= impratio=> and geom condim allow runs here: ONE list of constraint rows with explicit Jacobians
// in the global-memory slot of the env (next to the contact-record overflow), lanes split the rows.  Restates the oracle's
// make_constraint / update_constraint / solve / solve_pgs (oracle/oracle.hpp) row for row; semantics per upstream
// mj_instantiateContact, mj_makeImpedance, mj_constraintUpdate, HessianCone, mj_solPGS (MuJoCo 3.2.7, un-vendored).
// Limits of this path: the first GEN_MAX_CON contacts of an env get rows (more are counted as dropped).
// Included by engine_core.h between the specialised solver and the forward stages.
#pragma once

enum { GT_EQ = 0, GT_FRICTION = 1, GT_LIMIT = 2, GT_CONTACT = 3, GT_ELLIPTIC = 4 };       // row types (oracle EFC_*)
enum { GS_QUAD = 0, GS_SATISFIED = 1, GS_LINNEG = 2, GS_LINPOS = 3, GS_CONE = 4 };
// per-contact block: friction[5] (slide, slide, torsion, roll, roll), regularised mu, first row, dim, cone Hessian (6 x 6)
enum { GC_FRIC = 0, GC_MU = 5, GC_FIRST = 6, GC_DIM = 7, GC_HC = 8, GC_STRIDE = GEN_CON_STRIDE };
struct GenView { float *J, *W, *D, *R, *aref, *X, *V, *F, *B, *floss; int *type, *id, *state; float* con; };
DEV GenView gen_view(const ModelDev& m, WSP ws) {
  float* g = *(float* const*)(WSF(ws) + m.off[W_GPTR]) + m.gen_off;
  const size_t NR = (size_t)m.gen_rows, nv = (size_t)MD(nv);
  GenView v;
  v.J = g; g += NR * nv; v.W = g; g += NR * nv;
  v.D = g; g += NR; v.R = g; g += NR; v.aref = g; g += NR; v.X = g; g += NR; v.V = g; g += NR; v.F = g; g += NR; v.B = g; g += NR; v.floss = g; g += NR;
  v.type = (int*)g; g += NR; v.id = (int*)g; g += NR; v.state = (int*)g; g += NR;
  v.con = g;
  return v;
}
DEV int gen_condim(const ModelDev& m) { const int c = MD(condim); return (c == 1 || c == 4 || c == 6) ? c : 3; }
DEV int gen_rows_per_contact(const ModelDev& m) { const int d = gen_condim(m); return d == 1 ? 1 : (MD(cone) == 1 ? d : 2 * (d - 1)); }
DEV float gen_impratio(const ModelDev& m) { const float r = MO(impratio); return r > 0.f ? r : 1.f; }
DEV float gen_dot(const float* a, const float* b, int n) { float s = 0.f; NOUNROLL for (int k = 0; k < n; ++k) s += a[k] * b[k]; return s; }

// rotational Jacobian column of dof k for `body` (world frame): the motion axis if k moves the body
DEV void jac_col_rot(const ModelDev& m, WSP ws, int body, int k, float* jr) {
  if ((TB(body_dofmask)[body] >> k) & 1) { const float* cd = WS(W_CDOF) + 6 * k; jr[0] = cd[0]; jr[1] = cd[1]; jr[2] = cd[2]; }
  else { jr[0] = jr[1] = jr[2] = 0.f; }
}

// rows of the env: equality, dof friction, limits (copied from what make_constraint() prepared), then the contacts.
// Leaves the row count in W_CNT[CNT_ROWS] and returns it.
DEV_NOINLINE int gen_make_rows(const ModelDev& m, WSP ws, int ncon, int lane) { LANE_REFRESH();
  const int nv = MD(nv), njnt = MD(njnt), neq = MD(neq), cdim = gen_condim(m), ell = MD(cone) == 1, rpc = gen_rows_per_contact(m);
  const GenView G = gen_view(m, ws);
  const int ncg = imin(ncon, GEN_MAX_CON);
  const float* scom = WS(W_SCOM);
  // ---- row directory (one lane: the rows must be numbered in the oracle's order)
  if (lane == 0) {
    int r = 0;
    for (int e = 0; e < 3 * neq; ++e) { G.type[r] = GT_EQ; G.id[r] = e; ++r; }
    for (int k = 0; k < nv; ++k) if (WS(W_FR_D)[k] > 0.f) { G.type[r] = GT_FRICTION; G.id[r] = k; ++r; }
    for (int j = 0; j < njnt; ++j) if (WS(W_LM_SIGN)[j] != 0.f) { G.type[r] = GT_LIMIT; G.id[r] = j; ++r; }
    for (int c = 0; c < ncg; ++c) {
      ((int*)(G.con + (size_t)c * GC_STRIDE))[GC_FIRST] = r; ((int*)(G.con + (size_t)c * GC_STRIDE))[GC_DIM] = cdim;
      for (int j = 0; j < rpc; ++j) { G.type[r] = (ell && cdim > 1) ? GT_ELLIPTIC : GT_CONTACT; G.id[r] = c | (j << 16); ++r; }
    }
    WSI(W_CNT)[CNT_ROWS] = r;
    if (ncon > ncg) WSI(W_CNT)[CNT_DROPPED] += ncon - ncg;
  }
  SYNC();
  const int nefc = WSI(W_CNT)[CNT_ROWS];
  // ---- per-contact friction coefficients: max over the two geoms (ground = W_SCAL / ground_friction), floor mjMINMU
  NOUNROLL for (int c = lane; c < ncg; c += LANES) {
    const float* rec = CREC(c); float* cc = G.con + (size_t)c * GC_STRIDE;
    const int g2 = ((const int*)rec)[CR_GEOM], cell = ((const int*)rec)[CR_CELL];
    float tor = TB(geom_fr_random)[g2] ? WS(W_SCAL)[4] : LDG(TB(geom_friction) + 3 * g2 + 1), rol = TB(geom_fr_random)[g2] ? WS(W_SCAL)[5] : LDG(TB(geom_friction) + 3 * g2 + 2);
    if (cell <= -2) { const int g1 = -2 - cell;
      tor = fmaxf(tor, TB(geom_fr_random)[g1] ? WS(W_SCAL)[4] : LDG(TB(geom_friction) + 3 * g1 + 1)); rol = fmaxf(rol, TB(geom_fr_random)[g1] ? WS(W_SCAL)[5] : LDG(TB(geom_friction) + 3 * g1 + 2));
    } else { tor = fmaxf(tor, m.ground_friction[3] != 0.f ? WS(W_SCAL)[4] : m.ground_friction[1]); rol = fmaxf(rol, m.ground_friction[3] != 0.f ? WS(W_SCAL)[5] : m.ground_friction[2]); }
    const float sl = fmaxf(rec[CR_MU], 1e-5f);
    cc[GC_FRIC] = cc[GC_FRIC + 1] = sl; cc[GC_FRIC + 2] = fmaxf(tor, 1e-5f); cc[GC_FRIC + 3] = cc[GC_FRIC + 4] = fmaxf(rol, 1e-5f);
    cc[GC_MU] = sl / sqrtf(gen_impratio(m));
  }
  SYNC();
  // ---- Jacobian rows
  NOUNROLL for (int idx = lane; idx < nefc * nv; idx += LANES) {
    const int r = idx / nv, k = idx - r * nv, type = G.type[r], id = G.id[r];
    float v = 0.f;
    if (type == GT_EQ) v = WS(W_EQ_J)[(size_t)id * nv + k];
    else if (type == GT_FRICTION) v = (k == id) ? 1.f : 0.f;
    else if (type == GT_LIMIT) v = (k == TB(jnt_dofadr)[id]) ? WS(W_LM_SIGN)[id] : 0.f;
    else {
      const int c = id & 0xffff, sub = id >> 16;
      const float* rec = CREC(c); const float* cc = G.con + (size_t)c * GC_STRIDE; const float* fr = rec + CR_FRAME;
      const int b2 = ((const int*)rec)[CR_BODY], cell = ((const int*)rec)[CR_CELL], b1 = cell <= -2 ? TB(geom_body)[-2 - cell] : 0;
      float off[3], jp[3], jr[3]; v3sub(off, rec + CR_POS, scom);
      jac_col(m, ws, b2, k, off, jp); jac_col_rot(m, ws, b2, k, jr);
      if (b1 > 0) { float t[3]; jac_col(m, ws, b1, k, off, t); v3sub(jp, jp, t); jac_col_rot(m, ws, b1, k, t); v3sub(jr, jr, t); }
      // frame rows: 0..2 = (normal, t1, t2) . jp, 3..5 = (normal, t1, t2) . jr
      const float jn = v3dot(fr, jp);
      if (type == GT_ELLIPTIC) v = sub == 0 ? jn : (sub < 3 ? v3dot(fr + 3 * sub, jp) : v3dot(fr + 3 * (sub - 3), jr));
      else if (cdim == 1) v = jn;
      else { const int t = sub >> 1; const float sg = (sub & 1) ? -1.f : 1.f;
        const float jt = t < 2 ? v3dot(fr + 3 * (1 + t), jp) : v3dot(fr + 3 * (t - 2), jr);
        v = jn + sg * cc[GC_FRIC + t] * jt; }
    }
    G.J[(size_t)r * nv + k] = v;
  }
  SYNC();
  // ---- D, R, aref per row
  const float solref[2] = {MO(solref0), MO(solref1)};
  const float solimp[5] = {MO(solimp0), MO(solimp1), MO(solimp2), MO(solimp3), MO(solimp4)};
  float K, B; kb_params(m, solref, solimp, &K, &B);
  const float ipr = gen_impratio(m);
  NOUNROLL for (int r = lane; r < nefc; r += LANES) {
    const int type = G.type[r], id = G.id[r];
    float D = 0.f, aref = 0.f, fl = 0.f;
    if (type == GT_EQ) { D = WS(W_EQ_D)[id]; aref = WS(W_EQ_AREF)[id]; }
    else if (type == GT_FRICTION) { D = WS(W_FR_D)[id]; aref = WS(W_FR_AREF)[id]; fl = WS(W_FLOSS)[id]; }
    else if (type == GT_LIMIT) { D = WS(W_LM_D)[id]; aref = WS(W_LM_AREF)[id]; }
    else {
      const int c = id & 0xffff, sub = id >> 16;
      const float* rec = CREC(c); const float* cc = G.con + (size_t)c * GC_STRIDE;
      const float vel = gen_dot(G.J + (size_t)r * nv, WS(W_QVEL), nv), dist = rec[CR_DIST], imp = impedance(solimp, dist);
      float tran = WS(W_INVWB)[((const int*)rec)[CR_BODY]];
      { const int cell = ((const int*)rec)[CR_CELL]; if (cell <= -2) { const int b1 = TB(geom_body)[-2 - cell]; if (b1 > 0) tran += WS(W_INVWB)[b1]; } }
      const float f0 = cc[GC_FRIC], mu = cc[GC_MU];
      float R;
      if (type == GT_CONTACT) {
        aref = -B * vel - K * imp * dist;
        if (cdim == 1) R = fmaxf(MINVALF, (1.f - imp) * tran / imp);
        else R = 2.f * mu * mu * fmaxf(MINVALF, (1.f - imp) * (tran + f0 * f0 * tran) / imp);
      } else {
        const float R0 = fmaxf(MINVALF, (1.f - imp) * tran / imp), R1 = R0 / ipr;
        if (sub == 0) { aref = -B * vel - K * imp * dist; R = R0; }
        else { aref = -B * vel; R = sub == 1 ? R1 : R1 * f0 * f0 / (cc[GC_FRIC + sub - 1] * cc[GC_FRIC + sub - 1]); }
      }
      D = 1.f / R;
    }
    G.D[r] = D; G.R[r] = 1.f / D; G.aref[r] = aref; G.floss[r] = fl;
  }
  SYNC();
  return nefc;
}

// elliptic cone of one contact at the row residuals x[0..dim): zone 0 top (no force), 1 bottom (all rows quadratic), 2 middle.
// U = (x[0] mu, x[j] friction[j-1]), N = U[0], T = |U[1..]|, Dm = D[0] / (mu^2 (1 + mu^2))  (oracle cone_eval)
struct GenCone { int zone; float N, T, Dm, U[6]; };
DEV GenCone gen_cone(const float* cc, float D0, const float* x, int dim) {
  GenCone e; const float mu = cc[GC_MU];
  e.U[0] = x[0] * mu; float tt = 0.f;
  for (int j = 1; j < dim; ++j) { e.U[j] = x[j] * cc[GC_FRIC + j - 1]; tt += e.U[j] * e.U[j]; }
  e.N = e.U[0]; e.T = sqrtf(tt); e.Dm = D0 / (mu * mu * (1.f + mu * mu));
  if (e.N >= mu * e.T || (e.T <= 0.f && e.N >= 0.f)) e.zone = 0;
  else if (mu * e.N + e.T <= 0.f || (e.T <= 0.f && e.N < 0.f)) e.zone = 1;
  else e.zone = 2;
  return e;
}
// X = J q - aref
DEV_NOINLINE void gen_jaref(const ModelDev& m, WSP ws, int nefc, const float* q, int lane) { LANE_REFRESH();
  const GenView G = gen_view(m, ws); const int nv = MD(nv);
  NOUNROLL for (int r = lane; r < nefc; r += LANES) G.X[r] = gen_dot(G.J + (size_t)r * nv, q, nv) - G.aref[r];
  SYNC();
}
// forces and states from X; qfrc_constraint -> W_FCON; returns the constraint cost (oracle update_constraint)
DEV_NOINLINE float gen_update(const ModelDev& m, WSP ws, int nefc, int lane) { LANE_REFRESH();
  const GenView G = gen_view(m, ws); const int nv = MD(nv);
  float cost = 0.f;
  NOUNROLL for (int r = lane; r < nefc; r += LANES) {
    const int type = G.type[r];
    const float x = G.X[r], D = G.D[r];
    if (type == GT_ELLIPTIC) {
      if ((G.id[r] >> 16) != 0) continue;          // the lane of the cone's first row handles all its rows
      const float* cc = G.con + (size_t)(G.id[r] & 0xffff) * GC_STRIDE; const int dim = ((const int*)cc)[GC_DIM];
      const GenCone e = gen_cone(cc, D, G.X + r, dim);
      for (int j = 0; j < dim; ++j) {
        const float xj = G.X[r + j], Dj = G.D[r + j];
        if (e.zone == 0) { G.state[r + j] = GS_SATISFIED; G.F[r + j] = 0.f; }
        else if (e.zone == 1) { G.state[r + j] = GS_QUAD; G.F[r + j] = -Dj * xj; cost += 0.5f * Dj * xj * xj; }
        else {
          const float NmT = e.N - cc[GC_MU] * e.T;
          G.state[r + j] = GS_CONE;
          G.F[r + j] = j == 0 ? -e.Dm * NmT * cc[GC_MU] : e.Dm * NmT * cc[GC_MU] * e.U[j] * cc[GC_FRIC + j - 1] / e.T;
          if (j == 0) cost += 0.5f * e.Dm * NmT * NmT;
        }
      }
      continue;
    }
    int st = GS_QUAD; float force = -D * x, c = 0.5f * D * x * x;
    if (type == GT_FRICTION) {
      const float f = G.floss[r], Rf = G.R[r] * f;
      if (x <= -Rf) { st = GS_LINNEG; force = f; c = -0.5f * Rf * f - f * x; }
      else if (x >= Rf) { st = GS_LINPOS; force = -f; c = -0.5f * Rf * f + f * x; }
    } else if (type != GT_EQ) { if (x >= 0.f) { st = GS_SATISFIED; force = 0.f; c = 0.f; } }
    G.state[r] = st; G.F[r] = force; cost += c;
  }
  SYNC();
  FOR_LANE(k, nv) { float s = 0.f; NOUNROLL for (int r = 0; r < nefc; ++r) s += G.J[(size_t)r * nv + k] * G.F[r]; WS(W_FCON)[k] = s; }
  SYNC();
  return wsum(cost);
}
// rows at X + a V for the line search: cost and its first / second derivative along the search direction (oracle ls_eval)
DEV_NOINLINE RowSum gen_eval_rows(const ModelDev& m, WSP ws, int nefc, float a, int lane) { LANE_REFRESH();
  const GenView G = gen_view(m, ws);
  RowSum s = {0.f, 0.f, 0.f};
  NOUNROLL for (int r = lane; r < nefc; r += LANES) {
    const int type = G.type[r];
    const float jv = G.V[r], D = G.D[r], x = G.X[r] + a * jv;
    if (type == GT_ELLIPTIC) {
      if ((G.id[r] >> 16) != 0) continue;
      const float* cc = G.con + (size_t)(G.id[r] & 0xffff) * GC_STRIDE; const int dim = ((const int*)cc)[GC_DIM];
      float xa[6]; for (int j = 0; j < dim; ++j) xa[j] = G.X[r + j] + a * G.V[r + j];
      const GenCone e = gen_cone(cc, D, xa, dim);
      if (e.zone == 1) { for (int j = 0; j < dim; ++j) { const float Dj = G.D[r + j], vj = G.V[r + j]; s.cost += 0.5f * Dj * xa[j] * xa[j]; s.d0 += Dj * xa[j] * vj; s.d1 += Dj * vj * vj; } }
      else if (e.zone == 2) {
        const float mu = cc[GC_MU], V0 = G.V[r] * mu; float uv = 0.f, vv = 0.f;
        for (int j = 1; j < dim; ++j) { const float Vj = G.V[r + j] * cc[GC_FRIC + j - 1]; uv += e.U[j] * Vj; vv += Vj * Vj; }
        const float T1 = uv / e.T, T2 = (vv - T1 * T1) / e.T, NmT = e.N - mu * e.T, s1 = V0 - mu * T1;
        s.cost += 0.5f * e.Dm * NmT * NmT; s.d0 += e.Dm * NmT * s1; s.d1 += e.Dm * (s1 * s1 - NmT * mu * T2);
      }
      continue;
    }
    if (type == GT_FRICTION) {
      const float f = G.floss[r], Rf = G.R[r] * f;
      if (x <= -Rf) { s.cost += f * (-0.5f * Rf - x); s.d0 -= f * jv; continue; }
      if (x >= Rf) { s.cost += f * (-0.5f * Rf + x); s.d0 += f * jv; continue; }
    } else if (type != GT_EQ) { if (x >= 0.f) continue; }
    s.cost += 0.5f * D * x * x; s.d0 += D * x * jv; s.d1 += D * jv * jv;
  }
  return s;
}
// total cost at q (Gauss + constraints); leaves X / forces / W_FCON for q and M q in Mq (oracle solve(): mulM + jaref + update)
DEV_NOINLINE float gen_total_cost(const ModelDev& m, WSP ws, int nefc, const float* q, float* Mq, int lane) { LANE_REFRESH();
  const int nv = MD(nv);
  mat_vec(WS(W_M), q, Mq, nv, lane);
  if (q != WS(W_QACC)) { FOR_LANE(k, nv) WS(W_QACC)[k] = q[k]; SYNC(); }
  gen_jaref(m, ws, nefc, q, lane);
  const float c = gen_update(m, ws, nefc, lane);
  float g = 0.f;
  FOR_LANE(k, nv) g += (Mq[k] - WS(W_FSMOOTH)[k]) * (q[k] - WS(W_ASMOOTH)[k]);
  return c + 0.5f * wsum(g);
}
// grad = Ma - qfrc_smooth - qfrc_constraint; H = M + sum_quad D J'J + sum_cones Jc' Hc Jc; search = -H^-1 grad.  Returns |grad|
DEV_NOINLINE float gen_direction(const ModelDev& m, WSP ws, int nefc, int ncg, int lane, float* termnorm) { LANE_REFRESH();
  const GenView G = gen_view(m, ws); const int nv = MD(nv);
  float* grad = WS(W_GRAD); float* H = WS(W_A); const float* M = WS(W_M); const float* Ma = WS(W_MA);
  float gn = 0.f, fn = 0.f;
  FOR_LANE(k, nv) { const float a = Ma[k], b = WS(W_FSMOOTH)[k], c = WS(W_FCON)[k], g = a - b - c; grad[k] = g; gn += g * g; fn += a * a + b * b + c * c; WS(W_TMPV)[k] = -g; }
  gn = sqrtf(wsum(gn));
  if (termnorm) *termnorm = sqrtf(wsum(fn));
  // cone Hessians of the middle-zone contacts: Hc = Dm [v v' - mu (N - mu T) G], v = (mu, -mu U_j f_j / T), G = f_j f_k (delta_jk / T - U_j U_k / T^3)
  NOUNROLL for (int c = lane; c < ncg; c += LANES) {
    float* cc = G.con + (size_t)c * GC_STRIDE; const int r = ((const int*)cc)[GC_FIRST], dim = ((const int*)cc)[GC_DIM];
    if (G.type[r] != GT_ELLIPTIC || G.state[r] != GS_CONE) continue;
    const GenCone e = gen_cone(cc, G.D[r], G.X + r, dim);
    const float mu = cc[GC_MU], NmT = e.N - mu * e.T; float v[6];
    v[0] = mu; for (int j = 1; j < dim; ++j) v[j] = -mu * e.U[j] * cc[GC_FRIC + j - 1] / e.T;
    for (int a = 0; a < dim; ++a) for (int b = 0; b < dim; ++b) {
      float g = 0.f;
      if (a > 0 && b > 0) g = cc[GC_FRIC + a - 1] * cc[GC_FRIC + b - 1] * ((a == b ? 1.f / e.T : 0.f) - e.U[a] * e.U[b] / (e.T * e.T * e.T));
      cc[GC_HC + a * dim + b] = e.Dm * (v[a] * v[b] - mu * NmT * g);
    }
  }
  SYNC();
  const int npair = (nv * (nv + 1)) >> 1;
  NOUNROLL for (int idx = lane; idx < npair; idx += LANES) {          // lower triangle, one (i, j <= i) pair per lane
    const int t = TB(tri)[idx], i = t >> 8, j = t & 255;
    float h = M[i * nv + j];
    NOUNROLL for (int r = 0; r < nefc; ++r) {
      const int st = G.state[r];
      if (st == GS_QUAD) { const float ji = G.J[(size_t)r * nv + i]; if (ji != 0.f) h += G.D[r] * ji * G.J[(size_t)r * nv + j]; }
      else if (st == GS_CONE && (G.id[r] >> 16) == 0) {
        const float* cc = G.con + (size_t)(G.id[r] & 0xffff) * GC_STRIDE; const int dim = ((const int*)cc)[GC_DIM];
        for (int a = 0; a < dim; ++a) { const float ja = G.J[(size_t)(r + a) * nv + i]; if (ja == 0.f) continue;
          for (int b = 0; b < dim; ++b) h += ja * cc[GC_HC + a * dim + b] * G.J[(size_t)(r + b) * nv + j]; }
      }
    }
    H[i * nv + j] = h;
  }
  SYNC();
  chol_factor(m, H, WS(W_INVD), nv, lane, 0);      // general rows: dense variant
  chol_solve(H, WS(W_INVD), WS(W_TMPV), WS(W_BUF), WS(W_SEARCH), nv, lane);
  return gn;
}
// contact forces in the contact frame -> the records (CR_F: normal, t1, t2; CR_WF: torsion, roll1, roll2) for cfrc_ext; connect forces -> W_EQ_F
DEV_NOINLINE void gen_publish_forces(const ModelDev& m, WSP ws, int nefc, int ncon, int lane) { LANE_REFRESH();
  const GenView G = gen_view(m, ws); const int ncg = imin(ncon, GEN_MAX_CON), ell = MD(cone) == 1;
  NOUNROLL for (int c = lane; c < ncon; c += LANES) {
    float* rec = CREC(c); float lf[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < ncg) {
      const float* cc = G.con + (size_t)c * GC_STRIDE; const int r = ((const int*)cc)[GC_FIRST], dim = ((const int*)cc)[GC_DIM];
      const float* f = G.F + r;
      if (dim == 1) lf[0] = f[0];
      else if (!ell) { for (int k = 0; k < dim - 1; ++k) { lf[0] += f[2 * k] + f[2 * k + 1]; lf[1 + k] = (f[2 * k] - f[2 * k + 1]) * cc[GC_FRIC + k]; } }
      else for (int k = 0; k < dim; ++k) lf[k] = f[k];
    }
    for (int k = 0; k < 3; ++k) { rec[CR_F + k] = lf[k]; rec[CR_WF + k] = lf[3 + k]; }
  }
  NOUNROLL for (int r = lane; r < nefc; r += LANES) if (G.type[r] == GT_EQ) WS(W_EQ_F)[G.id[r]] = G.F[r];
  SYNC();
}

// Newton on the primal cost with the general rows; same control flow as newton_solve() / the oracle's solve()
DEV_NOINLINE int gen_newton(const ModelDev& m, WSP ws, int nefc, int ncon, int lane) { LANE_REFRESH();
  const GenView G = gen_view(m, ws); const int nv = MD(nv), ncg = imin(ncon, GEN_MAX_CON);
  float* qacc = WS(W_QACC); float* Ma = WS(W_MA); float* Mv = WS(W_MV); float* search = WS(W_SEARCH); const float* M = WS(W_M);
  const float cs = gen_total_cost(m, ws, nefc, WS(W_ASMOOTH), Ma, lane);
  float cost = gen_total_cost(m, ws, nefc, WS(W_WARM), Ma, lane);
  if (cost > cs) cost = gen_total_cost(m, ws, nefc, WS(W_ASMOOTH), Ma, lane);
  const float scale = 1.f / (WS(W_SCAL)[1] * (float)imax(1, nv)), tol = MO(tolerance);
  const int maxiter = MD(iterations);
  (void)gen_direction(m, ws, nefc, ncg, lane, nullptr);
  int iter = 0;
  while (iter < maxiter) {
    mat_vec(M, search, Mv, nv, lane);
    NOUNROLL for (int r = lane; r < nefc; r += LANES) G.V[r] = gen_dot(G.J + (size_t)r * nv, search, nv);
    SYNC();
    float q1 = 0.f, q2 = 0.f, sn = 0.f, gauss = 0.f, absterms = 0.f;
    FOR_LANE(k, nv) { const float r = Ma[k] - WS(W_FSMOOTH)[k]; q1 += search[k] * r; q2 += 0.5f * search[k] * Mv[k]; sn += search[k] * search[k];
      gauss += r * (qacc[k] - WS(W_ASMOOTH)[k]); absterms += fabsf(search[k] * r); }
    NOUNROLL for (int r = lane; r < nefc; r += LANES) absterms += fabsf(G.F[r] * G.V[r]);
    q1 = wsum(q1); q2 = wsum(q2); sn = sqrtf(wsum(sn)); gauss = 0.5f * wsum(gauss); absterms = wsum(absterms);
    const float alpha = linesearch(m, ws, nefc, gauss, q1, q2, sn, absterms, lane);      // ls_eval() dispatches to gen_eval_rows()
    if (alpha == 0.f) break;
    FOR_LANE(k, nv) { qacc[k] += alpha * search[k]; Ma[k] += alpha * Mv[k]; }
    NOUNROLL for (int r = lane; r < nefc; r += LANES) G.X[r] += alpha * G.V[r];
    SYNC();
    const float old = cost;
    float g = 0.f;
    FOR_LANE(k, nv) g += (Ma[k] - WS(W_FSMOOTH)[k]) * (qacc[k] - WS(W_ASMOOTH)[k]);
    cost = gen_update(m, ws, nefc, lane) + 0.5f * wsum(g);
    float fn;
    const float gn = gen_direction(m, ws, nefc, ncg, lane, &fn);
    ++iter;
    if ((old - cost) < fmaxf(tol / scale, COST_EPS * fabsf(old)) || gn < fmaxf(tol / scale, GRAD_EPS * fn)) break;      // fp32 floors as in newton_solve()
  }
  FOR_LANE(k, nv) WS(W_WARM)[k] = qacc[k];
  SYNC();
  return iter;
}

// min 1/2 x'Ax + x'b  s.t.  sum (x_j / fri_j)^2 <= r^2, n <= 5: Newton on the multiplier of the scaled problem (oracle qcqp)
DEV void gen_qcqp(const float* A, const float* b, const float* fri, float r, int n, float* x) {
  float As[25], bs[5], y[5], z[5], Lc[25];
  for (int i = 0; i < n; ++i) { bs[i] = b[i] * fri[i]; for (int j = 0; j < n; ++j) As[i * n + j] = A[i * n + j] * fri[i] * fri[j]; }
  float lam = 0.f;
  for (int it = 0; it <= 20; ++it) {
    for (int i = 0; i < n * n; ++i) Lc[i] = As[i];
    for (int i = 0; i < n; ++i) Lc[i * n + i] += lam;
    for (int j = 0; j < n; ++j) {
      float dd = Lc[j * n + j]; for (int k = 0; k < j; ++k) dd -= Lc[j * n + k] * Lc[j * n + k];
      dd = sqrtf(fmaxf(dd, MINVALF)); Lc[j * n + j] = dd;
      for (int i = j + 1; i < n; ++i) { float v = Lc[i * n + j]; for (int k = 0; k < j; ++k) v -= Lc[i * n + k] * Lc[j * n + k]; Lc[i * n + j] = v / dd; }
    }
    for (int i = 0; i < n; ++i) { float v = -bs[i]; for (int k = 0; k < i; ++k) v -= Lc[i * n + k] * y[k]; y[i] = v / Lc[i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { float v = y[i]; for (int k = i + 1; k < n; ++k) v -= Lc[k * n + i] * y[k]; y[i] = v / Lc[i * n + i]; }
    float yy = 0.f; for (int i = 0; i < n; ++i) yy += y[i] * y[i];
    if (yy <= r * r || it == 20) { if (yy > r * r && yy > 0.f) { const float sc = r / sqrtf(yy); for (int i = 0; i < n; ++i) y[i] *= sc; } break; }
    // z = (As + lam I)^-1 y;  phi = |y|^2 - r^2, phi' = -2 y'z
    for (int i = 0; i < n; ++i) { float v = y[i]; for (int k = 0; k < i; ++k) v -= Lc[i * n + k] * z[k]; z[i] = v / Lc[i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { float v = z[i]; for (int k = i + 1; k < n; ++k) v -= Lc[k * n + i] * z[k]; z[i] = v / Lc[i * n + i]; }
    float yz = 0.f; for (int i = 0; i < n; ++i) yz += y[i] * z[i];
    const float phi = yy - r * r, dphi = -2.f * yz;
    if (phi < 1e-6f * fmaxf(1.f, r * r) || dphi > -MINVALF) { const float sc = r / sqrtf(yy); for (int i = 0; i < n; ++i) y[i] *= sc; break; }
    lam = fmaxf(0.f, lam - phi / dphi);
  }
  for (int i = 0; i < n; ++i) x[i] = y[i] * fri[i];
}
// PGS on the dual (oracle solve_pgs): W = M^-1 J' row by row, running v = W f in W_SEARCH, Gauss-Seidel over the rows
DEV_NOINLINE int gen_pgs(const ModelDev& m, WSP ws, int nefc, int ncon, int lane) { LANE_REFRESH();
#ifdef COSIM_AB_NO_PGS
  return 0;
#endif
  const GenView G = gen_view(m, ws); const int nv = MD(nv);
  float* v = WS(W_SEARCH); float* Ma = WS(W_MA);
  // W rows and b; AR_ii kept in G.V.  W_A / W_INVD still hold the factor of M (stage 1): this solver builds no Hessian
  NOUNROLL for (int r = 0; r < nefc; ++r) {
    FOR_LANE(k, nv) WS(W_TMPV)[k] = G.J[(size_t)r * nv + k];
    SYNC();
    chol_solve(WS(W_A), WS(W_INVD), WS(W_TMPV), WS(W_BUF), WS(W_GRAD), nv, lane);
    FOR_LANE(k, nv) G.W[(size_t)r * nv + k] = WS(W_GRAD)[k];
    SYNC();
  }
  NOUNROLL for (int r = lane; r < nefc; r += LANES) {
    G.B[r] = gen_dot(G.J + (size_t)r * nv, WS(W_ASMOOTH), nv) - G.aref[r];
    G.V[r] = gen_dot(G.J + (size_t)r * nv, G.W + (size_t)r * nv, nv) + G.R[r];
  }
  SYNC();
  // warm start: forces of the primal update at qacc_warmstart, kept if their dual cost beats f = 0
  mat_vec(WS(W_M), WS(W_WARM), Ma, nv, lane);
  gen_jaref(m, ws, nefc, WS(W_WARM), lane);
  (void)gen_update(m, ws, nefc, lane);
  FOR_LANE(k, nv) { float s = 0.f; NOUNROLL for (int r = 0; r < nefc; ++r) s += G.W[(size_t)r * nv + k] * G.F[r]; v[k] = s; }
  SYNC();
  { float c = 0.f;
    NOUNROLL for (int r = lane; r < nefc; r += LANES) c += G.F[r] * (0.5f * (gen_dot(G.J + (size_t)r * nv, v, nv) + G.R[r] * G.F[r]) + G.B[r]);
    c = wsum(c);
    if (c > 0.f) { NOUNROLL for (int r = lane; r < nefc; r += LANES) G.F[r] = 0.f; FOR_LANE(k, nv) v[k] = 0.f; }
    SYNC(); }
  const float scale = 1.f / (WS(W_SCAL)[1] * (float)imax(1, nv)), tol = MO(tolerance);
  const int maxiter = MD(iterations);
  int iter = 0;
  while (iter < maxiter) {
    float improvement = 0.f;
    NOUNROLL for (int i = 0; i < nefc; ++i) {          // every lane runs the same scalar update (warp-uniform data); reads of v are plain loads
      const int type = G.type[i];
      const float* cc = type == GT_ELLIPTIC ? G.con + (size_t)(G.id[i] & 0xffff) * GC_STRIDE : nullptr;
      const int dm = type == GT_ELLIPTIC ? ((const int*)cc)[GC_DIM] : 1;
      float res[6], oldf[6], f[6], Ac[36];
      for (int j = 0; j < dm; ++j) { oldf[j] = G.F[i + j]; f[j] = oldf[j]; res[j] = gen_dot(G.J + (size_t)(i + j) * nv, v, nv) + G.R[i + j] * oldf[j] + G.B[i + j]; }
      if (dm == 1) {
        Ac[0] = G.V[i];
        float x = oldf[0] - res[0] / Ac[0];
        if (type == GT_FRICTION) x = fminf(G.floss[i], fmaxf(-G.floss[i], x));
        else if (type != GT_EQ) x = fmaxf(0.f, x);
        f[0] = x;
      } else {
        for (int a = 0; a < dm; ++a) for (int c2 = 0; c2 < dm; ++c2) Ac[a * dm + c2] = gen_dot(G.J + (size_t)(i + a) * nv, G.W + (size_t)(i + c2) * nv, nv) + (a == c2 ? G.R[i + a] : 0.f);
        if (f[0] < MINVALF) {          // apex: leave along the cone generator opposite to the scaled tangential residual, if the residual is outside the dual cone
          float tt = 0.f; for (int j = 1; j < dm; ++j) tt += cc[GC_FRIC + j - 1] * cc[GC_FRIC + j - 1] * res[j] * res[j];
          tt = sqrtf(tt);
          for (int j = 0; j < dm; ++j) f[j] = 0.f;
          if (res[0] < tt) {
            float g[6], den = 0.f; g[0] = 1.f;
            for (int j = 1; j < dm; ++j) g[j] = tt > 0.f ? -cc[GC_FRIC + j - 1] * cc[GC_FRIC + j - 1] * res[j] / tt : 0.f;
            for (int a = 0; a < dm; ++a) for (int c2 = 0; c2 < dm; ++c2) den += g[a] * Ac[a * dm + c2] * g[c2];
            if (den >= MINVALF) { const float x = -(res[0] - tt) / den; for (int j = 0; j < dm; ++j) f[j] = x * g[j]; }
          }
        } else {                       // step along the current force ray
          float den = 0.f, num = 0.f;
          for (int a = 0; a < dm; ++a) { float av = 0.f; for (int c2 = 0; c2 < dm; ++c2) av += Ac[a * dm + c2] * oldf[c2]; den += oldf[a] * av; num += oldf[a] * res[a]; }
          if (den >= MINVALF) { float x = -num / den; if (x < -1.f) x = -1.f; for (int j = 0; j < dm; ++j) f[j] += x * oldf[j]; }
        }
        if (f[0] < MINVALF) { for (int j = 1; j < dm; ++j) f[j] = 0.f; }
        else {                         // friction forces on the ellipse with the normal force fixed
          float At[25], bt[5], xs[5];
          for (int a = 1; a < dm; ++a) {
            float r0 = res[a] + Ac[a * dm] * (f[0] - oldf[0]);
            for (int c2 = 1; c2 < dm; ++c2) { At[(a - 1) * (dm - 1) + (c2 - 1)] = Ac[a * dm + c2]; r0 -= Ac[a * dm + c2] * oldf[c2]; }
            bt[a - 1] = r0;
          }
          gen_qcqp(At, bt, cc + GC_FRIC, f[0], dm - 1, xs);
          for (int j = 1; j < dm; ++j) f[j] = xs[j - 1];
        }
      }
      float change = 0.f;
      SYNC();
      for (int a = 0; a < dm; ++a) {
        const float da = f[a] - oldf[a];
        if (da == 0.f) continue;
        change += da * res[a];
        for (int c2 = 0; c2 < dm; ++c2) change += 0.5f * da * (dm == 1 ? Ac[0] : Ac[a * dm + c2]) * (f[c2] - oldf[c2]);
        FOR_LANE(k, nv) v[k] += G.W[(size_t)(i + a) * nv + k] * da;
        if (lane == 0) G.F[i + a] = f[a];
      }
      SYNC();
      improvement -= change;
      i += dm - 1;
    }
    ++iter;
    if (improvement * scale < tol) break;
  }
  FOR_LANE(k, nv) { float s = 0.f; NOUNROLL for (int r = 0; r < nefc; ++r) s += G.J[(size_t)r * nv + k] * G.F[r]; WS(W_FCON)[k] = s;
    const float a = WS(W_ASMOOTH)[k] + v[k]; WS(W_QACC)[k] = a; WS(W_WARM)[k] = a; }
  SYNC();
  gen_jaref(m, ws, nefc, WS(W_QACC), lane);
  return iter;
}

// constraint solve of the general path (stage 4); returns the solver iterations
DEV_NOINLINE int gen_solve(const ModelDev& m, WSP ws, int ncon, int lane) { LANE_REFRESH();
  const int nv = MD(nv), nefc = WSI(W_CNT)[CNT_ROWS];
  if (nefc == 0) {
    FOR_LANE(k, nv) { WS(W_QACC)[k] = WS(W_ASMOOTH)[k]; WS(W_WARM)[k] = WS(W_ASMOOTH)[k]; WS(W_FCON)[k] = 0.f; }
    SYNC(); return 0;
  }
  const int iters = MD(solver) == 1 ? gen_pgs(m, ws, nefc, ncon, lane) : gen_newton(m, ws, nefc, ncon, lane);
  gen_publish_forces(m, ws, nefc, ncon, lane);
  return iters;
}
