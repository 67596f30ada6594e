// oracle_math.hpp -- small fixed-size math + counter-based RNG used by the CPU oracle.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Conventions follow MuJoCo 3.2.7 (un-vendored dependency of the reference,
// /root/reference/requirements.txt:19): quaternions are (w,x,y,z); rotation matrices are
// row-major 3x3; spatial vectors are [angular(3); linear(3)].
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

template <class T> inline void v3set(T* r, T a, T b, T c) { r[0] = a; r[1] = b; r[2] = c; }
template <class T> inline void v3copy(T* r, const T* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
template <class T> inline void v3zero(T* r) { r[0] = r[1] = r[2] = T(0); }
template <class T> inline void v3add(T* r, const T* a, const T* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
template <class T> inline void v3sub(T* r, const T* a, const T* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
template <class T> inline void v3scl(T* r, const T* a, T s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
template <class T> inline void v3addscl(T* r, const T* a, const T* b, T s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
template <class T> inline T v3dot(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> inline void v3cross(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> inline T v3norm(const T* a) { return std::sqrt(v3dot(a, a)); }
template <class T> inline T v3normalize(T* a) {
  T n = v3norm(a);
  if (n < T(1e-15)) { a[0] = T(1); a[1] = a[2] = T(0); return n; }
  T s = T(1) / n; a[0] *= s; a[1] *= s; a[2] *= s; return n;
}
// r = M v, r = M^T v   (M row-major 3x3)
template <class T> inline void m3mulv(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> inline void m3tmulv(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> inline void m3mul(T* r, const T* a, const T* b) {
  T t[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  for (int i = 0; i < 9; ++i) r[i] = t[i];
}
template <class T> inline void quat_mul(T* r, const T* a, const T* b) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <class T> inline void quat_normalize(T* q) {
  T n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < T(1e-15)) { q[0] = T(1); q[1] = q[2] = q[3] = T(0); return; }
  T s = T(1) / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
template <class T> inline void quat_to_mat(T* m, const T* q) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = T(1) - T(2) * (y * y + z * z); m[1] = T(2) * (x * y - z * w); m[2] = T(2) * (x * z + y * w);
  m[3] = T(2) * (x * y + z * w); m[4] = T(1) - T(2) * (x * x + z * z); m[5] = T(2) * (y * z - x * w);
  m[6] = T(2) * (x * z - y * w); m[7] = T(2) * (y * z + x * w); m[8] = T(1) - T(2) * (x * x + y * y);
}
template <class T> inline void quat_rot(T* r, const T* q, const T* v) { T m[9]; quat_to_mat(m, q); m3mulv(r, m, v); }
template <class T> inline void axis_angle_quat(T* q, const T* axis, T angle) {
  T s = std::sin(angle * T(0.5));
  q[0] = std::cos(angle * T(0.5)); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

// ---- spatial algebra on [ang; lin] 6-vectors and 10-number inertias [Ixx,Iyy,Izz,Ixy,Ixz,Iyz, mh(3), m]
template <class T> inline void inert_mul(T* r, const T* I, const T* v) {
  const T* w = v; const T* l = v + 3; const T* mh = I + 6; T m = I[9];
  T Iw[3] = {I[0] * w[0] + I[3] * w[1] + I[4] * w[2], I[3] * w[0] + I[1] * w[1] + I[5] * w[2], I[4] * w[0] + I[5] * w[1] + I[2] * w[2]};
  T c1[3], c2[3];
  v3cross(c1, mh, l);   // (m h) x l
  v3cross(c2, mh, w);   // (m h) x w
  r[0] = Iw[0] + c1[0]; r[1] = Iw[1] + c1[1]; r[2] = Iw[2] + c1[2];
  r[3] = m * l[0] - c2[0]; r[4] = m * l[1] - c2[1]; r[5] = m * l[2] - c2[2];
}
template <class T> inline void cross_motion(T* r, const T* v, const T* s) {   // v x s (motion)
  T a[3], b[3], c[3];
  v3cross(a, v, s); v3cross(b, v, s + 3); v3cross(c, v + 3, s);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
template <class T> inline void cross_force(T* r, const T* v, const T* f) {    // v x* f (force)
  T a[3], b[3], c[3];
  v3cross(a, v, f); v3cross(b, v + 3, f + 3); v3cross(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (c0,c1,c2,c3), key = (k0,k1)
struct Philox {
  static inline void round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  static inline void gen(uint32_t out[4], uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed) {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int i = 0; i < 10; ++i) {
      round(c, k);
      k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    for (int i = 0; i < 4; ++i) out[i] = c[i];
  }
  // the idx-th 32-bit draw of (env, stream, step)
  static inline uint32_t draw(uint64_t seed, uint32_t env, uint32_t stream, uint32_t step, uint32_t idx) {
    uint32_t o[4]; gen(o, env, stream, step, idx >> 2, seed); return o[idx & 3];
  }
};
// uniform in [0,1): top 24 bits, exactly representable in float
inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// inverse normal CDF (Wichura AS241, PPND16)
inline double norm_ppf(double p) {
  const double q = p - 0.5;
  if (std::fabs(q) <= 0.425) {
    const double r = 0.180625 - q * q;
    return q * (((((((2509.0809287301226727 * r + 33430.575583588128105) * r + 67265.770927008700853) * r + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r + 133.14166789178437745) * r + 3.387132872796366608) /
           (((((((5226.495278852545925 * r + 28729.085735721942674) * r + 39307.89580009271061) * r + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r + 42.313330701600911252) * r + 1.0);
  }
  double r = q < 0 ? p : 1.0 - p;
  if (r <= 0) return q < 0 ? -1e30 : 1e30;
  r = std::sqrt(-std::log(r));
  double val;
  if (r <= 5.0) {
    r -= 1.6;
    val = (((((((7.7454501427834140764e-4 * r + 0.0227238449892691845833) * r + 0.24178072517745061177) * r + 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r + 4.6303378461565452959) * r + 1.42343711074968357734) /
          (((((((1.05075007164441684324e-9 * r + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r + 0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r + 2.05319162663775882187) * r + 1.0);
  } else {
    r -= 5.0;
    val = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r + 0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r + 5.4637849111641143699) * r + 6.6579046435011037772) /
          (((((((2.04426310338993978564e-15 * r + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r + 7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r + 0.59983224667449546444) * r + 1.0);
  }
  return q < 0 ? -val : val;
}

}  // namespace orc
