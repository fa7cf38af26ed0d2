// vm.cuh — fixed-size FP64 vector / rotation helpers shared by every kernel of libvils_b200.
// __host__ __device__ so that tests/hostcheck can compile the very same factor arithmetic for the CPU and
// compare it with the oracle without a GPU (the shipped library only ever calls them from kernels).
//
// Conventions (reference: vils_estimator/src/utility/utility.h:12-64, estimator.cpp:920-927):
//   quaternion storage x y z w, Hamilton product, body->world; tangent update q <- normalize(q (x) [1, dtheta/2]).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define VHD __host__ __device__ __forceinline__
#else
#define VHD inline
#endif

namespace vm {

// ---- forward-mode dual number with N partials (for the two autodiff functors of the reference) ----
template <int N>
struct Dual {
  double a;
  double v[N];
  VHD Dual() : a(0) {
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = 0;
  }
  VHD Dual(double s) : a(s) {  // NOLINT
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = 0;
  }
};
template <int N> VHD Dual<N> seed(double s, int k) { Dual<N> d(s); d.v[k] = 1.0; return d; }
template <int N> VHD Dual<N> operator+(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; h.a = f.a + g.a;
#pragma unroll
  for (int i = 0; i < N; i++) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> VHD Dual<N> operator-(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; h.a = f.a - g.a;
#pragma unroll
  for (int i = 0; i < N; i++) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> VHD Dual<N> operator-(const Dual<N>& f) { Dual<N> h; h.a = -f.a;
#pragma unroll
  for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h; }
template <int N> VHD Dual<N> operator*(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; h.a = f.a * g.a;
#pragma unroll
  for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> VHD Dual<N> operator/(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; const double gi = 1.0 / g.a, fg = f.a * gi; h.a = fg;
#pragma unroll
  for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - fg * g.v[i]) * gi; return h; }
template <int N> VHD Dual<N> chain(const Dual<N>& f, double val, double d) { Dual<N> h; h.a = val;
#pragma unroll
  for (int i = 0; i < N; i++) h.v[i] = d * f.v[i]; return h; }
template <int N> VHD Dual<N> dsqrt(const Dual<N>& f) { double s = sqrt(f.a); return chain(f, s, 0.5 / s); }
template <int N> VHD Dual<N> dsin(const Dual<N>& f) { return chain(f, sin(f.a), cos(f.a)); }
template <int N> VHD Dual<N> dacos(const Dual<N>& f) { return chain(f, acos(f.a), -1.0 / sqrt(1.0 - f.a * f.a)); }
VHD double dsqrt(double x) { return sqrt(x); }
VHD double dsin(double x) { return sin(x); }
VHD double dacos(double x) { return acos(x); }
VHD double val(double x) { return x; }
template <int N> VHD double val(const Dual<N>& f) { return f.a; }

// ---- 3-vectors / 3x3 ----
template <class T> struct V3 { T x, y, z; };
typedef V3<double> v3;
template <class T> VHD V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
VHD v3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }
template <class T> VHD V3<T> operator+(const V3<T>& a, const V3<T>& b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> VHD V3<T> operator-(const V3<T>& a, const V3<T>& b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> VHD V3<T> operator*(const V3<T>& a, const T& s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <class T> VHD T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> VHD V3<T> cross(const V3<T>& a, const V3<T>& b) { return mk<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

struct m3 { double m[3][3]; };
VHD m3 mul(const m3& a, const m3& b) { m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r; }
VHD m3 mulT(const m3& a, const m3& b) {  // a^T b
  m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[0][i] * b.m[0][j] + a.m[1][i] * b.m[1][j] + a.m[2][i] * b.m[2][j];
  return r; }
VHD m3 transpose(const m3& a) { m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
  return r; }
VHD v3 mul(const m3& a, const v3& v) { return mk(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z, a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z); }
VHD v3 mulT(const m3& a, const v3& v) { return mk(a.m[0][0] * v.x + a.m[1][0] * v.y + a.m[2][0] * v.z, a.m[0][1] * v.x + a.m[1][1] * v.y + a.m[2][1] * v.z, a.m[0][2] * v.x + a.m[1][2] * v.y + a.m[2][2] * v.z); }
VHD m3 skew(const v3& q) { m3 r; r.m[0][0] = 0; r.m[0][1] = -q.z; r.m[0][2] = q.y; r.m[1][0] = q.z; r.m[1][1] = 0; r.m[1][2] = -q.x; r.m[2][0] = -q.y; r.m[2][1] = q.x; r.m[2][2] = 0; return r; }
VHD m3 scale(const m3& a, double s) { m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] * s;
  return r; }
VHD m3 add(const m3& a, const m3& b) { m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r; }
VHD m3 sub(const m3& a, const m3& b) { m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j];
  return r; }
VHD m3 eye() { m3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i][j] = (i == j) ? 1.0 : 0.0;
  return r; }

// ---- quaternions (w x y z members; storage order in memory is x y z w) ----
template <class T> struct Q4 { T w, x, y, z; };
typedef Q4<double> q4;
template <class T> VHD Q4<T> mkq(T w, T x, T y, T z) { Q4<T> q; q.w = w; q.x = x; q.y = y; q.z = z; return q; }
VHD q4 ldq(const double* p) { return mkq(p[3], p[0], p[1], p[2]); }  // from x y z w
template <class T> VHD Q4<T> qmul(const Q4<T>& a, const Q4<T>& b) {
  return mkq<T>(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x); }
template <class T> VHD Q4<T> qconj(const Q4<T>& q) { return mkq<T>(q.w, -q.x, -q.y, -q.z); }
// Eigen QuaternionBase::inverse(): conjugate / squaredNorm (the autodiff functors call it on un-normalised slerps)
template <class T> VHD Q4<T> qinv(const Q4<T>& q) { T n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z; return mkq<T>(q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2); }
// FP64 specialisation: one reciprocal instead of four divides (a divide is ~20 instructions / ~120 cycles of latency on sm_100)
template <> VHD Q4<double> qinv<double>(const Q4<double>& q) { const double s = 1.0 / (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); return mkq<double>(q.w * s, -q.x * s, -q.y * s, -q.z * s); }
template <class T> VHD V3<T> qrot(const Q4<T>& q, const V3<T>& v) {  // v + 2w (u x v) + 2 u x (u x v)
  V3<T> u = mk<T>(q.x, q.y, q.z); V3<T> t = cross(u, v); t = t + t; return v + t * q.w + cross(u, t); }
VHD m3 q2R(const q4& q) {
  m3 r; const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.m[0][0] = 1 - (tyy + tzz); r.m[0][1] = txy - twz; r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz; r.m[1][1] = 1 - (txx + tzz); r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy; r.m[2][1] = tyz + twx; r.m[2][2] = 1 - (txx + tyy); return r; }
VHD q4 qnormalized(const q4& q) { double n = 1.0 / sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); return mkq(q.w * n, q.x * n, q.y * n, q.z * n); }
// Eigen slerp (shortest arc, linear blend when |dot| >= 1 - eps)
template <class T> VHD Q4<T> qslerp(const Q4<T>& a, double t, const Q4<T>& b) {
  T d = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
  T ad = val(d) < 0 ? -d : d;
  T s0, s1;
  if (val(ad) >= 1.0 - 2.220446049250313e-16) { s0 = T(1.0 - t); s1 = T(t); }
  else { T th = dacos(ad); T st = dsin(th); s0 = dsin(th * T(1.0 - t)) / st; s1 = dsin(th * T(t)) / st; }
  if (val(d) < 0) s1 = -s1;
  return mkq<T>(s0 * a.w + s1 * b.w, s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z); }

// PoseLocalParameterization::Plus (factor/pose_local_parameterization.cpp:3-19)
VHD void pose_plus(double* x, const double* d) {
  x[0] += d[0]; x[1] += d[1]; x[2] += d[2];
  q4 q = qnormalized(qmul(ldq(x + 3), mkq(1.0, d[3] * 0.5, d[4] * 0.5, d[5] * 0.5)));
  x[3] = q.x; x[4] = q.y; x[5] = q.z; x[6] = q.w; }

}  // namespace vm
