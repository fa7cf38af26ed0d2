// vo_math.h — tiny fixed-size linear algebra + forward-mode dual numbers for the CPU oracle.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md). Nothing in the product path may include this.
// No Eigen / Ceres exists in this environment, so the handful of Eigen/Ceres primitives the
// reference relies on are restated here:
//   * Eigen::Quaternion product / conjugate / rotate / toRotationMatrix / slerp  [upstream Eigen 3.3]
//   * ceres::Jet forward-mode autodiff (used by the reference for LidarICPConstraint_b,
//     LPSConstraint, LidarEdgeFactor, LidarPlaneNormFactor)                      [upstream Ceres]
//   * Utility::{deltaQ, skewSymmetric, Qleft, Qright, R2ypr, ypr2R}
//     (vils_estimator/src/utility/utility.h:12-108)
#pragma once
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace vo {

// ------------------------------------------------------------------------------------------
// Jet<N>: value + N partial derivatives (ceres::Jet restated).
// ------------------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; i++) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; }  // NOLINT implicit on purpose
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double fg = f.a * gi; h.a = fg;
  for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - fg * g.v[i]) * gi; return h; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> h = -f; h.a += s; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; i++) h.v[i] = f.v[i] * s; return h; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, double s) { return f * (1.0 / s); }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) { return Jet<N>(s) / g; }
template <int N> inline bool operator<(const Jet<N>& f, const Jet<N>& g) { return f.a < g.a; }
template <int N> inline bool operator<(const Jet<N>& f, double g) { return f.a < g; }
template <int N> inline bool operator>=(const Jet<N>& f, const Jet<N>& g) { return f.a >= g.a; }
template <int N> inline bool operator>=(const Jet<N>& f, double g) { return f.a >= g; }
template <int N> inline Jet<N> jchain(const Jet<N>& f, double val, double dval) { Jet<N> h; h.a = val; for (int i = 0; i < N; i++) h.v[i] = dval * f.v[i]; return h; }
template <int N> inline Jet<N> sqrt(const Jet<N>& f) { double s = std::sqrt(f.a); return jchain(f, s, 0.5 / s); }
template <int N> inline Jet<N> sin(const Jet<N>& f) { return jchain(f, std::sin(f.a), std::cos(f.a)); }
template <int N> inline Jet<N> cos(const Jet<N>& f) { return jchain(f, std::cos(f.a), -std::sin(f.a)); }
template <int N> inline Jet<N> acos(const Jet<N>& f) { return jchain(f, std::acos(f.a), -1.0 / std::sqrt(1.0 - f.a * f.a)); }
template <int N> inline Jet<N> abs(const Jet<N>& f) { return f.a < 0 ? -f : f; }
inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double acos(double x) { return std::acos(x); }
inline double abs(double x) { return std::fabs(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float acos(float x) { return std::acos(x); }
inline float abs(float x) { return std::fabs(x); }
inline double val(double x) { return x; }
inline float val(float x) { return x; }
template <int N> inline double val(const Jet<N>& f) { return f.a; }

template <class T> struct Eps { static T eps() { return std::numeric_limits<T>::epsilon(); } };
template <int N> struct Eps<Jet<N>> { static Jet<N> eps() { return Jet<N>(std::numeric_limits<double>::epsilon()); } };

// ------------------------------------------------------------------------------------------
// Vec3 / Mat3 / Quat
// ------------------------------------------------------------------------------------------
template <class T> struct Vec3T {
  T x, y, z;
  Vec3T() : x(T(0)), y(T(0)), z(T(0)) {}
  Vec3T(T a, T b, T c) : x(a), y(b), z(c) {}
  T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> inline Vec3T<T> operator+(const Vec3T<T>& a, const Vec3T<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline Vec3T<T> operator-(const Vec3T<T>& a, const Vec3T<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline Vec3T<T> operator-(const Vec3T<T>& a) { return {-a.x, -a.y, -a.z}; }
template <class T> inline Vec3T<T> operator*(const Vec3T<T>& a, const T& s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline Vec3T<T> operator*(const T& s, const Vec3T<T>& a) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline Vec3T<T> operator/(const Vec3T<T>& a, const T& s) { return {a.x / s, a.y / s, a.z / s}; }
template <class T> inline T dot(const Vec3T<T>& a, const Vec3T<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline Vec3T<T> cross(const Vec3T<T>& a, const Vec3T<T>& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
template <class T> inline T norm(const Vec3T<T>& a) { return sqrt(dot(a, a)); }
typedef Vec3T<double> Vec3;

template <class T> struct Mat3T {
  T m[3][3];
  Mat3T() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = T(0); }
  static Mat3T I() { Mat3T r; r.m[0][0] = r.m[1][1] = r.m[2][2] = T(1); return r; }
  T& operator()(int i, int j) { return m[i][j]; }
  const T& operator()(int i, int j) const { return m[i][j]; }
};
typedef Mat3T<double> Mat3;
template <class T> inline Mat3T<T> operator*(const Mat3T<T>& a, const Mat3T<T>& b) {
  Mat3T<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { T s = T(0); for (int k = 0; k < 3; k++) s = s + a.m[i][k] * b.m[k][j]; r.m[i][j] = s; } return r; }
template <class T> inline Vec3T<T> operator*(const Mat3T<T>& a, const Vec3T<T>& v) {
  return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
          a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z}; }
template <class T> inline Mat3T<T> operator*(const Mat3T<T>& a, const T& s) { Mat3T<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] * s; return r; }
template <class T> inline Mat3T<T> operator+(const Mat3T<T>& a, const Mat3T<T>& b) { Mat3T<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
template <class T> inline Mat3T<T> operator-(const Mat3T<T>& a, const Mat3T<T>& b) { Mat3T<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
template <class T> inline Mat3T<T> operator-(const Mat3T<T>& a) { Mat3T<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = -a.m[i][j]; return r; }
template <class T> inline Mat3T<T> transpose(const Mat3T<T>& a) { Mat3T<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
// Utility::skewSymmetric (utility.h:26-34)
template <class T> inline Mat3T<T> skew(const Vec3T<T>& q) {
  Mat3T<T> r; r.m[0][1] = -q.z; r.m[0][2] = q.y; r.m[1][0] = q.z; r.m[1][2] = -q.x; r.m[2][0] = -q.y; r.m[2][1] = q.x; return r; }

template <class T> struct QuatT {
  T w, x, y, z;
  QuatT() : w(T(1)), x(T(0)), y(T(0)), z(T(0)) {}
  QuatT(T w_, T x_, T y_, T z_) : w(w_), x(x_), y(y_), z(z_) {}  // Eigen ctor order (w,x,y,z)
  Vec3T<T> vec() const { return {x, y, z}; }
};
typedef QuatT<double> Quat;
// Eigen quaternion product
template <class T> inline QuatT<T> operator*(const QuatT<T>& a, const QuatT<T>& b) {
  return QuatT<T>(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                  a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x); }
template <class T> inline QuatT<T> conj(const QuatT<T>& q) { return QuatT<T>(q.w, -q.x, -q.y, -q.z); }
template <class T> inline T qnorm2(const QuatT<T>& q) { return q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z; }
// Eigen QuaternionBase::inverse(): conjugate / squaredNorm
template <class T> inline QuatT<T> inverse(const QuatT<T>& q) { T n2 = qnorm2(q); return QuatT<T>(q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2); }
template <class T> inline QuatT<T> normalized(const QuatT<T>& q) { T n = sqrt(qnorm2(q)); return QuatT<T>(q.w / n, q.x / n, q.y / n, q.z / n); }
// Eigen QuaternionBase::_transformVector: v + w*uv + cross(q.vec, uv), uv = 2 cross(q.vec, v)
template <class T> inline Vec3T<T> rotate(const QuatT<T>& q, const Vec3T<T>& v) {
  Vec3T<T> qv = q.vec(); Vec3T<T> uv = cross(qv, v); uv = uv + uv; return v + uv * q.w + cross(qv, uv); }
// Eigen QuaternionBase::toRotationMatrix
template <class T> inline Mat3T<T> toR(const QuatT<T>& q) {
  Mat3T<T> r; const T tx = T(2) * q.x, ty = T(2) * q.y, tz = T(2) * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.m[0][0] = T(1) - (tyy + tzz); r.m[0][1] = txy - twz; r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz; r.m[1][1] = T(1) - (txx + tzz); r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy; r.m[2][1] = tyz + twx; r.m[2][2] = T(1) - (txx + tyy); return r; }
// Eigen Quaternion(Matrix3) ctor: quaternionbase_assign_impl<Matrix3> (Shepperd-style branch on the trace)
inline Quat fromR(const Mat3& m) {
  Quat q; double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) { t = std::sqrt(t + 1.0); q.w = 0.5 * t; t = 0.5 / t; q.x = (m(2, 1) - m(1, 2)) * t; q.y = (m(0, 2) - m(2, 0)) * t; q.z = (m(1, 0) - m(0, 1)) * t; }
  else { int i = 0; if (m(1, 1) > m(0, 0)) i = 1; if (m(2, 2) > m(i, i)) i = 2; int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0); double qv[3]; qv[i] = 0.5 * t; t = 0.5 / t; q.w = (m(k, j) - m(j, k)) * t; qv[j] = (m(j, i) + m(i, j)) * t; qv[k] = (m(k, i) + m(i, k)) * t; q.x = qv[0]; q.y = qv[1]; q.z = qv[2]; }
  return q; }
// Eigen QuaternionBase::slerp [upstream Eigen 3.3 Quaternion.h]
template <class T> inline QuatT<T> slerp(const QuatT<T>& a, const T& t, const QuatT<T>& b) {
  const T one = T(1) - Eps<T>::eps();
  T d = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
  T absD = abs(d);
  T scale0, scale1;
  if (absD >= one) { scale0 = T(1) - t; scale1 = t; }
  else { T theta = acos(absD); T sinTheta = sin(theta); scale0 = sin((T(1) - t) * theta) / sinTheta; scale1 = sin(t * theta) / sinTheta; }
  if (d < T(0)) scale1 = -scale1;
  return QuatT<T>(scale0 * a.w + scale1 * b.w, scale0 * a.x + scale1 * b.x, scale0 * a.y + scale1 * b.y, scale0 * a.z + scale1 * b.z); }
// Utility::deltaQ (utility.h:11-24): first-order, NOT normalised
template <class T> inline QuatT<T> deltaQ(const Vec3T<T>& th) { return QuatT<T>(T(1), th.x / T(2), th.y / T(2), th.z / T(2)); }

// Utility::Qleft / Qright (utility.h:46-64), 4x4 scalar-first; callers take bottomRightCorner<3,3>().
inline Mat3 Qleft33(const Quat& q) { Mat3 r = Mat3::I() * q.w + skew(q.vec()); return r; }
inline Mat3 Qright33(const Quat& q) { Mat3 r = Mat3::I() * q.w - skew(q.vec()); return r; }
inline void Qleft44(const Quat& q, double L[4][4]) {
  L[0][0] = q.w; L[0][1] = -q.x; L[0][2] = -q.y; L[0][3] = -q.z; L[1][0] = q.x; L[2][0] = q.y; L[3][0] = q.z;
  Mat3 b = Qleft33(q); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) L[1 + i][1 + j] = b(i, j); }
inline void Qright44(const Quat& q, double L[4][4]) {
  L[0][0] = q.w; L[0][1] = -q.x; L[0][2] = -q.y; L[0][3] = -q.z; L[1][0] = q.x; L[2][0] = q.y; L[3][0] = q.z;
  Mat3 b = Qright33(q); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) L[1 + i][1 + j] = b(i, j); }

// Utility::R2ypr (utility.h:66-81), degrees
inline Vec3 R2ypr(const Mat3& R) {
  Vec3 n(R(0, 0), R(1, 0), R(2, 0)), o(R(0, 1), R(1, 1), R(2, 1)), a(R(0, 2), R(1, 2), R(2, 2));
  double y = std::atan2(n.y, n.x);
  double p = std::atan2(-n.z, n.x * std::cos(y) + n.y * std::sin(y));
  double r = std::atan2(a.x * std::sin(y) - a.y * std::cos(y), -o.x * std::sin(y) + o.y * std::cos(y));
  return Vec3(y / M_PI * 180.0, p / M_PI * 180.0, r / M_PI * 180.0); }
// Utility::ypr2R (utility.h:83-108), degrees in
inline Mat3 ypr2R(const Vec3& ypr) {
  double y = ypr.x / 180.0 * M_PI, p = ypr.y / 180.0 * M_PI, r = ypr.z / 180.0 * M_PI;
  Mat3 Rz, Ry, Rx;
  Rz(0, 0) = std::cos(y); Rz(0, 1) = -std::sin(y); Rz(1, 0) = std::sin(y); Rz(1, 1) = std::cos(y); Rz(2, 2) = 1;
  Ry(0, 0) = std::cos(p); Ry(0, 2) = std::sin(p); Ry(1, 1) = 1; Ry(2, 0) = -std::sin(p); Ry(2, 2) = std::cos(p);
  Rx(0, 0) = 1; Rx(1, 1) = std::cos(r); Rx(1, 2) = -std::sin(r); Rx(2, 1) = std::sin(r); Rx(2, 2) = std::cos(r);
  return Rz * Ry * Rx; }

// ------------------------------------------------------------------------------------------
// Dense helpers (row-major unless noted)
// ------------------------------------------------------------------------------------------
typedef std::vector<double> dvec;

// In-place Cholesky A = L L^T (lower, row-major n x n). Returns false if not positive definite.
inline bool cholesky_lower(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d); A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) A[i * n + j] = 0.0;
  return true; }
inline void chol_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i * n + k] * b[k]; b[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= L[k * n + i] * b[k]; b[i] = s / L[i * n + i]; } }
// General inverse via partial-pivot LU (what Eigen's MatrixBase::inverse() does for n > 4).
inline bool inverse_lu(const double* A, int n, double* Ainv) {
  dvec a(A, A + n * n); std::vector<int> piv(n);
  for (int i = 0; i < n; i++) piv[i] = i;
  for (int k = 0; k < n; k++) {
    int p = k; double best = std::fabs(a[k * n + k]);
    for (int i = k + 1; i < n; i++) if (std::fabs(a[i * n + k]) > best) { best = std::fabs(a[i * n + k]); p = i; }
    if (best == 0.0) return false;
    if (p != k) { for (int j = 0; j < n; j++) std::swap(a[k * n + j], a[p * n + j]); std::swap(piv[k], piv[p]); }
    for (int i = k + 1; i < n; i++) { a[i * n + k] /= a[k * n + k]; double l = a[i * n + k]; for (int j = k + 1; j < n; j++) a[i * n + j] -= l * a[k * n + j]; }
  }
  for (int c = 0; c < n; c++) {
    dvec x(n);
    for (int i = 0; i < n; i++) x[i] = (piv[i] == c) ? 1.0 : 0.0;
    for (int i = 0; i < n; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= a[i * n + k] * x[k]; x[i] = s; }
    for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < n; k++) s -= a[i * n + k] * x[k]; x[i] = s / a[i * n + i]; }
    for (int i = 0; i < n; i++) Ainv[i * n + c] = x[i];
  }
  return true; }
// Symmetric eigen-decomposition by cyclic Jacobi: A = V diag(w) V^T, V columns = eigenvectors (row-major V).
// Stands in for Eigen::SelfAdjointEigenSolver (marginalization_factor.cpp:275,301); eigenvalues ascending.
inline void eigh_jacobi(const double* Ain, int n, double* w, double* V) {
  dvec A(Ain, Ain + n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) { diag += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j]; }
    if (off <= 1e-60 * (diag + 1e-300) || off == 0.0) break;
    for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
      double apq = A[p * n + q]; if (apq == 0.0) continue;
      double app = A[p * n + p], aqq = A[q * n + q];
      double tau = (aqq - app) / (2.0 * apq);
      double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
      double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
      for (int k = 0; k < n; k++) { double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
      for (int k = 0; k < n; k++) { double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
    }
  }
  for (int i = 0; i < n; i++) w[i] = A[i * n + i];
  // sort ascending (Eigen convention)
  for (int i = 0; i < n - 1; i++) { int m = i; for (int j = i + 1; j < n; j++) if (w[j] < w[m]) m = j;
    if (m != i) { std::swap(w[i], w[m]); for (int k = 0; k < n; k++) std::swap(V[k * n + i], V[k * n + m]); } }
}

}  // namespace vo
