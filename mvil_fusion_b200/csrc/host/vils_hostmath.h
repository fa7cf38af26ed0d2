// vils_hostmath.h — small fixed-size quaternion / matrix helpers of the C++ host mirror (no Eigen in this image).
// Quaternions are stored x y z w like para_Pose (estimator.cpp:923-927); matrices are row-major.
#pragma once
#include <cmath>
#include <cstring>

namespace vils {
namespace hm {


inline void qmul(const double a[4], const double b[4], double o[4]) {   // x y z w, Hamilton
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[0] = aw * bx + ax * bw + ay * bz - az * by; o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx; o[3] = aw * bw - ax * bx - ay * by - az * bz;
}
inline void qrot(const double q[4], const double v[3], double o[3]) {
  const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
  double tx = 2 * (uy * v[2] - uz * v[1]), ty = 2 * (uz * v[0] - ux * v[2]), tz = 2 * (ux * v[1] - uy * v[0]);
  o[0] = v[0] + w * tx + (uy * tz - uz * ty); o[1] = v[1] + w * ty + (uz * tx - ux * tz); o[2] = v[2] + w * tz + (ux * ty - uy * tx);
}
inline void qrot_inv(const double q[4], const double v[3], double o[3]) { const double c[4] = {-q[0], -q[1], -q[2], q[3]}; qrot(c, v, o); }
inline void qnorm(double q[4]) { const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); for (int i = 0; i < 4; i++) q[i] /= n; }
inline void R2q(const double m[9], double q[4]) {   // Eigen Quaternion(Matrix3)
  double t = m[0] + m[4] + m[8];
  if (t > 0) { t = std::sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t; q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t; }
  else { int i = 0; if (m[4] > m[0]) i = 1; if (m[8] > m[i * 4]) i = 2; const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i * 4] - m[j * 4] - m[k * 4] + 1.0); q[i] = 0.5 * t; t = 0.5 / t; q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t; q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t; q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t; }
}

inline void q2R(const double* q, double* R) {   // Eigen toRotationMatrix, row-major
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
inline void mat3_mul(const double* A, const double* B, double* C) { double t[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j]; std::memcpy(C, t, sizeof(t)); }
inline void mat3_T(const double* A, double* B) { double t[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * j + i]; std::memcpy(B, t, sizeof(t)); }
inline void mat3_vec(const double* A, const double* v, double* o) { double t[3]; for (int i = 0; i < 3; i++) t[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2]; std::memcpy(o, t, sizeof(t)); }
inline void mat4_mul(const double* A, const double* B, double* C) { double t[16]; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double a = 0; for (int k = 0; k < 4; k++) a += A[4 * i + k] * B[4 * k + j]; t[4 * i + j] = a; } std::memcpy(C, t, sizeof(t)); }
inline void rigid4(const double* R, const double* t, double* T) { for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T[4 * i + j] = R[3 * i + j]; T[4 * i + 3] = t[i]; } T[12] = T[13] = T[14] = 0; T[15] = 1; }
inline void rigid4_inv(const double* T, double* Ti) {   // [R t]^-1 = [R^T  -R^T t]
  double R[9], t[3] = {T[3], T[7], T[11]}, Rt[9], o[3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = T[4 * i + j];
  mat3_T(R, Rt); mat3_vec(Rt, t, o); for (int i = 0; i < 3; i++) o[i] = -o[i];
  rigid4(Rt, o, Ti);
}
// Eigen QuaternionBase::slerp (shortest arc; linear blend when the quaternions are closer than epsilon)
inline void qslerp(const double a[4], const double b[4], double t, double o[4]) {
  const double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3], ad = std::fabs(d);
  double s0, s1;
  if (ad >= 1.0 - 2.220446049250313e-16) { s0 = 1.0 - t; s1 = t; }
  else { const double th = std::acos(ad), st = std::sin(th); s0 = std::sin((1.0 - t) * th) / st; s1 = std::sin(t * th) / st; }
  if (d < 0) s1 = -s1;
  for (int i = 0; i < 4; i++) o[i] = s0 * a[i] + s1 * b[i];
}
inline void qinv(const double q[4], double o[4]) { const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]; o[0] = -q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = q[3] / n2; }
inline double yaw_deg(const double* R) { return std::atan2(R[3], R[0]) / M_PI * 180.0; }   // Utility::R2ypr(R).x()


}  // namespace hm
}  // namespace vils
