// vils_preint.cu — batched IMU mid-point pre-integration: IntegrationBase::push_back / propagate / repropagate
// (vils_estimator/src/factor/integration_base.h:30-158).  One warp per interval; the 15x15 state-transition products
// (jacobian = F jacobian, covariance = F cov F^T + V noise V^T, :124-125) are spread over the lanes, the sample loop is
// sequential in time as in the reference.  F and V are kept dense in shared memory (15x15, 15x18).
#include "common.h"
#include "factors.cuh"

namespace {
using namespace vm;

struct PreintArgs {
  int n; const int32_t* off; const double* dt; const double* acc; const double* gyr; const double* acc0; const double* gyr0;
  const double* ba; const double* bg; double noise[4]; double* out;
};

__device__ void put(double* M, int ld, int r, int c, const m3& a, double s) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M[(r + i) * ld + c + j] = s * a.m[i][j];
}

__global__ void __launch_bounds__(32) preint_kernel(PreintArgs A) {
  const int k = blockIdx.x, lane = threadIdx.x;
  if (k >= A.n) return;
  __shared__ double J[225], P[225], F[225], V[270], T[225], T2[225];
  __shared__ double st[16];   // dp(3) dq(4: w x y z) dv(3) acc0(3) gyr0(3)
  for (int e = lane; e < 225; e += 32) { J[e] = (e / 15 == e % 15) ? 1.0 : 0.0; P[e] = 0.0; }   // :17
  if (lane == 0) {
    st[0] = st[1] = st[2] = 0; st[3] = 1; st[4] = st[5] = st[6] = 0; st[7] = st[8] = st[9] = 0;
    for (int a = 0; a < 3; a++) { st[10 + a] = A.acc0[3 * k + a]; st[13 + a] = A.gyr0[3 * k + a]; }
  }
  const v3 ba = ld3(A.ba + 3 * k), bg = ld3(A.bg + 3 * k);
  const double nz[6] = {A.noise[0] * A.noise[0], A.noise[1] * A.noise[1], A.noise[0] * A.noise[0], A.noise[1] * A.noise[1],
                        A.noise[2] * A.noise[2], A.noise[3] * A.noise[3]};                            // :21-27
  double sum_dt = 0;
  __syncwarp();
  for (int s = A.off[k]; s < A.off[k + 1]; s++) {
    const double dt = A.dt[s];
    for (int e = lane; e < 225; e += 32) F[e] = 0.0;
    for (int e = lane; e < 270; e += 32) V[e] = 0.0;
    __syncwarp();
    if (lane == 0) {
      const v3 a0 = mk(st[10], st[11], st[12]), g0 = mk(st[13], st[14], st[15]), a1 = ld3(A.acc + 3 * s), g1 = ld3(A.gyr + 3 * s);
      const q4 dq = mkq(st[3], st[4], st[5], st[6]);
      const v3 dp = mk(st[0], st[1], st[2]), dv = mk(st[7], st[8], st[9]);
      const v3 un_acc_0 = qrot(dq, a0 - ba);                                                            // :63
      const v3 w = (g0 + g1) * 0.5 - bg;                                                                // :64
      const q4 rq = qmul(dq, mkq(1.0, w.x * dt / 2, w.y * dt / 2, w.z * dt / 2));                       // :65
      const v3 un_acc_1 = qrot(rq, a1 - ba);
      const v3 un_acc = (un_acc_0 + un_acc_1) * 0.5;
      const v3 rp = dp + dv * dt + un_acc * (0.5 * dt * dt);
      const v3 rv = dv + un_acc * dt;
      const m3 Rw = skew(w), Ra0 = skew(a0 - ba), Ra1 = skew(a1 - ba), Rd = q2R(dq), Rr = q2R(rq), I3 = eye();
      const m3 IRw = sub(I3, scale(Rw, dt));
      const m3 RdA0 = mul(Rd, Ra0), RrA1 = mul(Rr, Ra1), RrA1I = mul(RrA1, IRw);
      put(F, 15, 0, 0, I3, 1.0);
      put(F, 15, 0, 3, add(scale(RdA0, -0.25 * dt * dt), scale(RrA1I, -0.25 * dt * dt)), 1.0);        // :91-92
      put(F, 15, 0, 6, I3, dt);
      put(F, 15, 0, 9, add(Rd, Rr), -0.25 * dt * dt);
      put(F, 15, 0, 12, RrA1, -0.25 * dt * dt * -dt);
      put(F, 15, 3, 3, IRw, 1.0);
      put(F, 15, 3, 12, I3, -dt);
      put(F, 15, 6, 3, add(scale(RdA0, -0.5 * dt), scale(RrA1I, -0.5 * dt)), 1.0);
      put(F, 15, 6, 6, I3, 1.0);
      put(F, 15, 6, 9, add(Rd, Rr), -0.5 * dt);
      put(F, 15, 6, 12, RrA1, -0.5 * dt * -dt);
      put(F, 15, 9, 9, I3, 1.0);
      put(F, 15, 12, 12, I3, 1.0);
      put(V, 18, 0, 0, Rd, 0.25 * dt * dt);                                                             // :108-120
      put(V, 18, 0, 3, RrA1, -0.25 * dt * dt * 0.5 * dt);
      put(V, 18, 0, 6, Rr, 0.25 * dt * dt);
      put(V, 18, 0, 9, RrA1, -0.25 * dt * dt * 0.5 * dt);
      put(V, 18, 3, 3, I3, 0.5 * dt);
      put(V, 18, 3, 9, I3, 0.5 * dt);
      put(V, 18, 6, 0, Rd, 0.5 * dt);
      put(V, 18, 6, 3, RrA1, -0.5 * dt * 0.5 * dt);
      put(V, 18, 6, 6, Rr, 0.5 * dt);
      put(V, 18, 6, 9, RrA1, -0.5 * dt * 0.5 * dt);
      put(V, 18, 9, 12, I3, dt);
      put(V, 18, 12, 15, I3, dt);
      const q4 nq = qnormalized(rq);                                                                     // :153
      st[0] = rp.x; st[1] = rp.y; st[2] = rp.z; st[3] = nq.w; st[4] = nq.x; st[5] = nq.y; st[6] = nq.z; st[7] = rv.x; st[8] = rv.y; st[9] = rv.z;
      st[10] = a1.x; st[11] = a1.y; st[12] = a1.z; st[13] = g1.x; st[14] = g1.y; st[15] = g1.z;
    }
    __syncwarp();
    for (int e = lane; e < 225; e += 32) {          // T = F J ; T2 = F P
      const int i = e / 15, j = e % 15; double a = 0, b = 0;
      for (int m = 0; m < 15; m++) { a += F[i * 15 + m] * J[m * 15 + j]; b += F[i * 15 + m] * P[m * 15 + j]; }
      T[e] = a; T2[e] = b;
    }
    __syncwarp();
    for (int e = lane; e < 225; e += 32) {          // J = T ; P = T2 F^T + V Q V^T
      const int i = e / 15, j = e % 15; double a = 0, b = 0;
      for (int m = 0; m < 15; m++) a += T2[i * 15 + m] * F[j * 15 + m];
      for (int m = 0; m < 18; m++) b += V[i * 18 + m] * nz[m / 3] * V[j * 18 + m];
      J[e] = T[e]; P[e] = a + b;
    }
    sum_dt += dt;
    __syncwarp();
  }
  double* o = A.out + (size_t)k * 467;
  if (lane == 0) {
    o[0] = st[0]; o[1] = st[1]; o[2] = st[2]; o[3] = st[4]; o[4] = st[5]; o[5] = st[6]; o[6] = st[3]; o[7] = st[7]; o[8] = st[8]; o[9] = st[9];
    o[10] = ba.x; o[11] = ba.y; o[12] = ba.z; o[13] = bg.x; o[14] = bg.y; o[15] = bg.z; o[16] = sum_dt;
  }
  for (int e = lane; e < 225; e += 32) { const int r = e / 15, c = e % 15; o[17 + c * 15 + r] = J[e]; o[242 + c * 15 + r] = P[e]; }   // column-major like Eigen
}

}  // namespace

extern "C" int vils_preintegrate(int32_t n, const int32_t* off, const double* dt, const double* acc, const double* gyr, const double* acc0,
                                 const double* gyr0, const double* ba, const double* bg, const double noise[4], vils_preint* out, int32_t device) {
  if (n < 0 || !off || !dt || !acc || !gyr || !acc0 || !gyr0 || !ba || !bg || !noise || !out) return vils::fail(VILS_ERR_BAD_ARG, "vils_preintegrate: null");
  int st = vils::require_device(device); if (st) return st;
  if (n == 0) return VILS_OK;
  const int ns = off[n];
  for (int k = 0; k < n; k++) if (off[k] > off[k + 1] || off[k] < 0) return vils::fail(VILS_ERR_BAD_ARG, "vils_preintegrate: offsets must be non-decreasing");
  char* d = nullptr;
  const size_t b_off = sizeof(int32_t) * (n + 1), b_s = sizeof(double) * ns, b_k = sizeof(double) * 3 * n, b_out = sizeof(double) * 467 * n;
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t o_off = 0, o_dt = al(b_off), o_acc = o_dt + al(b_s), o_gyr = o_acc + al(3 * b_s), o_a0 = o_gyr + al(3 * b_s), o_g0 = o_a0 + al(b_k),
               o_ba = o_g0 + al(b_k), o_bg = o_ba + al(b_k), o_out = o_bg + al(b_k), total = o_out + al(b_out);
  cudaError_t e = cudaMalloc(&d, total);
  if (e != cudaSuccess) return vils::fail_cuda(e, "vils_preintegrate alloc");
  cudaMemcpy(d + o_off, off, b_off, cudaMemcpyHostToDevice); cudaMemcpy(d + o_dt, dt, b_s, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_acc, acc, 3 * b_s, cudaMemcpyHostToDevice); cudaMemcpy(d + o_gyr, gyr, 3 * b_s, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_a0, acc0, b_k, cudaMemcpyHostToDevice); cudaMemcpy(d + o_g0, gyr0, b_k, cudaMemcpyHostToDevice);
  cudaMemcpy(d + o_ba, ba, b_k, cudaMemcpyHostToDevice); cudaMemcpy(d + o_bg, bg, b_k, cudaMemcpyHostToDevice);
  PreintArgs A; A.n = n; A.off = (const int32_t*)(d + o_off); A.dt = (const double*)(d + o_dt); A.acc = (const double*)(d + o_acc);
  A.gyr = (const double*)(d + o_gyr); A.acc0 = (const double*)(d + o_a0); A.gyr0 = (const double*)(d + o_g0); A.ba = (const double*)(d + o_ba);
  A.bg = (const double*)(d + o_bg); for (int i = 0; i < 4; i++) A.noise[i] = noise[i]; A.out = (double*)(d + o_out);
  preint_kernel<<<n, 32>>>(A);
  e = cudaMemcpy(out, d + o_out, b_out, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? VILS_OK : vils::fail_cuda(e, "vils_preintegrate");
}
