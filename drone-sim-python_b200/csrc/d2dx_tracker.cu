// 5-state LQR tracker on sampled references: Controllers.DiffFlatness.ComputeFlatness
// (Controllers.py:62-108), DiffController.ComputeGain (:159-186), implement_controller (10_opt_traj_tracking.py:72-89).
// One thread = one aircraft, state in registers, references read SoA from the sample table.
//
// The 5x5 Riccati equation.  In the path frame (T = rot(-psi_ref) (+) I3; q[0] = q[1]) the linearisation of
// d2d/dynamic.py:32-43 is  x' = dv,  y' = v psi,  psi' = a phi + b dv,  phi' = (u1 - phi)/tau_phi,  dv' = (u2 - dv)/tau_v
// (a, b the two entries the reference writes as g/va/(1+cos^2 phi) and g tan(phi)/va^2).  With s1 = tau_phi sqrt(r1),
// s2 = tau_v sqrt(r2) and kappa_j = (P_3j/s1, P_4j/s2) the Riccati entries (0,0), (1,1), (0,1) again force
// kappa_0 = sqrt(q)(C, S), kappa_1 = sqrt(q)(S, -C); the entries (2,2), (0,2), (0,3), (2,3) give P_12, P_01, P_02, P_22
// explicitly, and the remaining SIX unknowns (theta, P_23, P_24, P_33, P_34, P_44) solve the six equations
// (1,3), (3,3), (1,4), (2,4), (3,4), (4,4) -- Newton with the analytic 6x6 Jacobian, cold-started from the 3-state
// solution of d2dx_device.cuh, warm-started from the previous sample.  K' = [P_3./(tau_phi r1) ; P_4./(tau_v r2)],
// K = K' T.  Checked against scipy.linalg.solve_continuous_are for v in [2, 40], |phi| < 1.2, tau_phi in {0.01, 0.9667}:
// <= 7e-14 relative (prototype), and against the reference's own gains in tests/test_gpu_tracker.py.
#include "d2dx_device.cuh"
#include "d2dx_host.h"

namespace d2dx {

constexpr int kTrkThreads = 128;

struct Lqr5State { double C, S, p23, p24, p33, p34, p44; };    // al == 0 marks "cold"

struct Lqr5Par { double v, a, b, itp, itv, s1, s2, i1, i2, sq, q3, q4, q5; };

// solves the 6x6 system J d = -F without row exchanges: the equations are taken in the fixed order (1,4), (1,3), (2,4), (3,3), (3,4), (4,4)
// (rows 2, 0, 3, 1, 4, 5), the order in which plain elimination matched LAPACK to 2e-15 on 2700 Newton systems of this
// family (tau_phi 0.01 .. 0.97, v 4 .. 30 m/s, |phi| <= 1.1, condition numbers 11 .. 3900) -- the row exchanges of the
// pivoted version are 13 % of the tracker's instructions.  A pivot below 1e-9 of its row makes the step fail (the caller
// restarts cold and flags the aircraft if that fails too).
__device__ __forceinline__ bool solve6(double (&J)[6][6], double (&F)[6], double (&d)[6]) {
  constexpr int perm[6] = {2, 0, 3, 1, 4, 5};
  double A[6][7];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < 6; ++j) A[i][j] = J[perm[i]][j];
    A[i][6] = -F[perm[i]];
  }
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double amax = fabs(A[k][k]);
#pragma unroll
    for (int j = k + 1; j < 6; ++j) amax = fabs(A[k][j]) > amax ? fabs(A[k][j]) : amax;
    ok = ok && (fabs(A[k][k]) > 1e-9 * amax);
    const double ip = rcp_f(A[k][k]);
#pragma unroll
    for (int r = k + 1; r < 6; ++r) {
      const double f = A[r][k] * ip;
#pragma unroll
      for (int j = k + 1; j < 7; ++j) A[r][j] = fma(-f, A[k][j], A[r][j]);
    }
  }
  if (!ok) return false;                                    // also NaN: reported as not converged (flags), like a zero pivot
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = A[i][6];
#pragma unroll
    for (int j = i + 1; j < 6; ++j) s = fma(-A[i][j], d[j], s);
    d[i] = s * rcp_f(A[i][i]);
  }
  return true;
}

// Newton on (theta, p23, p24, p33, p34, p44); returns converged
__device__ bool lqr5_newton(const Lqr5Par& P, Lqr5State& u, int max_it) {
  for (int it = 0; it < max_it; ++it) {
    const double C = u.C, S = u.S, sq = P.sq;
    const double x2 = u.p23 * P.i1, y2 = u.p24 * P.i2, x3 = u.p33 * P.i1, y3 = u.p34 * P.i2, x4 = u.p34 * P.i1, y4 = u.p44 * P.i2;
    const double iv = rcp_f(P.v), ia = rcp_f(P.a);
    const double p12 = (x2 * x2 + y2 * y2 - P.q3) * 0.5 * iv;
    const double p01 = sq * (C * x2 + S * y2) * iv;
    const double p02 = sq * (P.s1 * C * P.itp + C * x3 + S * y3) * ia;
    const double p13 = P.s1 * sq * S;
    const double p22 = (u.p23 * P.itp + x2 * x3 + y2 * y3 - P.v * p13) * ia;
    double F[6], J[6][6], d[6];
    F[0] = P.a * p12 - p13 * P.itp - sq * (S * x3 - C * y3);
    F[1] = 2.0 * (P.a * u.p23 - u.p33 * P.itp) - (x3 * x3 + y3 * y3) + P.q4;
    F[2] = p01 + P.b * p12 + P.s2 * sq * C * P.itv - sq * (S * x4 - C * y4);
    F[3] = -P.v * P.s2 * sq * C + p02 + P.b * p22 - u.p24 * P.itv - (x2 * x4 + y2 * y4);
    F[4] = P.a * u.p24 + P.s1 * sq * C + P.b * u.p23 - u.p34 * (P.itp + P.itv) - (x3 * x4 + y3 * y4);
    F[5] = 2.0 * (P.s2 * sq * S + P.b * u.p24 - u.p44 * P.itv) - (x4 * x4 + y4 * y4) + P.q5;
    // d/dtheta
    const double dp01 = sq * (C * y2 - S * x2) * iv, dp02 = sq * (C * y3 - S * x3 - P.s1 * S * P.itp) * ia, dp22 = -P.v * P.s1 * sq * C * ia;
    J[0][0] = -P.s1 * sq * C * P.itp - sq * (C * x3 + S * y3);
    J[1][0] = 0.0;
    J[2][0] = dp01 - P.s2 * sq * S * P.itv - sq * (C * x4 + S * y4);
    J[3][0] = P.v * P.s2 * sq * S + dp02 + P.b * dp22;
    J[4][0] = -P.s1 * sq * S;
    J[5][0] = 2.0 * P.s2 * sq * C;
    // d/dp23
    J[0][1] = P.a * x2 * P.i1 * iv; J[1][1] = 2.0 * P.a; J[2][1] = (sq * C + P.b * x2) * P.i1 * iv;
    J[3][1] = P.b * (P.itp + x3 * P.i1) * ia - x4 * P.i1; J[4][1] = P.b; J[5][1] = 0.0;
    // d/dp24
    J[0][2] = P.a * y2 * P.i2 * iv; J[1][2] = 0.0; J[2][2] = (sq * S + P.b * y2) * P.i2 * iv;
    J[3][2] = P.b * y3 * P.i2 * ia - P.itv - y4 * P.i2; J[4][2] = P.a; J[5][2] = 2.0 * P.b;
    // d/dp33
    J[0][3] = -sq * S * P.i1; J[1][3] = -2.0 * (P.itp + x3 * P.i1); J[2][3] = 0.0;
    J[3][3] = (sq * C + P.b * x2) * P.i1 * ia; J[4][3] = -x4 * P.i1; J[5][3] = 0.0;
    // d/dp34
    J[0][4] = sq * C * P.i2; J[1][4] = -2.0 * y3 * P.i2; J[2][4] = -sq * S * P.i1;
    J[3][4] = (sq * S + P.b * y2) * P.i2 * ia - x2 * P.i1; J[4][4] = -(P.itp + P.itv) - (x3 * P.i1 + y4 * P.i2); J[5][4] = -2.0 * x4 * P.i1;
    // d/dp44
    J[0][5] = 0.0; J[1][5] = 0.0; J[2][5] = sq * C * P.i2; J[3][5] = -y2 * P.i2; J[4][5] = -y3 * P.i2; J[5][5] = -2.0 * (P.itv + y4 * P.i2);
    if (!solve6(J, F, d)) return false;
    const double Cn = C - S * d[0], Sn = S + C * d[0];
    const double nrm = rsqrt_f(Cn * Cn + Sn * Sn);
    u.C = Cn * nrm; u.S = Sn * nrm;
    u.p23 += d[1]; u.p24 += d[2]; u.p33 += d[3]; u.p34 += d[4]; u.p44 += d[5];
    const double scale = fabs(u.p23) + fabs(u.p24) + fabs(u.p33) + fabs(u.p34) + fabs(u.p44);
    const double step = fabs(d[1]) + fabs(d[2]) + fabs(d[3]) + fabs(d[4]) + fabs(d[5]);
    if (fabs(d[0]) < 3e-8 && step < 3e-8 * scale) return true;       // quadratic convergence: error ~1e-15 after this step
  }
  return false;
}

struct TrkRef { double xr[5], ur[2], k[10]; };

// ComputeFlatness, Controllers.py:62-108, formulas as written
__device__ __forceinline__ void flatness5(const double* Y /* x y xd yd xdd ydd xddd yddd */, double wx, double wy, double tau_phi,
                                          double tau_v, double* Xr, double* Ur, double& z, double& inv_va, double& cpsi, double& spsi) {
  const double v_ax = Y[2] - wx, v_ay = Y[3] - wy;
  const double v2 = v_ax * v_ax + v_ay * v_ay;
  inv_va = rsqrt_f(v2);
  const double va = v2 * inv_va;
  const double axd = Y[4], ayd = Y[5], axdd = Y[6], aydd = Y[7];
  const double num = v_ax * ayd - v_ay * axd;
  Xr[0] = Y[0]; Xr[1] = Y[1];
  Xr[2] = atan2_f(v_ay, v_ax);
  Xr[3] = atan2_f(num, kG * va);
  Xr[4] = va;
  const double va_dot = (v_ax * axd + v_ay * ayd) * inv_va;
  const double t1 = v_ax * axd - axd * v_ay;                         // as written at :95
  const double c1 = 1.0 + t1 * t1 * inv_va * inv_va;
  const double c2 = v_ax * aydd + axd * ayd - axdd * v_ay - axd * ayd;
  const double c3 = (v_ax * axd + v_ay * ayd) * num;
  const double phi_dot = rcp_f(c1) * (inv_va * inv_va) * (c2 * va - c3 * inv_va);
  Ur[0] = tau_phi * phi_dot + Xr[3];
  Ur[1] = tau_v * va_dot + va;
  z = num * inv_va * (1.0 / kG);                                     // tan(phi_ref) for kG va > 0
  cpsi = v_ax * inv_va; spsi = v_ay * inv_va;
}

// ComputeGain up to the gain: reference state / input and K (world frame, 2x5)
__device__ __forceinline__ void tracker_ref(const double* Y, double wx, double wy, double tau_phi, double tau_v,
                                            const d2dx_tracker_gains& g, Lqr5State& st, int& flags, TrkRef& r) {
  double z, inv_va, cpsi, spsi;
  flatness5(Y, wx, wy, tau_phi, tau_v, r.xr, r.ur, z, inv_va, cpsi, spsi);
  const double z2 = z * z;
  Lqr5Par P;
  P.v = r.xr[4];
  P.a = kG * inv_va * (1.0 + z2) * rcp_f(2.0 + z2);                  // g/va/(1+cos^2 phi), d2d/dynamic.py:38
  P.b = kG * inv_va * inv_va * z;                                    // g tan(phi)/va^2
  P.itp = 1.0 / tau_phi; P.itv = 1.0 / tau_v;
  P.sq = ::sqrt(g.q[0]); P.q3 = g.q[2]; P.q4 = g.q[3]; P.q5 = g.q[4];
  const double sr1 = ::sqrt(g.r[0]), sr2 = ::sqrt(g.r[1]);
  P.s1 = tau_phi * sr1; P.s2 = tau_v * sr2; P.i1 = 1.0 / P.s1; P.i2 = 1.0 / P.s2;
  bool ok = false;
  if (st.p23 > 0.0) {                                                // warm start from the previous sample
    Lqr5State w = st;
    ok = lqr5_newton(P, w, 12) && w.S > 0.0 && w.p23 > 0.0;
    if (ok) st = w;
  }
  if (!ok) {                                                         // cold: the 3-state gain as the first guess
    CareConst cc; cc.sq = P.sq; cc.q3 = P.q3; cc.sr1 = sr1; cc.sr2 = sr2; cc.isr1 = 1.0 / sr1; cc.isr2 = 1.0 / sr2;
    CareState c3 = {0.0, 1.0, 1.0, 0.0, 0.0};
    double K0[6];
    const double c1 = sr1 * rcp_f(P.a), e = P.b * c1;
    care_gain(cc, P.v, c1, e, c3, true, K0);
    const double al3 = K0[2] * sr1, be3 = K0[5] * sr2;
    st.C = c3.C; st.S = c3.S;
    st.p23 = P.s1 * al3; st.p24 = P.s2 * be3;
    st.p33 = tau_phi * (P.a * st.p23 + 0.5 * P.q4);
    st.p44 = tau_v * (P.s2 * P.sq * st.S + P.b * st.p24 + 0.5 * P.q5);
    st.p34 = (P.a * st.p24 + P.s1 * P.sq * st.C + P.b * st.p23) * rcp_f(P.itp + P.itv);
    ok = lqr5_newton(P, st, 40);
    if (!ok) flags |= 2;
  }
  // K' rows: P_3. / (tau_phi r1), P_4. / (tau_v r2); rotate the two position columns back to the world frame
  const double f1 = 1.0 / (tau_phi * g.r[0]), f2 = 1.0 / (tau_v * g.r[1]);
  const double k10 = P.s1 * P.sq * st.C * f1, k11 = P.s1 * P.sq * st.S * f1;
  const double k20 = P.s2 * P.sq * st.S * f2, k21 = -P.s2 * P.sq * st.C * f2;
  r.k[0] = k10 * cpsi - k11 * spsi; r.k[1] = k10 * spsi + k11 * cpsi; r.k[2] = st.p23 * f1; r.k[3] = st.p33 * f1; r.k[4] = st.p34 * f1;
  r.k[5] = k20 * cpsi - k21 * spsi; r.k[6] = k20 * spsi + k21 * cpsi; r.k[7] = st.p24 * f2; r.k[8] = st.p34 * f2; r.k[9] = st.p44 * f2;
}

// error, wraps (psi and phi, Controllers.py:165-166), saturations, feedback
__device__ __forceinline__ void tracker_feedback(const TrkRef& r, const double* X, const d2dx_tracker_gains& g, double* dX, double* U) {
#pragma unroll
  for (int k = 0; k < 5; ++k) dX[k] = X[k] - r.xr[k];
  dX[2] = wrap_pi(dX[2]); dX[3] = wrap_pi(dX[3]);
#pragma unroll
  for (int k = 0; k < 5; ++k) dX[k] = clip(dX[k], -g.err_sat[k], g.err_sat[k]);
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) s = fma(r.k[5 * m + k], dX[k], s);
    U[m] = clip(r.ur[m] - s, g.u_lo[m], g.u_hi[m]);
  }
}

__device__ __forceinline__ void load_state(const double* ls, size_t n, size_t i, Lqr5State& st) {
  st.C = 0.0; st.S = 1.0; st.p23 = 0.0; st.p24 = st.p33 = st.p34 = st.p44 = 0.0;
  if (ls) { st.C = ls[i]; st.S = ls[n + i]; st.p23 = ls[2 * n + i]; st.p24 = ls[3 * n + i]; st.p33 = ls[4 * n + i]; st.p34 = ls[5 * n + i]; st.p44 = ls[6 * n + i]; }
}
__device__ __forceinline__ void store_state(double* ls, size_t n, size_t i, const Lqr5State& st) {
  if (ls) { ls[i] = st.C; ls[n + i] = st.S; ls[2 * n + i] = st.p23; ls[3 * n + i] = st.p24; ls[4 * n + i] = st.p33; ls[5 * n + i] = st.p34; ls[6 * n + i] = st.p44; }
}

struct TrackerArgs {
  d2dx_tracker in;
  d2dx_tracker_out o;
  d2dx_tracker_gains g;
  int i_begin, i_end, nsub;
};

__global__ void __launch_bounds__(kTrkThreads) rollout_tracker_kernel(const __grid_constant__ TrackerArgs a) {
  const size_t M = a.in.M;
  const size_t j = (size_t)blockIdx.x * kTrkThreads + threadIdx.x;
  if (j >= M) return;
  const double wx = a.in.wind[j], wy = a.in.wind[M + j], tau_phi = a.in.ac[j], tau_v = a.in.ac[M + j];
  AcPar ap; ap.wx = wx; ap.wy = wy; ap.n_inv_tau_phi = -1.0 / tau_phi; ap.n_inv_tau_v = -1.0 / tau_v;
  double X[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) X[k] = a.in.X0[k * M + j];
  Lqr5State st;
  load_state(a.o.lqr_state, M, j, st);
  int flags = 0;
  if (a.o.X_log && a.i_begin == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) a.o.X_log[k * M + j] = X[k];
  }
  for (int i = a.i_begin + 1; i <= a.i_end; ++i) {
    double Y[8];
    const double* rp = a.in.ref + (size_t)i * 6 * M + j;
#pragma unroll
    for (int k = 0; k < 6; ++k) Y[k] = rp[k * M];
    Y[6] = 0.0; Y[7] = 0.0;                                          // Yddd_ref = [0, 0], 10_opt_traj_tracking.py:77
    TrkRef r;
    tracker_ref(Y, wx, wy, tau_phi, tau_v, a.g, st, flags, r);
    double dX[5], U[2];
    tracker_feedback(r, X, a.g, dX, U);
    const size_t row = (size_t)(i - 1);
    if (a.o.U_log) { a.o.U_log[(row * 2) * M + j] = U[0]; a.o.U_log[(row * 2 + 1) * M + j] = U[1]; }
    if (a.o.Xr_log) {
#pragma unroll
      for (int k = 0; k < 5; ++k) a.o.Xr_log[(row * 5 + k) * M + j] = r.xr[k];
    }
    if (a.o.dX_log) {
#pragma unroll
      for (int k = 0; k < 5; ++k) a.o.dX_log[(row * 5 + k) * M + j] = dX[k];
    }
    if (a.o.K_log) {
#pragma unroll
      for (int k = 0; k < 10; ++k) a.o.K_log[(row * 10 + k) * M + j] = r.k[k];
    }
    rk4_step(ap, X, U[0], U[1], a.in.dt, a.nsub);
    if (a.o.X_log) {
#pragma unroll
      for (int k = 0; k < 5; ++k) a.o.X_log[((size_t)i * 5 + k) * M + j] = X[k];
    }
  }
  if (!isfinite(X[0] + X[1] + X[2] + X[3] + X[4])) flags |= 1;
#pragma unroll
  for (int k = 0; k < 5; ++k) a.o.X_final[k * M + j] = X[k];
  if (a.o.flags) a.o.flags[j] |= flags;
  store_state(a.o.lqr_state, M, j, st);
}

__global__ void __launch_bounds__(kTrkThreads) flatness5_kernel(int n, const double* __restrict__ Ys, const double* __restrict__ W,
                                                                const double* __restrict__ ac, double* __restrict__ Xr, double* __restrict__ Ur) {
  const size_t i = (size_t)blockIdx.x * kTrkThreads + threadIdx.x;
  if (i >= (size_t)n) return;
  double Y[8], xr[5], ur[2], z, iva, c, s;
  for (int k = 0; k < 8; ++k) Y[k] = Ys[(size_t)k * n + i];
  flatness5(Y, W[i], W[n + i], ac[i], ac[n + i], xr, ur, z, iva, c, s);
  for (int k = 0; k < 5; ++k) Xr[(size_t)k * n + i] = xr[k];
  Ur[i] = ur[0]; Ur[(size_t)n + i] = ur[1];
}

__global__ void __launch_bounds__(kTrkThreads) tracker_control_kernel(int n, const double* __restrict__ X, const double* __restrict__ Ys,
                                                                      const double* __restrict__ W, const double* __restrict__ ac,
                                                                      const d2dx_tracker_gains g, double* __restrict__ U, double* __restrict__ Xr,
                                                                      double* __restrict__ dXo, double* __restrict__ K, double* __restrict__ ls) {
  const size_t i = (size_t)blockIdx.x * kTrkThreads + threadIdx.x;
  if (i >= (size_t)n) return;
  double Y[8], x[5], dX[5], u[2];
  for (int k = 0; k < 8; ++k) Y[k] = Ys[(size_t)k * n + i];
  for (int k = 0; k < 5; ++k) x[k] = X[(size_t)k * n + i];
  Lqr5State st;
  load_state(ls, n, i, st);
  int flags = 0;
  TrkRef r;
  tracker_ref(Y, W[i], W[n + i], ac[i], ac[n + i], g, st, flags, r);
  tracker_feedback(r, x, g, dX, u);
  U[i] = u[0]; U[(size_t)n + i] = u[1];
  if (Xr) for (int k = 0; k < 5; ++k) Xr[(size_t)k * n + i] = r.xr[k];
  if (dXo) for (int k = 0; k < 5; ++k) dXo[(size_t)k * n + i] = dX[k];
  if (K) for (int k = 0; k < 10; ++k) K[(size_t)k * n + i] = r.k[k];
  store_state(ls, n, i, st);
}

static int check_gains(const d2dx_tracker_gains& g, const char* who) {
  D2DX_CHECK_ARG(g.q[0] == g.q[1] && g.q[0] > 0 && g.q[2] > 0 && g.q[3] > 0 && g.q[4] > 0 && g.r[0] > 0 && g.r[1] > 0,
                 "%s: Q must be positive with q[0] == q[1], R positive", who);
  return D2DX_OK;
}

}  // namespace d2dx

using namespace d2dx;

extern "C" {

int d2dx_tracker_default_gains(d2dx_tracker_gains* g) {
  if (!g) return set_error(D2DX_EINVAL, "d2dx_tracker_default_gains: null");
  const double q[5] = {1, 1, 0.1, 0.01, 0.01};                       // Controllers.py:152
  for (int k = 0; k < 5; ++k) g->q[k] = q[k];
  g->r[0] = 8.0; g->r[1] = 1.0;
  g->err_sat[0] = 20.0; g->err_sat[1] = 20.0; g->err_sat[2] = kPi / 3; g->err_sat[3] = kPi / 4; g->err_sat[4] = 1.0;   // :147
  const double lim = 60.0 * (kPi / 180.0);                           // np.deg2rad(60), :149
  g->u_lo[0] = -lim; g->u_lo[1] = 4.0; g->u_hi[0] = lim; g->u_hi[1] = 20.0;
  return D2DX_OK;
}

int d2dx_flatness5(d2dx_handle* h, int32_t n, const double* Ys, const double* W, const double* ac, double* Xr, double* Ur, void* stream) {
  D2DX_CHECK_ARG(h && n > 0 && Ys && W && ac && Xr && Ur, "d2dx_flatness5: bad argument");
  D2DX_CUDA(cudaSetDevice(h->device));
  flatness5_kernel<<<(n + kTrkThreads - 1) / kTrkThreads, kTrkThreads, 0, as_stream(stream)>>>(n, Ys, W, ac, Xr, Ur);
  D2DX_LAUNCH_CHECK("flatness5_kernel");
  return D2DX_OK;
}

int d2dx_tracker_control(d2dx_handle* h, int32_t n, const double* X, const double* Ys, const double* W, const double* ac,
                         const d2dx_tracker_gains* gains_host, double* U, double* Xr, double* dX, double* K, double* lqr_state,
                         void* stream) {
  D2DX_CHECK_ARG(h && n > 0 && X && Ys && W && ac && U, "d2dx_tracker_control: bad argument");
  d2dx_tracker_gains g;
  if (gains_host) g = *gains_host; else d2dx_tracker_default_gains(&g);
  if (int rc = check_gains(g, "d2dx_tracker_control")) return rc;
  D2DX_CUDA(cudaSetDevice(h->device));
  tracker_control_kernel<<<(n + kTrkThreads - 1) / kTrkThreads, kTrkThreads, 0, as_stream(stream)>>>(n, X, Ys, W, ac, g, U, Xr, dX, K, lqr_state);
  D2DX_LAUNCH_CHECK("tracker_control_kernel");
  return D2DX_OK;
}

int d2dx_rollout_tracker(d2dx_handle* h, const d2dx_tracker* in, int32_t i_begin, int32_t i_end, int32_t nsub,
                         const d2dx_tracker_gains* gains_host, const d2dx_tracker_out* out, void* stream) {
  D2DX_CHECK_ARG(h && in && out, "d2dx_rollout_tracker: null argument");
  D2DX_CHECK_ARG(in->M > 0 && in->T >= 1 && in->ref && in->X0 && in->wind && in->ac && in->dt > 0 && out->X_final,
                 "d2dx_rollout_tracker: M=%d T=%d dt=%g or a missing array", in->M, in->T, in->dt);
  D2DX_CHECK_ARG(i_begin >= 0 && i_end >= i_begin && i_end < in->T && nsub >= 1, "d2dx_rollout_tracker: bad range [%d,%d] of %d or nsub=%d",
                 i_begin, i_end, in->T, nsub);
  TrackerArgs a;
  a.in = *in; a.o = *out; a.i_begin = i_begin; a.i_end = i_end; a.nsub = nsub;
  if (gains_host) a.g = *gains_host; else d2dx_tracker_default_gains(&a.g);
  if (int rc = check_gains(a.g, "d2dx_rollout_tracker")) return rc;
  D2DX_CUDA(cudaSetDevice(h->device));
  rollout_tracker_kernel<<<(in->M + kTrkThreads - 1) / kTrkThreads, kTrkThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("rollout_tracker_kernel");
  return D2DX_OK;
}

}  // extern "C"
