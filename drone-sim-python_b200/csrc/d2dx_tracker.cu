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
#include "d2dx_lqr5.cuh"
#include "d2dx_host.h"

namespace d2dx {

constexpr int kTrkThreads = 128;

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
  double Kp[10];
  if (!lqr5_gain(P, sr1, sr2, tau_phi, tau_v, g.r[0], g.r[1], st, Kp)) flags |= 2;
  // rotate the two position columns back to the world frame
  r.k[0] = Kp[0] * cpsi - Kp[1] * spsi; r.k[1] = Kp[0] * spsi + Kp[1] * cpsi; r.k[2] = Kp[2]; r.k[3] = Kp[3]; r.k[4] = Kp[4];
  r.k[5] = Kp[5] * cpsi - Kp[6] * spsi; r.k[6] = Kp[5] * spsi + Kp[6] * cpsi; r.k[7] = Kp[7]; r.k[8] = Kp[8]; r.k[9] = Kp[9];
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
  D2DX_NVTX("d2dx_flatness5");
  D2DX_CHECK_ARG(h && n > 0 && Ys && W && ac && Xr && Ur, "d2dx_flatness5: bad argument");
  D2DX_CUDA(cudaSetDevice(h->device));
  flatness5_kernel<<<(n + kTrkThreads - 1) / kTrkThreads, kTrkThreads, 0, as_stream(stream)>>>(n, Ys, W, ac, Xr, Ur);
  D2DX_LAUNCH_CHECK("flatness5_kernel");
  return D2DX_OK;
}

int d2dx_tracker_control(d2dx_handle* h, int32_t n, const double* X, const double* Ys, const double* W, const double* ac,
                         const d2dx_tracker_gains* gains_host, double* U, double* Xr, double* dX, double* K, double* lqr_state,
                         void* stream) {
  D2DX_NVTX("d2dx_tracker_control");
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
  D2DX_NVTX("d2dx_rollout_tracker");
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
