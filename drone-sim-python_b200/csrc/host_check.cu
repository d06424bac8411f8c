// Host instances of the engine's Riccati reductions (test infrastructure, not part of libd2dx.so): the SAME source the
// kernels inline -- care_gain (d2dx_device.cuh) and lqr5_gain (d2dx_lqr5.cuh) -- compiled for the CPU, so that
// tests/test_care_math.py can check them against scipy.linalg.solve_continuous_are without a GPU.
#include <vector>

#include "d2dx_ddp.cuh"
#include "d2dx_lqr5.cuh"

using namespace d2dx;

extern "C" {

// LQR gain of DFFFController.get (d2d/guidance.py:78-82) in the path frame (psi_ref = 0) at reference speed v and bank phi:
// K0[6] = row-major 2x3.  state5 (in/out): C, S, al, dth, dal; cold != 0 ignores it.  Returns 1 when converged.
int d2dx_host_care_gain(const double* qr4 /* q_pos, q_psi, r_phi, r_v */, double v, double phi, int cold, double* state5, double* K0) {
  d2dx_dfff_gains g;
  g.q_pos = qr4[0]; g.q_psi = qr4[1]; g.r_phi = qr4[2]; g.r_v = qr4[3];
  const CareConst cc = care_const(g);
  const double z = tan(phi), z2 = z * z, inv_va = 1.0 / v;
  const double c1 = cc.sr1 * (2.0 + z2) * rcp_f(kG * inv_va * (1.0 + z2));   // as make_ref forms them
  const double e = kG * inv_va * inv_va * z * c1;
  CareState st = {state5[0], state5[1], state5[2], state5[3], state5[4]};
  const bool ok = care_gain(cc, v, c1, e, st, cold != 0, K0);
  state5[0] = st.C; state5[1] = st.S; state5[2] = st.al; state5[3] = st.dth; state5[4] = st.dal;
  return ok ? 1 : 0;
}

// 5-state LQR gain of DiffController.ComputeGain (Controllers.py:159-186) in the path frame: Kp[10] = row-major 2x5.
// state7 (in/out): C, S, p23, p24, p33, p34, p44 (p23 <= 0 = cold).  Returns 1 when converged.
int d2dx_host_lqr5_gain(const double* q5, const double* r2, double v, double phi, double tau_phi, double tau_v, double* state7, double* Kp) {
  const double z = tan(phi), z2 = z * z, inv_va = 1.0 / v;
  Lqr5Par P;
  P.v = v;
  P.a = kG * inv_va * (1.0 + z2) * rcp_f(2.0 + z2);
  P.b = kG * inv_va * inv_va * z;
  P.itp = 1.0 / tau_phi; P.itv = 1.0 / tau_v;
  P.sq = ::sqrt(q5[0]); P.q3 = q5[2]; P.q4 = q5[3]; P.q5 = q5[4];
  const double sr1 = ::sqrt(r2[0]), sr2 = ::sqrt(r2[1]);
  P.s1 = tau_phi * sr1; P.s2 = tau_v * sr2; P.i1 = 1.0 / P.s1; P.i2 = 1.0 / P.s2;
  Lqr5State st = {state7[0], state7[1], state7[2], state7[3], state7[4], state7[5], state7[6]};
  const bool ok = lqr5_gain(P, sr1, sr2, tau_phi, tau_v, r2[0], r2[1], st, Kp);
  state7[0] = st.C; state7[1] = st.S; state7[2] = st.p23; state7[3] = st.p24; state7[4] = st.p33; state7[5] = st.p34; state7[6] = st.p44;
  return ok ? 1 : 0;
}

// The planner's second-order solver (d2dx_ddp.cuh) for ONE problem on the CPU: the code of ddp_solve_kernel's thread.
// prob9: N, h, wx, wy, vsp, kv, kb, kobs, obs_kind (weights already normalised as DdpProblem wants them); obs[n_obs][3];
// bounds4; box5 or NULL (x_lo, x_hi, y_lo, y_hi, w_box normalised); z0[3], zt[3]; u[2][N] in/out; xs[3][N] out; info[8] out.
int d2dx_host_ddp_solve(const double* prob9, int n_obs, const double* obs, const double* bounds4, const double* box5, const double* z0,
                        const double* zt, const d2dx_ddp_options* o, double* u, double* xs, double* info) {
  DdpProblem P;
  P.N = (int)prob9[0]; P.h = prob9[1]; P.wx = prob9[2]; P.wy = prob9[3]; P.vsp = prob9[4]; P.kv = prob9[5]; P.kb = prob9[6];
  P.kobs = prob9[7]; P.obs_kind = (int)prob9[8]; P.n_obs = n_obs;
  for (int k = 0; k < n_obs; ++k) { P.obs[k][0] = obs[3 * k]; P.obs[k][1] = obs[3 * k + 1]; P.obs[k][2] = obs[3 * k + 2]; }
  P.phi_lo = bounds4[0]; P.phi_hi = bounds4[1]; P.v_lo = bounds4[2]; P.v_hi = bounds4[3];
  P.has_box = box5 != nullptr;
  if (box5) { P.x_lo = box5[0]; P.x_hi = box5[1]; P.y_lo = box5[2]; P.y_hi = box5[3]; P.w_box = box5[4]; }
  else { P.x_lo = P.y_lo = -1e300; P.x_hi = P.y_hi = 1e300; P.w_box = 0.0; }
  const int N = P.N;
  std::vector<double> work(18 * (size_t)N);
  DdpWork W;
  W.N = N; W.stride = 1;
  W.u = work.data(); W.z = W.u + 2 * N; W.un = W.z + 3 * N; W.zn = W.un + 2 * N; W.k = W.zn + 3 * N; W.K = W.k + 2 * N;
  for (int i = 0; i < 2 * N; ++i) W.u[i] = u[i];
  d2dx_ddp_options opt;
  if (o) opt = *o;
  else {
    opt.max_iter = 400; opt.max_outer = 30; opt.max_inner = 40; opt.ls_max = 12; opt.ctol = 1e-8; opt.rel_tol = 1e-10; opt.abs_tol = 1e-14;
    opt.rho0 = 10.0; opt.rho_growth = 10.0; opt.rho_max = 1e8; opt.mu0 = 1e-6; opt.mu_min = 1e-8; opt.mu_max = 1e10; opt.mu_factor = 1.6; opt.reg_mode = 0; opt.min_solved = 0;
  }
  DdpSerial sweeps(P, W, zt);
  const DdpResult r = ddp_solve(sweeps, z0, opt);
  const double* us = (r.swaps & 1) ? W.un : W.u;
  const double* zs = (r.swaps & 1) ? W.zn : W.z;
  for (int i = 0; i < 2 * N; ++i) u[i] = us[i];
  for (int i = 0; i < 3 * N; ++i) xs[i] = zs[i];
  info[0] = r.flag; info[1] = r.iterations; info[2] = r.outer; info[3] = r.cost; info[4] = r.cmax; info[5] = r.lagr; info[6] = r.mu; info[7] = r.rho;
  return 0;
}

}  // extern "C"
