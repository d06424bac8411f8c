// Host instances of the engine's Riccati reductions (test infrastructure, not part of libd2dx.so): the SAME source the
// kernels inline -- care_gain (d2dx_device.cuh) and lqr5_gain (d2dx_lqr5.cuh) -- compiled for the CPU, so that
// tests/test_care_math.py can check them against scipy.linalg.solve_continuous_are without a GPU.
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

}  // extern "C"
