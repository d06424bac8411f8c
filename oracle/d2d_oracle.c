/* CPU oracle for the d2d closed-loop rollout, plain C -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A restatement of the reference's algorithm for path A (run_simulation, 05_test_simulation.py:21-34, with
 * DFFFController.get d2d/guidance.py:62-91, DiffFlatness :23-47, Aircraft.cont_dyn / cont_jac
 * d2d/dynamic.py:14-43 and the circle / line / min-snap trajectories of d2d/trajectory.py), fast enough for
 * parity checks on thousands of scenarios and for the CPU baseline of bench.py (OpenMP over scenarios).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Independence from the product: the LQR gain is obtained the generic way the reference does (a full 3x3
 * continuous algebraic Riccati equation on the un-rotated (A1, B1) of guidance.py:78), here by the matrix
 * sign function of the 6x6 Hamiltonian followed by two Newton-Kleinman refinements -- NOT by the reduced
 * two-unknown Newton iteration of the CUDA kernel.  Pinned against the golden vectors of the unmodified
 * reference (SciPy's Schur-based CARE) in tests/test_oracle_c.py. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define PI 3.141592653589793
#define G 9.81

static double wrap_pi(double v) { /* d2d/utils.py:7, NumPy floored modulo */
  double a = v + PI, m = fmod(a, 2 * PI);
  if (m != 0.0) { if (m < 0.0) m += 2 * PI; } else m = 0.0;
  return m - PI;
}
static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- small dense linear algebra (Gaussian elimination with partial pivoting) ---- */
static int solve_n(int n, double* A, double* b, int nrhs) { /* A n x n row-major, b n x nrhs row-major, in place */
  for (int k = 0; k < n; ++k) {
    int p = k; double best = fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i) if (fabs(A[i * n + k]) > best) { best = fabs(A[i * n + k]); p = i; }
    if (best == 0.0) return -1;
    if (p != k) {
      for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
      for (int j = 0; j < nrhs; ++j) { double t = b[k * nrhs + j]; b[k * nrhs + j] = b[p * nrhs + j]; b[p * nrhs + j] = t; }
    }
    for (int i = k + 1; i < n; ++i) {
      double f = A[i * n + k] / A[k * n + k];
      if (f == 0.0) continue;
      for (int j = k; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
      for (int j = 0; j < nrhs; ++j) b[i * nrhs + j] -= f * b[k * nrhs + j];
    }
  }
  for (int i = n - 1; i >= 0; --i)
    for (int j = 0; j < nrhs; ++j) {
      double s = b[i * nrhs + j];
      for (int k = i + 1; k < n; ++k) s -= A[i * n + k] * b[k * nrhs + j];
      b[i * nrhs + j] = s / A[i * n + i];
    }
  return 0;
}

/* K = lqr(A (3x3), B (3x2), diag(q), diag(r)): control.lqr of guidance.py:80 */
static int lqr3(const double A[9], const double B[6], const double q[3], const double r[2], double K[6]) {
  double Gm[9], H[36], Z[36], Zi[36], I6[36];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Gm[i * 3 + j] = B[i * 2] * B[j * 2] / r[0] + B[i * 2 + 1] * B[j * 2 + 1] / r[1];
  memset(H, 0, sizeof H);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    H[i * 6 + j] = A[i * 3 + j]; H[i * 6 + 3 + j] = -Gm[i * 3 + j];
    H[(3 + i) * 6 + 3 + j] = -A[j * 3 + i];
  }
  for (int i = 0; i < 3; ++i) H[(3 + i) * 6 + i] = -q[i];
  memcpy(Z, H, sizeof Z);
  for (int it = 0; it < 100; ++it) { /* sign iteration with determinant scaling */
    double M[36];
    memcpy(M, Z, sizeof M);
    memset(I6, 0, sizeof I6); for (int i = 0; i < 6; ++i) I6[i * 6 + i] = 1.0;
    /* determinant via a copy */
    double D[36]; memcpy(D, Z, sizeof D); double det = 1.0;
    for (int k = 0; k < 6; ++k) {
      int p = k; for (int i = k + 1; i < 6; ++i) if (fabs(D[i * 6 + k]) > fabs(D[p * 6 + k])) p = i;
      if (D[p * 6 + k] == 0.0) { det = 0.0; break; }
      if (p != k) { for (int j = 0; j < 6; ++j) { double t = D[k * 6 + j]; D[k * 6 + j] = D[p * 6 + j]; D[p * 6 + j] = t; } det = -det; }
      det *= D[k * 6 + k];
      for (int i = k + 1; i < 6; ++i) { double f = D[i * 6 + k] / D[k * 6 + k]; for (int j = k; j < 6; ++j) D[i * 6 + j] -= f * D[k * 6 + j]; }
    }
    if (det == 0.0) return -1;
    if (solve_n(6, M, I6, 6)) return -1;        /* I6 <- Z^-1 */
    memcpy(Zi, I6, sizeof Zi);
    double c = pow(fabs(det), -1.0 / 6.0), diff = 0.0, nrm = 0.0;
    for (int k = 0; k < 36; ++k) {
      double zn = 0.5 * (c * Z[k] + Zi[k] / c);
      diff += (zn - Z[k]) * (zn - Z[k]); nrm += zn * zn; Z[k] = zn;
    }
    if (diff <= 1e-28 * nrm) break;
  }
  /* [W12; W22 + I] P = -[W11 + I; W21]  (12 x 3 least squares via normal equations) */
  double Mm[36], Nn[36]; /* 12x3 each */
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    Mm[i * 3 + j] = Z[i * 6 + 3 + j]; Mm[(3 + i) * 3 + j] = Z[(3 + i) * 6 + 3 + j] + (i == j);
    Nn[i * 3 + j] = -(Z[i * 6 + j] + (i == j)); Nn[(3 + i) * 3 + j] = -Z[(3 + i) * 6 + j];
  }
  double MtM[9] = {0}, MtN[9] = {0};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 6; ++k) {
    MtM[i * 3 + j] += Mm[k * 3 + i] * Mm[k * 3 + j]; MtN[i * 3 + j] += Mm[k * 3 + i] * Nn[k * 3 + j];
  }
  if (solve_n(3, MtM, MtN, 3)) return -1;
  double P[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) P[i * 3 + j] = 0.5 * (MtN[i * 3 + j] + MtN[j * 3 + i]);
  /* Newton-Kleinman refinement: (A - G P)^T X + X (A - G P) = -(Q + P G P), 6 unknowns */
  for (int ref = 0; ref < 3; ++ref) {
    double Ac[9], PGP[9], GP[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += Gm[i * 3 + k] * P[k * 3 + j]; GP[i * 3 + j] = s; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += P[i * 3 + k] * GP[k * 3 + j]; PGP[i * 3 + j] = s; Ac[i * 3 + j] = A[i * 3 + j] - GP[i * 3 + j]; }
    static const int ui[6] = {0, 0, 0, 1, 1, 2}, uj[6] = {0, 1, 2, 1, 2, 2};
    double L[36], rhs[6];
    memset(L, 0, sizeof L);
    for (int e = 0; e < 6; ++e) { /* equation (i,j): sum_k Ac[k][i] X[k][j] + X[i][k] Ac[k][j] */
      int i = ui[e], j = uj[e];
      for (int k = 0; k < 3; ++k) {
        int a = k < j ? k : j, b = k < j ? j : k, u1 = -1, u2 = -1;
        for (int u = 0; u < 6; ++u) if (ui[u] == a && uj[u] == b) u1 = u;
        L[e * 6 + u1] += Ac[k * 3 + i];
        a = i < k ? i : k; b = i < k ? k : i;
        for (int u = 0; u < 6; ++u) if (ui[u] == a && uj[u] == b) u2 = u;
        L[e * 6 + u2] += Ac[k * 3 + j];
      }
      rhs[e] = -((i == j ? q[i] : 0.0) + PGP[i * 3 + j]);
    }
    if (solve_n(6, L, rhs, 1)) return -1;
    for (int u = 0; u < 6; ++u) { P[ui[u] * 3 + uj[u]] = rhs[u]; P[uj[u] * 3 + ui[u]] = rhs[u]; }
  }
  for (int m = 0; m < 2; ++m) for (int j = 0; j < 3; ++j) { /* K = R^-1 B^T P */
    double s = 0; for (int k = 0; k < 3; ++k) s += B[k * 2 + m] * P[k * 3 + j];
    K[m * 3 + j] = s / r[m];
  }
  return 0;
}

/* ---- trajectories: par layout = 17 doubles, slot 0 = t0 (same meaning as the reference attributes) ---- */
enum { T_LINE = 0, T_CIRCLE = 1, T_POLY = 3 };

static double arrn(int k, int n) { double a = 1; for (int i = n; i > n - k; --i) a *= i; return a; } /* trajectory.py:41-45 */

static void traj_get(int type, const double* p, double t, double Y[8]) { /* Y[2*k+c] */
  memset(Y, 0, 8 * sizeof(double));
  if (type == T_LINE) { /* trajectory.py:136-141 */
    double dt = t - p[0];
    Y[0] = p[1] + p[3] * dt; Y[1] = p[2] + p[4] * dt; Y[2] = p[3]; Y[3] = p[4];
  } else if (type == T_CIRCLE) { /* trajectory.py:153-160 */
    double r = p[3], om = p[4], alpha = (t - p[0]) * om + p[5], ca = cos(alpha), sa = sin(alpha);
    Y[0] = p[1] + r * ca; Y[1] = p[2] + r * sa;
    Y[2] = om * r * -sa; Y[3] = om * r * ca;
    Y[4] = om * om * r * -ca; Y[5] = om * om * r * -sa;
    Y[6] = om * om * om * r * sa; Y[7] = om * om * om * r * -ca;
  } else { /* T_POLY: trajectory.py:74-82,185-187 */
    double dt = t - p[0];
    for (int c = 0; c < 2; ++c) {
      const double* c0 = p + 1 + 8 * c;
      for (int d = 0; d < 4; ++d) {
        double v = 0.0; /* coefs[d][7] is a structural zero for d >= 1; for d = 0 Horner starts at coefs[0][7] */
        for (int j = 7; j >= 0; --j) {
          double cj = (j + d <= 7) ? (d == 0 ? c0[j] : arrn(d, j + d) * c0[j + d]) : 0.0;
          if (j == 7) v = cj; else { v *= dt; v += cj; }
        }
        Y[2 * d + c] = v;
      }
    }
  }
}

static void cont_dyn(const double X[5], const double U[2], const double W[2], double tau_phi, double tau_v, double d[5]) {
  d[0] = X[4] * cos(X[2]) + W[0];             /* dynamic.py:18-22 */
  d[1] = X[4] * sin(X[2]) + W[1];
  d[2] = G / X[4] * tan(X[3]);
  d[3] = -1 / tau_phi * (X[3] - U[0]);
  d[4] = -1 / tau_v * (X[4] - U[1]);
}

static void rk4(double X[5], const double U[2], const double W[2], double tau_phi, double tau_v, double dt, int nsub) {
  double h = dt / nsub, k1[5], k2[5], k3[5], k4[5], Y[5];
  for (int s = 0; s < nsub; ++s) {
    cont_dyn(X, U, W, tau_phi, tau_v, k1);
    for (int k = 0; k < 5; ++k) Y[k] = X[k] + 0.5 * h * k1[k];
    cont_dyn(Y, U, W, tau_phi, tau_v, k2);
    for (int k = 0; k < 5; ++k) Y[k] = X[k] + 0.5 * h * k2[k];
    cont_dyn(Y, U, W, tau_phi, tau_v, k3);
    for (int k = 0; k < 5; ++k) Y[k] = X[k] + h * k3[k];
    cont_dyn(Y, U, W, tau_phi, tau_v, k4);
    for (int k = 0; k < 5; ++k) X[k] = X[k] + (h / 6.0) * (k1[k] + 2.0 * k2[k] + 2.0 * k3[k] + k4[k]);
  }
  X[2] = wrap_pi(X[2]);
}

/* DFFFController.get, guidance.py:62-91 */
static int control(int type, const double* par, double t, const double X[5], const double W[2], double tau_phi, double tau_v,
                   double U[2], double Xr[5], double K[6]) {
  double Y[8];
  traj_get(type, par, t, Y);
  double vax = Y[2] - W[0], vay = Y[3] - W[1], va2 = vax * vax + vay * vay, va = sqrt(va2);
  double vadot = (vax * Y[4] + vay * Y[5]) / va;
  Xr[0] = Y[0]; Xr[1] = Y[1]; Xr[2] = atan2(vay, vax); Xr[4] = va;
  Xr[3] = atan((Y[5] * vax - Y[4] * vay) / va / 9.81);
  double Ur[2] = {tau_phi * 0.0 + Xr[3], tau_v * vadot + va};
  static const double sat[3] = {20, 20, PI / 3};
  double dX[3] = {X[0] - Xr[0], X[1] - Xr[1], wrap_pi(X[2] - Xr[2])};
  for (int k = 0; k < 3; ++k) dX[k] = clipd(dX[k], -sat[k], sat[k]);
  double sp = sin(Xr[2]), cp = cos(Xr[2]), cphi2 = cos(Xr[3]) * cos(Xr[3]), tphi = tan(Xr[3]);
  double A1[9] = {0, 0, -va * sp, 0, 0, va * cp, 0, 0, 0};            /* dynamic.py:36-38 */
  double B1[6] = {0, cp, 0, sp, G / va / (1 + cphi2), G / (va * va) * tphi};
  static const double q[3] = {1, 1, 0.1}, r[2] = {8, 1};
  int rc = lqr3(A1, B1, q, r, K);
  const double phisat = 45.0 * (PI / 180.0);
  U[0] = clipd(Ur[0] - (K[0] * dX[0] + K[1] * dX[1] + K[2] * dX[2]), -phisat, phisat);
  U[1] = clipd(Ur[1] - (K[3] * dX[0] + K[4] * dX[1] + K[5] * dX[2]), 4, 20);
  return rc;
}

/* run_simulation for B scenarios.  All arrays scenario-major: type[B], par[B][17], wind[B][2], X0[B][5];
 * logs X_log[B][n_rows][5], U_log[B][n_rows][2], K_log[B][n_rows][6] (NULL to skip), rows = samples i with
 * i % log_every == 0.  Returns the number of scenarios whose Riccati solve failed. */
typedef struct {
  int b0, b1, T, nsub, log_every, n_rows, bad;
  const double *time, *par, *wind, *X0;
  const int* type;
  double tau_phi, tau_v;
  double *X_log, *U_log, *K_log, *X_final, *sum_sq, *max_err;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (int b = j->b0; b < j->b1; ++b) {
    double X[5], U[2], Xr[5], K[6], ss = 0.0, mx = 0.0;
    memcpy(X, j->X0 + 5 * (size_t)b, sizeof X);
    const double* W = j->wind + 2 * (size_t)b;
    const double* p = j->par + 17 * (size_t)b;
    int fail = 0;
    for (int i = 0; i < j->T; ++i) {
      fail |= control(j->type[b], p, j->time[i], X, W, j->tau_phi, j->tau_v, U, Xr, K) != 0;
      double ex = X[0] - Xr[0], ey = X[1] - Xr[1], d2 = ex * ex + ey * ey;
      ss += d2; if (d2 > mx) mx = d2;
      if (i % j->log_every == 0) {
        size_t row = (size_t)b * j->n_rows + i / j->log_every;
        if (j->X_log) memcpy(j->X_log + 5 * row, X, sizeof X);
        if (j->U_log) memcpy(j->U_log + 2 * row, U, sizeof U);
        if (j->K_log) memcpy(j->K_log + 6 * row, K, sizeof K);
      }
      if (i == j->T - 1) break;
      rk4(X, U, W, j->tau_phi, j->tau_v, j->time[i + 1] - j->time[i], j->nsub);
    }
    memcpy(j->X_final + 5 * (size_t)b, X, sizeof X);
    if (j->sum_sq) j->sum_sq[b] = ss;
    if (j->max_err) j->max_err[b] = sqrt(mx);
    j->bad += fail;
  }
  return NULL;
}

int orc_max_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }

/* run_simulation for B scenarios on `nthreads` POSIX threads (0 = all online cores).  All arrays
 * scenario-major: type[B], par[B][17], wind[B][2], X0[B][5]; logs X_log[B][n_rows][5], U_log[B][n_rows][2],
 * K_log[B][n_rows][6] (NULL to skip), rows = samples i with i % log_every == 0.  Returns the number of
 * scenarios whose Riccati solve failed. */
int orc_rollout_dfff(int B, int T, const double* time, const int* type, const double* par, const double* wind,
                     const double* X0, double tau_phi, double tau_v, int nsub, int log_every, double* X_log,
                     double* U_log, double* K_log, double* X_final, double* sum_sq, double* max_err, int nthreads) {
  if (nthreads <= 0) nthreads = orc_max_threads();
  if (nthreads > B) nthreads = B;
  if (nthreads > 256) nthreads = 256;
  job_t jobs[256];
  pthread_t th[256];
  for (int t = 0; t < nthreads; ++t) {
    job_t* j = &jobs[t];
    j->b0 = (int)((long)B * t / nthreads); j->b1 = (int)((long)B * (t + 1) / nthreads);
    j->T = T; j->nsub = nsub; j->log_every = log_every; j->n_rows = (T - 1) / log_every + 1; j->bad = 0;
    j->time = time; j->par = par; j->wind = wind; j->X0 = X0; j->type = type; j->tau_phi = tau_phi; j->tau_v = tau_v;
    j->X_log = X_log; j->U_log = U_log; j->K_log = K_log; j->X_final = X_final; j->sum_sq = sum_sq; j->max_err = max_err;
    if (nthreads == 1) worker(j); else pthread_create(&th[t], NULL, worker, j);
  }
  int bad = 0;
  for (int t = 0; t < nthreads; ++t) { if (nthreads > 1) pthread_join(th[t], NULL); bad += jobs[t].bad; }
  return bad;
}

int orc_lqr3(const double* A, const double* B, double* K) {
  static const double q[3] = {1, 1, 0.1}, r[2] = {8, 1};
  return lqr3(A, B, q, r, K);
}

