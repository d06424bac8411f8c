// Single-shooting evaluation of the planner NLP: states from inputs on the backward-Euler grid of the
// collocation constraints (d2d/opty_utils.py:38-50), planner cost (d2d/opty_utils.py:55-165, d2d/multiopty_utils.py:29-174)
// and its exact gradient with respect to the inputs by one adjoint sweep.  One thread = one (problem, aircraft); the
// problem index is the fastest one in every array, so a warp's accesses are coalesced.
#include "d2dx_device.cuh"
#include "d2dx_host.h"

namespace d2dx {

constexpr int kShootThreads = 128;

struct ShootArgs {
  d2dx_colloc_problem p;
  int P;
  const double *u, *p0, *p1, *lam, *rho;
  double *uphys, *xs, *c, *cost, *lagr, *grad;
  int bounded;            // inputs are theta with phi = mid + half sin(theta) (same for v): box constraints by substitution
  double mid[2], half[2];
};

__device__ __forceinline__ bool on(double k) { return (k == k) && k != 0.0; }

__global__ void __launch_bounds__(kShootThreads) shoot_forward_kernel(const __grid_constant__ ShootArgs a) {
  const int N = a.p.N, n_ac = a.p.n_ac;
  const size_t P = a.P;
  const size_t t = (size_t)blockIdx.x * kShootThreads + threadIdx.x;
  if (t >= P * n_ac) return;
  const size_t p = t % P, ac = t / P;
  const double h = a.p.h, wx = a.p.wind[0], wy = a.p.wind[1];
  const size_t o_phi = ((0 * (size_t)n_ac + ac) * N) * P + p, o_v = ((1 * (size_t)n_ac + ac) * N) * P + p;
  const double* phi = a.u + o_phi;
  const double* v = a.u + o_v;
  double* xo = a.xs + ((0 * (size_t)n_ac + ac) * N) * P + p;
  double* yo = a.xs + ((1 * (size_t)n_ac + ac) * N) * P + p;
  double* po = a.xs + ((2 * (size_t)n_ac + ac) * N) * P + p;
  double x = a.p0[(0 * (size_t)n_ac + ac) * P + p], y = a.p0[(1 * (size_t)n_ac + ac) * P + p], psi = a.p0[(2 * (size_t)n_ac + ac) * P + p];
  xo[0] = x; yo[0] = y; po[0] = psi;
  for (int i = 0; i < N; ++i) {
    double ph = phi[(size_t)i * P], vv = v[(size_t)i * P];
    double sp, cp, s, c;
    if (a.bounded) {
      sincos_any(ph, s, c); ph = a.mid[0] + a.half[0] * s;
      sincos_any(vv, s, c); vv = a.mid[1] + a.half[1] * s;
      a.uphys[o_phi + (size_t)i * P] = ph; a.uphys[o_v + (size_t)i * P] = vv;
    }
    if (i == 0) continue;
    sincos_any(ph, sp, cp);
    psi += h * kG * sp * rcp_f(cp * vv);
    sincos_any(psi, s, c);
    x += h * (vv * c - wx);
    y += h * (vv * s - wy);
    xo[(size_t)i * P] = x; yo[(size_t)i * P] = y; po[(size_t)i * P] = psi;
  }
  a.c[(0 * (size_t)n_ac + ac) * P + p] = x - a.p1[(0 * (size_t)n_ac + ac) * P + p];
  a.c[(1 * (size_t)n_ac + ac) * P + p] = y - a.p1[(1 * (size_t)n_ac + ac) * P + p];
  a.c[(2 * (size_t)n_ac + ac) * P + p] = psi - a.p1[(2 * (size_t)n_ac + ac) * P + p];
}

__global__ void __launch_bounds__(kShootThreads) shoot_adjoint_kernel(const __grid_constant__ ShootArgs a) {
  const d2dx_colloc_problem& Q = a.p;
  const int N = Q.N, n_ac = Q.n_ac;
  const size_t P = a.P;
  const size_t t = (size_t)blockIdx.x * kShootThreads + threadIdx.x;
  if (t >= P * n_ac) return;
  const size_t p = t % P;
  const int ac = (int)(t / P);
  const double h = Q.h;
  const double sN = Q.obj_scale / N, norm_in = sN / Q.in_div;
  const bool use_obs = on(Q.kobs) && Q.n_obs > 0 && ac == 0;
  const bool use_col = on(Q.kcol) && n_ac > 1 && (Q.col_all_pairs || ac < 2);
  const double col_kr = Q.kcol_k / Q.rcol;
  const double* uin = a.bounded ? a.uphys : a.u;
  auto U = [&](int k, int b, int i) { return uin[((k * (size_t)n_ac + b) * N + i) * P + p]; };
  auto XS = [&](int k, int b, int i) { return a.xs[((k * (size_t)n_ac + b) * N + i) * P + p]; };
  const double rho = a.rho[p];
  double cterm[3], gl[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    cterm[k] = a.c[(k * (size_t)n_ac + ac) * P + p];
    gl[k] = a.lam[(k * (size_t)n_ac + ac) * P + p] + rho * cterm[k];          // d lagr / d terminal state
  }
  double Gx = gl[0], Gy = gl[1], Gp = gl[2];
  double cost = 0.0;
  for (int i = N - 1; i >= 0; --i) {
    const double x = XS(0, ac, i), y = XS(1, ac, i), psi = XS(2, ac, i), phi = U(0, ac, i), v = U(1, ac, i);
    // direct cost terms at this node and their position gradient
    const double dv = v - Q.vsp;
    cost += norm_in * (Q.kvel * dv * dv + Q.kbank * phi * phi);
    double ax = 0.0, ay = 0.0;
    if (use_obs) {
      for (int o = 0; o < Q.n_obs; ++o) {
        const double dx = x - Q.obs[o][0], dy = y - Q.obs[o][1], r = Q.obs[o][2];
        if (Q.obs_kind == 0) {
          const double raw = exp(r * r - (dx * dx + dy * dy));
          const double es = clip(raw, 0.0, 1e3);
          cost += Q.kobs * sN * es;
          if (raw < 1e3) { ax += Q.kobs * sN * -2.0 * dx * es; ay += Q.kobs * sN * -2.0 * dy * es; }   // flat where the clip is active
        } else {
          const double kr = 2.0 / r, ux = dx * kr, uy = dy * kr;
          const double es = fm::exp_neg(-(ux * ux + uy * uy));
          cost += Q.kobs * sN * es;
          ax += Q.kobs * sN * -2.0 * kr * kr * dx * es; ay += Q.kobs * sN * -2.0 * kr * kr * dy * es;
        }
      }
    }
    if (use_col) {
      const int b_lo = Q.col_all_pairs ? 0 : (ac == 0 ? 1 : 0), b_hi = Q.col_all_pairs ? n_ac : (ac == 0 ? 2 : 1);
      for (int b = b_lo; b < b_hi; ++b) {
        if (b == ac) continue;
        const double dx = x - XS(0, b, i), dy = y - XS(1, b, i);
        const double ux = dx * col_kr, uy = dy * col_kr;
        const double es = fm::exp_neg(-(ux * ux + uy * uy));
        if (ac < b) cost += Q.kcol * sN * es;                                   // each pair counted once
        ax += Q.kcol * sN * -2.0 * col_kr * col_kr * dx * es; ay += Q.kcol * sN * -2.0 * col_kr * col_kr * dy * es;
      }
    }
    double* gphi = a.grad + ((0 * (size_t)n_ac + ac) * N + i) * P + p;
    double* gv = a.grad + ((1 * (size_t)n_ac + ac) * N + i) * P + p;
    const double dphi_cost = norm_in * Q.kbank * 2.0 * phi, dv_cost = norm_in * Q.kvel * 2.0 * dv;
    double jphi = 1.0, jv = 1.0;                                                 // d(phi, v) / d theta
    if (a.bounded) {
      double s_, c_;
      sincos_any(a.u[((0 * (size_t)n_ac + ac) * N + i) * P + p], s_, c_); jphi = a.half[0] * c_;
      sincos_any(a.u[((1 * (size_t)n_ac + ac) * N + i) * P + p], s_, c_); jv = a.half[1] * c_;
    }
    if (i == 0) { *gphi = dphi_cost * jphi; *gv = dv_cost * jv; break; }                    // node 0: fixed state, inputs enter the cost only
    if (i < N - 1) { Gx += ax; Gy += ay; } else { Gx = gl[0] + ax; Gy = gl[1] + ay; }
    double s, c, sp, cp;
    sincos_any(psi, s, c);
    sincos_any(phi, sp, cp);
    Gp += Gx * (-h * v * s) + Gy * (h * v * c);                                  // x_i, y_i depend on psi_i
    const double icv = rcp_f(cp * v);
    *gphi = (Gp * h * kG * icv * rcp_f(cp) + dphi_cost) * jphi;                  // d psi_i / d phi_i = h g / (v cos^2 phi)
    *gv = (Gp * (-h * kG * sp * icv * rcp_f(v)) + Gx * h * c + Gy * h * s + dv_cost) * jv;
  }
  // lagrangian share of this aircraft
  double lg = cost;
#pragma unroll
  for (int k = 0; k < 3; ++k) lg += a.lam[(k * (size_t)n_ac + ac) * P + p] * cterm[k] + 0.5 * rho * cterm[k] * cterm[k];
  atomicAdd(a.cost + p, cost);
  atomicAdd(a.lagr + p, lg);
}

static void set_bounds(ShootArgs& a, const double* b) {
  a.bounded = b != nullptr;
  if (b) { a.mid[0] = 0.5 * (b[0] + b[1]); a.half[0] = 0.5 * (b[1] - b[0]); a.mid[1] = 0.5 * (b[2] + b[3]); a.half[1] = 0.5 * (b[3] - b[2]); }
}

static int check(const d2dx_colloc_problem* p, int P, const char* who) {
  D2DX_CHECK_ARG(p && P >= 1, "%s: null problem or P=%d", who, P);
  D2DX_CHECK_ARG(p->n_ac >= 1 && p->N >= 2 && p->h > 0 && p->in_div >= 1, "%s: n_ac=%d N=%d h=%g in_div=%d", who, p->n_ac, p->N, p->h, p->in_div);
  D2DX_CHECK_ARG(p->n_obs >= 0 && p->n_obs <= D2DX_MAX_OBSTACLES, "%s: n_obs=%d", who, p->n_obs);
  return D2DX_OK;
}

}  // namespace d2dx

using namespace d2dx;

extern "C" int d2dx_shoot_forward(d2dx_handle* h, const d2dx_colloc_problem* p, int32_t P, const double* u, const double* bounds,
                                  const double* p0, const double* p1, double* u_phys, double* xs, double* c, void* stream) {
  if (int rc = check(p, P, "d2dx_shoot_forward")) return rc;
  D2DX_CHECK_ARG(h && u && p0 && p1 && xs && c, "d2dx_shoot_forward: null array");
  D2DX_CHECK_ARG(!bounds || (u_phys && bounds[1] > bounds[0] && bounds[3] > bounds[2]), "d2dx_shoot_forward: bounds need u_phys and lo < hi");
  ShootArgs a = {};
  set_bounds(a, bounds);
  a.p = *p; a.P = P; a.u = u; a.p0 = p0; a.p1 = p1; a.uphys = u_phys; a.xs = xs; a.c = c;
  D2DX_CUDA(cudaSetDevice(h->device));
  const long n = (long)P * p->n_ac;
  shoot_forward_kernel<<<(unsigned)((n + kShootThreads - 1) / kShootThreads), kShootThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("shoot_forward_kernel");
  return D2DX_OK;
}

extern "C" int d2dx_shoot_adjoint(d2dx_handle* h, const d2dx_colloc_problem* p, int32_t P, const double* u, const double* bounds,
                                  const double* u_phys, const double* xs, const double* c, const double* lam, const double* rho,
                                  double* cost, double* lagr, double* grad, void* stream) {
  if (int rc = check(p, P, "d2dx_shoot_adjoint")) return rc;
  D2DX_CHECK_ARG(h && u && xs && c && lam && rho && cost && lagr && grad, "d2dx_shoot_adjoint: null array");
  D2DX_CHECK_ARG(!bounds || u_phys, "d2dx_shoot_adjoint: bounds need u_phys (from d2dx_shoot_forward)");
  ShootArgs a = {};
  set_bounds(a, bounds);
  a.uphys = const_cast<double*>(u_phys);
  a.p = *p; a.P = P; a.u = u; a.xs = const_cast<double*>(xs); a.c = const_cast<double*>(c); a.lam = lam; a.rho = rho; a.cost = cost; a.lagr = lagr; a.grad = grad;
  D2DX_CUDA(cudaSetDevice(h->device));
  D2DX_CUDA(cudaMemsetAsync(cost, 0, sizeof(double) * P, as_stream(stream)));
  D2DX_CUDA(cudaMemsetAsync(lagr, 0, sizeof(double) * P, as_stream(stream)));
  const long n = (long)P * p->n_ac;
  shoot_adjoint_kernel<<<(unsigned)((n + kShootThreads - 1) / kShootThreads), kShootThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("shoot_adjoint_kernel");
  return D2DX_OK;
}
