// Single-shooting evaluation of the planner NLP: states from inputs on the backward-Euler grid of the
// collocation constraints (d2d/opty_utils.py:38-50), planner cost (d2d/opty_utils.py:55-165, d2d/multiopty_utils.py:29-174)
// and its exact gradient with respect to the inputs by one adjoint sweep.
// One warp = one (problem, aircraft); lanes own consecutive nodes (coalesced rows of 32), and the recursions
//   psi_i = psi_{i-1} + dpsi_i,  x_i = x_{i-1} + dx_i(psi_i),  y_i = ...        (forward: prefix sums)
//   Gx_i = Gx_{i+1} + ax_i,  Gpsi_i = Gpsi_{i+1} + m_i(Gx_i, Gy_i)               (adjoint: suffix sums)
// are warp-shuffle scans with a carry between rows, so a 1001-node problem costs 32 row steps instead of 1000 serial ones.
#include "d2dx_device.cuh"
#include "d2dx_host.h"

namespace d2dx {

constexpr int kShootThreads = 128;
constexpr int kShootWarps = kShootThreads / 32;

struct ShootArgs {
  d2dx_colloc_problem p;
  int P;
  const double *u, *p0, *p1, *lam, *rho;
  double *uphys, *xs, *c, *cost, *lagr, *grad;
  int boxed;              // soft state box: weight/N * sum_i (excess of x_i, y_i over [lo, hi])^2 added to the cost
  double box[5];          // x_lo, x_hi, y_lo, y_hi, weight
  int bounded;            // inputs are theta with phi = mid + half sin(theta) (same for v): box constraints by substitution
  double mid[2], half[2];
};

__host__ __device__ __forceinline__ bool on(double k) { return (k == k) && k != 0.0; }

// inclusive prefix sum over the warp (lane order)
__device__ __forceinline__ double warp_prefix(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// inclusive suffix sum over the warp (lane l gets sum over lanes >= l)
__device__ __forceinline__ double warp_suffix(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

// One unit = one (problem, aircraft), run by W threads: W = 32 (a warp; kShootWarps units per block) when there are
// enough units to fill the machine, W = kShootWide (a whole block; scans go through shared memory) for few, long problems.
constexpr int kShootWide = 256;

template <int W> struct UnitScan {
  // inclusive prefix / suffix over the unit's threads; `total` = sum over the unit (same in every thread)
  static __device__ __forceinline__ double prefix(double v, double& total, double* red) {
    const int lane = threadIdx.x & 31;
    const double w = warp_prefix(v, lane);
    if (W == 32) { total = __shfl_sync(0xffffffffu, w, 31); return w; }
    const int wid = threadIdx.x >> 5;
    __syncthreads();                                 // the previous scan's readers are done with `red`
    if (lane == 31) red[wid] = w;
    __syncthreads();
    double off = 0.0, tot = 0.0;
#pragma unroll
    for (int k = 0; k < W / 32; ++k) { const double r = red[k]; tot += r; if (k < wid) off += r; }
    total = tot;
    return off + w;
  }
  static __device__ __forceinline__ double suffix(double v, double& total, double* red) {
    const int lane = threadIdx.x & 31;
    const double w = warp_suffix(v, lane);
    if (W == 32) { total = __shfl_sync(0xffffffffu, w, 0); return w; }
    const int wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = w;
    __syncthreads();
    double off = 0.0, tot = 0.0;
#pragma unroll
    for (int k = 0; k < W / 32; ++k) { const double r = red[k]; tot += r; if (k > wid) off += r; }
    total = tot;
    return off + w;
  }
};

template <int W>
__global__ void __launch_bounds__(W == 32 ? kShootThreads : W) shoot_forward_kernel(const __grid_constant__ ShootArgs a) {
  __shared__ double red[3][kShootWide / 32];
  const int N = a.p.N, n_ac = a.p.n_ac, lane = W == 32 ? (threadIdx.x & 31) : threadIdx.x;
  const long w = W == 32 ? (long)blockIdx.x * kShootWarps + (threadIdx.x >> 5) : (long)blockIdx.x;
  if (w >= (long)a.P * n_ac) return;
  const long p = w / n_ac;
  const int ac = (int)(w % n_ac);
  const double h = a.p.h, wx = a.p.wind[0], wy = a.p.wind[1];
  const size_t ou = (size_t)p * 2 * n_ac * N, ox = (size_t)p * 3 * n_ac * N, ob = ((size_t)p * 3) * n_ac;
  const double* phi_in = a.u + ou + (size_t)ac * N;
  const double* v_in = a.u + ou + (size_t)(n_ac + ac) * N;
  double* xo = a.xs + ox + (size_t)(0 * n_ac + ac) * N;
  double* yo = a.xs + ox + (size_t)(1 * n_ac + ac) * N;
  double* po = a.xs + ox + (size_t)(2 * n_ac + ac) * N;
  double cx = a.p0[ob + 0 * n_ac + ac], cy = a.p0[ob + 1 * n_ac + ac], cpsi = a.p0[ob + 2 * n_ac + ac];   // carries
  for (int base = 0; base < N; base += W) {
    const int i = base + lane;
    const bool valid = i < N;
    double ph = 0.0, vv = 1.0, s, c, tot;
    if (valid) {
      ph = phi_in[i]; vv = v_in[i];
      if (a.bounded) {
        sincos_any(ph, s, c); ph = a.mid[0] + a.half[0] * s;
        sincos_any(vv, s, c); vv = a.mid[1] + a.half[1] * s;
        a.uphys[ou + (size_t)ac * N + i] = ph; a.uphys[ou + (size_t)(n_ac + ac) * N + i] = vv;
      }
    }
    const bool moves = valid && i > 0;                       // node 0 is the initial state
    double sp, cp;
    sincos_any(ph, sp, cp);
    const double dpsi = moves ? h * kG * sp * rcp_f(cp * vv) : 0.0;
    const double psi = cpsi + UnitScan<W>::prefix(dpsi, tot, red[0]);
    cpsi += tot;
    sincos_any(psi, s, c);
    const double x = cx + UnitScan<W>::prefix(moves ? h * (vv * c - wx) : 0.0, tot, red[1]);
    cx += tot;
    const double y = cy + UnitScan<W>::prefix(moves ? h * (vv * s - wy) : 0.0, tot, red[2]);
    cy += tot;
    if (valid) { xo[i] = x; yo[i] = y; po[i] = psi; }
  }
  if (lane == 0) {
    a.c[ob + 0 * n_ac + ac] = cx - a.p1[ob + 0 * n_ac + ac];
    a.c[ob + 1 * n_ac + ac] = cy - a.p1[ob + 1 * n_ac + ac];
    a.c[ob + 2 * n_ac + ac] = cpsi - a.p1[ob + 2 * n_ac + ac];
  }
}

// EXTRA = obstacle and/or collision terms present; the input-cost-only instance keeps the position-gradient code out
template <bool EXTRA, int W>
__global__ void __launch_bounds__(W == 32 ? kShootThreads : W, W == 32 ? (EXTRA ? 4 : 8) : 1) shoot_adjoint_kernel(const __grid_constant__ ShootArgs a) {
  __shared__ double red[3][kShootWide / 32];
  const d2dx_colloc_problem& Q = a.p;
  const int N = Q.N, n_ac = Q.n_ac, lane = W == 32 ? (threadIdx.x & 31) : threadIdx.x;
  const long w = W == 32 ? (long)blockIdx.x * kShootWarps + (threadIdx.x >> 5) : (long)blockIdx.x;
  if (w >= (long)a.P * n_ac) return;
  const long p = w / n_ac;
  const int ac = (int)(w % n_ac);
  const double h = Q.h;
  const double sN = Q.obj_scale / N, norm_in = sN / Q.in_div;
  const bool use_obs = EXTRA && on(Q.kobs) && Q.n_obs > 0 && ac == 0;
  const bool use_col = EXTRA && on(Q.kcol) && n_ac > 1 && (Q.col_all_pairs || ac < 2);
  const int b_lo = Q.col_all_pairs ? 0 : (ac == 0 ? 1 : 0), b_hi = Q.col_all_pairs ? n_ac : (ac == 0 ? 2 : 1);
  const double col_kr = Q.kcol_k / Q.rcol;
  const size_t ou = (size_t)p * 2 * n_ac * N, ox = (size_t)p * 3 * n_ac * N, ob = ((size_t)p * 3) * n_ac;
  const double* uin = (a.bounded ? a.uphys : a.u) + ou;
  const double* xs = a.xs + ox;
  const double rho = a.rho[p];
  double cterm[3], gl[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    cterm[k] = a.c[ob + k * n_ac + ac];
    gl[k] = a.lam[ob + k * n_ac + ac] + rho * cterm[k];                         // d lagr / d terminal state
  }
  double Gx_c = gl[0], Gy_c = gl[1], Gp_c = gl[2];                              // carries: sums over the rows already done
  double cost = 0.0;
  for (int base = ((N - 1) / W) * W; base >= 0; base -= W) {
    const int i = base + lane;
    double tot;
    const bool valid = i < N, moves = valid && i > 0;
    double x = 0, y = 0, psi = 0, phi = 0, v = 1;
    if (valid) {
      x = xs[(size_t)(0 * n_ac + ac) * N + i]; y = xs[(size_t)(1 * n_ac + ac) * N + i]; psi = xs[(size_t)(2 * n_ac + ac) * N + i];
      phi = uin[(size_t)ac * N + i]; v = uin[(size_t)(n_ac + ac) * N + i];
    }
    const double dv = v - Q.vsp;
    double ax = 0.0, ay = 0.0;
    if (valid) {
      cost += norm_in * (Q.kvel * dv * dv + Q.kbank * phi * phi);
      if constexpr (EXTRA) {
      if (a.boxed) {                                                             // x/y_constraint of the planners as a soft box
        const double wN = a.box[4] * sN;
        const double ex = x > a.box[1] ? x - a.box[1] : (x < a.box[0] ? x - a.box[0] : 0.0);
        const double ey = y > a.box[3] ? y - a.box[3] : (y < a.box[2] ? y - a.box[2] : 0.0);
        cost += wN * (ex * ex + ey * ey);
        ax += 2.0 * wN * ex; ay += 2.0 * wN * ey;
      }
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
        for (int b = b_lo; b < b_hi; ++b) {
          if (b == ac) continue;
          const double dx = x - xs[(size_t)(0 * n_ac + b) * N + i], dy = y - xs[(size_t)(1 * n_ac + b) * N + i];
          const double ux = dx * col_kr, uy = dy * col_kr;
          const double es = fm::exp_neg(-(ux * ux + uy * uy));
          if (ac < b) cost += Q.kcol * sN * es;                                 // each pair counted once
          ax += Q.kcol * sN * -2.0 * col_kr * col_kr * dx * es; ay += Q.kcol * sN * -2.0 * col_kr * col_kr * dy * es;
        }
      }
      }
    }
    if (!moves) { ax = 0.0; ay = 0.0; }                                         // node 0 is fixed
    double Gx = Gx_c, Gy = Gy_c;                                                // without position costs Gx, Gy are the terminal multipliers
    if constexpr (EXTRA) {
      Gx += UnitScan<W>::suffix(ax, tot, red[0]); Gx_c += tot;
      Gy += UnitScan<W>::suffix(ay, tot, red[1]); Gy_c += tot;
    }
    double s, c, sp, cp;
    sincos_any(psi, s, c);
    sincos_any(phi, sp, cp);
    const double m = moves ? Gx * (-h * v * s) + Gy * (h * v * c) : 0.0;        // x_i, y_i depend on psi_i
    const double Gp = Gp_c + UnitScan<W>::suffix(m, tot, red[2]);
    Gp_c += tot;
    if (valid) {
      double jphi = 1.0, jv = 1.0;                                              // d(phi, v) / d theta
      if (a.bounded) {
        double s_, c_;
        sincos_any(a.u[ou + (size_t)ac * N + i], s_, c_); jphi = a.half[0] * c_;
        sincos_any(a.u[ou + (size_t)(n_ac + ac) * N + i], s_, c_); jv = a.half[1] * c_;
      }
      const double dphi_cost = norm_in * Q.kbank * 2.0 * phi, dv_cost = norm_in * Q.kvel * 2.0 * dv;
      double gphi = dphi_cost, gv = dv_cost;
      if (moves) {
        const double icv = rcp_f(cp * v);
        gphi += Gp * h * kG * icv * rcp_f(cp);                                  // d psi_i / d phi_i = h g / (v cos^2 phi)
        gv += Gp * (-h * kG * sp * icv * rcp_f(v)) + Gx * h * c + Gy * h * s;
      }
      a.grad[ou + (size_t)ac * N + i] = gphi * jphi;
      a.grad[ou + (size_t)(n_ac + ac) * N + i] = gv * jv;
    }
  }
  cost = warp_sum(cost);
  if (W != 32) {                                                                // unit = block: add the warps' shares in order
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = cost;
    __syncthreads();
    cost = 0.0;
#pragma unroll
    for (int k = 0; k < W / 32; ++k) cost += red[0][k];
  }
  if (lane == 0) {
    double lg = cost;
#pragma unroll
    for (int k = 0; k < 3; ++k) lg += a.lam[ob + k * n_ac + ac] * cterm[k] + 0.5 * rho * cterm[k] * cterm[k];
    a.cost[(size_t)p * n_ac + ac] = cost;                                       // per-aircraft shares: the caller sums them in order
    a.lagr[(size_t)p * n_ac + ac] = lg;
  }
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

static unsigned shoot_grid(const d2dx_colloc_problem* p, int P) { return (unsigned)(((long)P * p->n_ac + kShootWarps - 1) / kShootWarps); }
// few, long units: one block each (latency of one evaluation of a single problem: 32 serial row steps -> 4)
static bool shoot_wide(const d2dx_colloc_problem* p, int P) { return (long)P * p->n_ac <= 296 && p->N > 96; }

}  // namespace d2dx

using namespace d2dx;

extern "C" int d2dx_shoot_forward(d2dx_handle* h, const d2dx_colloc_problem* p, int32_t P, const double* u, const double* bounds,
                                  const double* p0, const double* p1, double* u_phys, double* xs, double* c, void* stream) {
  D2DX_NVTX("d2dx_shoot_forward");
  if (int rc = check(p, P, "d2dx_shoot_forward")) return rc;
  D2DX_CHECK_ARG(h && u && p0 && p1 && xs && c, "d2dx_shoot_forward: null array");
  D2DX_CHECK_ARG(!bounds || (u_phys && bounds[1] > bounds[0] && bounds[3] > bounds[2]), "d2dx_shoot_forward: bounds need u_phys and lo < hi");
  ShootArgs a = {};
  set_bounds(a, bounds);
  a.p = *p; a.P = P; a.u = u; a.p0 = p0; a.p1 = p1; a.uphys = u_phys; a.xs = xs; a.c = c;
  D2DX_CUDA(cudaSetDevice(h->device));
  if (shoot_wide(p, P)) shoot_forward_kernel<kShootWide><<<(unsigned)((long)P * p->n_ac), kShootWide, 0, as_stream(stream)>>>(a);
  else shoot_forward_kernel<32><<<shoot_grid(p, P), kShootThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("shoot_forward_kernel");
  return D2DX_OK;
}

extern "C" int d2dx_shoot_adjoint(d2dx_handle* h, const d2dx_colloc_problem* p, int32_t P, const double* u, const double* bounds,
                                  const double* state_box, const double* u_phys, const double* xs, const double* c, const double* lam,
                                  const double* rho, double* cost, double* lagr, double* grad, void* stream) {
  D2DX_NVTX("d2dx_shoot_adjoint");
  if (int rc = check(p, P, "d2dx_shoot_adjoint")) return rc;
  D2DX_CHECK_ARG(h && u && xs && c && lam && rho && cost && lagr && grad, "d2dx_shoot_adjoint: null array");
  D2DX_CHECK_ARG(!bounds || u_phys, "d2dx_shoot_adjoint: bounds need u_phys (from d2dx_shoot_forward)");
  ShootArgs a = {};
  set_bounds(a, bounds);
  a.uphys = const_cast<double*>(u_phys);
  a.p = *p; a.P = P; a.u = u; a.xs = const_cast<double*>(xs); a.c = const_cast<double*>(c); a.lam = lam;
  a.rho = rho; a.cost = cost; a.lagr = lagr; a.grad = grad;
  D2DX_CUDA(cudaSetDevice(h->device));
  a.boxed = state_box != nullptr;
  if (state_box) {
    D2DX_CHECK_ARG(state_box[1] >= state_box[0] && state_box[3] >= state_box[2] && state_box[4] >= 0.0, "d2dx_shoot_adjoint: state_box lo > hi or weight < 0");
    for (int k = 0; k < 5; ++k) a.box[k] = state_box[k];
  }
  const bool extra = a.boxed || (on(p->kobs) && p->n_obs > 0) || (on(p->kcol) && p->n_ac > 1);
  if (shoot_wide(p, P)) {
    const unsigned grid = (unsigned)((long)P * p->n_ac);
    if (extra) shoot_adjoint_kernel<true, kShootWide><<<grid, kShootWide, 0, as_stream(stream)>>>(a);
    else shoot_adjoint_kernel<false, kShootWide><<<grid, kShootWide, 0, as_stream(stream)>>>(a);
  } else {
    if (extra) shoot_adjoint_kernel<true, 32><<<shoot_grid(p, P), kShootThreads, 0, as_stream(stream)>>>(a);
    else shoot_adjoint_kernel<false, 32><<<shoot_grid(p, P), kShootThreads, 0, as_stream(stream)>>>(a);
  }
  D2DX_LAUNCH_CHECK("shoot_adjoint_kernel");
  return D2DX_OK;
}
