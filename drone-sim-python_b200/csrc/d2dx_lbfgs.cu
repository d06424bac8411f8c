// Batched augmented-Lagrangian L-BFGS "tick" (the role IPOPT plays behind prob.solve in
// 06_optyplan.py:117-125).  One thread block = one problem; every launch consumes one function evaluation
// (value, gradient, constraint values at the trial point the previous tick wrote) and produces the next trial point,
// so a solve is a fixed sequence [evaluate, tick] with no host decision inside -- it replays from a CUDA graph and
// problems of a population advance independently (own line search, own multiplier updates, own termination).
//
// Per problem state machine:  EVAL0 -> (DIR -> TRIAL ... accept) ... inner convergence -> multiplier update -> EVAL0 ... -> SOLVED
#include "d2dx_device.cuh"
#include "d2dx_host.h"

namespace d2dx {

constexpr int kLbThreads = 256;
enum { LB_EVAL0 = 0, LB_TRIAL = 1, LB_SOLVED = 2, LB_FAILED = 3 };
// scalar slots per problem
enum { SC_F = 0, SC_ALPHA, SC_GAMMA, SC_GD, SC_CPREV, SC_CMAX, SC_COST, SC_NSLOT = 8 };
// integer slots per problem
enum { MI_FLAG = 0, MI_HEAD, MI_CNT, MI_ITS, MI_NLS, MI_NFEV, MI_OUTER, MI_HAVE, MI_STEEP, MI_INNER, MI_TOTAL_ITS, MI_NSLOT = 16 };

struct LbLayout { long x, g, d, S, Y, rh, sc, fh, cacc, meta, total; };

static LbLayout lb_layout(long P, long n, long n_con, long m, long window) {
  LbLayout L; long o = 0;
  L.x = o; o += P * n;  L.g = o; o += P * n;  L.d = o; o += P * n;
  L.S = o; o += P * m * n;  L.Y = o; o += P * m * n;  L.rh = o; o += P * m;
  L.sc = o; o += P * SC_NSLOT;  L.fh = o; o += P * window;  L.cacc = o; o += P * n_con;
  L.meta = o; o += P * MI_NSLOT / 2;  L.total = o;
  return L;
}

struct LbArgs {
  int P, n, n_con, n_parts;
  d2dx_lbfgs_options o;
  LbLayout L;
  double* state;
  double* xt;            // [P][n] trial point (in: evaluated point, out: next point to evaluate)
  const double* fparts;  // [P][n_parts] value shares (summed in order)
  const double* cparts;  // [P][n_parts] cost shares (reported only) or null
  const double* gt;      // [P][n] gradient at xt
  const double* c;       // [P][n_con] constraint values at xt
  double* lam;           // [P][n_con]
  double* rho;           // [P]
  int* n_running;        // device counter: problems not yet SOLVED/FAILED (recomputed every tick)
};

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();                                   // protects `red` from the previous reduction's readers
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < kLbThreads / 32; ++k) t += red[k];
  return t;
}
__device__ __forceinline__ double block_max(double v, double* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = red[0];
#pragma unroll
  for (int k = 1; k < kLbThreads / 32; ++k) t = fmax(t, red[k]);
  return t;
}

__global__ void lbfgs_count_reset(int* n_running) { *n_running = 0; }

__global__ void lbfgs_init_kernel(int P, double* sc, double* rho, double rho0) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  sc[(size_t)p * SC_NSLOT + SC_CPREV] = -1.0;                // no previous constraint violation yet
  sc[(size_t)p * SC_NSLOT + SC_GAMMA] = 1.0;
  rho[p] = rho0;
}

__global__ void __launch_bounds__(kLbThreads) al_lbfgs_tick_kernel(const __grid_constant__ LbArgs a) {
  __shared__ double red[kLbThreads / 32];
  __shared__ double alpha_j[64];
  const int p = blockIdx.x, tid = threadIdx.x, n = a.n, m = a.o.m;
  double* st = a.state;
  double* x = st + a.L.x + (size_t)p * n;
  double* g = st + a.L.g + (size_t)p * n;
  double* d = st + a.L.d + (size_t)p * n;
  double* S = st + a.L.S + (size_t)p * m * n;
  double* Y = st + a.L.Y + (size_t)p * m * n;
  double* rh = st + a.L.rh + (size_t)p * m;
  double* sc = st + a.L.sc + (size_t)p * SC_NSLOT;
  double* fh = st + a.L.fh + (size_t)p * a.o.window;
  double* cacc = st + a.L.cacc + (size_t)p * a.n_con;
  int* mi = reinterpret_cast<int*>(st + a.L.meta) + (size_t)p * MI_NSLOT;
  double* xt = a.xt + (size_t)p * n;
  const double* gt = a.gt + (size_t)p * n;
  const double* c = a.c + (size_t)p * a.n_con;
  double* lam = a.lam + (size_t)p * a.n_con;

  int flag = mi[MI_FLAG];
  if (flag == LB_SOLVED || flag == LB_FAILED) return;        // xt already holds the solution
  double ft = 0.0, costt = 0.0;
  for (int k = 0; k < a.n_parts; ++k) { ft += a.fparts[(size_t)p * a.n_parts + k]; if (a.cparts) costt += a.cparts[(size_t)p * a.n_parts + k]; }
  int head = mi[MI_HEAD], cnt = mi[MI_CNT], have = mi[MI_HAVE], nls = mi[MI_NLS], inner = mi[MI_INNER];
  const int flag_in = flag, outer_in = mi[MI_OUTER], steep_in = mi[MI_STEEP];
  double f = sc[SC_F], alpha = sc[SC_ALPHA], gamma = sc[SC_GAMMA], gd = sc[SC_GD];
  bool accepted = false, new_dir = false;
  __syncthreads();                                           // everyone has read the integer state before thread 0 rewrites it

  if (flag == LB_EVAL0) {                                    // first evaluation of this (lam, rho): take it as the current point
    for (int i = tid; i < n; i += kLbThreads) { x[i] = xt[i]; g[i] = gt[i]; }
    f = ft; accepted = true; inner = 0;
  } else {                                                   // LB_TRIAL: Armijo test of xt = x + alpha d
    const bool ok = (ft <= f + 1e-4 * alpha * gd + 1e-15 * fabs(f)) && isfinite(ft);
    if (ok) {
      double sy = 0, ss = 0, yy = 0;
      for (int i = tid; i < n; i += kLbThreads) {
        const double s_ = xt[i] - x[i], y_ = gt[i] - g[i];
        sy += s_ * y_; ss += s_ * s_; yy += y_ * y_;
      }
      sy = block_sum(sy, red); ss = block_sum(ss, red); yy = block_sum(yy, red);
      const bool good = sy > 1e-10 * sqrt(ss * yy);
      double* Sh = S + (size_t)head * n; double* Yh = Y + (size_t)head * n;
      for (int i = tid; i < n; i += kLbThreads) {
        const double xi = xt[i], gi = gt[i];
        if (good) { Sh[i] = xi - x[i]; Yh[i] = gi - g[i]; }
        x[i] = xi; g[i] = gi;
      }
      if (good) { if (tid == 0) rh[head] = 1.0 / sy; gamma = sy / yy; have = 1; head = (head + 1) % m; cnt = min(cnt + 1, m); }
      f = ft; accepted = true; inner += 1;
    } else {
      nls += 1;
      if (nls >= a.o.ls_max) {                               // line search failed
        if (steep_in || !have) { flag = LB_FAILED; }     // even steepest descent cannot decrease: at the noise floor -> inner done
        cnt = 0; have = 0; head = 0; new_dir = true;         // drop the history, try steepest descent from x
      } else {
        alpha *= 0.5;
        for (int i = tid; i < n; i += kLbThreads) xt[i] = x[i] + alpha * d[i];
      }
    }
  }

  bool inner_done = false;
  if (accepted) {
    for (int k = tid; k < a.n_con; k += kLbThreads) cacc[k] = c[k];
    if (tid == 0) sc[SC_COST] = costt;
    double gn = 0.0;
    for (int i = tid; i < n; i += kLbThreads) gn = fmax(gn, fabs(g[i]));
    gn = block_max(gn, red);
    // windowed decrease test (ring of the last `window` accepted values)
    const int wdw = a.o.window;
    bool flat = false;
    if (inner >= wdw) flat = (fh[inner % wdw] - f) <= a.o.ftol * fmax(1.0, fabs(f));
    __syncthreads();
    if (tid == 0) fh[inner % wdw] = f;
    inner_done = gn <= a.o.gtol || flat || inner >= a.o.max_inner;
    new_dir = !inner_done;
    nls = 0;
  }
  if (flag == LB_FAILED) { inner_done = true; new_dir = false; flag = LB_TRIAL; }

  if (inner_done) {                                          // multiplier update (or termination) at the accepted point x
    __syncthreads();
    double cm = 0.0;
    for (int k = tid; k < a.n_con; k += kLbThreads) cm = fmax(cm, fabs(cacc[k]));
    cm = block_max(cm, red);
    const int outer = outer_in + 1;
    if (cm < a.o.ctol || outer >= a.o.max_outer) {
      flag = cm < a.o.ctol ? LB_SOLVED : LB_FAILED;
      for (int i = tid; i < n; i += kLbThreads) xt[i] = x[i];
    } else {
      const double rho = a.rho[p], cprev = sc[SC_CPREV];
      for (int k = tid; k < a.n_con; k += kLbThreads) lam[k] += rho * cacc[k];
      __syncthreads();
      if (tid == 0) {
        if (cprev < 0.0 || cm > 0.25 * cprev) a.rho[p] = fmin(rho * 3.0, a.o.rho_max);
        sc[SC_CPREV] = cm;
      }
      for (int i = tid; i < n; i += kLbThreads) xt[i] = x[i];
      flag = LB_EVAL0; cnt = 0; have = 0; head = 0;
    }
    if (tid == 0) { sc[SC_CMAX] = cm; mi[MI_OUTER] = outer; }
  }

  int steep = 0;
  if (new_dir) {                                             // two-loop recursion on d (thread i owns d[i], d[i + 256], ...)
    __syncthreads();
    for (int i = tid; i < n; i += kLbThreads) d[i] = g[i];
    for (int j = 0; j < cnt; ++j) {
      const int k = (head - 1 - j + 2 * m) % m;
      const double* Sk = S + (size_t)k * n; const double* Yk = Y + (size_t)k * n;
      double t = 0.0;
      for (int i = tid; i < n; i += kLbThreads) t += Sk[i] * d[i];
      t = rh[k] * block_sum(t, red);
      if (tid == 0) alpha_j[j] = t;
      for (int i = tid; i < n; i += kLbThreads) d[i] -= t * Yk[i];
    }
    for (int i = tid; i < n; i += kLbThreads) d[i] *= gamma;
    __syncthreads();
    for (int j = cnt - 1; j >= 0; --j) {
      const int k = (head - 1 - j + 2 * m) % m;
      const double* Sk = S + (size_t)k * n; const double* Yk = Y + (size_t)k * n;
      double t = 0.0;
      for (int i = tid; i < n; i += kLbThreads) t += Yk[i] * d[i];
      t = alpha_j[j] - rh[k] * block_sum(t, red);
      for (int i = tid; i < n; i += kLbThreads) d[i] += t * Sk[i];
    }
    double t = 0.0, g1 = 0.0, gm = 0.0;
    for (int i = tid; i < n; i += kLbThreads) { t += g[i] * d[i]; g1 += fabs(g[i]); gm = fmax(gm, fabs(g[i])); }
    t = block_sum(t, red); g1 = block_sum(g1, red); gm = block_max(gm, red);
    gd = -t;                                                  // d currently holds +H g
    double scale = -1.0;
    if (!(gd < -1e-14 * gm * gm) || !have) { steep = 1; scale = -(have ? gamma : 1.0 / fmax(g1, 1e-300)); }
    if (steep) {
      double gg = 0.0;
      for (int i = tid; i < n; i += kLbThreads) { d[i] = scale * g[i]; gg += g[i] * g[i]; }
      gd = scale * block_sum(gg, red);
    } else {
      for (int i = tid; i < n; i += kLbThreads) d[i] = -d[i];
    }
    alpha = 1.0; nls = 0;
    for (int i = tid; i < n; i += kLbThreads) xt[i] = x[i] + d[i];
    flag = LB_TRIAL;
  }

  if (tid == 0) {
    mi[MI_FLAG] = flag; mi[MI_HEAD] = head; mi[MI_CNT] = cnt; mi[MI_HAVE] = have; mi[MI_NLS] = nls; mi[MI_INNER] = inner;
    if (new_dir) mi[MI_STEEP] = steep;
    mi[MI_NFEV] += 1;
    if (accepted && flag_in == LB_TRIAL) mi[MI_TOTAL_ITS] += 1;
    sc[SC_F] = f; sc[SC_ALPHA] = alpha; sc[SC_GAMMA] = gamma; sc[SC_GD] = gd;
    if (flag != LB_SOLVED && flag != LB_FAILED) atomicAdd(a.n_running, 1);
  }
}

}  // namespace d2dx

using namespace d2dx;

static int lb_check(int P, int n, int n_con, const d2dx_lbfgs_options* o, const char* who) {
  D2DX_CHECK_ARG(o && P >= 1 && n >= 1 && n_con >= 0, "%s: P=%d n=%d n_con=%d", who, P, n, n_con);
  D2DX_CHECK_ARG(o->m >= 1 && o->m <= 64 && o->window >= 1 && o->ls_max >= 1 && o->max_inner >= 1 && o->max_outer >= 1,
                 "%s: m=%d (1..64) window=%d ls_max=%d max_inner=%d max_outer=%d", who, o->m, o->window, o->ls_max, o->max_inner, o->max_outer);
  return D2DX_OK;
}

extern "C" int d2dx_lbfgs_layout(int32_t P, int32_t n, int32_t n_con, const d2dx_lbfgs_options* o, int64_t offsets[8]) {
  if (int rc = lb_check(P, n, n_con, o, "d2dx_lbfgs_layout")) return rc;
  D2DX_CHECK_ARG(offsets, "d2dx_lbfgs_layout: null offsets");
  const LbLayout L = lb_layout(P, n, n_con, o->m, o->window);
  offsets[0] = L.total; offsets[1] = L.x; offsets[2] = L.g; offsets[3] = L.sc; offsets[4] = L.cacc; offsets[5] = L.meta;
  offsets[6] = SC_NSLOT; offsets[7] = MI_NSLOT;
  return D2DX_OK;
}

extern "C" int d2dx_lbfgs_init(d2dx_handle* h, int32_t P, int32_t n, int32_t n_con, const d2dx_lbfgs_options* o, double* state,
                               double* lam, double* rho, void* stream) {
  if (int rc = lb_check(P, n, n_con, o, "d2dx_lbfgs_init")) return rc;
  D2DX_CHECK_ARG(h && state && rho && (lam || n_con == 0), "d2dx_lbfgs_init: null array");
  const LbLayout L = lb_layout(P, n, n_con, o->m, o->window);
  D2DX_CUDA(cudaSetDevice(h->device));
  // integer and scalar state to zero (flag = EVAL0), c_prev = -1 (none yet), multipliers 0, rho = rho0
  D2DX_CUDA(cudaMemsetAsync(state + L.rh, 0, sizeof(double) * (L.total - L.rh), as_stream(stream)));
  lbfgs_init_kernel<<<(P + 127) / 128, 128, 0, as_stream(stream)>>>(P, state + L.sc, rho, o->rho0);
  D2DX_LAUNCH_CHECK("lbfgs_init_kernel");
  if (n_con) D2DX_CUDA(cudaMemsetAsync(lam, 0, sizeof(double) * (size_t)P * n_con, as_stream(stream)));
  return D2DX_OK;
}

extern "C" int d2dx_al_lbfgs_tick(d2dx_handle* h, int32_t P, int32_t n, int32_t n_con, const d2dx_lbfgs_options* o, double* state,
                                  double* x_trial, const double* f_parts, const double* cost_parts, int32_t n_parts, const double* grad,
                                  const double* c, double* lam, double* rho, int32_t* n_running, void* stream) {
  if (int rc = lb_check(P, n, n_con, o, "d2dx_al_lbfgs_tick")) return rc;
  D2DX_CHECK_ARG(h && state && x_trial && f_parts && grad && rho && n_running && n_parts >= 1 && (n_con == 0 || (c && lam)),
                 "d2dx_al_lbfgs_tick: null array or n_parts=%d", n_parts);
  LbArgs a = {};
  a.P = P; a.n = n; a.n_con = n_con; a.n_parts = n_parts; a.o = *o; a.L = lb_layout(P, n, n_con, o->m, o->window);
  a.state = state; a.xt = x_trial; a.fparts = f_parts; a.cparts = cost_parts; a.gt = grad; a.c = c; a.lam = lam; a.rho = rho;
  a.n_running = n_running;
  D2DX_CUDA(cudaSetDevice(h->device));
  lbfgs_count_reset<<<1, 1, 0, as_stream(stream)>>>(n_running);
  al_lbfgs_tick_kernel<<<P, kLbThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("al_lbfgs_tick_kernel");
  return D2DX_OK;
}
