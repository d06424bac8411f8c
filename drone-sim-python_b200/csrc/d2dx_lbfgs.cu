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

constexpr int kLbWide = 256, kLbWarpMaxN = 512;     // problems with n <= kLbWarpMaxN are run by one warp each
enum { LB_EVAL0 = 0, LB_TRIAL = 1, LB_SOLVED = 2, LB_FAILED = 3 };
// scalar slots per problem
enum { SC_F = 0, SC_ALPHA, SC_GAMMA, SC_GD, SC_CPREV, SC_CMAX, SC_COST, SC_GN, SC_G1, SC_NSLOT = 12 };
// integer slots per problem
enum { MI_FLAG = 0, MI_HEAD, MI_CNT, MI_ITS, MI_NLS, MI_NFEV, MI_OUTER, MI_HAVE, MI_STEEP, MI_INNER, MI_TOTAL_ITS, MI_NSLOT = 16 };

struct LbLayout { long x, g, d, S, Y, G, rh, sc, fh, cacc, meta, total; };

static LbLayout lb_layout(long P, long n, long n_con, long m, long window) {
  LbLayout L; long o = 0;
  L.x = o; o += P * n;  L.g = o; o += P * n;  L.d = o; o += P * n;
  L.S = o; o += P * m * n;  L.Y = o; o += P * m * n;  L.G = o; o += P * (2 * m + 1) * (2 * m + 1);  L.rh = o; o += P * m;
  L.sc = o; o += P * SC_NSLOT;  L.fh = o; o += P * window;  L.cacc = o; o += P * n_con;
  L.meta = o; o += P * MI_NSLOT / 2;  L.total = o;
  return L;
}

struct LbArgs {
  int P, n, n_con, n_parts;
  int gram;              // direction by Gram-matrix coefficients (two dependency-free passes: latency) or the classic two-loop (least work)
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

// T threads cooperate on one problem: T = 32 (one warp, no block barriers: small n) or T = 256
template <int T> __device__ __forceinline__ void team_sync() { if (T == 32) __syncwarp(); else __syncthreads(); }

template <int T> __device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  if (T == 32) return v;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();                                   // protects `red` from the previous reduction's readers
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < T / 32; ++k) t += red[k];
  return t;
}
template <int T> __device__ __forceinline__ double block_max(double v, double* red) {
  v = warp_max(v);
  if (T == 32) return v;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = red[0];
#pragma unroll
  for (int k = 1; k < T / 32; ++k) t = fmax(t, red[k]);
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

// dynamic shared memory (doubles): al[m] | dl[2m+1] | Gs[(2m+1)^2] | part[T/32][3(2m+1)]  (classic form: al only)
static size_t lb_smem_bytes(int m, int T, int gram) {
  const int B = 2 * m + 1;
  return sizeof(double) * (gram ? (size_t)B * B + (size_t)(T / 32) * 3 * B + B + m : (size_t)m);
}

template <int kLbThreads>
__global__ void __launch_bounds__(kLbThreads) al_lbfgs_tick_kernel(const __grid_constant__ LbArgs a) {
  extern __shared__ double smem[];
  __shared__ double red[kLbThreads / 32 + 1];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, n = a.n, m = a.o.m, B = 2 * m + 1;
  double* al = smem;                                         // two-loop alphas (both forms); the rest only in the Gram form
  double* dl = al + m;                                       // direction coefficients in the basis {s_k, y_k, g}: index k, m + k, 2m
  double* Gs = dl + B;                                       // Gram matrix of that basis
  double* part = Gs + B * B;
  double* st = a.state;
  double* x = st + a.L.x + (size_t)p * n;
  double* g = st + a.L.g + (size_t)p * n;
  double* d = st + a.L.d + (size_t)p * n;
  double* S = st + a.L.S + (size_t)p * m * n;
  double* Y = st + a.L.Y + (size_t)p * m * n;
  double* Gg = st + a.L.G + (size_t)p * B * B;
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
  double f = sc[SC_F], alpha = sc[SC_ALPHA], gamma = sc[SC_GAMMA], gd = sc[SC_GD], gn = sc[SC_GN], g1 = sc[SC_G1];
  bool accepted = false, new_dir = false, new_pair = false;
  int slot = 0;
  team_sync<kLbThreads>();                                   // everyone has read the state before thread 0 rewrites it

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
      sy = block_sum<kLbThreads>(sy, red); ss = block_sum<kLbThreads>(ss, red); yy = block_sum<kLbThreads>(yy, red);
      const bool good = sy > 1e-10 * sqrt(ss * yy);
      double* Sh = S + (size_t)head * n; double* Yh = Y + (size_t)head * n;
      for (int i = tid; i < n; i += kLbThreads) {
        const double xi = xt[i], gi = gt[i];
        if (good) { Sh[i] = xi - x[i]; Yh[i] = gi - g[i]; }
        x[i] = xi; g[i] = gi;
      }
      if (good) { if (tid == 0) rh[head] = 1.0 / sy; new_pair = true; slot = head; gamma = sy / yy; have = 1; head = (head + 1) % m; cnt = min(cnt + 1, m); }
      f = ft; accepted = true; inner += 1;
    } else {
      nls += 1;
      if (nls >= a.o.ls_max) {                               // line search failed
        if (steep_in || !have) { flag = LB_FAILED; }         // even steepest descent cannot decrease: at the noise floor -> inner done
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
    gn = 0.0; g1 = 0.0;
    for (int i = tid; i < n; i += kLbThreads) { const double ag = fabs(g[i]); gn = fmax(gn, ag); g1 += ag; }
    gn = block_max<kLbThreads>(gn, red); g1 = block_sum<kLbThreads>(g1, red);
    const int wdw = a.o.window;                              // windowed decrease test (ring of the last `window` accepted values)
    bool flat = false;
    if (inner >= wdw) flat = (fh[inner % wdw] - f) <= a.o.ftol * fmax(1.0, fabs(f));
    team_sync<kLbThreads>();
    if (tid == 0) fh[inner % wdw] = f;
    inner_done = gn <= a.o.gtol || flat || inner >= a.o.max_inner;
    new_dir = !inner_done;
    nls = 0;
  }
  if (flag == LB_FAILED) { inner_done = true; new_dir = false; flag = LB_TRIAL; }

  if (inner_done) {                                          // multiplier update (or termination) at the accepted point x
    team_sync<kLbThreads>();
    double cm = 0.0;
    for (int k = tid; k < a.n_con; k += kLbThreads) cm = fmax(cm, fabs(cacc[k]));
    cm = block_max<kLbThreads>(cm, red);
    const int outer = outer_in + 1;
    if (cm < a.o.ctol || outer >= a.o.max_outer) {
      flag = cm < a.o.ctol ? LB_SOLVED : LB_FAILED;
      for (int i = tid; i < n; i += kLbThreads) xt[i] = x[i];
    } else {
      const double rho = a.rho[p], cprev = sc[SC_CPREV];
      const bool raise = (cprev < 0.0 || cm > 0.25 * cprev) && rho < a.o.rho_max;
      for (int k = tid; k < a.n_con; k += kLbThreads) lam[k] += rho * cacc[k];
      team_sync<kLbThreads>();
      if (tid == 0) {
        if (raise) a.rho[p] = fmin(rho * 3.0, a.o.rho_max);
        sc[SC_CPREV] = cm;
      }
      for (int i = tid; i < n; i += kLbThreads) xt[i] = x[i];
      flag = LB_EVAL0;
      if (raise || !a.o.keep_history) { cnt = 0; have = 0; head = 0; }   // a multiplier update alone changes the Hessian little: keep the pairs
    }
    if (tid == 0) { sc[SC_CMAX] = cm; mi[MI_OUTER] = outer; }
  }

  int steep = 0;
  if (new_dir) {
    if (!a.gram) {
      // ---- classic two-loop recursion on d (thread i owns d[i], d[i + T], ...): 2 cnt dependent reductions, least work ----
      team_sync<kLbThreads>();
      for (int i = tid; i < n; i += kLbThreads) d[i] = g[i];
      for (int j = 0; j < cnt; ++j) {
        const int k = (head - 1 - j + 2 * m) % m;
        const double* Sk = S + (size_t)k * n; const double* Yk = Y + (size_t)k * n;
        double t = 0.0;
        for (int i = tid; i < n; i += kLbThreads) t += Sk[i] * d[i];
        t = rh[k] * block_sum<kLbThreads>(t, red);
        if (tid == 0) al[j] = t;
        for (int i = tid; i < n; i += kLbThreads) d[i] -= t * Yk[i];
      }
      for (int i = tid; i < n; i += kLbThreads) d[i] *= gamma;
      team_sync<kLbThreads>();
      for (int j = cnt - 1; j >= 0; --j) {
        const int k = (head - 1 - j + 2 * m) % m;
        const double* Sk = S + (size_t)k * n; const double* Yk = Y + (size_t)k * n;
        double t = 0.0;
        for (int i = tid; i < n; i += kLbThreads) t += Yk[i] * d[i];
        t = al[j] - rh[k] * block_sum<kLbThreads>(t, red);
        for (int i = tid; i < n; i += kLbThreads) d[i] += t * Sk[i];
      }
      double t = 0.0, gg = 0.0;
      for (int i = tid; i < n; i += kLbThreads) { t += g[i] * d[i]; gg += g[i] * g[i]; }
      t = block_sum<kLbThreads>(t, red); gg = block_sum<kLbThreads>(gg, red);
      gd = -t;                                               // d currently holds +H g
      if (!(gd < -1e-14 * gn * gn) || !have) steep = 1;
      if (steep) {
        const double scale = have ? gamma : 1.0 / fmax(g1, 1e-300);
        gd = -scale * gg;
        for (int i = tid; i < n; i += kLbThreads) { const double di = -scale * g[i]; d[i] = di; xt[i] = x[i] + di; }
      } else {
        for (int i = tid; i < n; i += kLbThreads) { const double di = -d[i]; d[i] = di; xt[i] = x[i] + di; }
      }
    } else {
  // ---- L-BFGS direction in the basis {s_k, y_k, g} (two-loop recursion on coefficients; the Gram matrix carries all
      //      dot products, so the vectors are read in two dependency-free passes) ----
      team_sync<kLbThreads>();
      for (int e = tid; e < B * B; e += kLbThreads) Gs[e] = Gg[e];
      const int nvec = 2 * cnt + 1;                            // valid basis vectors: s and y of the cnt newest slots, then g
      auto vec_index = [&](int v) { const int k = (head - cnt + (v >> 1) + 2 * m) % m; return v == nvec - 1 ? 2 * m : ((v & 1) ? m + k : k); };
      auto vec_ptr = [&](int idx) -> const double* { return idx == 2 * m ? g : (idx >= m ? Y + (size_t)(idx - m) * n : S + (size_t)idx * n); };
      if (accepted) {                                          // g changed (and maybe a new pair): refresh their Gram rows
        // each warp streams whole basis vectors (all lanes, coalesced, independent accumulators) and finishes their three dot
        // products with shuffles: no cross-warp reduction
        const double* sn = S + (size_t)slot * n; const double* yn = Y + (size_t)slot * n;
        for (int v = wid; v < nvec; v += kLbThreads / 32) {
          const int row = vec_index(v);
          const double* b = vec_ptr(row);
          double tg = 0.0, ts = 0.0, ty = 0.0;
          if (new_pair) for (int i = lane; i < n; i += 32) { const double bi = b[i]; tg += bi * g[i]; ts += bi * sn[i]; ty += bi * yn[i]; }
          else for (int i = lane; i < n; i += 32) tg += b[i] * g[i];
          tg = warp_sum(tg);
          if (new_pair) { ts = warp_sum(ts); ty = warp_sum(ty); }
          if (lane == 0) {
            Gs[row * B + 2 * m] = tg; Gs[2 * m * B + row] = tg; Gg[row * B + 2 * m] = tg; Gg[2 * m * B + row] = tg;
            if (new_pair) {
              Gs[row * B + slot] = ts; Gs[slot * B + row] = ts; Gg[row * B + slot] = ts; Gg[slot * B + row] = ts;
              Gs[row * B + m + slot] = ty; Gs[(m + slot) * B + row] = ty; Gg[row * B + m + slot] = ty; Gg[(m + slot) * B + row] = ty;
            }
          }
        }
      }
      team_sync<kLbThreads>();
      if (wid == 0) {                                          // coefficient two-loop: lanes over the basis index
        for (int l = lane; l < B; l += 32) dl[l] = l == 2 * m ? 1.0 : 0.0;
        __syncwarp();
        for (int i = cnt - 1; i >= 0; --i) {
          const int k = (head - cnt + i + 2 * m) % m;
          double t = 0.0;
          for (int l = lane; l < B; l += 32) t += dl[l] * Gs[k * B + l];
          t = warp_sum(t) / Gs[k * B + m + k];
          if (lane == 0) { al[i] = t; dl[m + k] -= t; }
          __syncwarp();
        }
        for (int l = lane; l < B; l += 32) dl[l] *= gamma;
        __syncwarp();
        for (int i = 0; i < cnt; ++i) {
          const int k = (head - cnt + i + 2 * m) % m;
          double t = 0.0;
          for (int l = lane; l < B; l += 32) t += dl[l] * Gs[(m + k) * B + l];
          t = warp_sum(t) / Gs[k * B + m + k];
          if (lane == 0) dl[k] += al[i] - t;
          __syncwarp();
        }
        double t = 0.0;
        for (int l = lane; l < B; l += 32) t += dl[l] * Gs[2 * m * B + l];
        t = warp_sum(t);
        if (lane == 0) red[kLbThreads / 32] = -t;              // g . d with d = -sum dl[l] b_l
      }
      team_sync<kLbThreads>();
      gd = red[kLbThreads / 32];
      if (!(gd < -1e-14 * gn * gn) || !have) steep = 1;
      if (steep) {
        const double scale = have ? gamma : 1.0 / fmax(g1, 1e-300);
        gd = -scale * Gs[2 * m * B + 2 * m];
        for (int i = tid; i < n; i += kLbThreads) { const double di = -scale * g[i]; d[i] = di; xt[i] = x[i] + di; }
      } else {
        double* vcoef = part;                                  // coefficient of basis vector v, in the order of the loop below
        for (int v = tid; v < nvec; v += kLbThreads) vcoef[v] = dl[vec_index(v)];
        team_sync<kLbThreads>();
        const int k0 = (head - cnt + 2 * m) % m;               // oldest valid slot
        for (int i = tid; i < n; i += kLbThreads) {
          double di = -vcoef[nvec - 1] * g[i];
          int k = k0;
          for (int v = 0; v + 1 < nvec; v += 2) {              // (s_k, y_k) pairs, oldest to newest
            di -= vcoef[v] * S[(size_t)k * n + i] + vcoef[v + 1] * Y[(size_t)k * n + i];
            k = k + 1 == m ? 0 : k + 1;
          }
          d[i] = di; xt[i] = x[i] + di;
        }
      }
    }
    alpha = 1.0; nls = 0;
    flag = LB_TRIAL;
  }

  if (tid == 0) {
    mi[MI_FLAG] = flag; mi[MI_HEAD] = head; mi[MI_CNT] = cnt; mi[MI_HAVE] = have; mi[MI_NLS] = nls; mi[MI_INNER] = inner;
    if (new_dir) mi[MI_STEEP] = steep;
    mi[MI_NFEV] += 1;
    if (accepted && flag_in == LB_TRIAL) mi[MI_TOTAL_ITS] += 1;
    sc[SC_F] = f; sc[SC_ALPHA] = alpha; sc[SC_GAMMA] = gamma; sc[SC_GD] = gd; sc[SC_GN] = gn; sc[SC_G1] = g1;
    if (flag != LB_SOLVED && flag != LB_FAILED) atomicAdd(a.n_running, 1);
  }
}

}  // namespace d2dx

using namespace d2dx;

static int lb_check(int P, int n, int n_con, const d2dx_lbfgs_options* o, const char* who) {
  D2DX_CHECK_ARG(o && P >= 1 && n >= 1 && n_con >= 0, "%s: P=%d n=%d n_con=%d", who, P, n, n_con);
  D2DX_CHECK_ARG(o->m >= 1 && o->m <= 32 && o->window >= 1 && o->ls_max >= 1 && o->max_inner >= 1 && o->max_outer >= 1,
                 "%s: m=%d (1..32) window=%d ls_max=%d max_inner=%d max_outer=%d", who, o->m, o->window, o->ls_max, o->max_inner, o->max_outer);
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
  D2DX_NVTX("d2dx_lbfgs_init");
  if (int rc = lb_check(P, n, n_con, o, "d2dx_lbfgs_init")) return rc;
  D2DX_CHECK_ARG(h && state && rho && (lam || n_con == 0), "d2dx_lbfgs_init: null array");
  const LbLayout L = lb_layout(P, n, n_con, o->m, o->window);
  D2DX_CUDA(cudaSetDevice(h->device));
  // integer and scalar state to zero (flag = EVAL0), c_prev = -1 (none yet), multipliers 0, rho = rho0
  D2DX_CUDA(cudaMemsetAsync(state + L.G, 0, sizeof(double) * (L.total - L.G), as_stream(stream)));
  lbfgs_init_kernel<<<(P + 127) / 128, 128, 0, as_stream(stream)>>>(P, state + L.sc, rho, o->rho0);
  D2DX_LAUNCH_CHECK("lbfgs_init_kernel");
  if (n_con) D2DX_CUDA(cudaMemsetAsync(lam, 0, sizeof(double) * (size_t)P * n_con, as_stream(stream)));
  return D2DX_OK;
}

extern "C" int d2dx_al_lbfgs_tick(d2dx_handle* h, int32_t P, int32_t n, int32_t n_con, const d2dx_lbfgs_options* o, double* state,
                                  double* x_trial, const double* f_parts, const double* cost_parts, int32_t n_parts, const double* grad,
                                  const double* c, double* lam, double* rho, int32_t* n_running, void* stream) {
  D2DX_NVTX("d2dx_al_lbfgs_tick");
  if (int rc = lb_check(P, n, n_con, o, "d2dx_al_lbfgs_tick")) return rc;
  D2DX_CHECK_ARG(h && state && x_trial && f_parts && grad && rho && n_running && n_parts >= 1 && (n_con == 0 || (c && lam)),
                 "d2dx_al_lbfgs_tick: null array or n_parts=%d", n_parts);
  LbArgs a = {};
  a.P = P; a.n = n; a.n_con = n_con; a.n_parts = n_parts; a.o = *o; a.L = lb_layout(P, n, n_con, o->m, o->window);
  a.state = state; a.xt = x_trial; a.fparts = f_parts; a.cparts = cost_parts; a.gt = grad; a.c = c; a.lam = lam; a.rho = rho;
  a.n_running = n_running;
  a.gram = (n >= 1024 && P <= 64) ? 1 : 0;   // measured: the Gram form wins only where the tick is latency-bound (profiles/README)
  D2DX_CUDA(cudaSetDevice(h->device));
  lbfgs_count_reset<<<1, 1, 0, as_stream(stream)>>>(n_running);
  if (n <= kLbWarpMaxN) {
    al_lbfgs_tick_kernel<32><<<P, 32, lb_smem_bytes(o->m, 32, a.gram), as_stream(stream)>>>(a);
  } else {                                                  // at most 47 kB of dynamic shared memory (m = 32): no opt-in needed
    al_lbfgs_tick_kernel<kLbWide><<<P, kLbWide, lb_smem_bytes(o->m, kLbWide, a.gram), as_stream(stream)>>>(a);
  }
  D2DX_LAUNCH_CHECK("al_lbfgs_tick_kernel");
  return D2DX_OK;
}
