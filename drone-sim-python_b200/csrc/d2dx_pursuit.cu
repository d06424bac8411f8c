// Pure-pursuit guidance on a sampled path (PurePursuitControler, d2d/guidance.py:204-245) and its
// closed-loop rollout (the run_simulation loop of 05_test_simulation.py:21-34 with ctl = PurePursuitControler).
// One warp = one aircraft: the lanes share the O(n_pts) nearest-point search of every control step (first index on
// ties, like np.argmin on np.linalg.norm), lane-redundant control law and RK4 step afterwards.
#include "d2dx_device.cuh"
#include "d2dx_host.h"

namespace d2dx {

constexpr int kPpThreads = 128;

struct PursuitArgs {
  d2dx_pursuit p;
  int B;
  const double *X0, *wind, *ac;
  double dt;
  int i_begin, i_end, nsub;
  double *X_log, *U_log, *X_final;
  int32_t* idx_log;
};

// nearest sample: np.argmin(np.linalg.norm(pts - X[:2], axis=1)) -- sqrt(dx*dx + dy*dy) without contraction, first minimum.
// The square root (correctly rounded, as NumPy's) is only taken for candidates: a point can only beat the running minimum
// if its squared distance is within a few ulp of, or below, the squared distance of the current best.
__device__ __forceinline__ int nearest_point(const d2dx_pursuit& p, double x, double y, int lane) {
  double best = __longlong_as_double(0x7ff0000000000000LL), lim = best;   // +inf
  int bi = 0x7fffffff;
  for (int j = lane; j < p.n_pts; j += 32) {
    const double dx = p.px[j] - x, dy = p.py[j] - y;
    const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    if (d2 <= lim) {
      const double d = sqrt(d2);
      if (d < best) { best = d; bi = j; lim = d2 * (1.0 + 8.9e-16); }    // ascending j per lane: strict < keeps the first
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  return bi == 0x7fffffff ? 0 : bi;                  // non-finite position: np.argmin of an all-NaN array is 0
}

__device__ __forceinline__ void pursuit_law(const d2dx_pursuit& p, const double* X, int idx, double& phi_sp, double& v_sp) {
  int ic = idx + p.lookahead;                                // :228-230
  if (ic >= p.n_pts) ic -= p.n_pts;
  const double pcx = p.px[ic] - X[0], pcy = p.py[ic] - X[1];
  const double err_psi = wrap_pi(X[2] - atan2(pcy, pcx));    // :237
  phi_sp = clip(-p.K * err_psi, -p.sat_phi, p.sat_phi);      // :239-240
  v_sp = p.v_sp;                                             // control_vel is False upstream (:217)
}

__global__ void __launch_bounds__(kPpThreads) pursuit_control_kernel(const __grid_constant__ PursuitArgs a) {
  const int lane = threadIdx.x & 31;
  const long g = (long)blockIdx.x * (kPpThreads / 32) + (threadIdx.x >> 5);
  if (g >= a.B) return;
  double X[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) X[k] = a.X0[(size_t)k * a.B + g];
  const int idx = nearest_point(a.p, X[0], X[1], lane);
  double phi_sp, v_sp;
  pursuit_law(a.p, X, idx, phi_sp, v_sp);
  if (lane == 0) {
    a.U_log[g] = phi_sp; a.U_log[(size_t)a.B + g] = v_sp;
    if (a.idx_log) a.idx_log[g] = idx;
  }
}

__global__ void __launch_bounds__(kPpThreads) rollout_pursuit_kernel(const __grid_constant__ PursuitArgs a) {
  const int lane = threadIdx.x & 31;
  const long g = (long)blockIdx.x * (kPpThreads / 32) + (threadIdx.x >> 5);
  if (g >= a.B) return;
  const size_t B = a.B;
  double X[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) X[k] = a.X0[(size_t)k * B + g];
  AcPar ap;
  ap.wx = a.wind[g]; ap.wy = a.wind[B + g];
  ap.n_inv_tau_phi = -1.0 / a.ac[g]; ap.n_inv_tau_v = -1.0 / a.ac[B + g];
  for (int i = a.i_begin; i < a.i_end; ++i) {
    if (lane == 0 && a.X_log) {
#pragma unroll
      for (int k = 0; k < 5; ++k) a.X_log[((size_t)i * 5 + k) * B + g] = X[k];
    }
    const int idx = nearest_point(a.p, X[0], X[1], lane);
    double phi_sp, v_sp;
    pursuit_law(a.p, X, idx, phi_sp, v_sp);
    if (lane == 0) {
      if (a.U_log) { a.U_log[((size_t)i * 2 + 0) * B + g] = phi_sp; a.U_log[((size_t)i * 2 + 1) * B + g] = v_sp; }
      if (a.idx_log) a.idx_log[(size_t)i * B + g] = idx;
    }
    rk4_generic_inplace(ap, X, phi_sp, v_sp, a.dt, a.nsub);  // every lane integrates the same state
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (a.X_log) a.X_log[((size_t)a.i_end * 5 + k) * B + g] = X[k];
      if (a.X_final) a.X_final[(size_t)k * B + g] = X[k];
    }
  }
}

static int pp_check(const d2dx_pursuit* p, const char* who) {
  D2DX_CHECK_ARG(p && p->n_pts >= 1 && p->px && p->py, "%s: empty path", who);
  D2DX_CHECK_ARG(p->lookahead >= 0 && p->lookahead <= p->n_pts, "%s: lookahead=%d for %d points", who, p->lookahead, p->n_pts);
  return D2DX_OK;
}

}  // namespace d2dx

using namespace d2dx;

extern "C" int d2dx_pursuit_control(d2dx_handle* h, const d2dx_pursuit* p, int32_t B, const double* X, double* U, int32_t* idx_closest,
                                    void* stream) {
  D2DX_NVTX("d2dx_pursuit_control");
  if (int rc = pp_check(p, "d2dx_pursuit_control")) return rc;
  D2DX_CHECK_ARG(h && B >= 1 && X && U, "d2dx_pursuit_control: bad argument");
  PursuitArgs a = {};
  a.p = *p; a.B = B; a.X0 = X; a.U_log = U; a.idx_log = idx_closest;
  D2DX_CUDA(cudaSetDevice(h->device));
  pursuit_control_kernel<<<(B + kPpThreads / 32 - 1) / (kPpThreads / 32), kPpThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("pursuit_control_kernel");
  return D2DX_OK;
}

extern "C" int d2dx_rollout_pursuit(d2dx_handle* h, const d2dx_pursuit* p, int32_t B, const double* X0, const double* wind, const double* ac,
                                    double dt, int32_t i_begin, int32_t i_end, int32_t nsub, double* X_log, double* U_log,
                                    int32_t* idx_log, double* X_final, void* stream) {
  D2DX_NVTX("d2dx_rollout_pursuit");
  if (int rc = pp_check(p, "d2dx_rollout_pursuit")) return rc;
  D2DX_CHECK_ARG(h && B >= 1 && X0 && wind && ac && dt > 0 && nsub >= 1 && i_begin >= 0 && i_end >= i_begin,
                 "d2dx_rollout_pursuit: B=%d dt=%g nsub=%d steps [%d, %d)", B, dt, nsub, i_begin, i_end);
  PursuitArgs a = {};
  a.p = *p; a.B = B; a.X0 = X0; a.wind = wind; a.ac = ac; a.dt = dt; a.i_begin = i_begin; a.i_end = i_end; a.nsub = nsub;
  a.X_log = X_log; a.U_log = U_log; a.idx_log = idx_log; a.X_final = X_final;
  D2DX_CUDA(cudaSetDevice(h->device));
  rollout_pursuit_kernel<<<(B + kPpThreads / 32 - 1) / (kPpThreads / 32), kPpThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("rollout_pursuit_kernel");
  return D2DX_OK;
}
