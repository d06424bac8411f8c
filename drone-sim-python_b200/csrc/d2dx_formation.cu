// Circular formation: distributed circular-formation controller (DCF) + guidance vector field (GVF)
// + fixed-step RK4.  One thread = one aircraft; the n_ac aircraft of a formation sit in consecutive
// lanes of one warp and exchange their phase angles with warp shuffles (the only coupling on path A').
// Replaces CircularFormationGVF, 08_CircularFormation_Full.py:21-97 / 09_CircularFormation_diffcentre.py:21-118.
#include "d2dx_device.cuh"
#include "d2dx_host.h"

namespace d2dx {

constexpr int kFormThreads = 128;
constexpr int kMaxAc = 32;

struct FormArgs {
  int F, n_ac, n_e, fpw;                 // fpw = formations per warp
  const double *X0, *c, *r, *ac;
  double ke, kd, kr, v_c, dt;
  int i_begin, i_end, nsub;
  d2dx_formation_out o;
  double zdes[kMaxAc];
  double Binc[kMaxAc * kMaxAc];          // row-major [n_ac][n_e]
};

// DCFController.get, d2d/guidance.py:103-126, for the formation whose first lane is `base`.
// Every lane of the warp must call it.  Returns U_r of this lane's aircraft.  Lane j of the formation owns the edges
// j, j + n_ac, j + 2 n_ac, ... (the reference accepts any incidence matrix: a complete graph has more edges than aircraft);
// `emit(edge, e)` receives each owned edge's wrapped error [rad].
template <typename Emit>
__device__ __forceinline__ double dcf_warp(const double* sB, const double* sz, int n_ac, int n_e, int base, int j,
                                           double theta, double kr, Emit emit) {
  double ur = 0.0;
  for (int e0 = 0; e0 < n_e; e0 += n_ac) {                 // one pass when n_e <= n_ac (chain and ring graphs)
    const int edge = e0 + j;
    double z = 0.0;
    for (int i = 0; i < n_ac; ++i) {                       // z = B^T theta
      const double th_i = __shfl_sync(0xffffffffu, theta, base + i);
      if (edge < n_e) z = fma(sB[i * n_e + edge], th_i, z);
    }
    double e = (edge < n_e) ? z - sz[edge] : 0.0;
    if (e > kPi) e -= kTwoPi;                              // :115-120, both tests in sequence
    if (e <= -kPi) e += kTwoPi;
    if (edge < n_e) emit(edge, e);
    const int kk_end = n_e - e0 < n_ac ? n_e - e0 : n_ac;
    for (int kk = 0; kk < kk_end; ++kk) {                  // U_r = -kr B e
      const double e_k = __shfl_sync(0xffffffffu, e, base + kk);
      ur = fma(sB[j * n_e + e0 + kk], e_k, ur);
    }
  }
  return -kr * ur;
}

__global__ void __launch_bounds__(kFormThreads) rollout_formation_kernel(const __grid_constant__ FormArgs a) {
  __shared__ double sB[kMaxAc * kMaxAc];
  __shared__ double sz[kMaxAc];
  __shared__ double slag[8][kFormThreads];               // stage factors of the bank / air-speed lags (rk4_step, LagCoef), per thread
  for (int k = threadIdx.x; k < a.n_ac * a.n_e; k += kFormThreads) sB[k] = a.Binc[k];
  for (int k = threadIdx.x; k < a.n_e; k += kFormThreads) sz[k] = a.zdes[k];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * kFormThreads + threadIdx.x) >> 5;
  const int lf = lane / a.n_ac, j = lane - lf * a.n_ac;
  const long f_raw = warp * a.fpw + lf;
  const bool active = lf < a.fpw && f_raw < a.F;
  const long f = active ? f_raw : 0;
  const int base = lf * a.n_ac < 32 ? lf * a.n_ac : 0;
  const size_t M = (size_t)a.F * a.n_ac;                 // aircraft in the batch
  const size_t g = (size_t)f * a.n_ac + j;

  AcPar ap;
  ap.wx = 0.0; ap.wy = 0.0;                              // WindField() default, 08_CircularFormation_Full.py:27
  ap.n_inv_tau_phi = -1.0 / a.ac[g]; ap.n_inv_tau_v = -1.0 / a.ac[M + g];
  {                                                      // dt, nsub and the time constants are fixed for the launch
    const LagCoef L = lag_coef(a.nsub == 1 ? a.dt : a.dt * (1.0 / a.nsub), ap);
    const int t = threadIdx.x;
    slag[0][t] = L.f2; slag[1][t] = L.f3; slag[2][t] = L.f4; slag[3][t] = L.ff; slag[4][t] = L.v2; slag[5][t] = L.v3; slag[6][t] = L.v4; slag[7][t] = L.vf;
  }
  const double cx = a.c[g], cy = a.c[M + g], R = a.r[g];
  double X[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) X[k] = a.X0[k * M + g];
  const int log_every = a.o.log_every > 0 ? a.o.log_every : 1;

  // countdown to the next logged step instead of an integer modulo and division per step (60 instructions)
  int log_in = (log_every - a.i_begin % log_every) % log_every;
  size_t row = (size_t)((a.i_begin + log_every - 1) / log_every);
  for (int i = a.i_begin; i < a.i_end; ++i) {
    const bool log_now = active && log_in == 0;
    if (log_now && a.o.X_log) {
#pragma unroll
      for (int k = 0; k < 5; ++k) a.o.X_log[(row * 5 + k) * M + g] = X[k];
    }
    const double theta = atan2_f(X[1] - cy, X[0] - cx);
    const bool log_eth = log_now && a.o.eth_log != nullptr;
    const double Ur = dcf_warp(sB, sz, a.n_ac, a.n_e, base, j, theta, a.kr, [&](int edge, double e) {
      if (log_eth) a.o.eth_log[row * ((size_t)a.F * a.n_e) + (size_t)f * a.n_e + edge] = e * (180.0 / kPi);
    });
    const double Rr = Ur + R;                            // 08_CircularFormation_Full.py:76
    double U, U1, U2;
    gvf_control(X[0], X[1], X[2], X[4], cx, cy, Rr, a.ke, a.kd, U, U1, U2);
    const double phi_c = atan_f(U * (1.0 / 9.81));       // :85  arctan(U/9.81)
    if (log_now) {
      if (a.o.U_log) a.o.U_log[row * M + g] = phi_c;
      if (a.o.Rr_log) a.o.Rr_log[row * M + g] = Rr;
    }
    {
      const int t = threadIdx.x;
      const LagCoef L = {slag[0][t], slag[1][t], slag[2][t], slag[3][t], slag[4][t], slag[5][t], slag[6][t], slag[7][t]};
      rk4_step<false, true>(ap, X, phi_c, a.v_c, a.dt, a.nsub, &L);   // :90
    }
    if (log_in == 0) { log_in = log_every; ++row; }
    --log_in;
  }
  if (active) {
    if (a.o.X_log && (a.i_end % log_every) == 0) {
      const size_t row = (size_t)(a.i_end / log_every);
#pragma unroll
      for (int k = 0; k < 5; ++k) a.o.X_log[(row * 5 + k) * M + g] = X[k];
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) a.o.X_final[k * M + g] = X[k];
    if (a.o.flags) a.o.flags[g] |= isfinite(X[0] + X[1] + X[2] + X[3] + X[4]) ? 0 : 1;
  }
}

// single DCF evaluation (DCFController.get)
struct DcfArgs {
  int F, n_ac, n_e, fpw;
  const double *p, *c;
  double kr;
  double *Ur, *e_deg;
  double zdes[kMaxAc];
  double Binc[kMaxAc * kMaxAc];
};

__global__ void __launch_bounds__(kFormThreads) dcf_kernel(const __grid_constant__ DcfArgs a) {
  __shared__ double sB[kMaxAc * kMaxAc];
  __shared__ double sz[kMaxAc];
  __shared__ double slag[8][kFormThreads];               // stage factors of the bank / air-speed lags (rk4_step, LagCoef), per thread
  for (int k = threadIdx.x; k < a.n_ac * a.n_e; k += kFormThreads) sB[k] = a.Binc[k];
  for (int k = threadIdx.x; k < a.n_e; k += kFormThreads) sz[k] = a.zdes[k];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * kFormThreads + threadIdx.x) >> 5;
  const int lf = lane / a.n_ac, j = lane - lf * a.n_ac;
  const long f_raw = warp * a.fpw + lf;
  const bool active = lf < a.fpw && f_raw < a.F;
  const long f = active ? f_raw : 0;
  const int base = lf * a.n_ac < 32 ? lf * a.n_ac : 0;
  const size_t M = (size_t)a.F * a.n_ac, g = (size_t)f * a.n_ac + j;
  const double theta = atan2_f(a.p[M + g] - a.c[M + g], a.p[g] - a.c[g]);
  const double Ur = dcf_warp(sB, sz, a.n_ac, a.n_e, base, j, theta, a.kr, [&](int edge, double e) {
    if (active) a.e_deg[(size_t)f * a.n_e + edge] = e * (180.0 / kPi);
  });
  if (active) a.Ur[g] = Ur;
}

__global__ void __launch_bounds__(kFormThreads) gvf_kernel(int n, const double* __restrict__ X, const double* __restrict__ c,
                                                            const double* __restrict__ r, double ke, double kd,
                                                            double* __restrict__ out) {
  const int i = blockIdx.x * kFormThreads + threadIdx.x;
  if (i >= n) return;
  double U, U1, U2;
  gvf_control(X[i], X[(size_t)n + i], X[2 * (size_t)n + i], X[4 * (size_t)n + i], c[i], c[(size_t)n + i], r[i], ke, kd, U, U1, U2);
  out[i] = U; out[(size_t)n + i] = U1; out[2 * (size_t)n + i] = U2;
}

__global__ void __launch_bounds__(kFormThreads) circle_implicit_kernel(int n, const double* __restrict__ X, const double* __restrict__ c,
                                                                        const double* __restrict__ r, double* __restrict__ out) {
  const int i = blockIdx.x * kFormThreads + threadIdx.x;
  if (i >= n) return;
  const double dx = X[i] - c[i], dy = X[(size_t)n + i] - c[(size_t)n + i];
  out[i] = (dx * dx + dy * dy) - r[i] * r[i];             // e = |p - c|^2 - r^2   (d2d/guidance.py:140)
  out[(size_t)n + i] = 2.0 * dx; out[2 * (size_t)n + i] = 2.0 * dy;   // n = grad e    (:144)
}

int formation_resident_threads_per_sm() {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rollout_formation_kernel, kFormThreads, 0);
  return nb * kFormThreads;
}

}  // namespace d2dx

using namespace d2dx;

extern "C" int d2dx_rollout_formation(d2dx_handle* h, const d2dx_formations* f, double dt, int32_t i_begin, int32_t i_end,
                                      int32_t nsub, const d2dx_formation_out* out, void* stream) {
  D2DX_NVTX("d2dx_rollout_formation");
  D2DX_CHECK_ARG(h && f && out, "d2dx_rollout_formation: null argument");
  D2DX_CHECK_ARG(f->F > 0 && f->n_ac >= 1 && f->n_ac <= kMaxAc && f->n_e >= 0 && f->n_e <= kMaxAc,
                 "d2dx_rollout_formation: F=%d n_ac=%d n_e=%d (n_ac, n_e <= %d)", f->F, f->n_ac, f->n_e, kMaxAc);
  D2DX_CHECK_ARG(f->X0 && f->c && f->r && f->ac && out->X_final && (f->n_e == 0 || (f->Binc_host && f->z_des_host)),
                 "d2dx_rollout_formation: missing array");
  D2DX_CHECK_ARG(i_begin >= 0 && i_end >= i_begin && nsub >= 1 && dt > 0, "d2dx_rollout_formation: bad range or step");
  FormArgs a;
  a.F = f->F; a.n_ac = f->n_ac; a.n_e = f->n_e; a.fpw = 32 / f->n_ac;
  a.X0 = f->X0; a.c = f->c; a.r = f->r; a.ac = f->ac;
  a.ke = f->ke; a.kd = f->kd; a.kr = f->kr; a.v_c = f->v_c; a.dt = dt;
  a.i_begin = i_begin; a.i_end = i_end; a.nsub = nsub; a.o = *out;
  for (int k = 0; k < f->n_e; ++k) a.zdes[k] = f->z_des_host[k];
  for (int k = 0; k < f->n_ac * f->n_e; ++k) a.Binc[k] = f->Binc_host[k];
  D2DX_CUDA(cudaSetDevice(h->device));
  const long warps = ((long)f->F + a.fpw - 1) / a.fpw;
  const int grid = (int)((warps * 32 + kFormThreads - 1) / kFormThreads);
  rollout_formation_kernel<<<grid, kFormThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("rollout_formation_kernel");
  return D2DX_OK;
}

extern "C" int d2dx_dcf(d2dx_handle* h, int32_t F, int32_t n_ac, int32_t n_e, const double* Binc_host,
                        const double* z_des_host, double kr, const double* p, const double* c, double* Ur,
                        double* e_deg, void* stream) {
  D2DX_NVTX("d2dx_dcf");
  D2DX_CHECK_ARG(h && F > 0 && n_ac >= 1 && n_ac <= kMaxAc && n_e >= 0 && n_e <= kMaxAc, "d2dx_dcf: bad sizes F=%d n_ac=%d n_e=%d", F, n_ac, n_e);
  D2DX_CHECK_ARG(p && c && Ur && e_deg && Binc_host && z_des_host, "d2dx_dcf: null array");
  DcfArgs a;
  a.F = F; a.n_ac = n_ac; a.n_e = n_e; a.fpw = 32 / n_ac; a.p = p; a.c = c; a.kr = kr; a.Ur = Ur; a.e_deg = e_deg;
  for (int k = 0; k < n_e; ++k) a.zdes[k] = z_des_host[k];
  for (int k = 0; k < n_ac * n_e; ++k) a.Binc[k] = Binc_host[k];
  D2DX_CUDA(cudaSetDevice(h->device));
  const long warps = ((long)F + a.fpw - 1) / a.fpw;
  dcf_kernel<<<(int)((warps * 32 + kFormThreads - 1) / kFormThreads), kFormThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("dcf_kernel");
  return D2DX_OK;
}

extern "C" int d2dx_circle_implicit(d2dx_handle* h, int32_t n, const double* X, const double* c, const double* r, double* out,
                                    void* stream) {
  D2DX_NVTX("d2dx_circle_implicit");
  D2DX_CHECK_ARG(h && n > 0 && X && c && r && out, "d2dx_circle_implicit: bad argument");
  D2DX_CUDA(cudaSetDevice(h->device));
  circle_implicit_kernel<<<(n + kFormThreads - 1) / kFormThreads, kFormThreads, 0, as_stream(stream)>>>(n, X, c, r, out);
  D2DX_LAUNCH_CHECK("circle_implicit_kernel");
  return D2DX_OK;
}

extern "C" int d2dx_gvf(d2dx_handle* h, int32_t n, const double* X, const double* c, const double* r, double ke, double kd,
                        double* out, void* stream) {
  D2DX_NVTX("d2dx_gvf");
  D2DX_CHECK_ARG(h && n > 0 && X && c && r && out, "d2dx_gvf: bad argument");
  D2DX_CUDA(cudaSetDevice(h->device));
  gvf_kernel<<<(n + kFormThreads - 1) / kFormThreads, kFormThreads, 0, as_stream(stream)>>>(n, X, c, r, ke, kd, out);
  D2DX_LAUNCH_CHECK("gvf_kernel");
  return D2DX_OK;
}
