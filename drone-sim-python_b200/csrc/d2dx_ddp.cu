// Batched second-order planner solve: one thread = one problem (d2dx_ddp.cuh), work arrays interleaved over the problems.
#include "d2dx_ddp.cuh"
#include "d2dx_host.h"

namespace d2dx {

constexpr int kDdpThreads = 64;

// fills the solver's problem description from the collocation problem + bounds + box (host and device builds share it)
__host__ inline int ddp_problem_from(const d2dx_colloc_problem* p, const double* bounds, const double* box, DdpProblem& P) {
  if (!p || p->n_ac != 1 || p->N < 2 || !(p->h > 0) || !bounds) return 1;
  const double sN = p->obj_scale / p->N, nin = sN / (p->in_div >= 1 ? p->in_div : 1);
  P.N = p->N; P.h = p->h; P.wx = p->wind[0]; P.wy = p->wind[1];
  P.vsp = p->vsp; P.kv = p->kvel * nin; P.kb = p->kbank * nin;
  const bool obs = (p->kobs == p->kobs) && p->kobs != 0.0 && p->n_obs > 0;
  P.kobs = obs ? p->kobs * sN : 0.0; P.n_obs = obs ? p->n_obs : 0; P.obs_kind = p->obs_kind;
  for (int o = 0; o < P.n_obs; ++o) { P.obs[o][0] = p->obs[o][0]; P.obs[o][1] = p->obs[o][1]; P.obs[o][2] = p->obs[o][2]; }
  P.phi_lo = bounds[0]; P.phi_hi = bounds[1]; P.v_lo = bounds[2]; P.v_hi = bounds[3];
  P.has_box = box != nullptr;
  if (box) { P.x_lo = box[0]; P.x_hi = box[1]; P.y_lo = box[2]; P.y_hi = box[3]; P.w_box = box[4] * sN; }
  else { P.x_lo = P.y_lo = -1e300; P.x_hi = P.y_hi = 1e300; P.w_box = 0.0; }
  return (P.phi_lo < P.phi_hi && P.v_lo < P.v_hi && P.v_lo > 0.0) ? 0 : 1;
}

struct DdpArgs {
  DdpProblem P;
  d2dx_ddp_options o;
  int n_prob;
  const double *p0, *p1;
  double *u, *xs, *info, *work;
};

__global__ void __launch_bounds__(kDdpThreads) ddp_solve_kernel(const __grid_constant__ DdpArgs a) {
  const int p = blockIdx.x * kDdpThreads + threadIdx.x;
  if (p >= a.n_prob) return;
  const int N = a.P.N;
  const long stride = a.n_prob;
  DdpWork W;
  W.N = N; W.stride = stride;
  double* base = a.work + p;                       // [array][k][node][problem]
  W.u = base; W.z = W.u + 2L * N * stride; W.un = W.z + 3L * N * stride; W.zn = W.un + 2L * N * stride;
  W.k = W.zn + 3L * N * stride; W.K = W.k + 2L * N * stride;
  const double* ui = a.u + (size_t)p * 2 * N;
  for (int i = 0; i < N; ++i) { W.at(W.u, 0, i) = ui[i]; W.at(W.u, 1, i) = ui[N + i]; }
  const double z0[3] = {a.p0[p * 3], a.p0[p * 3 + 1], a.p0[p * 3 + 2]}, zt[3] = {a.p1[p * 3], a.p1[p * 3 + 1], a.p1[p * 3 + 2]};
  const DdpResult r = ddp_solve(a.P, W, z0, zt, a.o);
  double* us = (r.swaps & 1) ? W.un : W.u;
  double* zs = (r.swaps & 1) ? W.zn : W.z;
  double* uo = a.u + (size_t)p * 2 * N;
  double* xo = a.xs + (size_t)p * 3 * N;
  for (int i = 0; i < N; ++i) {
    uo[i] = W.at(us, 0, i); uo[N + i] = W.at(us, 1, i);
    xo[i] = W.at(zs, 0, i); xo[N + i] = W.at(zs, 1, i); xo[2 * N + i] = W.at(zs, 2, i);
  }
  double* io = a.info + (size_t)p * 8;
  io[0] = r.flag; io[1] = r.iterations; io[2] = r.outer; io[3] = r.cost; io[4] = r.cmax; io[5] = r.lagr; io[6] = r.mu; io[7] = r.rho;
}

}  // namespace d2dx

using namespace d2dx;

extern "C" {

int d2dx_ddp_default_options(d2dx_ddp_options* o) {
  if (!o) return set_error(D2DX_EINVAL, "d2dx_ddp_default_options: null");
  o->max_iter = 400; o->max_outer = 30; o->max_inner = 40; o->ls_max = 12;
  o->ctol = 1e-8; o->rel_tol = 1e-10; o->abs_tol = 1e-14;
  o->rho0 = 10.0; o->rho_growth = 10.0; o->rho_max = 1e8;
  o->mu0 = 1e-6; o->mu_min = 1e-8; o->mu_max = 1e10; o->mu_factor = 1.6; o->reg_mode = 0;
  return D2DX_OK;
}

int64_t d2dx_ddp_work_size(int32_t P, int32_t N) { return (P < 1 || N < 2) ? 0 : 18LL * N * P; }

int d2dx_ddp_solve(d2dx_handle* h, const d2dx_colloc_problem* p, int32_t n_prob, const double* bounds_host4, const double* state_box_host5,
                   const double* p0, const double* p1, double* u, double* xs, double* info, double* work, const d2dx_ddp_options* o_host,
                   void* stream) {
  D2DX_NVTX("d2dx_ddp_solve");
  D2DX_CHECK_ARG(h && p && n_prob >= 1 && p0 && p1 && u && xs && info && work, "d2dx_ddp_solve: null argument or n_prob=%d", n_prob);
  DdpArgs a;
  D2DX_CHECK_ARG(ddp_problem_from(p, bounds_host4, state_box_host5, a.P) == 0,
                 "d2dx_ddp_solve: needs n_ac = 1, N >= 2, h > 0 and bounds phi_lo < phi_hi, 0 < v_lo < v_hi");
  if (o_host) a.o = *o_host; else d2dx_ddp_default_options(&a.o);
  D2DX_CHECK_ARG(a.o.max_iter >= 1 && a.o.max_outer >= 1 && a.o.max_inner >= 1 && a.o.ls_max >= 1 && a.o.mu_factor > 1.0 && a.o.rho0 > 0.0,
                 "d2dx_ddp_solve: bad options");
  a.n_prob = n_prob; a.p0 = p0; a.p1 = p1; a.u = u; a.xs = xs; a.info = info; a.work = work;
  D2DX_CUDA(cudaSetDevice(h->device));
  ddp_solve_kernel<<<(n_prob + kDdpThreads - 1) / kDdpThreads, kDdpThreads, 0, as_stream(stream)>>>(a);
  D2DX_LAUNCH_CHECK("ddp_solve_kernel");
  return D2DX_OK;
}

}  // extern "C"
