// Direct-collocation evaluation: backward-Euler defects of the d2d EoM (d2d/opty_utils.py:38-50), their
// sparse Jacobian, the planner cost and its gradient, over problems x aircraft x nodes.
// One thread = one (aircraft, node); a block covers a tile of nodes for every aircraft of one problem, the
// positions of all aircraft on that tile are staged in shared memory for the pairwise collision terms.
// HBM-bound: every load and store is coalesced along the node index.
#include "d2dx_device.cuh"
#include "d2dx_host.h"

namespace d2dx {

constexpr int kCollocThreads = 128;
constexpr int kMaxTickets = 65536;

struct CollocArgs {
  d2dx_colloc_problem p;
  int n_prob, layout;
  uint32_t what;
  const double* free_;
  double *res, *jac, *cost, *grad, *scratch;
  int32_t* tickets;
  int n_total, a_lo;          // shard context: owned aircraft are global [a_lo, a_lo + p.n_ac) of n_total
  const double* pos_all;      // [n_total][2][N] or NULL (positions come from free_)
  int TN, APP, ntiles;        // nodes per tile, aircraft per pass, tiles per problem
  int ticket_mode;            // 1: the last warp of a problem (atomic ticket) finishes the cost in this kernel;
                              // 0: per-warp partials only, colloc_cost_kernel finishes (large batches: no fence in the hot kernel)
  long n_free, n_con, nnz;
};

__device__ __forceinline__ bool enabled(double k) { return (k == k) && k != 0.0; }

// EXTRA = obstacle and / or pairwise collision terms present (exp, shared-memory pair loop); the plain instantiation is
// the lean HBM-bound path (residual + Jacobian + input cost + gradient) and is held to 64 registers for occupancy.
template <bool EXTRA>
__global__ void __launch_bounds__(kCollocThreads, EXTRA ? 5 : 8) colloc_kernel(const __grid_constant__ CollocArgs a) {
  extern __shared__ double spos[];                 // [n_total][2][TN] positions (+ two gradient accumulators of the same shape)
  const d2dx_colloc_problem& P = a.p;
  const int N = P.N, n_ac = P.n_ac, TN = a.TN;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x % a.ntiles, prob = blockIdx.x / a.ntiles;
  const int il = tid % TN, al = tid / TN;
  const int i = tile * TN + il;
  const double* fr = a.free_ + (size_t)prob * a.n_free;
  const bool want_cg = (a.what & (D2DX_EVAL_COST | D2DX_EVAL_GRAD)) != 0;
  const bool use_col = EXTRA && want_cg && enabled(P.kcol) && a.n_total > 1;
  const bool use_obs = EXTRA && want_cg && enabled(P.kobs) && P.n_obs > 0;
  // unsharded all-pairs mode: every unordered pair is evaluated ONCE (round k pairs aircraft a with a+k mod n) and its
  // gradient is scattered to both members through shared-memory accumulators, rounds separated by block barriers
  const bool sym_col = use_col && P.col_all_pairs && a.pos_all == nullptr;
  double* sown = spos + (size_t)a.n_total * 2 * a.TN;      // written only by the thread that owns (aircraft, node)
  double* sacc = sown + (size_t)a.n_total * 2 * a.TN;      // partner-side contributions

  if (use_col) {                                   // stage the tile's positions of ALL aircraft
    for (int idx = tid; idx < a.n_total * TN; idx += kCollocThreads) {
      const int g = idx / TN, ii = idx - g * TN, node = tile * TN + ii;
      double x = 0.0, y = 0.0;
      if (node < N) {
        if (a.pos_all) { x = a.pos_all[((size_t)g * 2) * N + node]; y = a.pos_all[((size_t)g * 2 + 1) * N + node]; }
        else { x = fr[(size_t)(3 * g) * N + node]; y = fr[(size_t)(3 * g + 1) * N + node]; }
      }
      spos[(g * 2) * TN + ii] = x; spos[(g * 2 + 1) * TN + ii] = y;
    }
    __syncthreads();
  }

  const int n = 3 * n_ac, q = 2 * n_ac;
  const double ih = 1.0 / P.h;
  const double sN = P.obj_scale / N;               // _p.obj_scale/_p.num_nodes
  const double norm_in = sN / P.in_div;
  const double col_kr = P.kcol_k / P.rcol;           // (k/r): one multiplication per pair component instead of dx / r * k
  double s_v = 0.0, s_phi = 0.0, s_obs = 0.0, s_col = 0.0;

  for (int a_l = al; a_l < n_ac; a_l += a.APP) {
    if (i >= N) break;
    const int bphi = P.perm_phi ? P.perm_phi[a_l] : a_l;
    const int bv = P.perm_v ? P.perm_v[a_l] : n_ac + a_l;
    const size_t ox = (size_t)(3 * a_l) * N + i, oy = ox + N, ops = oy + N;
    const size_t ophi = (size_t)(n + bphi) * N + i, ov = (size_t)(n + bv) * N + i;
    const double x = fr[ox], y = fr[oy], psi = fr[ops], phi = fr[ophi], v = fr[ov];

    if ((a.what & (D2DX_EVAL_RESIDUAL | D2DX_EVAL_JAC)) && i >= 1) {
      double s, c, sp, cp;
      sincos_any(psi, s, c);
      sincos_any(phi, sp, cp);
      const double iv = rcp_f(v);
      const double tn = sp * rcp_f(cp);            // tan(phi)
      const double gtv = kG * tn * iv;             // g tan(phi) / v
      if (a.what & D2DX_EVAL_RESIDUAL) {           // equation-major, node-minor (opty layout)
        const double xp = fr[ox - 1], yp = fr[oy - 1], pp = fr[ops - 1];
        double* r = a.res + (size_t)prob * a.n_con + (size_t)(3 * a_l) * (N - 1) + (i - 1);
        r[0] = (x - xp) * ih - v * c + P.wind[0];
        r[(size_t)(N - 1)] = (y - yp) * ih - v * s + P.wind[1];
        r[2 * (size_t)(N - 1)] = (psi - pp) * ih - gtv;
      }
      if (a.what & D2DX_EVAL_JAC) {
        const double j[12] = {ih, v * s, -ih, -c, ih, -v * c, -ih, -s, ih, -ih, -kG * fma(tn, tn, 1.0) * iv, gtv * iv};
        if (a.layout == D2DX_JAC_COMPACT) {        // [n_ac][12][N-1]: coalesced along the node
          double* jo = a.jac + (size_t)prob * a.nnz + (size_t)a_l * 12 * (N - 1) + (i - 1);
#pragma unroll
          for (int k = 0; k < 12; ++k) jo[(size_t)k * (N - 1)] = j[k];
        } else {                                   // opty-dense: [(N-1)][3 n_ac][8 n_ac]
          const int W = 2 * n + q;
          double* jo = a.jac + (size_t)prob * a.nnz + ((size_t)(i - 1) * n + 3 * a_l) * W;
          const int cx = 3 * a_l, cp = n + 3 * a_l, cphi = 2 * n + bphi, cv = 2 * n + bv;
          if (n_ac == 1) {                         // 24 contiguous values per node, structural zeros included
            jo[0] = j[0]; jo[1] = 0.0; jo[2] = j[1]; jo[3] = j[2]; jo[4] = 0.0; jo[5] = 0.0; jo[6] = 0.0; jo[7] = j[3];
            jo[8] = 0.0; jo[9] = j[4]; jo[10] = j[5]; jo[11] = 0.0; jo[12] = j[6]; jo[13] = 0.0; jo[14] = 0.0; jo[15] = j[7];
            jo[16] = 0.0; jo[17] = 0.0; jo[18] = j[8]; jo[19] = 0.0; jo[20] = 0.0; jo[21] = j[9]; jo[22] = j[10]; jo[23] = j[11];
          } else {                                 // non-zeros only; zeros were laid down by d2dx_colloc_init_dense
            jo[cx] = j[0]; jo[cx + 2] = j[1]; jo[cp] = j[2]; jo[cv] = j[3];
            jo[W + cx + 1] = j[4]; jo[W + cx + 2] = j[5]; jo[W + cp + 1] = j[6]; jo[W + cv] = j[7];
            jo[2 * W + cx + 2] = j[8]; jo[2 * W + cp + 2] = j[9]; jo[2 * W + cphi] = j[10]; jo[2 * W + cv] = j[11];
          }
        }
      }
    }

    if (want_cg) {
      const double dv = v - P.vsp;
      s_v += dv * dv; s_phi += phi * phi;
      double gx = 0.0, gy = 0.0;
      const int g_glob = a.a_lo + a_l;
      if (use_obs && g_glob == 0) {                // CostObstacle acts on aircraft 0 only (multiopty_utils.py:74)
        for (int o = 0; o < P.n_obs; ++o) {
          const double dx = x - P.obs[o][0], dy = y - P.obs[o][1], r = P.obs[o][2];
          double es, f = 1.0;
          if (P.obs_kind == 0) es = clip(exp(r * r - (dx * dx + dy * dy)), 0.0, 1e3);
          else {
            const double kr = 2.0 / r, ux = dx * kr, uy = dy * kr;
            es = fm::exp_neg(-(ux * ux + uy * uy));
            if (P.exact_grad) f = (2.0 / r) * (2.0 / r);
          }
          s_obs += es;
          gx += P.kobs * (sN * -2.0 * dx * es) * f;
          gy += P.kobs * (sN * -2.0 * dy * es) * f;
        }
      }
      if (sym_col) {                               // pair terms come in the rounds below; park the obstacle part
        sown[(a_l * 2) * TN + il] = gx; sown[(a_l * 2 + 1) * TN + il] = gy;
        sacc[(a_l * 2) * TN + il] = 0.0; sacc[(a_l * 2 + 1) * TN + il] = 0.0;
      } else if (use_col) {                        // CostCollision (multiopty_utils.py:120-153), pairs via shared memory
        const double f = P.exact_grad ? col_kr * col_kr : 1.0;
        const int b_lo = P.col_all_pairs ? 0 : (g_glob == 0 ? 1 : (g_glob == 1 ? 0 : 0));
        const int b_hi = P.col_all_pairs ? a.n_total : (g_glob == 0 ? 2 : (g_glob == 1 ? 1 : 0));
        for (int b = b_lo; b < b_hi; ++b) {
          if (b == g_glob) continue;
          const double dx = x - spos[(b * 2) * TN + il], dy = y - spos[(b * 2 + 1) * TN + il];
          const double ux = dx * col_kr, uy = dy * col_kr;
          const double es = fm::exp_neg(-(ux * ux + uy * uy));
          if (g_glob < b) s_col += es;
          gx += P.kcol * (sN * -2.0 * dx * es) * f;
          gy += P.kcol * (sN * -2.0 * dy * es) * f;
        }
      }
      if (a.what & D2DX_EVAL_GRAD) {
        double* go = a.grad + (size_t)prob * a.n_free;
        if (!sym_col) { go[ox] = gx; go[oy] = gy; }
        go[ops] = 0.0;
        go[ophi] = (P.kbank * 2.0 * phi) * norm_in;
        go[ov] = (P.kvel * 2.0 * dv) * norm_in;
      }
    }
  }

  if (EXTRA && sym_col) {                          // block-uniform condition: every thread takes part in the barriers
    const int n = n_ac, half = n / 2;
    const bool even = (n % 2) == 0;
    const double f = P.exact_grad ? col_kr * col_kr : 1.0;
    const double cw = P.kcol * sN * -2.0 * f;                     // weight of dx * es in the gradient
    const int row = 2 * TN;                                       // doubles per aircraft in the staged arrays
    for (int k = 1; k <= half; ++k) {
      __syncthreads();
      if (i < N) {
        for (int a_l = al; a_l < n; a_l += a.APP) {
          if (even && k == half && a_l >= half) continue;         // the antipodal pair is visited from its lower member only
          int b = a_l + k; if (b >= n) b -= n;
          const int oa = a_l * row + il, ob = b * row + il;
          const double dx = spos[oa] - spos[ob];
          const double dy = spos[oa + TN] - spos[ob + TN];
          const double ux = dx * col_kr, uy = dy * col_kr;
          const double es = fm::exp_neg(-(ux * ux + uy * uy));
          s_col += es;
          const double wx = cw * dx * es, wy = cw * dy * es;
          sown[oa] += wx; sown[oa + TN] += wy;
          sacc[ob] -= wx; sacc[ob + TN] -= wy;
        }
      }
    }
    __syncthreads();
    if ((a.what & D2DX_EVAL_GRAD) && i < N) {
      double* go = a.grad + (size_t)prob * a.n_free;
      for (int a_l = al; a_l < n; a_l += a.APP) {
        go[(size_t)(3 * a_l) * N + i] = sown[(a_l * 2) * TN + il] + sacc[(a_l * 2) * TN + il];
        go[(size_t)(3 * a_l + 1) * N + i] = sown[(a_l * 2 + 1) * TN + il] + sacc[(a_l * 2 + 1) * TN + il];
      }
    }
  }

  // instance constraints (06_optyplan.py:46-49) and their unit Jacobian entries: first tile of each problem
  if (tile == 0) {
    for (int k = tid; k < P.n_inst; k += kCollocThreads) {
      if (a.what & D2DX_EVAL_RESIDUAL)
        a.res[(size_t)prob * a.n_con + (size_t)n * (N - 1) + k] = fr[(size_t)P.inst_var[k] * N + P.inst_node[k]] - P.inst_val[k];
      if (a.what & D2DX_EVAL_JAC) a.jac[(size_t)prob * a.nnz + (a.nnz - P.n_inst) + k] = 1.0;
    }
  }

  if (a.what & D2DX_EVAL_COST) {                   // deterministic reduction: warp shuffles -> per-warp partials -> fixed-order sum
    const double v4[4] = {warp_sum(s_v), warp_sum(s_phi), warp_sum(s_obs), warp_sum(s_col)};
    const int nparts = a.ntiles * (kCollocThreads / 32);
    const int lane = tid & 31;
    double* part = a.scratch + ((size_t)prob * nparts + tile * (kCollocThreads / 32) + (tid >> 5)) * 4;
    if (lane == 0) { part[0] = v4[0]; part[1] = v4[1]; part[2] = v4[2]; part[3] = v4[3]; }
    if (a.ticket_mode) {
      int last = 0;
      if (lane == 0) {
        __threadfence();
        last = atomicAdd(&a.tickets[prob % kMaxTickets], 1) == nparts - 1;
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        __threadfence();
        const double* all = a.scratch + (size_t)prob * nparts * 4;
        double t4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int t = lane; t < nparts; t += 32)
          for (int k = 0; k < 4; ++k) t4[k] += __ldcg(all + t * 4 + k);
        for (int k = 0; k < 4; ++k) t4[k] = warp_sum(t4[k]);
        if (lane == 0) {
          double c = norm_in * (P.kvel * t4[0] + P.kbank * t4[1]);
          if (use_obs) c += P.kobs * (sN * t4[2]);
          if (use_col) c += P.kcol * (sN * t4[3]);
          a.cost[prob] = c;
          a.tickets[prob % kMaxTickets] = 0;
        }
      }
    }
  }
}

// second pass of the cost for large batches: one warp per problem sums the per-warp partials in a fixed order
__global__ void __launch_bounds__(128) colloc_cost_kernel(int n_prob, int nparts, const double* __restrict__ scratch, double norm_in,
                                                          double sN, double kvel, double kbank, double kobs, double kcol,
                                                          int use_obs, int use_col, double* __restrict__ cost) {
  const int prob = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (prob >= n_prob) return;
  const double* all = scratch + (size_t)prob * nparts * 4;
  double t4[4] = {0.0, 0.0, 0.0, 0.0};
  for (int t = lane; t < nparts; t += 32)
    for (int k = 0; k < 4; ++k) t4[k] += all[t * 4 + k];
  for (int k = 0; k < 4; ++k) t4[k] = warp_sum(t4[k]);
  if (lane == 0) {
    double c = norm_in * (kvel * t4[0] + kbank * t4[1]);
    if (use_obs) c += kobs * (sN * t4[2]);
    if (use_col) c += kcol * (sN * t4[3]);
    cost[prob] = c;
  }
}

// COO structure (opty convention)
__global__ void colloc_structure_kernel(int n_ac, int N, int layout, int n_inst, const int32_t* __restrict__ perm_phi,
                                        const int32_t* __restrict__ perm_v, const int32_t* __restrict__ inst_var,
                                        const int32_t* __restrict__ inst_node, long nnz, int64_t* __restrict__ rows,
                                        int64_t* __restrict__ cols) {
  const int n = 3 * n_ac, q = 2 * n_ac;
  const long nnz_eom = nnz - n_inst;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < nnz; idx += (long)gridDim.x * blockDim.x) {
    if (idx >= nnz_eom) {
      const int k = (int)(idx - nnz_eom);
      rows[idx] = (int64_t)n * (N - 1) + k;
      cols[idx] = (int64_t)inst_var[k] * N + inst_node[k];
      continue;
    }
    if (layout == D2DX_JAC_OPTY_DENSE) {
      const int W = 2 * n + q;
      const long node = idx / ((long)n * W);
      const int rem = (int)(idx - node * (long)n * W), e = rem / W, c = rem - e * W;
      rows[idx] = (int64_t)e * (N - 1) + node;
      int64_t col;
      if (c < n) col = (int64_t)c * N + node + 1;
      else if (c < 2 * n) col = (int64_t)(c - n) * N + node;
      else col = (int64_t)(n + (c - 2 * n)) * N + node + 1;
      cols[idx] = col;
    } else {
      const long per_ac = 12L * (N - 1);
      const int a_l = (int)(idx / per_ac);
      const int k = (int)((idx - a_l * per_ac) / (N - 1));
      const long node = idx - a_l * per_ac - (long)k * (N - 1);
      const int eq = k / 4;
      const int bphi = perm_phi ? perm_phi[a_l] : a_l, bv = perm_v ? perm_v[a_l] : n_ac + a_l;
      // local column of entry k: eq0 [x_i, psi_i, x_p, v_i]  eq1 [y_i, psi_i, y_p, v_i]  eq2 [psi_i, psi_p, phi_i, v_i]
      const int8_t kind[12] = {0, 2, 3, 7, 1, 2, 4, 7, 2, 5, 6, 7};
      const int kd = kind[k];
      int64_t col;
      if (kd < 3) col = (int64_t)(3 * a_l + kd) * N + node + 1;
      else if (kd < 6) col = (int64_t)(3 * a_l + kd - 3) * N + node;
      else if (kd == 6) col = (int64_t)(n + bphi) * N + node + 1;
      else col = (int64_t)(n + bv) * N + node + 1;
      rows[idx] = (int64_t)(3 * a_l + eq) * (N - 1) + node;
      cols[idx] = col;
    }
  }
}

__global__ void colloc_init_dense_kernel(long nnz, int n_inst, int n_prob, double* __restrict__ jac) {
  const long total = nnz * n_prob;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x)
    jac[idx] = (idx % nnz) >= nnz - n_inst ? 1.0 : 0.0;
}

__global__ void pack_positions_kernel(int n_ac, int N, const double* __restrict__ fr, double* __restrict__ pos) {
  const long total = (long)n_ac * 2 * N;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int g = (int)(idx / (2L * N));
    const long rem = idx - (long)g * 2 * N;        // c*N + node ; x,y slices are adjacent in the planner layout
    pos[idx] = fr[(size_t)(3 * g) * N + rem];
  }
}

int colloc_resident_threads_per_sm() {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, colloc_kernel<false>, kCollocThreads, 0);
  return nb * kCollocThreads;
}

static bool enabled_h(double k) { return (k == k) && k != 0.0; }

static void tile_shape(int n_ac, int& TN, int& APP) {
  TN = 128;
  while (TN > 32 && kCollocThreads / TN < n_ac) TN >>= 1;
  APP = kCollocThreads / TN;
}

static int sizes(const d2dx_colloc_problem* p, int layout, int64_t* s3) {
  const int64_t n_ac = p->n_ac, N = p->N;
  s3[0] = 5 * n_ac * N;
  s3[1] = 3 * n_ac * (N - 1) + p->n_inst;
  s3[2] = (layout == D2DX_JAC_OPTY_DENSE ? (N - 1) * 3 * n_ac * 8 * n_ac : 12 * n_ac * (N - 1)) + p->n_inst;
  return D2DX_OK;
}

static int check_problem(const d2dx_colloc_problem* p, const char* who) {
  D2DX_CHECK_ARG(p, "%s: null problem", who);
  D2DX_CHECK_ARG(p->n_ac >= 1 && p->N >= 2 && p->h > 0, "%s: n_ac=%d N=%d h=%g", who, p->n_ac, p->N, p->h);
  D2DX_CHECK_ARG(p->n_inst >= 0 && (p->n_inst == 0 || (p->inst_var && p->inst_node && p->inst_val)), "%s: instance constraints incomplete", who);
  D2DX_CHECK_ARG(p->n_obs >= 0 && p->n_obs <= D2DX_MAX_OBSTACLES, "%s: n_obs=%d (max %d)", who, p->n_obs, D2DX_MAX_OBSTACLES);
  D2DX_CHECK_ARG(p->in_div >= 1, "%s: in_div=%d", who, p->in_div);
  return D2DX_OK;
}

static int launch_eval(d2dx_handle* h, const d2dx_colloc_problem* p, int n_prob, int n_total, int a_lo, const double* free_,
                       const double* pos_all, int layout, uint32_t what, double* residual, double* jac, double* cost,
                       double* grad, double* scratch, void* stream, const char* who) {
  if (int rc = check_problem(p, who)) return rc;
  D2DX_CHECK_ARG(h && free_ && n_prob >= 1 && n_prob <= kMaxTickets, "%s: n_prob=%d (1..%d per call)", who, n_prob, kMaxTickets);
  D2DX_CHECK_ARG(layout == D2DX_JAC_COMPACT || layout == D2DX_JAC_OPTY_DENSE, "%s: unknown Jacobian layout %d", who, layout);
  D2DX_CHECK_ARG(!(what & D2DX_EVAL_RESIDUAL) || residual, "%s: residual requested but NULL", who);
  D2DX_CHECK_ARG(!(what & D2DX_EVAL_JAC) || jac, "%s: jacobian requested but NULL", who);
  D2DX_CHECK_ARG(!(what & D2DX_EVAL_COST) || (cost && scratch), "%s: cost requested but cost/scratch NULL", who);
  D2DX_CHECK_ARG(!(what & D2DX_EVAL_GRAD) || grad, "%s: gradient requested but NULL", who);
  CollocArgs a;
  a.p = *p; a.n_prob = n_prob; a.layout = layout; a.what = what; a.free_ = free_;
  a.res = residual; a.jac = jac; a.cost = cost; a.grad = grad; a.scratch = scratch; a.tickets = h->done_counter;
  a.n_total = n_total; a.a_lo = a_lo; a.pos_all = pos_all;
  tile_shape(p->n_ac, a.TN, a.APP);
  a.ntiles = (p->N + a.TN - 1) / a.TN;
  int64_t s3[3];
  sizes(p, layout, s3);
  a.n_free = s3[0]; a.n_con = s3[1]; a.nnz = s3[2];
  D2DX_CUDA(cudaSetDevice(h->device));
  a.ticket_mode = n_prob < 64;                 // latency-sensitive single evaluations stay one launch
  const bool want_cg = (what & (D2DX_EVAL_COST | D2DX_EVAL_GRAD)) != 0;
  const bool extra = want_cg && ((enabled_h(p->kcol) && n_total > 1) || (enabled_h(p->kobs) && p->n_obs > 0));
  const unsigned grid = (unsigned)((long)n_prob * a.ntiles);
  if (extra) {
    const bool col = enabled_h(p->kcol) && n_total > 1;
    const bool sym = col && p->col_all_pairs && pos_all == nullptr;
    const size_t smem = col ? (size_t)n_total * 2 * a.TN * sizeof(double) * (sym ? 3 : 1) : 0;
    if (smem > 48 * 1024) {
      D2DX_CHECK_ARG(smem <= 200 * 1024, "%s: %d aircraft need %zu B of shared memory", who, n_total, smem);
      D2DX_CUDA(cudaFuncSetAttribute(colloc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    colloc_kernel<true><<<grid, kCollocThreads, smem, as_stream(stream)>>>(a);
  } else {
    colloc_kernel<false><<<grid, kCollocThreads, 0, as_stream(stream)>>>(a);
  }
  D2DX_LAUNCH_CHECK("colloc_kernel");
  if ((what & D2DX_EVAL_COST) && !a.ticket_mode) {
    const double sN = p->obj_scale / p->N;
    colloc_cost_kernel<<<(n_prob + 3) / 4, 128, 0, as_stream(stream)>>>(
        n_prob, a.ntiles * (kCollocThreads / 32), scratch, sN / p->in_div, sN, p->kvel, p->kbank, p->kobs, p->kcol,
        enabled_h(p->kobs) && p->n_obs > 0, enabled_h(p->kcol) && n_total > 1, cost);
    D2DX_LAUNCH_CHECK("colloc_cost_kernel");
  }
  return D2DX_OK;
}

}  // namespace d2dx

using namespace d2dx;

extern "C" {

int d2dx_colloc_sizes(const d2dx_colloc_problem* p, int32_t layout, int64_t* s3) {
  D2DX_CHECK_ARG(p && s3, "d2dx_colloc_sizes: null argument");
  return sizes(p, layout, s3);
}

int64_t d2dx_colloc_scratch_size(const d2dx_colloc_problem* p, int32_t n_prob) {
  if (!p || n_prob < 1) return 0;
  int TN, APP;
  tile_shape(p->n_ac, TN, APP);
  return (int64_t)n_prob * ((p->N + TN - 1) / TN) * (kCollocThreads / 32) * 4;
}

int d2dx_colloc_structure(d2dx_handle* h, const d2dx_colloc_problem* p, int32_t layout, int64_t* rows, int64_t* cols, void* stream) {
  if (int rc = check_problem(p, "d2dx_colloc_structure")) return rc;
  D2DX_CHECK_ARG(h && rows && cols, "d2dx_colloc_structure: null argument");
  int64_t s3[3];
  sizes(p, layout, s3);
  D2DX_CUDA(cudaSetDevice(h->device));
  const int grid = (int)((s3[2] + 255) / 256 < 4096 ? (s3[2] + 255) / 256 : 4096);
  colloc_structure_kernel<<<grid, 256, 0, as_stream(stream)>>>(p->n_ac, p->N, layout, p->n_inst, p->perm_phi, p->perm_v,
                                                                 p->inst_var, p->inst_node, s3[2], rows, cols);
  D2DX_LAUNCH_CHECK("colloc_structure_kernel");
  return D2DX_OK;
}

int d2dx_colloc_init_dense(d2dx_handle* h, const d2dx_colloc_problem* p, int32_t n_prob, double* jac, void* stream) {
  if (int rc = check_problem(p, "d2dx_colloc_init_dense")) return rc;
  D2DX_CHECK_ARG(h && jac && n_prob >= 1, "d2dx_colloc_init_dense: bad argument");
  int64_t s3[3];
  sizes(p, D2DX_JAC_OPTY_DENSE, s3);
  D2DX_CUDA(cudaSetDevice(h->device));
  colloc_init_dense_kernel<<<h->sm_count * 8, 256, 0, as_stream(stream)>>>(s3[2], p->n_inst, n_prob, jac);
  D2DX_LAUNCH_CHECK("colloc_init_dense_kernel");
  return D2DX_OK;
}

int d2dx_colloc_eval(d2dx_handle* h, const d2dx_colloc_problem* p, int32_t n_prob, const double* free_, int32_t layout,
                     uint32_t what, double* residual, double* jac, double* cost, double* grad, double* scratch, void* stream) {
  return launch_eval(h, p, n_prob, p ? p->n_ac : 0, 0, free_, nullptr, layout, what, residual, jac, cost, grad, scratch, stream,
                     "d2dx_colloc_eval");
}

int d2dx_colloc_eval_shard(d2dx_handle* h, const d2dx_colloc_problem* p, int32_t n_ac_total, int32_t a_lo,
                           const double* free_local, const double* pos_all, uint32_t what, double* residual, double* jac,
                           double* cost, double* grad, double* scratch, void* stream) {
  D2DX_CHECK_ARG(p && pos_all && n_ac_total >= p->n_ac && a_lo >= 0 && a_lo + p->n_ac <= n_ac_total,
                 "d2dx_colloc_eval_shard: shard [%d,+%d) of %d", a_lo, p ? p->n_ac : -1, n_ac_total);
  return launch_eval(h, p, 1, n_ac_total, a_lo, free_local, pos_all, D2DX_JAC_COMPACT, what, residual, jac, cost, grad, scratch,
                     stream, "d2dx_colloc_eval_shard");
}

int d2dx_colloc_pack_positions(d2dx_handle* h, int32_t n_ac, int32_t N, const double* free_local, double* pos, void* stream) {
  D2DX_CHECK_ARG(h && n_ac >= 1 && N >= 1 && free_local && pos, "d2dx_colloc_pack_positions: bad argument");
  D2DX_CUDA(cudaSetDevice(h->device));
  const long total = (long)n_ac * 2 * N;
  pack_positions_kernel<<<(int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048), 256, 0, as_stream(stream)>>>(n_ac, N, free_local, pos);
  D2DX_LAUNCH_CHECK("pack_positions_kernel");
  return D2DX_OK;
}

}  // extern "C"
