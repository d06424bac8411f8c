// Aircraft-sharded collocation evaluation in ONE kernel per rank, exchanging over NVLink peer memory (C4, one
// problem -- or a batch of problems -- split by aircraft over the GPUs of one box).
//
// Every rank owns an exchange buffer (cudaMalloc) that all other ranks map (CUDA IPC between processes, plain pointers
// inside one process).  One evaluation =
//   phase 0  each block stores its tiles of the OWNED aircraft's x, y straight from free_local into every peer's position
//            table, then (one fence, one flag per tile and peer) publishes them -- for ALL its tiles before anything else;
//   phase 1  residual / Jacobian / input cost / gradient of psi, phi, v of the owned aircraft (needs no remote data: this is
//            what hides the NVLink latency);
//   phase 2  wait for the same tile of every peer, stage all positions in shared memory;
//   phase 3  collision terms of the owned aircraft against every other aircraft, gradient of x, y;
//   phase 4  the last block of a problem sums this rank's partials, writes them to every peer (fence + flag) and adds up
//            the partials of all ranks in rank order: every rank ends with the same, deterministic total cost.
// There is no collective call, no pack kernel and no host synchronisation; flags carry a monotonically increasing
// evaluation number kept on the device, so the launch can be captured in a CUDA graph and replayed.  Phase 4 doubles as the
// barrier that keeps evaluation e + 1 from overwriting tables a slower rank still reads in evaluation e.
// A peer that never answers costs `spin_cycles` per wait and is reported by d2dx_peer_status (never a hang).
#include <string.h>

#include "d2dx_colloc_dev.cuh"
#include "d2dx_host.h"

struct d2dx_peer {
  int device, world, rank, max_prob, n_total, N, ntiles;
  unsigned char* local;                       // this rank's exchange buffer
  unsigned char* base[D2DX_PEER_MAX_WORLD];   // every rank's buffer as mapped here (base[rank] == local)
  bool ipc_opened[D2DX_PEER_MAX_WORLD];
  bool connected;
  size_t bytes, off_tickets, off_posflag, off_costflag, off_cpart, off_lpart, off_pos;
  int resident_blocks;
};

namespace d2dx {

constexpr int kPeerWarps = 8;

struct PeerCtrl { uint32_t epoch, done_blocks, timeouts, pad; };

struct PeerArgs {
  CollocArgs c;                  // c.p = the local shard as a problem of n_own aircraft; c.n_total, c.a_lo; outputs shard-local
  int world, rank, max_prob, ntiles;
  unsigned char* base[D2DX_PEER_MAX_WORLD];
  size_t off_tickets, off_posflag, off_costflag, off_cpart, off_lpart, off_pos;
  long long spin_cycles;
};

__device__ __forceinline__ void flag_store(uint32_t* p, uint32_t v) {      // publishes everything this thread has observed
  __threadfence_system();
  *reinterpret_cast<volatile uint32_t*>(p) = v;
}

// waits until *p has reached `epoch`; false on timeout
__device__ __forceinline__ bool flag_wait(const uint32_t* p, uint32_t epoch, long long spin_cycles) {
  const volatile uint32_t* vp = reinterpret_cast<const volatile uint32_t*>(p);
  const long long t0 = clock64();
  while ((int32_t)(*vp - epoch) < 0) {
    if (clock64() - t0 > spin_cycles) return false;
  }
  __threadfence_system();
  return true;
}

__global__ void __launch_bounds__(kPeerWarps * 32, 3) colloc_peer_kernel(const __grid_constant__ PeerArgs g) {
  extern __shared__ double sm[];                   // [n_total][2][32] positions, then [W][4] reduction scratch
  const CollocArgs& a = g.c;
  const d2dx_colloc_problem& P = a.p;
  const int N = P.N, n_own = P.n_ac, n_total = a.n_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
  double* spos = sm + lane;
  double* sred = sm + n_total * 64;
  unsigned char* mine = g.base[g.rank];
  PeerCtrl* ctrl = reinterpret_cast<PeerCtrl*>(mine);
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(&ctrl->epoch) + 1u;
  const bool want_cg = (a.what & (D2DX_EVAL_COST | D2DX_EVAL_GRAD)) != 0;
  const bool use_col = want_cg && enabled(P.kcol) && n_total > 1;
  const bool use_obs = want_cg && enabled(P.kobs) && P.n_obs > 0;
  const int n_items = a.n_prob * g.ntiles;
  int32_t* tickets = reinterpret_cast<int32_t*>(mine + g.off_tickets);
  const uint32_t* posflag_in = reinterpret_cast<const uint32_t*>(mine + g.off_posflag);
  const uint32_t* costflag_in = reinterpret_cast<const uint32_t*>(mine + g.off_costflag);
  const double* pos_in = reinterpret_cast<const double*>(mine + g.off_pos);
  const double* cpart_in = reinterpret_cast<const double*>(mine + g.off_cpart);
  double* lpart = reinterpret_cast<double*>(mine + g.off_lpart);

  // ---- sweep A: publish the owned positions of EVERY tile this block will process, one fence, then the flags; by the time
  // the block comes to a tile's collision terms the peers' stores for it have long landed (the NVLink latency of a batch
  // is paid once, not per tile) ----
  if (use_col && g.world > 1) {
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int prob = item / g.ntiles, tile = item - prob * g.ntiles;
      const int i = tile * 32 + lane;
      if (i < N) {
        const double* fr = a.free_ + (size_t)prob * a.n_free;
        for (int a_l = w; a_l < n_own; a_l += W) {
          const double x = fr[(3 * a_l) * N + i], y = fr[(3 * a_l + 1) * N + i];
          const size_t o = (((size_t)prob * n_total + a.a_lo + a_l) * 2) * N + i;
          for (int r = 0; r < g.world; ++r) {
            if (r == g.rank) continue;
            double* dst = reinterpret_cast<double*>(g.base[r] + g.off_pos);
            dst[o] = x; dst[o + N] = y;
          }
        }
      }
    }
    __syncthreads();
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    bool fenced = false;
    for (int idx = threadIdx.x; idx < my_items * g.world; idx += blockDim.x) {
      const int r = idx % g.world, item = blockIdx.x + (idx / g.world) * gridDim.x;
      if (r == g.rank) continue;
      if (!fenced) { __threadfence_system(); fenced = true; }
      const int prob = item / g.ntiles, tile = item - prob * g.ntiles;
      uint32_t* f = reinterpret_cast<uint32_t*>(g.base[r] + g.off_posflag);
      *reinterpret_cast<volatile uint32_t*>(f + ((size_t)g.rank * g.max_prob + prob) * g.ntiles + tile) = epoch;
    }
  }

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {   // same order on every rank: see the deadlock note in DESIGN
    const int prob = item / g.ntiles, tile = item - prob * g.ntiles;
    const int i = tile * 32 + lane;
    const bool valid = i < N;
    const double* fr = a.free_ + (size_t)prob * a.n_free;
    __syncthreads();                               // shared memory of the previous item is free

    if (use_col) {                                 // the owned positions of this tile into shared memory
      for (int a_l = w; a_l < n_own; a_l += W) {
        const int gl = a.a_lo + a_l;
        double x = 0.0, y = 0.0;
        if (valid) { x = fr[(3 * a_l) * N + i]; y = fr[(3 * a_l + 1) * N + i]; }
        spos[(gl * 2) * 32] = x; spos[(gl * 2 + 1) * 32] = y;
      }
    }

    // ---- phase 1: everything that needs no remote data ----
    double s_v = 0.0, s_phi = 0.0, s_obs = 0.0, s_col = 0.0;
    if (valid) {
      for (int a_l = w; a_l < n_own; a_l += W) {
        const int ox = 3 * a_l * N + i;
        colloc_node<false>(a, fr, prob, a_l, i, fr[ox], fr[ox + N], 0.0, 0.0, want_cg, s_v, s_phi);
      }
    }

    // ---- phase 2: the peers' positions of this tile ----
    if (use_col) {
      if (threadIdx.x < g.world && threadIdx.x != g.rank) {
        if (!flag_wait(posflag_in + ((size_t)threadIdx.x * g.max_prob + prob) * g.ntiles + tile, epoch, g.spin_cycles))
          atomicAdd(&ctrl->timeouts, 1u);
      }
      __syncthreads();
      for (int gl = w; gl < n_total; gl += W) {
        if (gl >= a.a_lo && gl < a.a_lo + n_own) continue;
        double x = 0.0, y = 0.0;
        if (valid) {
          const size_t o = (((size_t)prob * n_total + gl) * 2) * N + i;
          x = __ldcg(pos_in + o); y = __ldcg(pos_in + o + N);
        }
        spos[(gl * 2) * 32] = x; spos[(gl * 2 + 1) * 32] = y;
      }
      __syncthreads();
    }

    // ---- phase 3: obstacle and collision terms of the owned aircraft, gradient of x, y ----
    if (want_cg && valid) {
      for (int a_l = w; a_l < n_own; a_l += W) {
        const int gl = a.a_lo + a_l;
        const int ox = 3 * a_l * N + i;
        const double x = fr[ox], y = fr[ox + N];
        double gx = 0.0, gy = 0.0;
        if (use_obs && gl == 0) obstacle_terms(P, a.sN, x, y, s_obs, gx, gy);
        if (use_col) {
          const int b_lo = P.col_all_pairs ? 0 : (gl == 0 ? 1 : 0);
          const int b_hi = P.col_all_pairs ? n_total : (gl == 0 ? 2 : (gl == 1 ? 1 : 0));
          const double* pb = spos + (b_lo * 2) * 32;
#pragma unroll 4
          for (int b = b_lo; b < b_hi; ++b, pb += 64) {
            if (b == gl) continue;
            const double dx = x - pb[0], dy = y - pb[32];
            const double es = fm::exp_neg(a.nkr2 * fma(dx, dx, dy * dy));
            if (gl < b) s_col += es;               // each pair counted once, at the owner of its lower-index aircraft
            const double wgt = a.cw * es;
            gx = fma(wgt, dx, gx); gy = fma(wgt, dy, gy);
          }
        }
        if (a.what & D2DX_EVAL_GRAD) {
          double* go = a.grad + (size_t)prob * a.n_free;
          go[ox] = gx; go[ox + N] = gy;
        }
      }
    }

    // instance constraints of the owned aircraft: first tile of each problem
    if (tile == 0) {
      for (int k = threadIdx.x; k < P.n_inst; k += blockDim.x) {
        if (a.what & D2DX_EVAL_RESIDUAL)
          a.res[(size_t)prob * a.n_con + 3 * n_own * (N - 1) + k] = fr[P.inst_var[k] * N + P.inst_node[k]] - P.inst_val[k];
        if (a.what & D2DX_EVAL_JAC) a.jac[(size_t)prob * a.nnz + (a.nnz - P.n_inst) + k] = 1.0;
      }
    }

    // ---- phase 4: cost partials; the last block of the problem exchanges them (also the end-of-evaluation barrier) ----
    const double v4[4] = {warp_sum(s_v), warp_sum(s_phi), warp_sum(s_obs), warp_sum(s_col)};
    if (lane == 0) { sred[w * 4 + 0] = v4[0]; sred[w * 4 + 1] = v4[1]; sred[w * 4 + 2] = v4[2]; sred[w * 4 + 3] = v4[3]; }
    __syncthreads();
    if (w == 0) {
      double b4[4] = {0.0, 0.0, 0.0, 0.0};
      for (int q = 0; q < W; ++q) { b4[0] += sred[q * 4]; b4[1] += sred[q * 4 + 1]; b4[2] += sred[q * 4 + 2]; b4[3] += sred[q * 4 + 3]; }
      double* parts = lpart + (size_t)prob * g.ntiles * 4;
      if (lane == 0) { parts[tile * 4] = b4[0]; parts[tile * 4 + 1] = b4[1]; parts[tile * 4 + 2] = b4[2]; parts[tile * 4 + 3] = b4[3]; }
      int last = 0;
      if (lane == 0) {
        __threadfence();
        last = atomicAdd(&tickets[prob], 1) == g.ntiles - 1;
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last) {
        __threadfence();
        double t4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int t = lane; t < g.ntiles; t += 32)
          for (int k = 0; k < 4; ++k) t4[k] += __ldcg(parts + t * 4 + k);
        for (int k = 0; k < 4; ++k) t4[k] = warp_sum(t4[k]);
        if (lane == 0) tickets[prob] = 0;
        if (lane < g.world) {                      // lane r: this rank's four sums -> rank r, then the flag
          double* cp = reinterpret_cast<double*>(g.base[lane] + g.off_cpart) + ((size_t)g.rank * g.max_prob + prob) * 4;
          cp[0] = t4[0]; cp[1] = t4[1]; cp[2] = t4[2]; cp[3] = t4[3];
          flag_store(reinterpret_cast<uint32_t*>(g.base[lane] + g.off_costflag) + (size_t)g.rank * g.max_prob + prob, epoch);
          if (!flag_wait(costflag_in + (size_t)lane * g.max_prob + prob, epoch, g.spin_cycles)) atomicAdd(&ctrl->timeouts, 1u);
        }
        __syncwarp();
        if (lane == 0 && (a.what & D2DX_EVAL_COST)) {
          double s4[4] = {0.0, 0.0, 0.0, 0.0};
          for (int r = 0; r < g.world; ++r) {
            const double* cp = cpart_in + ((size_t)r * g.max_prob + prob) * 4;
            for (int k = 0; k < 4; ++k) s4[k] += __ldcg(cp + k);
          }
          a.cost[prob] = colloc_cost_from_sums(a, s4, use_obs, use_col);
        }
      }
    }
  }

  // the last block to leave advances the evaluation number (all blocks read it long ago)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctrl->done_blocks, 1u) == gridDim.x - 1) {
      ctrl->done_blocks = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(&ctrl->epoch) = epoch;
    }
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace d2dx

using namespace d2dx;

extern "C" {

int d2dx_peer_create(d2dx_handle* h, int32_t world, int32_t rank, int32_t max_prob, int32_t n_ac_total, int32_t N, d2dx_peer** out) {
  D2DX_CHECK_ARG(h && out, "d2dx_peer_create: null argument");
  D2DX_CHECK_ARG(world >= 1 && world <= D2DX_PEER_MAX_WORLD && rank >= 0 && rank < world, "d2dx_peer_create: rank %d of %d (max %d)", rank, world,
                 D2DX_PEER_MAX_WORLD);
  D2DX_CHECK_ARG(max_prob >= 1 && n_ac_total >= 1 && N >= 2, "d2dx_peer_create: max_prob=%d n_ac_total=%d N=%d", max_prob, n_ac_total, N);
  D2DX_CHECK_ARG((size_t)n_ac_total * 64 * sizeof(double) + kPeerWarps * 4 * sizeof(double) <= 200 * 1024,
                 "d2dx_peer_create: %d aircraft do not fit the shared-memory position tile", n_ac_total);
  D2DX_CUDA(cudaSetDevice(h->device));
  d2dx_peer* p = new d2dx_peer;
  memset(p, 0, sizeof(*p));
  p->device = h->device; p->world = world; p->rank = rank; p->max_prob = max_prob; p->n_total = n_ac_total; p->N = N;
  p->ntiles = (N + 31) / 32;
  size_t o = align_up(sizeof(PeerCtrl), 256);
  p->off_tickets = o; o = align_up(o + sizeof(int32_t) * max_prob, 256);
  p->off_posflag = o; o = align_up(o + sizeof(uint32_t) * (size_t)world * max_prob * p->ntiles, 256);
  p->off_costflag = o; o = align_up(o + sizeof(uint32_t) * (size_t)world * max_prob, 256);
  p->off_cpart = o; o = align_up(o + sizeof(double) * (size_t)world * max_prob * 4, 256);
  p->off_lpart = o; o = align_up(o + sizeof(double) * (size_t)max_prob * p->ntiles * 4, 256);
  p->off_pos = o; o = align_up(o + sizeof(double) * (size_t)max_prob * n_ac_total * 2 * N, 256);
  p->bytes = o;
  void* buf = nullptr;
  cudaError_t e = cudaMalloc(&buf, p->bytes);
  if (e != cudaSuccess) { delete p; return set_error(D2DX_ECUDA, "d2dx_peer_create: cudaMalloc(%zu): %s", o, cudaGetErrorString(e)); }
  e = cudaMemset(buf, 0, p->bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(buf); delete p; return set_error(D2DX_ECUDA, "d2dx_peer_create: cudaMemset: %s", cudaGetErrorString(e)); }
  p->local = static_cast<unsigned char*>(buf);
  p->base[rank] = p->local;
  p->connected = world == 1;
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, colloc_peer_kernel, kPeerWarps * 32, 0);
  p->resident_blocks = (nb > 0 ? nb : 1) * h->sm_count;
  *out = p;
  return D2DX_OK;
}

int d2dx_peer_ipc_handle(d2dx_peer* p, void* handle_host64) {
  D2DX_CHECK_ARG(p && handle_host64, "d2dx_peer_ipc_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == D2DX_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  D2DX_CUDA(cudaSetDevice(p->device));
  cudaIpcMemHandle_t hd;
  D2DX_CUDA(cudaIpcGetMemHandle(&hd, p->local));
  memcpy(handle_host64, &hd, sizeof(hd));
  return D2DX_OK;
}

int d2dx_peer_connect_ipc(d2dx_peer* p, const void* handles_host) {
  D2DX_CHECK_ARG(p && handles_host, "d2dx_peer_connect_ipc: null argument");
  D2DX_CUDA(cudaSetDevice(p->device));
  const unsigned char* hs = static_cast<const unsigned char*>(handles_host);
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, hs + (size_t)r * D2DX_IPC_HANDLE_BYTES, sizeof(hd));
    void* ptr = nullptr;
    D2DX_CUDA(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    p->base[r] = static_cast<unsigned char*>(ptr);
    p->ipc_opened[r] = true;
  }
  p->connected = true;
  return D2DX_OK;
}

int d2dx_peer_connect_local(d2dx_peer* p, d2dx_peer* const* peers_host) {
  D2DX_CHECK_ARG(p && peers_host, "d2dx_peer_connect_local: null argument");
  for (int r = 0; r < p->world; ++r) {
    D2DX_CHECK_ARG(peers_host[r] && peers_host[r]->rank == r && peers_host[r]->world == p->world && peers_host[r]->bytes == p->bytes,
                   "d2dx_peer_connect_local: entry %d is not rank %d of the same exchange", r, r);
    if (peers_host[r]->device != p->device) {
      D2DX_CUDA(cudaSetDevice(p->device));
      cudaError_t e = cudaDeviceEnablePeerAccess(peers_host[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return set_error(D2DX_ECUDA, "d2dx_peer_connect_local: no peer access %d -> %d: %s", p->device, peers_host[r]->device, cudaGetErrorString(e));
      cudaGetLastError();
    }
    p->base[r] = peers_host[r]->local;
  }
  p->connected = true;
  return D2DX_OK;
}

int d2dx_peer_status(d2dx_peer* p, int32_t* status_host4) {
  D2DX_CHECK_ARG(p && status_host4, "d2dx_peer_status: null argument");
  D2DX_CUDA(cudaSetDevice(p->device));
  PeerCtrl c;
  D2DX_CUDA(cudaMemcpy(&c, p->local, sizeof(c), cudaMemcpyDeviceToHost));
  status_host4[0] = (int32_t)c.timeouts; status_host4[1] = (int32_t)c.epoch; status_host4[2] = p->resident_blocks;
  status_host4[3] = (int32_t)(p->bytes >> 10);
  return D2DX_OK;
}

int d2dx_peer_destroy(d2dx_peer* p) {
  if (!p) return D2DX_OK;
  cudaSetDevice(p->device);
  for (int r = 0; r < p->world; ++r)
    if (p->ipc_opened[r] && p->base[r]) cudaIpcCloseMemHandle(p->base[r]);
  if (p->local) cudaFree(p->local);
  delete p;
  return D2DX_OK;
}

int d2dx_colloc_eval_peer(d2dx_handle* h, d2dx_peer* peer, const d2dx_colloc_problem* p, int32_t n_prob, int32_t a_lo,
                          const double* free_local, uint32_t what, double* residual, double* jac, double* cost, double* grad,
                          void* stream) {
  D2DX_CHECK_ARG(h && peer && p && free_local, "d2dx_colloc_eval_peer: null argument");
  D2DX_CHECK_ARG(peer->connected, "d2dx_colloc_eval_peer: the exchange is not connected (d2dx_peer_connect_ipc / _local)");
  D2DX_CHECK_ARG(p->N == peer->N && n_prob >= 1 && n_prob <= peer->max_prob, "d2dx_colloc_eval_peer: N=%d (exchange %d), n_prob=%d (max %d)", p->N,
                 peer->N, n_prob, peer->max_prob);
  D2DX_CHECK_ARG(p->n_ac >= 1 && a_lo >= 0 && a_lo + p->n_ac <= peer->n_total, "d2dx_colloc_eval_peer: shard [%d,+%d) of %d", a_lo, p->n_ac,
                 peer->n_total);
  D2DX_CHECK_ARG(p->h > 0 && p->in_div >= 1 && p->n_obs >= 0 && p->n_obs <= D2DX_MAX_OBSTACLES, "d2dx_colloc_eval_peer: bad problem description");
  D2DX_CHECK_ARG(!(what & D2DX_EVAL_RESIDUAL) || residual, "d2dx_colloc_eval_peer: residual requested but NULL");
  D2DX_CHECK_ARG(!(what & D2DX_EVAL_JAC) || jac, "d2dx_colloc_eval_peer: jacobian requested but NULL");
  D2DX_CHECK_ARG(!(what & D2DX_EVAL_COST) || cost, "d2dx_colloc_eval_peer: cost requested but NULL");
  D2DX_CHECK_ARG(!(what & D2DX_EVAL_GRAD) || grad, "d2dx_colloc_eval_peer: gradient requested but NULL");
  PeerArgs g;
  memset(&g, 0, sizeof(g));
  CollocArgs& a = g.c;
  a.p = *p; a.n_prob = n_prob; a.layout = D2DX_JAC_COMPACT; a.what = what; a.free_ = free_local;
  a.res = residual; a.jac = jac; a.cost = cost; a.grad = grad; a.scratch = nullptr;
  a.n_total = peer->n_total; a.a_lo = a_lo; a.pos_all = nullptr;
  a.TN = 32; a.APP = kPeerWarps; a.ntiles = peer->ntiles; a.nparts = peer->ntiles; a.ticket_mode = 1;
  a.n_free = 5 * p->n_ac * p->N; a.n_con = 3 * p->n_ac * (p->N - 1) + p->n_inst; a.nnz = 12L * p->n_ac * (p->N - 1) + p->n_inst;
  colloc_constants(a);
  g.world = peer->world; g.rank = peer->rank; g.max_prob = peer->max_prob; g.ntiles = peer->ntiles;
  for (int r = 0; r < peer->world; ++r) g.base[r] = peer->base[r];
  g.off_tickets = peer->off_tickets; g.off_posflag = peer->off_posflag; g.off_costflag = peer->off_costflag;
  g.off_cpart = peer->off_cpart; g.off_lpart = peer->off_lpart; g.off_pos = peer->off_pos;
  g.spin_cycles = 2000000000LL;                  // ~1 s at 1.9 GHz
  D2DX_CUDA(cudaSetDevice(h->device));
  const size_t smem = ((size_t)peer->n_total * 64 + kPeerWarps * 4) * sizeof(double);
  if (smem > 48 * 1024) D2DX_CUDA(cudaFuncSetAttribute(colloc_peer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = n_prob * peer->ntiles;
  const int grid = items < peer->resident_blocks ? items : peer->resident_blocks;
  const int warps = p->n_ac < kPeerWarps ? p->n_ac : kPeerWarps;
  colloc_peer_kernel<<<grid, warps * 32, smem, as_stream(stream)>>>(g);
  D2DX_LAUNCH_CHECK("colloc_peer_kernel");
  return D2DX_OK;
}

}  // extern "C"
