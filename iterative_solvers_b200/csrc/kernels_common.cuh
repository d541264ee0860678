// Shared device helpers of libb200cg: modes, kernel arguments, reductions, scalar finalisation (sm_100a, fp64).
//
// The two-sweep scheme (stream_kernel.cuh; the single-sweep iteration of fused_kernel.cuh is the default where it applies) -
// two fused kernels per CG iteration (DESIGN.md "Kernels"):
//   dot phase    : p = r + beta*p_old on the fly, Ap = A p on the fly, reduces p.Ap and r.p       16 B/unknown
//   update phase : same p / Ap recomputed, x += alpha p, r -= alpha Ap, stores x, r, p,
//                  reduces r.r, |r|_inf, |dx|_inf (and |x-u|_inf)                                 48 B/unknown
// Ap is never stored and p is never re-read: 64 B per unknown per iteration instead of the 80 B of the
// store-Ap formulation. Every element-wise operation uses separately rounded multiplies and adds
// in the reference's order (matrix_free_system.cpp:216-266, :422-438), so the iterates differ from the
// reference's only through the summation order of the dot products.
#pragma once
#include <float.h>
#ifndef B200CG_STORE_HINT
#define B200CG_STORE_HINT 1
#endif
#include "common.cuh"

namespace b200cg {

enum { MODE_DOT = 0, MODE_UPD = 1, MODE_APPLY = 2 };
enum {
  F_U = 1,       // UPD / APPLY-report: also read the true solution u
  F_REPORT = 2,  // UPD: reduce |dx|_2, |x-u|_2.  APPLY: reduce |b - A v|_2, append the callback record, no store
  F_SUB_B = 4,   // APPLY: out = A v - b
  F_NOX = 8,     // UPD, x-deferral: even iteration, x is not touched (its update stays pending)
  F_X2 = 16,     // UPD, x-deferral: odd iteration, applies the pending update and this one
  F_MAXN = 32,   // single-sweep kernel under MSGSolver's max-norm rules: x every iteration, |r'|_inf, |dx|_inf, |x-u|_inf
  F_SHARD = 64   // single-sweep kernel on a sharded plan: two halo rows per side over peer memory
};

struct TileArgs {
  const double* r_in;  // DOT/UPD: residual.  APPLY with F_SUB_B / F_REPORT: rhs b
  const double* p_in;  // DOT/UPD: previous direction.  APPLY: input vector v
  double* x;           // UPD
  double* r_out;       // UPD
  double* p_out;       // UPD
  const double* u;     // F_U
  double* out;         // APPLY
  DevState* st;
  double* partials;    // [MAX_PARTIALS][gridDim.x]
  CbRecord* cb_log;
  const Tile* tiles;      // tile table, grouped per CTA
  const int* cta_begin;   // [gridDim.x + 1] offsets into tiles
  int defer;           // sharded plan: 1 = leave this rank's totals in st->loc_* for the NCCL all-reduce,
                       // 2 = publish them into every rank's PeerSync block over NVLink (peer-memory path)
  const PeerLinks* peers;  // defer == 2
  // peer-memory halo: UPD writes its first / last owned row of r' and p straight into the neighbours' halo rows
  double* nb_r_below;  // neighbour below: start of its top halo row in r_out (null: no neighbour / NCCL path)
  double* nb_p_below;
  double* nb_r_above;  // neighbour above: start of its bottom halo row
  double* nb_p_above;
  // single-sweep kernel on sharded plans: the neighbours' second halo rows (the two extra rows behind their stored rows)
  double* nb_r_below2;
  double* nb_p_below2;
  double* nb_r_above2;
  double* nb_p_above2;
  unsigned long long* cta_clock;  // [2 * gridDim.x] globaltimer at CTA start / end of its sweep (load balancing)
  const int* stop_flag;  // mapped host flag (requestStop, msg_solver.cpp:82), polled every STOP_POLL_EVERY-th iteration
  unsigned long long* peer_trace;  // diagnostics (B200CG_PEER_TRACE=1): [PEER_TRACE_CAP][4] globaltimer stamps of the
                                   // single sweep's cross-rank step - local sums ready, published, all flags seen, scalars formed
  Geom g;
};

// --------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }
// sweep outputs: written once, next read a whole sweep (GBs) later
__device__ __forceinline__ void st2_out(double* p, double2 v) {
#if B200CG_STORE_HINT == 1
  __stcs(reinterpret_cast<double2*>(p), v);  // st.global.cs: evict-first
#elif B200CG_STORE_HINT == 2
  __stwt(reinterpret_cast<double2*>(p), v);  // st.global.wt: write-through
#else
  *reinterpret_cast<double2*>(p) = v;
#endif
}

// The same store under a predicate, without a branch (the compiler wraps a conditional group of 16-byte stores in a
// divergent region: BSSY / BRA / BSYNC per row in the hot loop).
__device__ __forceinline__ void st2_out_if(bool pred, double* p, double2 v) {
#if B200CG_STORE_HINT == 1
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q st.global.cs.v2.f64 [%0], {%1, %2};\n\t}" ::"l"(p),
      "d"(v.x), "d"(v.y), "r"((int)pred)
      : "memory");
#else
  if (pred) st2_out(p, v);
#endif
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reduction of NS sums and NM maxima (fixed tree: deterministic). Result valid in thread 0.
template <int NS, int NM>
__device__ __forceinline__ void block_reduce(double (&s)[NS > 0 ? NS : 1], double (&mx)[NM > 0 ? NM : 1],
                                             double* scratch /* [(NS+NM) * 32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = warp_sum(s[k]);
#pragma unroll
  for (int k = 0; k < NM; ++k) mx[k] = warp_max(mx[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) scratch[k * 32 + warp] = s[k];
#pragma unroll
    for (int k = 0; k < NM; ++k) scratch[(NS + k) * 32 + warp] = mx[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      double v = lane < nwarp ? scratch[k * 32 + lane] : 0.0;
      s[k] = warp_sum(v);
    }
#pragma unroll
    for (int k = 0; k < NM; ++k) {
      double v = lane < nwarp ? scratch[(NS + k) * 32 + lane] : 0.0;
      mx[k] = warp_max(v);
    }
  }
  __syncthreads();
}

// Grid-wide reduction: every CTA publishes its partials; the last CTA to arrive (ticket) sums them in index
// order with a fixed tree, so the result does not depend on which CTA is last. Returns true in thread 0 of
// that CTA with the totals in s / mx. Scalars never leave the device.
template <int NS, int NM>
__device__ __forceinline__ bool grid_reduce(double (&s)[NS > 0 ? NS : 1], double (&mx)[NM > 0 ? NM : 1],
                                            double* partials, DevState* st, double* scratch) {
  __shared__ bool is_last;
  block_reduce<NS, NM>(s, mx, scratch);
  const unsigned int nb = gridDim.x;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) partials[(size_t)k * nb + blockIdx.x] = s[k];
#pragma unroll
    for (int k = 0; k < NM; ++k) partials[(size_t)(NS + k) * nb + blockIdx.x] = mx[k];
    __threadfence();
    unsigned int t = atomicAdd(&st->ticket, 1u);
    is_last = (t == nb - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    double v = 0.0;
    for (unsigned int i = threadIdx.x; i < nb; i += blockDim.x) v += __ldcg(&partials[(size_t)k * nb + i]);
    s[k] = v;
  }
#pragma unroll
  for (int k = 0; k < NM; ++k) {
    double v = 0.0;
    for (unsigned int i = threadIdx.x; i < nb; i += blockDim.x)
      v = fmax(v, __ldcg(&partials[(size_t)(NS + k) * nb + i]));
    mx[k] = v;
  }
  block_reduce<NS, NM>(s, mx, scratch);
  if (threadIdx.x == 0) st->ticket = 0u;
  return threadIdx.x == 0;
}

__device__ __forceinline__ void append_record(DevState* st, CbRecord* log, double it, double p, double r,
                                              double e) {
  unsigned int k = st->n_log;
  CbRecord rec;
  rec.it = it; rec.precision = p; rec.residual = r; rec.error = e;
  if (log) log[k % CB_LOG_CAP] = rec;  // (cluster kernel: only CTA 0 owns the log)
  st->n_log = k + 1;
}

// requestStop (msg_solver.cpp:82 polls its flag every iteration): the thread that forms the scalars looks at the host's
// mapped flag every STOP_POLL_EVERY-th iteration (a read over PCIe costs ~2 us), so a stop lands within that many
// iterations instead of at the next graph boundary. Call before st->it is advanced; on a sharded plan every rank
// looks at the same iterations and the maximum over the ranks decides, so all ranks stop together.
constexpr int STOP_POLL_EVERY = 16;
constexpr int PEER_TRACE_CAP = 4096;  // iterations kept by the peer-exchange trace (ring)
__device__ __forceinline__ bool poll_stop(const DevState* st, const int* stop_flag) {
  if (!stop_flag || ((st->it + 1) % STOP_POLL_EVERY) != 0) return false;
  return *reinterpret_cast<const volatile int*>(stop_flag) != 0;
}
__device__ __forceinline__ void apply_stop(DevState* st, bool stop_req) {
  if (stop_req && !st->done) {  // (a stop rule that fired in the same iteration wins)
    st->done = 1;
    st->converged = 0;
    st->stop_reason = 4;  // B200CG_STOP_INTERRUPTED
  }
}

// alpha = r.r / p.Ap (matrix_free_system.cpp:417-419) or r.z / Az.z (msg_solver.cpp:96-102)
__device__ __forceinline__ void finalize_dot(DevState* st, double pAp, double rz) {
  st->pAp = pAp;
  st->rz = rz;
  st->alpha = (st->rule == 0) ? st->rr / pAp : rz / pAp;
}

// iteration_callback(iterations, precision, residual_norm, error_norm), matrix_free_system.cpp:457-468
__device__ __forceinline__ void finalize_report(DevState* st, CbRecord* log, double res2, double err2, bool has_u) {
  st->res_l2 = sqrt(res2);
  if (has_u) st->err_l2 = sqrt(err2);
  append_record(st, log, (double)(st->it - 1), st->dx_l2, st->res_l2, st->err_l2);
  st->report_pending = 0;
}

// Peer-memory publication of a sweep kernel's totals (thread 0 of the last CTA): values first, then - after a
// system-scope fence - the epoch flag, into every rank's PeerSync block.
template <int NS, int NM>
__device__ __forceinline__ void peer_publish(const PeerLinks* pl, DevState* st, int phase, const double (&s)[NS > 0 ? NS : 1],
                                             const double (&mx)[NM > 0 ? NM : 1], bool stop_req = false) {
  static_assert(NM <= 3, "max slot 3 carries the stop request");
  const unsigned long long epoch = st->epoch[phase] + 1ull;
  for (int dst = 0; dst < pl->world; ++dst) {
    double* v = pl->sync[dst]->vals[phase][pl->rank];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = k < NS ? s[k < NS ? k : 0] : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) v[4 + k] = k < NM ? mx[k < NM ? k : 0] : 0.0;
    v[7] = stop_req ? 1.0 : 0.0;
  }
  __threadfence_system();
  for (int dst = 0; dst < pl->world; ++dst)
    *reinterpret_cast<volatile unsigned long long*>(&pl->sync[dst]->flag[phase][pl->rank]) = epoch;
}

// x-deferral bookkeeping after an update phase: a NOX iteration leaves x += alpha*p pending.
__device__ __forceinline__ void note_x_deferral(DevState* st, int flags) {
  if (flags & F_NOX) {
    st->alpha_prev = st->alpha;
    st->x_pending = 1;
  } else {
    st->x_pending = 0;
  }
}

// Stop rules, evaluated by one thread after the update phase.
__device__ __forceinline__ void finalize_update(DevState* st, CbRecord* log, double rr_new, double r_max,
                                                double dx_max, double err_max, double dx2, double err2,
                                                bool report) {
  const int it = st->it + 1;
  st->it = it;
  const double r_norm = sqrt(rr_new);
  st->r_norm = r_norm;
  st->r_max = r_max;
  st->dx_max = dx_max;
  if (st->has_u) st->err_max = err_max;
  if (report) {
    st->dx_l2 = sqrt(dx2);
    st->err_l2 = sqrt(err2);
    st->report_pending = 1;  // a report kernel follows
  }
  if (st->rule == 0) {
    // MatrixFreeSolver, matrix_free_system.cpp:409,432-441,472
    st->beta = rr_new / st->rr;
    st->rr = rr_new;
    const bool go = (it < st->max_it) && (r_norm > st->eps_rel * st->r0_norm);
    if (!go) {
      st->done = 1;
      st->converged = (r_norm <= st->eps_rel * st->r0_norm) ? 1 : 0;
      st->stop_reason = st->converged ? 2 : 0;
    }
  } else {
    // MSGSolver, msg_solver.cpp:144-183
    int done = 0;
    if (st->eps_p > 0 && dx_max < st->eps_p) { done = 1; st->converged = 1; st->stop_reason = 1; }
    else if (st->eps_r > 0 && r_max < st->eps_r) { done = 1; st->converged = 1; st->stop_reason = 2; }
    else if (st->eps_e > 0 && st->has_u && err_max < st->eps_e) { done = 1; st->converged = 1; st->stop_reason = 3; }
    if (!done) {
      st->beta = (r_norm * r_norm) / st->rz;
      st->rr = rr_new;
      if (st->callback_every > 0 && (it % st->callback_every == 0 || it == 1))
        append_record(st, log, (double)it, dx_max, r_max, st->err_max);
      if (it >= st->max_it) { done = 1; st->converged = 0; st->stop_reason = 0; }
    }
    st->done = done;
  }
}

// Peer-memory path, second half (same thread, right after peer_publish): wait until every rank's publication of
// this epoch has landed in OUR PeerSync block, reduce the slots in rank order (identical on every rank) and form the
// scalars. which: 1 = dot phase, 2 = update phase. Every rank publishes before it waits, so the wait cannot
// deadlock; a flag that does not arrive within PEER_TIMEOUT_NS ends the solve with comm_error.
constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void peer_finalize(DevState* st, CbRecord* log, const PeerLinks* pl, int which, int flags) {
  const int phase = which - 1;
  const unsigned long long epoch = st->epoch[phase] + 1ull;
  const PeerSync* mine = pl->sync[pl->rank];
  bool ok = true;
  const unsigned long long t0 = global_ns();
  for (int r = 0; r < pl->world && ok; ++r) {
    const volatile unsigned long long* f = &mine->flag[phase][r];
    while (*f < epoch) {
      if (global_ns() - t0 > PEER_TIMEOUT_NS) {
        ok = false;
        break;
      }
    }
  }
  __threadfence_system();
  st->epoch[phase] = epoch;
  if (!ok) {
    st->comm_error = 1;
    st->done = 1;
    st->converged = 0;
    return;
  }
  double s[4] = {0.0, 0.0, 0.0, 0.0}, mx[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = 0; r < pl->world; ++r) {
    const volatile double* v = mine->vals[phase][r];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[k] += v[k];
      mx[k] = fmax(mx[k], v[4 + k]);
    }
  }
  const bool has_u = (flags & F_U) != 0, report = (flags & F_REPORT) != 0;
  if (which == 1) {
    finalize_dot(st, s[0], s[1]);
  } else {
    finalize_update(st, log, s[0], mx[0], mx[1], has_u ? mx[2] : DBL_MAX, report ? s[1] : 0.0,
                    (report && has_u) ? s[2] : 0.0, report);
    note_x_deferral(st, flags);
    apply_stop(st, mx[3] > 0.0);
  }
}

// The wait half of peer_finalize alone (single-sweep kernel): true when every rank's publication of this epoch of
// `phase` has landed; the summed slots come back in s[0..3]. On a timeout the solve is ended with comm_error.
__device__ __forceinline__ bool peer_collect(DevState* st, const PeerLinks* pl, int phase, double (&s)[4], bool* stop_req,
                                             double* mx = nullptr /* [3]: the maxima over the ranks of slots 4..6 */) {
  const unsigned long long epoch = st->epoch[phase] + 1ull;
  const PeerSync* mine = pl->sync[pl->rank];
  bool ok = true;
  const unsigned long long t0 = global_ns();
  for (int r = 0; r < pl->world && ok; ++r) {
    const volatile unsigned long long* f = &mine->flag[phase][r];
    while (*f < epoch) {
      if (global_ns() - t0 > PEER_TIMEOUT_NS) {
        ok = false;
        break;
      }
    }
  }
  __threadfence_system();
  st->epoch[phase] = epoch;
  if (!ok) {
    st->comm_error = 1;
    st->done = 1;
    st->converged = 0;
    return false;
  }
  s[0] = s[1] = s[2] = s[3] = 0.0;
  double stop = 0.0, m[3] = {0.0, 0.0, 0.0};
  for (int r = 0; r < pl->world; ++r) {
    const volatile double* v = mine->vals[phase][r];
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] += v[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) m[k] = fmax(m[k], v[4 + k]);
    stop = fmax(stop, v[7]);
  }
  if (mx) {
    mx[0] = m[0];
    mx[1] = m[1];
    mx[2] = m[2];
  }
  *stop_req = stop > 0.0;
  return true;
}

}  // namespace b200cg
