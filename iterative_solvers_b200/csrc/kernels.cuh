// Device kernels of libb200cg: the matrix-free CG hot path (sm_100a, fp64).
//
// Two fused kernels per CG iteration (DESIGN.md "Kernels"):
//   dot phase    : p = r + beta*p_old on the fly, Ap = A p on the fly, reduces p.Ap and r.p       16 B/unknown
//   update phase : same p / Ap recomputed, x += alpha p, r -= alpha Ap, stores x, r, p,
//                  reduces r.r, |r|_inf, |dx|_inf (and |x-u|_inf)                                 48 B/unknown
// Ap is never stored and p is never re-read: 64 B per unknown per iteration instead of the 80 B of the
// store-Ap formulation. Every element-wise operation uses separately rounded multiplies and adds
// in the reference's order (matrix_free_system.cpp:216-266, :422-438), so the iterates differ from the
// reference's only through the summation order of the dot products.
#pragma once
#include <float.h>
#include "common.cuh"

namespace b200cg {

enum { MODE_DOT = 0, MODE_UPD = 1, MODE_APPLY = 2 };
enum {
  F_U = 1,       // UPD / APPLY-report: also read the true solution u
  F_REPORT = 2,  // UPD: reduce |dx|_2, |x-u|_2.  APPLY: reduce |b - A v|_2, append the callback record, no store
  F_SUB_B = 4    // APPLY: out = A v - b
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
  int defer;           // sharded plan: publish this rank's totals in st->loc_*, finalize after the all-reduce
  Geom g;
};

// --------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ double2 ld_ro2(const double* p, bool ok) {
  // read-only for the lifetime of the kernel: non-coherent path
  return ok ? __ldg(reinterpret_cast<const double2*>(p)) : make_double2(0.0, 0.0);
}
__device__ __forceinline__ double2 ld_rw2(const double* p, bool ok) {
  return ok ? *reinterpret_cast<const double2*>(p) : make_double2(0.0, 0.0);
}
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

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
  log[k % CB_LOG_CAP] = rec;
  st->n_log = k + 1;
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

// --------------------------------------------------------------------------------------------- tile kernel
// PF = rows prefetched ahead (register ring).
template <int MODE, int PF, int FLAGS>
__global__ void __launch_bounds__(CTA_THREADS) cg_tile_kernel(const TileArgs a) {
  constexpr bool LOAD_R = (MODE != MODE_APPLY) || (FLAGS & (F_SUB_B | F_REPORT));
  constexpr bool LOAD_X = (MODE == MODE_UPD);
  constexpr bool LOAD_U = (FLAGS & F_U) != 0;
  constexpr bool REPORT = (FLAGS & F_REPORT) != 0;
  constexpr int NS = (MODE == MODE_DOT) ? 2 : (MODE == MODE_UPD ? (REPORT ? 3 : 1) : (REPORT ? 2 : 0));
  constexpr int NM = (MODE == MODE_UPD) ? (LOAD_U ? 3 : 2) : 0;

  const Geom& g = a.g;
  DevState* st = a.st;
  if (MODE == MODE_APPLY) {
    if (REPORT && st->report_pending == 0) return;  // no report pending
  } else {
    if (st->done) return;
  }

  __shared__ __align__(16) double sp[2][STRIP_LOAD];
  __shared__ double scratch[(NS + NM > 0 ? NS + NM : 1) * 32];

  // ---- tile decode: block B tiles first, then block U; strips of one row chunk are adjacent in blockIdx
  int strip, ya, yb, xlo;
  {
    const int tile = blockIdx.x;
    if (tile < g.tilesB) {
      const int nsB = g.strips - g.stripB0;
      const int chunk = tile / nsB;
      strip = g.stripB0 + (tile - chunk * nsB);
      ya = g.yB0 + chunk * g.tile_rows;
      yb = min(ya + g.tile_rows, g.yB1);
      xlo = g.xsplit + 1;
    } else {
      const int t = tile - g.tilesB;
      const int chunk = t / g.strips;
      strip = t - chunk * g.strips;
      ya = g.yU0 + chunk * g.tile_rows;
      yb = min(ya + g.tile_rows, g.yU1);
      xlo = 1;
    }
  }
  const int tid = threadIdx.x;
  const int col = strip * STRIP_OUT + 2 * tid;  // storage column of node x0 = col - XOFF
  const int x0 = col - XOFF;
  const bool ld_ok = col < g.pitch;
  const bool is_out = (tid >= STRIP_HALO / 2) && (tid < CTA_THREADS - STRIP_HALO / 2);
  const bool v0 = is_out && (x0 >= xlo) && (x0 <= g.n - 1);
  const bool v1 = is_out && (x0 + 1 >= xlo) && (x0 + 1 <= g.n - 1);
  const bool st_ok = v0 || v1;

  const double cA = g.A, cxk = g.xk, cyk = g.yk;
  double alpha = 0.0, beta = 0.0;
  if (MODE != MODE_APPLY) {
    beta = st->beta;
    if (MODE == MODE_UPD) alpha = st->alpha;
  }

  const size_t pitch = (size_t)g.pitch;
  const int S = yb - ya + 2;  // rows ya-1 .. yb
  size_t off = (size_t)(ya - 1 - g.ybase) * pitch + (size_t)col;  // offset of the next row to load

  double2 qr[PF], qp[PF], qx[PF], qu[PF];
#pragma unroll
  for (int j = 0; j < PF; ++j) {
    const bool ok = ld_ok && (j < S);
    const bool inner = ok && (j >= 1) && (j < S - 1);
    qp[j] = ld_ro2(a.p_in + off, ok);
    if (LOAD_R) qr[j] = ld_ro2(a.r_in + off, ok);
    if (LOAD_X) qx[j] = ld_rw2(a.x + off, inner);
    if (LOAD_U) qu[j] = ld_ro2(a.u + off, inner);
    off += pitch;
  }

  double2 pm = make_double2(0.0, 0.0), pc = make_double2(0.0, 0.0);
  double2 r_prev = make_double2(0.0, 0.0), x_prev = make_double2(0.0, 0.0), u_prev = make_double2(0.0, 0.0);
  double Lp = 0.0, Rp = 0.0;
  double acc_s[NS > 0 ? NS : 1] = {0.0};
  double acc_m[NM > 0 ? NM : 1] = {0.0};
  size_t eoff = (size_t)(ya - g.ybase) * pitch + (size_t)col;  // offset of the next row to emit

  for (int base = 0; base < S; base += PF) {
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int s = base + j;
      if (s >= S) break;
      const double2 cur_p = qp[j];
      double2 cur_r = make_double2(0.0, 0.0), cur_x = make_double2(0.0, 0.0), cur_u = make_double2(0.0, 0.0);
      if (LOAD_R) cur_r = qr[j];
      if (LOAD_X) cur_x = qx[j];
      if (LOAD_U) cur_u = qu[j];
      {  // refill the slot with row s + PF
        const int sn = s + PF;
        const bool ok = ld_ok && (sn < S);
        const bool inner = ok && (sn < S - 1);
        qp[j] = ld_ro2(a.p_in + off, ok);
        if (LOAD_R) qr[j] = ld_ro2(a.r_in + off, ok);
        if (LOAD_X) qx[j] = ld_rw2(a.x + off, inner);
        if (LOAD_U) qu[j] = ld_ro2(a.u + off, inner);
        off += pitch;
      }
      // direction of this row: p = r + beta * p_old (matrix_free_system.cpp:436-438)
      double2 pn;
      if (MODE == MODE_APPLY) {
        pn = cur_p;
      } else {
        pn.x = __dadd_rn(cur_r.x, __dmul_rn(beta, cur_p.x));
        pn.y = __dadd_rn(cur_r.y, __dmul_rn(beta, cur_p.y));
      }
      double* srow = sp[s & 1];
      st2(srow + 2 * tid, pn);
      __syncthreads();
      const double L = (tid > 0) ? srow[2 * tid - 1] : 0.0;
      const double R = (tid < CTA_THREADS - 1) ? srow[2 * tid + 2] : 0.0;

      if (s >= 2) {
        // row y-1: centre pc, bottom pm, top pn; accumulation order diag, left, right, top, bottom
        // (matrix_free_system.cpp:216-266), each term a rounded multiply then a rounded add.
        double ap0 = __dmul_rn(cA, pc.x);
        ap0 = __dadd_rn(ap0, __dmul_rn(cxk, Lp));
        ap0 = __dadd_rn(ap0, __dmul_rn(cxk, pc.y));
        ap0 = __dadd_rn(ap0, __dmul_rn(cyk, pn.x));
        ap0 = __dadd_rn(ap0, __dmul_rn(cyk, pm.x));
        double ap1 = __dmul_rn(cA, pc.y);
        ap1 = __dadd_rn(ap1, __dmul_rn(cxk, pc.x));
        ap1 = __dadd_rn(ap1, __dmul_rn(cxk, Rp));
        ap1 = __dadd_rn(ap1, __dmul_rn(cyk, pn.y));
        ap1 = __dadd_rn(ap1, __dmul_rn(cyk, pm.y));
        const double p0 = v0 ? pc.x : 0.0, p1 = v1 ? pc.y : 0.0;
        if (MODE == MODE_DOT) {
          acc_s[0] = fma(p0, ap0, acc_s[0]);
          acc_s[0] = fma(p1, ap1, acc_s[0]);
          acc_s[1] = fma(r_prev.x, p0, acc_s[1]);
          acc_s[1] = fma(r_prev.y, p1, acc_s[1]);
        } else if (MODE == MODE_UPD) {
          // x += alpha p; r -= alpha Ap (matrix_free_system.cpp:422-429)
          double2 xn, rn;
          xn.x = v0 ? __dadd_rn(x_prev.x, __dmul_rn(alpha, pc.x)) : 0.0;
          xn.y = v1 ? __dadd_rn(x_prev.y, __dmul_rn(alpha, pc.y)) : 0.0;
          rn.x = v0 ? __dsub_rn(r_prev.x, __dmul_rn(alpha, ap0)) : 0.0;
          rn.y = v1 ? __dsub_rn(r_prev.y, __dmul_rn(alpha, ap1)) : 0.0;
          if (st_ok) {
            st2(a.x + eoff, xn);
            st2(a.r_out + eoff, rn);
            st2(a.p_out + eoff, make_double2(p0, p1));
          }
          acc_s[0] = fma(rn.x, rn.x, acc_s[0]);
          acc_s[0] = fma(rn.y, rn.y, acc_s[0]);
          acc_m[0] = fmax(acc_m[0], fmax(fabs(rn.x), fabs(rn.y)));
          const double d0 = v0 ? __dsub_rn(xn.x, x_prev.x) : 0.0;  // msg_solver.cpp:124-129
          const double d1 = v1 ? __dsub_rn(xn.y, x_prev.y) : 0.0;
          acc_m[1] = fmax(acc_m[1], fmax(fabs(d0), fabs(d1)));
          if (REPORT) {
            acc_s[1] = fma(d0, d0, acc_s[1]);
            acc_s[1] = fma(d1, d1, acc_s[1]);
          }
          if (LOAD_U) {
            const double e0 = v0 ? __dsub_rn(xn.x, u_prev.x) : 0.0;  // msg_solver.cpp:132-139
            const double e1 = v1 ? __dsub_rn(xn.y, u_prev.y) : 0.0;
            acc_m[NM - 1] = fmax(acc_m[NM - 1], fmax(fabs(e0), fabs(e1)));
            if (REPORT) {
              acc_s[2] = fma(e0, e0, acc_s[2]);
              acc_s[2] = fma(e1, e1, acc_s[2]);
            }
          }
        } else {
          if (REPORT) {
            const double d0 = v0 ? __dsub_rn(r_prev.x, ap0) : 0.0;  // b - A x, matrix_free_system.cpp:459-463
            const double d1 = v1 ? __dsub_rn(r_prev.y, ap1) : 0.0;
            acc_s[0] = fma(d0, d0, acc_s[0]);
            acc_s[0] = fma(d1, d1, acc_s[0]);
            if (LOAD_U) {
              const double e0 = v0 ? __dsub_rn(pc.x, u_prev.x) : 0.0;
              const double e1 = v1 ? __dsub_rn(pc.y, u_prev.y) : 0.0;
              acc_s[1] = fma(e0, e0, acc_s[1]);
              acc_s[1] = fma(e1, e1, acc_s[1]);
            }
          } else if (st_ok) {
            double2 o;
            if (FLAGS & F_SUB_B) {
              o.x = v0 ? __dsub_rn(ap0, r_prev.x) : 0.0;  // A x - b, dirichlet_solver.cpp:156-158
              o.y = v1 ? __dsub_rn(ap1, r_prev.y) : 0.0;
            } else {
              o.x = v0 ? ap0 : 0.0;
              o.y = v1 ? ap1 : 0.0;
            }
            st2(a.out + eoff, o);
          }
        }
        eoff += pitch;
      }
      pm = pc;
      pc = pn;
      Lp = L;
      Rp = R;
      r_prev = cur_r;
      x_prev = cur_x;
      u_prev = cur_u;
    }
  }

  if (NS + NM == 0) return;
  if (!grid_reduce<NS, NM>(acc_s, acc_m, a.partials, st, scratch)) return;
  // ---- one thread: turn the totals into the next scalars
  if (a.defer) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      st->loc_s[k] = k < NS ? acc_s[k < NS ? k : 0] : 0.0;
      st->loc_m[k] = k < NM ? acc_m[k < NM ? k : 0] : 0.0;
    }
    return;
  }
  if (MODE == MODE_DOT) {
    finalize_dot(st, acc_s[0], acc_s[1]);
  } else if (MODE == MODE_UPD) {
    finalize_update(st, a.cb_log, acc_s[0], acc_m[0], acc_m[1], LOAD_U ? acc_m[NM - 1] : DBL_MAX,
                    REPORT ? acc_s[1] : 0.0, (REPORT && LOAD_U) ? acc_s[2] : 0.0, REPORT);
  } else if (REPORT) {
    finalize_report(st, a.cb_log, acc_s[0], LOAD_U ? acc_s[1] : 0.0, LOAD_U);
  }
}

// --------------------------------------------------------------------------------------------- init
// Arms the device-side state from ||r0||: matrix_free_system.cpp:399-409, msg_solver.cpp:42-77.
__device__ __forceinline__ void finalize_init(DevState* st, CbRecord* log, double rr0, double rmax0, double umax,
                                              bool has_u) {
  const double r0 = sqrt(rr0);
  st->rr = rr0;
  st->rz = rr0;
  st->pAp = 0.0;
  st->alpha = 0.0;
  st->beta = 0.0;
  st->r0_norm = r0;
  st->r_norm = r0;
  st->r_max = rmax0;
  st->dx_max = DBL_MAX;
  st->err_max = has_u ? umax : DBL_MAX;  // |x0 - u|_inf with x0 = 0
  st->dx_l2 = st->err_l2 = st->res_l2 = 0.0;
  st->it = 0;
  st->converged = 0;
  st->stop_reason = 0;
  st->report_pending = 0;
  int done = 0;
  if (st->rule == 0) {
    if (!(0 < st->max_it && r0 > st->eps_rel * r0)) {
      done = 1;
      st->converged = (r0 <= st->eps_rel * r0) ? 1 : 0;
      st->stop_reason = st->converged ? 2 : 0;
    }
  } else {
    if (st->callback_every > 0) append_record(st, log, 0.0, DBL_MAX, rmax0, st->err_max);
    if (st->max_it <= 0) done = 1;
  }
  st->done = done;
}
// r0 = b, x0 = 0, p_old = 0 over the rows this rank owns; reduces r0.r0, |r0|_inf, |0 - u|_inf and arms the
// device-side state (matrix_free_system.cpp:387-400, msg_solver.cpp:33-77).
struct InitArgs {
  const double* b;
  const double* u;  // may be null
  double* r;
  double* p;
  double* x;
  DevState* st;
  double* partials;
  CbRecord* cb_log;
  size_t begin, count;  // in doubles, multiples of 2
  int defer;
};

__global__ void __launch_bounds__(CTA_THREADS) cg_init_kernel(const InitArgs a) {
  __shared__ double scratch[3 * 32];
  double s[1] = {0.0}, mx[2] = {0.0, 0.0};
  const size_t n2 = a.count / 2;
  const double2 z = make_double2(0.0, 0.0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = a.begin + 2 * i;
    const double2 bv = __ldg(reinterpret_cast<const double2*>(a.b + o));
    st2(a.r + o, bv);
    st2(a.p + o, z);
    st2(a.x + o, z);
    s[0] = fma(bv.x, bv.x, s[0]);
    s[0] = fma(bv.y, bv.y, s[0]);
    mx[0] = fmax(mx[0], fmax(fabs(bv.x), fabs(bv.y)));
    if (a.u) {
      const double2 uv = __ldg(reinterpret_cast<const double2*>(a.u + o));
      mx[1] = fmax(mx[1], fmax(fabs(uv.x), fabs(uv.y)));
    }
  }
  DevState* st = a.st;
  if (!grid_reduce<1, 2>(s, mx, a.partials, st, scratch)) return;
  if (a.defer) {
    st->loc_s[0] = s[0]; st->loc_s[1] = st->loc_s[2] = st->loc_s[3] = 0.0;
    st->loc_m[0] = mx[0]; st->loc_m[1] = mx[1]; st->loc_m[2] = st->loc_m[3] = 0.0;
    return;
  }
  finalize_init(st, a.cb_log, s[0], mx[0], mx[1], a.u != nullptr);
}

// After an all-reduce of st->loc_s (sum) and st->loc_m (max) on a sharded plan: every rank forms the same scalars.
// which: 0 = init, 1 = dot phase, 2 = update phase, 3 = report
__global__ void finalize_kernel(DevState* st, CbRecord* log, int which, int flags) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const bool has_u = (flags & F_U) != 0, report = (flags & F_REPORT) != 0;
  if (which == 0) {
    finalize_init(st, log, st->loc_s[0], st->loc_m[0], st->loc_m[1], has_u);
  } else if (which == 1) {
    if (st->done) return;
    finalize_dot(st, st->loc_s[0], st->loc_s[1]);
  } else if (which == 2) {
    if (st->done) return;
    finalize_update(st, log, st->loc_s[0], st->loc_m[0], st->loc_m[1], has_u ? st->loc_m[2] : DBL_MAX,
                    report ? st->loc_s[1] : 0.0, (report && has_u) ? st->loc_s[2] : 0.0, report);
  } else {
    if (!st->report_pending) return;
    finalize_report(st, log, st->loc_s[0], has_u ? st->loc_s[1] : 0.0, has_u);
  }
}

// --------------------------------------------------------------------------------------------- layout
// compact (reference order, grid_system.cpp:84-111) <-> pitched device layout
__device__ __forceinline__ void decode_compact(const Geom& g, long long idx, int& x, int& y) {
  if (g.ysplit == 0) {  // RECT: row-major
    const long long row = idx / g.wU;
    y = (int)row + 1;
    x = (int)(idx - row * g.wU) + 1;
  } else if (idx < g.NB) {
    const long long row = idx / g.wB;
    y = (int)row + 1;
    x = (int)(idx - row * g.wB) + g.xsplit + 1;
  } else {
    const long long k = idx - g.NB;
    const long long row = k / g.wU;
    y = (int)row + g.ysplit + 1;
    x = (int)(k - row * g.wU) + 1;
  }
}
__device__ __forceinline__ size_t pitched_off(const Geom& g, int x, int y) {
  return (size_t)(y - g.ybase) * (size_t)g.pitch + (size_t)(x + XOFF);
}

__global__ void scatter_compact_kernel(const double* __restrict__ compact, double* __restrict__ pitched,
                                       const Geom g) {
  const long long cnt = g.hi - g.lo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt;
       i += (long long)gridDim.x * blockDim.x) {
    int x, y;
    decode_compact(g, g.lo + i, x, y);
    pitched[pitched_off(g, x, y)] = compact[i];
  }
}
__global__ void gather_compact_kernel(const double* __restrict__ pitched, double* __restrict__ compact,
                                      const Geom g) {
  const long long cnt = g.hi - g.lo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt;
       i += (long long)gridDim.x * blockDim.x) {
    int x, y;
    decode_compact(g, g.lo + i, x, y);
    compact[i] = pitched[pitched_off(g, x, y)];
  }
}
// error = x - u on the compact staging buffer (dirichlet_solver.cpp:172-174)
__global__ void gather_diff_kernel(const double* __restrict__ pa, const double* __restrict__ pb,
                                   double* __restrict__ compact, const Geom g) {
  const long long cnt = g.hi - g.lo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt;
       i += (long long)gridDim.x * blockDim.x) {
    int x, y;
    decode_compact(g, g.lo + i, x, y);
    const size_t o = pitched_off(g, x, y);
    compact[i] = __dsub_rn(pa[o], pb[o]);
  }
}

// --------------------------------------------------------------------------------------------- K0 setup
// f, u and the boundary predicates of the reference (grid_system.cpp:8-67)
// (explicitly rounded products/sums: the host reference has no FMA contraction)
__device__ __forceinline__ double f_rhs(double x, double y) {
  const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y);
  return __dmul_rn(__dmul_rn(4.0, __dadd_rn(xx, yy)), exp(__dsub_rn(xx, yy)));
}
__device__ __forceinline__ double u_exact(double x, double y) {
  return exp(__dsub_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
}
__device__ __forceinline__ bool left_bdry(const Geom& g, int x, int y) {
  if (g.ysplit == 0) return x == 0;
  return (x == 0 && y >= g.ysplit && y <= g.m) || (x == g.xsplit && y >= 0 && y <= g.ysplit);
}
__device__ __forceinline__ bool bottom_bdry(const Geom& g, int x, int y) {
  if (g.ysplit == 0) return y == 0;
  return (y == 0 && x >= g.xsplit && x <= g.n) || (y == g.ysplit && x >= 0 && x <= g.xsplit);
}

// what: 0 = rhs b (calculate_value, grid_system.cpp:45-67), 1 = true solution u, 2 = x coordinate, 3 = y coordinate.
// Written to the pitched vector (dst_pitched) and/or the compact staging buffer (dst_compact).
__global__ void setup_kernel(double* __restrict__ dst_pitched, double* __restrict__ dst_compact, const Geom g,
                             int what) {
  const long long cnt = g.hi - g.lo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt;
       i += (long long)gridDim.x * blockDim.x) {
    int x, y;
    decode_compact(g, g.lo + i, x, y);
    // x * hx and a + (.) rounded separately, as the reference's calculate_x (grid_system.cpp:69-72)
    const double xp = __dadd_rn(g.a, __dmul_rn((double)x, g.hx));
    const double yp = __dadd_rn(g.c, __dmul_rn((double)y, g.hy));
    double v;
    if (what == 0) {
      v = f_rhs(xp, yp);
      if (left_bdry(g, x - 1, y)) v = __dsub_rn(v, __dmul_rn(g.xk, u_exact(__dadd_rn(g.a, __dmul_rn((double)(x - 1), g.hx)), yp)));
      if (x + 1 == g.n) v = __dsub_rn(v, __dmul_rn(g.xk, u_exact(__dadd_rn(g.a, __dmul_rn((double)(x + 1), g.hx)), yp)));
      if (y + 1 == g.m) v = __dsub_rn(v, __dmul_rn(g.yk, u_exact(xp, __dadd_rn(g.c, __dmul_rn((double)(y + 1), g.hy)))));
      if (bottom_bdry(g, x, y - 1)) v = __dsub_rn(v, __dmul_rn(g.yk, u_exact(xp, __dadd_rn(g.c, __dmul_rn((double)(y - 1), g.hy)))));
    } else if (what == 1) {
      v = u_exact(xp, yp);
    } else if (what == 2) {
      v = xp;
    } else {
      v = yp;
    }
    if (dst_pitched) dst_pitched[pitched_off(g, x, y)] = v;
    if (dst_compact) dst_compact[i] = v;
  }
}

}  // namespace b200cg
