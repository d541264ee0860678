// Device kernels of libb200cg besides the hot sweep (stream_kernel.cuh): initialisation, layout conversion,
// K0 setup (rhs, true solution, coordinates).
#pragma once
#include "kernels_common.cuh"
#include "stream_kernel.cuh"

namespace b200cg {

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
  st->x_pending = 0;
  st->alpha_prev = 0.0;
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

static __global__ void __launch_bounds__(CTA_THREADS) cg_init_kernel(const InitArgs a) {
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
static __global__ void finalize_kernel(DevState* st, CbRecord* log, int which, int flags) {
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
    note_x_deferral(st, flags);
  } else {
    if (!st->report_pending) return;
    finalize_report(st, log, st->loc_s[0], has_u ? st->loc_s[1] : 0.0, has_u);
  }
}

// x += alpha_prev * p over the owned rows: settles the update an even (NOX) last iteration left pending.
static __global__ void __launch_bounds__(CTA_THREADS) x_flush_kernel(double* __restrict__ x, const double* __restrict__ p,
                                                             DevState* st, size_t begin, size_t count) {
  if (!st->x_pending) return;
  const double alpha = st->alpha_prev;
  const size_t n2 = count / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = begin + 2 * i;
    const double2 pv = __ldg(reinterpret_cast<const double2*>(p + o));
    double2 xv = *reinterpret_cast<const double2*>(x + o);
    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, pv.x));
    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, pv.y));
    st2(x + o, xv);
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

static __global__ void scatter_compact_kernel(const double* __restrict__ compact, double* __restrict__ pitched,
                                       const Geom g) {
  const long long cnt = g.hi - g.lo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt;
       i += (long long)gridDim.x * blockDim.x) {
    int x, y;
    decode_compact(g, g.lo + i, x, y);
    pitched[pitched_off(g, x, y)] = compact[i];
  }
}
static __global__ void gather_compact_kernel(const double* __restrict__ pitched, double* __restrict__ compact,
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
static __global__ void gather_diff_kernel(const double* __restrict__ pa, const double* __restrict__ pb,
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
static __global__ void setup_kernel(double* __restrict__ dst_pitched, double* __restrict__ dst_compact, const Geom g,
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
