// Opt-in preconditioner behind the solver interface (the reference keeps solver/solver.hpp:17-66 as the
// abstract base for exactly this and ships none): a geometric-multigrid V-cycle on the node grid as the preconditioner
// of CG. Plain CG needs ~2.7 n iterations (44 000 at 16384^2); with the V-cycle the count is 7 independent of n.
//
//   levels   : n, m halved while both are even, >= 8 and (L-shape) the re-entrant lines x = n/2, y = m/2 stay on coarse
//              grid lines; coarse operators by rediscretisation (the reference's coefficients with the doubled step,
//              grid_system.cpp:314-318)
//   smoother : damped Jacobi, omega = 0.8 (the diagonal is constant): x <- x + (omega / A_diag) (b - A x)
//   V-cycle  : 2 pre-sweeps (the first from x = 0), residual, full-weighting restriction, recursion, bilinear
//              prolongation, 2 post-sweeps; coarsest level: 16 sweeps. Symmetric, hence a valid CG preconditioner.
// Every vector is a pitched copy of its level's node grid with exact zeros outside the unknowns (the layout of the CG
// vectors, common.cuh), so the stencil needs no boundary predicates - only the outputs are masked. Elementwise
// arithmetic is separately rounded in a fixed order; the test suite restates it in numpy operation by operation.
// These kernels are plain one-thread-per-node-pair sweeps: a V-cycle costs ~190 B per fine unknown and the whole
// 16384^2 solve a few hundred milliseconds against a minute of plain CG, so they are not the tuned hot path.
#pragma once
#include "kernels_common.cuh"

namespace b200cg {

constexpr double MG_OMEGA = 0.8;
constexpr int MG_NU_PRE = 2, MG_NU_POST = 2, MG_NU_COARSEST = 16;
constexpr int MG_THREADS = 128;  // one node pair per thread: 256 columns per block

struct MgGeom {
  int n, m, xsplit, ysplit, pitch;
  int nbx, nby;        // block decomposition of the 1-D grid: block b handles column block b % nbx, rows b / nbx + 1 + k nby
  double A, xk, yk, w;  // stencil coefficients of this level, w = omega / A
};

__device__ __forceinline__ bool mg_unknown(const MgGeom& g, int x, int y) {
  if (x < 1 || x > g.n - 1 || y < 1 || y > g.m - 1) return false;
  if (g.ysplit && y <= g.ysplit) return x > g.xsplit;
  return true;
}
__device__ __forceinline__ size_t mg_off(const MgGeom& g, int x, int y) { return (size_t)y * g.pitch + (size_t)(x + XOFF); }

// A v at the node pair (x0, x0 + 1) of row y; accumulation order diag, left, right, top, bottom
// (matrix_free_system.cpp:216-266), each term a rounded multiply then a rounded add.
__device__ __forceinline__ double2 mg_stencil(const MgGeom& g, const double* __restrict__ v, size_t o) {
  const double2 c = *reinterpret_cast<const double2*>(v + o);
  const double l = v[o - 1], r = v[o + 2];
  const double2 up = *reinterpret_cast<const double2*>(v + o + g.pitch);
  const double2 dn = *reinterpret_cast<const double2*>(v + o - g.pitch);
  double2 a;
  a.x = __dmul_rn(g.A, c.x);
  a.x = __dadd_rn(a.x, __dmul_rn(g.xk, l));
  a.x = __dadd_rn(a.x, __dmul_rn(g.xk, c.y));
  a.x = __dadd_rn(a.x, __dmul_rn(g.yk, up.x));
  a.x = __dadd_rn(a.x, __dmul_rn(g.yk, dn.x));
  a.y = __dmul_rn(g.A, c.y);
  a.y = __dadd_rn(a.y, __dmul_rn(g.xk, c.x));
  a.y = __dadd_rn(a.y, __dmul_rn(g.xk, r));
  a.y = __dadd_rn(a.y, __dmul_rn(g.yk, up.y));
  a.y = __dadd_rn(a.y, __dmul_rn(g.yk, dn.y));
  return a;
}

enum { MG_JACOBI_FIRST = 0, MG_JACOBI = 1, MG_RESIDUAL = 2, MG_APPLY_DOT = 3 };

// OP 0: out = w b            (first Jacobi sweep from x = 0)
// OP 1: out = x + w (b - A x)
// OP 2: out = b - A x
// OP 3: out = A x, reduces x.out and forms alpha = r.z / p.Ap (the PCG operator application; x = p)
template <int OP>
__global__ void __launch_bounds__(MG_THREADS) mg_stencil_kernel(const MgGeom g, const double* __restrict__ x,
                                                                const double* __restrict__ b, double* __restrict__ out,
                                                                DevState* st, double* partials) {
  __shared__ double scratch[32];
  const int bx = blockIdx.x % g.nbx, by = blockIdx.x / g.nbx;
  const int col = 2 * (bx * MG_THREADS + threadIdx.x);  // storage column of the pair
  const int x0 = col - XOFF;
  double acc[1] = {0.0}, none[1] = {0.0};
  if (col + 1 < g.pitch) {
    for (int y = 1 + by; y <= g.m - 1; y += g.nby) {
      const bool v0 = mg_unknown(g, x0, y), v1 = mg_unknown(g, x0 + 1, y);
      if (!(v0 || v1)) continue;
      const size_t o = mg_off(g, x0, y);
      double2 res;
      if (OP == MG_JACOBI_FIRST) {
        const double2 bv = *reinterpret_cast<const double2*>(b + o);
        res.x = __dmul_rn(g.w, bv.x);
        res.y = __dmul_rn(g.w, bv.y);
      } else {
        const double2 ax = mg_stencil(g, x, o);
        if (OP == MG_APPLY_DOT) {
          res = ax;
        } else {
          const double2 bv = *reinterpret_cast<const double2*>(b + o);
          res.x = __dsub_rn(bv.x, ax.x);
          res.y = __dsub_rn(bv.y, ax.y);
          if (OP == MG_JACOBI) {
            const double2 xv = *reinterpret_cast<const double2*>(x + o);
            res.x = __dadd_rn(xv.x, __dmul_rn(g.w, res.x));
            res.y = __dadd_rn(xv.y, __dmul_rn(g.w, res.y));
          }
        }
      }
      res.x = v0 ? res.x : 0.0;
      res.y = v1 ? res.y : 0.0;
      st2(out + o, res);
      if (OP == MG_APPLY_DOT) {
        const double2 xv = *reinterpret_cast<const double2*>(x + o);
        acc[0] = fma(xv.x, res.x, acc[0]);
        acc[0] = fma(xv.y, res.y, acc[0]);
      }
    }
  }
  if (OP != MG_APPLY_DOT) return;
  if (!grid_reduce<1, 0>(acc, none, partials, st, scratch)) return;
  st->pAp = acc[0];
  st->alpha = st->rz / acc[0];
}

// Full weighting of the fine residual onto the coarse unknowns: (4 c + 2 ((W + E) + N + S) + ((NW + NE) + SW + SE)) / 16.
__global__ void __launch_bounds__(MG_THREADS) mg_restrict_kernel(const MgGeom gf, const MgGeom gc,
                                                                 const double* __restrict__ r, double* __restrict__ bc) {
  const int bx = blockIdx.x % gc.nbx, by = blockIdx.x / gc.nbx;
  const int X0 = 2 * (bx * MG_THREADS + threadIdx.x) - XOFF;
  if (X0 + XOFF + 1 >= gc.pitch) return;
  for (int Y = 1 + by; Y <= gc.m - 1; Y += gc.nby) {
    double2 res = make_double2(0.0, 0.0);
    bool any = false;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int X = X0 + k;
      if (!mg_unknown(gc, X, Y)) continue;
      any = true;
      const size_t o = mg_off(gf, 2 * X, 2 * Y);
      const size_t p = (size_t)gf.pitch;
      const double s1 = __dadd_rn(__dadd_rn(__dadd_rn(r[o - 1], r[o + 1]), r[o + p]), r[o - p]);
      const double s2 = __dadd_rn(__dadd_rn(__dadd_rn(r[o + p - 1], r[o + p + 1]), r[o - p - 1]), r[o - p + 1]);
      const double v = __dmul_rn(0.0625, __dadd_rn(__dadd_rn(__dmul_rn(4.0, r[o]), __dmul_rn(2.0, s1)), s2));
      if (k == 0) res.x = v;
      else res.y = v;
    }
    if (any) st2(bc + mg_off(gc, X0, Y), res);
  }
}

// x += bilinear interpolation of the coarse correction (zero on the coarse boundary), on the fine unknowns.
__global__ void __launch_bounds__(MG_THREADS) mg_prolong_add_kernel(const MgGeom gf, const MgGeom gc,
                                                                    const double* __restrict__ e, double* __restrict__ x) {
  const int bx = blockIdx.x % gf.nbx, by = blockIdx.x / gf.nbx;
  const int x0 = 2 * (bx * MG_THREADS + threadIdx.x) - XOFF;  // even node of the pair
  if (x0 + XOFF + 1 >= gf.pitch) return;
  for (int y = 1 + by; y <= gf.m - 1; y += gf.nby) {
    const bool v0 = mg_unknown(gf, x0, y), v1 = mg_unknown(gf, x0 + 1, y);
    if (!(v0 || v1)) continue;
    const int X = x0 >> 1, Y = y >> 1;
    const size_t oc = mg_off(gc, X, Y), pc = (size_t)gc.pitch;
    const double e00 = e[oc], e10 = e[oc + 1];  // e(X, Y), e(X + 1, Y)
    double p0, p1;
    if ((y & 1) == 0) {
      p0 = e00;
      p1 = __dmul_rn(0.5, __dadd_rn(e00, e10));
    } else {
      const double e01 = e[oc + pc], e11 = e[oc + pc + 1];  // e(X, Y + 1), e(X + 1, Y + 1)
      p0 = __dmul_rn(0.5, __dadd_rn(e00, e01));
      p1 = __dmul_rn(0.25, __dadd_rn(__dadd_rn(__dadd_rn(e00, e10), e01), e11));
    }
    const size_t o = mg_off(gf, x0, y);
    double2 xv = *reinterpret_cast<const double2*>(x + o);
    xv.x = v0 ? __dadd_rn(xv.x, p0) : 0.0;
    xv.y = v1 ? __dadd_rn(xv.y, p1) : 0.0;
    st2(x + o, xv);
  }
}

// ---- the PCG vector operations on whole pitched vectors (zeros outside the unknowns stay zeros: no masks)
// x += alpha p, r -= alpha Ap, reduces r.r; the last CTA advances the iteration and evaluates the stop rule of
// MatrixFreeSolver::solve (matrix_free_system.cpp:409).
__global__ void __launch_bounds__(CTA_THREADS) pcg_update_kernel(double* __restrict__ x, double* __restrict__ r,
                                                                 const double* __restrict__ p, const double* __restrict__ Ap,
                                                                 size_t begin, size_t count, DevState* st, double* partials,
                                                                 const int* stop_flag) {
  __shared__ double scratch[32];
  const double alpha = st->alpha;
  double acc[1] = {0.0}, none[1] = {0.0};
  const size_t n2 = count / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = begin + 2 * i;
    const double2 pv = *reinterpret_cast<const double2*>(p + o);
    const double2 av = *reinterpret_cast<const double2*>(Ap + o);
    double2 xv = *reinterpret_cast<const double2*>(x + o);
    double2 rv = *reinterpret_cast<const double2*>(r + o);
    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, pv.x));
    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, pv.y));
    rv.x = __dsub_rn(rv.x, __dmul_rn(alpha, av.x));
    rv.y = __dsub_rn(rv.y, __dmul_rn(alpha, av.y));
    st2(x + o, xv);
    st2(r + o, rv);
    acc[0] = fma(rv.x, rv.x, acc[0]);
    acc[0] = fma(rv.y, rv.y, acc[0]);
  }
  if (!grid_reduce<1, 0>(acc, none, partials, st, scratch)) return;
  const bool stop_req = poll_stop(st, stop_flag);
  const int it = st->it + 1;
  st->it = it;
  st->rr = acc[0];
  const double r_norm = sqrt(acc[0]);
  st->r_norm = r_norm;
  const bool go = (it < st->max_it) && (r_norm > st->eps_rel * st->r0_norm);
  if (!go) {
    st->done = 1;
    st->converged = (r_norm <= st->eps_rel * st->r0_norm) ? 1 : 0;
    st->stop_reason = st->converged ? 2 : 0;
  }
  apply_stop(st, stop_req);
}

// r.z -> beta = r'.z' / r.z (first = 1: the initial r0.z0, beta stays 0)
__global__ void __launch_bounds__(CTA_THREADS) pcg_dot_rz_kernel(const double* __restrict__ r, const double* __restrict__ z,
                                                                 size_t begin, size_t count, DevState* st, double* partials,
                                                                 int first) {
  __shared__ double scratch[32];
  double acc[1] = {0.0}, none[1] = {0.0};
  const size_t n2 = count / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = begin + 2 * i;
    const double2 rv = *reinterpret_cast<const double2*>(r + o);
    const double2 zv = *reinterpret_cast<const double2*>(z + o);
    acc[0] = fma(rv.x, zv.x, acc[0]);
    acc[0] = fma(rv.y, zv.y, acc[0]);
  }
  if (!grid_reduce<1, 0>(acc, none, partials, st, scratch)) return;
  st->beta = first ? 0.0 : acc[0] / st->rz;
  st->rz = acc[0];
}

// p = z + beta p
__global__ void __launch_bounds__(CTA_THREADS) pcg_direction_kernel(double* __restrict__ p, const double* __restrict__ z,
                                                                    size_t begin, size_t count, const DevState* st) {
  const double beta = st->beta;
  const size_t n2 = count / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = begin + 2 * i;
    const double2 zv = *reinterpret_cast<const double2*>(z + o);
    double2 pv = *reinterpret_cast<const double2*>(p + o);
    pv.x = __dadd_rn(zv.x, __dmul_rn(beta, pv.x));
    pv.y = __dadd_rn(zv.y, __dmul_rn(beta, pv.y));
    st2(p + o, pv);
  }
}

}  // namespace b200cg
