// Small-grid path (sm_100a): the whole CG solve in ONE launch of ONE thread-block cluster.
//
// Grids up to ~300^2 are pure latency: two kernels + two grid-wide reductions per iteration cost ~17 us in the
// graph path even though the data fits on chip. Here a cluster of 8 (or 16) CTAs keeps r, p and x of its row band
// in shared memory for the whole solve, reads the neighbour bands' boundary rows of p through distributed
// shared memory (DSMEM), reduces the dot products through DSMEM slots, and separates the phases with
// barrier.cluster (three per iteration) instead of kernel boundaries. alpha, beta and the stop verdict are formed
// redundantly and identically by every CTA from the same slots, so nothing leaves the cluster until the end.
// Arithmetic is the same as in stream_kernel.cuh: separately rounded multiply/add in the reference's order
// (matrix_free_system.cpp:216-266, :422-438); reductions use a fixed tree and a fixed CTA order (deterministic).
#pragma once
#include <cooperative_groups.h>
#include "kernels_common.cuh"

namespace b200cg {

namespace cg = cooperative_groups;

constexpr int CL_THREADS = 512;
constexpr int CL_MAX_CTAS = 16;
constexpr int CL_SLOTS = 8;          // values per CTA per reduction
constexpr int CL_POLL_EVERY = 128;   // iterations between looks at the host's interrupt flag

struct ClusterArgs {
  const double* b;   // pitched rhs
  const double* u;   // pitched true solution or null
  double* x;         // pitched solution (written once, at the end)
  DevState* st;      // parameters in, final state out
  CbRecord* cb_log;
  const volatile int* stop_flag;  // mapped host memory, may be null
  Geom g;
  int rows_per_cta;
};

static inline size_t cluster_smem_bytes(int rows_per_cta, int pitch) {
  return (size_t)3 * rows_per_cta * pitch * sizeof(double) + 128;
}

// Sum / max of the per-CTA slot rows, in CTA order: every thread of every CTA gets the same value.
__device__ __forceinline__ double slots_sum(const double (*slots)[CL_SLOTS], int nctas, int k) {
  double v = 0.0;
  for (int c = 0; c < nctas; ++c) v += slots[c][k];
  return v;
}
__device__ __forceinline__ double slots_max(const double (*slots)[CL_SLOTS], int nctas, int k) {
  double v = 0.0;
  for (int c = 0; c < nctas; ++c) v = fmax(v, slots[c][k]);
  return v;
}

static __global__ void __launch_bounds__(CL_THREADS, 1) cg_cluster_kernel(const ClusterArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int nctas = (int)cluster.num_blocks();
  const Geom& g = a.g;
  const int pitch = g.pitch;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  extern __shared__ __align__(16) unsigned char cl_smem[];
  const int band_nodes = a.rows_per_cta * pitch;
  double* sp = reinterpret_cast<double*>(cl_smem);  // direction p
  double* sr = sp + band_nodes;                     // residual r
  double* sx = sr + band_nodes;                     // solution x
  __shared__ double slotsA[CL_MAX_CTAS][CL_SLOTS];  // dot-phase partials of every CTA (written remotely)
  __shared__ double slotsB[CL_MAX_CTAS][CL_SLOTS];  // update-phase partials
  __shared__ double scratch[CL_SLOTS * 32];
  __shared__ DevState sst;                          // this CTA's copy of the scalar state (all copies agree)

  // band of grid rows owned by this CTA
  const int y0 = 1 + rank * a.rows_per_cta;
  const int y1 = min(y0 + a.rows_per_cta, g.m);
  const int rows = max(y1 - y0, 0);
  const int my_nodes = rows * pitch;
  const double* p_below = (rank > 0) ? cluster.map_shared_rank(sp, rank - 1) : nullptr;
  const double* p_above = (rank + 1 < nctas) ? cluster.map_shared_rank(sp, rank + 1) : nullptr;
  const int below_rows = a.rows_per_cta;  // every lower band is full

  if (tid == 0) sst = *a.st;
  // ---- init: r0 = b, x0 = 0, p = 0 (matrix_free_system.cpp:387-400, msg_solver.cpp:33-39)
  double s0 = 0.0, m0 = 0.0, m1 = 0.0;
  for (int i = tid; i < band_nodes; i += CL_THREADS) {
    double bv = 0.0, uv = 0.0;
    if (i < my_nodes) {
      const size_t go = (size_t)(y0 - g.ybase) * pitch + i;
      bv = a.b[go];
      if (a.u) uv = a.u[go];
    }
    sr[i] = bv;
    sp[i] = 0.0;
    sx[i] = 0.0;
    s0 = fma(bv, bv, s0);
    m0 = fmax(m0, fabs(bv));
    m1 = fmax(m1, fabs(uv));
  }
  {
    double s[1] = {s0}, mx[2] = {m0, m1};
    block_reduce<1, 2>(s, mx, scratch);
    if (tid == 0) {
      for (int c = 0; c < nctas; ++c) {
        double(*dst)[CL_SLOTS] = cluster.map_shared_rank(slotsB, c);
        dst[rank][0] = s[0];
        dst[rank][1] = mx[0];
        dst[rank][2] = mx[1];
        dst[rank][3] = 0.0;
      }
    }
  }
  cluster.sync();
  if (tid == 0) {
    finalize_init(&sst, rank == 0 ? a.cb_log : nullptr, slots_sum(slotsB, nctas, 0), slots_max(slotsB, nctas, 1),
                  slots_max(slotsB, nctas, 2), a.u != nullptr);
  }
  __syncthreads();

  const double cA = g.A, cxk = g.xk, cyk = g.yk;
  const bool has_u = a.u != nullptr;
  int interrupted = 0;

  // A p at band node i (row, col), with p rows outside the band read from the neighbour CTAs through DSMEM.
  auto apply_at = [&](int i, int row, bool valid, double pc) -> double {
    const double left = sp[i - 1], right = sp[i + 1];
    double top, bottom;
    if (row + 1 < rows) top = sp[i + pitch];
    else top = p_above ? p_above[i - row * pitch] : 0.0;  // row 0 of the band above
    if (row > 0) bottom = sp[i - pitch];
    else bottom = p_below ? p_below[(below_rows - 1) * pitch + i] : 0.0;  // last row of the band below
    double ap = __dmul_rn(cA, pc);
    ap = __dadd_rn(ap, __dmul_rn(cxk, left));
    ap = __dadd_rn(ap, __dmul_rn(cxk, right));
    ap = __dadd_rn(ap, __dmul_rn(cyk, top));
    ap = __dadd_rn(ap, __dmul_rn(cyk, bottom));
    return valid ? ap : 0.0;
  };

  while (!sst.done) {
    const double beta = sst.beta;
    // ---- direction: p = r + beta p (own band; zeros stay zeros)
    // warp w sweeps rows w, w + 16, ...; lane l the unknown columns x = xlo + l, xlo + l + 32, ... of a row
    for (int row = warp; row < rows; row += CL_THREADS / 32) {
      const int xlo = (g.ysplit && y0 + row <= g.ysplit) ? g.xsplit + 1 : 1;
      for (int x = xlo + lane; x <= g.n - 1; x += 32) {
        const int i = row * pitch + x + XOFF;
        sp[i] = __dadd_rn(sr[i], __dmul_rn(beta, sp[i]));
      }
    }
    cluster.sync();  // every band's p is complete

    // ---- dot phase: p.Ap and r.p
    double d0 = 0.0, d1 = 0.0;
    for (int row = warp; row < rows; row += CL_THREADS / 32) {
      const int xlo = (g.ysplit && y0 + row <= g.ysplit) ? g.xsplit + 1 : 1;
      for (int x = xlo + lane; x <= g.n - 1; x += 32) {
        const int i = row * pitch + x + XOFF;
        const double pc = sp[i];
        const double ap = apply_at(i, row, true, pc);
        d0 = fma(pc, ap, d0);
        d1 = fma(sr[i], pc, d1);
      }
    }
    {
      double s[2] = {d0, d1}, mx[1] = {0.0};
      block_reduce<2, 0>(s, mx, scratch);  // every lane of warp 0 holds the totals
      if (tid < nctas) {                   // lane c publishes them in CTA c's slot row (one DSMEM store each)
        double(*dst)[CL_SLOTS] = cluster.map_shared_rank(slotsA, tid);
        dst[rank][0] = s[0];
        dst[rank][1] = s[1];
      }
    }
    cluster.sync();
    if (tid == 0) finalize_dot(&sst, slots_sum(slotsA, nctas, 0), slots_sum(slotsA, nctas, 1));
    __syncthreads();
    const double alpha = sst.alpha;

    // ---- update phase: x += alpha p, r -= alpha Ap (Ap recomputed from the same p: bit-identical)
    double rr = 0.0, rmax = 0.0, dxmax = 0.0, emax = 0.0;
    for (int row = warp; row < rows; row += CL_THREADS / 32) {
      const int xlo = (g.ysplit && y0 + row <= g.ysplit) ? g.xsplit + 1 : 1;
      for (int x = xlo + lane; x <= g.n - 1; x += 32) {
        const int i = row * pitch + x + XOFF;
        const double pc = sp[i];
        const double ap = apply_at(i, row, true, pc);
        const double xo = sx[i];
        const double xn = __dadd_rn(xo, __dmul_rn(alpha, pc));
        const double rn = __dsub_rn(sr[i], __dmul_rn(alpha, ap));
        sx[i] = xn;
        // r is only read at node i by this thread: safe to update in place; p is not touched here
        sr[i] = rn;
        rr = fma(rn, rn, rr);
        rmax = fmax(rmax, fabs(rn));
        dxmax = fmax(dxmax, fabs(__dsub_rn(xn, xo)));
        if (has_u) emax = fmax(emax, fabs(__dsub_rn(xn, a.u[(size_t)(y0 + row - g.ybase) * pitch + x + XOFF])));
      }
    }
    {
      double s[1] = {rr}, mx[4] = {rmax, dxmax, emax, 0.0};
      if (rank == 0 && tid == 0 && a.stop_flag && ((sst.it + 1) % CL_POLL_EVERY) == 0) mx[3] = (*a.stop_flag != 0) ? 1.0 : 0.0;
      block_reduce<1, 4>(s, mx, scratch);
      if (tid < nctas) {
        double(*dst)[CL_SLOTS] = cluster.map_shared_rank(slotsB, tid);
        dst[rank][0] = s[0];
        dst[rank][1] = mx[0];
        dst[rank][2] = mx[1];
        dst[rank][3] = mx[2];
        dst[rank][4] = mx[3];
      }
    }
    cluster.sync();
    if (tid == 0) {
      finalize_update(&sst, rank == 0 ? a.cb_log : nullptr, slots_sum(slotsB, nctas, 0), slots_max(slotsB, nctas, 1),
                      slots_max(slotsB, nctas, 2), has_u ? slots_max(slotsB, nctas, 3) : DBL_MAX, 0.0, 0.0, false);
      if (!sst.done && slots_max(slotsB, nctas, 4) > 0.0) {
        sst.done = 1;
        interrupted = 1;
      }
    }
    __syncthreads();
  }

  // ---- results: x of the band, final state (all CTAs hold the same one; CTA 0 publishes it)
  for (int i = tid; i < my_nodes; i += CL_THREADS) a.x[(size_t)(y0 - g.ybase) * pitch + i] = sx[i];
  if (rank == 0 && tid == 0) {
    if (interrupted) {
      sst.converged = 0;
      sst.stop_reason = 4;
    }
    sst.ticket = 0u;
    *a.st = sst;
  }
  cluster.sync();  // no CTA may exit while a neighbour can still read its shared memory
}

}  // namespace b200cg
