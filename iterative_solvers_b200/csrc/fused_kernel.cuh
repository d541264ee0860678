// Single-sweep CG iteration (sm_100a): one kernel per iteration, 40 B per unknown instead of 56.
//
// The two-sweep scheme (stream_kernel.cuh) needs p.Ap before it may update r, hence two passes over r and p.
// Here alpha comes from the recurrence of the single-reduction CG (Chronopoulos & Gear):
//
//     gamma_k = r_k.r_k,   delta_k = r_k.A r_k,   beta_k = gamma_k / gamma_{k-1},
//     alpha_k = gamma_k / (delta_k - beta_k gamma_k / alpha_{k-1})          (= gamma_k / p_k.A p_k in exact arithmetic)
//
// so that one pass can do   p = r + beta p_old,  x += alpha p,  r' = r - alpha A p,  gamma' = r'.r',  delta' = r'.A r'.
// A r' needs r' at the four neighbours, i.e. A p one node further out, i.e. p two nodes out: the sweep carries a halo
// of two rows / two columns and recomputes p and r' there. Element-wise arithmetic is unchanged (separately rounded
// multiply/add in the reference's order, matrix_free_system.cpp:216-266, :422-438); only the way alpha is formed
// differs, and tests/studies/single_reduction_cg.py shows the iterates stay within 4e-14 of the reference's on every
// golden grid (same iteration counts). The default iteration of the REL_L2 rule without a report callback and - as the
// F_MAXN flavour, which touches x every iteration and also reduces the three max-norms - of MSGSolver's rules, on single
// and on sharded (peer-memory) plans; b200cg_params.single_sweep = 2 / B200CG_SINGLE_SWEEP=0 select the two-sweep one.
//
// Structure: the producer warp / mbarrier stage ring of stream_kernel.cuh. Consumers differ in the column mapping:
// every warp owns a window of 64 staged columns and writes the inner 60, so all horizontal neighbours (two levels)
// come from warp shuffles and no lane needs another warp's data. A CTA is 7 consumer warps + the producer = 8 warps:
// at 2 CTAs/SM that leaves 128 registers per thread (a 9-warp CTA is capped at 96 and the x-touching flavour spilled,
// profiles/r2_single_sweep.md), 7 x 60 = 420 written columns per strip of 424 staged ones. Rows run through a two-deep
// register pipeline: when row y arrives, A p and r' of row y-1 and A r' of row y-2 become computable.
#pragma once
#include "stream_kernel.cuh"

namespace b200cg {

constexpr int FUSED_CW = 7;                                  // consumer warps per CTA (default geometry)
constexpr int FUSED_WARP_STEP = 60;                          // columns written per consumer warp (64 processed)
constexpr int FUSED_COL_SHIFT = 2;                           // a strip's first staged storage column is strip * width + this
__host__ __device__ constexpr int fused_strip_out(int cw) { return FUSED_WARP_STEP * cw; }   // columns written per strip: 420 (840)
__host__ __device__ constexpr int fused_row(int cw) { return fused_strip_out(cw) + 4; }      // staged columns = doubles between staged rows
constexpr int FUSED_STRIP_OUT = fused_strip_out(FUSED_CW);
// CW = 14 (B200CG_FUSED_CW=14, experiment): one 15-warp CTA per SM on 840-column strips - half as many strip edges
// (partial 128-byte lines, halo columns) per byte; profiles/r2_single_sweep.md

template <int FLAGS>
struct FusedCfg {
  static constexpr bool X2 = (FLAGS & F_X2) != 0;   // odd iteration: x += alpha_prev * p_old + alpha * p
  // F_MAXN (MSGSolver's max-norm rules, msg_solver.cpp:105-162): x += alpha * p EVERY iteration and the sweep also
  // reduces |r'|_inf, |x' - x|_inf and, with F_U, |x' - u|_inf - the rules look at them after every iteration, so there
  // is no x-deferral: 48 B per unknown-iteration (56 with u) against 64 (72) of the dot sweep + update sweep
  static constexpr bool MAXN = (FLAGS & F_MAXN) != 0;
  static constexpr bool XS = X2 || MAXN;            // x is streamed (emit rows only)
  static constexpr bool LOAD_U = MAXN && (FLAGS & F_U) != 0;  // so is the true solution
  static constexpr int NSTREAM = 2 + (XS ? 1 : 0) + (LOAD_U ? 1 : 0);  // p, r, [x], [u]
  static constexpr int NS = 2;                      // gamma' = r'.r', delta' = r'.A r'
  static constexpr int NM = MAXN ? (LOAD_U ? 3 : 2) : 0;  // |r'|_inf, |dx|_inf, [|x' - u|_inf]'
  static constexpr int ROWS_BELOW = 2;              // rows streamed below a tile's first emit row
  // F_SHARD (sharded plans, peer memory): the slab's two first / last rows of r' and p also go to the neighbours - into
  // their halo row and into one of the two extra rows every pitched vector carries behind its stored rows (row ylo-2 at
  // index yrows, row yhi+1 at index yrows+1) - and the iteration's two sums cross the ranks through the PeerSync slots,
  // alternating the slot by iteration parity: ONE publish-and-wait per iteration.
  static constexpr bool SHARD = (FLAGS & F_SHARD) != 0;
  static_assert(!(X2 && MAXN), "x-deferral and the max-norm rules exclude each other");
};

template <int FLAGS, int HS, int NST, int CW>
constexpr size_t fused_smem_bytes() {
  return (size_t)NST * HS * FusedCfg<FLAGS>::NSTREAM * fused_row(CW) * 8 + (size_t)NST * (16 + sizeof(StageMeta)) + 128;
}

// The scalars of the next iteration from gamma' = r'.r' and delta' = r'.A r' (one thread, after the grid reduction).
// Stop test as MatrixFreeSolver's (matrix_free_system.cpp:409, :432-441).
__device__ __forceinline__ void finalize_fused(DevState* st, double gamma_new, double delta_new, int flags) {
  const int it = st->it + 1;
  st->it = it;
  const double r_norm = sqrt(gamma_new);
  st->r_norm = r_norm;
  st->r_max = 0.0;
  st->dx_max = 0.0;
  note_x_deferral(st, flags);  // remembers this iteration's alpha before it is replaced
  const bool go = (it < st->max_it) && (r_norm > st->eps_rel * st->r0_norm);
  if (!go) {
    st->rr = gamma_new;
    st->done = 1;
    st->converged = (r_norm <= st->eps_rel * st->r0_norm) ? 1 : 0;
    st->stop_reason = st->converged ? 2 : 0;
    return;
  }
  const double beta = gamma_new / st->rr;
  const double alpha = gamma_new / (delta_new - beta * gamma_new / st->alpha);
  st->rr = gamma_new;
  st->beta = beta;
  st->alpha = alpha;
  st->pAp = gamma_new / alpha;
}

// The same for MSGSolver's rules (msg_solver.cpp:115-183): the three max-norm tests in the reference's order, then
// beta = (|r'|_2)^2 / r.z with r.z = gamma of the previous iteration (r.z = r.r in exact arithmetic: z - r is a multiple
// of the previous direction, to which r is orthogonal) and alpha from the single-reduction recurrence.
__device__ __forceinline__ void finalize_fused_maxnorm(DevState* st, CbRecord* log, double gamma_new, double delta_new,
                                                       double r_max, double dx_max, double err_max) {
  const int it = st->it + 1;
  st->it = it;
  const double r_norm = sqrt(gamma_new);
  st->r_norm = r_norm;
  st->r_max = r_max;
  st->dx_max = dx_max;
  if (st->has_u) st->err_max = err_max;
  st->x_pending = 0;
  int done = 0;
  if (st->eps_p > 0 && dx_max < st->eps_p) { done = 1; st->converged = 1; st->stop_reason = 1; }
  else if (st->eps_r > 0 && r_max < st->eps_r) { done = 1; st->converged = 1; st->stop_reason = 2; }
  else if (st->eps_e > 0 && st->has_u && err_max < st->eps_e) { done = 1; st->converged = 1; st->stop_reason = 3; }
  if (!done) {
    const double beta = (r_norm * r_norm) / st->rz;
    const double alpha = gamma_new / (delta_new - beta * gamma_new / st->alpha);
    st->rz = gamma_new;
    st->rr = gamma_new;
    st->beta = beta;
    st->alpha = alpha;
    st->pAp = gamma_new / alpha;
    if (st->callback_every > 0 && (it % st->callback_every == 0 || it == 1))
      append_record(st, log, (double)it, dx_max, r_max, st->err_max);
    if (it >= st->max_it) { done = 1; st->converged = 0; st->stop_reason = 0; }
  }
  st->done = done;
}

template <int FLAGS, int HS, int NST, int CW = FUSED_CW>
__global__ void __launch_bounds__((CW + 1) * 32, CW == FUSED_CW ? 2 : 1) cg_fused_kernel(const TileArgs a) {
  constexpr int FUSED_ROW = fused_row(CW), FUSED_STRIP_COLS = fused_row(CW);
  using Cfg = FusedCfg<FLAGS>;
  constexpr bool X2 = Cfg::X2, SHARD = Cfg::SHARD, MAXN = Cfg::MAXN, XS = Cfg::XS, LOAD_U = Cfg::LOAD_U;
  constexpr int NSTREAM = Cfg::NSTREAM, NS = Cfg::NS, NM = Cfg::NM, LO = Cfg::ROWS_BELOW;
  constexpr int STAGE_DOUBLES = HS * NSTREAM * FUSED_ROW;
  constexpr int OFF_P = 0, OFF_R = HS * FUSED_ROW, OFF_X = 2 * HS * FUSED_ROW, OFF_U = 3 * HS * FUSED_ROW;
  constexpr int MI_DX = NM > 1 ? 1 : 0, MI_E = NM > 0 ? NM - 1 : 0;  // slots of |dx|_inf and |x' - u|_inf in acc_m

  const Geom& g = a.g;
  DevState* st = a.st;
  if (st->done) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage_data = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NST * STAGE_DOUBLES * 8);
  uint64_t* empty = full + NST;
  StageMeta* meta = reinterpret_cast<StageMeta*>(empty + NST);
  __shared__ double scratch[(NS + NM) * 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CW);
    }
    mbar_fence_init();
    if (a.cta_clock) a.cta_clock[2 * blockIdx.x] = global_ns();
  }
  __syncthreads();

  double acc_s[NS] = {0.0};  // gamma', delta'
  double acc_m[NM > 0 ? NM : 1] = {0.0};  // F_MAXN: |r'|_inf, |dx|_inf, [|x' - u|_inf]
  bool sent_halo = false;  // F_SHARD: this thread stored into a neighbour rank's rows
  const int y_store_lo = g.ybase, y_store_hi = g.ybase + g.yrows;  // stored rows [lo, hi)
  // Row index of grid row y inside a pitched vector, or -1 where the row does not exist (beyond the domain boundary).
  auto row_index = [&](int y) -> int {
    if (y >= y_store_lo && y < y_store_hi) return y - g.ybase;
    if (SHARD && y == g.ylo - 2 && a.nb_r_below) return g.yrows;      // second halo row below (from the neighbour)
    if (SHARD && y == g.yhi + 1 && a.nb_r_above) return g.yrows + 1;  // second halo row above
    return -1;
  };

  if (warp == CW) {
    // ================================================================ producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const size_t pitch = (size_t)g.pitch;
      const int t_end = a.cta_begin[blockIdx.x + 1];
      for (int t = a.cta_begin[blockIdx.x]; t < t_end; ++t) {
        const Tile tl = a.tiles[t];
        const int col0 = tl.col0, ya = tl.ya, yb = tl.yb;
        const uint32_t row_bytes = (uint32_t)min(FUSED_STRIP_COLS, g.pitch - col0) * 8u;
        const int S = yb - ya + 2 + LO;  // rows ya-2 .. yb+1
        for (int s0 = 0; s0 < S; s0 += HS) {
          mbar_wait(&empty[stage], phase ^ 1u);
          const int nrows = min(HS, S - s0);
          const int y0 = ya - LO + s0;
          StageMeta m;
          m.col0 = col0; m.y0 = y0; m.nrows = nrows; m.flags = (s0 == 0) ? META_TILE_FIRST : 0;
          m.ya = ya; m.yb = yb; m.xlo = tl.xlo; m.pad = 0;
          meta[stage] = m;
          // rows outside the stored range (below the first / above the last boundary row) are not copied: the
          // consumers take them as zeros. Emit rows also need x.
          int stored = 0, inner = 0;
          for (int j = 0; j < nrows; ++j) {
            const int y = y0 + j;
            stored += (row_index(y) >= 0) ? 1 : 0;
            inner += (y >= ya && y < yb) ? 1 : 0;
          }
          const uint32_t bytes = row_bytes * (uint32_t)(2 * stored + (XS ? inner : 0) + (LOAD_U ? inner : 0));
          mbar_arrive_expect_tx(&full[stage], bytes);
          double* sd = stage_data + (size_t)stage * STAGE_DOUBLES;
          for (int j = 0; j < nrows; ++j) {
            const int y = y0 + j;
            const int ri = row_index(y);
            if (ri < 0) continue;
            const size_t off = (size_t)ri * pitch + (size_t)col0;
            bulk_g2s(sd + OFF_P + j * FUSED_ROW, a.p_in + off, row_bytes, &full[stage]);
            bulk_g2s(sd + OFF_R + j * FUSED_ROW, a.r_in + off, row_bytes, &full[stage]);
            if (XS && y >= ya && y < yb) bulk_g2s(sd + OFF_X + j * FUSED_ROW, a.x + off, row_bytes, &full[stage]);
            if (LOAD_U && y >= ya && y < yb) bulk_g2s(sd + OFF_U + j * FUSED_ROW, a.u + off, row_bytes, &full[stage]);
          }
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
      mbar_wait(&empty[stage], phase ^ 1u);
      StageMeta m;
      m.col0 = 0; m.y0 = 0; m.nrows = 0; m.flags = META_END; m.ya = 0; m.yb = 0; m.xlo = 0; m.pad = 0;
      meta[stage] = m;
      mbar_arrive(&full[stage]);
    }
    __syncwarp();
  } else {
    // ================================================================ consumers
    const double cA = g.A, cxk = g.xk, cyk = g.yk;
    const double beta = st->beta, alpha = st->alpha;
    double alpha_prev = 0.0;
    if (X2) alpha_prev = st->alpha_prev;
    const size_t pitch = (size_t)g.pitch;
    const int sc = FUSED_WARP_STEP * warp + 2 * lane;   // this thread's first staged column
    const bool writer = (lane >= 1) && (lane <= 30);    // lanes 0 and 31 only feed their neighbours
    const double2 zero2 = make_double2(0.0, 0.0);

    // the five-point operator at one node, accumulated in the reference's order: diag, left, right, top, bottom
    auto stencil = [&](double c, double l, double r, double t, double b) -> double {
      double v = __dmul_rn(cA, c);
      v = __dadd_rn(v, __dmul_rn(cxk, l));
      v = __dadd_rn(v, __dmul_rn(cxk, r));
      v = __dadd_rn(v, __dmul_rn(cyk, t));
      v = __dadd_rn(v, __dmul_rn(cyk, b));
      return v;
    };

    int stage = 0;
    uint32_t phase = 0;
    // per-tile pipeline state; "1" = one row back, "2" = two rows back
    int ya = 0, yb = 0, x0 = 0;
    size_t col_off = 0;
    double2 P1 = zero2, P2 = zero2;        // p of rows y-1, y-2
    double LP1 = 0.0, RP1 = 0.0;           // horizontal neighbours of P1
    double2 R1 = zero2, R2 = zero2;        // r' of rows y-2, y-3
    double LR1 = 0.0, RR1 = 0.0;           // horizontal neighbours of R1
    double2 r1 = zero2, x1 = zero2, q1 = zero2;  // staged r, x, p_old of row y-1 (masked)
    double2 u1 = zero2;                    // F_U: staged true solution of row y-1
    bool k1a = false, k1b = false;         // row y-1: are this thread's two nodes unknowns
    bool va = false, vb = false;           // the same for the rows of the tile itself (all in one block of the L)
    bool warp_on = true;                   // this warp writes at least one unknown of the tile
    bool st_ok = false;                    // this thread stores in the tile's rows (a writer lane with an unknown)
    bool ma = false, mb = false;           // this thread's two nodes enter the maxima of the tile's rows (writer && va / vb)

    for (;;) {
      mbar_wait(&full[stage], phase);
      const StageMeta m = meta[stage];
      if (m.flags & META_END) break;
      if (m.flags & META_TILE_FIRST) {
        ya = m.ya;
        yb = m.yb;
        x0 = m.col0 + sc - XOFF;
        col_off = (size_t)(m.col0 + sc);
        P1 = P2 = R1 = R2 = r1 = x1 = q1 = u1 = zero2;
        LP1 = RP1 = LR1 = RR1 = 0.0;
        k1a = k1b = false;
        va = (x0 >= m.xlo) && (x0 <= g.n - 1);
        vb = (x0 + 1 >= m.xlo) && (x0 + 1 <= g.n - 1);
        // A warp's results depend on its own 64-column window only, so a warp none of whose written columns
        // (window columns 2 .. 61) is an unknown of this tile has nothing to do: it only hands the stages back. That
        // is 6 of 7 warps in a last strip of a few columns (16383 = 39 x 420 + 3) and the left warps of the strip the
        // re-entrant edge of the L cuts.
        const int xw = m.col0 + FUSED_WARP_STEP * warp - XOFF;  // node of the window's first staged column
        warp_on = (xw + 61 >= m.xlo) && (xw + 2 <= g.n - 1);
        st_ok = writer && (va || vb);
        ma = writer && va;
        mb = writer && vb;
      }
      const double* sd = stage_data + (size_t)stage * STAGE_DOUBLES;
      // One staged row. full_tag: every row of the stage lies in [ya+2, yb), so it is stored, it is a row of the
      // tile, rows y-1 and y-2 are emit rows and x is staged: the per-row predicates fold away. There the staged inputs
      // need no masks either: outside the unknowns r, p_old and x hold exact zeros in memory (so p = 0 + beta*0 is a
      // zero too), and columns beyond the row pitch - stale shared memory - only ever feed results that are masked.
      // Only r' must be masked, because A p does not vanish on boundary nodes.
      auto do_row = [&](const int j, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int y = m.y0 + j;
        // ---- row y arrives: its direction p = r + beta * p_old, zero outside the unknowns
        bool k0a, k0b;
        double2 cur_p = zero2, cur_r = zero2, cur_x = zero2, cur_u = zero2;
        if (FULL) {
          k0a = va;
          k0b = vb;
          cur_p = *reinterpret_cast<const double2*>(sd + OFF_P + j * FUSED_ROW + sc);
          cur_r = *reinterpret_cast<const double2*>(sd + OFF_R + j * FUSED_ROW + sc);
          if (XS) cur_x = *reinterpret_cast<const double2*>(sd + OFF_X + j * FUSED_ROW + sc);
          if (LOAD_U) cur_u = *reinterpret_cast<const double2*>(sd + OFF_U + j * FUSED_ROW + sc);
        } else {
          const bool row_stored = row_index(y) >= 0;
          const bool row_ok = (y >= 1) && (y <= g.m - 1);
          const int xlo = (g.ysplit != 0 && y <= g.ysplit) ? g.xsplit + 1 : 1;
          k0a = row_ok && (x0 >= xlo) && (x0 <= g.n - 1);
          k0b = row_ok && (x0 + 1 >= xlo) && (x0 + 1 <= g.n - 1);
          if (row_stored) {
            cur_p = *reinterpret_cast<const double2*>(sd + OFF_P + j * FUSED_ROW + sc);
            cur_r = *reinterpret_cast<const double2*>(sd + OFF_R + j * FUSED_ROW + sc);
            if (XS && y >= ya && y < yb) cur_x = *reinterpret_cast<const double2*>(sd + OFF_X + j * FUSED_ROW + sc);
            if (LOAD_U && y >= ya && y < yb) cur_u = *reinterpret_cast<const double2*>(sd + OFF_U + j * FUSED_ROW + sc);
          }
          cur_p.x = k0a ? cur_p.x : 0.0;  cur_p.y = k0b ? cur_p.y : 0.0;
          cur_r.x = k0a ? cur_r.x : 0.0;  cur_r.y = k0b ? cur_r.y : 0.0;
          cur_x.x = k0a ? cur_x.x : 0.0;  cur_x.y = k0b ? cur_x.y : 0.0;
          cur_u.x = k0a ? cur_u.x : 0.0;  cur_u.y = k0b ? cur_u.y : 0.0;
        }
        double2 P0;
        P0.x = __dadd_rn(cur_r.x, __dmul_rn(beta, cur_p.x));
        P0.y = __dadd_rn(cur_r.y, __dmul_rn(beta, cur_p.y));
        const double LP0 = __shfl_up_sync(0xffffffffu, P0.y, 1);
        const double RP0 = __shfl_down_sync(0xffffffffu, P0.x, 1);

        // ---- row y-1: A p and r' = r - alpha A p (zero outside the unknowns)
        double2 R0;
        {
          const double ap0 = stencil(P1.x, LP1, P1.y, P0.x, P2.x);
          const double ap1 = stencil(P1.y, P1.x, RP1, P0.y, P2.y);
          R0.x = k1a ? __dsub_rn(r1.x, __dmul_rn(alpha, ap0)) : 0.0;
          R0.y = k1b ? __dsub_rn(r1.y, __dmul_rn(alpha, ap1)) : 0.0;
        }
        const bool emit1 = FULL || ((y - 1 >= ya) && (y - 1 < yb));
        if (emit1 && writer) {
          double2 xm = zero2;  // F_MAXN: x' = x + alpha * p (msg_solver.cpp:105-107)
          if (MAXN) {
            xm.x = __dadd_rn(x1.x, __dmul_rn(alpha, P1.x));
            xm.y = __dadd_rn(x1.y, __dmul_rn(alpha, P1.y));
          }
          if (k1a || k1b) {
            const size_t o = (size_t)(y - 1 - g.ybase) * pitch + col_off;
            st2_out(a.r_out + o, R0);
            st2_out(a.p_out + o, P1);
            if (MAXN) st2_out(a.x + o, xm);
            if (X2) {  // x += alpha_prev * p_old, then += alpha * p: the reference's order of additions
              double2 xn;
              xn.x = __dadd_rn(__dadd_rn(x1.x, __dmul_rn(alpha_prev, q1.x)), __dmul_rn(alpha, P1.x));
              xn.y = __dadd_rn(__dadd_rn(x1.y, __dmul_rn(alpha_prev, q1.y)), __dmul_rn(alpha, P1.y));
              st2_out(a.x + o, xn);
            }
            if (SHARD && !FULL) {
              // the slab's two first / last rows also land in the neighbours' halo rows (NVLink stores)
              const int ye = y - 1;
              if (a.nb_r_below && (ye == g.ylo || ye == g.ylo + 1)) {
                double* dr = (ye == g.ylo ? a.nb_r_below : a.nb_r_below2) + col_off;
                double* dp = (ye == g.ylo ? a.nb_p_below : a.nb_p_below2) + col_off;
                st2(dr, R0);
                st2(dp, P1);
                sent_halo = true;
              }
              if (a.nb_r_above && (ye == g.yhi - 1 || ye == g.yhi - 2)) {
                double* dr = (ye == g.yhi - 1 ? a.nb_r_above : a.nb_r_above2) + col_off;
                double* dp = (ye == g.yhi - 1 ? a.nb_p_above : a.nb_p_above2) + col_off;
                st2(dr, R0);
                st2(dp, P1);
                sent_halo = true;
              }
            }
          }
          acc_s[0] = fma(R0.x, R0.x, acc_s[0]);
          acc_s[0] = fma(R0.y, R0.y, acc_s[0]);
          if (MAXN) {
            // |r'|_inf, |x' - x|_inf, |x' - u|_inf over the unknowns (msg_solver.cpp:121-139); after a FULL stage x1 / u1 are
            // unmasked, hence the selects
            acc_m[0] = fmax(acc_m[0], fmax(fabs(R0.x), fabs(R0.y)));
            const double d0 = k1a ? fabs(__dsub_rn(xm.x, x1.x)) : 0.0;
            const double d1 = k1b ? fabs(__dsub_rn(xm.y, x1.y)) : 0.0;
            acc_m[MI_DX] = fmax(acc_m[MI_DX], fmax(d0, d1));
            if (LOAD_U) {
              const double e0 = k1a ? fabs(__dsub_rn(xm.x, u1.x)) : 0.0;
              const double e1 = k1b ? fabs(__dsub_rn(xm.y, u1.y)) : 0.0;
              acc_m[MI_E] = fmax(acc_m[MI_E], fmax(e0, e1));
            }
          }
        }
        const double RR0 = __shfl_down_sync(0xffffffffu, R0.x, 1);
        const double LR0 = __shfl_up_sync(0xffffffffu, R0.y, 1);
        // ---- row y-2: A r' and delta' += r'.A r'
        const bool emit2 = FULL || ((y - 2 >= ya) && (y - 2 < yb));
        if (emit2 && writer) {
          const double w0 = stencil(R1.x, LR1, R1.y, R0.x, R2.x);
          const double w1 = stencil(R1.y, R1.x, RR1, R0.y, R2.y);
          acc_s[1] = fma(R1.x, w0, acc_s[1]);
          acc_s[1] = fma(R1.y, w1, acc_s[1]);
        }
        R2 = R1;  R1 = R0;  LR1 = LR0;  RR1 = RR0;

        // ---- shift the pipeline
        P2 = P1;  P1 = P0;  LP1 = LP0;  RP1 = RP0;
        r1 = cur_r;  x1 = cur_x;  q1 = cur_p;  u1 = cur_u;
        k1a = k0a;  k1b = k0b;
      };
      // The same row for a FULL stage (all but a tile's first and last), written branch-free: the ncu source page of the
      // generic form showed 8 BRA + 5.5 BSYNC per warp-row from the lane-divergent `writer` regions and 10 % of the
      // samples in branch resolution, in a kernel that waits for data only 12 % of the time (even flavour). Every lane
      // computes everything; lanes 0 / 31 (and columns that are no unknowns) are taken out with selects on the two sums
      // and a predicate on the stores. Same operations on the same values in the same order as do_row: bit-identical.
      double2 G1 = zero2;  // r' of row y-2 as it enters the sums: zero in the lanes that do not write
      size_t eo = 0;       // offset of the row being emitted
      auto do_row_full = [&](const int j) {
        const double2 cur_p = *reinterpret_cast<const double2*>(sd + OFF_P + j * FUSED_ROW + sc);
        const double2 cur_r = *reinterpret_cast<const double2*>(sd + OFF_R + j * FUSED_ROW + sc);
        double2 cur_x = zero2, cur_u = zero2;
        if (XS) cur_x = *reinterpret_cast<const double2*>(sd + OFF_X + j * FUSED_ROW + sc);
        if (LOAD_U) cur_u = *reinterpret_cast<const double2*>(sd + OFF_U + j * FUSED_ROW + sc);
        double2 P0;
        P0.x = __dadd_rn(cur_r.x, __dmul_rn(beta, cur_p.x));
        P0.y = __dadd_rn(cur_r.y, __dmul_rn(beta, cur_p.y));
        const double LP0 = __shfl_up_sync(0xffffffffu, P0.y, 1);
        const double RP0 = __shfl_down_sync(0xffffffffu, P0.x, 1);
        const double ap0 = stencil(P1.x, LP1, P1.y, P0.x, P2.x);
        const double ap1 = stencil(P1.y, P1.x, RP1, P0.y, P2.y);
        double2 R0;
        R0.x = va ? __dsub_rn(r1.x, __dmul_rn(alpha, ap0)) : 0.0;
        R0.y = vb ? __dsub_rn(r1.y, __dmul_rn(alpha, ap1)) : 0.0;
        double2 xn = zero2;
        if (X2) {  // x += alpha_prev * p_old, then += alpha * p: the reference's order of additions
          xn.x = __dadd_rn(__dadd_rn(x1.x, __dmul_rn(alpha_prev, q1.x)), __dmul_rn(alpha, P1.x));
          xn.y = __dadd_rn(__dadd_rn(x1.y, __dmul_rn(alpha_prev, q1.y)), __dmul_rn(alpha, P1.y));
        }
        if (MAXN) {  // x' = x + alpha * p; its distance from x and from u enters the maxima (selects: unmasked inputs)
          xn.x = __dadd_rn(x1.x, __dmul_rn(alpha, P1.x));
          xn.y = __dadd_rn(x1.y, __dmul_rn(alpha, P1.y));
          const double d0 = ma ? fabs(__dsub_rn(xn.x, x1.x)) : 0.0;
          const double d1 = mb ? fabs(__dsub_rn(xn.y, x1.y)) : 0.0;
          acc_m[MI_DX] = fmax(acc_m[MI_DX], fmax(d0, d1));
          if (LOAD_U) {
            const double e0 = ma ? fabs(__dsub_rn(xn.x, u1.x)) : 0.0;
            const double e1 = mb ? fabs(__dsub_rn(xn.y, u1.y)) : 0.0;
            acc_m[MI_E] = fmax(acc_m[MI_E], fmax(e0, e1));
          }
        }
        st2_out_if(st_ok, a.r_out + eo, R0);
        st2_out_if(st_ok, a.p_out + eo, P1);
        if (XS) st2_out_if(st_ok, a.x + eo, xn);
        eo += pitch;
        const double RR0 = __shfl_down_sync(0xffffffffu, R0.x, 1);
        const double LR0 = __shfl_up_sync(0xffffffffu, R0.y, 1);
        const double w0 = stencil(R1.x, LR1, R1.y, R0.x, R2.x);
        const double w1 = stencil(R1.y, R1.x, RR1, R0.y, R2.y);
        double2 G0;
        G0.x = writer ? R0.x : 0.0;
        G0.y = writer ? R0.y : 0.0;
        acc_s[0] = fma(G0.x, G0.x, acc_s[0]);
        acc_s[0] = fma(G0.y, G0.y, acc_s[0]);
        acc_s[1] = fma(G1.x, w0, acc_s[1]);
        acc_s[1] = fma(G1.y, w1, acc_s[1]);
        if (MAXN) acc_m[0] = fmax(acc_m[0], fmax(fabs(G0.x), fabs(G0.y)));
        G1 = G0;
        R2 = R1;  R1 = R0;  LR1 = LR0;  RR1 = RR0;
        P2 = P1;  P1 = P0;  LP1 = LP0;  RP1 = RP0;
        r1 = cur_r;  x1 = cur_x;  q1 = cur_p;  u1 = cur_u;
      };
      // (sharded plans: the emit rows ya+1 and yb-2 may be rows the neighbours need, so FULL stays clear of them)
      if (!warp_on) {
        // nothing to compute for this tile
      } else if (m.nrows == HS && m.y0 >= ya + (SHARD ? 3 : 2) && m.y0 + HS <= yb - (SHARD ? 1 : 0)) {
        G1.x = writer ? R1.x : 0.0;
        G1.y = writer ? R1.y : 0.0;
        eo = (size_t)(m.y0 - 1 - g.ybase) * pitch + col_off;
#pragma unroll
        for (int j = 0; j < HS; ++j) do_row_full(j);
        k1a = va;  // (what the generic form would have left behind)
        k1b = vb;
      } else {
#pragma unroll 1
        for (int j = 0; j < m.nrows; ++j) do_row(j, cuda::std::false_type{});
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == NST) { stage = 0; phase ^= 1u; }
    }
  }

  if (tid == 0 && a.cta_clock) a.cta_clock[2 * blockIdx.x + 1] = global_ns();
  if (SHARD && sent_halo) __threadfence_system();  // remote halo stores before the exit ticket
  if (!grid_reduce<NS, NM>(acc_s, acc_m, a.partials, st, scratch)) return;
  bool stop_req = poll_stop(st, a.stop_flag);
  double gamma = acc_s[0];
  double delta = acc_s[1];
  if (SHARD) {
    // this rank's two sums (and, under the max-norm rules, its maxima) to every rank, everyone's back; the slot alternates
    // from iteration to iteration (x-deferral: with the flavour; max-norm rules: with the iteration count) so that a fast
    // rank's next publication cannot overwrite values a slow rank is still reading
    const int phase = MAXN ? (st->it & 1) : (X2 ? 1 : 0);
    double mine[2] = {gamma, delta}, total[4], mx_all[3];
    unsigned long long* tr = a.peer_trace ? a.peer_trace + 4 * (size_t)(st->it % PEER_TRACE_CAP) : nullptr;
    if (tr) tr[0] = global_ns();
    peer_publish<2, NM>(a.peers, st, phase, mine, acc_m, stop_req);
    if (tr) tr[1] = global_ns();
    if (!peer_collect(st, a.peers, phase, total, &stop_req, mx_all)) return;
    if (tr) tr[2] = global_ns();
    gamma = total[0];
    delta = total[1];
    if (MAXN) finalize_fused_maxnorm(st, a.cb_log, gamma, delta, mx_all[0], mx_all[MI_DX], LOAD_U ? mx_all[MI_E] : DBL_MAX);
    else finalize_fused(st, gamma, delta, FLAGS);
    apply_stop(st, stop_req);
    if (tr) tr[3] = global_ns();
    return;
  }
  if (MAXN) {
    finalize_fused_maxnorm(st, a.cb_log, gamma, delta, acc_m[0], acc_m[MI_DX], LOAD_U ? acc_m[MI_E] : DBL_MAX);
    apply_stop(st, stop_req);
    return;
  }
  finalize_fused(st, gamma, delta, FLAGS);
  apply_stop(st, stop_req);
}

}  // namespace b200cg
