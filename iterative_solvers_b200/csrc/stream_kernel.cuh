// The matrix-free hot kernel (sm_100a): a persistent, warp-specialised, bulk-async-copy fed stencil sweep.
//
//   producer warp : one elected lane walks this CTA's tiles and feeds a ring of shared-memory stages with
//                   cp.async.bulk (UBLKCP) row copies of r, p_old (and x, u), completion counted by mbarriers
//                   (expect_tx). No registers hold in-flight data; NST stages of HS rows are in flight per CTA.
//   consumer warps: 8 warps, each thread owns 2 adjacent columns of a 512-column strip and marches down the
//                   rows: p = r + beta*p_old for its columns, horizontal neighbours by warp shuffle (warp-edge
//                   lanes recompute theirs from the staged r, p_old), vertical neighbours from registers.
//                   Warps only meet at the stage mbarriers - there is no per-row __syncthreads.
// A strip loads 4 extra columns on each side (one 32-byte sector), so every output column finds its
// neighbours inside the CTA. Stage metadata written by the producer tells consumers what a stage holds, so the
// tile order is the producer's business alone: it walks the list of equal-work tiles the host dealt to this CTA.
#pragma once
#include <cuda/std/cstdint>
#include <cuda/std/type_traits>
#include "kernels_common.cuh"

namespace b200cg {

constexpr int CONS_WARPS = 8;
constexpr int CONS_THREADS = CONS_WARPS * 32;  // 256: 2 columns each = STRIP_LOAD
constexpr int STREAM_THREADS = CONS_THREADS + 32;
constexpr int ROW_BYTES = STRIP_LOAD * 8;  // 4096

struct StageMeta {
  int col0;    // storage column of the strip's first loaded column
  int y0;      // grid row of the stage's first row
  int nrows;   // rows in this stage (<= HS)
  int flags;   // META_*
  int ya, yb;  // emit rows of the tile: [ya, yb)
  int xlo;     // first unknown x in these rows
  int pad;
};
enum { META_TILE_FIRST = 1, META_END = 2 };

// ---------------------------------------------------------------------------------- mbarrier / bulk copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(addr),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int MODE, int FLAGS>
struct StreamCfg {
  static constexpr bool LOAD_R = (MODE != MODE_APPLY) || (FLAGS & (F_SUB_B | F_REPORT));
  static constexpr bool LOAD_X = (MODE == MODE_UPD) && !(FLAGS & F_NOX);
  static constexpr bool LOAD_U = (FLAGS & F_U) != 0;
  static constexpr bool REPORT = (FLAGS & F_REPORT) != 0;
  static constexpr int NSTREAM = 1 + (LOAD_R ? 1 : 0) + (LOAD_X ? 1 : 0) + (LOAD_U ? 1 : 0);
  static constexpr int NS = (MODE == MODE_DOT) ? 2 : (MODE == MODE_UPD ? (REPORT ? 3 : 1) : (REPORT ? 2 : 0));
  // max-norms (|r|, |dx|, |x - u|) feed the MAXNORM stop rules only; the x-deferral flavours exist only under REL_L2
  static constexpr bool DEFERRAL = (FLAGS & (F_NOX | F_X2)) != 0;
  static constexpr int NM = (MODE == MODE_UPD && !DEFERRAL) ? (LOAD_U ? 3 : 2) : 0;
};

template <int MODE, int FLAGS, int HS, int NST>
constexpr size_t stream_smem_bytes() {  // (independent of CTAS)
  return (size_t)NST * HS * StreamCfg<MODE, FLAGS>::NSTREAM * ROW_BYTES + (size_t)NST * (16 + sizeof(StageMeta)) + 128;
}

template <int MODE, int FLAGS, int HS, int NST, int CTAS>
__global__ void __launch_bounds__(STREAM_THREADS, CTAS) cg_stream_kernel(const TileArgs a) {
  // x-deferral (REL_L2 rule without report): x is only touched every other iteration. F_NOX = even iteration,
  // x untouched, the state remembers alpha; F_X2 = odd iteration, x += alpha_prev*p_old + alpha*p in the
  // reference's order of additions, so x is bit-identical to updating it every iteration.
  constexpr bool NOX = (FLAGS & F_NOX) != 0, X2 = (FLAGS & F_X2) != 0;
  using Cfg = StreamCfg<MODE, FLAGS>;
  constexpr bool LOAD_R = Cfg::LOAD_R, LOAD_X = Cfg::LOAD_X, LOAD_U = Cfg::LOAD_U, REPORT = Cfg::REPORT;
  constexpr int NSTREAM = Cfg::NSTREAM, NS = Cfg::NS, NM = Cfg::NM;
  constexpr int STAGE_DOUBLES = HS * NSTREAM * STRIP_LOAD;
  // stream order inside a stage: p, [r], [x], [u]; each HS rows of STRIP_LOAD doubles
  constexpr int OFF_P = 0, OFF_R = HS * STRIP_LOAD, OFF_X = (1 + (LOAD_R ? 1 : 0)) * HS * STRIP_LOAD,
                OFF_U = (1 + (LOAD_R ? 1 : 0) + (LOAD_X ? 1 : 0)) * HS * STRIP_LOAD;

  const Geom& g = a.g;
  DevState* st = a.st;
  if (MODE == MODE_APPLY) {
    if (REPORT && st->report_pending == 0) return;  // no report pending
  } else {
    if (st->done) return;
  }

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage_data = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NST * STAGE_DOUBLES * 8);
  uint64_t* empty = full + NST;
  StageMeta* meta = reinterpret_cast<StageMeta*>(empty + NST);
  __shared__ double scratch[(NS + NM > 0 ? NS + NM : 1) * 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], CONS_WARPS);
    }
    mbar_fence_init();
    if (a.cta_clock) a.cta_clock[2 * blockIdx.x] = global_ns();
  }
  __syncthreads();

  double acc_s[NS > 0 ? NS : 1] = {0.0};
  double acc_m[NM > 0 ? NM : 1] = {0.0};
  bool sent_halo = false;  // this thread stored into a neighbour rank's halo row (peer memory)

  if (warp == CONS_WARPS) {
    // ================================================================ producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const size_t pitch = (size_t)g.pitch;
      const int t_end = a.cta_begin[blockIdx.x + 1];
      for (int t = a.cta_begin[blockIdx.x]; t < t_end; ++t) {
        const Tile tl = a.tiles[t];
        const int col0 = tl.col0, ya = tl.ya, yb = tl.yb, xlo = tl.xlo;
        const uint32_t row_bytes = (uint32_t)min(STRIP_LOAD, g.pitch - col0) * 8u;
        const int S = yb - ya + 2;  // rows ya-1 .. yb
        for (int s0 = 0; s0 < S; s0 += HS) {
          mbar_wait(&empty[stage], phase ^ 1u);
          const int nrows = min(HS, S - s0);
          const int y0 = ya - 1 + s0;
          StageMeta m;
          m.col0 = col0; m.y0 = y0; m.nrows = nrows; m.flags = (s0 == 0) ? META_TILE_FIRST : 0;
          m.ya = ya; m.yb = yb; m.xlo = xlo; m.pad = 0;
          meta[stage] = m;
          // inner rows (emit rows) also need x / u
          int inner = 0;
          for (int j = 0; j < nrows; ++j) inner += (y0 + j >= ya && y0 + j < yb) ? 1 : 0;
          const uint32_t bytes = row_bytes * (uint32_t)(nrows * (1 + (LOAD_R ? 1 : 0)) + inner * ((LOAD_X ? 1 : 0) + (LOAD_U ? 1 : 0)));
          mbar_arrive_expect_tx(&full[stage], bytes);
          double* sd = stage_data + (size_t)stage * STAGE_DOUBLES;
          for (int j = 0; j < nrows; ++j) {
            const int y = y0 + j;
            const size_t off = (size_t)(y - g.ybase) * pitch + (size_t)col0;
            bulk_g2s(sd + OFF_P + j * STRIP_LOAD, a.p_in + off, row_bytes, &full[stage]);
            if (LOAD_R) bulk_g2s(sd + OFF_R + j * STRIP_LOAD, a.r_in + off, row_bytes, &full[stage]);
            if (y >= ya && y < yb) {
              if (LOAD_X) bulk_g2s(sd + OFF_X + j * STRIP_LOAD, a.x + off, row_bytes, &full[stage]);
              if (LOAD_U) bulk_g2s(sd + OFF_U + j * STRIP_LOAD, a.u + off, row_bytes, &full[stage]);
            }
          }
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
      // end marker
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
    double alpha = 0.0, beta = 0.0, alpha_prev = 0.0;
    if (MODE != MODE_APPLY) {
      beta = st->beta;
      if (MODE == MODE_UPD) alpha = st->alpha;
      if (X2) alpha_prev = st->alpha_prev;
    }
    const size_t pitch = (size_t)g.pitch;
    const bool is_out = (tid >= STRIP_HALO / 2) && (tid < CONS_THREADS - STRIP_HALO / 2);
    const int c2 = 2 * tid;  // column inside the strip
    // the one staged column a warp-edge lane needs from outside its warp
    const int c_edge = (lane == 0) ? max(c2 - 1, 0) : ((lane == 31) ? min(c2 + 2, STRIP_LOAD - 1) : c2);

    int stage = 0;
    uint32_t phase = 0;
    // per-tile state
    bool v0 = false, v1 = false;
    bool warp_on = true;  // some column of this warp is an unknown of the tile
    int ya = 0;
    size_t eoff = 0;
    double2 pm = make_double2(0.0, 0.0), pc = make_double2(0.0, 0.0);
    double2 r_prev = make_double2(0.0, 0.0), x_prev = make_double2(0.0, 0.0), u_prev = make_double2(0.0, 0.0);
    double2 q_prev = make_double2(0.0, 0.0);  // p_old of the row being emitted (X2 only)
    double Lp = 0.0, Rp = 0.0;

    for (;;) {
      mbar_wait(&full[stage], phase);
      const StageMeta m = meta[stage];
      if (m.flags & META_END) break;
      if (m.flags & META_TILE_FIRST) {
        const int x0 = m.col0 + c2 - XOFF;
        v0 = is_out && (x0 >= m.xlo) && (x0 <= g.n - 1);
        v1 = is_out && (x0 + 1 >= m.xlo) && (x0 + 1 <= g.n - 1);
        ya = m.ya;
        eoff = (size_t)(ya - g.ybase) * pitch + (size_t)(m.col0 + c2);
        pm = pc = make_double2(0.0, 0.0);
        // a warp's results depend on its own 64 columns (and staged neighbours) only: a warp without a single unknown
        // in this tile - the right half of a half-empty last strip, the columns left of the L's re-entrant edge - only
        // hands the stages back
        warp_on = __any_sync(0xffffffffu, v0 || v1);
      }
      const double* sd = stage_data + (size_t)stage * STAGE_DOUBLES;
      // One staged row. full_tag: every row of the stage is an interior row of its tile (emit always, x and u
      // staged), so the per-row predicates fold away.
      auto do_row = [&](const int j, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int y = m.y0 + j;
        const double* prow = sd + OFF_P + j * STRIP_LOAD;
        const double* rrow = sd + OFF_R + j * STRIP_LOAD;
        const double2 cur_p = *reinterpret_cast<const double2*>(prow + c2);
        double2 cur_r = make_double2(0.0, 0.0), cur_x = make_double2(0.0, 0.0), cur_u = make_double2(0.0, 0.0);
        if (LOAD_R) cur_r = *reinterpret_cast<const double2*>(rrow + c2);
        const bool inner = FULL || ((y >= ya) && (y < m.yb));
        if (LOAD_X && inner) cur_x = *reinterpret_cast<const double2*>(sd + OFF_X + j * STRIP_LOAD + c2);
        if (LOAD_U && inner) cur_u = *reinterpret_cast<const double2*>(sd + OFF_U + j * STRIP_LOAD + c2);
        // direction of this row: p = r + beta * p_old (matrix_free_system.cpp:436-438)
        double2 pn;
        if (MODE == MODE_APPLY) {
          pn = cur_p;
        } else {
          pn.x = __dadd_rn(cur_r.x, __dmul_rn(beta, cur_p.x));
          pn.y = __dadd_rn(cur_r.y, __dmul_rn(beta, cur_p.y));
        }
        // horizontal neighbours: shuffle inside the warp; a warp-edge lane recomputes the one it lacks from the
        // staged rows (the strip's two outermost threads, which have no such neighbour, are halo threads whose
        // results are masked below)
        double pe = 0.0;
        if (lane == 0 || lane == 31) {  // (all 32 lanes loading 8-byte words at a 16-byte stride = a 2-way bank conflict)
          pe = prow[c_edge];
          if (MODE != MODE_APPLY) pe = __dadd_rn(rrow[c_edge], __dmul_rn(beta, pe));
        }
        double L = __shfl_up_sync(0xffffffffu, pn.y, 1);
        double R = __shfl_down_sync(0xffffffffu, pn.x, 1);
        L = (lane == 0) ? pe : L;
        R = (lane == 31) ? pe : R;

        if (FULL || y > ya) {
          // emit row y-1: centre pc, bottom pm, top pn; accumulation order diag, left, right, top, bottom
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
          // Columns that are not unknowns of this tile (boundary, excluded quadrant, halo columns, columns past
          // the row pitch whose staged values are stale) contribute exact zeros: mask every operand once.
          ap0 = v0 ? ap0 : 0.0;
          ap1 = v1 ? ap1 : 0.0;
          const double p0 = v0 ? pc.x : 0.0, p1 = v1 ? pc.y : 0.0;
          const double r0 = v0 ? r_prev.x : 0.0, r1 = v1 ? r_prev.y : 0.0;
          const bool st_ok = v0 || v1;
          if (MODE == MODE_DOT) {
            acc_s[0] = fma(p0, ap0, acc_s[0]);
            acc_s[0] = fma(p1, ap1, acc_s[0]);
            acc_s[1] = fma(r0, p0, acc_s[1]);
            acc_s[1] = fma(r1, p1, acc_s[1]);
          } else if (MODE == MODE_UPD) {
            // x += alpha p; r -= alpha Ap (matrix_free_system.cpp:422-429)
            double xo0 = v0 ? x_prev.x : 0.0, xo1 = v1 ? x_prev.y : 0.0;
            if (X2) {  // the update the previous (NOX) iteration left pending: x += alpha_prev * p_old
              xo0 = __dadd_rn(xo0, __dmul_rn(alpha_prev, v0 ? q_prev.x : 0.0));
              xo1 = __dadd_rn(xo1, __dmul_rn(alpha_prev, v1 ? q_prev.y : 0.0));
            }
            double2 xn, rn;
            xn.x = __dadd_rn(xo0, __dmul_rn(alpha, p0));
            xn.y = __dadd_rn(xo1, __dmul_rn(alpha, p1));
            rn.x = __dsub_rn(r0, __dmul_rn(alpha, ap0));
            rn.y = __dsub_rn(r1, __dmul_rn(alpha, ap1));
            if (FULL) {  // predicated stores: no divergent region in the unrolled path
              if (!NOX) st2_out_if(st_ok, a.x + eoff, xn);
              st2_out_if(st_ok, a.r_out + eoff, rn);
              st2_out_if(st_ok, a.p_out + eoff, make_double2(p0, p1));
            } else if (st_ok) {
              if (!NOX) st2_out(a.x + eoff, xn);
              st2_out(a.r_out + eoff, rn);
              st2_out(a.p_out + eoff, make_double2(p0, p1));
              // peer-memory halo: the slab's first / last row also lands in the neighbour's halo row (NVLink stores).
              // Those are a tile's first / last emit rows, which a FULL stage never holds.
              const int ye = y - 1;
              const int cs = m.col0 + c2;
              if (!FULL && a.nb_r_below && ye == g.ylo) {
                st2(a.nb_r_below + cs, rn);
                st2(a.nb_p_below + cs, make_double2(p0, p1));
                sent_halo = true;
              }
              if (!FULL && a.nb_r_above && ye == g.yhi - 1) {
                st2(a.nb_r_above + cs, rn);
                st2(a.nb_p_above + cs, make_double2(p0, p1));
                sent_halo = true;
              }
            }
            acc_s[0] = fma(rn.x, rn.x, acc_s[0]);
            acc_s[0] = fma(rn.y, rn.y, acc_s[0]);
            const double d0 = __dsub_rn(xn.x, xo0);  // msg_solver.cpp:124-129
            const double d1 = __dsub_rn(xn.y, xo1);
            if (NM > 0) {
              acc_m[0] = fmax(acc_m[0], fmax(fabs(rn.x), fabs(rn.y)));
              acc_m[NM > 1 ? 1 : 0] = fmax(acc_m[NM > 1 ? 1 : 0], fmax(fabs(d0), fabs(d1)));
            }
            if (REPORT) {
              acc_s[1] = fma(d0, d0, acc_s[1]);
              acc_s[1] = fma(d1, d1, acc_s[1]);
            }
            if (LOAD_U) {
              const double e0 = v0 ? __dsub_rn(xn.x, u_prev.x) : 0.0;  // msg_solver.cpp:132-139
              const double e1 = v1 ? __dsub_rn(xn.y, u_prev.y) : 0.0;
              acc_m[NM > 0 ? NM - 1 : 0] = fmax(acc_m[NM > 0 ? NM - 1 : 0], fmax(fabs(e0), fabs(e1)));
              if (REPORT) {
                acc_s[2] = fma(e0, e0, acc_s[2]);
                acc_s[2] = fma(e1, e1, acc_s[2]);
              }
            }
          } else {
            if (REPORT) {
              const double d0 = __dsub_rn(r0, ap0);  // b - A x, matrix_free_system.cpp:459-463
              const double d1 = __dsub_rn(r1, ap1);
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
                o.x = __dsub_rn(ap0, r0);  // A x - b, dirichlet_solver.cpp:156-158
                o.y = __dsub_rn(ap1, r1);
              } else {
                o.x = ap0;
                o.y = ap1;
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
        if (X2) q_prev = cur_p;
      };
      // FULL: rows y0 .. y0+HS-1 all in [ya, yb) and the emitted rows y0-1 .. y0+HS-2 all in (ya, yb-1), so neither
      // the tile's first emit row (a slab's first row goes to the neighbour's halo) nor its last is handled here
      if (!warp_on) {
        // nothing to compute for this tile
      } else if (m.nrows == HS && !(m.flags & META_TILE_FIRST) && m.y0 - 1 > m.ya && m.y0 + HS <= m.yb) {
#pragma unroll
        for (int j = 0; j < HS; ++j) do_row(j, cuda::std::true_type{});
      } else {
#pragma unroll 1
        for (int j = 0; j < m.nrows; ++j) do_row(j, cuda::std::false_type{});
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == NST) { stage = 0; phase ^= 1u; }
    }
  }

  // end of this CTA's sweep: the first consumer warp is as good a witness as any (all finish within a stage)
  if (tid == 0 && a.cta_clock) a.cta_clock[2 * blockIdx.x + 1] = global_ns();
  if (MODE == MODE_UPD && sent_halo) __threadfence_system();  // remote halo stores before the exit ticket
  if (NS + NM == 0) return;
  if (!grid_reduce<NS, NM>(acc_s, acc_m, a.partials, st, scratch)) return;
  // ---- one thread: turn the totals into the next scalars
  if (a.defer == 2) {  // peer memory: publish this rank's totals to every rank, then collect everyone's
    peer_publish<NS, NM>(a.peers, st, MODE == MODE_DOT ? 0 : 1, acc_s, acc_m, MODE == MODE_UPD && poll_stop(st, a.stop_flag));
    peer_finalize(st, a.cb_log, a.peers, MODE == MODE_DOT ? 1 : 2, FLAGS);
    return;
  }
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
    const bool stop_req = poll_stop(st, a.stop_flag);
    finalize_update(st, a.cb_log, acc_s[0], NM > 0 ? acc_m[0] : 0.0, NM > 1 ? acc_m[NM > 1 ? 1 : 0] : 0.0,
                    LOAD_U ? acc_m[NM > 0 ? NM - 1 : 0] : DBL_MAX,
                    REPORT ? acc_s[1] : 0.0, (REPORT && LOAD_U) ? acc_s[2] : 0.0, REPORT);
    note_x_deferral(st, FLAGS);
    apply_stop(st, stop_req);
  } else if (REPORT) {
    finalize_report(st, a.cb_log, acc_s[0], LOAD_U ? acc_s[1] : 0.0, LOAD_U);
  }
}

}  // namespace b200cg
