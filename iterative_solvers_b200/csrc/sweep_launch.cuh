// Launch side of the sweep kernel: launch shapes, flavour -> tile table, common kernel arguments, and the
// NCCL-fallback reduce-and-finalize step of sharded plans. Header-only so that every translation unit instantiates
// just the kernel flavours it launches.
#pragma once
#include "plan.h"

namespace b200cg {

// Launch shapes of the sweep kernel. A shape = rows per stage (HS), stages (NST), resident CTAs per SM (CTAS);
// the bulk-copy destinations take ~96 KB per CTA at 2 CTAs/SM and ~64-72 KB at 3. Shape 0 is the default; the
// others exist for the hot flavours only and are selected per plan with B200CG_SHAPE_DOT / _UPD / _NOX
// (tuning knobs).
template <int NSTREAM, int SHAPE>
struct ShapeOf;
template <int NSTREAM>
struct ShapeOf<NSTREAM, 0> {  // HS = 2, 2 CTAs/SM
  static constexpr int HS = 2, CTAS = 2, NST = NSTREAM == 1 ? 8 : (NSTREAM == 2 ? 6 : (NSTREAM == 3 ? 4 : 3));
};
template <int NSTREAM>
struct ShapeOf<NSTREAM, 1> {  // HS = 2, 3 CTAs/SM
  static constexpr int HS = 2, CTAS = 3, NST = NSTREAM == 1 ? 8 : (NSTREAM == 2 ? 4 : 3);
};
template <int NSTREAM>
struct ShapeOf<NSTREAM, 2> {  // HS = 4, 2 CTAs/SM
  static constexpr int HS = 4, CTAS = 2, NST = NSTREAM == 1 ? 6 : (NSTREAM == 2 ? 3 : 2);
};
template <int NSTREAM>
struct ShapeOf<NSTREAM, 3> {  // HS = 4, 3 CTAs/SM (two-stream flavours only)
  static constexpr int HS = 4, CTAS = 3, NST = 2;
};

template <int MODE, int FLAGS, int SHAPE>
static int launch_shape(b200cg_plan_s* P, TileArgs a, cudaStream_t s) {
  using Sh = ShapeOf<StreamCfg<MODE, FLAGS>::NSTREAM, SHAPE>;
  auto kernel = cg_stream_kernel<MODE, FLAGS, Sh::HS, Sh::NST, Sh::CTAS>;
  constexpr size_t smem = stream_smem_bytes<MODE, FLAGS, Sh::HS, Sh::NST>();
  static thread_local bool configured[64] = {};
  const int dev = P->desc.device & 63;
  if (!configured[dev]) {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev] = true;
  }
  // flavour: 0 = dot phase, 1 = update without x (NOX), 2 = everything else (its table is cut for 2 CTAs/SM)
  constexpr int fl = (MODE == MODE_DOT && FLAGS == 0) ? 0 : ((MODE == MODE_UPD && FLAGS == F_NOX) ? 1 : 2);
  const TileTable& tt = P->tile_tab[fl];
  if (tt.n_tiles <= 0) return B200CG_OK;
  a.tiles = tt.d_tiles;
  a.cta_begin = tt.d_cta_begin;
  // only the three hot kernels stamp their CTAs (flavour 2 = the x-touching update of the default path)
  constexpr bool stamped = fl < 2 || (MODE == MODE_UPD && (FLAGS == F_X2 || FLAGS == 0));
  a.cta_clock = stamped ? P->d_clock[fl] : nullptr;
  if (stamped) P->clock_ctas[fl] = tt.grid;
  kernel<<<tt.grid, STREAM_THREADS, smem, s>>>(a);
  CU(cudaGetLastError());
  return B200CG_OK;
}

// hot flavours get every shape; the rest run shape 0
template <int MODE, int FLAGS>
static int launch_tile(b200cg_plan_s* P, const TileArgs& a, cudaStream_t s) {
  constexpr bool hot = (MODE == MODE_DOT && FLAGS == 0) ||
                       (MODE == MODE_UPD && (FLAGS == 0 || FLAGS == F_NOX || FLAGS == F_X2));
  if constexpr (hot) {
    constexpr int nstream = StreamCfg<MODE, FLAGS>::NSTREAM;
    const int shape = MODE == MODE_DOT ? P->shape_dot : (FLAGS == F_NOX ? P->shape_nox : P->shape_upd);
    switch (shape) {
      case 1: return launch_shape<MODE, FLAGS, 1>(P, a, s);
      case 2: return launch_shape<MODE, FLAGS, 2>(P, a, s);
      case 3:
        if constexpr (nstream == 2) return launch_shape<MODE, FLAGS, 3>(P, a, s);
        break;
      default: break;
    }
  }
  return launch_shape<MODE, FLAGS, 0>(P, a, s);
}

// The single-sweep iteration (fused_kernel.cuh): 8-warp CTAs, 2 CTAs/SM, 80-109 KB of copy destinations per CTA.
template <int FLAGS, int HS, int NST, int CW>
static int launch_fused_cfg(b200cg_plan_s* P, TileArgs a, cudaStream_t s) {
  auto kernel = cg_fused_kernel<FLAGS, HS, NST, CW>;
  constexpr size_t smem = fused_smem_bytes<FLAGS, HS, NST, CW>();
  static_assert(smem + 1024 <= (CW == FUSED_CW ? 114 : 227) * 1024, "two CTAs per SM (one for the wide geometry)");
  static thread_local bool configured[64] = {};
  const int dev = P->desc.device & 63;
  if (!configured[dev]) {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev] = true;
  }
  const TileTable& tt = P->tile_tab[3];
  if (tt.n_tiles <= 0) return B200CG_OK;
  a.tiles = tt.d_tiles;
  a.cta_begin = tt.d_cta_begin;
  a.cta_clock = P->d_clock[3];
  P->clock_ctas[3] = tt.grid;
  kernel<<<tt.grid, (CW + 1) * 32, smem, s>>>(a);
  CU(cudaGetLastError());
  return B200CG_OK;
}
// Stage shapes (rows per stage x stages): even iterations stream r, p_old in 4-row stages x 3 (81 KB per CTA), odd ones
// also x in 4-row stages x 2 (81 KB). Measured against deeper and finer rings (4 x 4, 2 x 8 | 2 x 5, 3 x 3): within
// 1.5 %, these two on top (profiles/r2_single_sweep.md) - the ring is deep enough, so the alternatives are gone.
template <int FLAGS>
static int launch_fused_flags(b200cg_plan_s* P, const TileArgs& a, cudaStream_t s) {
  constexpr bool with_x = (FLAGS & (F_X2 | F_MAXN)) != 0;  // three (four with u) streams: two stages
  if (P->fused_cw == 14) {  // slabs of >= 4 M unknowns: one wide CTA per SM
    if constexpr (with_x) return launch_fused_cfg<FLAGS, 4, 2, 14>(P, a, s);
    else return launch_fused_cfg<FLAGS, 4, 4, 14>(P, a, s);
  }
  if constexpr (with_x) return launch_fused_cfg<FLAGS, 4, 2, FUSED_CW>(P, a, s);
  else return launch_fused_cfg<FLAGS, 4, 3, FUSED_CW>(P, a, s);
}
template <int FLAGS>
static int launch_fused(b200cg_plan_s* P, const TileArgs& a, cudaStream_t s) {
  if (a.defer == 2) return launch_fused_flags<FLAGS | F_SHARD>(P, a, s);  // sharded plan, peer memory
  return launch_fused_flags<FLAGS>(P, a, s);
}

static TileArgs base_args(b200cg_plan_s* P) {
  TileArgs a;
  memset(&a, 0, sizeof(a));
  a.st = P->d_state;
  a.partials = P->d_partials;
  a.cb_log = P->d_log;
  a.defer = P->desc.world > 1 ? 1 : 0;  // build_graph switches the loop kernels to 2 (peer memory) when it can
  a.g = P->g;
  return a;
}

// sharded plans: all-reduce this rank's totals, then every rank forms the same scalars (finalize_kernel)
static int reduce_and_finalize(b200cg_plan_s* P, int which, int flags, bool with_max, cudaStream_t s) {
  if (P->desc.world <= 1) return B200CG_OK;
  std::string err;
  if (!comm_allreduce_state(&P->comm, P->d_state, with_max, s, &err)) return fail(B200CG_ERR_COMM, "%s", err.c_str());
  finalize_kernel<<<1, 32, 0, s>>>(P->d_state, P->d_log, which, flags);
  CU(cudaGetLastError());
  return B200CG_OK;
}

}  // namespace b200cg
