// libb200cg, host side 3/3: the opt-in multigrid-preconditioned CG (b200cg_params.preconditioner = 1) - level hierarchy,
// V-cycle and the PCG loop around it. Kernels: mg_kernels.cuh. (The test suite checks it against a numpy restatement.)
#include "mg.h"

#include "mg_kernels.cuh"

using namespace b200cg;

namespace b200cg {

struct MgLevel {
  MgGeom g;
  double* x = nullptr;  // V: the level's iterate / correction (level 0: the caller's z)
  double* b = nullptr;  // right-hand side (level 0: the caller's r)
  double* t = nullptr;  // U: ping-pong partner of x, also holds the residual before it is restricted
  size_t elems = 0;
  int grid = 0;
};

struct MgHierarchy {
  std::vector<MgLevel> levels;
};

}  // namespace b200cg

static MgGeom level_geom(const b200cg_plan_s* P, int n, int m, bool lshape) {
  const b200cg_plan_desc& d = P->desc;
  MgGeom g;
  memset(&g, 0, sizeof(g));
  g.n = n;
  g.m = m;
  g.xsplit = lshape ? n / 2 : 0;
  g.ysplit = lshape ? m / 2 : 0;
  g.pitch = ((n + 1 + XOFF) + 15) / 16 * 16;
  const double hx = (d.b - d.a) / n, hy = (d.d - d.c) / m;  // grid_system.cpp:314-318 with this level's step
  g.A = -2 * (1 / (hx * hx) + 1 / (hy * hy));
  g.xk = 1 / (hx * hx);
  g.yk = 1 / (hy * hy);
  g.w = MG_OMEGA / g.A;
  g.nbx = (g.pitch / 2 + MG_THREADS - 1) / MG_THREADS;
  const int cap = P->sms * 16;  // <= the reduction's partial slots (sms * 16 + 64)
  g.nby = std::max(1, std::min(m - 1, cap / g.nbx));
  return g;
}

static bool can_coarsen(const MgGeom& g) {
  bool ok = g.n % 2 == 0 && g.m % 2 == 0 && g.n >= 8 && g.m >= 8;
  if (g.ysplit) ok = ok && g.xsplit % 2 == 0 && g.ysplit % 2 == 0;
  return ok;
}

void b200cg::mg_free(b200cg_plan_s* P) {
  if (!P->mg) return;
  for (size_t l = 1; l < P->mg->levels.size(); ++l) {
    cudaFree(P->mg->levels[l].x);
    cudaFree(P->mg->levels[l].b);
    cudaFree(P->mg->levels[l].t);
  }
  delete P->mg;
  P->mg = nullptr;
}

// Level 0 lives in the plan's own vectors (rhs = the CG residual, iterate = z, partner = a scratch vector); the coarse
// levels get three zero-initialised pitched vectors each (together a third of the fine level's).
static int mg_setup(b200cg_plan_s* P) {
  if (P->mg) return B200CG_OK;
  if (P->generic || P->desc.world > 1) return fail(B200CG_ERR_UNSUPPORTED, "the multigrid preconditioner needs a geometric single-GPU plan");
  const bool lshape = P->g.ysplit != 0;
  if (lshape && (P->g.xsplit != P->g.n / 2 || P->g.ysplit != P->g.m / 2))
    return fail(B200CG_ERR_UNSUPPORTED, "unexpected L-shape split");
  RET(ensure_scratch(P));
  MgHierarchy* H = new MgHierarchy();
  MgLevel l0;
  l0.g = level_geom(P, P->g.n, P->g.m, lshape);
  if (l0.g.pitch != P->g.pitch || P->g.ybase != 0) {
    delete H;
    return fail(B200CG_ERR_STATE, "multigrid level 0 does not match the plan's layout");
  }
  l0.elems = (size_t)(P->g.m + 1) * l0.g.pitch;
  l0.grid = l0.g.nbx * l0.g.nby;
  H->levels.push_back(l0);
  P->mg = H;
  while (can_coarsen(H->levels.back().g)) {
    const MgGeom& f = H->levels.back().g;
    MgLevel c;
    c.g = level_geom(P, f.n / 2, f.m / 2, lshape);
    c.elems = (size_t)(c.g.m + 1) * c.g.pitch;
    c.grid = c.g.nbx * c.g.nby;
    double** vs[3] = {&c.x, &c.b, &c.t};
    for (double** v : vs) {
      CU(cudaMalloc(v, c.elems * sizeof(double)));
      CU(cudaMemsetAsync(*v, 0, c.elems * sizeof(double), P->stream));
    }
    H->levels.push_back(c);
  }
  return B200CG_OK;
}

template <int OP>
static void launch_stencil(b200cg_plan_s* P, const MgLevel& L, const double* x, const double* b, double* out) {
  mg_stencil_kernel<OP><<<L.grid, MG_THREADS, 0, P->stream>>>(L.g, x, b, out, P->d_state, P->d_partials);
}

// z = M r: one V-cycle from level `l` down and back. Buffers per level: b (rhs), x ("V"), t ("U").
//   first sweep b -> U, second U -> V; residual (V, b) -> U; restrict U -> coarse b; recursion; V += P e; post V -> U -> V.
static int vcycle(b200cg_plan_s* P, int l, int64_t* launches) {
  MgHierarchy* H = P->mg;
  MgLevel& L = H->levels[l];
  static_assert(MG_NU_PRE == 2 && MG_NU_POST == 2 && MG_NU_COARSEST % 2 == 0, "buffer roles below assume these counts");
  launch_stencil<MG_JACOBI_FIRST>(P, L, nullptr, L.b, L.t);
  launch_stencil<MG_JACOBI>(P, L, L.t, L.b, L.x);
  *launches += 2;
  if (l + 1 == (int)H->levels.size()) {
    for (int k = 2; k < MG_NU_COARSEST; k += 2) {
      launch_stencil<MG_JACOBI>(P, L, L.x, L.b, L.t);
      launch_stencil<MG_JACOBI>(P, L, L.t, L.b, L.x);
      *launches += 2;
    }
    CU(cudaGetLastError());
    return B200CG_OK;
  }
  MgLevel& C = H->levels[l + 1];
  launch_stencil<MG_RESIDUAL>(P, L, L.x, L.b, L.t);
  mg_restrict_kernel<<<C.grid, MG_THREADS, 0, P->stream>>>(L.g, C.g, L.t, C.b);
  *launches += 2;
  RET(vcycle(P, l + 1, launches));
  mg_prolong_add_kernel<<<L.grid, MG_THREADS, 0, P->stream>>>(L.g, C.g, C.x, L.x);
  launch_stencil<MG_JACOBI>(P, L, L.x, L.b, L.t);
  launch_stencil<MG_JACOBI>(P, L, L.t, L.b, L.x);
  *launches += 3;
  CU(cudaGetLastError());
  return B200CG_OK;
}

int b200cg::mg_levels(const b200cg_plan_s* P) { return P->mg ? (int)P->mg->levels.size() : 0; }
int b200cg::mg_prepare(b200cg_plan_s* P) { return mg_setup(P); }

// The PCG loop. On entry the init kernel has run (r[0] = b, x = 0, p[0] = 0, ||r0|| in the device state and its verdict for
// max_it = 0 / b = 0); on exit the device state mirror holds the result like after the plain loop.
//   z = M r; p = z; loop { Ap = A p; alpha = r.z / p.Ap; x += alpha p; r -= alpha Ap; stop?; z = M r; beta; p = z + beta p }
int b200cg::mg_pcg_solve(b200cg_plan_s* P, const volatile int* stop_flag, int64_t* launches, bool* interrupted) {
  RET(mg_setup(P));
  cudaStream_t s = P->stream;
  MgHierarchy* H = P->mg;
  double *r = P->r[0], *p = P->p[0], *Ap = P->r[1], *z = P->p[1], *x = P->x;
  H->levels[0].b = r;
  H->levels[0].x = z;
  H->levels[0].t = P->va;
  const MgLevel& L0 = H->levels[0];
  const Geom& g = P->g;
  const size_t begin = (size_t)(g.ylo - g.ybase) * g.pitch, count = (size_t)(g.yhi - g.ylo) * g.pitch;
  const int ew = ew_grid(P, (long long)(count / 2));
  *P->h_stop = 0;
  auto read_state = [&]() -> int {
    CU(cudaMemcpyAsync(P->h_state, P->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return B200CG_OK;
  };
  RET(read_state());
  if (P->h_state->done) return B200CG_OK;
  RET(vcycle(P, 0, launches));
  pcg_dot_rz_kernel<<<ew, CTA_THREADS, 0, s>>>(r, z, begin, count, P->d_state, P->d_partials, 1);
  pcg_direction_kernel<<<ew, CTA_THREADS, 0, s>>>(p, z, begin, count, P->d_state);  // p = z + 0 * 0
  *launches += 2;
  for (;;) {
    if (stop_flag && *stop_flag) *P->h_stop = 1;
    launch_stencil<MG_APPLY_DOT>(P, L0, p, nullptr, Ap);
    pcg_update_kernel<<<ew, CTA_THREADS, 0, s>>>(x, r, p, Ap, begin, count, P->d_state, P->d_partials, P->d_stop);
    *launches += 2;
    CU(cudaGetLastError());
    RET(read_state());
    if (P->h_state->done) break;
    if (stop_flag && *stop_flag) {
      *interrupted = true;
      break;
    }
    RET(vcycle(P, 0, launches));
    pcg_dot_rz_kernel<<<ew, CTA_THREADS, 0, s>>>(r, z, begin, count, P->d_state, P->d_partials, 0);
    pcg_direction_kernel<<<ew, CTA_THREADS, 0, s>>>(p, z, begin, count, P->d_state);
    *launches += 2;
  }
  if (P->h_state->stop_reason == B200CG_STOP_INTERRUPTED) *interrupted = true;
  return B200CG_OK;
}
