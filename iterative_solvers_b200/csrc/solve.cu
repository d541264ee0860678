// libb200cg, host side 2/2: the operator and solve entry points - CUDA-graph captured CG loop with device-resident
// scalars, the single-cluster small-grid path, feedback balancing between launches, post-processing.
#include "sweep_launch.cuh"
#include "mg.h"

using namespace b200cg;

extern "C" int b200cg_apply(b200cg_plan_t P, const double* x_host, double* y_host) {
  if (!P || !x_host || !y_host) return fail(B200CG_ERR_INVALID_ARG, "plan/x_host/y_host is NULL");
  NEED_GEOMETRY(P);
  CU(cudaSetDevice(P->desc.device));
  RET(ensure_scratch(P));
  RET(upload_vector(P, x_host, P->va));
  RET(exchange_halo(P, P->va));
  TileArgs a = base_args(P);
  a.p_in = P->va;
  a.out = P->vb;
  RET((launch_tile<MODE_APPLY, 0>(P, a, P->stream)));
  RET(download_vector(P, P->vb, y_host));
  CU(cudaStreamSynchronize(P->stream));
  return B200CG_OK;
}

// ------------------------------------------------------------------------------------------- solve
// Captures `iters` CG iterations (even, so the ping-pong buffers return to their start) into one executable graph.
// timed: event-record nodes bracket the kernels of the first two iterations (the variant a solve launches first; the
// launches after it are pipelined and must not re-record events the host is still reading).
// Event nodes bracket iterations 2 and 3 of a timed graph (0 and 1 of a very short one): the first iteration of a
// launch still carries the launch skew between the ranks of a sharded plan.
static int timed_first_iteration(int iters) { return iters >= 4 ? 2 : 0; }

static int build_graph(b200cg_plan_s* P, int variant, int iters, bool timed, GraphEntry* out) {
  cudaStream_t s = P->stream;
  const bool with_u = variant & V_U, report = variant & V_REPORT, csr = variant & V_CSR, xdefer = variant & V_XDEFER;
  const bool fused = variant & V_FUSED, maxn = variant & V_MAXN;
  int kernels = 0;
  const int k0 = timed_first_iteration(iters);  // the two iterations whose kernels are bracketed by event nodes
  auto EV = [&](cudaEvent_t e, cudaStream_t st, unsigned int flags) {
    if (timed) cudaEventRecordWithFlags(e, st, flags);
  };
  CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = B200CG_OK;
  for (int k = 0; k < iters && rc == B200CG_OK; ++k) {
    const int par = k & 1;
    if (k == k0) EV(P->ev[0], s, cudaEventRecordExternal);
    if (csr) {
      // assembled path: p update + SpMV + dots, then the shared update pass
      CsrArgs ca = csr_args(&P->csr, P->d_state, P->d_partials, P->d_log, par);
      ca.stop_flag = P->d_stop;
      csr_spmv_kernel<1><<<csr_grid(P->csr.nrows, P->sms), CTA_THREADS, 0, s>>>(ca);
      ++kernels;
      if (k == k0) EV(P->ev[1], s, cudaEventRecordExternal);
      if (with_u)
        csr_update_kernel<1><<<csr_grid(P->csr.nrows, P->sms), CTA_THREADS, 0, s>>>(ca);
      else
        csr_update_kernel<0><<<csr_grid(P->csr.nrows, P->sms), CTA_THREADS, 0, s>>>(ca);
      ++kernels;
      if (k == k0) EV(P->ev[2], s, cudaEventRecordExternal);
      continue;
    }
    TileArgs a = base_args(P);
    a.stop_flag = P->d_stop;
    a.r_in = P->r[par];
    a.p_in = P->p[par];
    a.x = P->x;
    a.r_out = P->r[par ^ 1];
    a.p_out = P->p[par ^ 1];
    a.u = P->u;
    const int fl = xdefer ? ((k & 1) ? F_X2 : F_NOX) : ((with_u ? F_U : 0) | (report ? F_REPORT : 0));
    // sharded plans: reductions and halo rows over NVLink peer memory (no NCCL call in the loop); the per-iteration
    // report variant keeps the NCCL exchange
    const bool peer = P->desc.world > 1 && P->peer_mode && !report;
    if (peer) {
      a.defer = 2;
      a.peers = P->d_links;
      a.peer_trace = P->d_peer_trace;
      const Geom& g = P->g;
      const int rank = P->desc.rank;
      if (rank > 0) {  // neighbour below: its top halo row is its last stored row
        const size_t rows_below = (size_t)(P->ycuts[rank] - P->ycuts[rank - 1]) + 2;
        a.nb_r_below = P->nb_below_r[par ^ 1] + (rows_below - 1) * g.pitch;
        a.nb_p_below = P->nb_below_p[par ^ 1] + (rows_below - 1) * g.pitch;
        a.nb_r_below2 = P->nb_below_r[par ^ 1] + (rows_below + 1) * g.pitch;  // its second extra row (its yhi+1)
        a.nb_p_below2 = P->nb_below_p[par ^ 1] + (rows_below + 1) * g.pitch;
      }
      if (rank + 1 < P->desc.world) {  // neighbour above: its bottom halo row is its stored row 0
        const size_t rows_above = (size_t)(P->ycuts[rank + 2] - P->ycuts[rank + 1]) + 2;
        a.nb_r_above = P->nb_above_r[par ^ 1];
        a.nb_p_above = P->nb_above_p[par ^ 1];
        a.nb_r_above2 = P->nb_above_r[par ^ 1] + rows_above * g.pitch;  // its first extra row (its ylo-2)
        a.nb_p_above2 = P->nb_above_p[par ^ 1] + rows_above * g.pitch;
      }
    }
    if (fused) {
      // single-sweep iteration: one kernel; under the REL_L2 rule x is touched on odd iterations only, under the max-norm
      // rules every iteration (the event slots of the absent dot phase collapse to zero length)
      if (k == k0) EV(P->ev[1], s, cudaEventRecordExternal);
      if (k == k0 + 1) EV(P->ev[8], s, cudaEventRecordExternal);
      if (maxn) rc = with_u ? launch_fused<F_MAXN | F_U>(P, a, s) : launch_fused<F_MAXN>(P, a, s);  // max-norm rules: x every iteration
      else rc = (k & 1) ? launch_fused<F_X2>(P, a, s) : launch_fused<F_NOX>(P, a, s);
      ++kernels;
      if (k == k0) EV(P->ev[2], s, cudaEventRecordExternal);
      if (k == k0 + 1) EV(P->ev[9], s, cudaEventRecordExternal);
      continue;
    }
    rc = launch_tile<MODE_DOT, 0>(P, a, s);
    ++kernels;
    if (rc == B200CG_OK && !peer && P->desc.world > 1) {  // (peer memory: the sweep's last CTA did it)
      rc = reduce_and_finalize(P, 1, fl, false, s);
      ++kernels;
    }
    if (k == k0) EV(P->ev[1], s, cudaEventRecordExternal);
    if (k == k0 + 1) EV(P->ev[8], s, cudaEventRecordExternal);
    if (rc != B200CG_OK) break;
    if (xdefer) rc = (k & 1) ? launch_tile<MODE_UPD, F_X2>(P, a, s) : launch_tile<MODE_UPD, F_NOX>(P, a, s);
    else if (report && with_u) rc = launch_tile<MODE_UPD, F_REPORT | F_U>(P, a, s);
    else if (report) rc = launch_tile<MODE_UPD, F_REPORT>(P, a, s);
    else if (with_u) rc = launch_tile<MODE_UPD, F_U>(P, a, s);
    else rc = launch_tile<MODE_UPD, 0>(P, a, s);
    ++kernels;
    if (rc == B200CG_OK && !peer && P->desc.world > 1) {
      rc = reduce_and_finalize(P, 2, fl, /*with_max=*/!xdefer, s);
      ++kernels;
      if (rc == B200CG_OK) rc = exchange_halo2(P, P->r[par ^ 1], P->p[par ^ 1]);
    }
    if (k == k0) EV(P->ev[2], s, cudaEventRecordExternal);
    if (k == k0 + 1 && !report) EV(P->ev[9], s, cudaEventRecordExternal);
    if (rc == B200CG_OK && report) {
      TileArgs ra = base_args(P);
      ra.p_in = P->x;
      ra.r_in = P->b;
      ra.u = P->u;
      if (P->desc.world > 1) rc = exchange_halo(P, P->x);
      if (rc == B200CG_OK)
        rc = with_u ? launch_tile<MODE_APPLY, F_REPORT | F_U>(P, ra, s)
                    : launch_tile<MODE_APPLY, F_REPORT>(P, ra, s);
      ++kernels;
      if (rc == B200CG_OK && P->desc.world > 1) {
        rc = reduce_and_finalize(P, 3, fl, false, s);
        ++kernels;
      }
    }
  }
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(s, &graph);
  if (rc != B200CG_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(B200CG_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(B200CG_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  out->exec = exec;
  out->iters = iters;
  out->kernels = kernels;
  return B200CG_OK;
}

// Small-grid path: how many CTAs a cluster needs to hold r, p, x of the grid in shared memory (0 = does not fit).
static int cluster_ctas_for(const b200cg_plan_s* P, int* rows_per_cta, size_t* smem) {
  if (P->generic || P->desc.world > 1) return 0;
  const Geom& g = P->g;
  // 8 CTAs (portable cluster size) for the smallest grids, where the per-iteration cost is the three cluster
  // barriers; 16 CTAs (non-portable size, B200 allows it) once a band would exceed 16 rows: the sweep over the band
  // is what takes the time there (n = 250: 11.6 us/iteration with 8 CTAs).
  const int order[2] = {(g.m - 1 > 128 && P->cluster16_ok) ? 16 : 8, (g.m - 1 > 128 && P->cluster16_ok) ? 8 : 16};
  for (int c : order) {
    if (c == 16 && !P->cluster16_ok) continue;
    const int rpc = (g.m - 1 + c - 1) / c;
    const size_t bytes = cluster_smem_bytes(rpc, g.pitch);
    if (bytes <= 200 * 1024) {
      *rows_per_cta = rpc;
      *smem = bytes;
      return c;
    }
  }
  return 0;
}

static int launch_cluster_solve(b200cg_plan_s* P, int ctas, int rows_per_cta, size_t smem, bool with_u, cudaStream_t s) {
  CU(cudaFuncSetAttribute(cg_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ctas > 8) CU(cudaFuncSetAttribute(cg_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  ClusterArgs ca;
  ca.b = P->b;
  ca.u = with_u ? P->u : nullptr;
  ca.x = P->x;
  ca.st = P->d_state;
  ca.cb_log = P->d_log;
  ca.stop_flag = P->d_stop;
  ca.g = P->g;
  ca.rows_per_cta = rows_per_cta;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas, 1, 1);
  cfg.blockDim = dim3(CL_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CU(cudaLaunchKernelEx(&cfg, cg_cluster_kernel, ca));
  return B200CG_OK;
}

static int default_iters_per_graph(const b200cg_plan_s* P) {
  // A launch should last ~50 ms: every graph boundary costs a host round trip (~0.1 ms on one GPU, ~0.4 ms as the
  // maximum over 8 ranks - profiles/r2_scaling.md) even though consecutive launches are pipelined, and nothing else
  // depends on the length any more (interrupts are polled on the device, callback records are logged there).
  const double t_it_us = 15.0 + (double)local_count(P) * 40.0 / 5.5e6;  // 40 B per unknown at 5.5 TB/s + fixed cost
  const int k = (int)(50e3 / t_it_us);
  return std::max(20, std::min(200, k & ~1));
}

// One solve of a b200cg_solve_batch queue: its right-hand side is already on its way into P->compact on the input copy
// stream, the next one follows as soon as the staging buffer is free, the solution leaves on the output copy stream.
struct BatchStep {
  BatchIo* io;
  const double* next_b;  // right-hand side of the following solve (nullptr: none)
  int index;
};

// Everything one b200cg_solve call carries between its phases.
struct SolveCall {
  b200cg_plan_s* P;
  const b200cg_params* prm;
  b200cg_info* info;
  b200cg_iter_cb cb;
  void* user;
  const volatile int* stop_flag;
  bool csr, with_u, report;
  bool xdefer = false, fused = false, fused_maxn = false, use_cluster = false, interrupted = false;
  unsigned int consumed = 0;  // callback records already delivered
  double dot_ms = 0.0, upd_even_ms = 0.0, upd_odd_ms = 0.0;
  int samples = 0;
  long long count;  // unknowns of this rank
  const BatchStep* batch = nullptr;
};

static int check_solve_args(b200cg_plan_t P, const b200cg_params* prm, const double* b_host, double* x_host,
                            b200cg_info* info, b200cg_iter_cb cb) {
  if (!P || !prm || !info) return fail(B200CG_ERR_INVALID_ARG, "plan/params/info is NULL");
  if (prm->op != B200CG_OP_MATRIX_FREE && prm->op != B200CG_OP_CSR) return fail(B200CG_ERR_INVALID_ARG, "unknown operator %d", prm->op);
  if (prm->rule != B200CG_RULE_REL_L2 && prm->rule != B200CG_RULE_MAXNORM) return fail(B200CG_ERR_INVALID_ARG, "unknown rule %d", prm->rule);
  if (!prm->rhs_on_device && !b_host) return fail(B200CG_ERR_INVALID_ARG, "b_host is NULL and rhs_on_device is 0");
  if (prm->rhs_on_device && !P->have_rhs) return fail(B200CG_ERR_STATE, "rhs_on_device set but the plan holds no rhs");
  if (!prm->keep_x_on_device && !x_host) return fail(B200CG_ERR_INVALID_ARG, "x_host is NULL and keep_x_on_device is 0");
  const bool csr = prm->op == B200CG_OP_CSR;
  if (!csr) NEED_GEOMETRY(P);
  if (prm->preconditioner != B200CG_PRECOND_NONE && prm->preconditioner != B200CG_PRECOND_MULTIGRID)
    return fail(B200CG_ERR_INVALID_ARG, "unknown preconditioner %d", prm->preconditioner);
  if (prm->preconditioner == B200CG_PRECOND_MULTIGRID && (csr || prm->rule != B200CG_RULE_REL_L2 || cb || P->desc.world > 1))
    return fail(B200CG_ERR_UNSUPPORTED, "the multigrid preconditioner serves the matrix-free operator under the relative-residual "
                                        "rule without a report callback on a single-GPU plan");
  if (csr && !P->csr.row_map) return fail(B200CG_ERR_STATE, "CSR solve without a matrix: call b200cg_set_csr / b200cg_assemble_csr");
  if (csr && prm->rule == B200CG_RULE_REL_L2 && cb) return fail(B200CG_ERR_UNSUPPORTED, "per-iteration report callbacks exist only on the matrix-free path");
  return B200CG_OK;
}

// Host vectors -> device (the assembled path keeps compact vectors, the matrix-free path pitched ones).
static int upload_inputs(SolveCall& c, const double* b_host, const double* u_host) {
  b200cg_plan_s* P = c.P;
  cudaStream_t s = P->stream;
  const int64_t bytes = c.count * (int64_t)sizeof(double);
  if (c.csr) {
    std::string err;
    int rc = csr_ensure_vectors(&P->csr, s, &err);
    if (rc) return fail(rc, "%s", err.c_str());
    if (!c.prm->rhs_on_device) {
      CU(cudaMemcpyAsync(P->csr.b, b_host, bytes, cudaMemcpyHostToDevice, s));
      c.info->h2d_bytes += bytes;
    } else {
      gather_compact_kernel<<<ew_grid(P, c.count), CTA_THREADS, 0, s>>>(P->b, P->csr.b, P->g);
      c.info->kernel_launches += 1;
    }
    if (c.with_u) {
      CU(cudaMemcpyAsync(P->csr.u, u_host, bytes, cudaMemcpyHostToDevice, s));
      c.info->h2d_bytes += bytes;
    }
    P->csr.has_u = c.with_u;
  } else {
    if (c.batch) {
      // the copy was enqueued on the input copy stream while the previous solve iterated
      BatchIo* io = c.batch->io;
      CU(cudaStreamWaitEvent(s, io->h2d_done, 0));
      scatter_compact_kernel<<<ew_grid(P, c.count), CTA_THREADS, 0, s>>>(P->compact, P->b, P->g);
      CU(cudaGetLastError());
      CU(cudaEventRecord(io->in_free, s));
      if (c.batch->next_b) {
        CU(cudaStreamWaitEvent(io->s_in, io->in_free, 0));
        CU(cudaMemcpyAsync(P->compact, c.batch->next_b, bytes, cudaMemcpyHostToDevice, io->s_in));
        CU(cudaEventRecord(io->h2d_done, io->s_in));
      }
      P->have_rhs = true;
      c.info->h2d_bytes += bytes;
      c.info->kernel_launches += 1;
    } else if (!c.prm->rhs_on_device) {
      RET(upload_vector(P, b_host, P->b));
      P->have_rhs = true;
      c.info->h2d_bytes += bytes;
      c.info->kernel_launches += 1;
    }
    if (c.with_u) {
      RET(ensure_u(P));
      RET(upload_vector(P, u_host, P->u));
      c.info->h2d_bytes += bytes;
      c.info->kernel_launches += 1;
    }
  }
  P->have_u = c.with_u;
  return B200CG_OK;
}

// Solver parameters into the device-side state; everything else in it is armed by the init kernel.
static int arm_device_state(SolveCall& c) {
  b200cg_plan_s* P = c.P;
  const b200cg_params* prm = c.prm;
  DevState hs;
  memset(&hs, 0, sizeof(hs));
  hs.eps_rel = prm->eps_rel;
  hs.eps_p = prm->eps_p;
  hs.eps_r = prm->eps_r;
  hs.eps_e = prm->eps_e;
  hs.max_it = prm->max_it;
  hs.rule = prm->rule;
  hs.has_u = c.with_u ? 1 : 0;
  hs.callback_every = (c.cb && prm->rule == B200CG_RULE_MAXNORM) ? (prm->callback_every > 0 ? prm->callback_every : 100) : 0;
  hs.epoch[0] = P->peer_epoch[0];  // the PeerSync flags are monotonic over the plan's life
  hs.epoch[1] = P->peer_epoch[1];
  *P->h_state = hs;
  CU(cudaMemcpyAsync(P->d_state, P->h_state, sizeof(DevState), cudaMemcpyHostToDevice, P->stream));
  return B200CG_OK;
}

static void deliver_callbacks(SolveCall& c, const DevState& st, const CbRecord* log) {
  if (c.cb) {
    // (the launch sizes keep a launch's records within the ring; should one ever lap it, deliver the surviving tail only)
    if (st.n_log - c.consumed > (unsigned int)CB_LOG_CAP) c.consumed = st.n_log - CB_LOG_CAP;
    for (; c.consumed < st.n_log; ++c.consumed) {
      const CbRecord& rec = log[c.consumed % CB_LOG_CAP];
      c.cb(c.user, (int)rec.it, rec.precision, rec.residual, rec.error);
    }
  } else {
    c.consumed = st.n_log;
  }
}

// Small-grid path: the whole solve is one launch of one thread-block cluster (cluster_kernel.cuh).
static int run_cluster_solve(SolveCall& c, int ctas, int rows_per_cta, size_t smem) {
  b200cg_plan_s* P = c.P;
  cudaStream_t s = P->stream;
  *P->h_stop = 0;
  CU(cudaEventRecord(P->ev[5], s));
  RET(launch_cluster_solve(P, ctas, rows_per_cta, smem, c.with_u, s));
  c.info->kernel_launches += 1;
  CU(cudaMemcpyAsync(P->h_state, P->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(P->h_log, P->d_log, sizeof(CbRecord) * CB_LOG_CAP, cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(P->ev[8], s));
  // the kernel polls the mapped flag every CL_POLL_EVERY iterations; forward the caller's stop request
  for (int spins = 0; cudaEventQuery(P->ev[8]) == cudaErrorNotReady; ++spins) {
    if (c.stop_flag && *c.stop_flag) *P->h_stop = 1;
    if (spins > 2000) std::this_thread::sleep_for(std::chrono::microseconds(50));  // long solve: stop burning a core
  }
  CU(cudaStreamSynchronize(s));
  c.interrupted = P->h_state->stop_reason == B200CG_STOP_INTERRUPTED;
  deliver_callbacks(c, *P->h_state, P->h_log);
  return B200CG_OK;
}

static int launch_init(SolveCall& c) {
  b200cg_plan_s* P = c.P;
  cudaStream_t s = P->stream;
  if (c.csr) {
    csr_init_kernel<<<csr_grid(c.count, P->sms), CTA_THREADS, 0, s>>>(csr_args(&P->csr, P->d_state, P->d_partials, P->d_log, 0));
    CU(cudaGetLastError());
    c.info->kernel_launches += 1;
    return B200CG_OK;
  }
  const Geom& g = P->g;
  InitArgs ia;
  ia.b = P->b;
  ia.u = c.with_u ? P->u : nullptr;
  ia.r = P->r[0];
  ia.p = P->p[0];
  ia.x = P->x;
  ia.st = P->d_state;
  ia.partials = P->d_partials;
  ia.cb_log = P->d_log;
  ia.begin = (size_t)(g.ylo - g.ybase) * g.pitch;
  ia.count = (size_t)(g.yhi - g.ylo) * g.pitch;
  ia.defer = P->desc.world > 1 ? 1 : 0;
  cg_init_kernel<<<ew_grid(P, (long long)(ia.count / 2)), CTA_THREADS, 0, s>>>(ia);
  CU(cudaGetLastError());
  c.info->kernel_launches += 1;
  if (P->desc.world > 1) {  // once per solve: NCCL
    RET(reduce_and_finalize(P, 0, c.with_u ? F_U : 0, true, s));
    RET(exchange_halo2(P, P->r[0], P->p[0]));
    c.info->kernel_launches += 1;
  }
  if (c.fused) {
    // single-sweep iteration: alpha_0 = r0.r0 / r0.A r0 needs one operator application up front; the dot sweep with
    // beta = 0 and p_old = 0 is exactly that (p = r0)
    TileArgs a = base_args(P);
    a.r_in = P->r[0];
    a.p_in = P->p[0];
    RET((launch_tile<MODE_DOT, 0>(P, a, s)));
    c.info->kernel_launches += 1;
    if (P->desc.world > 1) {
      // sharded plan: the dot sweep left this rank's sums in loc_s (defer = 1); then the second halo rows of the first
      // sweep's inputs: r0 from the neighbours (their second first / last rows), p = 0
      RET(reduce_and_finalize(P, 1, 0, false, s));
      c.info->kernel_launches += 1;
      const Geom& g = P->g;
      double* r0 = P->r[0];
      double* extra_r = r0 + (size_t)g.yrows * g.pitch;
      std::string err;
      if (!comm_halo(&P->comm, r0 + (size_t)2 * g.pitch, r0 + (size_t)(g.yrows - 3) * g.pitch, extra_r, extra_r + g.pitch,
                     g.pitch, s, &err))
        return fail(B200CG_ERR_COMM, "%s", err.c_str());
      CU(cudaMemsetAsync(P->p[0] + (size_t)g.yrows * g.pitch, 0, (size_t)2 * g.pitch * sizeof(double), s));
    }
  }
  return B200CG_OK;
}

// Event nodes bracket the kernels of the first two captured iterations of a graph launch; a sample counts only if
// those iterations really ran in this launch (`advanced` = iterations completed by it).
static void sample_kernel_times(SolveCall& c, int advanced, int K) {
  advanced -= timed_first_iteration(K);  // iterations completed from the first bracketed one on
  b200cg_plan_s* P = c.P;
  float d0 = 0.f, u0 = 0.f, d1 = 0.f, u1 = 0.f;
  if (!c.csr && !c.report && advanced >= 2) {
    if (cudaEventElapsedTime(&d0, P->ev[0], P->ev[1]) == cudaSuccess && cudaEventElapsedTime(&u0, P->ev[1], P->ev[2]) == cudaSuccess &&
        cudaEventElapsedTime(&d1, P->ev[2], P->ev[8]) == cudaSuccess && cudaEventElapsedTime(&u1, P->ev[8], P->ev[9]) == cudaSuccess) {
      c.dot_ms += 0.5 * (d0 + d1);
      c.upd_even_ms += u0;
      c.upd_odd_ms += u1;
      ++c.samples;
    }
  } else if ((c.csr || c.report) && advanced >= 1) {
    if (cudaEventElapsedTime(&d0, P->ev[0], P->ev[1]) == cudaSuccess && cudaEventElapsedTime(&u0, P->ev[1], P->ev[2]) == cudaSuccess) {
      c.dot_ms += d0;
      c.upd_even_ms += u0;
      c.upd_odd_ms += u0;
      ++c.samples;
    }
  }
}

// The general path: init kernel, then graph launches of K captured iterations until the device says done.
static int run_graph_solve(SolveCall& c) {
  b200cg_plan_s* P = c.P;
  const b200cg_params* prm = c.prm;
  cudaStream_t s = P->stream;
  // single-sweep iteration (the default): the relative-residual rule without report and MSGSolver's max-norm rules;
  // sharded plans need the peer-memory exchange and at least 4 rows per rank (every rank sees all cuts, so all ranks
  // decide alike)
  const bool want_fused = prm->single_sweep == 1 || (prm->single_sweep == 0 && P->single_sweep_default);
  bool fused_ok = P->desc.world <= 1;
  if (!fused_ok && P->peer_mode) {
    fused_ok = true;
    for (int r = 0; r < P->desc.world; ++r) fused_ok = fused_ok && (P->ycuts[r + 1] - P->ycuts[r] >= 4);
  }
  c.fused = want_fused && fused_ok && !c.csr && !c.report;
  c.fused_maxn = c.fused && prm->rule == B200CG_RULE_MAXNORM;
  RET(launch_init(c));
  int K = prm->iters_per_graph > 0 ? prm->iters_per_graph : default_iters_per_graph(P);
  if (prm->max_it > 0) K = std::min(K, prm->max_it + 1);
  K = std::max(2, (K + 1) & ~1);
  // callback records of one launch must fit the ring: report = one per iteration, MAXNORM with callback_every = 1 one per
  // iteration plus the init record
  K = std::min(K, c.report ? CB_LOG_CAP / 2 : (c.cb ? CB_LOG_CAP - 2 : CB_LOG_CAP));
  // x-deferral: the relative-residual rule never looks at x, so x is only touched every other iteration
  c.xdefer = (c.fused && !c.fused_maxn) || (P->x_deferral && !c.csr && !c.report && prm->rule == B200CG_RULE_REL_L2);
  const int variant = c.fused_maxn ? (V_FUSED | V_MAXN | (c.with_u ? V_U : 0))
                      : c.fused    ? V_FUSED
                                   : (c.xdefer ? V_XDEFER : ((c.with_u ? V_U : 0) | (c.report ? V_REPORT : 0) | (c.csr ? V_CSR : 0)));
  // two cached graphs per (variant, K): with event nodes (first launch of a solve: kernel times) and without
  GraphEntry& g_timed = P->graphs[(variant | V_TIMED) * 4096 + K];
  GraphEntry& g_plain = P->graphs[variant * 4096 + K];
  if (!g_timed.exec) RET(build_graph(P, variant, K, true, &g_timed));
  if (!g_plain.exec) RET(build_graph(P, variant, K, false, &g_plain));

  CU(cudaEventRecord(P->ev[5], s));
  *P->h_stop = 0;
  // Interrupts (requestStop): the caller's flag is forwarded into the mapped flag the loop kernels poll every
  // STOP_POLL_EVERY-th iteration; the device then ends the solve with INTERRUPTED - on a sharded peer-memory plan on
  // every rank at the same iteration (the request travels with the reduction slots). Only the NCCL exchange (fallback,
  // per-iteration report) lacks that channel: there the ranks agree between graph launches.
  const bool device_stop = P->desc.world <= 1 || (P->peer_mode && !c.report);
  // Launches are pipelined: launch k+1 is enqueued before the host waits for launch k's read-back (two pinned mirrors),
  // so the GPU never idles across a graph boundary and the ranks of a sharded plan do not pick up each other's host
  // jitter. A launch enqueued after the stop rule fired drains harmlessly (kernels begin with `if (st->done) return`).
  // Not pipelined: the solve's first launch (its event nodes are read), launches followed by a feedback-balancing
  // step (needs an idle stream), and the NCCL exchange (the ranks agree on stop requests between launches).
  int last_slot = 0;  // read-back target of the most recently enqueued launch
  auto enqueue = [&](int slot, GraphEntry& ge) -> int {
    last_slot = slot;
    if (c.stop_flag && *c.stop_flag) *P->h_stop = 1;
    CU(cudaGraphLaunch(ge.exec, s));
    c.info->kernel_launches += ge.kernels;
    CU(cudaMemcpyAsync(P->h_state_m[slot], P->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(P->h_log_m[slot], P->d_log, sizeof(CbRecord) * CB_LOG_CAP, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(P->ev_launch[slot], s));
    return B200CG_OK;
  };
  auto wait_for = [&](int slot) -> int {
    if (c.stop_flag) {
      for (int spins = 0; cudaEventQuery(P->ev_launch[slot]) == cudaErrorNotReady; ++spins) {
        if (*c.stop_flag) *P->h_stop = 1;
        if (spins > 20000) std::this_thread::sleep_for(std::chrono::microseconds(20));
      }
    }
    CU(cudaEventSynchronize(P->ev_launch[slot]));
    return B200CG_OK;
  };
  int& rounds = c.fused ? P->balance_rounds_fused : P->balance_rounds;
  int it_before = 0, cur = 0;
  bool next_enqueued = false;
  // the init kernel's verdict (0 iterations) and its callback record come back with the first launch
  RET(enqueue(cur, g_timed));
  for (bool first = true;; first = false) {
    const bool balancing = rounds > 0 && !c.csr;
    const bool pipelined = !first && !balancing && device_stop;
    if (pipelined && !next_enqueued) {
      RET(enqueue(cur ^ 1, g_plain));
      next_enqueued = true;
    }
    RET(wait_for(cur));
    const DevState& st = *P->h_state_m[cur];
    const int advanced = st.it - it_before;
    it_before = st.it;
    if (first) sample_kernel_times(c, advanced, K);
    deliver_callbacks(c, st, P->h_log_m[cur]);
    if (st.done) break;
    if (balancing && advanced >= 2) {
      // young plan: correct the static split from the measured per-CTA sweep times (the stream is idle here); only
      // the flavours this loop launches carry fresh stamps
      if (c.fused) {
        RET(rebalance_tiles(P, 3));
      } else {
        for (int fl = 0; fl < 3; ++fl) RET(rebalance_tiles(P, fl));
      }
      --rounds;
    }
    bool stop = c.stop_flag && *c.stop_flag;
    if (stop) *P->h_stop = 1;
    if (!device_stop) {  // NCCL exchange: every rank learns whether any rank wants to stop (all ranks call this)
      bool none = true;
      std::string err;
      if (!comm_all_agree(&P->comm, !stop, &none, s, &err)) return fail(B200CG_ERR_COMM, "%s", err.c_str());
      stop = !none;
    } else if (P->desc.world > 1) {
      stop = false;  // the device verdict is the collective one
    }
    if (stop) {
      c.interrupted = true;
      break;
    }
    if (!next_enqueued) RET(enqueue(cur ^ 1, g_plain));
    next_enqueued = false;
    cur ^= 1;
  }
  // A launch enqueued ahead drains: after the stop rule fired it does nothing; after a host-side interrupt it may have
  // advanced a few more iterations (until the device saw the flag) - either way its read-back is the final state.
  CU(cudaStreamSynchronize(s));
  deliver_callbacks(c, *P->h_state_m[last_slot], P->h_log_m[last_slot]);
  *P->h_state = *P->h_state_m[last_slot];
  if (P->h_state->stop_reason == B200CG_STOP_INTERRUPTED) c.interrupted = true;
  return B200CG_OK;
}

// x of the last iterate to the host (settling a pending x-deferral update first).
static int collect_solution(SolveCall& c, const DevState& st, double* x_host) {
  b200cg_plan_s* P = c.P;
  cudaStream_t s = P->stream;
  if (c.xdefer && st.x_pending) {  // the loop ended on an even iteration: x += alpha * p is still owed
    const Geom& g = P->g;
    const size_t begin = (size_t)(g.ylo - g.ybase) * g.pitch, count = (size_t)(g.yhi - g.ylo) * g.pitch;
    x_flush_kernel<<<ew_grid(P, (long long)(count / 2)), CTA_THREADS, 0, s>>>(P->x, P->p[st.it & 1], P->d_state, begin, count);
    CU(cudaGetLastError());
    c.info->kernel_launches += 1;
  }
  P->solution_in_csr = c.csr;
  if (c.batch) {
    // gather into the second staging buffer once the previous solution has left it; the copy to the host runs on the
    // output copy stream under the next solve's iterations
    BatchIo* io = c.batch->io;
    const int i = c.batch->index;
    if (i > 0) CU(cudaStreamWaitEvent(s, io->d2h_done[(i - 1) & 1], 0));
    gather_compact_kernel<<<ew_grid(P, c.count), CTA_THREADS, 0, s>>>(P->x, io->stage_out, P->g);
    CU(cudaGetLastError());
    CU(cudaEventRecord(io->gathered, s));
    CU(cudaStreamWaitEvent(io->s_out, io->gathered, 0));
    CU(cudaMemcpyAsync(x_host, io->stage_out, c.count * sizeof(double), cudaMemcpyDeviceToHost, io->s_out));
    CU(cudaEventRecord(io->d2h_done[i & 1], io->s_out));
    c.info->kernel_launches += 1;
    c.info->d2h_bytes += c.count * (int64_t)sizeof(double);
  } else if (!c.prm->keep_x_on_device) {
    if (c.csr) {
      CU(cudaMemcpyAsync(x_host, P->csr.x, c.count * sizeof(double), cudaMemcpyDeviceToHost, s));
    } else {
      RET(download_vector(P, P->x, x_host));
      c.info->kernel_launches += 1;
    }
    c.info->d2h_bytes += c.count * (int64_t)sizeof(double);
  }
  return B200CG_OK;
}

static void fill_info(const SolveCall& c, const DevState& st) {
  b200cg_plan_s* P = c.P;
  b200cg_info* info = c.info;
  info->iterations = st.it;
  info->converged = c.interrupted ? 0 : st.converged;
  info->stop_reason = c.interrupted ? B200CG_STOP_INTERRUPTED : st.stop_reason;
  info->r0_l2 = st.r0_norm;
  info->r_l2 = st.r_norm;
  info->r_max = st.r_max;
  info->dx_max = st.dx_max;
  // the x-deferral flavours (REL_L2 without callback, one or two sweeps) do not track |x - u|_inf: nothing reads it there
  info->err_max = c.xdefer ? DBL_MAX : st.err_max;
  auto span = [&](int a, int b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, P->ev[a], P->ev[b]);
    return (double)ms;
  };
  info->h2d_ms = span(3, 4);
  info->solve_ms = span(5, 6);
  info->d2h_ms = span(6, 7);
  info->device_ms = span(3, 7);
  info->dot_kernel_ms = c.samples ? c.dot_ms / c.samples : 0.0;
  info->upd_kernel_ms = c.samples ? 0.5 * (c.upd_even_ms + c.upd_odd_ms) / c.samples : 0.0;
  info->upd_even_ms = c.samples ? c.upd_even_ms / c.samples : 0.0;
  info->upd_odd_ms = c.samples ? c.upd_odd_ms / c.samples : 0.0;
  info->kernel_samples = c.samples;
  info->x_deferral = c.xdefer ? 1 : 0;
  info->single_sweep = c.fused ? 1 : 0;
  info->cluster_path = c.use_cluster ? 1 : 0;
  info->peer_exchange = (P->desc.world > 1 && P->peer_mode && !c.report && !c.use_cluster) ? 1 : 0;
}

static int solve_impl(b200cg_plan_t P, const b200cg_params* prm, const double* b_host, const double* u_host,
                      double* x_host, b200cg_info* info, b200cg_iter_cb cb, void* user, const volatile int* stop_flag,
                      const BatchStep* batch) {
  RET(check_solve_args(P, prm, b_host, x_host, info, cb));
  memset(info, 0, sizeof(*info));
  const double t_begin = now_ms();
  CU(cudaSetDevice(P->desc.device));
  cudaStream_t s = P->stream;
  SolveCall c{P, prm, info, cb, user, stop_flag,
              /*csr=*/prm->op == B200CG_OP_CSR, /*with_u=*/u_host != nullptr,
              /*report=*/prm->rule == B200CG_RULE_REL_L2 && cb != nullptr};
  c.count = local_count(P);
  c.batch = batch;
  info->local_unknowns = c.count;
  P->have_solution = false;

  CU(cudaEventRecord(P->ev[3], s));
  RET(upload_inputs(c, b_host, u_host));
  CU(cudaEventRecord(P->ev[4], s));
  RET(arm_device_state(c));

  // grids that fit one thread-block cluster's shared memory run as a single resident kernel
  int cl_rows = 0, cl_ctas = 0;
  size_t cl_smem = 0;
  {
    // the single launch appends every record of the solve to the ring: it 0, it 1, every callback_every-th (the value
    // arm_device_state hands the device), so a dense cadence or a long solve stays on the graph path
    const int every = prm->callback_every > 0 ? prm->callback_every : 100;
    const long long cb_records = 3 + (long long)std::max(prm->max_it, 0) / every;
    const bool eligible = !c.csr && !c.report && prm->small_grid_path != 1 && P->cluster_enabled && !(cb && cb_records > CB_LOG_CAP);
    if (eligible) cl_ctas = cluster_ctas_for(P, &cl_rows, &cl_smem);
    if (prm->small_grid_path == 2 && cl_ctas == 0)
      return fail(B200CG_ERR_UNSUPPORTED, "small_grid_path = 2 but this solve cannot run in one cluster "
                                          "(grid too large, sharded plan, CSR operator or per-iteration report)");
  }
  const bool mg = prm->preconditioner == B200CG_PRECOND_MULTIGRID;
  c.use_cluster = cl_ctas > 0 && !mg;
  if (mg) {
    // opt-in: CG preconditioned by a multigrid V-cycle (mg.cu) - its own loop, same init kernel and stop rule
    RET(mg_prepare(P));
    RET(launch_init(c));
    CU(cudaEventRecord(P->ev[5], s));
    RET(mg_pcg_solve(P, stop_flag, &info->kernel_launches, &c.interrupted));
    info->preconditioner = B200CG_PRECOND_MULTIGRID;
    info->mg_levels = mg_levels(P);
  } else if (c.use_cluster) {
    RET(run_cluster_solve(c, cl_ctas, cl_rows, cl_smem));
  } else {
    RET(run_graph_solve(c));
  }
  CU(cudaEventRecord(P->ev[6], s));

  const DevState st = *P->h_state;
  P->peer_epoch[0] = st.epoch[0];
  P->peer_epoch[1] = st.epoch[1];
  if (st.comm_error) return fail(B200CG_ERR_COMM, "peer-memory exchange timed out after %d iterations (a rank stopped publishing)", st.it);
  RET(collect_solution(c, st, x_host));
  CU(cudaEventRecord(P->ev[7], s));
  CU(cudaStreamSynchronize(s));
  P->have_solution = true;
  fill_info(c, st);
  // MSGSolver fires one more callback after the loop with the final values (msg_solver.cpp:193-195)
  if (cb && prm->rule == B200CG_RULE_MAXNORM) cb(user, st.it, st.dx_max, st.r_max, st.err_max);
  info->total_ms = now_ms() - t_begin;
  return B200CG_OK;
}

extern "C" int b200cg_solve(b200cg_plan_t P, const b200cg_params* prm, const double* b_host, const double* u_host,
                            double* x_host, b200cg_info* info, b200cg_iter_cb cb, void* user,
                            const volatile int* stop_flag) {
  return solve_impl(P, prm, b_host, u_host, x_host, info, cb, user, stop_flag, nullptr);
}

// ------------------------------------------------------------------------------------------- batch of right-hand sides
static int batch_io(b200cg_plan_s* P, BatchIo** out) {
  if (!P->batch) P->batch = new BatchIo();
  BatchIo* io = P->batch;
  if (!io->s_in) CU(cudaStreamCreateWithFlags(&io->s_in, cudaStreamNonBlocking));
  if (!io->s_out) CU(cudaStreamCreateWithFlags(&io->s_out, cudaStreamNonBlocking));
  if (!io->stage_out) CU(cudaMalloc(&io->stage_out, std::max<long long>(local_count(P), 1) * sizeof(double)));
  for (cudaEvent_t* e : {&io->h2d_done, &io->in_free, &io->gathered, &io->d2h_done[0], &io->d2h_done[1]})
    if (!*e) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  *out = io;
  return B200CG_OK;
}

extern "C" int b200cg_solve_batch(b200cg_plan_t P, const b200cg_params* prm, int count, const double* const* b_hosts,
                                  double* const* x_hosts, b200cg_info* infos, b200cg_batch_cb done, void* user,
                                  const volatile int* stop_flag) {
  if (!P || !prm || count < 0 || (count > 0 && (!b_hosts || !x_hosts || !infos)))
    return fail(B200CG_ERR_INVALID_ARG, "plan/params/b_hosts/x_hosts/infos is NULL or count < 0");
  if (prm->op != B200CG_OP_MATRIX_FREE) return fail(B200CG_ERR_UNSUPPORTED, "b200cg_solve_batch serves the matrix-free operator");
  if (prm->rhs_on_device || prm->keep_x_on_device)
    return fail(B200CG_ERR_INVALID_ARG, "b200cg_solve_batch moves every rhs and every solution: rhs_on_device / keep_x_on_device must be 0");
  NEED_GEOMETRY(P);
  for (int i = 0; i < count; ++i)
    if (!b_hosts[i] || !x_hosts[i]) return fail(B200CG_ERR_INVALID_ARG, "b_hosts[%d] / x_hosts[%d] is NULL", i, i);
  if (count == 0) return B200CG_OK;
  CU(cudaSetDevice(P->desc.device));
  BatchIo* io = nullptr;
  RET(batch_io(P, &io));
  memset(infos, 0, sizeof(b200cg_info) * (size_t)count);
  const size_t bytes = (size_t)local_count(P) * sizeof(double);
  // (anything an earlier call left in the staging buffers has been consumed: every entry point ends synchronised)
  int rc = B200CG_OK;
  auto cuda_ok = [&](cudaError_t e, const char* what) {
    if (e == cudaSuccess || rc != B200CG_OK) return;
    rc = fail(B200CG_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
  };
  cuda_ok(cudaMemcpyAsync(P->compact, b_hosts[0], bytes, cudaMemcpyHostToDevice, io->s_in), "H2D of the first right-hand side");
  cuda_ok(cudaEventRecord(io->h2d_done, io->s_in), "cudaEventRecord");
  int solved = 0;
  for (int i = 0; i < count && rc == B200CG_OK; ++i) {
    const BatchStep step{io, i + 1 < count ? b_hosts[i + 1] : nullptr, i};
    rc = solve_impl(P, prm, b_hosts[i], nullptr, x_hosts[i], &infos[i], nullptr, nullptr, stop_flag, &step);
    if (rc != B200CG_OK) break;
    ++solved;
    if (i > 0) {  // solution i-1 left under this solve's iterations
      cuda_ok(cudaEventSynchronize(io->d2h_done[(i - 1) & 1]), "D2H of a solution");
      if (rc == B200CG_OK && done) done(user, i - 1, &infos[i - 1]);
    }
    if (infos[i].stop_reason == B200CG_STOP_INTERRUPTED) break;  // requestStop ends the queue with this solve
  }
  // nothing may stay in flight on the caller's buffers, whatever happened
  const cudaError_t e_in = cudaStreamSynchronize(io->s_in), e_out = cudaStreamSynchronize(io->s_out);
  cuda_ok(e_in, "input copy stream");
  cuda_ok(e_out, "output copy stream");
  if (rc == B200CG_OK && solved > 0 && done) done(user, solved - 1, &infos[solved - 1]);
  for (int i = solved; i < count; ++i) infos[i].stop_reason = B200CG_STOP_INTERRUPTED;
  return rc;
}

extern "C" int b200cg_get_solution(b200cg_plan_t P, double* x_host) {
  if (!P || !x_host) return fail(B200CG_ERR_INVALID_ARG, "plan/x_host is NULL");
  if (!P->have_solution) return fail(B200CG_ERR_STATE, "no solution in the plan: call b200cg_solve first");
  CU(cudaSetDevice(P->desc.device));
  if (P->solution_in_csr)
    CU(cudaMemcpyAsync(x_host, P->csr.x, local_count(P) * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
  else
    RET(download_vector(P, P->x, x_host));
  CU(cudaStreamSynchronize(P->stream));
  return B200CG_OK;
}

extern "C" int b200cg_postprocess(b200cg_plan_t P, int op, double* residual_host, double* error_host) {
  if (!P) return fail(B200CG_ERR_INVALID_ARG, "plan is NULL");
  if (!P->have_solution) return fail(B200CG_ERR_STATE, "no solution in the plan: call b200cg_solve first");
  CU(cudaSetDevice(P->desc.device));
  cudaStream_t s = P->stream;
  const long long cnt = local_count(P);
  if ((op == B200CG_OP_CSR) != P->solution_in_csr)
    return fail(B200CG_ERR_STATE, "postprocess operator differs from the operator of the last solve");
  if (residual_host) {
    if (op == B200CG_OP_CSR) {
      if (!P->csr.row_map) return fail(B200CG_ERR_STATE, "no CSR matrix in the plan");
      // A x - b with the assembled matrix (dirichlet_solver.cpp:147-161): z[0] <- x, Az <- A x - b
      csr_residual_kernel<<<csr_grid(cnt, P->sms), CTA_THREADS, 0, s>>>(csr_args(&P->csr, P->d_state, P->d_partials, P->d_log, 0));
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(residual_host, P->csr.Az, cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
    } else {
      RET(ensure_scratch(P));
      RET(exchange_halo(P, P->x));
      TileArgs a = base_args(P);
      a.p_in = P->x;
      a.r_in = P->b;
      a.out = P->vb;
      RET((launch_tile<MODE_APPLY, F_SUB_B>(P, a, s)));
      RET(download_vector(P, P->vb, residual_host));
    }
  }
  if (error_host) {
    if (!P->have_u) return fail(B200CG_ERR_STATE, "error = x - u needs the true solution passed to the last solve");
    if (op == B200CG_OP_CSR) {
      csr_error_kernel<<<csr_grid(cnt, P->sms), CTA_THREADS, 0, s>>>(csr_args(&P->csr, P->d_state, P->d_partials, P->d_log, 0));
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(error_host, P->csr.Az, cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
    } else {
      gather_diff_kernel<<<ew_grid(P, cnt), CTA_THREADS, 0, s>>>(P->x, P->u, P->compact, P->g);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(error_host, P->compact, cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
  }
  CU(cudaStreamSynchronize(s));
  return B200CG_OK;
}

