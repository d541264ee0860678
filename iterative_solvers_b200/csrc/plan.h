// Internal host-side declarations of libb200cg shared by plan.cu (geometry, plan life cycle, data movement, CSR entry
// points) and solve.cu (graph loop, small-grid path, post-processing). Not part of the public ABI (include/b200cg.h).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200cg.h"
#include "kernels.cuh"
#include "csr_kernels.cuh"
#include "cluster_kernel.cuh"
#include "fused_kernel.cuh"
#include "comm.h"

// ------------------------------------------------------------------------------------------- errors
// Records the message for b200cg_last_error() (thread-local) and returns `code`.
int b200cg_fail(int code, const char* fmt, ...);
#define fail b200cg_fail

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? B200CG_ERR_NO_DEVICE \
                                                                                 : B200CG_ERR_CUDA, \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);    \
  } while (0)

#define RET(call)                       \
  do {                                  \
    int rc__ = (call);                  \
    if (rc__ != B200CG_OK) return rc__; \
  } while (0)


#define NEED_GEOMETRY(P)                                                                                      \
  do {                                                                                                        \
    if ((P)->generic)                                                                                         \
      return fail(B200CG_ERR_UNSUPPORTED, "%s needs a geometric plan (this one is B200CG_DOMAIN_GENERIC)", __func__); \
  } while (0)

namespace b200cg {

struct MgHierarchy;  // mg.cu

double now_ms();

// ------------------------------------------------------------------------------------------- plan
// Graph variants (key of b200cg_plan_s::graphs = variant * 4096 + iterations per graph)
enum { V_U = 1, V_REPORT = 2, V_CSR = 4, V_XDEFER = 8, V_FUSED = 16, V_TIMED = 32, V_MAXN = 64 };

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  int iters = 0;
  int kernels = 0;
};

struct TileTable {  // one per sweep flavour: 0 = dot phase, 1 = update without x, 2 = everything else, 3 = single sweep
  Tile* d_tiles = nullptr;
  int* d_cta_begin = nullptr;
  size_t tile_capacity = 0;
  int ctas_per_sm = 2;
  int strip_out = STRIP_OUT;   // columns written per strip; a strip's first staged storage column is
  int col_shift = 0;           //   strip * strip_out + col_shift  (fused_kernel.cuh: 480 / 2)
  int grid = 0, n_tiles = 0;
  bool balanced = false;       // feedback balancing applies (long marches)
  std::vector<double> weight;  // relative share of the sweep per CTA
};

// b200cg_solve_batch: what overlaps the copies of neighbouring solves with the iterations of the current one (lazy)
struct BatchIo {
  cudaStream_t s_in = nullptr, s_out = nullptr;  // H2D of the next right-hand side / D2H of the previous solution
  double* stage_out = nullptr;                   // second compact staging buffer (P->compact stages the inputs)
  cudaEvent_t h2d_done = nullptr;                // the next rhs has landed in P->compact
  cudaEvent_t in_free = nullptr;                 // P->compact has been scattered into P->b
  cudaEvent_t gathered = nullptr;                // the solution has been gathered into stage_out
  cudaEvent_t d2h_done[2] = {nullptr, nullptr};  // solution i is complete in host memory (slot i & 1)
};

}  // namespace b200cg

using namespace b200cg;  // internal header: the plan struct is the C ABI's opaque type and lives at global scope

struct b200cg_plan_s {
  b200cg_plan_desc desc;
  Geom g;
  int sms = 148;
  TileTable tile_tab[4];  // sweep work lists per flavour
  int balance_rounds = 0;  // feedback-balancing steps still to do (the first graph launches of the plan)
  int balance_rounds_fused = 0;  // the same for the single-sweep flavour, counted from its own first launches
  int fused_cw = FUSED_CW;                          // consumer warps of the single-sweep kernel: 7 (2 CTAs/SM) or 14 (1 CTA/SM, experiment)
  int shape_dot = 3, shape_upd = 2, shape_nox = 2;  // launch shapes of the hot flavours (launch_tile); measured best at 16384^2
  bool x_deferral = true;                           // REL_L2 without report: touch x every other iteration
  bool cluster_enabled = true;                      // small-grid path allowed (B200CG_CLUSTER=0 disables)
  bool single_sweep_default = true;                 // single-sweep iteration wherever it applies (B200CG_SINGLE_SWEEP=0: two sweeps)
  bool cluster16_ok = false;                        // a 16-CTA cluster of the small-grid kernel can be scheduled
  cudaStream_t stream = nullptr;
  size_t vec_elems = 0;  // doubles per pitched vector
  double* r[2] = {nullptr, nullptr};
  double* p[2] = {nullptr, nullptr};
  double* x = nullptr;
  double* b = nullptr;
  double* u = nullptr;
  double* va = nullptr;  // scratch vectors for apply / postprocess (lazy)
  double* vb = nullptr;
  double* compact = nullptr;  // staging buffer in the reference's compact ordering (local range)
  DevState* d_state = nullptr;
  DevState* h_state = nullptr;  // pinned mirror (the canonical copy the solve's tail reads)
  DevState* h_state_m[2] = {nullptr, nullptr};  // the graph loop's two read-back targets (launches are pipelined)
  CbRecord* h_log_m[2] = {nullptr, nullptr};
  cudaEvent_t ev_launch[2] = {};                // read-back of launch slot k complete
  CbRecord* d_log = nullptr;
  CbRecord* h_log = nullptr;  // pinned mirror
  int* h_stop = nullptr;      // mapped flag the cluster kernel polls (interrupt requests)
  unsigned long long* d_clock[4] = {nullptr, nullptr, nullptr, nullptr};  // per-CTA start/end stamps per sweep flavour
  int clock_ctas[4] = {0, 0, 0, 0};
  int* d_stop = nullptr;
  double* d_partials = nullptr;
  int partial_slots = 0;
  cudaEvent_t ev[11] = {};
  bool have_rhs = false, have_u = false, have_solution = false;
  bool generic = false;          // B200CG_DOMAIN_GENERIC: CSR entry points only, no pitched vectors
  bool solution_in_csr = false;  // the last solve ran on the assembled path
  std::map<int, GraphEntry> graphs;
  CsrData csr;
  Comm comm;
  // NVLink peer-memory exchange (sharded plans): IPC-mapped neighbour vectors and every rank's PeerSync block
  bool peer_mode = false;
  PeerSync* d_sync = nullptr;        // this rank's block
  PeerLinks* d_links = nullptr;      // device table of all ranks' blocks
  std::vector<void*> ipc_opened;     // everything cudaIpcOpenMemHandle returned (closed at destroy)
  double* nb_below_r[2] = {nullptr, nullptr};  // neighbour vectors (base pointers of its pitched buffers)
  double* nb_below_p[2] = {nullptr, nullptr};
  double* nb_above_r[2] = {nullptr, nullptr};
  double* nb_above_p[2] = {nullptr, nullptr};
  unsigned long long peer_epoch[2] = {0, 0};
  unsigned long long* d_peer_trace = nullptr;  // B200CG_PEER_TRACE=1: stamps of the single sweep's cross-rank step
  int64_t n_global = 0;
  std::vector<int> ycuts;  // row cuts of all ranks
  BatchIo* batch = nullptr;   // copy streams / staging of b200cg_solve_batch (created by its first call)
  MgHierarchy* mg = nullptr;  // level hierarchy of the opt-in multigrid preconditioner (built by the first solve that asks)
};

namespace b200cg {

// ---- shared helpers (defined in plan.cu)
int ew_grid(const b200cg_plan_s* P, long long work_items);       // grid of an element-wise kernel
long long local_count(const b200cg_plan_s* P);                   // unknowns owned by this rank
int upload_vector(b200cg_plan_s* P, const double* host, double* pitched);
int download_vector(b200cg_plan_s* P, const double* pitched, double* host);
int ensure_u(b200cg_plan_s* P);
int ensure_scratch(b200cg_plan_s* P);
int exchange_halo(b200cg_plan_s* P, double* v);                  // one-row halo over NCCL (no-op on one GPU)
int exchange_halo2(b200cg_plan_s* P, double* v0, double* v1);
int rebalance_tiles(b200cg_plan_s* P, int flavour);              // one feedback step of the work split
void drop_graphs(b200cg_plan_s* P, int variant_mask);            // destroys the cached graphs whose variant has a bit of the mask

}  // namespace b200cg
