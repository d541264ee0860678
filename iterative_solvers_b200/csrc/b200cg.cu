// libb200cg: C ABI (include/b200cg.h) over the sm_100a kernels in kernels.cuh / csr_kernels.cuh.
// Host side: plan (geometry + device buffers), CUDA-graph captured CG loop with device-resident scalars.
// No CPU fallback: every compute entry point fails loudly without a CUDA device.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200cg.h"
#include "kernels.cuh"
#include "csr_kernels.cuh"
#include "cluster_kernel.cuh"
#include "comm.h"

using namespace b200cg;

// ------------------------------------------------------------------------------------------- errors
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

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

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------- plan
struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  int iters = 0;
  int kernels = 0;
};

struct TileTable {  // one per sweep flavour: 0 = dot phase, 1 = update without x, 2 = everything else
  Tile* d_tiles = nullptr;
  int* d_cta_begin = nullptr;
  size_t tile_capacity = 0;
  int ctas_per_sm = 2;
  int grid = 0, n_tiles = 0;
  bool balanced = false;       // feedback balancing applies (long marches)
  std::vector<double> weight;  // relative share of the sweep per CTA
};

struct b200cg_plan_s {
  b200cg_plan_desc desc;
  Geom g;
  int sms = 148;
  TileTable tile_tab[3];  // sweep work lists per flavour
  int balance_rounds = 0;  // feedback-balancing steps still to do (the first graph launches of the plan)
  int shape_dot = 3, shape_upd = 2, shape_nox = 3;  // launch shapes of the hot flavours (launch_tile); measured best at 16384^2
  bool x_deferral = true;                           // REL_L2 without report: touch x every other iteration
  bool cluster_enabled = true;                      // small-grid path allowed (B200CG_CLUSTER=0 disables)
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
  DevState* h_state = nullptr;  // pinned mirror
  CbRecord* d_log = nullptr;
  CbRecord* h_log = nullptr;  // pinned mirror
  int* h_stop = nullptr;      // mapped flag the cluster kernel polls (interrupt requests)
  unsigned long long* d_clock[3] = {nullptr, nullptr, nullptr};  // per-CTA start/end stamps per sweep flavour
  int clock_ctas[3] = {0, 0, 0};
  int* d_stop = nullptr;
  double* d_partials = nullptr;
  int partial_slots = 0;
  cudaEvent_t ev[10] = {};
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
  int64_t n_global = 0;
  std::vector<int> ycuts;  // row cuts of all ranks
};

static long long row_start(const Geom& g, int y) {  // compact index of the first unknown of row y
  if (g.ysplit && y <= g.ysplit) return (long long)(y - 1) * g.wB;
  return g.NB + (long long)(y - g.ysplit - 1) * g.wU;
}

static int setup_geometry(b200cg_plan_s* P) {
  const b200cg_plan_desc& d = P->desc;
  Geom& g = P->g;
  memset(&g, 0, sizeof(g));
  if (d.domain == B200CG_DOMAIN_GENERIC) {
    if (d.generic_rows <= 0 || d.generic_rows > 2147483647LL)
      return fail(B200CG_ERR_INVALID_ARG, "generic plan needs 0 < generic_rows < 2^31 (int32 CSR indices)");
    if (d.world > 1) return fail(B200CG_ERR_UNSUPPORTED, "the CSR comparison path is single-GPU");
    P->n_global = d.generic_rows;
    g.lo = 0;
    g.hi = d.generic_rows;
    P->ycuts.assign(2, 0);
    return B200CG_OK;
  }
  if (d.domain == B200CG_DOMAIN_LSHAPE) {
    if (d.n != d.m || (d.n % 2) != 0 || d.n < 4)
      return fail(B200CG_ERR_INVALID_ARG,
                  "L-shaped domain needs even n == m >= 4 (got n=%d, m=%d): the reference numbering "
                  "(grid_system.cpp:103-111) is only self-consistent there",
                  d.n, d.m);
  } else if (d.domain == B200CG_DOMAIN_LSHAPE_ANY) {
    if (d.n < 4 || d.m < 4) return fail(B200CG_ERR_INVALID_ARG, "L-shaped domain needs n, m >= 4");
  } else if (d.domain == B200CG_DOMAIN_RECT) {
    if (d.n < 2 || d.m < 2) return fail(B200CG_ERR_INVALID_ARG, "RECT domain needs n, m >= 2");
  } else {
    return fail(B200CG_ERR_INVALID_ARG, "unknown domain kind %d", d.domain);
  }
  if (!(d.b > d.a) || !(d.d > d.c)) return fail(B200CG_ERR_INVALID_ARG, "empty domain [a,b]x[c,d]");
  g.n = d.n;
  g.m = d.m;
  g.a = d.a;
  g.c = d.c;
  g.hx = (d.b - d.a) / (d.n);  // grid_system.cpp:314-318
  g.hy = (d.d - d.c) / (d.m);
  g.A = -2 * (1 / (g.hx * g.hx) + 1 / (g.hy * g.hy));
  g.xk = 1 / (g.hx * g.hx);
  g.yk = 1 / (g.hy * g.hy);
  if (d.domain == B200CG_DOMAIN_LSHAPE || d.domain == B200CG_DOMAIN_LSHAPE_ANY) {
    g.xsplit = d.n / 2;
    g.ysplit = d.m / 2;
    g.wB = d.n - 1 - d.n / 2;  // = n/2 - 1 for even n (grid_system.cpp:108-111)
    g.wU = d.n - 1;
    g.NB = (long long)g.wB * (d.m / 2);
  } else {
    g.xsplit = 0;
    g.ysplit = 0;
    g.wB = 0;
    g.wU = d.n - 1;
    g.NB = 0;
  }
  P->n_global = row_start(g, d.m - 1) + g.wU;

  // row slabs balanced by unknowns (block B rows are narrower than block U rows)
  const int world = d.world > 1 ? d.world : 1;
  const int rank = d.world > 1 ? d.rank : 0;
  if (rank < 0 || rank >= world) return fail(B200CG_ERR_INVALID_ARG, "rank %d outside world %d", rank, world);
  if (world > d.m - 1) return fail(B200CG_ERR_INVALID_ARG, "more ranks (%d) than unknown rows (%d)", world, d.m - 1);
  P->ycuts.assign(world + 1, 1);
  {
    int y = 1;
    for (int k = 1; k < world; ++k) {
      const long long target = (long long)((double)P->n_global * k / world);
      while (y < d.m - 1 && row_start(g, y + 1) <= target) ++y;
      // keep at least one row per rank
      y = std::max(y, P->ycuts[k - 1] + 1);
      y = std::min(y, d.m - 1 - (world - k) + 1);
      P->ycuts[k] = y;
    }
    P->ycuts[world] = d.m;
  }
  g.ylo = P->ycuts[rank];
  g.yhi = P->ycuts[rank + 1];
  g.ybase = g.ylo - 1;
  g.yrows = g.yhi - g.ylo + 2;
  g.lo = row_start(g, g.ylo);
  g.hi = (g.yhi >= d.m) ? P->n_global : row_start(g, g.yhi);
  g.pitch = ((d.n + 1 + XOFF) + 15) / 16 * 16;

  return B200CG_OK;
}

// Cuts the sweep over this rank's rows into tiles and deals them to the resident CTAs.
// All (strip, row) pairs are linearised strip-major and cut into grid * tiles_per_cta ranges (split where a range
// crosses a strip end); CTA c gets the ranges c, c + grid, ... . A range's length is proportional to the weight of
// its CTA: 1 at first (equal split), then corrected from the measured per-CTA sweep times (rebalance_tiles) - SMs
// differ by 10-15 % in achieved memory throughput on the write-heavy sweeps, and the pattern is stable from launch
// to launch (profiles/r1_scheduling_experiments.md). A tile's two halo rows are amortised over its height.
// desc.tile_rows > 0 forces fixed-height tiles instead (tests: ragged heights, many tiles per CTA).
static void build_tiles(b200cg_plan_s* P, TileTable* tt, std::vector<Tile>* tiles, std::vector<int>* cta_begin) {
  const Geom& g = P->g;
  struct Col { int col0, y0, y1, xlo; };
  std::vector<Col> cols;  // one entry per (block, strip)
  const int strips = (g.n - 1) / STRIP_OUT + 1;
  const int yB0 = g.ylo, yB1 = g.ysplit ? std::max(g.ylo, std::min(g.yhi, g.ysplit + 1)) : g.ylo;
  const int yU0 = std::max(g.ylo, g.ysplit + 1), yU1 = std::max(yU0, g.yhi);
  if (yB1 > yB0)
    for (int s = (g.xsplit + 1) / STRIP_OUT; s < strips; ++s) cols.push_back({s * STRIP_OUT, yB0, yB1, g.xsplit + 1});
  if (yU1 > yU0)
    for (int s = 0; s < strips; ++s) cols.push_back({s * STRIP_OUT, yU0, yU1, 1});
  long long total = 0;
  for (const Col& c : cols) total += c.y1 - c.y0;
  const int max_grid = P->sms * tt->ctas_per_sm;
  std::vector<std::vector<Tile>> per_cta;
  tt->balanced = false;
  if (P->desc.tile_rows > 0) {
    std::vector<Tile> all;
    for (const Col& c : cols)
      for (int y = c.y0; y < c.y1; y += P->desc.tile_rows)
        all.push_back({c.col0, y, std::min(y + P->desc.tile_rows, c.y1), c.xlo});
    const int grid = (int)std::max<size_t>(1, std::min<size_t>(all.size(), (size_t)max_grid));
    per_cta.resize(grid);
    for (size_t i = 0; i < all.size(); ++i) per_cta[i % grid].push_back(all[i]);
  } else {
    const int MIN_ROWS = 4;       // below this the two halo rows dominate
    // ranges per CTA: every range costs two halo rows, so only long marches are split (4 x >= 128 rows)
    int tiles_per_cta = (int)std::max<long long>(1, std::min<long long>(4, total / ((long long)max_grid * 128)));
    if (const char* env = getenv("B200CG_TILES_PER_CTA")) tiles_per_cta = std::max(1, atoi(env));
    long long nranges = std::min<long long>((long long)max_grid * tiles_per_cta, std::max<long long>(1, total / MIN_ROWS));
    const int grid = (int)std::min<long long>(max_grid, nranges);
    if (nranges > grid) nranges = (nranges / grid) * grid;  // same count for every CTA
    per_cta.resize(std::max(grid, 1));
    if ((int)tt->weight.size() != grid) tt->weight.assign(grid, 1.0);
    tt->balanced = total / grid >= 64;  // worth balancing only when every CTA has a long march
    // cumulative weight at the range boundaries -> boundaries in rows
    double wsum = 0.0;
    for (long long r = 0; r < nranges; ++r) wsum += tt->weight[r % grid];
    size_t ci = 0;
    long long pos = 0;  // linear position of cols[ci].y0
    double wacc = 0.0;
    long long lo = 0;
    for (long long r = 0; r < nranges; ++r) {
      wacc += tt->weight[r % grid];
      long long hi = (r + 1 == nranges) ? total : (long long)std::llround((double)total * (wacc / wsum));
      hi = std::max(hi, lo);
      while (lo < hi) {
        while (ci < cols.size() && pos + (cols[ci].y1 - cols[ci].y0) <= lo) {
          pos += cols[ci].y1 - cols[ci].y0;
          ++ci;
        }
        const Col& c = cols[ci];
        const long long seg_hi = std::min<long long>(hi, pos + (c.y1 - c.y0));
        per_cta[r % per_cta.size()].push_back({c.col0, (int)(c.y0 + (lo - pos)), (int)(c.y0 + (seg_hi - pos)), c.xlo});
        lo = seg_hi;
      }
    }
  }
  tiles->clear();
  cta_begin->assign(1, 0);
  for (const auto& v : per_cta) {
    tiles->insert(tiles->end(), v.begin(), v.end());
    cta_begin->push_back((int)tiles->size());
  }
  tt->grid = (int)per_cta.size();
  tt->n_tiles = (int)tiles->size();
}

// (Re)builds a flavour's tile table on the host and puts it into its device arrays (allocated once, with slack:
// the graphs hold these pointers).
static int upload_tiles(b200cg_plan_s* P, TileTable* tt) {
  std::vector<Tile> tiles;
  std::vector<int> cta_begin;
  build_tiles(P, tt, &tiles, &cta_begin);
  if (!tt->d_tiles) {
    tt->tile_capacity = tiles.size() + 4 * ((size_t)(P->g.n - 1) / STRIP_OUT + 2) + 64;  // + strip-end splits
    CU(cudaMalloc(&tt->d_tiles, tt->tile_capacity * sizeof(Tile)));
    CU(cudaMalloc(&tt->d_cta_begin, ((size_t)P->sms * tt->ctas_per_sm + 1) * sizeof(int)));
  }
  if (tiles.size() > tt->tile_capacity) return fail(B200CG_ERR_STATE, "tile table overflow (%zu > %zu)", tiles.size(), tt->tile_capacity);
  if (!tiles.empty()) CU(cudaMemcpy(tt->d_tiles, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(tt->d_cta_begin, cta_begin.data(), cta_begin.size() * sizeof(int), cudaMemcpyHostToDevice));
  return B200CG_OK;
}

// One feedback step: every CTA swept work proportional to its weight and took T_c; give it weight * (mean T / T_c)
// (damped), so that all CTAs finish together. Called between graph launches while a plan is young.
static int rebalance_tiles(b200cg_plan_s* P, int flavour) {
  TileTable* tt = &P->tile_tab[flavour];
  if (!tt->balanced || tt->grid <= 1) return B200CG_OK;
  std::vector<unsigned long long> clk(2 * (size_t)tt->grid);
  CU(cudaMemcpy(clk.data(), P->d_clock[flavour], clk.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  unsigned long long t0 = ~0ull;
  for (int c = 0; c < tt->grid; ++c) {
    if (clk[2 * c] == 0 || clk[2 * c + 1] <= clk[2 * c]) return B200CG_OK;  // flavour did not run in this launch
    t0 = std::min(t0, clk[2 * c]);
  }
  double mean = 0.0;
  std::vector<double> T(tt->grid);
  for (int c = 0; c < tt->grid; ++c) {
    T[c] = (double)(clk[2 * c + 1] - t0);
    mean += T[c] / tt->grid;
  }
  double wsum = 0.0;
  for (int c = 0; c < tt->grid; ++c) {
    const double f = std::min(1.25, std::max(0.8, mean / T[c]));
    tt->weight[c] *= 1.0 + 0.8 * (f - 1.0);
    wsum += tt->weight[c];
  }
  for (double& w : tt->weight) w *= tt->grid / wsum;
  return upload_tiles(P, tt);
}

static int ew_grid(const b200cg_plan_s* P, long long work_items) {
  long long blocks = (work_items + CTA_THREADS - 1) / CTA_THREADS;
  long long cap = (long long)P->sms * 16;
  return (int)std::max(1LL, std::min(blocks, cap));
}

// ------------------------------------------------------------------------------------------- library
extern "C" const char* b200cg_last_error(void) { return g_last_error.c_str(); }
extern "C" int b200cg_version(void) { return B200CG_VERSION; }

extern "C" int b200cg_device_count(int* count) {
  if (!count) return fail(B200CG_ERR_INVALID_ARG, "count is NULL");
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
    return fail(B200CG_ERR_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(e));
  }
  return B200CG_OK;
}

extern "C" int b200cg_alloc_pinned(void** ptr, size_t bytes) {
  if (!ptr) return fail(B200CG_ERR_INVALID_ARG, "ptr is NULL");
  CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return B200CG_OK;
}
extern "C" int b200cg_free_pinned(void* ptr) {
  if (ptr) CU(cudaFreeHost(ptr));
  return B200CG_OK;
}

extern "C" int b200cg_comm_unique_id(void* id128) {
  if (!id128) return fail(B200CG_ERR_INVALID_ARG, "id128 is NULL");
  std::string err;
  if (!comm_unique_id(id128, &err)) return fail(B200CG_ERR_COMM, "%s", err.c_str());
  return B200CG_OK;
}

// Peer-memory exchange: swap CUDA IPC handles of r[2], p[2] and the PeerSync block over NCCL (once), map the two
// neighbours' vectors and every rank's block. All ranks then agree (min-reduction) whether the mapping worked
// everywhere; otherwise the plan keeps the NCCL path. B200CG_PEER=0 forces the NCCL path.
static int setup_peer_memory(b200cg_plan_s* P) {
  const int world = P->desc.world, rank = P->desc.rank;
  std::string err;
  bool mine = true;
  {
    const char* env = getenv("B200CG_PEER");
    if ((env && atoi(env) == 0) || world > PEER_MAX_RANKS) mine = false;
  }
  CU(cudaMalloc(&P->d_sync, sizeof(PeerSync)));
  CU(cudaMemset(P->d_sync, 0, sizeof(PeerSync)));
  constexpr int NH = 5;  // handles per rank: r0, r1, p0, p1, sync
  std::vector<cudaIpcMemHandle_t> all((size_t)world * NH);
  {
    cudaIpcMemHandle_t h[NH];
    void* ptrs[NH] = {P->r[0], P->r[1], P->p[0], P->p[1], P->d_sync};
    for (int k = 0; k < NH; ++k)
      if (cudaIpcGetMemHandle(&h[k], ptrs[k]) != cudaSuccess) {
        mine = false;
        memset(&h[k], 0, sizeof(h[k]));
        cudaGetLastError();
      }
    unsigned char *d_send = nullptr, *d_recv = nullptr;
    CU(cudaMalloc(&d_send, sizeof(h)));
    CU(cudaMalloc(&d_recv, sizeof(h) * world));
    CU(cudaMemcpyAsync(d_send, h, sizeof(h), cudaMemcpyHostToDevice, P->stream));
    if (!comm_allgather_bytes(&P->comm, d_send, d_recv, sizeof(h), P->stream, &err)) return fail(B200CG_ERR_COMM, "%s", err.c_str());
    CU(cudaMemcpyAsync(all.data(), d_recv, sizeof(h) * world, cudaMemcpyDeviceToHost, P->stream));
    CU(cudaStreamSynchronize(P->stream));
    cudaFree(d_send);
    cudaFree(d_recv);
  }
  PeerLinks links;
  memset(&links, 0, sizeof(links));
  links.rank = rank;
  links.world = world;
  auto open = [&](int r, int k) -> void* {
    void* q = nullptr;
    if (cudaIpcOpenMemHandle(&q, all[(size_t)r * NH + k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      mine = false;
      return nullptr;
    }
    P->ipc_opened.push_back(q);
    return q;
  };
  if (mine) {
    for (int r = 0; r < world && mine; ++r) links.sync[r] = (r == rank) ? P->d_sync : static_cast<PeerSync*>(open(r, 4));
    if (rank > 0)
      for (int k = 0; k < 2 && mine; ++k) {
        P->nb_below_r[k] = static_cast<double*>(open(rank - 1, k));
        P->nb_below_p[k] = static_cast<double*>(open(rank - 1, 2 + k));
      }
    if (rank + 1 < world)
      for (int k = 0; k < 2 && mine; ++k) {
        P->nb_above_r[k] = static_cast<double*>(open(rank + 1, k));
        P->nb_above_p[k] = static_cast<double*>(open(rank + 1, 2 + k));
      }
  }
  bool all_ok = false;
  if (!comm_all_agree(&P->comm, mine, &all_ok, P->stream, &err)) return fail(B200CG_ERR_COMM, "%s", err.c_str());
  P->peer_mode = all_ok;
  if (all_ok) {
    CU(cudaMalloc(&P->d_links, sizeof(PeerLinks)));
    CU(cudaMemcpy(P->d_links, &links, sizeof(links), cudaMemcpyHostToDevice));
  }
  return B200CG_OK;
}

// ------------------------------------------------------------------------------------------- plan API
static void free_plan(b200cg_plan_s* P) {
  if (!P) return;
  cudaSetDevice(P->desc.device);
  for (auto& kv : P->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void* q : P->ipc_opened) cudaIpcCloseMemHandle(q);
  cudaFree(P->d_sync);
  cudaFree(P->d_links);
  comm_destroy(&P->comm);
  csr_free(&P->csr);
  for (int i = 0; i < 2; ++i) {
    cudaFree(P->r[i]);
    cudaFree(P->p[i]);
  }
  cudaFree(P->x);
  cudaFree(P->b);
  cudaFree(P->u);
  cudaFree(P->va);
  cudaFree(P->vb);
  cudaFree(P->compact);
  cudaFree(P->d_state);
  cudaFree(P->d_log);
  cudaFree(P->d_partials);
  for (auto& tt : P->tile_tab) {
    cudaFree(tt.d_tiles);
    cudaFree(tt.d_cta_begin);
  }
  if (P->h_state) cudaFreeHost(P->h_state);
  if (P->h_log) cudaFreeHost(P->h_log);
  if (P->h_stop) cudaFreeHost(P->h_stop);
  for (auto& c : P->d_clock) cudaFree(c);
  for (auto& e : P->ev)
    if (e) cudaEventDestroy(e);
  if (P->stream) cudaStreamDestroy(P->stream);
  delete P;
}

static int alloc_vec(b200cg_plan_s* P, double** v) {
  CU(cudaMalloc(v, P->vec_elems * sizeof(double)));
  CU(cudaMemsetAsync(*v, 0, P->vec_elems * sizeof(double), P->stream));
  return B200CG_OK;
}

static int plan_create_impl(b200cg_plan_s* P) {
  RET(setup_geometry(P));  // argument errors first: they are reported even on a machine without a GPU
  int ndev = 0;
  RET(b200cg_device_count(&ndev));
  if (ndev <= 0) return fail(B200CG_ERR_NO_DEVICE, "no CUDA device visible: libb200cg has no CPU fallback");
  if (P->desc.device < 0 || P->desc.device >= ndev)
    return fail(B200CG_ERR_INVALID_ARG, "device %d outside [0, %d)", P->desc.device, ndev);
  CU(cudaSetDevice(P->desc.device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, P->desc.device));
  if (prop.major < 10)
    return fail(B200CG_ERR_UNSUPPORTED, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name,
                prop.major, prop.minor);
  P->sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking));
  for (auto& e : P->ev) CU(cudaEventCreate(&e));
  const Geom& g = P->g;
  P->generic = P->desc.domain == B200CG_DOMAIN_GENERIC;
  if (!P->generic) {
    P->vec_elems = (size_t)g.yrows * (size_t)g.pitch;
    for (int i = 0; i < 2; ++i) {
      RET(alloc_vec(P, &P->r[i]));
      RET(alloc_vec(P, &P->p[i]));
    }
    RET(alloc_vec(P, &P->x));
    RET(alloc_vec(P, &P->b));
    CU(cudaMalloc(&P->compact, std::max<long long>(g.hi - g.lo, 1) * sizeof(double)));
  }
  CU(cudaMalloc(&P->d_state, sizeof(DevState)));
  CU(cudaMemsetAsync(P->d_state, 0, sizeof(DevState), P->stream));
  CU(cudaHostAlloc(&P->h_state, sizeof(DevState), cudaHostAllocDefault));
  memset(P->h_state, 0, sizeof(DevState));
  CU(cudaMalloc(&P->d_log, sizeof(CbRecord) * CB_LOG_CAP));
  CU(cudaHostAlloc(&P->h_log, sizeof(CbRecord) * CB_LOG_CAP, cudaHostAllocDefault));
  for (auto& c : P->d_clock) {
    CU(cudaMalloc(&c, sizeof(unsigned long long) * 2 * (size_t)P->sms * 3));
    CU(cudaMemsetAsync(c, 0, sizeof(unsigned long long) * 2 * (size_t)P->sms * 3, P->stream));
  }
  CU(cudaHostAlloc(&P->h_stop, sizeof(int), cudaHostAllocMapped));
  *P->h_stop = 0;
  CU(cudaHostGetDevicePointer(&P->d_stop, P->h_stop, 0));
  {
    auto env_int = [](const char* name, int dflt) {
      const char* v = getenv(name);
      return v ? atoi(v) : dflt;
    };
    P->shape_dot = env_int("B200CG_SHAPE_DOT", P->shape_dot);
    P->shape_upd = env_int("B200CG_SHAPE_UPD", P->shape_upd);
    P->shape_nox = env_int("B200CG_SHAPE_NOX", P->shape_nox);
    P->x_deferral = env_int("B200CG_XDEFER", 1) != 0;
    P->balance_rounds = env_int("B200CG_BALANCE", 4);
    P->cluster_enabled = env_int("B200CG_CLUSTER", 1) != 0;
    if (!P->generic && P->cluster_enabled) {
      // probe once whether the non-portable 16-CTA cluster is schedulable with a full shared-memory carve-out
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(16, 1, 1);
      cfg.blockDim = dim3(CL_THREADS, 1, 1);
      cfg.dynamicSmemBytes = 200 * 1024;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 16;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaFuncSetAttribute(cg_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess &&
          cudaFuncSetAttribute(cg_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
          cudaOccupancyMaxActiveClusters(&nclusters, cg_cluster_kernel, &cfg) == cudaSuccess)
        P->cluster16_ok = nclusters > 0;
      cudaGetLastError();
    }
  }
  if (!P->generic) {
    // resident CTAs per SM of each flavour's launch shape (ShapeOf<>::CTAS)
    auto ctas_of = [](int shape) { return (shape == 1 || shape == 3) ? 3 : 2; };
    P->tile_tab[0].ctas_per_sm = ctas_of(P->shape_dot);
    P->tile_tab[1].ctas_per_sm = ctas_of(P->shape_nox);
    P->tile_tab[2].ctas_per_sm = 2;  // every other flavour runs a 2-CTAs/SM shape
    if (P->shape_upd == 1) P->shape_upd = 0;
    for (auto& tt : P->tile_tab) RET(upload_tiles(P, &tt));
  }
  P->partial_slots = P->sms * 16 + 64;
  CU(cudaMalloc(&P->d_partials, sizeof(double) * MAX_PARTIALS * (size_t)P->partial_slots));
  if (P->desc.world > 1) {
    std::string err;
    if (!P->desc.comm_id) return fail(B200CG_ERR_INVALID_ARG, "world > 1 needs comm_id (b200cg_comm_unique_id)");
    if (!comm_init(&P->comm, P->desc.comm_id, P->desc.rank, P->desc.world, P->stream, &err))
      return fail(B200CG_ERR_COMM, "%s", err.c_str());
    RET(setup_peer_memory(P));
  }
  CU(cudaStreamSynchronize(P->stream));
  return B200CG_OK;
}

extern "C" int b200cg_plan_create(b200cg_plan_t* plan, const b200cg_plan_desc* desc) {
  if (!plan || !desc) return fail(B200CG_ERR_INVALID_ARG, "plan/desc is NULL");
  *plan = nullptr;
  b200cg_plan_s* P = new b200cg_plan_s();
  P->desc = *desc;
  int rc = plan_create_impl(P);
  if (rc != B200CG_OK) {
    std::string keep = g_last_error;
    free_plan(P);
    cudaGetLastError();
    g_last_error = keep;
    return rc;
  }
  *plan = P;
  return B200CG_OK;
}

extern "C" int b200cg_plan_destroy(b200cg_plan_t plan) {
  free_plan(plan);
  return B200CG_OK;
}

extern "C" int b200cg_size(b200cg_plan_t P, int64_t* n) {
  if (!P || !n) return fail(B200CG_ERR_INVALID_ARG, "plan/n is NULL");
  *n = P->n_global;
  return B200CG_OK;
}
extern "C" int b200cg_local_range(b200cg_plan_t P, int64_t* lo, int64_t* hi) {
  if (!P || !lo || !hi) return fail(B200CG_ERR_INVALID_ARG, "plan/lo/hi is NULL");
  *lo = P->g.lo;
  *hi = P->g.hi;
  return B200CG_OK;
}

extern "C" int b200cg_partition(const b200cg_plan_desc* desc, int* y_lo, int* y_hi, int64_t* lo, int64_t* hi,
                                int64_t* n_unknowns) {
  if (!desc) return fail(B200CG_ERR_INVALID_ARG, "desc is NULL");
  b200cg_plan_s tmp;
  tmp.desc = *desc;
  RET(setup_geometry(&tmp));
  if (y_lo) *y_lo = tmp.g.ylo;
  if (y_hi) *y_hi = tmp.g.yhi;
  if (lo) *lo = tmp.g.lo;
  if (hi) *hi = tmp.g.hi;
  if (n_unknowns) *n_unknowns = tmp.n_global;
  return B200CG_OK;
}

// ------------------------------------------------------------------------------------------- data movement
static long long local_count(const b200cg_plan_s* P) { return P->g.hi - P->g.lo; }

static int upload_vector(b200cg_plan_s* P, const double* host, double* pitched) {
  const long long cnt = local_count(P);
  CU(cudaMemcpyAsync(P->compact, host, cnt * sizeof(double), cudaMemcpyHostToDevice, P->stream));
  scatter_compact_kernel<<<ew_grid(P, cnt), CTA_THREADS, 0, P->stream>>>(P->compact, pitched, P->g);
  CU(cudaGetLastError());
  return B200CG_OK;
}
static int download_vector(b200cg_plan_s* P, const double* pitched, double* host) {
  const long long cnt = local_count(P);
  gather_compact_kernel<<<ew_grid(P, cnt), CTA_THREADS, 0, P->stream>>>(pitched, P->compact, P->g);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(host, P->compact, cnt * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
  return B200CG_OK;
}
static int ensure_u(b200cg_plan_s* P) {
  if (!P->u) RET(alloc_vec(P, &P->u));
  return B200CG_OK;
}
static int ensure_scratch(b200cg_plan_s* P) {
  if (!P->va) RET(alloc_vec(P, &P->va));
  if (!P->vb) RET(alloc_vec(P, &P->vb));
  return B200CG_OK;
}
// one-row halo exchange of a pitched vector with the slab neighbours (no-op on a single GPU)
static int exchange_halo(b200cg_plan_s* P, double* v) {
  if (P->desc.world <= 1) return B200CG_OK;
  std::string err;
  const Geom& g = P->g;
  double* first_owned = v + (size_t)1 * g.pitch;
  double* last_owned = v + (size_t)(g.yrows - 2) * g.pitch;
  double* halo_below = v;
  double* halo_above = v + (size_t)(g.yrows - 1) * g.pitch;
  if (!comm_halo(&P->comm, first_owned, last_owned, halo_below, halo_above, g.pitch, P->stream, &err))
    return fail(B200CG_ERR_COMM, "%s", err.c_str());
  return B200CG_OK;
}

#define NEED_GEOMETRY(P)                                                                                      \
  do {                                                                                                        \
    if ((P)->generic)                                                                                         \
      return fail(B200CG_ERR_UNSUPPORTED, "%s needs a geometric plan (this one is B200CG_DOMAIN_GENERIC)", __func__); \
  } while (0)

static int exchange_halo2(b200cg_plan_s* P, double* v0, double* v1) {
  if (P->desc.world <= 1) return B200CG_OK;
  std::string err;
  if (!comm_halo2(&P->comm, v0, v1, P->g.yrows, P->g.pitch, P->stream, &err)) return fail(B200CG_ERR_COMM, "%s", err.c_str());
  return B200CG_OK;
}

extern "C" int b200cg_build_rhs(b200cg_plan_t P) {
  if (!P) return fail(B200CG_ERR_INVALID_ARG, "plan is NULL");
  NEED_GEOMETRY(P);
  CU(cudaSetDevice(P->desc.device));
  setup_kernel<<<ew_grid(P, local_count(P)), CTA_THREADS, 0, P->stream>>>(P->b, nullptr, P->g, 0);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(P->stream));
  P->have_rhs = true;
  return B200CG_OK;
}
extern "C" int b200cg_set_rhs(b200cg_plan_t P, const double* b_host) {
  if (!P || !b_host) return fail(B200CG_ERR_INVALID_ARG, "plan/b_host is NULL");
  NEED_GEOMETRY(P);
  CU(cudaSetDevice(P->desc.device));
  RET(upload_vector(P, b_host, P->b));
  CU(cudaStreamSynchronize(P->stream));
  P->have_rhs = true;
  return B200CG_OK;
}
extern "C" int b200cg_get_rhs(b200cg_plan_t P, double* b_host) {
  if (!P || !b_host) return fail(B200CG_ERR_INVALID_ARG, "plan/b_host is NULL");
  NEED_GEOMETRY(P);
  if (!P->have_rhs) return fail(B200CG_ERR_STATE, "no rhs in the plan: call b200cg_build_rhs or b200cg_set_rhs first");
  CU(cudaSetDevice(P->desc.device));
  RET(download_vector(P, P->b, b_host));
  CU(cudaStreamSynchronize(P->stream));
  return B200CG_OK;
}
static int setup_to_host(b200cg_plan_s* P, int what, double* host) {
  const long long cnt = local_count(P);
  setup_kernel<<<ew_grid(P, cnt), CTA_THREADS, 0, P->stream>>>(nullptr, P->compact, P->g, what);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(host, P->compact, cnt * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
  CU(cudaStreamSynchronize(P->stream));
  return B200CG_OK;
}
extern "C" int b200cg_get_true_solution(b200cg_plan_t P, double* u_host) {
  if (!P || !u_host) return fail(B200CG_ERR_INVALID_ARG, "plan/u_host is NULL");
  NEED_GEOMETRY(P);
  CU(cudaSetDevice(P->desc.device));
  return setup_to_host(P, 1, u_host);
}
extern "C" int b200cg_get_coords(b200cg_plan_t P, double* xs, double* ys) {
  if (!P || !xs || !ys) return fail(B200CG_ERR_INVALID_ARG, "plan/xs/ys is NULL");
  NEED_GEOMETRY(P);
  CU(cudaSetDevice(P->desc.device));
  RET(setup_to_host(P, 2, xs));
  return setup_to_host(P, 3, ys);
}

// ------------------------------------------------------------------------------------------- operator
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

// ------------------------------------------------------------------------------------------- CSR entry points
extern "C" int b200cg_set_csr(b200cg_plan_t P, int64_t nrows, int64_t nnz, const int* row_map, const int* entries,
                              const double* values) {
  if (!P || !row_map || !entries || !values) return fail(B200CG_ERR_INVALID_ARG, "NULL argument");
  if (P->desc.world > 1) return fail(B200CG_ERR_UNSUPPORTED, "the CSR comparison path is single-GPU");
  if (nrows != P->n_global) return fail(B200CG_ERR_INVALID_ARG, "nrows %lld != unknowns %lld", (long long)nrows, (long long)P->n_global);
  CU(cudaSetDevice(P->desc.device));
  std::string err;
  int rc = csr_upload(&P->csr, nrows, nnz, row_map, entries, values, P->stream, &err);
  if (rc) return fail(rc, "%s", err.c_str());
  return B200CG_OK;
}
extern "C" int b200cg_assemble_csr(b200cg_plan_t P, int64_t* nnz) {
  if (!P) return fail(B200CG_ERR_INVALID_ARG, "plan is NULL");
  NEED_GEOMETRY(P);
  if (P->desc.world > 1) return fail(B200CG_ERR_UNSUPPORTED, "the CSR comparison path is single-GPU");
  CU(cudaSetDevice(P->desc.device));
  std::string err;
  int rc = csr_assemble(&P->csr, P->g, P->n_global, P->sms, P->stream, &err);
  if (rc) return fail(rc, "%s", err.c_str());
  if (nnz) *nnz = P->csr.nnz;
  return B200CG_OK;
}
extern "C" int b200cg_get_csr(b200cg_plan_t P, int* row_map, int* entries, double* values) {
  if (!P) return fail(B200CG_ERR_INVALID_ARG, "plan is NULL");
  if (!P->csr.row_map) return fail(B200CG_ERR_STATE, "no CSR matrix in the plan");
  CU(cudaSetDevice(P->desc.device));
  if (row_map) CU(cudaMemcpy(row_map, P->csr.row_map, (P->csr.nrows + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  if (entries) CU(cudaMemcpy(entries, P->csr.entries, P->csr.nnz * sizeof(int), cudaMemcpyDeviceToHost));
  if (values) CU(cudaMemcpy(values, P->csr.values, P->csr.nnz * sizeof(double), cudaMemcpyDeviceToHost));
  return B200CG_OK;
}
extern "C" int b200cg_csr_apply(b200cg_plan_t P, const double* x_host, double* y_host) {
  if (!P || !x_host || !y_host) return fail(B200CG_ERR_INVALID_ARG, "plan/x_host/y_host is NULL");
  if (!P->csr.row_map) return fail(B200CG_ERR_STATE, "no CSR matrix in the plan");
  CU(cudaSetDevice(P->desc.device));
  std::string err;
  int rc = csr_ensure_vectors(&P->csr, P->stream, &err);
  if (rc) return fail(rc, "%s", err.c_str());
  const long long N = P->csr.nrows;
  CU(cudaMemcpyAsync(P->csr.z[0], x_host, N * sizeof(double), cudaMemcpyHostToDevice, P->stream));
  csr_spmv_kernel<0><<<csr_grid(N, P->sms), CTA_THREADS, 0, P->stream>>>(csr_args(&P->csr, P->d_state, P->d_partials, P->d_log, 0));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(y_host, P->csr.Az, N * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
  CU(cudaStreamSynchronize(P->stream));
  return B200CG_OK;
}

// ------------------------------------------------------------------------------------------- solve
enum { V_U = 1, V_REPORT = 2, V_CSR = 4, V_XDEFER = 8 };

// Captures `iters` CG iterations (even, so the ping-pong buffers return to their start) plus the status
// read-back into one executable graph. Event-record nodes bracket the kernels of the first iteration.
static int build_graph(b200cg_plan_s* P, int variant, int iters, GraphEntry* out) {
  cudaStream_t s = P->stream;
  const bool with_u = variant & V_U, report = variant & V_REPORT, csr = variant & V_CSR, xdefer = variant & V_XDEFER;
  int kernels = 0;
  CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = B200CG_OK;
  for (int k = 0; k < iters && rc == B200CG_OK; ++k) {
    const int par = k & 1;
    if (k == 0) cudaEventRecordWithFlags(P->ev[0], s, cudaEventRecordExternal);
    if (csr) {
      // assembled path: p update + SpMV + dots, then the shared update pass
      csr_spmv_kernel<1><<<csr_grid(P->csr.nrows, P->sms), CTA_THREADS, 0, s>>>(
          csr_args(&P->csr, P->d_state, P->d_partials, P->d_log, par));
      ++kernels;
      if (k == 0) cudaEventRecordWithFlags(P->ev[1], s, cudaEventRecordExternal);
      if (with_u)
        csr_update_kernel<1><<<csr_grid(P->csr.nrows, P->sms), CTA_THREADS, 0, s>>>(
            csr_args(&P->csr, P->d_state, P->d_partials, P->d_log, par));
      else
        csr_update_kernel<0><<<csr_grid(P->csr.nrows, P->sms), CTA_THREADS, 0, s>>>(
            csr_args(&P->csr, P->d_state, P->d_partials, P->d_log, par));
      ++kernels;
      if (k == 0) cudaEventRecordWithFlags(P->ev[2], s, cudaEventRecordExternal);
      continue;
    }
    TileArgs a = base_args(P);
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
      const Geom& g = P->g;
      const int rank = P->desc.rank;
      if (rank > 0) {  // neighbour below: its top halo row is its last stored row
        const size_t rows_below = (size_t)(P->ycuts[rank] - P->ycuts[rank - 1]) + 2;
        a.nb_r_below = P->nb_below_r[par ^ 1] + (rows_below - 1) * g.pitch;
        a.nb_p_below = P->nb_below_p[par ^ 1] + (rows_below - 1) * g.pitch;
      }
      if (rank + 1 < P->desc.world) {  // neighbour above: its bottom halo row is its stored row 0
        a.nb_r_above = P->nb_above_r[par ^ 1];
        a.nb_p_above = P->nb_above_p[par ^ 1];
      }
    }
    rc = launch_tile<MODE_DOT, 0>(P, a, s);
    ++kernels;
    if (rc == B200CG_OK && peer) {
      peer_finalize_kernel<<<1, 32, 0, s>>>(P->d_state, P->d_log, P->d_links, 1, fl);
      ++kernels;
    } else if (rc == B200CG_OK && P->desc.world > 1) {
      rc = reduce_and_finalize(P, 1, fl, false, s);
      ++kernels;
    }
    if (k == 0) cudaEventRecordWithFlags(P->ev[1], s, cudaEventRecordExternal);
    if (k == 1) cudaEventRecordWithFlags(P->ev[8], s, cudaEventRecordExternal);
    if (rc != B200CG_OK) break;
    if (xdefer) rc = (k & 1) ? launch_tile<MODE_UPD, F_X2>(P, a, s) : launch_tile<MODE_UPD, F_NOX>(P, a, s);
    else if (report && with_u) rc = launch_tile<MODE_UPD, F_REPORT | F_U>(P, a, s);
    else if (report) rc = launch_tile<MODE_UPD, F_REPORT>(P, a, s);
    else if (with_u) rc = launch_tile<MODE_UPD, F_U>(P, a, s);
    else rc = launch_tile<MODE_UPD, 0>(P, a, s);
    ++kernels;
    if (rc == B200CG_OK && peer) {
      peer_finalize_kernel<<<1, 32, 0, s>>>(P->d_state, P->d_log, P->d_links, 2, fl);
      ++kernels;
    } else if (rc == B200CG_OK && P->desc.world > 1) {
      rc = reduce_and_finalize(P, 2, fl, /*with_max=*/!xdefer, s);
      ++kernels;
      if (rc == B200CG_OK) rc = exchange_halo2(P, P->r[par ^ 1], P->p[par ^ 1]);
    }
    if (k == 0) cudaEventRecordWithFlags(P->ev[2], s, cudaEventRecordExternal);
    if (k == 1 && !report) cudaEventRecordWithFlags(P->ev[9], s, cudaEventRecordExternal);
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
  cudaMemcpyAsync(P->h_state, P->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(P->h_log, P->d_log, sizeof(CbRecord) * CB_LOG_CAP, cudaMemcpyDeviceToHost, s);
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
  // small grids are launch-bound: long graphs; big grids: keep the stop/interrupt latency around 0.1 s
  const long long n = local_count(P);
  if (n <= (1LL << 20)) return 100;
  if (n <= (1LL << 24)) return 50;
  return 20;
}

extern "C" int b200cg_solve(b200cg_plan_t P, const b200cg_params* prm, const double* b_host, const double* u_host,
                            double* x_host, b200cg_info* info, b200cg_iter_cb cb, void* user,
                            const volatile int* stop_flag) {
  if (!P || !prm || !info) return fail(B200CG_ERR_INVALID_ARG, "plan/params/info is NULL");
  if (prm->op != B200CG_OP_MATRIX_FREE && prm->op != B200CG_OP_CSR) return fail(B200CG_ERR_INVALID_ARG, "unknown operator %d", prm->op);
  if (prm->rule != B200CG_RULE_REL_L2 && prm->rule != B200CG_RULE_MAXNORM) return fail(B200CG_ERR_INVALID_ARG, "unknown rule %d", prm->rule);
  if (!prm->rhs_on_device && !b_host) return fail(B200CG_ERR_INVALID_ARG, "b_host is NULL and rhs_on_device is 0");
  if (prm->rhs_on_device && !P->have_rhs) return fail(B200CG_ERR_STATE, "rhs_on_device set but the plan holds no rhs");
  if (!prm->keep_x_on_device && !x_host) return fail(B200CG_ERR_INVALID_ARG, "x_host is NULL and keep_x_on_device is 0");
  const bool csr = prm->op == B200CG_OP_CSR;
  if (!csr) NEED_GEOMETRY(P);
  if (csr && !P->csr.row_map) return fail(B200CG_ERR_STATE, "CSR solve without a matrix: call b200cg_set_csr / b200cg_assemble_csr");
  if (csr && prm->rule == B200CG_RULE_REL_L2 && cb) return fail(B200CG_ERR_UNSUPPORTED, "per-iteration report callbacks exist only on the matrix-free path");
  memset(info, 0, sizeof(*info));
  const double t_begin = now_ms();
  CU(cudaSetDevice(P->desc.device));
  cudaStream_t s = P->stream;
  const long long cnt = local_count(P);
  info->local_unknowns = cnt;
  P->have_solution = false;

  // ---- inputs
  CU(cudaEventRecord(P->ev[3], s));
  const bool with_u = (u_host != nullptr);
  if (csr) {
    // the assembled path works on compact vectors: host data goes straight into them
    std::string err;
    int rc = csr_ensure_vectors(&P->csr, s, &err);
    if (rc) return fail(rc, "%s", err.c_str());
    if (!prm->rhs_on_device) {
      CU(cudaMemcpyAsync(P->csr.b, b_host, cnt * sizeof(double), cudaMemcpyHostToDevice, s));
      info->h2d_bytes += cnt * (int64_t)sizeof(double);
    } else {
      gather_compact_kernel<<<ew_grid(P, cnt), CTA_THREADS, 0, s>>>(P->b, P->csr.b, P->g);
      info->kernel_launches += 1;
    }
    if (with_u) {
      CU(cudaMemcpyAsync(P->csr.u, u_host, cnt * sizeof(double), cudaMemcpyHostToDevice, s));
      info->h2d_bytes += cnt * (int64_t)sizeof(double);
    }
    P->csr.has_u = with_u;
  } else {
    if (!prm->rhs_on_device) {
      RET(upload_vector(P, b_host, P->b));
      P->have_rhs = true;
      info->h2d_bytes += cnt * (int64_t)sizeof(double);
      info->kernel_launches += 1;
    }
    if (with_u) {
      RET(ensure_u(P));
      RET(upload_vector(P, u_host, P->u));
      info->h2d_bytes += cnt * (int64_t)sizeof(double);
      info->kernel_launches += 1;
    }
  }
  P->have_u = with_u;
  CU(cudaEventRecord(P->ev[4], s));

  // ---- device-side solver state
  const bool report = (prm->rule == B200CG_RULE_REL_L2) && (cb != nullptr);
  DevState hs;
  memset(&hs, 0, sizeof(hs));
  hs.eps_rel = prm->eps_rel;
  hs.eps_p = prm->eps_p;
  hs.eps_r = prm->eps_r;
  hs.eps_e = prm->eps_e;
  hs.max_it = prm->max_it;
  hs.rule = prm->rule;
  hs.has_u = with_u ? 1 : 0;
  hs.callback_every = (cb && prm->rule == B200CG_RULE_MAXNORM) ? (prm->callback_every > 0 ? prm->callback_every : 100) : 0;
  hs.epoch[0] = P->peer_epoch[0];  // the PeerSync flags are monotonic over the plan's life
  hs.epoch[1] = P->peer_epoch[1];
  *P->h_state = hs;
  CU(cudaMemcpyAsync(P->d_state, P->h_state, sizeof(DevState), cudaMemcpyHostToDevice, s));

  // ---- path: grids that fit one thread-block cluster's shared memory run as a single resident kernel
  int cl_rows = 0;
  size_t cl_smem = 0;
  int cl_ctas = 0;
  {
    const long long cb_records = 2 + (long long)std::max(prm->max_it, 0) / 100;
    const bool eligible = !csr && !report && prm->small_grid_path != 1 && P->cluster_enabled &&
                          !(cb && cb_records > CB_LOG_CAP);
    if (eligible) cl_ctas = cluster_ctas_for(P, &cl_rows, &cl_smem);
    if (prm->small_grid_path == 2 && cl_ctas == 0)
      return fail(B200CG_ERR_UNSUPPORTED, "small_grid_path = 2 but this solve cannot run in one cluster "
                                          "(grid too large, sharded plan, CSR operator or per-iteration report)");
  }
  const bool use_cluster = cl_ctas > 0;

  bool xdefer = false;
  unsigned int consumed = 0;
  bool interrupted = false;
  double dot_ms = 0.0, upd_even = 0.0, upd_odd = 0.0;
  int samples = 0, it_before = 0, it_launch0 = 0;
  (void)it_before;
  (void)it_launch0;
  if (use_cluster) {
    *P->h_stop = 0;
    CU(cudaEventRecord(P->ev[5], s));
    RET(launch_cluster_solve(P, cl_ctas, cl_rows, cl_smem, with_u, s));
    info->kernel_launches += 1;
    CU(cudaMemcpyAsync(P->h_state, P->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(P->h_log, P->d_log, sizeof(CbRecord) * CB_LOG_CAP, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(P->ev[8], s));
    // the kernel polls the mapped flag every CL_POLL_EVERY iterations; forward the caller's stop request
    for (int spins = 0; cudaEventQuery(P->ev[8]) == cudaErrorNotReady; ++spins) {
      if (stop_flag && *stop_flag) *P->h_stop = 1;
      if (spins > 2000) std::this_thread::sleep_for(std::chrono::microseconds(50));  // long solve: stop burning a core
    }
    CU(cudaStreamSynchronize(s));
    const DevState& st = *P->h_state;
    interrupted = st.stop_reason == B200CG_STOP_INTERRUPTED;
    if (cb)
      for (; consumed < st.n_log; ++consumed) {
        const CbRecord& rec = P->h_log[consumed % CB_LOG_CAP];
        cb(user, (int)rec.it, rec.precision, rec.residual, rec.error);
      }
  } else {
    if (csr) {
      csr_init_kernel<<<csr_grid(cnt, P->sms), CTA_THREADS, 0, s>>>(csr_args(&P->csr, P->d_state, P->d_partials, P->d_log, 0));
      CU(cudaGetLastError());
      info->kernel_launches += 1;
    } else {
      const Geom& g = P->g;
      InitArgs ia;
      ia.b = P->b;
      ia.u = with_u ? P->u : nullptr;
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
      info->kernel_launches += 1;
      if (P->desc.world > 1) {
        RET(reduce_and_finalize(P, 0, with_u ? F_U : 0, true, s));
        RET(exchange_halo2(P, P->r[0], P->p[0]));
        info->kernel_launches += 1;
      }
    }

    // ---- the captured loop
    int K = prm->iters_per_graph > 0 ? prm->iters_per_graph : default_iters_per_graph(P);
    if (prm->max_it > 0) K = std::min(K, prm->max_it + 1);
    K = std::max(2, (K + 1) & ~1);
    K = std::min(K, report ? CB_LOG_CAP / 2 : CB_LOG_CAP);
    // x-deferral: the relative-residual rule never looks at x, so x is only touched every other iteration
    xdefer = P->x_deferral && !csr && !report && prm->rule == B200CG_RULE_REL_L2;
    const int variant = xdefer ? V_XDEFER : ((with_u ? V_U : 0) | (report ? V_REPORT : 0) | (csr ? V_CSR : 0));
    const int key = variant * 4096 + K;
    GraphEntry& ge = P->graphs[key];
    if (!ge.exec) RET(build_graph(P, variant, K, &ge));

    CU(cudaEventRecord(P->ev[5], s));
    // the init kernel's verdict (0 iterations) and record come back with the first graph launch
    for (;;) {
      CU(cudaGraphLaunch(ge.exec, s));
      info->kernel_launches += ge.kernels;
      CU(cudaStreamSynchronize(s));
      const DevState& st = *P->h_state;
      // Event nodes bracket the kernels of the first two captured iterations; count the sample only if those
      // iterations really ran in this launch (it advanced by at least 2 and the run was not already over).
      if (!csr && !report && st.it - it_before >= 2) {
        float d0 = 0.f, u0 = 0.f, d1 = 0.f, u1 = 0.f;
        if (cudaEventElapsedTime(&d0, P->ev[0], P->ev[1]) == cudaSuccess &&
            cudaEventElapsedTime(&u0, P->ev[1], P->ev[2]) == cudaSuccess &&
            cudaEventElapsedTime(&d1, P->ev[2], P->ev[8]) == cudaSuccess &&
            cudaEventElapsedTime(&u1, P->ev[8], P->ev[9]) == cudaSuccess) {
          dot_ms += 0.5 * (d0 + d1);
          upd_even += u0;
          upd_odd += u1;
          ++samples;
        }
      } else if ((csr || report) && st.it - it_before >= 1) {
        float d0 = 0.f, u0 = 0.f;
        if (cudaEventElapsedTime(&d0, P->ev[0], P->ev[1]) == cudaSuccess &&
            cudaEventElapsedTime(&u0, P->ev[1], P->ev[2]) == cudaSuccess) {
          dot_ms += d0;
          upd_even += u0;
          upd_odd += u0;
          ++samples;
        }
      }
      it_before = st.it;
      if (cb) {
        for (; consumed < st.n_log; ++consumed) {
          const CbRecord& rec = P->h_log[consumed % CB_LOG_CAP];
          cb(user, (int)rec.it, rec.precision, rec.residual, rec.error);
        }
      } else {
        consumed = st.n_log ? st.n_log : 1;
      }
      if (consumed == 0) consumed = 1;
      if (st.done) break;
      if (P->balance_rounds > 0 && !csr && st.it - it_launch0 >= 2) {
        // young plan: correct the static split from the measured per-CTA sweep times (the stream is idle here)
        for (int fl = 0; fl < 3; ++fl) RET(rebalance_tiles(P, fl));
        --P->balance_rounds;
      }
      it_launch0 = st.it;
      if (stop_flag && *stop_flag) {
        interrupted = true;
        break;
      }
    }
}
  CU(cudaEventRecord(P->ev[6], s));

  // ---- outputs
  const DevState st = *P->h_state;
  P->peer_epoch[0] = st.epoch[0];
  P->peer_epoch[1] = st.epoch[1];
  if (st.comm_error) return fail(B200CG_ERR_COMM, "peer-memory exchange timed out after %d iterations (a rank stopped publishing)", st.it);
  if (xdefer && st.x_pending) {  // the loop ended on an even iteration: settle x += alpha * p
    const Geom& g = P->g;
    const size_t begin = (size_t)(g.ylo - g.ybase) * g.pitch, count = (size_t)(g.yhi - g.ylo) * g.pitch;
    x_flush_kernel<<<ew_grid(P, (long long)(count / 2)), CTA_THREADS, 0, s>>>(P->x, P->p[st.it & 1], P->d_state, begin, count);
    CU(cudaGetLastError());
    info->kernel_launches += 1;
  }
  P->solution_in_csr = csr;
  if (!prm->keep_x_on_device) {
    if (csr) {
      CU(cudaMemcpyAsync(x_host, P->csr.x, cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
    } else {
      RET(download_vector(P, P->x, x_host));
      info->kernel_launches += 1;
    }
    info->d2h_bytes += cnt * (int64_t)sizeof(double);
  }
  CU(cudaEventRecord(P->ev[7], s));
  CU(cudaStreamSynchronize(s));
  P->have_solution = true;

  info->iterations = st.it;
  info->converged = interrupted ? 0 : st.converged;
  info->stop_reason = interrupted ? B200CG_STOP_INTERRUPTED : st.stop_reason;
  info->r0_l2 = st.r0_norm;
  info->r_l2 = st.r_norm;
  info->r_max = st.r_max;
  info->dx_max = st.dx_max;
  info->err_max = st.err_max;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, P->ev[3], P->ev[4]);
  info->h2d_ms = ms;
  cudaEventElapsedTime(&ms, P->ev[5], P->ev[6]);
  info->solve_ms = ms;
  cudaEventElapsedTime(&ms, P->ev[6], P->ev[7]);
  info->d2h_ms = ms;
  cudaEventElapsedTime(&ms, P->ev[3], P->ev[7]);
  info->device_ms = ms;
  info->dot_kernel_ms = samples ? dot_ms / samples : 0.0;
  info->upd_kernel_ms = samples ? 0.5 * (upd_even + upd_odd) / samples : 0.0;
  info->upd_even_ms = samples ? upd_even / samples : 0.0;
  info->upd_odd_ms = samples ? upd_odd / samples : 0.0;
  info->x_deferral = xdefer ? 1 : 0;
  info->cluster_path = use_cluster ? 1 : 0;
  info->peer_exchange = (P->desc.world > 1 && P->peer_mode && !report && !use_cluster) ? 1 : 0;
  info->kernel_samples = samples;
  // MSGSolver fires one more callback after the loop with the final values (msg_solver.cpp:193-195)
  if (cb && prm->rule == B200CG_RULE_MAXNORM) cb(user, st.it, st.dx_max, st.r_max, st.err_max);
  info->total_ms = now_ms() - t_begin;
  return B200CG_OK;
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

extern "C" int b200cg_cta_times(b200cg_plan_t P, int flavour, uint64_t* out, int capacity, int* n_ctas) {
  if (!P || !out || !n_ctas) return fail(B200CG_ERR_INVALID_ARG, "plan/out/n_ctas is NULL");
  if (flavour < 0 || flavour > 2) return fail(B200CG_ERR_INVALID_ARG, "flavour %d outside [0, 2]", flavour);
  NEED_GEOMETRY(P);
  const int n = std::min(P->clock_ctas[flavour], capacity);
  CU(cudaSetDevice(P->desc.device));
  CU(cudaMemcpy(out, P->d_clock[flavour], sizeof(unsigned long long) * 2 * (size_t)n, cudaMemcpyDeviceToHost));
  *n_ctas = n;
  return B200CG_OK;
}
