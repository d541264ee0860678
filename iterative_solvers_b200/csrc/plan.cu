// libb200cg, host side 1/2: errors, geometry and work split, plan life cycle, peer-memory setup, data movement and
// the CSR entry points of the C ABI (include/b200cg.h). The solve loop lives in solve.cu.
// No CPU fallback: every compute entry point fails loudly without a CUDA device.
#include "plan.h"
#include "mg.h"

using namespace b200cg;

// ------------------------------------------------------------------------------------------- errors
static thread_local std::string g_last_error;

int b200cg_fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

double b200cg::now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

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
  const int strip_out = tt->strip_out, shift = tt->col_shift;
  const int strips = (g.n - 1) / strip_out + 1;
  const int yB0 = g.ylo, yB1 = g.ysplit ? std::max(g.ylo, std::min(g.yhi, g.ysplit + 1)) : g.ylo;
  const int yU0 = std::max(g.ylo, g.ysplit + 1), yU1 = std::max(yU0, g.yhi);
  if (yB1 > yB0)
    for (int s = (g.xsplit + 1) / strip_out; s < strips; ++s) cols.push_back({s * strip_out + shift, yB0, yB1, g.xsplit + 1});
  if (yU1 > yU0)
    for (int s = 0; s < strips; ++s) cols.push_back({s * strip_out + shift, yU0, yU1, 1});
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
    tt->tile_capacity = tiles.size() + 4 * ((size_t)(P->g.n - 1) / tt->strip_out + 2) + 64;  // + strip-end splits
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
int b200cg::rebalance_tiles(b200cg_plan_s* P, int flavour) {
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
  // consumed: a later solve that does not launch this flavour must not re-apply these stamps
  CU(cudaMemset(P->d_clock[flavour], 0, clk.size() * sizeof(unsigned long long)));
  return upload_tiles(P, tt);
}

// Cached graphs hold the device pointers of the buffers they were captured with: whoever frees such buffers drops the
// graphs that captured them.
void b200cg::drop_graphs(b200cg_plan_s* P, int variant_mask) {
  for (auto it = P->graphs.begin(); it != P->graphs.end();) {
    if ((it->first / 4096) & variant_mask) {
      if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
      it = P->graphs.erase(it);
    } else {
      ++it;
    }
  }
}

int b200cg::ew_grid(const b200cg_plan_s* P, long long work_items) {
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
  if (all_ok && getenv("B200CG_PEER_TRACE") && atoi(getenv("B200CG_PEER_TRACE")) != 0) {
    CU(cudaMalloc(&P->d_peer_trace, sizeof(unsigned long long) * 4 * PEER_TRACE_CAP));
    CU(cudaMemset(P->d_peer_trace, 0, sizeof(unsigned long long) * 4 * PEER_TRACE_CAP));
  }
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
  cudaFree(P->d_peer_trace);
  comm_destroy(&P->comm);
  csr_free(&P->csr);
  mg_free(P);
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
  if (P->batch) {
    BatchIo* io = P->batch;
    if (io->s_in) cudaStreamDestroy(io->s_in);
    if (io->s_out) cudaStreamDestroy(io->s_out);
    cudaFree(io->stage_out);
    for (cudaEvent_t e : {io->h2d_done, io->in_free, io->gathered, io->d2h_done[0], io->d2h_done[1]})
      if (e) cudaEventDestroy(e);
    delete io;
  }
  cudaFree(P->d_state);
  cudaFree(P->d_log);
  cudaFree(P->d_partials);
  for (auto& tt : P->tile_tab) {
    cudaFree(tt.d_tiles);
    cudaFree(tt.d_cta_begin);
  }
  for (int k = 0; k < 2; ++k) {
    if (P->h_state_m[k]) cudaFreeHost(P->h_state_m[k]);
    if (P->h_log_m[k]) cudaFreeHost(P->h_log_m[k]);
    if (P->ev_launch[k]) cudaEventDestroy(P->ev_launch[k]);
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
    // + two rows behind the stored ones: the second halo rows (ylo-2, yhi+1) of the sharded single-sweep iteration
    P->vec_elems = (size_t)(g.yrows + 2) * (size_t)g.pitch;
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
  for (int k = 0; k < 2; ++k) {
    CU(cudaHostAlloc(&P->h_state_m[k], sizeof(DevState), cudaHostAllocDefault));
    CU(cudaHostAlloc(&P->h_log_m[k], sizeof(CbRecord) * CB_LOG_CAP, cudaHostAllocDefault));
    CU(cudaEventCreateWithFlags(&P->ev_launch[k], cudaEventDisableTiming));
  }
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
    // single-sweep geometry: one 15-warp CTA per SM on 840-column strips for large slabs (half as many strip edges per
    // byte: +1.8 % at 16384^2, +1.5 % at 4096^2), two 8-warp CTAs per SM on 420-column strips below ~4 M unknowns, where
    // the number of CTAs matters more (1024^2: 49.7 against 44.3 GDOF-it/s) - profiles/r2_single_sweep.md
    P->fused_cw = (!P->generic && P->g.hi - P->g.lo >= (4LL << 20)) ? 14 : FUSED_CW;
    if (const char* env = getenv("B200CG_FUSED_CW")) P->fused_cw = atoi(env) == 14 ? 14 : FUSED_CW;
    P->x_deferral = env_int("B200CG_XDEFER", 1) != 0;
    P->balance_rounds = env_int("B200CG_BALANCE", 4);
    P->balance_rounds_fused = P->balance_rounds;
    P->cluster_enabled = env_int("B200CG_CLUSTER", 1) != 0;
    P->single_sweep_default = env_int("B200CG_SINGLE_SWEEP", 1) != 0;
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
    P->tile_tab[3].ctas_per_sm = P->fused_cw == FUSED_CW ? 2 : 1;  // single-sweep iteration: its own strip geometry (fused_kernel.cuh)
    P->tile_tab[3].strip_out = fused_strip_out(P->fused_cw);
    P->tile_tab[3].col_shift = FUSED_COL_SHIFT;
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

extern "C" int b200cg_work_split(const b200cg_plan_desc* desc, int sms, int ctas_per_sm, const double* weights,
                                 int n_weights, int* tiles_out, int64_t capacity, int64_t* n_tiles, int* cta_begin_out,
                                 int* grid_out) {
  if (!desc || !n_tiles || !grid_out) return fail(B200CG_ERR_INVALID_ARG, "desc/n_tiles/grid is NULL");
  if (sms < 1 || ctas_per_sm < 1) return fail(B200CG_ERR_INVALID_ARG, "sms and ctas_per_sm must be positive");
  if (desc->domain == B200CG_DOMAIN_GENERIC) return fail(B200CG_ERR_UNSUPPORTED, "a generic (CSR-only) plan has no sweep tiles");
  b200cg_plan_s tmp;
  tmp.desc = *desc;
  RET(setup_geometry(&tmp));
  tmp.sms = sms;
  TileTable tt;
  tt.ctas_per_sm = ctas_per_sm;
  if (desc->reserved0 == 1 || desc->reserved0 == 2) {  // the single-sweep kernel's strip geometries (2: the wide one)
    tt.strip_out = fused_strip_out(desc->reserved0 == 2 ? 14 : FUSED_CW);
    tt.col_shift = FUSED_COL_SHIFT;
  }
  if (weights && n_weights > 0) tt.weight.assign(weights, weights + n_weights);
  std::vector<Tile> tiles;
  std::vector<int> cta_begin;
  build_tiles(&tmp, &tt, &tiles, &cta_begin);
  *n_tiles = (int64_t)tiles.size();
  *grid_out = tt.grid;
  if (tiles_out) {
    if ((int64_t)tiles.size() > capacity) return fail(B200CG_ERR_INVALID_ARG, "%zu tiles do not fit capacity %lld", tiles.size(), (long long)capacity);
    for (size_t i = 0; i < tiles.size(); ++i) {
      tiles_out[4 * i + 0] = tiles[i].col0;
      tiles_out[4 * i + 1] = tiles[i].ya;
      tiles_out[4 * i + 2] = tiles[i].yb;
      tiles_out[4 * i + 3] = tiles[i].xlo;
    }
  }
  if (cta_begin_out) std::copy(cta_begin.begin(), cta_begin.end(), cta_begin_out);
  return B200CG_OK;
}

// ------------------------------------------------------------------------------------------- data movement
long long b200cg::local_count(const b200cg_plan_s* P) { return P->g.hi - P->g.lo; }

int b200cg::upload_vector(b200cg_plan_s* P, const double* host, double* pitched) {
  const long long cnt = local_count(P);
  CU(cudaMemcpyAsync(P->compact, host, cnt * sizeof(double), cudaMemcpyHostToDevice, P->stream));
  scatter_compact_kernel<<<ew_grid(P, cnt), CTA_THREADS, 0, P->stream>>>(P->compact, pitched, P->g);
  CU(cudaGetLastError());
  return B200CG_OK;
}
int b200cg::download_vector(b200cg_plan_s* P, const double* pitched, double* host) {
  const long long cnt = local_count(P);
  gather_compact_kernel<<<ew_grid(P, cnt), CTA_THREADS, 0, P->stream>>>(pitched, P->compact, P->g);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(host, P->compact, cnt * sizeof(double), cudaMemcpyDeviceToHost, P->stream));
  return B200CG_OK;
}
int b200cg::ensure_u(b200cg_plan_s* P) {
  if (!P->u) RET(alloc_vec(P, &P->u));
  return B200CG_OK;
}
int b200cg::ensure_scratch(b200cg_plan_s* P) {
  if (!P->va) RET(alloc_vec(P, &P->va));
  if (!P->vb) RET(alloc_vec(P, &P->vb));
  return B200CG_OK;
}
// one-row halo exchange of a pitched vector with the slab neighbours (no-op on a single GPU)
int b200cg::exchange_halo(b200cg_plan_s* P, double* v) {
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

int b200cg::exchange_halo2(b200cg_plan_s* P, double* v0, double* v1) {
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

// ------------------------------------------------------------------------------------------- CSR entry points
extern "C" int b200cg_set_csr(b200cg_plan_t P, int64_t nrows, int64_t nnz, const int* row_map, const int* entries,
                              const double* values) {
  if (!P || !row_map || !entries || !values) return fail(B200CG_ERR_INVALID_ARG, "NULL argument");
  if (P->desc.world > 1) return fail(B200CG_ERR_UNSUPPORTED, "the CSR comparison path is single-GPU");
  if (nrows != P->n_global) return fail(B200CG_ERR_INVALID_ARG, "nrows %lld != unknowns %lld", (long long)nrows, (long long)P->n_global);
  CU(cudaSetDevice(P->desc.device));
  std::string err;
  bool reallocated = false;
  int rc = csr_upload(&P->csr, nrows, nnz, row_map, entries, values, P->stream, &err, &reallocated);
  if (reallocated) {
    drop_graphs(P, V_CSR);  // they captured the buffers csr_upload freed
    if (P->solution_in_csr) P->have_solution = P->solution_in_csr = false;  // the solution lived in the freed vectors
  }
  if (rc) return fail(rc, "%s", err.c_str());
  return B200CG_OK;
}
extern "C" int b200cg_assemble_csr(b200cg_plan_t P, int64_t* nnz) {
  if (!P) return fail(B200CG_ERR_INVALID_ARG, "plan is NULL");
  NEED_GEOMETRY(P);
  if (P->desc.world > 1) return fail(B200CG_ERR_UNSUPPORTED, "the CSR comparison path is single-GPU");
  CU(cudaSetDevice(P->desc.device));
  drop_graphs(P, V_CSR);  // they captured the buffers csr_assemble frees
  if (P->solution_in_csr) P->have_solution = P->solution_in_csr = false;  // the solution lived in the freed vectors
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

extern "C" int b200cg_peer_trace(b200cg_plan_t P, uint64_t* out, int capacity, int* n_iterations) {
  if (!P || !out || !n_iterations) return fail(B200CG_ERR_INVALID_ARG, "plan/out/n_iterations is NULL");
  *n_iterations = 0;
  if (!P->d_peer_trace) return fail(B200CG_ERR_STATE, "no peer-exchange trace: create the sharded plan with B200CG_PEER_TRACE=1");
  const int n = std::min(capacity, PEER_TRACE_CAP);
  CU(cudaSetDevice(P->desc.device));
  CU(cudaMemcpy(out, P->d_peer_trace, sizeof(unsigned long long) * 4 * (size_t)n, cudaMemcpyDeviceToHost));
  *n_iterations = n;
  return B200CG_OK;
}

extern "C" int b200cg_cta_times(b200cg_plan_t P, int flavour, uint64_t* out, int capacity, int* n_ctas) {
  if (!P || !out || !n_ctas) return fail(B200CG_ERR_INVALID_ARG, "plan/out/n_ctas is NULL");
  if (flavour < 0 || flavour > 3) return fail(B200CG_ERR_INVALID_ARG, "flavour %d outside [0, 3]", flavour);
  NEED_GEOMETRY(P);
  const int n = std::min(P->clock_ctas[flavour], capacity);
  CU(cudaSetDevice(P->desc.device));
  CU(cudaMemcpy(out, P->d_clock[flavour], sizeof(unsigned long long) * 2 * (size_t)n, cudaMemcpyDeviceToHost));
  *n_ctas = n;
  return B200CG_OK;
}

