// Shared device/host definitions of libb200cg (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200cg {

// ---------------------------------------------------------------------------------------------------
// Tile geometry of the matrix-free kernels. One CTA marches a strip of STRIP_LOAD columns down TILE rows.
// It loads 4 extra columns on each side (one 32-byte sector) so that the horizontal neighbours of every
// output column are produced inside the CTA: no halo threads, no divergent loads.
constexpr int CTA_THREADS = 256;
constexpr int STRIP_LOAD = 2 * CTA_THREADS;      // 512 columns loaded (one double2 per thread per row)
constexpr int STRIP_HALO = 4;                    // over-read columns on each side
constexpr int STRIP_OUT = STRIP_LOAD - 2 * STRIP_HALO;  // 504 columns written
constexpr int XOFF = STRIP_HALO;                 // storage column of grid node x is x + XOFF
constexpr int MAX_PARTIALS = 8;                  // reduction slots per CTA

// Device layout of every CG vector: a pitched copy of the node grid, rows y = ybase .. ybase+yrows-1,
// node (x, y) at (y - ybase) * pitch + x + XOFF. Boundary nodes, the excluded quadrant of the L-shape and
// all padding hold exact zeros and are never written with anything else, so the 5-point stencil needs no
// boundary predicates: a dropped Dirichlet neighbour (matrix_free_system.cpp:221-266) is "+ k * 0.0".
struct Geom {
  int n, m;            // intervals (grid nodes 0..n x 0..m)
  int xsplit, ysplit;  // L-shape: n/2, m/2 (rows y <= ysplit start at x = xsplit+1); RECT: 0, 0
  int pitch;           // doubles per stored row (multiple of 16)
  int ybase, yrows;    // stored rows
  int ylo, yhi;        // unknown rows owned by this rank: [ylo, yhi) within [1, m)
  double A, xk, yk;    // -2(xk+yk), 1/hx^2, 1/hy^2 (grid_system.cpp:316-318)
  double a, c, hx, hy; // node coordinates x = a + i*hx, y = c + j*hy
  // compact (reference) ordering
  long long NB;        // unknowns in block B (global)
  int wB, wU;          // row widths of block B / block U (RECT: wB unused, wU = n-1)
  long long lo, hi;    // compact index range owned by this rank
};

// One unit of sweep work: rows [ya, yb) of one 512-column strip. The host cuts the sweep into equal-work tiles
// and deals them to the resident CTAs (plan.cu: build_tiles); the producer warp just walks its list.
struct Tile {
  int col0;  // storage column of the first loaded column (strip * STRIP_OUT; single-sweep kernel: strip * 480 + 2)
  int ya, yb;
  int xlo;   // first unknown x in these rows (1, or xsplit+1 in block B)
};

// Scalars of the iteration live on the device; the host never sees alpha/beta.
struct DevState {
  double rr;         // r.r of the current residual
  double rz;         // r.p   (MSG flavour numerator, msg_solver.cpp:96)
  double pAp;        // p.Ap
  double alpha, beta;
  double alpha_prev; // x-deferral: alpha of the iteration whose x update is pending
  double r0_norm, r_norm;
  double r_max, dx_max, err_max;
  double dx_l2, err_l2, res_l2;  // MatrixFreeSolver callback quantities (matrix_free_system.cpp:444-463)
  double loc_s[4], loc_m[4];     // sharded plans: this rank's totals, all-reduced before the scalars are formed
  // solver parameters (written by the host before the first graph launch)
  double eps_rel, eps_p, eps_r, eps_e;
  int max_it;
  int rule;          // b200cg_rule
  int has_u;
  int callback_every;
  // progress
  int it;            // completed iterations
  int done;          // 0 = running
  int converged;
  int stop_reason;
  unsigned int ticket;
  unsigned int n_log; // callback records appended so far
  unsigned long long epoch[2];  // peer-memory exchange: publications consumed so far per phase
  int comm_error;     // a peer flag did not arrive in time
  int pad_comm;
  int report_pending; // MatrixFreeSolver callback: the update phase asks the report kernel to run
  int x_pending;      // x-deferral: the last iteration did not touch x; x += alpha_prev * p is owed
};

// ---- NVLink peer-memory exchange of the sharded plans (one process per GPU, buffers shared through CUDA IPC) ----
// Every rank owns one PeerSync block; every rank's sweep kernels write their reduction partials straight into
// all ranks' blocks (remote stores over NVLink) and then raise a flag with the epoch number; the finalize kernel
// of each rank waits for all flags of the epoch and sums the slots in rank order (same result on every rank).
constexpr int PEER_MAX_RANKS = 16;
constexpr int PEER_VALS = 8;  // [s0 s1 s2 s3 | m0 m1 m2 m3]
struct PeerSync {
  double vals[2][PEER_MAX_RANKS][PEER_VALS];        // [phase: 0 = dot, 1 = update][source rank]
  unsigned long long flag[2][PEER_MAX_RANKS];       // epoch of the last complete publication
};
struct PeerLinks {                                  // device-resident table of mapped peer pointers
  PeerSync* sync[PEER_MAX_RANKS];                   // every rank's block (own one included)
  int rank, world;
};

struct CbRecord {
  double it, precision, residual, error;
};

constexpr int CB_LOG_CAP = 1024;  // ring of callback records; >= iterations per graph launch

}  // namespace b200cg
