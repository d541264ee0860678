// Assembled-matrix comparison path: CSR SpMV CG on compact vectors, sharing the device-side
// scalar state, stop rules and graph loop of the matrix-free path.
//   csr_spmv_kernel<1> : z = r + beta*z_old (stored), Az = A z with z gathered as r[c] + beta*z_old[c]
//                        on the fly, reduces Az.z and r.z        (msg_solver.cpp:93-99,167-169)
//   csr_update_kernel  : x += alpha z, r -= alpha Az, reduces r.r, |r|_inf, |dx|_inf(, |x-u|_inf)
//                        (msg_solver.cpp:105-139)
// Row sums accumulate values[k]*x[entries[k]] in stored order with separately rounded multiply and add, as the
// serial KokkosSparse::spmv restatement does (oracle/shim/KokkosSparse_spmv.hpp).
#pragma once
#include <string>
#include <cub/device/device_scan.cuh>
#include "kernels.cuh"

namespace b200cg {

struct CsrData {
  long long nrows = 0, nnz = 0;
  int* row_map = nullptr;
  int* entries = nullptr;
  double* values = nullptr;
  // CG vectors in the compact ordering
  double* x = nullptr;
  double* r = nullptr;
  double* z[2] = {nullptr, nullptr};
  double* Az = nullptr;
  double* b = nullptr;
  double* u = nullptr;
  bool has_u = false;
};

struct CsrArgs {
  long long nrows;
  const int* row_map;
  const int* entries;
  const double* values;
  double* x;
  double* r;
  const double* z_old;
  double* z_new;
  double* Az;
  const double* b;
  const double* u;
  DevState* st;
  double* partials;
  CbRecord* cb_log;
  const int* stop_flag;  // mapped host flag (requestStop), see poll_stop; null outside the solve loop
};

static inline CsrArgs csr_args(CsrData* c, DevState* st, double* partials, CbRecord* log, int par) {
  CsrArgs a;
  a.nrows = c->nrows;
  a.row_map = c->row_map;
  a.entries = c->entries;
  a.values = c->values;
  a.x = c->x;
  a.r = c->r;
  a.z_old = c->z[par];
  a.z_new = c->z[par ^ 1];
  a.Az = c->Az;
  a.b = c->b;
  a.u = c->has_u ? c->u : nullptr;
  a.st = st;
  a.partials = partials;
  a.cb_log = log;
  a.stop_flag = nullptr;
  return a;
}

static inline int csr_grid(long long n, int sms) {
  long long blocks = (n + CTA_THREADS - 1) / CTA_THREADS;
  long long cap = (long long)sms * 16;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

static inline void csr_free(CsrData* c) {
  cudaFree(c->row_map);
  cudaFree(c->entries);
  cudaFree(c->values);
  cudaFree(c->x);
  cudaFree(c->r);
  cudaFree(c->z[0]);
  cudaFree(c->z[1]);
  cudaFree(c->Az);
  cudaFree(c->b);
  cudaFree(c->u);
  *c = CsrData();
}

#define CSR_CU(call)                                                                         \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      *err = std::string(#call) + " failed: " + cudaGetErrorString(e__);                     \
      return 3;                                                                              \
    }                                                                                        \
  } while (0)

static inline int csr_ensure_vectors(CsrData* c, cudaStream_t s, std::string* err) {
  const size_t bytes = (size_t)(c->nrows > 0 ? c->nrows : 1) * sizeof(double);
  double** vs[] = {&c->x, &c->r, &c->z[0], &c->z[1], &c->Az, &c->b, &c->u};
  for (double** v : vs) {
    if (!*v) {
      CSR_CU(cudaMalloc(v, bytes));
      CSR_CU(cudaMemsetAsync(*v, 0, bytes, s));
    }
  }
  return 0;
}

// A matrix of the shape already resident is copied into the existing buffers (*reallocated = false: graphs captured
// with these pointers stay valid); any other shape frees and reallocates everything.
static inline int csr_upload(CsrData* c, long long nrows, long long nnz, const int* row_map, const int* entries,
                             const double* values, cudaStream_t s, std::string* err, bool* reallocated) {
  *reallocated = !(c->row_map && c->nrows == nrows && c->nnz == nnz);
  if (*reallocated) {
    csr_free(c);
    c->nrows = nrows;
    c->nnz = nnz;
    CSR_CU(cudaMalloc(&c->row_map, (size_t)(nrows + 1) * sizeof(int)));
    CSR_CU(cudaMalloc(&c->entries, (size_t)(nnz > 0 ? nnz : 1) * sizeof(int)));
    CSR_CU(cudaMalloc(&c->values, (size_t)(nnz > 0 ? nnz : 1) * sizeof(double)));
  }
  CSR_CU(cudaMemcpyAsync(c->row_map, row_map, (size_t)(nrows + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
  CSR_CU(cudaMemcpyAsync(c->entries, entries, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, s));
  CSR_CU(cudaMemcpyAsync(c->values, values, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, s));
  CSR_CU(cudaStreamSynchronize(s));
  return 0;
}

// --------------------------------------------------------------------------------------------- assembly
// GridSystem::initiate_matrix on the device (grid_system.cpp:157-274): per row diag, left, right, top, bottom;
// a neighbour is present iff it is an unknown (a dropped neighbour is a Dirichlet node).
__device__ __forceinline__ bool is_unknown(const Geom& g, int x, int y) {
  if (x < 1 || x > g.n - 1 || y < 1 || y > g.m - 1) return false;
  if (g.ysplit && y <= g.ysplit) return x > g.xsplit;
  return true;
}
__device__ __forceinline__ long long compact_index(const Geom& g, int x, int y) {
  if (g.ysplit && y <= g.ysplit) return (long long)(y - 1) * g.wB + (x - g.xsplit - 1);
  return g.NB + (long long)(y - g.ysplit - 1) * g.wU + (x - 1);
}

static __global__ void csr_count_kernel(int* __restrict__ counts, const Geom g, long long nrows) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= nrows;
       i += (long long)gridDim.x * blockDim.x) {
    int c = 0;
    if (i < nrows) {
      int x, y;
      decode_compact(g, i, x, y);
      c = 1 + is_unknown(g, x - 1, y) + is_unknown(g, x + 1, y) + is_unknown(g, x, y + 1) + is_unknown(g, x, y - 1);
    }
    counts[i] = c;
  }
}

static __global__ void csr_fill_kernel(const int* __restrict__ row_map, int* __restrict__ entries,
                                double* __restrict__ values, const Geom g, long long nrows) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nrows;
       i += (long long)gridDim.x * blockDim.x) {
    int x, y;
    decode_compact(g, i, x, y);
    int k = row_map[i];
    entries[k] = (int)i; values[k++] = g.A;
    if (is_unknown(g, x - 1, y)) { entries[k] = (int)compact_index(g, x - 1, y); values[k++] = g.xk; }
    if (is_unknown(g, x + 1, y)) { entries[k] = (int)compact_index(g, x + 1, y); values[k++] = g.xk; }
    if (is_unknown(g, x, y + 1)) { entries[k] = (int)compact_index(g, x, y + 1); values[k++] = g.yk; }
    if (is_unknown(g, x, y - 1)) { entries[k] = (int)compact_index(g, x, y - 1); values[k++] = g.yk; }
  }
}

static inline int csr_assemble(CsrData* c, const Geom& g, long long nrows, int sms, cudaStream_t s,
                               std::string* err) {
  if (nrows * 5 > 2147483647LL) {
    *err = "CSR assembly: more than 2^31-1 non-zeros do not fit the reference's int32 row_map";
    return 6;
  }
  csr_free(c);
  c->nrows = nrows;
  CSR_CU(cudaMalloc(&c->row_map, (size_t)(nrows + 1) * sizeof(int)));
  int* counts = nullptr;
  CSR_CU(cudaMalloc(&counts, (size_t)(nrows + 1) * sizeof(int)));
  csr_count_kernel<<<csr_grid(nrows + 1, sms), CTA_THREADS, 0, s>>>(counts, g, nrows);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, c->row_map, (int)(nrows + 1), s);
  void* tmp = nullptr;
  CSR_CU(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
  CSR_CU(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, c->row_map, (int)(nrows + 1), s));
  int nnz = 0;
  CSR_CU(cudaMemcpyAsync(&nnz, c->row_map + nrows, sizeof(int), cudaMemcpyDeviceToHost, s));
  CSR_CU(cudaStreamSynchronize(s));
  cudaFree(tmp);
  cudaFree(counts);
  c->nnz = nnz;
  CSR_CU(cudaMalloc(&c->entries, (size_t)(nnz > 0 ? nnz : 1) * sizeof(int)));
  CSR_CU(cudaMalloc(&c->values, (size_t)(nnz > 0 ? nnz : 1) * sizeof(double)));
  csr_fill_kernel<<<csr_grid(nrows, sms), CTA_THREADS, 0, s>>>(c->row_map, c->entries, c->values, g, nrows);
  CSR_CU(cudaGetLastError());
  CSR_CU(cudaStreamSynchronize(s));
  return 0;
}

// --------------------------------------------------------------------------------------------- CG kernels
static __global__ void __launch_bounds__(CTA_THREADS) csr_init_kernel(const CsrArgs a) {
  __shared__ double scratch[3 * 32];
  double s[1] = {0.0}, mx[2] = {0.0, 0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nrows;
       i += (long long)gridDim.x * blockDim.x) {
    const double bv = a.b[i];
    a.r[i] = bv;           // msg_solver.cpp:36
    a.x[i] = 0.0;          // msg_solver.cpp:33
    ((double*)a.z_old)[i] = 0.0;  // first direction: z = r + 0 * 0
    s[0] = fma(bv, bv, s[0]);
    mx[0] = fmax(mx[0], fabs(bv));
    if (a.u) mx[1] = fmax(mx[1], fabs(a.u[i]));
  }
  if (!grid_reduce<1, 2>(s, mx, a.partials, a.st, scratch)) return;
  finalize_init(a.st, a.cb_log, s[0], mx[0], mx[1], a.u != nullptr);
}

// CG = 0: Az = A z_old (plain SpMV, b200cg_csr_apply).  CG = 1: the fused direction update + SpMV + dots.
// One thread per row walking its row in global memory. The 40-byte lane stride looks wasteful, but the L1 absorbs it:
// DRAM traffic is 4.87 GB per launch at 8192^2 against 5.23 GB of model bytes and the kernel runs at 0.86 of the measured
// copy peak. A warp-cooperative variant (the warp's 160 non-zeros copied into shared memory with coalesced loads, lanes
// walking their rows there) was measured and was 34 % SLOWER (1.25 ms against 0.93 ms per launch, profiles/r2_csr.md).
template <int CG>
static __global__ void __launch_bounds__(CTA_THREADS) csr_spmv_kernel(const CsrArgs a) {
  __shared__ double scratch[2 * 32];
  DevState* st = a.st;
  double beta = 0.0;
  if (CG) {
    if (st->done) return;
    beta = st->beta;
  }
  double s[2] = {0.0, 0.0}, mx[1] = {0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nrows;
       i += (long long)gridDim.x * blockDim.x) {
    const int k0 = a.row_map[i], k1 = a.row_map[i + 1];
    double sum = 0.0;
    for (int k = k0; k < k1; ++k) {
      const int c = a.entries[k];
      double zc;
      if (CG) zc = __dadd_rn(a.r[c], __dmul_rn(beta, a.z_old[c]));  // z = r + beta z, msg_solver.cpp:167-169
      else zc = a.z_old[c];
      sum = __dadd_rn(sum, __dmul_rn(a.values[k], zc));
    }
    a.Az[i] = sum;  // alpha = 1, beta = 0: y = 1.0 * sum
    if (CG) {
      const double ri = a.r[i];
      const double zi = __dadd_rn(ri, __dmul_rn(beta, a.z_old[i]));
      a.z_new[i] = zi;
      s[0] = fma(sum, zi, s[0]);  // (A z, z), msg_solver.cpp:99
      s[1] = fma(ri, zi, s[1]);   // (r, z),   msg_solver.cpp:96
    }
  }
  if (!CG) return;
  if (!grid_reduce<2, 0>(s, mx, a.partials, st, scratch)) return;
  finalize_dot(st, s[0], s[1]);
}

template <int WITH_U>
static __global__ void __launch_bounds__(CTA_THREADS) csr_update_kernel(const CsrArgs a) {
  __shared__ double scratch[4 * 32];
  DevState* st = a.st;
  if (st->done) return;
  const double alpha = st->alpha;
  double s[1] = {0.0}, mx[3] = {0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nrows;
       i += (long long)gridDim.x * blockDim.x) {
    const double xo = a.x[i];
    const double xn = __dadd_rn(xo, __dmul_rn(alpha, a.z_new[i]));     // msg_solver.cpp:105-107
    const double rn = __dsub_rn(a.r[i], __dmul_rn(alpha, a.Az[i]));    // msg_solver.cpp:110-112
    a.x[i] = xn;
    a.r[i] = rn;
    s[0] = fma(rn, rn, s[0]);
    mx[0] = fmax(mx[0], fabs(rn));
    mx[1] = fmax(mx[1], fabs(__dsub_rn(xn, xo)));                       // msg_solver.cpp:124-129
    if (WITH_U) mx[2] = fmax(mx[2], fabs(__dsub_rn(xn, a.u[i])));       // msg_solver.cpp:132-139
  }
  if (!grid_reduce<1, 3>(s, mx, a.partials, st, scratch)) return;
  const bool stop_req = poll_stop(st, a.stop_flag);
  finalize_update(st, a.cb_log, s[0], mx[0], mx[1], WITH_U ? mx[2] : DBL_MAX, 0.0, 0.0, false);
  apply_stop(st, stop_req);
}

// Az <- A x - b (dirichlet_solver.cpp:147-161)
static __global__ void __launch_bounds__(CTA_THREADS) csr_residual_kernel(const CsrArgs a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nrows;
       i += (long long)gridDim.x * blockDim.x) {
    double sum = 0.0;
    for (int k = a.row_map[i]; k < a.row_map[i + 1]; ++k)
      sum = __dadd_rn(sum, __dmul_rn(a.values[k], a.x[a.entries[k]]));
    a.Az[i] = __dsub_rn(sum, a.b[i]);
  }
}

// error = x - u on compact vectors (dirichlet_solver.cpp:172-174); result in Az
static __global__ void __launch_bounds__(CTA_THREADS) csr_error_kernel(const CsrArgs a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nrows;
       i += (long long)gridDim.x * blockDim.x)
    a.Az[i] = __dsub_rn(a.x[i], a.u[i]);
}

}  // namespace b200cg
