// TEST INFRASTRUCTURE ONLY (oracle/): restatement of the one KokkosKernels 4.0.01 call the
// reference makes, KokkosSparse::spmv("N", alpha, A, x, beta, y) (msg_solver.cpp:93,236;
// dirichlet_solver.cpp:153). Published algorithm: y_i = beta*y_i + alpha * sum_k values[k]*x[entries[k]]
// over the row's stored entries. Rows are accumulated in stored order (the reference stores
// diag, left, right, top, bottom - grid_system.cpp:193-217), which is also the order
// MatrixFreeSystem::apply uses (matrix_free_system.cpp:216-266). Parity at this boundary is pinned
// independently of Kokkos by check_debug.py:13-36 and py_debug.txt:5,12,14,15.
#pragma once
#include "KokkosSparse_CrsMatrix.hpp"

namespace KokkosSparse {

template <class AMatrix, class XVector, class YVector>
inline void spmv(const char mode[], double alpha, const AMatrix& A, const XVector& x, double beta,
                 const YVector& y) {
  (void)mode;  // the reference only uses "N"
  const long nrows = A.numRows();
  for (long i = 0; i < nrows; ++i) {
    double sum = 0.0;
    for (long k = A.graph.row_map(i); k < A.graph.row_map(i + 1); ++k)
      sum += A.values(k) * x(A.graph.entries(k));
    if (beta == 0.0)
      y(i) = alpha * sum;
    else
      y(i) = beta * y(i) + alpha * sum;
  }
}

}  // namespace KokkosSparse
