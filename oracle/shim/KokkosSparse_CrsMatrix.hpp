// TEST INFRASTRUCTURE ONLY (oracle/): serial stand-in for KokkosSparse::CrsMatrix
// (KokkosKernels 4.0.01, pinned at /root/reference/solver/CMakeLists.txt:92, not vendored).
#pragma once
#include "Kokkos_Core.hpp"

namespace KokkosSparse {

template <class Scalar, class Ordinal, class Device, class MemTraits, class SizeType>
class CrsMatrix {
 public:
  using values_type = Kokkos::View<Scalar*, Kokkos::HostSpace>;
  using row_map_type = Kokkos::View<SizeType*, Kokkos::HostSpace>;
  using index_type = Kokkos::View<Ordinal*, Kokkos::HostSpace>;
  struct StaticCrsGraphType {
    row_map_type row_map;
    index_type entries;
  };
  StaticCrsGraphType graph;
  values_type values;

  CrsMatrix() = default;
  CrsMatrix(const std::string&, Ordinal nrows, Ordinal ncols, SizeType nnz, const values_type& vals,
            const row_map_type& rows, const index_type& cols)
      : values(vals), nrows_(nrows), ncols_(ncols), nnz_(nnz) {
    graph.row_map = rows;
    graph.entries = cols;
  }
  Ordinal numRows() const { return nrows_; }
  Ordinal numCols() const { return ncols_; }
  SizeType nnz() const { return nnz_; }

 private:
  Ordinal nrows_ = 0, ncols_ = 0;
  SizeType nnz_ = 0;
};

}  // namespace KokkosSparse
