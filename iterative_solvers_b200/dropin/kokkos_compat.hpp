// Host-side stand-ins for the few Kokkos / KokkosKernels names that appear in the reference's public headers
// (solver.hpp:12-15: KokkosVector, KokkosCrsMatrix) and in its one consumer (qt_gui/src/mainwindow.cpp:81-82,
// :166-167: Kokkos::initialize / is_initialized / finalize). The B200 build does not depend on Kokkos: these are
// plain reference-counted host buffers; all arithmetic happens on the GPU behind the C ABI (include/b200cg.h).
#pragma once

#include <cstddef>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#ifndef KOKKOS_LAMBDA
#define KOKKOS_LAMBDA [=]
#endif

namespace Kokkos {

struct HostSpace {};
struct B200Device {};
using DefaultExecutionSpace = B200Device;

namespace b200_detail {
inline bool& runtime_flag() {
  static bool on = false;
  return on;
}
}  // namespace b200_detail

// Lazy, idempotent runtime switch (grid_system.cpp:304-306 calls it on first use).
inline void initialize() { b200_detail::runtime_flag() = true; }
inline void initialize(int&, char**) { b200_detail::runtime_flag() = true; }
inline bool is_initialized() { return b200_detail::runtime_flag(); }
inline void finalize() { b200_detail::runtime_flag() = false; }

// Rank-1 host array with shared ownership: copies are shallow, element access works through a const handle,
// a labelled constructor zero-fills - the behaviours the reference's callers rely on.
template <class DataType, class... Properties>
class View {
 public:
  using value_type = typename std::remove_pointer<DataType>::type;

  View() = default;
  View(const std::string& label, std::size_t n) : label_(label), store_(std::make_shared<std::vector<value_type>>(n)) {}

  std::size_t extent(int dim) const { return (dim == 0 && store_) ? store_->size() : (dim == 0 ? 0 : 1); }
  std::size_t size() const { return extent(0); }
  value_type& operator()(std::size_t i) const { return (*store_)[i]; }
  value_type& operator[](std::size_t i) const { return (*store_)[i]; }
  value_type* data() const { return store_ ? store_->data() : nullptr; }
  const std::string& label() const { return label_; }
  bool is_allocated() const { return static_cast<bool>(store_); }

 private:
  std::string label_;
  std::shared_ptr<std::vector<value_type>> store_;
};

template <class ViewType>
inline ViewType create_mirror_view(const ViewType& v) {
  return v;  // everything already lives on the host
}

template <class D, class... P, class S, class... Q>
inline void deep_copy(const View<D, P...>& dst, const View<S, Q...>& src) {
  if (static_cast<const void*>(dst.data()) == static_cast<const void*>(src.data())) return;
  const std::size_t n = dst.extent(0) < src.extent(0) ? dst.extent(0) : src.extent(0);
  for (std::size_t i = 0; i < n; ++i) dst(i) = src(i);
}
template <class D, class... P>
inline void deep_copy(const View<D, P...>& dst, const typename View<D, P...>::value_type& value) {
  for (std::size_t i = 0; i < dst.extent(0); ++i) dst(i) = value;
}

template <class... Traits>
struct RangePolicy {
  RangePolicy(long first, long last) : first_(first), last_(last) {}
  long first_, last_;
};

// Host loop for callers that still post-process with parallel_for; the solver itself never uses it.
template <class Policy, class Functor>
inline void parallel_for(const Policy& policy, const Functor& body) {
  for (long i = policy.first_; i < policy.last_; ++i) body(static_cast<int>(i));
}
template <class Policy, class Functor>
inline void parallel_for(const std::string&, const Policy& policy, const Functor& body) {
  parallel_for(policy, body);
}
inline void fence() {}

}  // namespace Kokkos

namespace KokkosSparse {

// CSR container with the member names the reference reads (grid_system.cpp:148, dirichlet_solver.cpp:415-429):
// graph.row_map, graph.entries, values, numRows(), numCols(), nnz().
template <class Scalar, class Ordinal, class Device, class MemoryTraits, class SizeType>
class CrsMatrix {
 public:
  using values_type = Kokkos::View<Scalar*, Kokkos::HostSpace>;
  using row_map_type = Kokkos::View<SizeType*, Kokkos::HostSpace>;
  using index_type = Kokkos::View<Ordinal*, Kokkos::HostSpace>;
  struct Graph {
    row_map_type row_map;
    index_type entries;
  };

  CrsMatrix() = default;
  CrsMatrix(const std::string& label, Ordinal rows, Ordinal cols, SizeType nonzeros, const values_type& vals,
            const row_map_type& rows_view, const index_type& cols_view)
      : values(vals), label_(label), rows_(rows), cols_(cols), nnz_(nonzeros) {
    graph.row_map = rows_view;
    graph.entries = cols_view;
  }

  Ordinal numRows() const { return rows_; }
  Ordinal numCols() const { return cols_; }
  SizeType nnz() const { return nnz_; }

  Graph graph;
  values_type values;

 private:
  std::string label_;
  Ordinal rows_ = 0, cols_ = 0;
  SizeType nnz_ = 0;
};

}  // namespace KokkosSparse
