// TEST INFRASTRUCTURE ONLY (oracle/): a serial, host-only stand-in for the tiny
// subset of Kokkos 4.0.01 that the reference sources use, so that the UNMODIFIED
// files under /root/reference/solver/ compile here (Kokkos is not installed and
// cannot be fetched). It is never included by the product library.
//
// Subset (usage sites in the reference): grid_system.cpp:133-148,277-297,304-305;
// msg_solver.cpp:19-39,67,90-93,105-167,218-253; dirichlet_solver.cpp:19-20,135-176,415-429.
#pragma once
#include <cstddef>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#define KOKKOS_LAMBDA [=]
#define KOKKOS_INLINE_FUNCTION inline

namespace Kokkos {

struct HostSpace {};
struct Serial {};
using DefaultExecutionSpace = Serial;
using DefaultHostExecutionSpace = Serial;

inline bool& shim_initialized_flag() { static bool f = false; return f; }
inline void initialize() { shim_initialized_flag() = true; }
inline void initialize(int&, char**) { shim_initialized_flag() = true; }
inline bool is_initialized() { return shim_initialized_flag(); }
inline void finalize() { shim_initialized_flag() = false; }

// Rank-1 view with shared ownership and shallow copies; writes go through a const handle.
template <class DataType, class... Props>
class View {
 public:
  using value_type = typename std::remove_pointer<DataType>::type;
  View() = default;
  View(const std::string&, std::size_t n) : buf_(std::make_shared<std::vector<value_type>>(n)) {}
  std::size_t extent(int) const { return buf_ ? buf_->size() : 0; }
  std::size_t size() const { return extent(0); }
  value_type& operator()(std::size_t i) const { return (*buf_)[i]; }
  value_type* data() const { return buf_ ? buf_->data() : nullptr; }
 private:
  std::shared_ptr<std::vector<value_type>> buf_;
};

template <class V>
inline V create_mirror_view(const V& v) { return v; }

template <class D, class... P, class S, class... Q>
inline void deep_copy(const View<D, P...>& dst, const View<S, Q...>& src) {
  if (dst.data() == src.data()) return;
  for (std::size_t i = 0; i < dst.extent(0); ++i) dst(i) = src(i);
}
template <class D, class... P>
inline void deep_copy(const View<D, P...>& dst, const typename View<D, P...>::value_type& v) {
  for (std::size_t i = 0; i < dst.extent(0); ++i) dst(i) = v;
}

template <class... Props>
struct RangePolicy {
  long begin_, end_;
  RangePolicy(long b, long e) : begin_(b), end_(e) {}
};

template <class Policy, class F>
inline void parallel_for(const Policy& p, const F& f) {
  for (long i = p.begin_; i < p.end_; ++i) f(static_cast<int>(i));
}
template <class Policy, class F>
inline void parallel_for(const std::string&, const Policy& p, const F& f) { parallel_for(p, f); }

inline void fence() {}

}  // namespace Kokkos
