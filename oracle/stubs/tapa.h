// TEST INFRASTRUCTURE (oracle/) -- stub of the one TAPA symbol the reference's host library
// uses (tapa::aligned_allocator, /root/reference/common/include/spmv-helper.h:30), so that
// /root/reference/common/src/spmv-helper.cpp compiles without the TAPA toolchain.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
namespace tapa {
template <typename T>
struct aligned_allocator {
  using value_type = T;
  aligned_allocator() = default;
  template <typename U>
  aligned_allocator(const aligned_allocator<U>&) {}
  T* allocate(std::size_t n) {
    void* p = nullptr;
    if (posix_memalign(&p, 4096, n * sizeof(T) ? n * sizeof(T) : 1)) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, std::size_t) { free(p); }
  template <typename U>
  bool operator==(const aligned_allocator<U>&) const { return true; }
  template <typename U>
  bool operator!=(const aligned_allocator<U>&) const { return false; }
};
}  // namespace tapa
