// Shared host-side plumbing for the library: error reporting and device buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

namespace g16 {

void set_last_error(const std::string& msg);   // capi.cu
const char* get_last_error();

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define G16_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      throw g16::Error(2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                              __FILE__ + ":" + std::to_string(__LINE__));                      \
  } while (0)

#define G16_REQUIRE(cond, msg)                       \
  do {                                               \
    if (!(cond)) throw g16::Error(1, std::string(msg)); \
  } while (0)

// every kernel launch of the library is followed by exactly one G16_LAUNCH_CHECK()
void count_launch();   // capi.cu
#define G16_LAUNCH_CHECK()       \
  do {                           \
    g16::count_launch();         \
    G16_CUDA(cudaGetLastError()); \
  } while (0)

// RAII device allocation (grow-only reuse through ensure()).
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  void ensure(size_t n) {
    if (n <= bytes) return;
    release();
    if (n == 0) return;
    G16_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

static inline int ceil_log2_sz(size_t x) {
  int l = 0;
  while (((size_t)1 << l) < x) l++;
  return l;
}

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace g16
