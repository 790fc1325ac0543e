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
void count_launches(uint64_t n);   // capi.cu: kernels replayed by a CUDA graph launch
uint64_t launches_so_far();        // capi.cu
#define G16_LAUNCH_CHECK()       \
  do {                           \
    g16::count_launch();         \
    G16_CUDA(cudaGetLastError()); \
  } while (0)

// RAII device allocation (grow-only reuse through ensure()).
//
// Memory comes from the device's stream-ordered pool (cudaMallocAsync) with the release threshold lifted, so what a
// destroyed context frees stays cached in the process: a host that creates a context per proof -- the shape of
// the reference's generateProofWithMask(zkey, witness), cli_main.nim:193-210 -- pays for cudaMalloc / cudaFree of
// gigabytes once, not per call.  Semantics stay those of cudaMalloc / cudaFree: an allocation is usable on any
// stream when ensure() returns, and release() waits for the device first (as cudaFree does implicitly).
cudaStream_t pool_stream();            // capi.cu: per-device internal stream; configures the pool on first use
void pool_trim();                      // capi.cu: returns the cached memory of the current device's pool to the driver
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int dev = -1;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      int cur = 0;
      cudaGetDevice(&cur);
      if (dev >= 0 && dev != cur) cudaSetDevice(dev);
      cudaDeviceSynchronize();
      cudaStream_t s = pool_stream();
      if (cudaFreeAsync(p, s) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
      }
      if (dev >= 0 && dev != cur) cudaSetDevice(cur);
    }
    p = nullptr;
    bytes = 0;
  }
  void ensure(size_t n) {
    if (n <= bytes) return;
    release();
    if (n == 0) return;
    G16_CUDA(cudaGetDevice(&dev));
    cudaStream_t s = pool_stream();
    cudaError_t e = cudaMallocAsync(&p, n, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      // the pool could not serve the request (fragmented cache, or memory held by other processes): give the cached
      // memory back to the driver and take a plain allocation
      cudaGetLastError();
      p = nullptr;
      pool_trim();
      e = cudaMalloc(&p, n);
      if (e != cudaSuccess) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        cudaGetLastError();
        throw Error(2, std::string("device allocation of ") + std::to_string(n) + " bytes failed on device " +
                           std::to_string(dev) + " (" + cudaGetErrorString(e) + "; free " + std::to_string(fr) + " of " +
                           std::to_string(tot) + ")");
      }
    }
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
