// (template implementation, included by msm_g1.cu and msm_g2.cu)
// GPU Pippenger multi-scalar multiplication for BN254 G1 / G2 on sm_100a.
//
// Replaces groth16/bn128/msm.nim:35-59 (msmConstantineG1), :63-83 (msmConstantineG2) and the
// chunk-per-thread driver msm.nim:89-158; the result is the same canonical group element.
//
// Pipeline (DESIGN.md "MSM"):
//   1. k_msm_digits        signed-digit windows of each scalar -> (bucket key, point index|sign)
//   2. radix sort          (key, value) pairs by bucket key (CUB device radix sort, key bits only)
//   3. k_bucket_bounds     first sorted position of every bucket
//   4. k_bucket_accumulate one thread per bucket, XYZZ mixed additions over gathered affine points
//   5. k_bucket_reduce     per window: segmented running sums + block tree reduction
//   6. k_window_combine    Horner over the windows -> one XYZZ point
#include <cub/device/device_radix_sort.cuh>
#pragma once
#include "msm.cuh"
#include "msm_digits.cuh"

namespace g16 {

// ---------------------------------------------------------------------------------------
// vectorised loads / stores of whole structs (sizes are multiples of 16 bytes)
// ---------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T ldg_vec(const T* p) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiple expected");
  T r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(q + i);
  return r;
}
template <class T>
__device__ __forceinline__ T ld_vec(const T* p) {
  T r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = q[i];
  return r;
}
template <class T>
__device__ __forceinline__ void st_vec(T* p, const T& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) q[i] = s[i];
}

// ---------------------------------------------------------------------------------------
// 1. digits
// ---------------------------------------------------------------------------------------
static __global__ void k_msm_digits(const Fr* __restrict__ scalars, uint32_t n, int mont, int c, int nwin, uint32_t nb,
                             uint32_t key_none, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = ld_vec(scalars + i);
  if (mont) s = from_mont(s);                    // msm.nim:44 toBig()
  int carry = 0;
  for (int w = 0; w < nwin; w++) {
    int d = msm_signed_digit(s.v, c, w, nwin, carry);
    uint32_t key = key_none, val = i;
    if (d > 0) key = (uint32_t)w * nb + (uint32_t)(d - 1);
    else if (d < 0) {
      key = (uint32_t)w * nb + (uint32_t)(-d - 1);
      val |= 0x80000000u;
    }
    keys[(size_t)w * n + i] = key;
    vals[(size_t)w * n + i] = val;
  }
}

// ---------------------------------------------------------------------------------------
// 3. bucket boundaries: start[b] = first j with keys[j] >= b
// ---------------------------------------------------------------------------------------
static __global__ void k_bucket_bounds(const uint32_t* __restrict__ keys, size_t m, uint32_t nbuckets,
                                uint32_t* __restrict__ start) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nbuckets) return;
  size_t lo = 0, hi = m;
  while (lo < hi) {
    size_t mid = (lo + hi) >> 1;
    if (keys[mid] < b) lo = mid + 1;
    else hi = mid;
  }
  start[b] = (uint32_t)lo;
}

// ---------------------------------------------------------------------------------------
// 4. bucket accumulation
// ---------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_bucket_accumulate(const uint32_t* __restrict__ vals,
                                                           const uint32_t* __restrict__ start,
                                                           const Affine<F>* __restrict__ points,
                                                           XYZZ<F>* __restrict__ buckets, uint32_t nbuckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  uint32_t j0 = start[b], j1 = start[b + 1];
  XYZZ<F> acc = xyzz_inf<F>();
  for (uint32_t j = j0; j < j1; j++) {
    uint32_t v = vals[j];
    Affine<F> p = ldg_vec(points + (v & 0x7fffffffu));
    if (v & 0x80000000u) p.y = fneg(p.y);
    acc = xyzz_madd(acc, p);
  }
  st_vec(buckets + b, acc);
}

// ---------------------------------------------------------------------------------------
// 5. bucket reduction: window sum = sum_k (k+1) * B_k
//    thread t of a window owns buckets [t*L, (t+1)*L): running sums give
//    S_t = sum B_k and R_t = sum (k - t*L + 1) B_k; its contribution is R_t + (t*L) * S_t.
// ---------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_bucket_reduce(const XYZZ<F>* __restrict__ buckets, uint32_t nb, uint32_t L,
                                                       XYZZ<F>* __restrict__ winpart) {
  extern __shared__ uint4 red_raw[];
  XYZZ<F>* red = reinterpret_cast<XYZZ<F>*>(red_raw);
  const uint32_t w = blockIdx.y;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;     // segment index inside the window
  const XYZZ<F>* B = buckets + (size_t)w * nb + (size_t)t * L;
  XYZZ<F> running = xyzz_inf<F>(), sum = xyzz_inf<F>();
  for (int k = (int)L - 1; k >= 0; k--) {
    XYZZ<F> b = ld_vec(B + k);
    xyzz_add_ni(running, running, b);
    xyzz_add_ni(sum, sum, running);
  }
  if (t) {
    XYZZ<F> off = xyzz_mul_u32(t * L, running);
    xyzz_add_ni(sum, sum, off);
  }
  red[threadIdx.x] = sum;
  __syncthreads();
  for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      XYZZ<F> o = red[threadIdx.x + s];
      xyzz_add_ni(sum, sum, o);
      red[threadIdx.x] = sum;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_vec(winpart + (size_t)w * gridDim.x + blockIdx.x, red[0]);
}

// ---------------------------------------------------------------------------------------
// 6. window combine: result = sum_w 2^(c w) * W_w  (Horner from the top window)
// ---------------------------------------------------------------------------------------
template <class F>
__global__ void k_window_combine(const XYZZ<F>* __restrict__ winpart, uint32_t bpw, int nwin, int c,
                                 XYZZ<F>* __restrict__ result) {
  extern __shared__ uint4 red_raw[];
  XYZZ<F>* win = reinterpret_cast<XYZZ<F>*>(red_raw);
  int w = threadIdx.x;
  if (w < nwin) {
    XYZZ<F> acc = xyzz_inf<F>();
    for (uint32_t i = 0; i < bpw; i++) {
      XYZZ<F> o = ld_vec(winpart + (size_t)w * bpw + i);
      xyzz_add_ni(acc, acc, o);
    }
    win[w] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    XYZZ<F> acc = win[nwin - 1];
    for (int i = nwin - 2; i >= 0; i--) {
      for (int j = 0; j < c; j++) xyzz_dbl_ni(acc, acc);
      XYZZ<F> o = win[i];
      xyzz_add_ni(acc, acc, o);
    }
    st_vec(result, acc);
  }
}

template <class F>
__global__ void k_set_inf(XYZZ<F>* p) {
  if (threadIdx.x == 0 && blockIdx.x == 0) st_vec(p, xyzz_inf<F>());
}

template <class F>
__global__ void k_xyzz_sum_to_affine(const XYZZ<F>* parts, int count, Affine<F>* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  XYZZ<F> acc = xyzz_inf<F>();
  for (int i = 0; i < count; i++) {
    XYZZ<F> o = ld_vec(parts + i);
    xyzz_add_ni(acc, acc, o);
  }
  Affine<F> a;
  xyzz_to_affine_ni(a, acc);
  st_vec(out, a);
}

template <class F>
__global__ void k_affine_sum_to_xyzz(const Affine<F>* parts, int count, XYZZ<F>* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  XYZZ<F> acc = xyzz_inf<F>();
  for (int i = 0; i < count; i++) {                                          // msm.nim:117-119
    Affine<F> o = ld_vec(parts + i);
    xyzz_madd_ni(acc, acc, o);
  }
  st_vec(out, acc);
}

// ---------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------
template <class F>
size_t Msm<F>::workspace_bytes() const {
  return keys_[0].bytes + keys_[1].bytes + vals_[0].bytes + vals_[1].bytes + start_.bytes + buckets_.bytes +
         winpart_.bytes + cub_tmp_.bytes;
}

template <class F>
Msm<F>::~Msm() {
  for (int i = 0; i < 4; i++)
    if (pev_[i]) cudaEventDestroy(pev_[i]);
}
template <class F>
float Msm<F>::last_accum_ms() const {
  float ms = 0.f;
  if (pev_[1] && pev_[2]) cudaEventElapsedTime(&ms, pev_[1], pev_[2]);
  return ms;
}
template <class F>
float Msm<F>::last_total_ms() const {
  float ms = 0.f;
  if (pev_[0] && pev_[3]) cudaEventElapsedTime(&ms, pev_[0], pev_[3]);
  return ms;
}

template <class F>
void Msm<F>::run(const Fr* scalars, bool scalars_mont, const Affine<F>* points, size_t n, XYZZ<F>* result,
                 cudaStream_t stream, const MsmConfig& cfg) {
  G16_REQUIRE(n < ((size_t)1 << 31), "MSM size must be below 2^31");
  if (n == 0) {                                   // zero-length MSM (C1 when nvars = npubs+1)
    k_set_inf<F><<<1, 32, 0, stream>>>(result);
    G16_LAUNCH_CHECK();
    return;
  }
  const int c = cfg.c ? cfg.c : msm_pick_window(n, sizeof(F) > sizeof(Fp));
  G16_REQUIRE(c >= 2 && c <= 22, "MSM window must be 2..22 bits");
  const int nwin = msm_num_windows(c);
  const uint32_t nb = 1u << (c - 1);
  const uint32_t nbuckets = (uint32_t)nwin * nb;
  const size_t m = (size_t)nwin * n;
  G16_REQUIRE(m < ((size_t)1 << 32), "MSM pair count must fit 32 bits");
  last_c = c;
  last_nwin = nwin;
  if (profile) {
    for (int i = 0; i < 4; i++)
      if (!pev_[i]) G16_CUDA(cudaEventCreate(&pev_[i]));
    G16_CUDA(cudaEventRecord(pev_[0], stream));
  }

  keys_[0].ensure(m * 4);
  keys_[1].ensure(m * 4);
  vals_[0].ensure(m * 4);
  vals_[1].ensure(m * 4);
  start_.ensure(((size_t)nbuckets + 2) * 4);
  buckets_.ensure((size_t)nbuckets * sizeof(XYZZ<F>));

  // bucket reduction geometry
  uint32_t nseg = nb < 2048u ? nb : 2048u;       // segments (threads) per window
  uint32_t L = nb / nseg;
  uint32_t tpb = nseg < 128u ? nseg : 128u;
  uint32_t bpw = nseg / tpb;
  winpart_.ensure((size_t)nwin * bpw * sizeof(XYZZ<F>));

  k_msm_digits<<<div_up(n, 256), 256, 0, stream>>>(scalars, (uint32_t)n, scalars_mont ? 1 : 0, c, nwin, nb, nbuckets,
                                                   keys_[0].as<uint32_t>(), vals_[0].as<uint32_t>());
  G16_LAUNCH_CHECK();

  int end_bit = 1;
  while (((uint64_t)1 << end_bit) <= (uint64_t)nbuckets) end_bit++;
  size_t tmp_bytes = 0;
  G16_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_[0].as<uint32_t>(), keys_[1].as<uint32_t>(),
                                           vals_[0].as<uint32_t>(), vals_[1].as<uint32_t>(), (int64_t)m, 0, end_bit,
                                           stream));
  cub_tmp_.ensure(tmp_bytes);
  G16_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp_.p, tmp_bytes, keys_[0].as<uint32_t>(), keys_[1].as<uint32_t>(),
                                           vals_[0].as<uint32_t>(), vals_[1].as<uint32_t>(), (int64_t)m, 0, end_bit,
                                           stream));

  k_bucket_bounds<<<div_up((size_t)nbuckets + 1, 256), 256, 0, stream>>>(keys_[1].as<uint32_t>(), m, nbuckets,
                                                                         start_.as<uint32_t>());
  G16_LAUNCH_CHECK();
  if (profile) G16_CUDA(cudaEventRecord(pev_[1], stream));
  k_bucket_accumulate<F><<<div_up(nbuckets, 128), 128, 0, stream>>>(vals_[1].as<uint32_t>(), start_.as<uint32_t>(),
                                                                    points, buckets_.as<XYZZ<F>>(), nbuckets);
  G16_LAUNCH_CHECK();
  if (profile) G16_CUDA(cudaEventRecord(pev_[2], stream));
  dim3 rgrid(bpw, (unsigned)nwin);
  k_bucket_reduce<F><<<rgrid, tpb, tpb * sizeof(XYZZ<F>), stream>>>(buckets_.as<XYZZ<F>>(), nb, L,
                                                                    winpart_.as<XYZZ<F>>());
  G16_LAUNCH_CHECK();
  int cthreads = ((nwin + 31) / 32) * 32;
  k_window_combine<F><<<1, cthreads, (size_t)nwin * sizeof(XYZZ<F>), stream>>>(winpart_.as<XYZZ<F>>(), bpw, nwin, c,
                                                                               result);
  G16_LAUNCH_CHECK();
  if (profile) {
    G16_CUDA(cudaEventRecord(pev_[3], stream));
    uint32_t pairs = 0;   // start[nbuckets] = number of sorted pairs with a real bucket key
    G16_CUDA(cudaMemcpyAsync(&pairs, start_.as<uint32_t>() + nbuckets, 4, cudaMemcpyDeviceToHost, stream));
    G16_CUDA(cudaStreamSynchronize(stream));
    last_pairs = pairs;
  }
}

template <class F>
void xyzz_sum_to_affine(const XYZZ<F>* parts, int count, Affine<F>* out, cudaStream_t stream) {
  k_xyzz_sum_to_affine<F><<<1, 32, 0, stream>>>(parts, count, out);
  G16_LAUNCH_CHECK();
}
template <class F>
void affine_sum_to_xyzz(const Affine<F>* parts, int count, XYZZ<F>* out, cudaStream_t stream) {
  k_affine_sum_to_xyzz<F><<<1, 32, 0, stream>>>(parts, count, out);
  G16_LAUNCH_CHECK();
}

}  // namespace g16
