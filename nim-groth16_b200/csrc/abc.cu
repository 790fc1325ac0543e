// Sparse evaluation of A.w and B.w from the zkey coefficient list, on sm_100a.
//
// Replaces groth16/prover.nim:56-73 (buildABC): Az[row] += coeff * w[col] for matrix-A entries, the
// same for B, then Cz = Az o Bz (C.w is never evaluated; a matrix-C entry is an error, prover.nim:67).
// The reference scatters sequentially over the coefficient list; here the list is sorted once per
// zkey into rows (CSR) and every proof runs one thread per row: sums in Fr are exact, so the result
// is bit-identical regardless of the summation order.
#include <cub/device/device_radix_sort.cuh>
#include "abc.cuh"

namespace g16 {

static __device__ __forceinline__ Fr ld_fr32(const void* p) {   // 4-byte aligned source
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = q[i];
  return r;
}
static __device__ __forceinline__ Fr ld_fr16(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
static __device__ __forceinline__ void st_fr16(Fr* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// record layout helpers
static __device__ __forceinline__ const uint32_t* rec_ptr(const void* base, size_t i, int format) {
  size_t stride = (format == COEFF_PACKED44_R2) ? 44 : 48;
  return reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(base) + i * stride);
}

__global__ void k_coeff_keys(const void* recs, uint32_t nnz, int format, uint32_t n, uint32_t nvars, uint32_t* keys,
                             uint32_t* idx, int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const uint32_t* r = rec_ptr(recs, i, format);
  uint32_t m = r[0], row = r[1], col = r[2];
  if (m > 1) atomicOr(err, m == 2 ? 1 : 2);        // 1: matrix C (prover.nim:67), 2: invalid selector
  if (row >= n) atomicOr(err, 4);                   // zkey.nim:186
  if (col >= nvars) atomicOr(err, 8);               // zkey.nim:187
  keys[i] = (m > 1 || row >= n) ? 0u : m * n + row;
  idx[i] = i;
}

__global__ void k_coeff_gather(const void* recs, const uint32_t* perm, uint32_t nnz, int format, uint32_t* cols,
                               Fr* vals) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  uint32_t i = perm[j];
  const uint32_t* r = rec_ptr(recs, i, format);
  cols[j] = r[2];
  Fr v = ld_fr32(r + (format == COEFF_PACKED44_R2 ? 3 : 4));
  if (format == COEFF_STRUCT48_MONT) v = fmul(v, Fr::rsquared());   // c*R -> c*R^2 (io.nim:134-139 encoding)
  st_fr16(vals + j, v);
}

__global__ void k_coo_gather(const uint32_t* other_in, const Fr* vals_in, const uint32_t* perm, uint32_t nnz,
                             uint32_t* other, Fr* vals) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  uint32_t i = perm[j];
  other[j] = other_in[i];
  st_fr16(vals + j, ld_fr16(vals_in + i));
}

__global__ void k_iota(uint32_t* idx, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = i;
}

__global__ void k_lower_bounds(const uint32_t* keys, uint32_t nnz, uint32_t nkeys, uint32_t* ptr) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nkeys) return;
  uint32_t lo = 0, hi = nnz;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1;
    else hi = mid;
  }
  ptr[k] = lo;
}

static void sort_by_key(DevBuf& keys_sorted, DevBuf& perm_sorted, uint32_t* keys, uint32_t* idx, size_t nnz,
                        size_t nkeys, cudaStream_t stream) {
  keys_sorted.ensure(nnz * 4 + 4);
  perm_sorted.ensure(nnz * 4 + 4);
  int end_bit = 1;
  while (((uint64_t)1 << end_bit) <= (uint64_t)nkeys) end_bit++;
  size_t tmp_bytes = 0;
  DevBuf tmp;
  G16_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_sorted.as<uint32_t>(), idx,
                                           perm_sorted.as<uint32_t>(), (int64_t)nnz, 0, end_bit, stream));
  tmp.ensure(tmp_bytes);
  G16_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys, keys_sorted.as<uint32_t>(), idx,
                                           perm_sorted.as<uint32_t>(), (int64_t)nnz, 0, end_bit, stream));
  G16_CUDA(cudaStreamSynchronize(stream));   // tmp is released on return
}

void coeffs_to_csr(SparseCsr& out, const void* dev_records, size_t nnz, int format, int log_n, size_t nvars,
                   cudaStream_t stream) {
  G16_REQUIRE(format == COEFF_PACKED44_R2 || format == COEFF_STRUCT48_MONT, "unknown coefficient record format");
  G16_REQUIRE(nnz < ((size_t)1 << 31), "too many coefficients");
  size_t n = (size_t)1 << log_n;
  out.nnz = nnz;
  out.nkeys = 2 * n;
  out.ptr.ensure((2 * n + 2) * 4);
  out.other.ensure(nnz * 4 + 4);
  out.vals.ensure(nnz * sizeof(Fr) + sizeof(Fr));
  DevBuf keys, idx, keys_sorted, perm, err;
  keys.ensure(nnz * 4 + 4);
  idx.ensure(nnz * 4 + 4);
  err.ensure(4);
  G16_CUDA(cudaMemsetAsync(err.p, 0, 4, stream));
  if (nnz) {
    k_coeff_keys<<<div_up(nnz, 256), 256, 0, stream>>>(dev_records, (uint32_t)nnz, format, (uint32_t)n,
                                                       (uint32_t)nvars, keys.as<uint32_t>(), idx.as<uint32_t>(),
                                                       err.as<int>());
    G16_LAUNCH_CHECK();
  }
  int herr = 0;
  G16_CUDA(cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, stream));
  G16_CUDA(cudaStreamSynchronize(stream));
  G16_REQUIRE(!(herr & 1), "fatal error: matrix C coefficient in the zkey (prover.nim:67)");
  G16_REQUIRE(!(herr & 2), "invalid matrix selector");
  G16_REQUIRE(!(herr & 4), "row index out of range");
  G16_REQUIRE(!(herr & 8), "column index out of range");
  if (nnz) {
    sort_by_key(keys_sorted, perm, keys.as<uint32_t>(), idx.as<uint32_t>(), nnz, 2 * n, stream);
    k_coeff_gather<<<div_up(nnz, 256), 256, 0, stream>>>(dev_records, perm.as<uint32_t>(), (uint32_t)nnz, format,
                                                         out.other.as<uint32_t>(), out.vals.as<Fr>());
    G16_LAUNCH_CHECK();
  } else {
    keys_sorted.ensure(4);
  }
  k_lower_bounds<<<div_up(2 * n + 1, 256), 256, 0, stream>>>(keys_sorted.as<uint32_t>(), (uint32_t)nnz,
                                                             (uint32_t)(2 * n), out.ptr.as<uint32_t>());
  G16_LAUNCH_CHECK();
  G16_CUDA(cudaStreamSynchronize(stream));
}

void coo_to_csr(SparseCsr& out, const uint32_t* dev_keys, const uint32_t* dev_other, const Fr* dev_vals, size_t nnz,
                size_t nkeys, cudaStream_t stream) {
  out.nnz = nnz;
  out.nkeys = nkeys;
  out.ptr.ensure((nkeys + 2) * 4);
  out.other.ensure(nnz * 4 + 4);
  out.vals.ensure(nnz * sizeof(Fr) + sizeof(Fr));
  DevBuf idx, keys_sorted, perm, keys_copy;
  idx.ensure(nnz * 4 + 4);
  keys_copy.ensure(nnz * 4 + 4);
  if (nnz) {
    G16_CUDA(cudaMemcpyAsync(keys_copy.p, dev_keys, nnz * 4, cudaMemcpyDeviceToDevice, stream));
    k_iota<<<div_up(nnz, 256), 256, 0, stream>>>(idx.as<uint32_t>(), (uint32_t)nnz);
    G16_LAUNCH_CHECK();
    sort_by_key(keys_sorted, perm, keys_copy.as<uint32_t>(), idx.as<uint32_t>(), nnz, nkeys, stream);
    k_coo_gather<<<div_up(nnz, 256), 256, 0, stream>>>(dev_other, dev_vals, perm.as<uint32_t>(), (uint32_t)nnz,
                                                       out.other.as<uint32_t>(), out.vals.as<Fr>());
    G16_LAUNCH_CHECK();
  } else {
    keys_sorted.ensure(4);
  }
  k_lower_bounds<<<div_up(nkeys + 1, 256), 256, 0, stream>>>(keys_sorted.as<uint32_t>(), (uint32_t)nnz,
                                                             (uint32_t)nkeys, out.ptr.as<uint32_t>());
  G16_LAUNCH_CHECK();
  G16_CUDA(cudaStreamSynchronize(stream));
}

// one thread per constraint row: Az[i], Bz[i], Cz[i]
__global__ void k_build_abc(const uint32_t* __restrict__ ptr, const uint32_t* __restrict__ cols,
                            const Fr* __restrict__ vals, const Fr* __restrict__ w, Fr* __restrict__ abc, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr a = Fr::zero(), b = Fr::zero();
  for (uint32_t j = ptr[i]; j < ptr[i + 1]; j++) a = fadd(a, fmul(ld_fr16(vals + j), ld_fr16(w + cols[j])));
  for (uint32_t j = ptr[n + i]; j < ptr[n + i + 1]; j++) b = fadd(b, fmul(ld_fr16(vals + j), ld_fr16(w + cols[j])));
  st_fr16(abc + i, a);
  st_fr16(abc + n + i, b);
  st_fr16(abc + 2 * (size_t)n + i, fmul(a, b));     // prover.nim:69-71
}

void build_abc(const SparseCsr& csr, const Fr* witness_std, Fr* abc, int log_n, cudaStream_t stream) {
  size_t n = (size_t)1 << log_n;
  G16_REQUIRE(csr.nkeys == 2 * n, "coefficient structure does not match the domain size");
  k_build_abc<<<div_up(n, 128), 128, 0, stream>>>(csr.ptr.as<uint32_t>(), csr.other.as<uint32_t>(), csr.vals.as<Fr>(),
                                                  witness_std, abc, (uint32_t)n);
  G16_LAUNCH_CHECK();
}

__global__ void k_fr_convert(const Fr* in, Fr* out, uint32_t n, int to_m) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Fr x = ld_fr16(in + i);
    st_fr16(out + i, to_m ? to_mont(x) : from_mont(x));
  }
}
static unsigned conv_grid(size_t n) {
  size_t g = (n + 255) / 256;
  return (unsigned)(g < 148 * 16 ? (g ? g : 1) : 148 * 16);
}
void fr_from_mont(const Fr* in, Fr* out, size_t n, cudaStream_t stream) {
  if (!n) return;
  k_fr_convert<<<conv_grid(n), 256, 0, stream>>>(in, out, (uint32_t)n, 0);
  G16_LAUNCH_CHECK();
}
// standard-form integers that are not below r (a malformed .wtns; io.nim:141-145 fromBig reduces them) are
// reduced in place; canonical values are only read.  Keeps buildABC (which reduces) and the MSM digits (which
// assume < 2^254) consistent.
__global__ void k_fr_reduce_std(Fr* x, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Fr v = ld_fr16(x + i);
    bool ge = true;                                   // v >= r ?
#pragma unroll
    for (int k = 7; k >= 0; k--) {
      const uint32_t m = FrParams::mod(k);
      if (v.v[k] != m) {
        ge = v.v[k] > m;
        break;
      }
    }
    if (ge) st_fr16(x + i, from_mont(to_mont(v)));
  }
}
void fr_reduce_std(Fr* x, size_t n, cudaStream_t stream) {
  if (!n) return;
  k_fr_reduce_std<<<conv_grid(n), 256, 0, stream>>>(x, (uint32_t)n);
  G16_LAUNCH_CHECK();
}
void fr_to_mont(const Fr* in, Fr* out, size_t n, cudaStream_t stream) {
  if (!n) return;
  k_fr_convert<<<conv_grid(n), 256, 0, stream>>>(in, out, (uint32_t)n, 1);
  G16_LAUNCH_CHECK();
}

}  // namespace g16
