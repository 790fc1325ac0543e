// (template implementation, included by msm_g1.cu and msm_g2.cu)
// MSM back end on sm_100a: bucket accumulation over length-balanced work items, bucket reduction and the
// final combination, for up to three point sets sharing one MsmSorter run.
//
// Replaces the per-chunk Pippenger the reference delegates to constantine (groth16/bn128/msm.nim:49 / :76)
// and the partial-sum loop of msm.nim:117-119; the result is the same canonical group element.
//
//   k_bucket_accumulate  one thread per work item (<= T mixed additions, items sorted by length), gathered
//                        affine loads (16-byte vectors), XYZZ accumulator in registers
//   k_bucket_fixup       buckets that were split into several items: block-wide tree sum of their partials
//   k_reduce_level       sum_k (k+1) B_k per bucket set, level 1: running sums over 16 consecutive buckets
//   k_reduce_bits/final  the 16-times shorter array left by level 1, summed bit by bit of the index (3 launches
//                        in total instead of a chain of latency-bound levels)
//   k_window_combine     Horner over the windows (plain layout; a single window in the precomputed layout)
//   k_build_table        precomputed layout: 2^(c w) P_i for every window, batch-normalised to affine
#pragma once
#include "msm.cuh"
#include <stdlib.h>
#include "msm_digits.cuh"
#ifdef G16_EXPERIMENTS
#include "experiments/msm_tree.cuh"   // batched-affine accumulation, measured and rejected
#endif
#include "msm_reduce_plan.cuh"

#ifndef G16_G2_MINB_DEFAULT
#define G16_G2_MINB_DEFAULT 3
#endif

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

template <class F>
struct AccSets {
  const Affine<F>* points[MsmAccumulator<F>::MAX_SETS];
  XYZZ<F>* buckets[MsmAccumulator<F>::MAX_SETS];
  XYZZ<F>* partials[MsmAccumulator<F>::MAX_SETS];
  XYZZ<F>* winpart[MsmAccumulator<F>::MAX_SETS];
  XYZZ<F>* result[MsmAccumulator<F>::MAX_SETS];
};

// ---------------------------------------------------------------------------------------
// bucket accumulation
// ---------------------------------------------------------------------------------------
template <class F, int MINB>
__global__ void __launch_bounds__(128, MINB) k_bucket_accumulate(AccSets<F> sets, const uint32_t* __restrict__ vals,
                                                           const uint32_t* __restrict__ start,
                                                           const uint32_t* __restrict__ item_start,
                                                           const uint32_t* __restrict__ item_bucket,
                                                           const uint32_t* __restrict__ items_sorted,
                                                           uint32_t nbuckets, uint32_t max_items, uint32_t T) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= max_items) return;
  uint32_t id = items_sorted[t];
  if (id >= item_start[nbuckets]) return;          // padding
  const int set = blockIdx.y;
  const Affine<F>* __restrict__ points = sets.points[set];
  uint32_t b = item_bucket[id];
  uint32_t i0 = item_start[b];
  uint32_t nitems = item_start[b + 1] - i0;
  uint32_t j0 = start[b] + (id - i0) * T;
  uint32_t j1 = start[b + 1];
  if (j1 > j0 + T) j1 = j0 + T;
  XYZZ<F> acc = xyzz_inf<F>();
  for (uint32_t j = j0; j < j1; j++) {
    uint32_t v = vals[j];
    Affine<F> p = ldg_vec(points + (v & 0x7fffffffu));
    if (v & 0x80000000u) p.y = fneg(p.y);
    acc = xyzz_madd(acc, p);
  }
  if (nitems == 1) st_vec(sets.buckets[set] + b, acc);
  else st_vec(sets.partials[set] + id, acc);
}

// buckets split into 2..MSM_FIXUP_SMALL_MAX work items (e.g. the heavier low buckets fed by the short top
// window): one thread adds the few partial sums.  grid.x covers nbuckets threads, grid.y = point sets.
template <class F>
__global__ void __launch_bounds__(128) k_bucket_fixup_small(AccSets<F> sets, const uint32_t* __restrict__ item_start,
                                                            const uint32_t* __restrict__ multi) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= multi[0]) return;
  const int set = blockIdx.y;
  const uint32_t b = multi[2 + i];
  const uint32_t i0 = item_start[b], i1 = item_start[b + 1];
  XYZZ<F> acc = ld_vec(sets.partials[set] + i0);
  for (uint32_t k = i0 + 1; k < i1; k++) {
    XYZZ<F> o = ld_vec(sets.partials[set] + k);
    xyzz_add_ni(acc, acc, o);
  }
  st_vec(sets.buckets[set] + b, acc);
}

// giant buckets (skewed scalar distributions, degenerate top window): a block tree-sums the partials
template <class F>
__global__ void __launch_bounds__(128) k_bucket_fixup(AccSets<F> sets, const uint32_t* __restrict__ item_start,
                                                      const uint32_t* __restrict__ multi, uint32_t nbuckets) {
  extern __shared__ uint4 red_raw[];
  XYZZ<F>* red = reinterpret_cast<XYZZ<F>*>(red_raw);
  const int set = blockIdx.y;
  const uint32_t count = multi[1];
  const uint32_t* list = multi + 2 + nbuckets;
  for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
    uint32_t b = list[i];
    uint32_t i0 = item_start[b], i1 = item_start[b + 1];
    XYZZ<F> acc = xyzz_inf<F>();
    for (uint32_t k = i0 + threadIdx.x; k < i1; k += blockDim.x) {
      XYZZ<F> o = ld_vec(sets.partials[set] + k);
      xyzz_add_ni(acc, acc, o);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
      if (threadIdx.x < s) {
        XYZZ<F> o = red[threadIdx.x + s];
        xyzz_add_ni(acc, acc, o);
        red[threadIdx.x] = acc;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) st_vec(sets.buckets[set] + b, acc);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// bucket reduction: per bucket set, sum_k (k+1) * B_k.
//   level 1: thread t owns L consecutive buckets; running sums give S_t = sum_j B_{tL+j} and
//            R_t = sum_j (j+1) B_{tL+j}.  Then sum_k (k+1) B_k = sum_t R_t + L * sum_t t * S_t.
//   sum_t t * S_t is the same problem on the L-times shorter array S with 0-based weights; it is summed bit by
//   bit of t (k_reduce_bits) and k_reduce_final assembles  R + L * sum_j 2^j U_j.
// No per-thread scalar multiplication.   grid = (blocks, bucket sets (windows), point sets)
// ---------------------------------------------------------------------------------------
template <class F>
struct ReduceLevel {
  const XYZZ<F>* in[MsmAccumulator<F>::MAX_SETS];    // n_in entries per bucket set
  XYZZ<F>* out_s[MsmAccumulator<F>::MAX_SETS];       // n_in / L entries per bucket set (input of the next level)
  XYZZ<F>* out_r[MsmAccumulator<F>::MAX_SETS];       // gridDim.x block sums of R per bucket set
  uint32_t n_in, L;
  uint32_t out_r_stride;                             // entries between the R partials of consecutive bucket sets
  int weight_one_based;                              // level 1: weights j+1; level 2: weights j
};

template <class F>
__global__ void __launch_bounds__(128) k_reduce_level(ReduceLevel<F> a) {
  extern __shared__ uint4 red_raw[];
  XYZZ<F>* red = reinterpret_cast<XYZZ<F>*>(red_raw);
  const uint32_t w = blockIdx.y;
  const int set = blockIdx.z;
  const uint32_t nthreads = a.n_in / a.L;                       // per bucket set
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  XYZZ<F> running = xyzz_inf<F>(), sum = xyzz_inf<F>();
  if (t < nthreads) {
    const XYZZ<F>* B = a.in[set] + (size_t)w * a.n_in + (size_t)t * a.L;
    for (int k = (int)a.L - 1; k >= 0; k--) {
      XYZZ<F> b = ld_vec(B + k);
      xyzz_add_ni(running, running, b);
      if (k > 0 || a.weight_one_based) xyzz_add_ni(sum, sum, running);
    }
    st_vec(a.out_s[set] + (size_t)w * nthreads + t, running);
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
  if (threadIdx.x == 0) st_vec(a.out_r[set] + (size_t)w * a.out_r_stride + blockIdx.x, sum);
}

// The array S left by the running-sum levels (n entries, 0-based weights t) is not reduced by further levels --
// they are chains of tiny latency-bound launches on the critical path of every MSM -- but bit by bit:
//   sum_t t * S_t = sum_j 2^j * U_j,   U_j = sum of the S_t whose index has bit j set.
// k_reduce_bits: block (j, chunk of 1024 entries) -> one partial of U_j.  grid = (nbits * nchunks, bucket sets,
// point sets).  About log2(n)/2 additions per entry instead of 2, all of them independent.
// Partials of the "small levels" (R of level 2, then U_0 .. U_{nbits-1}) live in one array indexed
// [(bucket set * nsmall + level) * stride + k].
template <class F>
struct ReduceBits {
  const XYZZ<F>* in[MsmAccumulator<F>::MAX_SETS];    // n entries per bucket set
  XYZZ<F>* out[MsmAccumulator<F>::MAX_SETS];         // small-level array
  uint32_t n, nbits, nchunks, nsmall, first_level, stride;
};

template <class F>
__global__ void __launch_bounds__(128) k_reduce_bits(ReduceBits<F> a) {
  extern __shared__ uint4 red_raw[];
  XYZZ<F>* red = reinterpret_cast<XYZZ<F>*>(red_raw);
  const uint32_t chunk = blockIdx.x % a.nchunks, j = blockIdx.x / a.nchunks;
  const uint32_t w = blockIdx.y;
  const int set = blockIdx.z;
  const uint32_t base = chunk * REDUCE_BITS_CHUNK;
  XYZZ<F>* dst = a.out[set] + ((size_t)w * a.nsmall + a.first_level + j) * a.stride + chunk;
  if ((1u << j) >= REDUCE_BITS_CHUNK && !((base >> j) & 1u)) {   // bit j is constant over the chunk, and clear
    if (threadIdx.x == 0) st_vec(dst, xyzz_inf<F>());
    return;
  }
  const XYZZ<F>* S = a.in[set] + (size_t)w * a.n;
  XYZZ<F> acc = xyzz_inf<F>();
  for (uint32_t k = 0; k < REDUCE_BITS_CHUNK / 128; k++) {
    const uint32_t t = base + k * 128 + threadIdx.x;
    if (t < a.n && ((t >> j) & 1u)) {
      XYZZ<F> o = ld_vec(S + t);
      xyzz_add_ni(acc, acc, o);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      XYZZ<F> o = red[threadIdx.x + s];
      xyzz_add_ni(acc, acc, o);
      red[threadIdx.x] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_vec(dst, acc);
}

// total of one bucket set = R1 + sum over the small levels of 2^shift * (sum of the level's partials).
// grid = (bucket sets, point sets), 128 threads: 48 sum the R1 partials, 4 per small level sum its partials;
// the doublings of the levels run in parallel, a tree adds the levels.
constexpr uint32_t REDUCE_FINAL_RTHREADS = 48;
template <class F>
struct ReduceFinal {
  const XYZZ<F>* r1[MsmAccumulator<F>::MAX_SETS];
  const XYZZ<F>* small[MsmAccumulator<F>::MAX_SETS];
  XYZZ<F>* wintot[MsmAccumulator<F>::MAX_SETS];
  uint32_t blocks1, nsmall, stride;
  uint8_t shift[REDUCE_MAX_SMALL];
  uint16_t count[REDUCE_MAX_SMALL];
};

template <class F>
__global__ void __launch_bounds__(128) k_reduce_final(ReduceFinal<F> a) {
  extern __shared__ uint4 red_raw[];
  XYZZ<F>* red = reinterpret_cast<XYZZ<F>*>(red_raw);
  const uint32_t w = blockIdx.x;
  const int set = blockIdx.y;
  const uint32_t tid = threadIdx.x;
  uint32_t seg_base, local, width;
  XYZZ<F> acc = xyzz_inf<F>();
  if (tid < REDUCE_FINAL_RTHREADS) {
    seg_base = 0;
    local = tid;
    width = REDUCE_FINAL_RTHREADS;
    for (uint32_t k = local; k < a.blocks1; k += REDUCE_FINAL_RTHREADS) {
      XYZZ<F> o = ld_vec(a.r1[set] + (size_t)w * a.blocks1 + k);
      xyzz_add_ni(acc, acc, o);
    }
  } else {
    const uint32_t l = (tid - REDUCE_FINAL_RTHREADS) >> 2;
    seg_base = REDUCE_FINAL_RTHREADS + 4 * l;
    local = (tid - REDUCE_FINAL_RTHREADS) & 3u;
    width = 4;
    if (l < a.nsmall)
      for (uint32_t k = local; k < a.count[l]; k += 4) {
        XYZZ<F> o = ld_vec(a.small[set] + ((size_t)w * a.nsmall + l) * a.stride + k);
        xyzz_add_ni(acc, acc, o);
      }
  }
  red[tid] = acc;
  __syncthreads();
  for (uint32_t s = 32; s > 0; s >>= 1) {
    if (local < s && local + s < width) {
      XYZZ<F> o = red[seg_base + local + s];
      xyzz_add_ni(acc, acc, o);
      red[tid] = acc;
    }
    __syncthreads();
  }
  const XYZZ<F> rsum = acc;                   // meaningful in thread 0
  XYZZ<F> v = xyzz_inf<F>();
  if (tid < a.nsmall) {
    v = red[REDUCE_FINAL_RTHREADS + 4 * tid];
    for (uint32_t i = 0; i < a.shift[tid]; i++) xyzz_dbl_ni(v, v);
  }
  __syncthreads();
  red[tid] = v;
  __syncthreads();
  for (uint32_t s = 16; s > 0; s >>= 1) {
    if (tid < s && tid + s < a.nsmall) {
      XYZZ<F> o = red[tid + s];
      xyzz_add_ni(v, v, o);
      red[tid] = v;
    }
    __syncthreads();
  }
  if (tid == 0) {
    xyzz_add_ni(v, v, rsum);
    st_vec(a.wintot[set] + w, v);
  }
}

// result = sum_w 2^(c w) * W_w  (Horner from the top window; one window in the precomputed layout)
// grid.x = point sets; sets.winpart[] points at the window totals
template <class F>
__global__ void k_window_combine(AccSets<F> sets, int nwin, int c) {
  if (threadIdx.x != 0) return;
  const int set = blockIdx.x;
  XYZZ<F> acc = ld_vec(sets.winpart[set] + (nwin - 1));
  for (int i = nwin - 2; i >= 0; i--) {
    for (int j = 0; j < c; j++) xyzz_dbl_ni(acc, acc);
    XYZZ<F> o = ld_vec(sets.winpart[set] + i);
    xyzz_add_ni(acc, acc, o);
  }
  st_vec(sets.result[set], acc);
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

// ---------------------------------------------------------------------------------------
// precomputed window table: one thread per point, c doublings per window, one batched inversion
// ---------------------------------------------------------------------------------------
constexpr int MSM_MAX_WINDOWS = 128;

template <class F>
__global__ void __launch_bounds__(128) k_build_table(const Affine<F>* __restrict__ points, uint32_t n_points,
                                                     uint32_t pad_front, int c, int nwin,
                                                     Affine<F>* __restrict__ table) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = n_points + pad_front;
  if (i >= n) return;
  if (i < pad_front) {
    for (int w = 0; w < nwin; w++) st_vec(table + (size_t)w * n + i, aff_inf<F>());
    return;
  }
  Affine<F> p = ldg_vec(points + (i - pad_front));
  st_vec(table + i, p);
  if (aff_is_inf(p)) {
    for (int w = 1; w < nwin; w++) st_vec(table + (size_t)w * n + i, aff_inf<F>());
    return;
  }
  // Montgomery's batch-inversion trick over the nwin-1 values t_w = zz_w * zzz_w of this point's doubling
  // chain, with the table itself as scratch (no extra memory): the chain is walked twice.
  // pass 1: forward walk; slot_w = (prefix product t_1..t_{w-1}, t_w)
  XYZZ<F> q = xyzz_from_affine(p);
  F prod = F::one();
  for (int w = 1; w < nwin; w++) {
    for (int j = 0; j < c; j++) xyzz_dbl_ni(q, q);
    // BN254 G1/G2 have odd order: 2^k P is never infinity for P != infinity, so t_w != 0
    Affine<F> slot;
    slot.x = prod;
    slot.y = fmul(q.zz, q.zzz);
    st_vec(table + (size_t)w * n + i, slot);
    prod = fmul(prod, slot.y);
  }
  F inv_all = finv(prod);
  // pass 2: backward over the slots; 1/t_w = inv_all * prefix_w * suffix_w
  q = xyzz_from_affine(p);
  F suffix = F::one();
  for (int w = nwin - 1; w >= 1; w--) {
    Affine<F> slot = ld_vec(table + (size_t)w * n + i);
    F t = slot.y;
    slot.y = fmul(fmul(inv_all, slot.x), suffix);   // = 1 / t_w
    st_vec(table + (size_t)w * n + i, slot);
    suffix = fmul(suffix, t);
  }
  // pass 3: forward walk again, normalising each multiple with its stored inverse
  for (int w = 1; w < nwin; w++) {
    for (int j = 0; j < c; j++) xyzz_dbl_ni(q, q);
    Affine<F> slot = ld_vec(table + (size_t)w * n + i);
    F it = slot.y;                                   // 1 / (zz * zzz)
    Affine<F> a;
    a.x = fmul(q.x, fmul(it, q.zzz));                // X / zz
    a.y = fmul(q.y, fmul(it, q.zz));                 // Y / zzz
    st_vec(table + (size_t)w * n + i, a);
  }
}

// ---------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------
template <class F>
MsmAccumulator<F>::~MsmAccumulator() {
  for (int i = 0; i < 2; i++)
    if (pev_[i]) cudaEventDestroy(pev_[i]);
  if (lev_) cudaEventDestroy(lev_);
}
template <class F>
size_t MsmAccumulator<F>::workspace_bytes() const {
  return buckets_.bytes + partials_.bytes + winpart_.bytes + tree_w_.bytes + tree_m_.bytes + tree_nodes_.bytes +
         tree_bp_.bytes;
}
template <class F>
float MsmAccumulator<F>::last_accum_ms() const {
  float ms = 0.f;
  if (pev_[0] && pev_[1]) cudaEventElapsedTime(&ms, pev_[0], pev_[1]);
  return ms;
}

// occupancy variant of the accumulate kernel: G1 fits 4 CTAs/SM by itself; for G2 the register budget is a
// trade-off between spills and resident warps (G16_G2_MINB overrides the default for experiments)
template <class F>
static void launch_accumulate(dim3 grid, cudaStream_t stream, const AccSets<F>& sets, const MsmSorter& sorter,
                              const MsmGeometry& g) {
  static int minb = -1;
  if (minb < 0) {
    const char* e = getenv("G16_G2_MINB");
    minb = e ? atoi(e) : G16_G2_MINB_DEFAULT;
  }
#define G16_ACC_ARGS sets, sorter.vals(), sorter.start(), sorter.item_start(), sorter.item_bucket(), \
                     sorter.items_sorted(), g.nbuckets, g.max_items, g.T
  if constexpr (sizeof(F) == sizeof(Fp)) {
    k_bucket_accumulate<F, 1><<<grid, 128, 0, stream>>>(G16_ACC_ARGS);       // 126 registers: 4 CTAs/SM anyway
  } else {
    if (minb == 2) k_bucket_accumulate<F, 2><<<grid, 128, 0, stream>>>(G16_ACC_ARGS);
    else if (minb == 4) k_bucket_accumulate<F, 4><<<grid, 128, 0, stream>>>(G16_ACC_ARGS);   // 128 registers, heavy spills: 7.8 ms
    else k_bucket_accumulate<F, 3><<<grid, 128, 0, stream>>>(G16_ACC_ARGS);   // 168 registers, a few spills: 6.9 vs 7.3 ms
  }
#undef G16_ACC_ARGS
}

#ifdef G16_EXPERIMENTS
// batched-affine bucket accumulation (msm_tree.cuh): tree_log rounds of three launches, then the chunk heads
// become the bucket sums
template <class F>
void MsmAccumulator<F>::run_tree(const MsmSorter& sorter, const MsmPointSet<F>* in, int nsets,
                                 XYZZ<F>* const* buckets, cudaStream_t stream) {
  const MsmGeometry& g = sorter.geom();
  const size_t cap0 = sorter.tree_cap(0);
  const size_t blocks0 = (cap0 + TREE_PER_BLOCK - 1) / TREE_PER_BLOCK;
  const size_t w_bytes = g.m * sizeof(Affine<F>);
  const size_t m_bytes = cap0 * sizeof(F);
  const size_t n_bytes = blocks0 * 2 * TREE_TPB * sizeof(F);
  const size_t b_bytes = blocks0 * sizeof(F);
  tree_w_.ensure(w_bytes * nsets);
  tree_m_.ensure(m_bytes * nsets);
  tree_nodes_.ensure(n_bytes * nsets);
  tree_bp_.ensure(2 * b_bytes * nsets);
  TreeSets<F> ts;
  for (int s = 0; s < MAX_SETS; s++) {
    int k = s < nsets ? s : 0;
    ts.points[s] = in[k].points;
    ts.W[s] = reinterpret_cast<Affine<F>*>(tree_w_.as<char>() + w_bytes * k);
    ts.M[s] = reinterpret_cast<F*>(tree_m_.as<char>() + m_bytes * k);
    ts.tree[s] = reinterpret_cast<F*>(tree_nodes_.as<char>() + n_bytes * k);
    ts.bp[s] = reinterpret_cast<F*>(tree_bp_.as<char>() + 2 * b_bytes * k);
    ts.ibp[s] = reinterpret_cast<F*>(tree_bp_.as<char>() + 2 * b_bytes * k + b_bytes);
    ts.buckets[s] = buckets[k];
  }
  for (int r = 0; r < g.tree_log; r++) {
    dim3 grid(div_up(sorter.tree_cap(r), TREE_PER_BLOCK), (unsigned)nsets);
    if (r == 0)
      k_tree_prepare<F, true><<<grid, TREE_TPB, 0, stream>>>(ts, sorter.vals(), sorter.tree_list(r), sorter.tree_count(r), r);
    else
      k_tree_prepare<F, false><<<grid, TREE_TPB, 0, stream>>>(ts, sorter.vals(), sorter.tree_list(r), sorter.tree_count(r), r);
    G16_LAUNCH_CHECK();
    k_tree_invert<F><<<nsets, TREE_TPB, 0, stream>>>(ts, sorter.tree_count(r));
    G16_LAUNCH_CHECK();
    k_tree_finish<F><<<grid, TREE_TPB, 0, stream>>>(ts, sorter.tree_list(r), sorter.tree_count(r), r);
    G16_LAUNCH_CHECK();
  }
  dim3 bgrid(div_up(g.nbuckets, 128), (unsigned)nsets);
  k_tree_finalize<F><<<bgrid, 128, 0, stream>>>(ts, sorter.start(), sorter.item_start(), g.nbuckets);
  G16_LAUNCH_CHECK();
  k_tree_fixup_small<F><<<bgrid, 128, 0, stream>>>(ts, sorter.start(), sorter.item_start(), sorter.multi(),
                                                    g.tree_log);
  G16_LAUNCH_CHECK();
  dim3 fgrid(148, (unsigned)nsets);
  k_tree_fixup_big<F><<<fgrid, 128, 128 * sizeof(XYZZ<F>), stream>>>(ts, sorter.start(), sorter.item_start(),
                                                                     sorter.multi(), g.nbuckets, g.tree_log);
  G16_LAUNCH_CHECK();
}

#endif

template <class F>
void MsmAccumulator<F>::run(const MsmSorter& sorter, const MsmPointSet<F>* in, int nsets, cudaStream_t stream,
                            cudaStream_t tail) {
  G16_REQUIRE(nsets >= 1 && nsets <= MAX_SETS, "MsmAccumulator: 1..3 point sets");
  const MsmGeometry& g = sorter.geom();
  G16_REQUIRE(g.nwin <= MSM_MAX_WINDOWS, "too many windows");
  // bucket reduction plan (msm_reduce_plan.cuh): running sums over 16 consecutive buckets (level 1), over 8
  // consecutive entries of the result when it is still long (level 2), then the bit-sliced sum of what is left
  const uint32_t nsetsB = g.precomp ? 1u : (uint32_t)g.nwin;
  static int l1_max = -1, l2_min = -1;
  if (l1_max < 0) {
    const char* e = getenv("G16_REDUCE_L1");          // experiment knobs: length of the level-1 running sums ...
    l1_max = e ? atoi(e) : 16;
    const char* e2 = getenv("G16_REDUCE_L2MIN");      // ... and the array length from which level 2 runs
    l2_min = e2 ? atoi(e2) : 8192;
  }
  const ReducePlan rp = msm_reduce_plan(g.nb, (uint32_t)l1_max, (uint32_t)l2_min);
  G16_REQUIRE(rp.ok, "too many buckets for the reduction");
  const uint32_t L1 = rp.L1, n1 = rp.n1, L2 = rp.L2, n2 = rp.n2, tpb1 = rp.tpb1, blocks1 = rp.blocks1, tpb2 = rp.tpb2,
                 blocks2 = rp.blocks2, nbits = rp.nbits, nchunks = rp.nchunks, first_bit_level = rp.first_bit_level,
                 nsmall = rp.nsmall, stride = rp.stride;
  // scratch per point set: S1 | S2 | R1 partials | small-level partials | one total per bucket set
  const size_t s1_entries = (size_t)nsetsB * n1;
  const size_t s2_entries = L2 > 1 ? (size_t)nsetsB * n2 : 0;
  const size_t r_entries = (size_t)nsetsB * blocks1;
  const size_t u_entries = (size_t)nsetsB * nsmall * stride;

  size_t bucket_bytes = (size_t)g.nbuckets * sizeof(XYZZ<F>);
  size_t partial_bytes = g.tree_log ? 0 : (size_t)g.max_items * sizeof(XYZZ<F>);
  size_t winpart_bytes = (s1_entries + s2_entries + r_entries + u_entries + nsetsB) * sizeof(XYZZ<F>);
  buckets_.ensure(bucket_bytes * nsets);
  partials_.ensure(partial_bytes * nsets);
  winpart_.ensure(winpart_bytes * nsets);
  AccSets<F> sets;
  for (int s = 0; s < MAX_SETS; s++) {
    int k = s < nsets ? s : 0;
    sets.points[s] = in[k].points;
    sets.result[s] = in[k].result;
    sets.buckets[s] = reinterpret_cast<XYZZ<F>*>(buckets_.as<char>() + bucket_bytes * k);
    sets.partials[s] = reinterpret_cast<XYZZ<F>*>(partials_.as<char>() + partial_bytes * k);
    sets.winpart[s] = reinterpret_cast<XYZZ<F>*>(winpart_.as<char>() + winpart_bytes * k);
  }
  // empty buckets are never written by a work item: zero bytes == infinity
  G16_CUDA(cudaMemsetAsync(buckets_.p, 0, bucket_bytes * nsets, stream));
  if (profile) {
    for (int i = 0; i < 2; i++)
      if (!pev_[i]) G16_CUDA(cudaEventCreate(&pev_[i]));
    G16_CUDA(cudaEventRecord(pev_[0], stream));
  }
  if (g.tree_log) {
#ifdef G16_EXPERIMENTS
    run_tree(sorter, in, nsets, sets.buckets, stream);
    if (profile) G16_CUDA(cudaEventRecord(pev_[1], stream));
#else
    G16_REQUIRE(false, "batched-affine tree mode needs a library built with `make EXPERIMENTS=1`");
#endif
  } else {
    dim3 agrid(div_up(g.max_items, 128), (unsigned)nsets);
    launch_accumulate<F>(agrid, stream, sets, sorter, g);
    G16_LAUNCH_CHECK();
    if (profile) G16_CUDA(cudaEventRecord(pev_[1], stream));
    dim3 sgrid(div_up(g.nbuckets, 128), (unsigned)nsets);
    k_bucket_fixup_small<F><<<sgrid, 128, 0, stream>>>(sets, sorter.item_start(), sorter.multi());
    G16_LAUNCH_CHECK();
    dim3 fgrid(148, (unsigned)nsets);
    k_bucket_fixup<F><<<fgrid, 128, 128 * sizeof(XYZZ<F>), stream>>>(sets, sorter.item_start(), sorter.multi(),
                                                                     g.nbuckets);
    G16_LAUNCH_CHECK();
  }
  // bucket reduction: three or four launches
  const size_t off_s2 = s1_entries, off_r = s1_entries + s2_entries, off_u = off_r + r_entries, off_t = off_u + u_entries;
  ReduceLevel<F> lv1, lv2;
  lv1.n_in = g.nb;
  lv1.L = L1;
  lv1.out_r_stride = blocks1;
  lv1.weight_one_based = 1;
  lv2.n_in = n1;
  lv2.L = L2;
  lv2.out_r_stride = nsmall * stride;
  lv2.weight_one_based = 0;
  ReduceBits<F> rb;
  rb.n = n2;
  rb.nbits = nbits;
  rb.nchunks = nchunks;
  rb.nsmall = nsmall;
  rb.first_level = first_bit_level;
  rb.stride = stride;
  ReduceFinal<F> fin;
  fin.blocks1 = blocks1;
  fin.nsmall = nsmall;
  fin.stride = stride;
  for (uint32_t l = 0; l < REDUCE_MAX_SMALL; l++) {
    fin.shift[l] = rp.shift[l];
    fin.count[l] = rp.count[l];
  }
  for (int s = 0; s < MAX_SETS; s++) {
    XYZZ<F>* base = sets.winpart[s];
    lv1.in[s] = sets.buckets[s];
    lv1.out_s[s] = base;
    lv1.out_r[s] = base + off_r;
    lv2.in[s] = base;
    lv2.out_s[s] = base + off_s2;
    lv2.out_r[s] = base + off_u;                      // small level 0
    rb.in[s] = L2 > 1 ? base + off_s2 : base;
    rb.out[s] = base + off_u;
    fin.r1[s] = base + off_r;
    fin.small[s] = base + off_u;
    fin.wintot[s] = base + off_t;
  }
  k_reduce_level<F><<<dim3(blocks1, nsetsB, (unsigned)nsets), tpb1, tpb1 * sizeof(XYZZ<F>), stream>>>(lv1);
  G16_LAUNCH_CHECK();
  if (tail && tail != stream) {           // everything below is latency-bound: continue on the high-priority stream
    if (!lev_) G16_CUDA(cudaEventCreateWithFlags(&lev_, cudaEventDisableTiming));
    G16_CUDA(cudaEventRecord(lev_, stream));
    G16_CUDA(cudaStreamWaitEvent(tail, lev_, 0));
    stream = tail;
  }
  if (L2 > 1) {
    k_reduce_level<F><<<dim3(blocks2, nsetsB, (unsigned)nsets), tpb2, tpb2 * sizeof(XYZZ<F>), stream>>>(lv2);
    G16_LAUNCH_CHECK();
  }
  if (nbits) {
    k_reduce_bits<F><<<dim3(nbits * nchunks, nsetsB, (unsigned)nsets), 128, 128 * sizeof(XYZZ<F>), stream>>>(rb);
    G16_LAUNCH_CHECK();
  }
  k_reduce_final<F><<<dim3(nsetsB, (unsigned)nsets), 128, 128 * sizeof(XYZZ<F>), stream>>>(fin);
  G16_LAUNCH_CHECK();
  AccSets<F> wsets = sets;
  for (int s = 0; s < MAX_SETS; s++) wsets.winpart[s] = sets.winpart[s] + off_t;
  k_window_combine<F><<<nsets, 32, 0, stream>>>(wsets, g.precomp ? 1 : g.nwin, g.c);
  G16_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------
template <class F>
Msm<F>::~Msm() {
  for (int i = 0; i < 2; i++)
    if (tev_[i]) cudaEventDestroy(tev_[i]);
}
template <class F>
float Msm<F>::last_total_ms() const {
  float ms = 0.f;
  if (tev_[0] && tev_[1]) cudaEventElapsedTime(&ms, tev_[0], tev_[1]);
  return ms;
}

template <class F>
void Msm<F>::go(const Fr* scalars, bool mont, const Affine<F>* pts, const MsmGeometry& g, XYZZ<F>* result,
                cudaStream_t stream) {
  last_c = g.c;
  last_nwin = g.nwin;
  if (profile_) {
    for (int i = 0; i < 2; i++)
      if (!tev_[i]) G16_CUDA(cudaEventCreate(&tev_[i]));
    G16_CUDA(cudaEventRecord(tev_[0], stream));
  }
  sorter_.run(scalars, mont, g, stream);
  MsmPointSet<F> set;
  set.points = pts;
  set.result = result;
  acc_.run(sorter_, &set, 1, stream);
  if (profile_) {
    G16_CUDA(cudaEventRecord(tev_[1], stream));
    uint32_t pairs = 0;   // start[nbuckets] = number of sorted pairs with a real bucket key
    G16_CUDA(cudaMemcpyAsync(&pairs, sorter_.start() + g.nbuckets, 4, cudaMemcpyDeviceToHost, stream));
    G16_CUDA(cudaStreamSynchronize(stream));
    last_pairs = pairs;
  }
}

template <class F>
void Msm<F>::run(const Fr* scalars, bool scalars_mont, const Affine<F>* points, size_t n, XYZZ<F>* result,
                 cudaStream_t stream, int window_bits) {
  if (n == 0) {                                   // zero-length MSM (C1 when nvars = npubs+1)
    k_set_inf<F><<<1, 32, 0, stream>>>(result);
    G16_LAUNCH_CHECK();
    return;
  }
  int c = window_bits ? window_bits : msm_pick_window(n, false);
  go(scalars, scalars_mont, points, msm_geometry(n, c, false), result, stream);
}

template <class F>
void Msm<F>::run_precomp(const Fr* scalars, bool scalars_mont, const Affine<F>* table, size_t n, int window_bits,
                         XYZZ<F>* result, cudaStream_t stream) {
  if (n == 0) {
    k_set_inf<F><<<1, 32, 0, stream>>>(result);
    G16_LAUNCH_CHECK();
    return;
  }
  go(scalars, scalars_mont, table, msm_geometry(n, window_bits, true), result, stream);
}

template <class F>
void msm_build_table(const Affine<F>* points, size_t n_points, size_t pad_front, int c, Affine<F>* table,
                     cudaStream_t stream) {
  size_t n = n_points + pad_front;
  if (!n) return;
  int nwin = msm_num_windows(c);
  k_build_table<F><<<div_up(n, 128), 128, 0, stream>>>(points, (uint32_t)n_points, (uint32_t)pad_front, c, nwin,
                                                       table);
  G16_LAUNCH_CHECK();
}

template <class F>
void xyzz_sum_to_affine(const XYZZ<F>* parts, int count, Affine<F>* out, cudaStream_t stream) {
  k_xyzz_sum_to_affine<F><<<1, 32, 0, stream>>>(parts, count, out);
  G16_LAUNCH_CHECK();
}

}  // namespace g16
