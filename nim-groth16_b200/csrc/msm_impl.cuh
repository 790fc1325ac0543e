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
//   k_reduce_level/final sum_k (k+1) B_k per bucket set as a recursion over levels of running sums
//   k_window_combine     Horner over the windows (plain layout; a single window in the precomputed layout)
//   k_build_table        precomputed layout: 2^(c w) P_i for every window, batch-normalised to affine
#pragma once
#include "msm.cuh"
#include <stdlib.h>
#include "msm_digits.cuh"
#include "msm_tree.cuh"

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
// bucket reduction: per bucket set, sum_k (k+1) * B_k, as a recursion over levels.
//   level 1: thread t owns L consecutive buckets; running sums give S_t = sum_j B_{tL+j} and
//            R_t = sum_j (j+1) B_{tL+j}.  Then sum_k (k+1) B_k = sum_t R_t + L * sum_t t * S_t,
//   and sum_t t * S_t is the same problem on the L-times shorter array S with 0-based weights (level 2, ...).
//   Result = R^(1) + L * (R^(2) + L * (R^(3) + ...)),  R^(l) = plain sum of the level's local weighted sums.
// No per-thread scalar multiplication; each level is one launch and the tiny upper levels overlap with the
// other streams' work.   grid = (blocks, bucket sets (windows), point sets)
// ---------------------------------------------------------------------------------------
template <class F>
struct ReduceLevel {
  const XYZZ<F>* in[MsmAccumulator<F>::MAX_SETS];    // n_in entries per bucket set
  XYZZ<F>* out_s[MsmAccumulator<F>::MAX_SETS];       // n_in / L entries per bucket set (input of the next level)
  XYZZ<F>* out_r[MsmAccumulator<F>::MAX_SETS];       // gridDim.x block sums of R per bucket set
  uint32_t n_in, L;
  int weight_one_based;                              // level 1: weights j+1; upper levels: weights j
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
  if (threadIdx.x == 0) st_vec(a.out_r[set] + (size_t)w * gridDim.x + blockIdx.x, sum);
}

constexpr int MSM_MAX_LEVELS = 12;
template <class F>
struct ReduceFinal {
  const XYZZ<F>* r[MsmAccumulator<F>::MAX_SETS][MSM_MAX_LEVELS];   // block sums of each level
  XYZZ<F>* wintot[MsmAccumulator<F>::MAX_SETS];                    // one total per bucket set (window)
  uint32_t blocks[MSM_MAX_LEVELS];
  uint32_t logL[MSM_MAX_LEVELS];
  int nlevels;
};

// total of one bucket set: Horner over the levels; grid = (bucket sets, point sets), 128 threads
template <class F>
__global__ void __launch_bounds__(128) k_reduce_final(ReduceFinal<F> a) {
  extern __shared__ uint4 red_raw[];
  XYZZ<F>* red = reinterpret_cast<XYZZ<F>*>(red_raw);
  const uint32_t w = blockIdx.x;
  const int set = blockIdx.y;
  XYZZ<F> total = xyzz_inf<F>();          // only meaningful in thread 0
  for (int l = a.nlevels - 1; l >= 0; l--) {
    XYZZ<F> acc = xyzz_inf<F>();
    for (uint32_t k = threadIdx.x; k < a.blocks[l]; k += blockDim.x) {
      XYZZ<F> o = ld_vec(a.r[set][l] + (size_t)w * a.blocks[l] + k);
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
    if (threadIdx.x == 0) {
      // total = R^(l) + L_l * total   (the levels above act through the weights of this level)
      for (uint32_t j = 0; j < a.logL[l]; j++) xyzz_dbl_ni(total, total);
      xyzz_add_ni(total, total, acc);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_vec(a.wintot[set] + w, total);
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
    else k_bucket_accumulate<F, 3><<<grid, 128, 0, stream>>>(G16_ACC_ARGS);   // 168 registers, a few spills: 6.9 vs 7.3 ms
  }
#undef G16_ACC_ARGS
}

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

template <class F>
void MsmAccumulator<F>::run(const MsmSorter& sorter, const MsmPointSet<F>* in, int nsets, cudaStream_t stream) {
  G16_REQUIRE(nsets >= 1 && nsets <= MAX_SETS, "MsmAccumulator: 1..3 point sets");
  const MsmGeometry& g = sorter.geom();
  G16_REQUIRE(g.nwin <= MSM_MAX_WINDOWS, "too many windows");
  // reduction levels: n_0 = nb buckets; level l maps n_l entries to n_l / L_l entries, down to one
  const uint32_t nsetsB = g.precomp ? 1u : (uint32_t)g.nwin;
  int nlevels = 0;
  uint32_t lvl_n[MSM_MAX_LEVELS + 1], lvl_L[MSM_MAX_LEVELS], lvl_tpb[MSM_MAX_LEVELS], lvl_blocks[MSM_MAX_LEVELS];
  lvl_n[0] = g.nb;
  while (lvl_n[nlevels] > 1) {
    G16_REQUIRE(nlevels < MSM_MAX_LEVELS, "too many reduction levels");
    // level 1 is throughput-bound (fan-in 16 keeps its work at 2.1 additions per bucket); the upper levels are
    // latency-bound chains of tiny launches, so they use fan-in 4 (7 sequential additions per level)
    uint32_t want = nlevels == 0 ? 16u : 4u;
    uint32_t L = lvl_n[nlevels] >= want ? want : lvl_n[nlevels];
    uint32_t threads = lvl_n[nlevels] / L;
    lvl_L[nlevels] = L;
    lvl_tpb[nlevels] = threads < 128u ? (threads < 32u ? 32u : threads) : 128u;
    lvl_blocks[nlevels] = (threads + lvl_tpb[nlevels] - 1) / lvl_tpb[nlevels];
    lvl_n[nlevels + 1] = threads;
    nlevels++;
  }
  // scratch per point set: the S array and the block sums of every level, then one total per bucket set
  size_t s_entries = 0, r_entries = 0;
  for (int l = 0; l < nlevels; l++) {
    s_entries += (size_t)nsetsB * lvl_n[l + 1];
    r_entries += (size_t)nsetsB * lvl_blocks[l];
  }

  size_t bucket_bytes = (size_t)g.nbuckets * sizeof(XYZZ<F>);
  size_t partial_bytes = g.tree_log ? 0 : (size_t)g.max_items * sizeof(XYZZ<F>);
  size_t winpart_bytes = (s_entries + r_entries + nsetsB) * sizeof(XYZZ<F>);
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
    run_tree(sorter, in, nsets, sets.buckets, stream);
    if (profile) G16_CUDA(cudaEventRecord(pev_[1], stream));
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
  // bucket reduction: one launch per level, then the per-set Horner over the levels
  ReduceFinal<F> fin;
  fin.nlevels = nlevels;
  size_t off_s = 0, off_r = s_entries;
  const size_t off_t = s_entries + r_entries;
  for (int l = 0; l < nlevels; l++) {
    ReduceLevel<F> lv;
    lv.n_in = lvl_n[l];
    lv.L = lvl_L[l];
    lv.weight_one_based = (l == 0) ? 1 : 0;
    for (int s = 0; s < MAX_SETS; s++) {
      XYZZ<F>* base = sets.winpart[s];
      lv.in[s] = (l == 0) ? sets.buckets[s] : base + (off_s - (size_t)nsetsB * lvl_n[l]);
      lv.out_s[s] = base + off_s;
      lv.out_r[s] = base + off_r;
      fin.r[s][l] = base + off_r;
      fin.wintot[s] = base + off_t;
    }
    fin.blocks[l] = lvl_blocks[l];
    uint32_t lg = 0;
    while ((1u << lg) < lvl_L[l]) lg++;
    fin.logL[l] = lg;
    dim3 rgrid(lvl_blocks[l], nsetsB, (unsigned)nsets);
    k_reduce_level<F><<<rgrid, lvl_tpb[l], lvl_tpb[l] * sizeof(XYZZ<F>), stream>>>(lv);
    G16_LAUNCH_CHECK();
    off_s += (size_t)nsetsB * lvl_n[l + 1];
    off_r += (size_t)nsetsB * lvl_blocks[l];
  }
  for (int l = nlevels; l < MSM_MAX_LEVELS; l++) {
    fin.blocks[l] = 0;
    fin.logL[l] = 0;
    for (int s = 0; s < MAX_SETS; s++) fin.r[s][l] = nullptr;
  }
  dim3 fgrid2(nsetsB, (unsigned)nsets);
  k_reduce_final<F><<<fgrid2, 128, 128 * sizeof(XYZZ<F>), stream>>>(fin);
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
