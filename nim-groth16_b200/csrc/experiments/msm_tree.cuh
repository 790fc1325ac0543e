// Batched-affine bucket accumulation ("tree" mode of MsmAccumulator).
//
// The XYZZ accumulate kernel runs at the integer-pipe roof, so the only way to make bucket accumulation faster
// is to spend fewer multiplications per addition.  An affine addition costs 1 inversion + 3 multiplications;
// with Montgomery's trick the inversion is shared by a whole launch and becomes 3 more multiplications:
// ~6.4 modular multiplications per addition instead of 10 (G1), ~18 Fp multiplications instead of 28 (G2).
//
// The sorted pair array is cut into chunks of 2^tree_log slots relative to each bucket's start (the work items
// of MsmSorter with T = 2^tree_log).  Round r adds, inside every chunk, the element at relative position q to
// the one at q + 2^r for all q divisible by 2^(r+1): after tree_log rounds the head slot of a chunk holds the
// chunk's sum.  Every addition of a round is independent of the others, so a round is three launches:
//   k_tree_prepare  thread = 8 additions: denominators d_i (x2 - x1, or 2y for a doubling, or 1 when one side is
//                   infinity / the sum is infinity), exclusive prefix products -> M, thread totals -> block
//                   product tree -> BP[block]
//   k_tree_invert   one block per point set: all block products inverted with ONE field inversion
//   k_tree_finish   block tree walked down -> 1/thread total -> 1/d_i -> lambda, x3, y3 -> W[left slot]
// Round 0's prepare gathers the points through the sorted references (window table, sign applied) and stores
// them in the working array W; everything after that reads and writes W in place.  The lists of additions per round depend only on the bucket
// structure and are built once per sorter run (MsmSorter::tree_*), shared by all point sets.
// Buckets made of one chunk are converted to XYZZ by k_tree_finalize; longer buckets (skewed scalars) add their
// chunk heads with mixed additions: k_tree_fixup_small (one thread) / k_tree_fixup_big (one block).
#pragma once
#include "msm.cuh"

namespace g16 {

constexpr int TREE_K = 8;          // additions per thread
constexpr int TREE_TPB = 128;      // threads per block
constexpr int TREE_PER_BLOCK = TREE_K * TREE_TPB;

template <class F>
struct TreeSets {
  const Affine<F>* points[MsmAccumulator<F>::MAX_SETS];
  Affine<F>* W[MsmAccumulator<F>::MAX_SETS];      // m slots
  F* M[MsmAccumulator<F>::MAX_SETS];              // prefix products, one per list entry of the round
  F* tree[MsmAccumulator<F>::MAX_SETS];           // 2*TREE_TPB nodes per block
  F* bp[MsmAccumulator<F>::MAX_SETS];             // block products
  F* ibp[MsmAccumulator<F>::MAX_SETS];            // their inverses
  XYZZ<F>* buckets[MsmAccumulator<F>::MAX_SETS];
};

template <class F>
__device__ __forceinline__ Affine<F> tree_load(const Affine<F>* __restrict__ points, const Affine<F>* W,
                                               const uint32_t* __restrict__ vals, int r, uint32_t slot) {
  if (r == 0) {
    uint32_t v = vals[slot];
    Affine<F> p = ldg_vec(points + (v & 0x7fffffffu));
    if (v & 0x80000000u) p.y = fneg(p.y);
    return p;
  }
  return ld_vec(W + slot);
}

// kind: 0 = generic addition (d = x2 - x1), 1 = doubling (d = 2 y1), 2 = result is p1, 3 = result is p2,
// 4 = result is infinity; d = 1 for kinds 2..4 so that the shared product stays invertible
template <class F>
__device__ __forceinline__ int tree_pair_op(const Affine<F>& p1, const Affine<F>& p2, bool has2, F& d) {
  if (!has2 || aff_is_inf(p2)) return 2;
  if (aff_is_inf(p1)) return 3;
  d = fsub(p2.x, p1.x);
  if (!fis_zero(d)) return 0;
  if (feq(p1.y, p2.y) && !fis_zero(p1.y)) {
    d = fdbl(p1.y);
    return 1;
  }
  return 4;
}

// rare operand patterns (infinity, equal x) of a listed pair, out of line to keep the hot loops small
template <class F>
__device__ __noinline__ int tree_pair_op_ni(const Affine<F>* a1, const Affine<F>* a2, F& d) {
  Affine<F> p1 = ld_vec(a1), p2 = ld_vec(a2);
  return tree_pair_op(p1, p2, true, d);
}

// Round 0 gathers the operands through the sorted references exactly once: the (sign-applied) points are
// written to their slots of W, so that k_tree_finish and the later rounds stream W instead of gathering again.
template <class F, bool ROUND0>
__global__ void __launch_bounds__(TREE_TPB) k_tree_prepare(TreeSets<F> ts, const uint32_t* __restrict__ vals,
                                                           const uint32_t* __restrict__ list,
                                                           const uint32_t* __restrict__ count, int r) {
  __shared__ F nodes[2 * TREE_TPB];
  const uint32_t n = *count;
  if (blockIdx.x * TREE_PER_BLOCK >= n) return;
  const int set = blockIdx.y;
  const Affine<F>* __restrict__ points = ts.points[set];
  Affine<F>* W = ts.W[set];
  const uint32_t e0 = (blockIdx.x * TREE_TPB + threadIdx.x) * TREE_K;
  F run = F::one();
#pragma unroll 2
  for (int i = 0; i < TREE_K; i++) {
    const uint32_t e = e0 + i;
    if (e >= n) break;
    const uint32_t ent = list[e];
    const uint32_t slot = ent & 0x7fffffffu;
    const bool has2 = !(ent >> 31);
    F d = F::one();
    int kind = 2;
    if (ROUND0) {
      uint32_t v1 = vals[slot];
      Affine<F> p1 = ldg_vec(points + (v1 & 0x7fffffffu));
      if (v1 & 0x80000000u) p1.y = fneg(p1.y);
      st_vec(W + slot, p1);
      if (has2) {
        uint32_t v2 = vals[slot + 1];
        Affine<F> p2 = ldg_vec(points + (v2 & 0x7fffffffu));
        if (v2 & 0x80000000u) p2.y = fneg(p2.y);
        st_vec(W + slot + 1, p2);
        kind = tree_pair_op(p1, p2, true, d);
      }
    } else {
      // only the x coordinates (one sector each); anything unusual reloads the full points
      const Affine<F>* a1 = W + slot;
      const Affine<F>* a2 = W + slot + (1u << r);
      F x1 = ld_vec(&a1->x), x2 = ld_vec(&a2->x);
      d = fsub(x2, x1);
      kind = 0;
      if (fis_zero(d) || fis_zero(x1) || fis_zero(x2)) kind = tree_pair_op_ni(a1, a2, d);
    }
    ts.M[set][e] = run;
    if (kind < 2) run = fmul(run, d);
  }
  nodes[TREE_TPB + threadIdx.x] = run;
  __syncthreads();
  for (int s = TREE_TPB >> 1; s >= 1; s >>= 1) {
    if ((int)threadIdx.x < s) nodes[s + threadIdx.x] = fmul(nodes[2 * (s + threadIdx.x)], nodes[2 * (s + threadIdx.x) + 1]);
    __syncthreads();
  }
  F* tree = ts.tree[set] + (size_t)blockIdx.x * (2 * TREE_TPB);
  tree[threadIdx.x] = nodes[threadIdx.x];
  tree[TREE_TPB + threadIdx.x] = nodes[TREE_TPB + threadIdx.x];
  if (threadIdx.x == 0) ts.bp[set][blockIdx.x] = nodes[1];
}

// ibp[i] = 1 / bp[i] for i < ceil(count / TREE_PER_BLOCK): per-thread prefix products over a contiguous segment,
// a product tree over the threads, one inversion, and the way back.  grid.x = point sets.
template <class F>
__global__ void __launch_bounds__(TREE_TPB) k_tree_invert(TreeSets<F> ts, const uint32_t* __restrict__ count) {
  __shared__ F nodes[2 * TREE_TPB];
  __shared__ F inv[2 * TREE_TPB];
  const uint32_t n = *count;
  const uint32_t nblk = (n + TREE_PER_BLOCK - 1) / TREE_PER_BLOCK;
  if (nblk == 0) return;
  const int set = blockIdx.x;
  const F* bp = ts.bp[set];
  F* ibp = ts.ibp[set];
  const uint32_t seg = (nblk + TREE_TPB - 1) / TREE_TPB;
  const uint32_t i0 = threadIdx.x * seg;
  uint32_t i1 = i0 + seg;
  if (i1 > nblk) i1 = nblk;
  F run = F::one();
  for (uint32_t i = i0; i < i1; i++) {
    ibp[i] = run;
    run = fmul(run, bp[i]);
  }
  nodes[TREE_TPB + threadIdx.x] = run;
  __syncthreads();
  for (int s = TREE_TPB >> 1; s >= 1; s >>= 1) {
    if ((int)threadIdx.x < s) nodes[s + threadIdx.x] = fmul(nodes[2 * (s + threadIdx.x)], nodes[2 * (s + threadIdx.x) + 1]);
    __syncthreads();
  }
  if (threadIdx.x == 0) inv[1] = finv(nodes[1]);
  __syncthreads();
  for (int s = 1; s < TREE_TPB; s <<= 1) {          // children 2s .. 4s-1
    for (int c = 2 * s + threadIdx.x; c < 4 * s; c += TREE_TPB) inv[c] = fmul(inv[c >> 1], nodes[c ^ 1]);
    __syncthreads();
  }
  F inv_run = inv[TREE_TPB + threadIdx.x];
  for (uint32_t i = i1; i > i0; i--) {
    F prefix = ibp[i - 1];
    ibp[i - 1] = fmul(inv_run, prefix);
    inv_run = fmul(inv_run, bp[i - 1]);
  }
}

// everything except the generic addition, out of line
template <class F>
__device__ __noinline__ void tree_finish_rare(const Affine<F>& p1, const Affine<F>& p2, int kind, const F& inv_d,
                                              Affine<F>& out) {
  if (kind == 1) {
    F x2 = fsqr(p1.x);
    F lam = fmul(fadd(fdbl(x2), x2), inv_d);
    out.x = fsub(fsqr(lam), fdbl(p1.x));
    out.y = fsub(fmul(lam, fsub(p1.x, out.x)), p1.y);
  } else if (kind == 2) {
    out = p1;
  } else if (kind == 3) {
    out = p2;
  } else {
    out = aff_inf<F>();
  }
}

template <class F>
__global__ void __launch_bounds__(TREE_TPB) k_tree_finish(TreeSets<F> ts, const uint32_t* __restrict__ list,
                                                          const uint32_t* __restrict__ count, int r) {
  __shared__ F nodes[2 * TREE_TPB];
  __shared__ F inv[2 * TREE_TPB];
  const uint32_t n = *count;
  if (blockIdx.x * TREE_PER_BLOCK >= n) return;
  const int set = blockIdx.y;
  Affine<F>* W = ts.W[set];
  const F* tree = ts.tree[set] + (size_t)blockIdx.x * (2 * TREE_TPB);
  nodes[threadIdx.x] = tree[threadIdx.x];
  nodes[TREE_TPB + threadIdx.x] = tree[TREE_TPB + threadIdx.x];
  if (threadIdx.x == 0) inv[1] = ts.ibp[set][blockIdx.x];
  __syncthreads();
  for (int s = 1; s < TREE_TPB; s <<= 1) {
    for (int c = 2 * s + threadIdx.x; c < 4 * s; c += TREE_TPB) inv[c] = fmul(inv[c >> 1], nodes[c ^ 1]);
    __syncthreads();
  }
  F inv_run = inv[TREE_TPB + threadIdx.x];
  const uint32_t e0 = (blockIdx.x * TREE_TPB + threadIdx.x) * TREE_K;
  const F* M = ts.M[set];
#pragma unroll 2
  for (int i = TREE_K - 1; i >= 0; i--) {
    const uint32_t e = e0 + i;
    if (e >= n) continue;
    const uint32_t ent = list[e];
    const uint32_t slot = ent & 0x7fffffffu;
    const bool has2 = !(ent >> 31);
    if (!has2) continue;                       // round 0 copy: k_tree_prepare already stored it
    Affine<F> p1 = ld_vec(W + slot);
    Affine<F> p2 = ld_vec(W + slot + (1u << r));
    F d;
    int kind = tree_pair_op(p1, p2, true, d);
    Affine<F> out;
    F inv_d = fmul(inv_run, M[e]);
    if (kind == 0) {
      inv_run = fmul(inv_run, d);
      F lam = fmul(fsub(p2.y, p1.y), inv_d);
      out.x = fsub(fsqr(lam), fadd(p1.x, p2.x));
      out.y = fsub(fmul(lam, fsub(p1.x, out.x)), p1.y);
    } else {
      if (kind == 1) inv_run = fmul(inv_run, d);
      tree_finish_rare(p1, p2, kind, inv_d, out);
    }
    st_vec(W + slot, out);
  }
}

// buckets made of exactly one chunk; grid = (nbuckets / 128, point sets)
template <class F>
__global__ void __launch_bounds__(128) k_tree_finalize(TreeSets<F> ts, const uint32_t* __restrict__ start,
                                                       const uint32_t* __restrict__ item_start, uint32_t nbuckets) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  if (item_start[b + 1] - item_start[b] != 1) return;
  const int set = blockIdx.y;
  Affine<F> p = ld_vec(ts.W[set] + start[b]);
  st_vec(ts.buckets[set] + b, xyzz_from_affine(p));
}

template <class F>
__global__ void __launch_bounds__(128) k_tree_fixup_small(TreeSets<F> ts, const uint32_t* __restrict__ start,
                                                          const uint32_t* __restrict__ item_start,
                                                          const uint32_t* __restrict__ multi, int tree_log) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= multi[0]) return;
  const int set = blockIdx.y;
  const uint32_t b = multi[2 + i];
  const uint32_t nitems = item_start[b + 1] - item_start[b];
  const Affine<F>* W = ts.W[set] + start[b];
  XYZZ<F> acc = xyzz_from_affine(ld_vec(W));
  for (uint32_t k = 1; k < nitems; k++) {
    Affine<F> o = ld_vec(W + ((size_t)k << tree_log));
    xyzz_madd_ni(acc, acc, o);
  }
  st_vec(ts.buckets[set] + b, acc);
}

template <class F>
__global__ void __launch_bounds__(128) k_tree_fixup_big(TreeSets<F> ts, const uint32_t* __restrict__ start,
                                                        const uint32_t* __restrict__ item_start,
                                                        const uint32_t* __restrict__ multi, uint32_t nbuckets,
                                                        int tree_log) {
  extern __shared__ uint4 red_raw[];
  XYZZ<F>* red = reinterpret_cast<XYZZ<F>*>(red_raw);
  const int set = blockIdx.y;
  const uint32_t count = multi[1];
  const uint32_t* lst = multi + 2 + nbuckets;
  for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
    const uint32_t b = lst[i];
    const uint32_t nitems = item_start[b + 1] - item_start[b];
    const Affine<F>* W = ts.W[set] + start[b];
    XYZZ<F> acc = xyzz_inf<F>();
    for (uint32_t k = threadIdx.x; k < nitems; k += blockDim.x) {
      Affine<F> o = ld_vec(W + ((size_t)k << tree_log));
      xyzz_madd_ni(acc, acc, o);
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
    if (threadIdx.x == 0) st_vec(ts.buckets[set] + b, acc);
    __syncthreads();
  }
}

}  // namespace g16
