// MSM front end shared by every point set over the same scalars: signed-digit decomposition, radix sort of
// (bucket, point reference) pairs, bucket boundaries and length-balanced work items.
//
//   k_msm_digits       one thread per scalar: window digits -> keys[w*n+i] (bucket), vals[w*n+i] (point | sign)
//   CUB radix sort     pairs by bucket key (only the significant key bits)
//   k_bucket_bounds    start[b] = first sorted position with key >= b
//   k_bucket_chunks    a bucket longer than T additions is split into ceil(len/T) work items
//   CUB exclusive sum  item_start[b]
//   k_make_items       item -> bucket, sort key = T - length (longest first); multi-item buckets are listed
//   CUB radix sort     items by length so the 32 lanes of a warp run equally long loops
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "msm.cuh"
#include "msm_digits.cuh"
#include <stdlib.h>

#ifndef G16_MSM_TREE_DEFAULT
#define G16_MSM_TREE_DEFAULT 0
#endif

namespace g16 {

int msm_pick_window(size_t n, bool precomp) {
  double best = 1e300;
  int best_c = 4;
  for (int c = 4; c <= 22; c++) {
    double W = (double)msm_num_windows(c);
    double nb = (double)((size_t)1 << (c - 1));
    double sets = precomp ? 1.0 : W;
    // mixed additions (10 modmul) per pair + running sums (2 full additions of 14 modmul per bucket)
    // + sort / launch overheads that grow with the pair and bucket counts
    double cost = 10.0 * W * (double)n + 28.0 * sets * nb + 0.6 * W * (double)n + 4000.0 * sets;
    if (!precomp && c > 18) continue;
    if (W * (double)n >= 2147483648.0) continue;   // point references are 31 bits
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

static int default_tree_log() {
  static int v = -1;
  if (v < 0) {
#ifdef G16_EXPERIMENTS
    const char* e = getenv("G16_MSM_TREE");
    v = e ? atoi(e) : G16_MSM_TREE_DEFAULT;
    if (v != 0 && (v < 3 || v > MSM_TREE_MAX_LOG)) v = 0;
#else
    v = 0;                                           // the tree mode is not compiled into the default library
#endif
  }
  return v;
}

MsmGeometry msm_geometry(size_t n, int c, bool precomp, int tree_log) {
  G16_REQUIRE(n < ((size_t)1 << 31), "MSM size must be below 2^31");
  G16_REQUIRE(c >= 2 && c <= 22, "MSM window must be 2..22 bits");
  MsmGeometry g;
  g.n = n;
  g.c = c;
  g.nwin = msm_num_windows(c);
  g.precomp = precomp;
  g.nb = 1u << (c - 1);
  g.nbuckets = precomp ? g.nb : (uint32_t)g.nwin * g.nb;
  g.m = (size_t)g.nwin * n;
  G16_REQUIRE(g.m < ((size_t)1 << 31), "MSM pair count must fit 31 bits");
  // A work item is at most T additions.  Items are sorted by length so lanes are balanced for any T; T bounds
  // the serial latency of one thread (~11 us per G1 addition with the SM fully occupied): a launch processes
  // m additions on ~75k resident threads, so an item longer than about m / 150k additions would outlive the
  // rest of the kernel.  It must stay above twice the mean bucket length so that regular buckets are not
  // split.  Heavier buckets -- the top window of a 254-bit scalar has few significant bits and concentrates
  // its n digits on 2^(254 mod c) buckets, and real witnesses are full of 0/1/small values -- become several
  // items whose partial sums are merged by k_bucket_fixup_small / k_bucket_fixup.
  size_t avg = g.m / g.nbuckets + 1;
  size_t T = g.m / 150000;
  if (T < 2 * avg + 16) T = 2 * avg + 16;
  if (T > 32768) T = 32768;
  // Few buckets (a small window: the pieces of a sharded proof, mid-size MSMs): one item per bucket would launch fewer
  // threads than the GPU holds for a few waves (4 CTAs x 128 threads x 148 SMs = 76k), and the kernel then runs at
  // ~64 % of its rate (measured: 2^18-point shards at c = 17).  Regular buckets are then split as well, into items of
  // m / 300k additions (at least 16, so the extra full additions of the fixup stay below ~9 % of the bucket's work).
  static int target_items = -1;
  if (target_items < 0) {
    const char* e = getenv("G16_MSM_TARGET_ITEMS");
    target_items = e ? atoi(e) : 300000;
  }
  if (target_items > 0 && g.nbuckets < (uint32_t)target_items) {
    size_t t2 = g.m / (size_t)target_items;
    if (t2 < 16) t2 = 16;
    if (t2 < T) T = t2;
  }
  g.tree_log = tree_log < 0 ? default_tree_log() : tree_log;
  if (g.tree_log) T = (size_t)1 << g.tree_log;   // chunks of the batched-affine tree (msm_tree.cuh)
  g.T = (uint32_t)T;
  g.max_items = g.nbuckets + (uint32_t)(g.m / T) + 1;
  return g;
}

// ---------------------------------------------------------------------------------------
static __device__ __forceinline__ Fr ld_scalar(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

__global__ void k_msm_digits(const Fr* __restrict__ scalars, uint32_t n, int mont, int c, int nwin, uint32_t nb,
                             int precomp, uint32_t key_none, uint32_t* __restrict__ keys,
                             uint32_t* __restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = ld_scalar(scalars + i);
  if (mont) s = from_mont(s);                    // msm.nim:44 toBig()
  int carry = 0;
  for (int w = 0; w < nwin; w++) {
    int d = msm_signed_digit(s.v, c, w, nwin, carry);
    uint32_t ref = precomp ? (uint32_t)w * n + i : i;
    uint32_t base = precomp ? 0u : (uint32_t)w * nb;
    uint32_t key = key_none;
    if (d > 0) key = base + (uint32_t)(d - 1);
    else if (d < 0) {
      key = base + (uint32_t)(-d - 1);
      ref |= 0x80000000u;
    }
    keys[(size_t)w * n + i] = key;
    vals[(size_t)w * n + i] = ref;
  }
}

__global__ void k_bucket_bounds(const uint32_t* __restrict__ keys, uint32_t m, uint32_t nbuckets,
                                uint32_t* __restrict__ start) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nbuckets) return;
  uint32_t lo = 0, hi = m;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (keys[mid] < b) lo = mid + 1;
    else hi = mid;
  }
  start[b] = lo;
}

__global__ void k_bucket_chunks(const uint32_t* __restrict__ start, uint32_t nbuckets, uint32_t T,
                                uint32_t* __restrict__ chunks) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nbuckets) return;
  uint32_t len = b < nbuckets ? start[b + 1] - start[b] : 0;
  chunks[b] = (len + T - 1) / T;
}

__global__ void k_items_init(uint32_t* __restrict__ key, uint32_t* __restrict__ idx, uint32_t max_items, uint32_t T,
                             uint32_t* __restrict__ multi_count) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) {
    multi_count[0] = 0;
    multi_count[1] = 0;
  }
  if (t >= max_items) return;
  key[t] = T;        // padding sorts after every real item (real keys are T - len <= T - 1)
  idx[t] = t;
}

__global__ void k_make_items(const uint32_t* __restrict__ start, const uint32_t* __restrict__ item_start,
                             uint32_t nbuckets, uint32_t T, uint32_t* __restrict__ item_bucket,
                             uint32_t* __restrict__ item_key, uint32_t* __restrict__ multi) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  uint32_t i0 = item_start[b], i1 = item_start[b + 1];
  if (i1 == i0) return;
  uint32_t len = start[b + 1] - start[b];
  for (uint32_t k = 0; i0 + k < i1; k++) {
    uint32_t l = len - k * T;
    if (l > T) l = T;
    item_bucket[i0 + k] = b;
    item_key[i0 + k] = T - l;
  }
  if (i1 - i0 > 1) {           // layout of `multi`: [count_small, count_big, small[nbuckets], big[nbuckets]]
    if (i1 - i0 <= MSM_FIXUP_SMALL_MAX) {
      uint32_t slot = atomicAdd(multi, 1u);
      multi[2 + slot] = b;
    } else {
      uint32_t slot = atomicAdd(multi + 1, 1u);
      multi[2 + nbuckets + slot] = b;
    }
  }
}

#ifdef G16_EXPERIMENTS
// Lists of the additions of every tree round (msm_tree.cuh): slot j with chunk-relative position q is a left
// operand of round r when q is a multiple of 2^(r+1) and q + 2^r is still inside the chunk.  Round 0 also lists
// the last element of odd-length chunks (bit 31: no partner) so that it is copied into the working array.
constexpr int TREE_LIST_TPB = 1024;
constexpr int TREE_LIST_SPT = 4;      // slots per thread
struct TreeListOffsets {
  size_t off[MSM_TREE_MAX_LOG];
};
__global__ void __launch_bounds__(TREE_LIST_TPB) k_tree_lists(const uint32_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ start, uint32_t m,
                                                              uint32_t nbuckets, int tree_log,
                                                              uint32_t* __restrict__ lists, TreeListOffsets offs,
                                                              uint32_t* __restrict__ cnt) {
  // one global atomic per block and round: counts per warp -> block scan in shared memory -> base
  __shared__ uint32_t warp_cnt[MSM_TREE_MAX_LOG][TREE_LIST_TPB / 32];
  __shared__ uint32_t block_base[MSM_TREE_MAX_LOG];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t T = 1u << tree_log;
  uint32_t q[TREE_LIST_SPT], lc[TREE_LIST_SPT];
  const uint32_t j0 = (blockIdx.x * TREE_LIST_TPB + threadIdx.x) * TREE_LIST_SPT;
#pragma unroll
  for (int k = 0; k < TREE_LIST_SPT; k++) {
    const uint32_t j = j0 + k;
    q[k] = 1;        // odd position with an empty chunk: never listed
    lc[k] = 0;
    if (j < m) {
      uint32_t b = keys[j];
      if (b < nbuckets) {
        uint32_t s0 = start[b];
        uint32_t rel = j - s0, len = start[b + 1] - s0;
        q[k] = rel & (T - 1u);
        lc[k] = len - (rel - q[k]);
        if (lc[k] > T) lc[k] = T;
      }
    }
  }
  // pass 1: counts
  uint32_t mine[MSM_TREE_MAX_LOG];
#pragma unroll
  for (int r = 0; r < MSM_TREE_MAX_LOG; r++) {
    uint32_t c = 0;
    if (r < tree_log) {
#pragma unroll
      for (int k = 0; k < TREE_LIST_SPT; k++) {
        bool left = lc[k] && (q[k] & ((2u << r) - 1u)) == 0;
        c += (left && (r == 0 || q[k] + (1u << r) < lc[k])) ? 1u : 0u;
      }
    }
    mine[r] = c;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((int)lane >= d) incl += o;
    }
    if (lane == 31) warp_cnt[r][warp] = incl;
    mine[r] = incl - c;     // exclusive prefix inside the warp
  }
  __syncthreads();
  if (threadIdx.x < (unsigned)tree_log) {
    const int r = threadIdx.x;
    uint32_t run = 0;
    for (int w = 0; w < TREE_LIST_TPB / 32; w++) {
      uint32_t c = warp_cnt[r][w];
      warp_cnt[r][w] = run;
      run += c;
    }
    block_base[r] = run ? atomicAdd(cnt + r, run) : 0u;
  }
  __syncthreads();
  // pass 2: write
  for (int r = 0; r < tree_log; r++) {
    uint32_t pos = block_base[r] + warp_cnt[r][warp] + mine[r];
    uint32_t* out = lists + offs.off[r];
#pragma unroll
    for (int k = 0; k < TREE_LIST_SPT; k++) {
      bool left = lc[k] && (q[k] & ((2u << r) - 1u)) == 0;
      bool pair = left && (q[k] + (1u << r) < lc[k]);
      if (pair || (r == 0 && left)) out[pos++] = (j0 + k) | (pair ? 0u : 0x80000000u);
    }
  }
}

#endif

MsmSorter::~MsmSorter() {}

size_t MsmSorter::workspace_bytes() const {
  size_t t = 0;
  const DevBuf* all[] = {&keys_[0], &keys_[1], &vals_[0], &vals_[1], &start_, &chunks_, &item_start_, &item_bucket_,
                         &item_key_[0], &item_key_[1], &item_idx_[0], &item_idx_[1], &multi_, &cub_tmp_,
                         &tree_list_, &tree_cnt_};
  for (auto* b : all) t += b->bytes;
  return t;
}

static int bits_for(uint64_t max_value) {
  int b = 1;
  while (((uint64_t)1 << b) <= max_value) b++;
  return b;
}

void MsmSorter::run(const Fr* scalars, bool scalars_mont, const MsmGeometry& g, cudaStream_t stream) {
  g_ = g;
  G16_REQUIRE(g.n > 0, "MsmSorter: empty input");
  const size_t m = g.m;
  keys_[0].ensure(m * 4);
  keys_[1].ensure(m * 4);
  vals_[0].ensure(m * 4);
  vals_[1].ensure(m * 4);
  start_.ensure(((size_t)g.nbuckets + 2) * 4);
  chunks_.ensure(((size_t)g.nbuckets + 2) * 4);
  item_start_.ensure(((size_t)g.nbuckets + 2) * 4);
  item_bucket_.ensure((size_t)g.max_items * 4);
  for (int i = 0; i < 2; i++) {
    item_key_[i].ensure((size_t)g.max_items * 4);
    item_idx_[i].ensure((size_t)g.max_items * 4);
  }
  multi_.ensure((2 * (size_t)g.nbuckets + 4) * 4);

  k_msm_digits<<<div_up(g.n, 256), 256, 0, stream>>>(scalars, (uint32_t)g.n, scalars_mont ? 1 : 0, g.c, g.nwin, g.nb,
                                                     g.precomp ? 1 : 0, g.nbuckets, keys_[0].as<uint32_t>(),
                                                     vals_[0].as<uint32_t>());
  G16_LAUNCH_CHECK();

  // temp storage: the larger of the three CUB calls
  int key_bits = bits_for(g.nbuckets);
  int item_bits = bits_for(g.T);
  size_t t1 = 0, t2 = 0, t3 = 0;
  G16_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, keys_[0].as<uint32_t>(), keys_[1].as<uint32_t>(),
                                           vals_[0].as<uint32_t>(), vals_[1].as<uint32_t>(), (int64_t)m, 0, key_bits,
                                           stream));
  G16_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, chunks_.as<uint32_t>(), item_start_.as<uint32_t>(),
                                         (int)(g.nbuckets + 1), stream));
  G16_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t3, item_key_[0].as<uint32_t>(), item_key_[1].as<uint32_t>(),
                                           item_idx_[0].as<uint32_t>(), item_idx_[1].as<uint32_t>(),
                                           (int64_t)g.max_items, 0, item_bits, stream));
  size_t tmp = t1 > t2 ? t1 : t2;
  if (t3 > tmp) tmp = t3;
  cub_tmp_.ensure(tmp);

  G16_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp_.p, t1, keys_[0].as<uint32_t>(), keys_[1].as<uint32_t>(),
                                           vals_[0].as<uint32_t>(), vals_[1].as<uint32_t>(), (int64_t)m, 0, key_bits,
                                           stream));
  k_bucket_bounds<<<div_up((size_t)g.nbuckets + 1, 256), 256, 0, stream>>>(keys_[1].as<uint32_t>(), (uint32_t)m,
                                                                           g.nbuckets, start_.as<uint32_t>());
  G16_LAUNCH_CHECK();
  k_bucket_chunks<<<div_up((size_t)g.nbuckets + 1, 256), 256, 0, stream>>>(start_.as<uint32_t>(), g.nbuckets, g.T,
                                                                           chunks_.as<uint32_t>());
  G16_LAUNCH_CHECK();
  G16_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp_.p, t2, chunks_.as<uint32_t>(), item_start_.as<uint32_t>(),
                                         (int)(g.nbuckets + 1), stream));
  k_items_init<<<div_up(g.max_items, 256), 256, 0, stream>>>(item_key_[0].as<uint32_t>(), item_idx_[0].as<uint32_t>(),
                                                             g.max_items, g.T, multi_.as<uint32_t>());
  G16_LAUNCH_CHECK();
  k_make_items<<<div_up(g.nbuckets, 256), 256, 0, stream>>>(start_.as<uint32_t>(), item_start_.as<uint32_t>(),
                                                            g.nbuckets, g.T, item_bucket_.as<uint32_t>(),
                                                            item_key_[0].as<uint32_t>(), multi_.as<uint32_t>());
  G16_LAUNCH_CHECK();
  if (!g.tree_log) {
    G16_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp_.p, t3, item_key_[0].as<uint32_t>(), item_key_[1].as<uint32_t>(),
                                             item_idx_[0].as<uint32_t>(), item_idx_[1].as<uint32_t>(),
                                             (int64_t)g.max_items, 0, item_bits, stream));
    return;
  }
#ifdef G16_EXPERIMENTS
  // batched-affine tree: per-round addition lists
  size_t total = 0;
  for (int r = 0; r < MSM_TREE_MAX_LOG; r++) {
    tree_off_[r] = total;
    tree_cap_[r] = r < g.tree_log ? (uint32_t)((m >> (r + 1)) + g.max_items + 32) : 0;
    total += tree_cap_[r];
  }
  tree_list_.ensure(total * 4);
  tree_cnt_.ensure(MSM_TREE_MAX_LOG * 4);
  G16_CUDA(cudaMemsetAsync(tree_cnt_.p, 0, MSM_TREE_MAX_LOG * 4, stream));
  TreeListOffsets offs;
  for (int r = 0; r < MSM_TREE_MAX_LOG; r++) offs.off[r] = tree_off_[r];
  k_tree_lists<<<div_up(m, TREE_LIST_TPB * TREE_LIST_SPT), TREE_LIST_TPB, 0, stream>>>(
      keys_[1].as<uint32_t>(), start_.as<uint32_t>(), (uint32_t)m, g.nbuckets, g.tree_log,
      tree_list_.as<uint32_t>(), offs, tree_cnt_.as<uint32_t>());
  G16_LAUNCH_CHECK();
#else
  G16_REQUIRE(false, "batched-affine tree mode needs a library built with `make EXPERIMENTS=1`");
#endif
}

}  // namespace g16
