// Library-internal interface of the MSM engine: GPU Pippenger over G1 (Fp) and G2 (Fp2).
// Replaces groth16/bn128/msm.nim:35-59,63-83 (msmConstantineG1/G2 -> constantine multiScalarMul*)
// and the thread chunking of msm.nim:89-158.
//
// Two layers:
//   MsmSorter        scalars -> signed window digits -> radix sort by bucket -> bucket bounds ->
//                    length-balanced work items.  Depends only on the scalars, so the four witness MSMs of a
//                    proof (A1, B1, B2, C1) share one sorter run.
//   MsmAccumulator   sorted pairs + one point set -> bucket sums -> bucket reduction -> one XYZZ point.
// Two point layouts:
//   plain            n affine points; every window has its own 2^(c-1) buckets; Horner over the windows.
//   precomputed      a table of W*n affine points, table[w*n + i] = 2^(c*w) * P_i, resident in HBM; all
//                    windows share ONE bucket set, so there is no window combination and c can be larger.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "ec.cuh"
#include "msm_digits.cuh"

namespace g16 {

struct MsmGeometry {
  size_t n = 0;            // scalars / points
  int c = 0;               // window bits
  int nwin = 0;            // windows = ceil(255 / c)
  bool precomp = false;
  uint32_t nb = 0;         // buckets per window = 2^(c-1)
  uint32_t nbuckets = 0;   // bucket sets * nb  (precomp: nb; plain: nwin * nb)
  size_t m = 0;            // (scalar, window) pairs = nwin * n
  uint32_t T = 0;          // maximum additions per work item
  uint32_t max_items = 0;  // upper bound on work items
  int tree_log = 0;        // 0: XYZZ work items; 3..6: batched-affine tree over chunks of 2^tree_log slots (T)
};
constexpr int MSM_TREE_MAX_LOG = 6;

constexpr uint32_t MSM_FIXUP_SMALL_MAX = 16;   // buckets with at most this many items are merged by one thread

int msm_pick_window(size_t n, bool precomp);
// tree_log < 0: take the library default (environment G16_MSM_TREE, 0 = XYZZ work items)
MsmGeometry msm_geometry(size_t n, int c, bool precomp, int tree_log = -1);

class MsmSorter {
 public:
  MsmSorter() {}
  ~MsmSorter();
  // scalars: n x 32 bytes on the device; scalars_mont: Montgomery residues (the reference's in-memory Fr,
  // msm.nim:42-44 toBig) or standard-form integers (.wtns bytes).
  void run(const Fr* scalars, bool scalars_mont, const MsmGeometry& g, cudaStream_t stream);
  const MsmGeometry& geom() const { return g_; }
  size_t workspace_bytes() const;
  // device views, valid after run() (stream order)
  const uint32_t* vals() const { return vals_[1].as<uint32_t>(); }
  const uint32_t* start() const { return start_.as<uint32_t>(); }
  const uint32_t* item_start() const { return item_start_.as<uint32_t>(); }
  const uint32_t* item_bucket() const { return item_bucket_.as<uint32_t>(); }
  const uint32_t* items_sorted() const { return item_idx_[1].as<uint32_t>(); }
  // buckets split into several items: [count_small, count_big, small[nbuckets], big[nbuckets]]
  const uint32_t* multi() const { return multi_.as<uint32_t>(); }
  // batched-affine tree (geom().tree_log > 0): per round r the list of left slots (bit 31: no partner, round 0
  // only) and its length; the lists depend only on the bucket structure
  const uint32_t* tree_list(int r) const { return tree_list_.as<uint32_t>() + tree_off_[r]; }
  const uint32_t* tree_count(int r) const { return tree_cnt_.as<uint32_t>() + r; }
  uint32_t tree_cap(int r) const { return tree_cap_[r]; }

 private:
  MsmGeometry g_;
  DevBuf keys_[2], vals_[2], start_, chunks_, item_start_, item_bucket_, item_key_[2], item_idx_[2], multi_, cub_tmp_;
  DevBuf tree_list_, tree_cnt_;
  size_t tree_off_[MSM_TREE_MAX_LOG] = {0, 0, 0, 0, 0, 0};
  uint32_t tree_cap_[MSM_TREE_MAX_LOG] = {0, 0, 0, 0, 0, 0};
};

template <class F>
struct MsmPointSet {
  const Affine<F>* points = nullptr;   // plain: n points; precomputed: nwin * n table
  XYZZ<F>* result = nullptr;           // device, one XYZZ point
};

template <class F>
class MsmAccumulator {
 public:
  static constexpr int MAX_SETS = 3;
  MsmAccumulator() {}
  ~MsmAccumulator();
  // up to MAX_SETS point sets over the same sorted pairs, accumulated in the same launches.
  // `tail`: optional second stream (created with a HIGHER priority than `stream`) for the latency-bound end of the
  // bucket reduction -- a handful of one-block kernels.  On the same stream (or one of equal priority) their blocks
  // queue behind the thousands of pending blocks of whatever accumulation kernel another proof in flight has
  // launched, and the tail of one proof no longer overlaps with the accumulation of the next.  With a tail stream
  // the results are complete when `tail` reaches this point, `stream` is not held back.
  void run(const MsmSorter& sorter, const MsmPointSet<F>* sets, int nsets, cudaStream_t stream,
           cudaStream_t tail = nullptr);
  size_t workspace_bytes() const;
  bool profile = false;                // CUDA events around the accumulate kernel on `stream`
  float last_accum_ms() const;

 private:
  void run_tree(const MsmSorter& sorter, const MsmPointSet<F>* sets, int nsets, XYZZ<F>* const* buckets,
                cudaStream_t stream);
  DevBuf buckets_, partials_, winpart_;
  DevBuf tree_w_, tree_m_, tree_nodes_, tree_bp_;
  cudaEvent_t pev_[2] = {nullptr, nullptr};
  cudaEvent_t lev_ = nullptr;          // level 1 of the reduction done (the tail stream waits for it)
};

// The reference-shaped single MSM (plain layout): sorter + accumulator.
template <class F>
class Msm {
 public:
  // result (XYZZ, device memory) = sum_i scalars[i] * points[i]
  void run(const Fr* scalars, bool scalars_mont, const Affine<F>* points, size_t n, XYZZ<F>* result,
           cudaStream_t stream, int window_bits = 0);
  // same with a precomputed table built by msm_build_table (window bits fixed by the table)
  void run_precomp(const Fr* scalars, bool scalars_mont, const Affine<F>* table, size_t n, int window_bits,
                   XYZZ<F>* result, cudaStream_t stream);
  size_t workspace_bytes() const { return sorter_.workspace_bytes() + acc_.workspace_bytes(); }
  void set_profile(bool on) { profile_ = on; acc_.profile = on; }
  float last_accum_ms() const { return acc_.last_accum_ms(); }
  float last_total_ms() const;
  uint64_t last_pairs = 0;
  int last_c = 0, last_nwin = 0;
  ~Msm();

 private:
  void go(const Fr* scalars, bool mont, const Affine<F>* pts, const MsmGeometry& g, XYZZ<F>* result, cudaStream_t s);
  MsmSorter sorter_;
  MsmAccumulator<F> acc_;
  bool profile_ = false;
  cudaEvent_t tev_[2] = {nullptr, nullptr};
};

// table[w * n + i] = 2^(c*w) * points[i] for w < ceil(255/c), affine (infinity stays (0,0)).
// `pad_front` leading entries of every window are written as infinity (used to align the C1 points with
// witness indices: prover.nim:262-264 zs = witness[npubs+1 ..]).
template <class F>
void msm_build_table(const Affine<F>* points, size_t n_points, size_t pad_front, int c, Affine<F>* table,
                     cudaStream_t stream);

// parts[0..count) summed and normalised to affine (infinity -> (0,0)); one tiny kernel.
template <class F>
void xyzz_sum_to_affine(const XYZZ<F>* parts, int count, Affine<F>* out, cudaStream_t stream);

extern template class MsmAccumulator<Fp>;
extern template class MsmAccumulator<Fp2>;
extern template class Msm<Fp>;
extern template class Msm<Fp2>;

}  // namespace g16
