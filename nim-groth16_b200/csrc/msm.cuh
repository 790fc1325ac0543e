// Library-internal interface of msm.cu: GPU Pippenger over G1 (Fp) and G2 (Fp2).
// Replaces groth16/bn128/msm.nim:35-59,63-83 (msmConstantineG1/G2 -> constantine multiScalarMul*)
// and the thread chunking of msm.nim:89-158.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "ec.cuh"

namespace g16 {

struct MsmConfig {
  int c = 0;          // window bits (0 = pick from n)
};

template <class F>
class Msm {
 public:
  Msm() {}
  // result (XYZZ, device memory) = sum_i scalars[i] * points[i].
  // scalars: n x 32 bytes on the device; `scalars_mont` says whether they are Montgomery residues
  // (the reference's in-memory Fr, msm.nim:42-44 toBig) or standard-form integers (.wtns bytes).
  void run(const Fr* scalars, bool scalars_mont, const Affine<F>* points, size_t n, XYZZ<F>* result,
           cudaStream_t stream, const MsmConfig& cfg = MsmConfig());
  // device memory currently held by the workspace
  size_t workspace_bytes() const;
  int last_c = 0, last_nwin = 0;
  // optional per-kernel timing of the dominant kernel (bucket accumulation), CUDA events on `stream`
  bool profile = false;
  float last_accum_ms() const;     // valid after the stream has been synchronised
  float last_total_ms() const;
  uint64_t last_pairs = 0;         // non-zero (scalar, window) digits of the last run (filled when profiling)
  ~Msm();

 private:
  DevBuf keys_[2], vals_[2], start_, buckets_, winpart_, cub_tmp_;
  cudaEvent_t pev_[4] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t nbuckets_last_ = 0;
};

// parts[0..count) summed and normalised to affine (infinity -> (0,0)); one tiny kernel.
template <class F>
void xyzz_sum_to_affine(const XYZZ<F>* parts, int count, Affine<F>* out, cudaStream_t stream);

// affine partial sums (one per shard, msm.nim:117-119) -> XYZZ sum
template <class F>
void affine_sum_to_xyzz(const Affine<F>* parts, int count, XYZZ<F>* out, cudaStream_t stream);

int msm_pick_window(size_t n, bool g2);

extern template class Msm<Fp>;
extern template class Msm<Fp2>;

}  // namespace g16
