// Library-internal interface of prover.cu: the resident proving context.
#pragma once
#include <cuda_runtime.h>
#include <memory>
#include "../../include/g16b200.h"
#include "abc.cuh"
#include "common.cuh"
#include "ec.cuh"
#include "msm.cuh"

namespace g16 {

struct alignas(16) PartialsAffine {   // == g16_partials
  G1Affine a1, b1, h1, c1;
  G2Affine b2;
};
static_assert(sizeof(PartialsAffine) == sizeof(g16_partials), "g16_partials layout");

struct alignas(16) MsmResults {       // XYZZ sums of the five MSMs
  G1XYZZ a1, b1, h1, c1;
  G2XYZZ b2;
};

struct alignas(16) MaskTerms {        // prover.nim:279-300, everything that does not depend on an MSM
  G1XYZZ t_a;    // alpha1 + r * delta1
  G1XYZZ t_b1;   // beta1 + s * delta1
  G1XYZZ t_c;    // (-r*s) * delta1
  G2XYZZ t_b2;   // beta2 + s * delta2
  G1XYZZ t_sa;   // s * alpha1     (masked partials only, see Prover::set_mask)
  G1XYZZ t_rb;   // r * beta1
  uint32_t r[8], s[8];
};

struct alignas(16) SpecPointsDev {    // SpecPoints (zkey_types.nim:24-31) needed by the prover
  G1Affine alpha1, beta1, delta1;
  G2Affine beta2, delta2;
};

// {v_lo, v_hi, h_lo, h_hi} of rank k of G (policy: see prover.cu)
void shard_ranges(size_t nvars, size_t n, int k, int G, size_t out[4]);

// Everything that depends only on the zkey: built once, read-only afterwards, shared by all proofs in flight.
struct Resident {
  Resident(const g16_zkey_view& zk, int shard_index, int shard_count);
  int shard_index, shard_count;
  uint32_t nvars, npubs, log_n, flavour;
  size_t n;
  size_t v_lo, v_hi, h_lo, h_hi;       // this shard's ranges of the witness-indexed arrays and of H1 (msm.nim:107-111)
  // window tables 2^(c w) P_i: A1, B1, C1 (padded to witness indices), H1 in G1; B2 in G2
  DevBuf tabA1, tabB1, tabC1, tabH1, tabB2;
  MsmGeometry gw, gh;
  SparseCsr csr;
  DevBuf spec;                         // SpecPointsDev
  DevBuf dtab1, dtab2;                 // 2^j * delta1 / delta2
  DevBuf atab1, btab1;                 // 2^j * alpha1 / beta1
  size_t bytes() const;
};

// One proof in flight: streams, scratch and the sorter / accumulator workspaces.
class Prover {
 public:
  Prover(const g16_zkey_view& zk, int shard_index, int shard_count);
  explicit Prover(std::shared_ptr<Resident> resident);   // another slot over the same resident key
  ~Prover();
  std::shared_ptr<Resident> resident() const { return R; }
  // witness upload (host or device source) into the resident standard-form buffer
  void load_witness(const void* w, int form, int mem_kind);
  void run_msms(g16_stats* stats);                     // ABC, quotient, five MSMs -> results_ (asynchronous)
  void collect_stats(g16_stats* stats);                // phase times of the last run (after completion)
  void partials_to_affine(void* partials_dev);         // results_ -> g16_partials (device), synchronous
  void partials_to_affine_async(void* partials_dev);
  void partials_wait(g16_stats* stats);
  void finish_async();                                 // enqueue assembly + copy-out
  void wait(g16_proof* proof, g16_stats* stats);       // completion of finish_async()
  bool in_flight() const { return in_flight_; }
  void sum_partials(const void* gathered_dev, int count);   // gathered g16_partials -> results_
  void start_mask(const uint64_t r[4], const uint64_t s[4]);   // mask terms on their own stream
  // announce the masks BEFORE the partial sums: this shard folds s*A_k + r*B1_k into its c1 partial, so the
  // finish needs no MSM-dependent scalar multiplication (see k_shard_early)
  void set_mask(const uint64_t r[4], const uint64_t s[4]);
  bool masked_partials() const { return masked_partials_; }
  bool same_mask(const uint64_t r[4], const uint64_t s[4]) const;
  void finish(g16_proof* proof, g16_stats* stats);             // assemble (waits for start_mask)
  int shard_count() const { return R->shard_count; }
  uint32_t nvars() const { return R->nvars; }
  Fr* witness_dev() { return witness_.as<Fr>(); }
  void sync();
  void timer_start();                 // CUDA event on the context's main stream
  float timer_stop();                 // records, synchronises, returns elapsed ms since timer_start()
  size_t resident_bytes() const;

 private:
  void init_slot();
  std::shared_ptr<Resident> R;
  MsmSorter sortW_, sortH_;
  MsmAccumulator<Fp> accW_, accH_;
  MsmAccumulator<Fp2> accB2_;
  DevBuf witness_, staging_, abc_, qs_, results_, mask_, proof_, early_;
  bool mask_started_ = false, early_done_ = false, in_flight_ = false, masked_partials_ = false;
  uint64_t mask_host_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaStream_t main_ = nullptr, st_mask_ = nullptr, st_[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_[24];
  cudaEvent_t tev_[2] = {nullptr, nullptr};
  g16_proof* proof_pinned_ = nullptr;
};

}  // namespace g16
