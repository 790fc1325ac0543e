// Library-internal interface of prover.cu: the resident proving context.
#pragma once
#include <cuda_runtime.h>
#include <memory>
#include <utility>
#include <vector>
#include "../../include/g16b200.h"
#include "abc.cuh"
#include "common.cuh"
#include "ec.cuh"
#include "msm.cuh"

namespace g16 {

struct alignas(16) PartialsAffine {   // == g16_partials
  G1Affine a1, b1, h1, c1;
  G2Affine b2;
  uint64_t tag[2];                    // [0]: 1 = masked record (c1 = C_k + s A_k + r B1_k); [1]: hash of (r, s) if masked
};
struct alignas(16) ProofOut {         // what travels back to the host: the proof and a status word
  g16_proof proof;
  uint32_t status;                    // 0 = ok; 1 = the gathered partial records disagree on the mask
  uint32_t pad[3];
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
  // GLV split of r ([0]) and s ([1]) (csrc/glv.h): |k1| in words 0..3, |k2| in 4..7, their signs in 8, 9; glv_ok = 0
  // sends the witness-dependent scalar multiplications down the plain 256-bit path
  uint32_t glv[2][10];
  uint32_t glv_ok, pad_[3];
};

struct alignas(16) SpecPointsDev {    // SpecPoints (zkey_types.nim:24-31) needed by the prover
  G1Affine alpha1, beta1, delta1;
  G2Affine beta2, delta2;
};

// What rank k of G owns: a contiguous point range of each of the five MSMs (msm.nim:107-111 chunking, applied per
// MSM instead of to all of them alike).  A1 / B1 / B2 / C1 ranges are witness indices (C1[j - npubs - 1]
// multiplies witness[j], prover.nim:262-264); the H range indexes the domain.  Policy and cost model: prover.cu.
struct ShardPlan {
  size_t a1_lo, a1_hi, b1_lo, b1_hi, c1_lo, c1_hi, b2_lo, b2_hi, h_lo, h_hi;
};
void shard_plan(size_t nvars, size_t npubs, size_t n, int k, int G, ShardPlan& out);

// The witness-indexed pieces of a shard grouped by range: pieces with the same [lo, hi) share one digit/sort pass
// and (in G1) the accumulate and reduce launches.
struct WitnessGroup {
  size_t lo = 0, hi = 0;
  MsmGeometry geom;
  int nsets1 = 0;                      // G1 point sets of this group (A1, B1, C1 in that order when present)
  int which1[3] = {0, 0, 0};           // 0 = A1, 1 = B1, 2 = C1
  DevBuf tab1[3];
  bool has_b2 = false;
  DevBuf tabB2;
};

// Everything that depends only on the zkey: built once, read-only afterwards, shared by all proofs in flight.
struct Resident {
  Resident(const g16_zkey_view& zk, int shard_index, int shard_count);
  int shard_index, shard_count;
  uint32_t nvars, npubs, log_n, flavour;
  size_t n;
  ShardPlan plan;                      // this shard's ranges
  bool precomp = true;                 // window tables (resident key) or plain points (one-shot context)
  bool owns_ab = false;                // has A1 or B1 points (its masked partials need k_shard_early)
  // per group: window tables 2^(c w) P_i (or the plain points) of A1, B1, C1 (padded to witness indices) and B2
  std::vector<std::unique_ptr<WitnessGroup>> groups;
  DevBuf tabH1;
  MsmGeometry gh;
  std::vector<std::pair<size_t, size_t>> witness_needs;   // merged witness index intervals this shard reads
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
  // The same in three steps, for a host witness that several devices of one process share (multi.cu): every device
  // uploads ONE slice over its own PCIe link, the devices then exchange slices over NVLink.
  //   witness_begin     start of the phase (timing event), staging buffer for Montgomery input
  //   witness_raw       the buffer raw values land in (staging for Montgomery input, the witness buffer otherwise)
  //   witness_upload    H2D of elements [lo, hi) into witness_raw on the main stream; witness_uploaded() fires after it
  //   witness_finish    after the missing parts of the needed intervals have been copied in on the main stream:
  //                     conversion / reduction of the needed intervals
  void witness_begin(int form);
  Fr* witness_raw(int form) { return form == G16_FORM_MONT ? staging_.as<Fr>() : witness_.as<Fr>(); }
  void witness_upload(const void* w_host, int form, size_t lo, size_t hi);
  cudaEvent_t witness_uploaded() const { return ev_[17]; }
  void witness_finish(int form, size_t h2d_bytes);
  const std::vector<std::pair<size_t, size_t>>& witness_needs() const { return R->witness_needs; }
  void run_msms(g16_stats* stats);                     // ABC, quotient, five MSMs -> results_ (asynchronous)
  void run_msms_body(bool capturing);                  // the launches themselves (replayed from a CUDA graph when enabled)
  void collect_stats(g16_stats* stats);                // phase times of the last run (after completion)
  void partials_to_affine(void* partials_dev);         // results_ -> g16_partials (device), synchronous
  // ordering against a caller-owned stream (the collective between partials and finish): direction 0 makes
  // `ext` wait for this context's partial record, 1 makes this context wait for what is enqueued on `ext`
  void order_stream(cudaStream_t ext, int direction);
  cudaStream_t main_stream() const { return main_; }
  cudaEvent_t partials_event() const { return ev_[10]; }
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
  size_t last_witness_bytes() const { return h2d_bytes_; }
  const ShardPlan& plan() const { return R->plan; }
  bool table_layout() const { return R->precomp; }

 private:
  void init_slot();
  std::shared_ptr<Resident> R;
  struct GroupWork {                   // per-proof workspaces of one WitnessGroup
    MsmSorter sort;
    MsmAccumulator<Fp> acc1;
    MsmAccumulator<Fp2> acc2;
  };
  std::vector<std::unique_ptr<GroupWork>> gw_;
  MsmSorter sortH_;
  MsmAccumulator<Fp> accH_;
  DevBuf witness_, staging_, abc_, qs_, results_, mask_, proof_, early_;
  bool mask_started_ = false, early_done_ = false, in_flight_ = false, masked_partials_ = false, mask_used_ = false;
  uint64_t mask_hash_ = 0;
  uint64_t mask_host_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaStream_t main_ = nullptr, st_mask_ = nullptr, st_[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_[24];
  cudaEvent_t gev_[4] = {nullptr, nullptr, nullptr, nullptr};   // sort of group i done (the G2 stream waits on it)
  cudaEvent_t gdone_[4] = {nullptr, nullptr, nullptr, nullptr}; // G1 work of group i done
  cudaStream_t st_g_[4] = {nullptr, nullptr, nullptr, nullptr}; // streams of the groups after the first
  // high-priority streams for the latency-bound tails: [0] H chain, [1] B2, [2..5] the G1 work of group 0..3
  cudaStream_t tail_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t tdone_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t h2d_bytes_ = 0;                // witness bytes copied by the last load_witness()
  // CUDA graph of run_msms_body per mode (0: no mask kernel, 1: masked partials, 2: early assembly); opt-in G16_GRAPH=1
  struct GraphSlot { cudaGraphExec_t exec = nullptr; uint64_t launches = 0; bool failed = false; };
  GraphSlot graphs_[3];
  int use_graph_ = 0;
  uint64_t runs_ = 0;
  cudaEvent_t tev_[2] = {nullptr, nullptr};
  ProofOut* proof_pinned_ = nullptr;
};

}  // namespace g16
