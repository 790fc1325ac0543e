// In-library multi-GPU prover: one process, N devices, behind the unchanged generateProofWithMask-shaped call.
//
// The reference's prover is one proc whose parallelism is internal (groth16/bn128/msm.nim:96-124: chunks handed
// to a thread pool, partial sums added by the caller, :117-119).  MultiProver is that shape across devices: shard
// k of the ShardPlan lives on device k, every proof sends each device the witness intervals it reads, the 400-byte
// partial records return to the first device by peer copy, ordered with events, and the first device assembles.
// No NCCL, no host synchronisation between the partial sums and the finish.
#pragma once
#include <memory>
#include <vector>
#include "prover.cuh"

namespace g16 {

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (dev != prev) G16_CUDA(cudaSetDevice(dev));
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

class MultiProver {
 public:
  MultiProver(const g16_zkey_view& zk, const std::vector<int>& devices);
  explicit MultiProver(const MultiProver& base);       // another proof slot over the same resident shards
  ~MultiProver();
  void submit(const void* witness, int form, int mem_kind, const uint64_t r[4], const uint64_t s[4]);
  void wait(g16_proof* proof, g16_stats* stats);
  bool in_flight() const { return shard_[0]->in_flight(); }
  int count() const { return (int)shard_.size(); }
  int device(int k) const { return dev_[k]; }
  Prover& shard(int k) { return *shard_[k]; }
  size_t last_witness_bytes() const;
  size_t resident_bytes() const;
  void timer_start();
  float timer_stop();

 private:
  void init_slot();
  std::vector<int> dev_;
  std::vector<std::unique_ptr<Prover>> shard_;
  std::vector<std::unique_ptr<DevBuf>> local_;          // device k: its own partial record
  std::vector<cudaEvent_t> sent_;                       // record k has arrived in gathered_
  DevBuf gathered_;                                     // first device: count() records
  bool scatter_witness_ = true;                         // host witness: one PCIe upload in G slices + NVLink exchange
};

}  // namespace g16
