// In-library multi-GPU prover (see multi.cuh).  Replaces the thread-pool parallelism inside
// generateProofWithMask (groth16/prover.nim:215-304 with groth16/bn128/msm.nim:96-124) by devices of one process.
#include "multi.cuh"
#include <stdlib.h>
#include <string.h>
#include <exception>
#include <functional>
#include <thread>

namespace g16 {

MultiProver::MultiProver(const g16_zkey_view& zk, const std::vector<int>& devices) : dev_(devices) {
  const int G = (int)dev_.size();
  G16_REQUIRE(G >= 1, "multi-device context needs at least one device");
  int have = 0;
  G16_CUDA(cudaGetDeviceCount(&have));
  for (int d : dev_) G16_REQUIRE(d >= 0 && d < have, "multi-device context: no such device");
  // NVLink peer access between every pair of devices: the partial records go to the first device, the witness
  // slices go from every device to every device that reads them.  DevBuf memory comes from the stream-ordered pools,
  // which need their own access grant (without it the runtime stages peer copies through the host).
  for (int a = 0; a < G; a++)
    for (int b2 = 0; b2 < G; b2++) {
      const int owner = dev_[a], peer = dev_[b2];
      if (owner == peer) continue;                  // the same device listed twice: a test configuration
      int can = 0;
      cudaDeviceCanAccessPeer(&can, peer, owner);
      if (!can) continue;                           // peer copies are then staged through the host by the runtime
      {
        DeviceGuard g(peer);                        // `peer` maps the memory of `owner`
        cudaError_t e = cudaDeviceEnablePeerAccess(owner, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) G16_CUDA(e);
        cudaGetLastError();
      }
      cudaMemPool_t pool = nullptr;
      if (cudaDeviceGetDefaultMemPool(&pool, owner) != cudaSuccess) continue;
      cudaMemAccessDesc d;
      memset(&d, 0, sizeof(d));
      d.location.type = cudaMemLocationTypeDevice;
      d.location.id = peer;
      d.flags = cudaMemAccessFlagsProtReadWrite;
      if (cudaMemPoolSetAccess(pool, &d, 1) != cudaSuccess) cudaGetLastError();
    }
  for (int k = 0; k < G; k++) {
    DeviceGuard g(dev_[k]);
    shard_.emplace_back(new Prover(zk, k, G));
  }
  init_slot();
}

MultiProver::MultiProver(const MultiProver& base) : dev_(base.dev_) {
  for (size_t k = 0; k < dev_.size(); k++) {
    DeviceGuard g(dev_[k]);
    shard_.emplace_back(new Prover(base.shard_[k]->resident()));
  }
  init_slot();
}

void MultiProver::init_slot() {
  const int G = (int)dev_.size();
  if (const char* e = getenv("G16_WITNESS_SCATTER")) scatter_witness_ = e[0] != '0';
  for (int k = 0; k < G; k++) {
    DeviceGuard g(dev_[k]);
    local_.emplace_back(new DevBuf());
    local_.back()->ensure(sizeof(PartialsAffine));
    cudaEvent_t e = nullptr;
    G16_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    sent_.push_back(e);
  }
  DeviceGuard g(dev_[0]);
  gathered_.ensure((size_t)G * sizeof(PartialsAffine));
}

MultiProver::~MultiProver() {
  for (size_t k = 0; k < shard_.size(); k++) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(dev_[k]);
    shard_[k].reset();                    // ~Prover synchronises its device
    local_[k].reset();
    if (sent_[k]) cudaEventDestroy(sent_[k]);
    cudaSetDevice(prev);
  }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(dev_[0]);
  gathered_.release();
  cudaSetDevice(prev);
}

void MultiProver::submit(const void* witness, int form, int mem_kind, const uint64_t r[4], const uint64_t s[4]) {
  const int G = (int)dev_.size();
  G16_REQUIRE(!in_flight(), "a proof is already in flight on this context");
  PartialsAffine* all = gathered_.as<PartialsAffine>();
  // A host witness travels ONCE over PCIe: device k uploads the k-th of G slices over its own link, then every device
  // pulls the parts of the intervals it reads from the devices that hold them (NVLink peer copies ordered by events)
  // -- instead of up to G copies of the whole witness competing for the host's PCIe lanes.
  const bool scatter = mem_kind == G16_MEM_HOST && G > 1 && scatter_witness_;
  const size_t nv = shard_[0]->nvars();
  auto slice = [&](int k, size_t& lo, size_t& hi) {
    lo = nv * (size_t)k / (size_t)G;
    hi = k == G - 1 ? nv : nv * (size_t)(k + 1) / (size_t)G;
  };
  auto run_on_all = [&](const std::function<void(int)>& fn) {
    // one host thread per device (the calling thread takes the first) enqueues that device's part -- about 70
    // launches per proof each: eight devices are fed in the time of one
    std::vector<std::thread> workers;
    std::vector<std::exception_ptr> errors((size_t)G);
    for (int k = G - 1; k >= 1; k--)
      workers.emplace_back([&, k] {
        try {
          fn(k);
        } catch (...) {
          errors[(size_t)k] = std::current_exception();
        }
      });
    try {
      fn(0);
    } catch (...) {
      errors[0] = std::current_exception();
    }
    for (auto& w : workers) w.join();
    for (auto& e : errors)
      if (e) std::rethrow_exception(e);
  };
  if (scatter) {
    G16_REQUIRE(witness != nullptr, "witness is null");
    // phase 1: every device uploads its slice (a pageable source is staged by the calling thread, so the slices are
    // staged in parallel as well); the peer copies of phase 2 need these events to have been recorded
    run_on_all([&](int k) {
      DeviceGuard g(dev_[k]);
      size_t lo, hi;
      slice(k, lo, hi);
      shard_[k]->witness_begin(form);
      shard_[k]->witness_upload(witness, form, lo, hi);
    });
  }
  auto enqueue = [&](int k) {
    DeviceGuard g(dev_[k]);
    Prover& p = *shard_[k];
    p.set_mask(r, s);                     // every shard folds s*A_k + r*B1_k into its record next to its MSMs
    if (scatter) {
      for (const auto& iv : p.witness_needs())
        for (int j = 0; j < G; j++) {
          if (j == k) continue;
          size_t lo, hi;
          slice(j, lo, hi);
          if (lo < iv.first) lo = iv.first;
          if (hi > iv.second) hi = iv.second;
          if (hi <= lo) continue;
          G16_CUDA(cudaStreamWaitEvent(p.main_stream(), shard_[j]->witness_uploaded(), 0));
          G16_CUDA(cudaMemcpyPeerAsync(p.witness_raw(form) + lo, dev_[k], shard_[j]->witness_raw(form) + lo, dev_[j],
                                       (hi - lo) * sizeof(Fr), p.main_stream()));
        }
      size_t lo, hi;
      slice(k, lo, hi);
      p.witness_finish(form, (hi - lo) * sizeof(Fr));
    } else {
      p.load_witness(witness, form, mem_kind);
    }
    p.run_msms(nullptr);
    p.partials_to_affine_async(local_[k]->p);
    G16_CUDA(cudaMemcpyPeerAsync(all + k, dev_[0], local_[k]->p, dev_[k], sizeof(PartialsAffine), p.main_stream()));
    G16_CUDA(cudaEventRecord(sent_[k], p.main_stream()));
  };
  run_on_all(enqueue);
  DeviceGuard g(dev_[0]);
  Prover& head = *shard_[0];
  for (int k = 1; k < G; k++) G16_CUDA(cudaStreamWaitEvent(head.main_stream(), sent_[k], 0));
  head.sum_partials(all, G);
  head.finish_async();
}

void MultiProver::wait(g16_proof* proof, g16_stats* stats) {
  {
    DeviceGuard g(dev_[0]);
    shard_[0]->wait(proof, stats);
  }
  if (!stats) return;
  // the head waited for every record, so the other shards' events have completed: maximum per phase
  for (size_t k = 1; k < shard_.size(); k++) {
    g16_stats st;
    memset(&st, 0, sizeof(st));
    DeviceGuard g(dev_[k]);
    shard_[k]->collect_stats(&st);
    float* a = &stats->ms_h2d;
    const float* b = &st.ms_h2d;
    for (int i = 0; i < 8; i++)           // ms_h2d .. reserved_ms
      if (b[i] > a[i]) a[i] = b[i];
  }
}

size_t MultiProver::last_witness_bytes() const {
  size_t t = 0;
  for (auto& p : shard_) t += p->last_witness_bytes();
  return t;
}
size_t MultiProver::resident_bytes() const {
  size_t t = 0;
  for (auto& p : shard_) t += p->resident_bytes();
  return t;
}
void MultiProver::timer_start() {
  DeviceGuard g(dev_[0]);
  shard_[0]->timer_start();
}
float MultiProver::timer_stop() {
  DeviceGuard g(dev_[0]);
  return shard_[0]->timer_stop();
}

}  // namespace g16
