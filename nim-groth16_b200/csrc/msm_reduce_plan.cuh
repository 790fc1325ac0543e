// Plan of the bucket reduction sum_k (k+1) B_k of one bucket set (host side, no CUDA types): which running-sum
// levels run, how many bit-sliced sums follow, and with which power of two every partial enters the total.
// Shared by MsmAccumulator::run (msm_impl.cuh) and the CPU emulation of the host tests (tests/hostemu).
//
//   level 1: thread t owns L1 consecutive buckets: S1_t = sum_j B_{t L1 + j},  R1 += sum_j (j+1) B_{t L1 + j}
//   level 2: (long arrays only) thread t owns L2 consecutive S1: S2_t = sum_j S1_{t L2 + j}, R2 += sum_j j S1_{t L2 + j}
//   bits:    U_j = sum of the S_t (S2, or S1 without level 2) whose index t has bit j set
//   total  = R1 + L1 * (R2 + L2 * sum_j 2^j U_j)
// "Small levels" are R2 (when present) followed by U_0 .. U_{nbits-1}; small level l enters the total multiplied
// by 2^shift[l] and is stored as count[l] partial sums.
#pragma once
#include <stdint.h>

namespace g16 {

constexpr uint32_t REDUCE_BITS_CHUNK = 1024;
constexpr uint32_t REDUCE_MAX_SMALL = 20;

struct ReducePlan {
  uint32_t L1, n1, L2, n2;
  uint32_t tpb1, blocks1, tpb2, blocks2;
  uint32_t nbits, nchunks, first_bit_level, nsmall, stride;
  uint8_t shift[REDUCE_MAX_SMALL];
  uint16_t count[REDUCE_MAX_SMALL];
  bool ok;
};

inline ReducePlan msm_reduce_plan(uint32_t nb, uint32_t l1_max = 16, uint32_t l2_min_n1 = 8192) {
  auto ilog2 = [](uint32_t v) { uint32_t l = 0; while ((1u << l) < v) l++; return l; };
  auto tpb_for = [](uint32_t threads) { return threads < 128u ? (threads < 32u ? 32u : threads) : 128u; };
  ReducePlan p;
  // (shorter running sums with more bit slicing were tried for small bucket sets: less latency per MSM but more
  // work, and a sharded proof is throughput-bound in aggregate: 1/8 shard 5.27 -> 5.60 ms; not kept)
  p.L1 = nb >= l1_max ? l1_max : nb;
  p.n1 = nb / p.L1;
  p.L2 = p.n1 >= l2_min_n1 ? 8u : 1u;             // level 2 only pays for itself on long arrays
  p.n2 = p.n1 / p.L2;
  p.tpb1 = tpb_for(p.n1);
  p.blocks1 = (p.n1 + p.tpb1 - 1) / p.tpb1;
  p.tpb2 = tpb_for(p.n2);
  p.blocks2 = p.L2 > 1 ? (p.n2 + p.tpb2 - 1) / p.tpb2 : 0;
  p.nbits = ilog2(p.n2);
  p.nchunks = (p.n2 + REDUCE_BITS_CHUNK - 1) / REDUCE_BITS_CHUNK;
  p.first_bit_level = p.L2 > 1 ? 1u : 0u;
  p.nsmall = p.first_bit_level + p.nbits;
  p.stride = p.blocks2 > p.nchunks ? p.blocks2 : p.nchunks;
  p.ok = p.nsmall <= REDUCE_MAX_SMALL;
  const uint32_t logL1 = ilog2(p.L1), logL2 = ilog2(p.L2);
  for (uint32_t l = 0; l < REDUCE_MAX_SMALL; l++) {
    p.shift[l] = 0;
    p.count[l] = 0;
  }
  if (p.L2 > 1) {
    p.shift[0] = (uint8_t)logL1;
    p.count[0] = (uint16_t)p.blocks2;
  }
  for (uint32_t j = 0; j < p.nbits && p.first_bit_level + j < REDUCE_MAX_SMALL; j++) {
    p.shift[p.first_bit_level + j] = (uint8_t)(logL1 + logL2 + j);
    p.count[p.first_bit_level + j] = (uint16_t)p.nchunks;
  }
  return p;
}

}  // namespace g16
