// Pass planning and tile index arithmetic for the multi-pass shared-memory NTT.
// Shared between the CUDA kernels (ntt.cu) and the host emulation used by the CPU tests, so the
// index math is exercised without a GPU.
//
// A transform over n = 2^log_n points is split into passes; a pass owns the index bits
// [t_lo, t_lo + k) and runs those k radix-2 stages on tiles held in shared memory.  A tile is
// 2^k "rows" (the k owned bits) by C = 2^logC adjacent columns (low bits), so global accesses are
// runs of C*32 contiguous bytes.  Decimation-in-frequency walks bits from high to low (natural
// in, bit-reversed out); decimation-in-time walks them low to high (bit-reversed in, natural out).
#pragma once
#include <stdint.h>
#include "field.cuh"

namespace g16 {

#ifndef G16_NTT_TILE_LOG
#define G16_NTT_TILE_LOG 11
#endif
constexpr int NTT_TILE_LOG = G16_NTT_TILE_LOG;   // 2048 elements * 32 B = 64 KiB of shared memory per CTA
constexpr int NTT_MIN_LOGC = 2;    // strided passes read runs of >= 4 elements = 128 B
constexpr int NTT_MAX_PASSES = 8;

struct NttPass {
  int t_lo;   // lowest index bit owned by the pass
  int k;      // number of stages (bits) in the pass
  int logC;   // log2 of adjacent columns per tile
};

struct NttPlan {
  int log_n;
  int npass;
  NttPass pass[NTT_MAX_PASSES];   // ordered from low bits to high bits
};

// Every pass of a transform of 2^11 points or more works on full tiles (k + logC == NTT_TILE_LOG): the contiguous
// pass owns the low 11 bits, the remaining bits are split evenly over strided passes of at most 9 bits.
inline NttPlan ntt_make_plan(int log_n) {
  NttPlan pl;
  pl.log_n = log_n;
  const int kc = NTT_TILE_LOG;                  // contiguous pass: up to 11 bits
  const int ks = NTT_TILE_LOG - NTT_MIN_LOGC;   // strided pass: up to 9 bits
  if (log_n <= kc) {
    pl.npass = 1;
    pl.pass[0] = NttPass{0, log_n, 0};
    return pl;
  }
  const int rest = log_n - kc;
  const int S = (rest + ks - 1) / ks;           // strided passes
  pl.npass = 1 + S;
  pl.pass[0] = NttPass{0, kc, 0};
  int rem = rest, t = kc;
  for (int i = 1; i <= S; i++) {
    int k = (rem + (S - i + 1) - 1) / (S - i + 1);
    pl.pass[i] = NttPass{t, k, NTT_TILE_LOG - k};
    t += k;
    rem -= k;
  }
  return pl;
}

// element p of tile `tile` -> global index
G16_HD uint32_t ntt_global_index(uint32_t tile, uint32_t p, int t_lo, int k, int logC) {
  uint32_t groups_log = (uint32_t)(t_lo - logC);            // lo-groups per hi value = 2^(t_lo-logC)
  uint32_t lo0 = (tile & ((1u << groups_log) - 1u)) << logC;
  uint32_t hi = tile >> groups_log;
  uint32_t mid = p >> logC;
  uint32_t lo = lo0 | (p & ((1u << logC) - 1u));
  return (hi << (t_lo + k)) | (mid << t_lo) | lo;
}

// butterfly q of stage s (owned bit s, global bit t_lo + s): tile positions and twiddle exponent
G16_HD void ntt_butterfly_index(uint32_t tile, uint32_t q, int s, int t_lo, int logC, int log_n,
                                uint32_t& p_u, uint32_t& p_v, uint32_t& tw_index) {
  uint32_t groups_log = (uint32_t)(t_lo - logC);
  uint32_t lo0 = (tile & ((1u << groups_log) - 1u)) << logC;
  uint32_t lo = lo0 | (q & ((1u << logC) - 1u));
  uint32_t m = q >> logC;                                   // k-1 bits: the row with bit s removed
  uint32_t low_s = m & ((1u << s) - 1u);
  uint32_t mid_u = ((m >> s) << (s + 1)) | low_s;
  p_u = (mid_u << logC) | (q & ((1u << logC) - 1u));
  p_v = p_u + (1u << (logC + s));
  uint32_t t = (uint32_t)(t_lo + s);
  uint32_t g_mod = (low_s << t_lo) | lo;                    // global index mod 2^t
  tw_index = g_mod << (log_n - 1 - (int)t);                 // exponent of omega, < n/2
}

G16_HD uint32_t ntt_bitrev(uint32_t x, int log_n) {
  uint32_t r = 0;
  for (int i = 0; i < log_n; i++) r |= ((x >> i) & 1u) << (log_n - 1 - i);
  return r;
}

}  // namespace g16
