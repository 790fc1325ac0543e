// Signed-digit window decomposition of a 254-bit standard-form scalar (GPU Pippenger front end).
// digit_w in [-(2^(c-1) - 1), 2^(c-1)];  sum_w digit_w * 2^(c*w) == k  provided nwin*c >= 255.
#pragma once
#include "field.cuh"

namespace g16 {

// bits [pos, pos+c) of the 256-bit little-endian integer k (zero beyond bit 255), c <= 24
G16_HD uint32_t msm_extract_bits(const uint32_t k[8], int pos, int c) {
  if (pos >= 256) return 0;
  int limb = pos >> 5, sh = pos & 31;
  uint64_t w = k[limb];
  if (limb + 1 < 8) w |= (uint64_t)k[limb + 1] << 32;
  return (uint32_t)(w >> sh) & ((1u << c) - 1u);
}

G16_HD int msm_signed_digit(const uint32_t k[8], int c, int w, int nwin, int& carry) {
  (void)nwin;
  uint32_t v = msm_extract_bits(k, w * c, c) + (uint32_t)carry;
  if (v > (1u << (c - 1))) {
    carry = 1;
    return (int)v - (1 << c);
  }
  carry = 0;
  return (int)v;
}

// number of windows needed so that the top window never carries out (scalars < 2^254)
G16_HD int msm_num_windows(int c) { return (255 + c - 1) / c; }

}  // namespace g16
