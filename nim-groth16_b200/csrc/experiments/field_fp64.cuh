// BN254 base-field multiplication on the FP64 pipe (device only).  EXPERIMENT, not on the product path:
// measured at 91 % of the IMAD multiplier alone and +2 % co-scheduled (profiles/r1_v6_fp64_pipe_probe.json).
//
// The IMAD path of field.cuh saturates the fma pipe (IMAD.WIDE is half rate) while the FP64 pipe and most of
// the ALU pipe idle.  This header provides a second, independent Montgomery multiplier built from DFMA:
// an element is five 52-bit limbs held as doubles; a limb product a*b < 2^104 is split exactly into its
// high and low 52-bit halves with two round-toward-zero FMAs
//     hi = fma_rz(a, b, 2^104)                 -> bits = 0x467<<52 | floor(a*b / 2^52)
//     lo = fma_rz(a, b, (2^104 + 2^52) - hi)   -> bits = 0x433<<52 | (a*b mod 2^52)
// and the raw bit patterns are accumulated into 64-bit integer columns whose initial values pre-subtract the
// exponent fields.  Montgomery radix is R' = 2^260 (five limbs), i.e. dfmul(a, b) = a*b*2^-260 mod p: against
// the 2^256 Montgomery form of the rest of the library every product carries an extra factor k = 2^-4, which
// the XYZZ formulas could absorb as a weighted projective rescaling (a - z = 1, b - zzz = 1 in exponents of k).
// Reduction is lazy: operands may be any values < 2^256 with normalized limbs, results are < 2p.
// The carry-exact integer model of this file is tools/emu_dfmul.py.
#pragma once
#include "field.cuh"

namespace g16 {

struct Fd {
  double v[5];
};

namespace fd {
constexpr unsigned long long M52 = (1ull << 52) - 1ull;
constexpr unsigned long long LOX = 0x433ull << 52;   // exponent field of 2^52
constexpr unsigned long long HIX = 0x467ull << 52;   // exponent field of 2^104
constexpr unsigned long long STEP = 2ull * LOX + 2ull * HIX;
// p in 52-bit limbs, -p^-1 mod 2^52
constexpr unsigned long long PL0 = 0x8c16d87cfd47ull, PL1 = 0x916871ca8d3c2ull, PL2 = 0x181585d97816aull,
                             PL3 = 0xa029b85045b68ull, PL4 = 0x30644e72e131ull, PINV = 0x20782e4866389ull;
__device__ __forceinline__ double plimb(int j) {
  return j == 0 ? (double)PL0 : j == 1 ? (double)PL1 : j == 2 ? (double)PL2 : j == 3 ? (double)PL3 : (double)PL4;
}
__device__ __forceinline__ unsigned long long plimb_u(int j) {
  return j == 0 ? PL0 : j == 1 ? PL1 : j == 2 ? PL2 : j == 3 ? PL3 : PL4;
}
__device__ __forceinline__ void split(double a, double b, unsigned long long& hi, unsigned long long& lo) {
  const double c1 = 0x1p104, c2 = 0x1p104 + 0x1p52;
  double h = __fma_rz(a, b, c1);
  double l = __fma_rz(a, b, c2 - h);
  hi = (unsigned long long)__double_as_longlong(h);
  lo = (unsigned long long)__double_as_longlong(l);
}
// integer < 2^52 -> double
__device__ __forceinline__ double to_double(unsigned long long x) {
  return __hiloint2double((int)((uint32_t)(x >> 32) | 0x43300000u), (int)(uint32_t)x) - 0x1p52;
}
// exact integer-valued double in [0, 2^52) -> integer
__device__ __forceinline__ unsigned long long to_u64(double x) {
  return (unsigned long long)__double_as_longlong(x + 0x1p52) & M52;
}
}  // namespace fd

// 8 x u32 value (< 2^256) -> five 52-bit limbs
__device__ __forceinline__ void fd_limbs_from_words(const uint32_t* w, unsigned long long* L) {
  unsigned long long W0 = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
  unsigned long long W1 = (unsigned long long)w[2] | ((unsigned long long)w[3] << 32);
  unsigned long long W2 = (unsigned long long)w[4] | ((unsigned long long)w[5] << 32);
  unsigned long long W3 = (unsigned long long)w[6] | ((unsigned long long)w[7] << 32);
  L[0] = W0 & fd::M52;
  L[1] = ((W0 >> 52) | (W1 << 12)) & fd::M52;
  L[2] = ((W1 >> 40) | (W2 << 24)) & fd::M52;
  L[3] = ((W2 >> 28) | (W3 << 36)) & fd::M52;
  L[4] = W3 >> 16;
}
__device__ __forceinline__ Fd fd_from_fp(const Fp& a) {
  unsigned long long L[5];
  fd_limbs_from_words(a.v, L);
  Fd r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.v[i] = fd::to_double(L[i]);
  return r;
}
// normalized limbs, value < 2^256 -> canonical residue (< p) in 8 x u32
__device__ __forceinline__ Fp fd_to_fp(const Fd& a) {
  unsigned long long L[5];
#pragma unroll
  for (int i = 0; i < 5; i++) L[i] = fd::to_u64(a.v[i]);
  unsigned long long W0 = L[0] | (L[1] << 52);
  unsigned long long W1 = (L[1] >> 12) | (L[2] << 40);
  unsigned long long W2 = (L[2] >> 24) | (L[3] << 28);
  unsigned long long W3 = (L[3] >> 36) | (L[4] << 16);
  Fp t;
  t.v[0] = (uint32_t)W0; t.v[1] = (uint32_t)(W0 >> 32);
  t.v[2] = (uint32_t)W1; t.v[3] = (uint32_t)(W1 >> 32);
  t.v[4] = (uint32_t)W2; t.v[5] = (uint32_t)(W2 >> 32);
  t.v[6] = (uint32_t)W3; t.v[7] = (uint32_t)(W3 >> 32);
#pragma unroll 1
  for (int it = 0; it < 6; it++) {   // 2^256 < 6p
    Fp u;
    long long bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      bw += (long long)t.v[i] - (long long)FpParams::mod(i);
      u.v[i] = (uint32_t)bw;
      bw >>= 32;
    }
    if (bw) break;
    t = u;
  }
  return t;
}

// a*b*2^-260 mod p; limbs of a and b normalized (< 2^52), values < 2^256; result < 2p, limbs normalized
__device__ __forceinline__ Fd dfmul(const Fd& a, const Fd& b) {
  using namespace fd;
  const double c1 = 0x1p104, c2 = 0x1p104 + 0x1p52;
  unsigned long long t[6];
#pragma unroll
  for (int k = 0; k < 5; k++) t[k] = 0ull - ((unsigned long long)k * STEP + 2ull * LOX);
  t[5] = 0ull - (2ull * HIX + 4ull * STEP);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    unsigned long long h[5], l[5], hq[5], lq[5];
#pragma unroll
    for (int j = 0; j < 5; j++) split(a.v[i], b.v[j], h[j], l[j]);
    t[0] += l[0];
    // q = (t0 mod 2^52) * (-p^-1) mod 2^52 : the exponent offsets only touch bits >= 52
    double t0d = __hiloint2double((int)(((uint32_t)(t[0] >> 32) & 0xFFFFFu) | 0x43300000u), (int)(uint32_t)t[0]) - 0x1p52;
    double qh = __fma_rz(t0d, (double)PINV, c1);
    double q = __fma_rz(t0d, (double)PINV, c2 - qh) - 0x1p52;
#pragma unroll
    for (int j = 0; j < 5; j++) split(q, plimb(j), hq[j], lq[j]);
    t[0] += lq[0];
#pragma unroll
    for (int j = 1; j < 5; j++) t[j] += (l[j] + lq[j]) + (h[j - 1] + hq[j - 1]);
    t[5] += h[4] + hq[4];
    unsigned long long carry = t[0] >> 52;
    t[0] = t[1] + carry;
    t[1] = t[2];
    t[2] = t[3];
    t[3] = t[4];
    t[4] = t[5];
    t[5] = (i < 4) ? 0ull - (2ull * HIX + (unsigned long long)(3 - i) * STEP) : 0ull;
  }
  Fd r;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    t[k + 1] += t[k] >> 52;
    r.v[k] = __hiloint2double((int)(((uint32_t)(t[k] >> 32) & 0xFFFFFu) | 0x43300000u), (int)(uint32_t)t[k]) - 0x1p52;
  }
  r.v[4] = to_double(t[4]);
  return r;
}

}  // namespace g16
