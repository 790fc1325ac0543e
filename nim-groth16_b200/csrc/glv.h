// GLV decomposition of a BN254 scalar (host arithmetic, no CUDA types).
//
// G1 has the endomorphism phi(x, y) = (beta x, y) = lambda (x, y) with beta^3 = 1 in Fp and lambda^3 = 1 in Fr.  A scalar
// k < r is written k = k1 + k2 lambda (mod r) with |k1|, |k2| < 2^127, so that k P = k1 P + k2 phi(P) needs 127 instead of
// 254 sequential doublings.  Used by the two scalar multiplications of a proof whose base point depends on the witness
// (s ** pi_A and r ** rho, prover.nim:298-299): their doubling chain is latency-bound single-thread work on the
// critical path of a sequential proof.
//   lattice basis  v1 = (a1, b1),  v2 = (a2, b2),  a_i + b_i lambda = 0 (mod r),  a1 b2 - a2 b1 = r
//   c1 = round(b2 k / r),  c2 = round(-b1 k / r)   (fixed point: c_i = (k g_i + 2^379) >> 380)
//   k1 = k - c1 a1 - c2 a2,   k2 = -c1 b1 - c2 b2
// Constants derived from r (fields.nim:37) with the extended Euclid of Gallant-Lambert-Vanstone; checked by
// tests/test_abi.py (k1 + k2 lambda = k mod r, both below 2^127, edge scalars) and on the GPU by every proof test.
#pragma once
#include <stdint.h>

namespace g16 {

struct GlvSplit {
  uint64_t k1[2], k2[2];   // magnitudes
  uint32_t neg1, neg2;     // signs
};

namespace glv_detail {
typedef unsigned __int128 u128;
// out[0 .. na+nb) = a * b
inline void mul(const uint64_t* a, int na, const uint64_t* b, int nb, uint64_t* out) {
  for (int i = 0; i < na + nb; i++) out[i] = 0;
  for (int i = 0; i < na; i++) {
    u128 c = 0;
    for (int j = 0; j < nb; j++) {
      c += (u128)a[i] * b[j] + out[i + j];
      out[i + j] = (uint64_t)c;
      c >>= 64;
    }
    out[i + nb] = (uint64_t)c;
  }
}
// (k * g + 2^379) >> 380 for 4-limb k, g; result in 3 limbs
inline void mul_shift(const uint64_t k[4], const uint64_t g[4], uint64_t c[3]) {
  uint64_t t[8];
  mul(k, 4, g, 4, t);
  u128 carry = (u128)t[5] + ((uint64_t)1 << 59);     // 379 = 5 * 64 + 59
  t[5] = (uint64_t)carry;
  carry >>= 64;
  for (int i = 6; i < 8 && carry; i++) {
    carry += t[i];
    t[i] = (uint64_t)carry;
    carry >>= 64;
  }
  // >> 380 = drop 5 limbs, then 60 bits
  c[0] = (t[5] >> 60) | (t[6] << 4);
  c[1] = (t[6] >> 60) | (t[7] << 4);
  c[2] = t[7] >> 60;
}
// 5-limb two's complement helpers
inline void add5(uint64_t* x, const uint64_t* y) {
  u128 c = 0;
  for (int i = 0; i < 5; i++) {
    c += (u128)x[i] + y[i];
    x[i] = (uint64_t)c;
    c >>= 64;
  }
}
inline void neg5(uint64_t* x) {
  u128 c = 1;
  for (int i = 0; i < 5; i++) {
    c += (uint64_t)~x[i];
    x[i] = (uint64_t)c;
    c >>= 64;
  }
}
inline void split_sign(uint64_t* x, uint64_t out[2], uint32_t& neg) {
  neg = (uint32_t)(x[4] >> 63);
  if (neg) neg5(x);
  out[0] = x[0];
  out[1] = x[1];
}
}  // namespace glv_detail

// k: standard-form integer below r, 4 little-endian limbs.  ok = false if a magnitude does not fit 128 bits (cannot
// happen for k < r; the caller then uses the plain 254-bit path).
inline GlvSplit glv_decompose(const uint64_t k[4], bool* ok = nullptr) {
  using namespace glv_detail;
  static const uint64_t A1[1] = {0x89d3256894d213e3ull};                                   // a1 = b2
  static const uint64_t A2[2] = {0x0be4e1541221250bull, 0x6f4d8248eeb859fdull};            // a2
  static const uint64_t NB1[2] = {0x8211bbeb7d4f1128ull, 0x6f4d8248eeb859fcull};           // -b1
  static const uint64_t G1C[4] = {0x28fa7d32d2fafba6ull, 0x76eb9c714773a6efull, 0x2d91d232ec7e0b3dull, 0};   // round(2^380 b2 / r)
  static const uint64_t G2C[4] = {0x9869375169b9be00ull, 0xda5e38cfb5eaa26dull, 0xf7a7bd9d4391eb18ull,
                                  0x24ccef014a773d2cull};                                   // round(2^380 (-b1) / r)
  uint64_t c1[3], c2[3];
  mul_shift(k, G1C, c1);
  mul_shift(k, G2C, c2);
  uint64_t p[5], q[5], t[6];
  // k1 = k - c1 a1 - c2 a2
  uint64_t k1[5] = {k[0], k[1], k[2], k[3], 0};
  mul(c1, 3, A1, 1, t);
  for (int i = 0; i < 5; i++) p[i] = i < 4 ? t[i] : 0;
  mul(c2, 3, A2, 2, t);
  for (int i = 0; i < 5; i++) q[i] = t[i];
  add5(p, q);
  neg5(p);
  add5(k1, p);
  // k2 = c1 (-b1) - c2 b2
  uint64_t k2[5];
  mul(c1, 3, NB1, 2, t);
  for (int i = 0; i < 5; i++) k2[i] = t[i];
  mul(c2, 3, A1, 1, t);
  for (int i = 0; i < 5; i++) q[i] = i < 4 ? t[i] : 0;
  neg5(q);
  add5(k2, q);
  GlvSplit r;
  split_sign(k1, r.k1, r.neg1);
  split_sign(k2, r.k2, r.neg2);
  if (ok) *ok = !(k1[2] | k1[3] | k1[4] | k2[2] | k2[3] | k2[4]);
  return r;
}

}  // namespace g16
