// 254-bit Montgomery arithmetic for BN254 Fp / Fr (R = 2^256), 8 x 32-bit limbs.
//
// Replaces the constantine field layer the reference reaches through
// groth16/bn128/fields.nim:23-28,110-133 (types + operator sugar) with sm_100a device code.
// In-memory layout is the reference's: 32 little-endian bytes holding the Montgomery residue
// (io.nim:87-92,103-114).
//
// Device path: even/odd accumulator chains on mad.lo.cc/madc.hi.cc (ptxas fuses each lo/hi
// pair into one IMAD.WIDE.U32[.X]); see tools/emu_montmul2.py for the carry-exact model.
// Host path (G16_HOST_EMU or plain g++): portable u64 code with identical results, used only
// by tests/hostemu to check the formulas layered on top.  The shipped library has no host
// compute path.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define G16_HD __host__ __device__ __forceinline__
#define G16_D __device__ __forceinline__
#else
#define G16_HD inline
#define G16_D inline
#endif

namespace g16 {

// ---------------------------------------------------------------------------------------
// moduli and derived constants (fields.nim:36-37)
// ---------------------------------------------------------------------------------------
struct FpParams {
  static G16_HD constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
  static G16_HD constexpr uint32_t one(int i) {  // 2^256 mod p  (io.nim:87)
    constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  static G16_HD constexpr uint32_t r2(int i) {  // 2^512 mod p
    constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return m[i];
  }
};

struct FrParams {
  static G16_HD constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
  static G16_HD constexpr uint32_t one(int i) {  // 2^256 mod r  (io.nim:91)
    constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  static G16_HD constexpr uint32_t r2(int i) {  // 2^512 mod r
    constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
  }
};

// ---------------------------------------------------------------------------------------
// field element
// ---------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fe {
  uint32_t v[8];
  typedef P Params;

  static G16_HD Fe zero() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  static G16_HD Fe one() {  // Montgomery form of 1
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
    return r;
  }
  static G16_HD Fe rsquared() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
    return r;
  }
  static G16_HD Fe modulus() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::mod(i);
    return r;
  }
};

typedef Fe<FpParams> Fp;
typedef Fe<FrParams> Fr;

template <class P>
G16_HD bool fis_zero(const Fe<P>& a) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) t |= a.v[i];
  return t == 0;
}

template <class P>
G16_HD bool feq(const Fe<P>& a, const Fe<P>& b) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) t |= a.v[i] ^ b.v[i];
  return t == 0;
}


#if defined(__CUDA_ARCH__)
// Every carry chain is ONE asm statement: the CC flag is invisible to the compiler, so
// separate statements could legally be reordered.
// All operands that are written are "+r" (in place): with plain "=r" outputs the compiler may
// give an output the register of an input that a later instruction of the block still reads.
// t += b (no overflow: both < 2^254, or b is a masked modulus)
G16_D void add8_ip(uint32_t* t, const uint32_t* b) {
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, %15;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7])
      : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}
// t -= b, returns 0xffffffff when the subtraction borrowed, else 0
G16_D uint32_t sub8_ip(uint32_t* t, const uint32_t* b) {
  uint32_t borrow;
  asm("sub.cc.u32 %0, %0, %9;\n\t"
      "subc.cc.u32 %1, %1, %10;\n\t"
      "subc.cc.u32 %2, %2, %11;\n\t"
      "subc.cc.u32 %3, %3, %12;\n\t"
      "subc.cc.u32 %4, %4, %13;\n\t"
      "subc.cc.u32 %5, %5, %14;\n\t"
      "subc.cc.u32 %6, %6, %15;\n\t"
      "subc.cc.u32 %7, %7, %16;\n\t"
      "subc.u32 %8, 0, 0;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]),
        "=r"(borrow)
      : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return borrow;
}
// y0 += X[1];  X := (X >> 64) + {a1,a3,a5,a7} * b + carry      (see tools/emu_montmul2.py)
G16_D void mad_row_shift(uint32_t* X, uint32_t& y0, uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7,
                         uint32_t b) {
  asm("add.cc.u32 %8, %8, %1;\n\t"
      "madc.lo.cc.u32 %0, %9, %13, %2;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %4;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %6;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, 0;\n\t"
      "madc.hi.u32 %7, %12, %13, 0;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]),
        "+r"(y0)
      : "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
}
// Y += {a0,a2,a4,a6} * b ; x7 += carry-out
G16_D void mad_row_carry(uint32_t* Y, uint32_t& x7, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6,
                         uint32_t b) {
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]), "+r"(Y[6]), "+r"(Y[7]),
        "+r"(x7)
      : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}
// X += {p1,p3,p5,p7} * m   (no carry-out by construction)
G16_D void mad_row_nc(uint32_t* X, uint32_t p1, uint32_t p3, uint32_t p5, uint32_t p7, uint32_t m) {
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
      "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7])
      : "r"(p1), "r"(p3), "r"(p5), "r"(p7), "r"(m));
}
// D += (E >> 32)  (E[0] == 0), 8 limbs, no carry-out
G16_D void merge8_ip(uint32_t* D, const uint32_t* E) {
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(D[0]), "+r"(D[1]), "+r"(D[2]), "+r"(D[3]), "+r"(D[4]), "+r"(D[5]), "+r"(D[6]), "+r"(D[7])
      : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]));
}
// ---- wide (unreduced) arithmetic for the lazily reduced Fp2 multiplication; model: tools/emu_lazy_fp2.py ----
// reduction-only shift step: y0 += X[1]; X := X >> 64 (with the carry)
G16_D void redc_shift(uint32_t* X, uint32_t& y0) {
  asm("add.cc.u32 %8, %8, %1;\n\t"
      "addc.cc.u32 %0, %2, 0;\n\t"
      "addc.cc.u32 %1, %3, 0;\n\t"
      "addc.cc.u32 %2, %4, 0;\n\t"
      "addc.cc.u32 %3, %5, 0;\n\t"
      "addc.cc.u32 %4, %6, 0;\n\t"
      "addc.cc.u32 %5, %7, 0;\n\t"
      "addc.u32 %6, 0, 0;\n\t"
      "mov.u32 %7, 0;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(y0));
}
// t[0..7] += b[0..7] with carry-in c (0 / 1) and carry-out returned (0 / 1)
G16_D uint32_t add8_c(uint32_t* t, const uint32_t* b, uint32_t c) {
  uint32_t co;
  asm("add.cc.u32 %8, %17, 0xffffffff;\n\t"
      "addc.cc.u32 %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, %10;\n\t"
      "addc.cc.u32 %2, %2, %11;\n\t"
      "addc.cc.u32 %3, %3, %12;\n\t"
      "addc.cc.u32 %4, %4, %13;\n\t"
      "addc.cc.u32 %5, %5, %14;\n\t"
      "addc.cc.u32 %6, %6, %15;\n\t"
      "addc.cc.u32 %7, %7, %16;\n\t"
      "addc.u32 %8, 0, 0;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "=r"(co)
      : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(c));
  return co;
}
// t[0..7] -= b[0..7] with borrow-in bw (0 / 1) and borrow-out returned (0 / 1)
G16_D uint32_t sub8_b(uint32_t* t, const uint32_t* b, uint32_t bw) {
  uint32_t bo;
  asm("sub.cc.u32 %8, 0, %17;\n\t"
      "subc.cc.u32 %0, %0, %9;\n\t"
      "subc.cc.u32 %1, %1, %10;\n\t"
      "subc.cc.u32 %2, %2, %11;\n\t"
      "subc.cc.u32 %3, %3, %12;\n\t"
      "subc.cc.u32 %4, %4, %13;\n\t"
      "subc.cc.u32 %5, %5, %14;\n\t"
      "subc.cc.u32 %6, %6, %15;\n\t"
      "subc.cc.u32 %7, %7, %16;\n\t"
      "subc.u32 %8, 0, 0;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "=r"(bo)
      : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(bw));
  return bo & 1u;
}
// out[0..15] = a * b for 8-limb a, b < 2^256: even-column products accumulate in E, odd-column products in O
// (O[k] sits at column k+1); every row is one carry chain of four lo/hi pairs (IMAD.WIDE after fusion)
G16_D void mul_wide(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  uint32_t E[16], O[16];
#pragma unroll
  for (int k = 0; k < 16; k++) E[k] = O[k] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if ((i & 1) == 0) {
      mad_row_carry(E + i, E[i + 8], a[0], a[2], a[4], a[6], b[i]);
      mad_row_carry(O + i, O[i + 8], a[1], a[3], a[5], a[7], b[i]);
    } else {
      mad_row_carry(O + (i - 1), O[i + 7], a[0], a[2], a[4], a[6], b[i]);
      if (i < 7) mad_row_carry(E + (i + 1), E[i + 9], a[1], a[3], a[5], a[7], b[i]);
      else mad_row_nc(E + 8, a[1], a[3], a[5], a[7], b[7]);          // the product is below 2^512: no carry-out
    }
  }
  // out = E + (O << 32)
  out[0] = E[0];
  uint32_t c = add8_c(E + 1, O, 0);          // limbs 1..8
  uint32_t hi[8];
#pragma unroll
  for (int k = 0; k < 7; k++) hi[k] = O[8 + k];
  hi[7] = 0;                                  // O[15] == 0: O * 2^32 < 2^512
  uint32_t e2[8];
#pragma unroll
  for (int k = 0; k < 7; k++) e2[k] = E[9 + k];
  e2[7] = 0;
  add8_c(e2, hi, c);                          // limbs 9..15 (e2[7] stays 0)
#pragma unroll
  for (int k = 1; k < 9; k++) out[k] = E[k];
#pragma unroll
  for (int k = 0; k < 7; k++) out[9 + k] = e2[k];
}
#endif

// ---------------------------------------------------------------------------------------
// add / sub (inputs and outputs fully reduced: < modulus)
// ---------------------------------------------------------------------------------------
template <class P>
G16_HD Fe<P> fadd(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> t, u;
#if defined(__CUDA_ARCH__)
  uint32_t m[8];
#pragma unroll
  for (int i = 0; i < 8; i++) m[i] = P::mod(i);
  t = a;
  add8_ip(t.v, b.v);
  u = t;
  uint32_t borrow = sub8_ip(u.v, m);
#pragma unroll
  for (int i = 0; i < 8; i++) t.v[i] = borrow ? t.v[i] : u.v[i];
  return t;
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.v[i] + b.v[i];
    t.v[i] = (uint32_t)c;
    c >>= 32;
  }
  int64_t bw = 0;
  for (int i = 0; i < 8; i++) {
    bw += (int64_t)t.v[i] - (int64_t)P::mod(i);
    u.v[i] = (uint32_t)bw;
    bw >>= 32;
  }
  return bw ? t : u;
#endif
}

template <class P>
G16_HD Fe<P> fsub(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> t;
#if defined(__CUDA_ARCH__)
  t = a;
  uint32_t borrow = sub8_ip(t.v, b.v);
  uint32_t m[8];
#pragma unroll
  for (int i = 0; i < 8; i++) m[i] = P::mod(i) & borrow;
  add8_ip(t.v, m);
  return t;
#else
  int64_t bw = 0;
  for (int i = 0; i < 8; i++) {
    bw += (int64_t)a.v[i] - (int64_t)b.v[i];
    t.v[i] = (uint32_t)bw;
    bw >>= 32;
  }
  if (bw) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
      c += (uint64_t)t.v[i] + P::mod(i);
      t.v[i] = (uint32_t)c;
      c >>= 32;
    }
  }
  return t;
#endif
}

template <class P>
G16_HD Fe<P> fneg(const Fe<P>& a) {
  return fis_zero(a) ? a : fsub(Fe<P>::modulus(), a);
}

template <class P>
G16_HD Fe<P> fdbl(const Fe<P>& a) {
  return fadd(a, a);
}

// a/2 mod p (ntt.nim:121 div2)
template <class P>
G16_HD Fe<P> fhalve(const Fe<P>& a) {
  Fe<P> t = a;
  uint64_t c = 0;
  uint32_t odd = 0u - (a.v[0] & 1u);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)t.v[i] + (P::mod(i) & odd);
    t.v[i] = (uint32_t)c;
    c >>= 32;
  }
#pragma unroll
  for (int i = 0; i < 7; i++) t.v[i] = (t.v[i] >> 1) | (t.v[i + 1] << 31);
  t.v[7] = (t.v[7] >> 1) | ((uint32_t)c << 31);
  return t;
}

// ---------------------------------------------------------------------------------------
// Montgomery multiplication: returns a*b/2^256 mod p, fully reduced.
// ---------------------------------------------------------------------------------------
// LAZY = true skips the final conditional subtraction: for operands < 2p the result is < 2p as well
// ((2p)^2 / 2^256 + p < 1.76 p because p < 2^254), congruent to a*b/2^256.
template <bool LAZY, class P>
G16_HD Fe<P> fmul_core(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
#if defined(__CUDA_ARCH__)
  // t[0] / t[1] alternate between the "even" (column k) and "odd" (column k+1) roles.
  uint32_t t[2][8];
#pragma unroll
  for (int k = 0; k < 8; k++) t[0][k] = t[1][k] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t* X = t[i & 1];        // odd-aligned array of this iteration (shifted in place)
    uint32_t* Y = t[(i + 1) & 1];  // even-aligned array of this iteration
    const uint32_t bi = b.v[i];
    mad_row_shift(X, Y[0], a.v[1], a.v[3], a.v[5], a.v[7], bi);
    mad_row_carry(Y, X[7], a.v[0], a.v[2], a.v[4], a.v[6], bi);
    const uint32_t m = Y[0] * P::INV;
    mad_row_nc(X, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
    mad_row_carry(Y, X[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
  }
  // iteration 7 had X = t[1] (odd role) and Y = t[0] (even role, limb 0 now zero)
  merge8_ip(t[1], t[0]);
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = t[1][k];
  if (LAZY) return r;
  Fe<P> u = r;
  uint32_t pm[8];
#pragma unroll
  for (int k = 0; k < 8; k++) pm[k] = P::mod(k);
  uint32_t borrow = sub8_ip(u.v, pm);
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = borrow ? r.v[k] : u.v[k];
  return r;
#else
  // portable CIOS on 32-bit limbs with 64-bit accumulators
  uint32_t t[10];
  for (int k = 0; k < 10; k++) t[k] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) {
      c += (uint64_t)a.v[j] * b.v[i] + t[j];
      t[j] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[8] = (uint32_t)c;
    t[9] = (uint32_t)(c >> 32);
    uint32_t m = t[0] * P::INV;
    c = (uint64_t)m * P::mod(0) + t[0];
    c >>= 32;
    for (int j = 1; j < 8; j++) {
      c += (uint64_t)m * P::mod(j) + t[j];
      t[j - 1] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[7] = (uint32_t)c;
    t[8] = t[9] + (uint32_t)(c >> 32);
  }
  for (int k = 0; k < 8; k++) r.v[k] = t[k];
  if (LAZY) return r;
  Fe<P> u;
  int64_t bw = 0;
  for (int k = 0; k < 8; k++) {
    bw += (int64_t)r.v[k] - (int64_t)P::mod(k);
    u.v[k] = (uint32_t)bw;
    bw >>= 32;
  }
  return (bw && !t[8]) ? r : u;
#endif
}
template <class P>
G16_HD Fe<P> fmul(const Fe<P>& a, const Fe<P>& b) {
  return fmul_core<false>(a, b);
}

// Two independent Montgomery products with their carry chains interleaved row by row: the rows of one
// product do not depend on the rows of the other, so a warp exposes two independent IMAD.WIDE chains to the
// scheduler (used by the Fp2 layer, whose kernels run at only two warps per scheduler).
template <class P>
G16_HD void fmul2(const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d, Fe<P>& r1, Fe<P>& r2) {
#if defined(__CUDA_ARCH__)
  uint32_t t[2][8], u[2][8];
#pragma unroll
  for (int k = 0; k < 8; k++) t[0][k] = t[1][k] = u[0][k] = u[1][k] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t* X = t[i & 1];
    uint32_t* Y = t[(i + 1) & 1];
    uint32_t* X2 = u[i & 1];
    uint32_t* Y2 = u[(i + 1) & 1];
    const uint32_t bi = b.v[i], di = d.v[i];
    mad_row_shift(X, Y[0], a.v[1], a.v[3], a.v[5], a.v[7], bi);
    mad_row_shift(X2, Y2[0], c.v[1], c.v[3], c.v[5], c.v[7], di);
    mad_row_carry(Y, X[7], a.v[0], a.v[2], a.v[4], a.v[6], bi);
    mad_row_carry(Y2, X2[7], c.v[0], c.v[2], c.v[4], c.v[6], di);
    const uint32_t m = Y[0] * P::INV;
    const uint32_t m2 = Y2[0] * P::INV;
    mad_row_nc(X, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
    mad_row_nc(X2, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m2);
    mad_row_carry(Y, X[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
    mad_row_carry(Y2, X2[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m2);
  }
  merge8_ip(t[1], t[0]);
  merge8_ip(u[1], u[0]);
  uint32_t pm[8];
#pragma unroll
  for (int k = 0; k < 8; k++) pm[k] = P::mod(k);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    r1.v[k] = t[1][k];
    r2.v[k] = u[1][k];
  }
  Fe<P> v1 = r1, v2 = r2;
  uint32_t bw1 = sub8_ip(v1.v, pm);
  uint32_t bw2 = sub8_ip(v2.v, pm);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    r1.v[k] = bw1 ? r1.v[k] : v1.v[k];
    r2.v[k] = bw2 ? r2.v[k] : v2.v[k];
  }
#else
  r1 = fmul(a, b);
  r2 = fmul(c, d);
#endif
}

#if defined(__CUDA_ARCH__)
// Montgomery reduction of a 16-limb value T < p * 2^256: T / 2^256 mod p, fully reduced.  The reduction rows of
// fmul_core without its product rows: T's low half seeds the even accumulator, the high half is added at the end.
template <class P>
G16_D Fe<P> redc_wide(const uint32_t* T) {
  uint32_t t[2][8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    t[0][k] = T[k];
    t[1][k] = 0;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t* Y = t[i & 1];        // even-aligned array of this iteration
    uint32_t* X = t[(i + 1) & 1];  // odd-aligned array
    if (i > 0) redc_shift(X, Y[0]);
    const uint32_t m = Y[0] * P::INV;
    mad_row_nc(X, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
    mad_row_carry(Y, X[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
  }
  merge8_ip(t[0], t[1]);           // iteration 7: t[1] even role (limb 0 zero), t[0] odd role
  add8_ip(t[0], T + 8);            // + T_hi: below 2p
  Fe<P> r, u;
  uint32_t pm[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    r.v[k] = t[0][k];
    pm[k] = P::mod(k);
  }
  u = r;
  uint32_t borrow = sub8_ip(u.v, pm);
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = borrow ? r.v[k] : u.v[k];
  return r;
}
#endif

template <class P>
G16_HD Fe<P> fsqr(const Fe<P>& a) {
  return fmul(a, a);
}

// Montgomery form <-> standard form
template <class P>
G16_HD Fe<P> to_mont(const Fe<P>& a) {
  return fmul(a, Fe<P>::rsquared());
}
template <class P>
G16_HD Fe<P> from_mont(const Fe<P>& a) {
  Fe<P> one = Fe<P>::zero();
  one.v[0] = 1;
  return fmul(a, one);
}

// a^(p-2) (Fermat); a = 0 -> 0.  Not unrolled: it is used once per MSM / proof, not per element.
template <class P>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
Fe<P> finv(const Fe<P>& a) {
  uint32_t e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = P::mod(i);
  e[0] -= 2;  // both moduli end in ...01 / ...47: no borrow
  Fe<P> acc = Fe<P>::one();
#pragma unroll 1
  for (int i = 253; i >= 0; i--) {
    acc = fsqr(acc);
    if ((e[i >> 5] >> (i & 31)) & 1u) acc = fmul(acc, a);
  }
  return acc;
}

// small-exponent power (fields.nim:139-147 smallPowFr)
template <class P>
G16_HD Fe<P> fpow_u64(const Fe<P>& base, uint64_t e) {
  Fe<P> a = Fe<P>::one();
  Fe<P> s = base;
#pragma unroll 1
  while (e) {
    if (e & 1) a = fmul(a, s);
    e >>= 1;
    s = fsqr(s);
  }
  return a;
}

// ---------------------------------------------------------------------------------------
// Fp2 = Fp[u]/(u^2 + 1)   (fields.nim:27,30-32; c0 then c1 in memory)
// ---------------------------------------------------------------------------------------
struct alignas(16) Fp2 {
  Fp c0, c1;
  static G16_HD Fp2 zero() {
    Fp2 r;
    r.c0 = Fp::zero();
    r.c1 = Fp::zero();
    return r;
  }
  static G16_HD Fp2 one() {
    Fp2 r;
    r.c0 = Fp::one();
    r.c1 = Fp::zero();
    return r;
  }
};

G16_HD bool fis_zero(const Fp2& a) { return fis_zero(a.c0) && fis_zero(a.c1); }
G16_HD bool feq(const Fp2& a, const Fp2& b) { return feq(a.c0, b.c0) && feq(a.c1, b.c1); }
G16_HD Fp2 fadd(const Fp2& a, const Fp2& b) {
  Fp2 r;
  r.c0 = fadd(a.c0, b.c0);
  r.c1 = fadd(a.c1, b.c1);
  return r;
}
G16_HD Fp2 fsub(const Fp2& a, const Fp2& b) {
  Fp2 r;
  r.c0 = fsub(a.c0, b.c0);
  r.c1 = fsub(a.c1, b.c1);
  return r;
}
G16_HD Fp2 fneg(const Fp2& a) {
  Fp2 r;
  r.c0 = fneg(a.c0);
  r.c1 = fneg(a.c1);
  return r;
}
G16_HD Fp2 fdbl(const Fp2& a) { return fadd(a, a); }
// Out-of-line Fp multiply for the Fp2 layer: a G2 mixed addition contains 28 Fp multiplications; inlining
// them all makes a ~100 KB loop body that misses the instruction cache (ncu: no_instruction stalls), so the
// Fp2 operations call one shared copy instead (by-value arguments travel in registers).
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
Fp fmul_call(Fp a, Fp b) { return fmul(a, b); }

struct FpPair {
  Fp a, b;
};
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
FpPair fmul2_call(Fp a, Fp b, Fp c, Fp d) {
  FpPair r;
  fmul2(a, b, c, d, r.a, r.b);
  return r;
}

#ifndef G16_FP2_INLINE_MUL
#define G16_FP2_MUL(a, b) fmul_call(a, b)
#else
#define G16_FP2_MUL(a, b) fmul(a, b)
#endif

// G16_FP2_WHOLE_CALL: one out-of-line call per Fp2 operation (Karatsuba additions inside the callee) instead of
// one per Fp product pair: fewer argument moves, which ptxas places on the fma pipe as IMAD.MOV.
#ifdef G16_FP2_WHOLE_CALL
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
Fp2 fp2_mul_call(Fp2 a, Fp2 b) {
#if defined(__CUDA_ARCH__) && !defined(G16_FP2_NO_LAZY)
  // Lazy reduction across the Karatsuba: three unreduced 512-bit products, the additions and subtractions on the
  // wide values, two Montgomery reductions instead of three (336 instead of 408 MAC32 per Fp2 multiplication).
  //   c1 = REDC(T2 - T0 - T1),  T2 = (a0 + a1)(b0 + b1) with unreduced sums (< 2p < 2^255);  0 <= c1-value < 2 p^2
  //   c0 = REDC(T0 - T1 [+ p 2^256 if negative]);  both arguments stay below p 2^256, which REDC needs
  uint32_t T0[16], T1[16], T2[16], sa[8], sb[8];
  mul_wide(a.c0.v, b.c0.v, T0);
  mul_wide(a.c1.v, b.c1.v, T1);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    sa[k] = a.c0.v[k];
    sb[k] = b.c0.v[k];
  }
  add8_ip(sa, a.c1.v);
  add8_ip(sb, b.c1.v);
  mul_wide(sa, sb, T2);
  uint32_t bw = sub8_b(T2, T0, 0);
  sub8_b(T2 + 8, T0 + 8, bw);
  bw = sub8_b(T2, T1, 0);
  sub8_b(T2 + 8, T1 + 8, bw);
  bw = sub8_b(T0, T1, 0);
  bw = sub8_b(T0 + 8, T1 + 8, bw);
  uint32_t pm[8];
#pragma unroll
  for (int k = 0; k < 8; k++) pm[k] = bw ? FpParams::mod(k) : 0u;
  add8_ip(T0 + 8, pm);
  Fp2 r;
  r.c0 = redc_wide<FpParams>(T0);
  r.c1 = redc_wide<FpParams>(T2);
  return r;
#else
  Fp t0, t1;
  fmul2(a.c0, b.c0, a.c1, b.c1, t0, t1);
  Fp s = fmul(fadd(a.c0, a.c1), fadd(b.c0, b.c1));
  Fp2 r;
  r.c0 = fsub(t0, t1);
  r.c1 = fsub(fsub(s, t0), t1);
  return r;
#endif
}
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
Fp2 fp2_sqr_call(Fp2 a) {
  Fp t, u;
  fmul2(a.c0, a.c1, fadd(a.c0, a.c1), fsub(a.c0, a.c1), t, u);
  Fp2 r;
  r.c0 = u;
  r.c1 = fdbl(t);
  return r;
}
G16_HD Fp2 fmul(const Fp2& a, const Fp2& b) { return fp2_mul_call(a, b); }
G16_HD Fp2 fsqr(const Fp2& a) { return fp2_sqr_call(a); }
#else
G16_HD Fp2 fmul(const Fp2& a, const Fp2& b) {  // Karatsuba, 3 Fp mul (two of them interleaved)
#ifdef G16_FP2_NO_PAIRING
  Fp t0 = G16_FP2_MUL(a.c0, b.c0);
  Fp t1 = G16_FP2_MUL(a.c1, b.c1);
#else
  FpPair pr = fmul2_call(a.c0, b.c0, a.c1, b.c1);
  Fp t0 = pr.a, t1 = pr.b;
#endif
  Fp s = G16_FP2_MUL(fadd(a.c0, a.c1), fadd(b.c0, b.c1));
  Fp2 r;
  r.c0 = fsub(t0, t1);
  r.c1 = fsub(fsub(s, t0), t1);
  return r;
}
G16_HD Fp2 fsqr(const Fp2& a) {  // 2 Fp mul, interleaved
#ifdef G16_FP2_NO_PAIRING
  Fp t = G16_FP2_MUL(a.c0, a.c1);
  Fp2 r;
  r.c0 = G16_FP2_MUL(fadd(a.c0, a.c1), fsub(a.c0, a.c1));
#else
  FpPair pr = fmul2_call(a.c0, a.c1, fadd(a.c0, a.c1), fsub(a.c0, a.c1));
  Fp t = pr.a;
  Fp2 r;
  r.c0 = pr.b;
#endif
  r.c1 = fdbl(t);
  return r;
}
#endif
G16_HD Fp2 finv(const Fp2& a) {
  Fp d = finv(fadd(fsqr(a.c0), fsqr(a.c1)));
  Fp2 r;
  r.c0 = fmul(a.c0, d);
  r.c1 = fneg(fmul(a.c1, d));
  return r;
}

}  // namespace g16
