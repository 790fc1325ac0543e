// Short-Weierstrass arithmetic for BN254 G1 (over Fp) and G2 (over Fp2), a = 0.
//
// Replaces the constantine EC layer reached through groth16/bn128/curves.nim:33-37 (types),
// :136-154 (addG1/addG2), :182-214 (`**`) and msm.nim:54,81 (prj.affine).  Affine points use the
// reference's convention: infinity = (0,0) (curves.nim:49-50), x then y in memory.
// Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2);
// infinity <=> ZZ == 0, so zero-filled memory is a valid array of infinities.
#pragma once
#include "field.cuh"

#if defined(__CUDACC__)
#define G16_NI __host__ __device__ __noinline__
#else
#define G16_NI inline
#endif

namespace g16 {

template <class F>
struct alignas(16) Affine {
  F x, y;
};

template <class F>
struct alignas(16) XYZZ {
  F x, y, zz, zzz;
};

typedef Affine<Fp> G1Affine;   // 64 bytes  (curves.nim:33)
typedef Affine<Fp2> G2Affine;  // 128 bytes (curves.nim:34)
typedef XYZZ<Fp> G1XYZZ;
typedef XYZZ<Fp2> G2XYZZ;

template <class F>
G16_HD bool aff_is_inf(const Affine<F>& p) {
  return fis_zero(p.x) && fis_zero(p.y);
}
template <class F>
G16_HD Affine<F> aff_inf() {
  Affine<F> r;
  r.x = F::zero();
  r.y = F::zero();
  return r;
}
template <class F>
G16_HD Affine<F> aff_neg(const Affine<F>& p) {
  Affine<F> r;
  r.x = p.x;
  r.y = fneg(p.y);
  return r;
}

template <class F>
G16_HD bool xyzz_is_inf(const XYZZ<F>& p) {
  return fis_zero(p.zz);
}
template <class F>
G16_HD XYZZ<F> xyzz_inf() {
  XYZZ<F> r;
  r.x = F::zero();
  r.y = F::zero();
  r.zz = F::zero();
  r.zzz = F::zero();
  return r;
}
template <class F>
G16_HD XYZZ<F> xyzz_from_affine(const Affine<F>& p) {
  if (aff_is_inf(p)) return xyzz_inf<F>();
  XYZZ<F> r;
  r.x = p.x;
  r.y = p.y;
  r.zz = F::one();
  r.zzz = F::one();
  return r;
}
template <class F>
G16_HD XYZZ<F> xyzz_neg(const XYZZ<F>& p) {
  XYZZ<F> r = p;
  r.y = fneg(p.y);
  return r;
}

// 2 * (affine point)   [mdbl-2008-s-1: 2M + 3S... here 3S + 3M]
template <class F>
G16_HD XYZZ<F> xyzz_dbl_affine(const Affine<F>& p) {
  if (aff_is_inf(p) || fis_zero(p.y)) return xyzz_inf<F>();
  XYZZ<F> r;
  F U = fdbl(p.y);
  F V = fsqr(U);
  F W = fmul(U, V);
  F S = fmul(p.x, V);
  F X2 = fsqr(p.x);
  F M = fadd(fdbl(X2), X2);
  r.x = fsub(fsqr(M), fdbl(S));
  r.y = fsub(fmul(M, fsub(S, r.x)), fmul(W, p.y));
  r.zz = V;
  r.zzz = W;
  return r;
}

// 2 * P   [dbl-2008-s-1]
template <class F>
G16_HD XYZZ<F> xyzz_dbl(const XYZZ<F>& p) {
  if (xyzz_is_inf(p) || fis_zero(p.y)) return xyzz_inf<F>();
  XYZZ<F> r;
  F U = fdbl(p.y);
  F V = fsqr(U);
  F W = fmul(U, V);
  F S = fmul(p.x, V);
  F X2 = fsqr(p.x);
  F M = fadd(fdbl(X2), X2);
  r.x = fsub(fsqr(M), fdbl(S));
  r.y = fsub(fmul(M, fsub(S, r.x)), fmul(W, p.y));
  r.zz = fmul(V, p.zz);
  r.zzz = fmul(W, p.zzz);
  return r;
}

// Out-of-line copies for cold paths (rare doubling branches, reductions, scalar-mul loops): keeps
// the hot loops small and the compile time / register pressure of the Fp2 kernels in check.
template <class F>
G16_NI void xyzz_dbl_affine_ni(XYZZ<F>& r, const Affine<F>& p) { r = xyzz_dbl_affine(p); }
template <class F>
G16_NI void xyzz_dbl_ni(XYZZ<F>& r, const XYZZ<F>& p) { r = xyzz_dbl(p); }

// acc + (affine q)   [madd-2008-s: 8M + 2S], all special cases handled
template <class F>
G16_HD XYZZ<F> xyzz_madd(const XYZZ<F>& acc, const Affine<F>& q) {
  if (aff_is_inf(q)) return acc;
  if (xyzz_is_inf(acc)) return xyzz_from_affine(q);
  F U2 = fmul(q.x, acc.zz);
  F S2 = fmul(q.y, acc.zzz);
  F Pv = fsub(U2, acc.x);
  F Rv = fsub(S2, acc.y);
  if (fis_zero(Pv)) {
    XYZZ<F> d = xyzz_inf<F>();
    if (fis_zero(Rv)) xyzz_dbl_affine_ni(d, q);
    return d;
  }
  F PP = fsqr(Pv);
  F PPP = fmul(Pv, PP);
  F Q = fmul(acc.x, PP);
  XYZZ<F> r;
  r.x = fsub(fsub(fsqr(Rv), PPP), fdbl(Q));
  r.y = fsub(fmul(Rv, fsub(Q, r.x)), fmul(acc.y, PPP));
  r.zz = fmul(acc.zz, PP);
  r.zzz = fmul(acc.zzz, PPP);
  return r;
}

// p + q   [add-2008-s: 12M + 2S], all special cases handled
template <class F>
G16_HD XYZZ<F> xyzz_add(const XYZZ<F>& p, const XYZZ<F>& q) {
  if (xyzz_is_inf(q)) return p;
  if (xyzz_is_inf(p)) return q;
  F U1 = fmul(p.x, q.zz);
  F U2 = fmul(q.x, p.zz);
  F S1 = fmul(p.y, q.zzz);
  F S2 = fmul(q.y, p.zzz);
  F Pv = fsub(U2, U1);
  F Rv = fsub(S2, S1);
  if (fis_zero(Pv)) {
    XYZZ<F> d = xyzz_inf<F>();
    if (fis_zero(Rv)) xyzz_dbl_ni(d, p);
    return d;
  }
  F PP = fsqr(Pv);
  F PPP = fmul(Pv, PP);
  F Q = fmul(U1, PP);
  XYZZ<F> r;
  r.x = fsub(fsub(fsqr(Rv), PPP), fdbl(Q));
  r.y = fsub(fmul(Rv, fsub(Q, r.x)), fmul(S1, PPP));
  r.zz = fmul(fmul(p.zz, q.zz), PP);
  r.zzz = fmul(fmul(p.zzz, q.zzz), PPP);
  return r;
}

template <class F>
G16_NI void xyzz_add_ni(XYZZ<F>& r, const XYZZ<F>& p, const XYZZ<F>& q) { r = xyzz_add(p, q); }
template <class F>
G16_NI void xyzz_madd_ni(XYZZ<F>& r, const XYZZ<F>& p, const Affine<F>& q) { r = xyzz_madd(p, q); }

// XYZZ -> affine, infinity -> (0,0)   (msm.nim:54 prj.affine; curves.nim:49-50)
template <class F>
G16_HD Affine<F> xyzz_to_affine(const XYZZ<F>& p) {
  if (xyzz_is_inf(p)) return aff_inf<F>();
  F i = finv(fmul(p.zz, p.zzz));
  Affine<F> r;
  r.x = fmul(p.x, fmul(i, p.zzz));  // X / ZZ
  r.y = fmul(p.y, fmul(i, p.zz));   // Y / ZZZ
  return r;
}

template <class F>
G16_NI void xyzz_to_affine_ni(Affine<F>& r, const XYZZ<F>& p) { r = xyzz_to_affine(p); }

// k * P for a standard-form (non-Montgomery) 256-bit scalar, MSB-first double-and-add.
// (curves.nim:182-196 `**`; used for the six mask terms of prover.nim:279-300.)
template <class F>
G16_HD XYZZ<F> xyzz_scalar_mul(const uint32_t k[8], const Affine<F>& p) {
  XYZZ<F> acc = xyzz_inf<F>();
  int top = 255;
  while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1u)) top--;
#pragma unroll 1
  for (int i = top; i >= 0; i--) {
    xyzz_dbl_ni(acc, acc);
    if ((k[i >> 5] >> (i & 31)) & 1u) xyzz_madd_ni(acc, acc, p);
  }
  return acc;
}

// small-integer multiple of an XYZZ point (bucket-reduction segment offsets)
template <class F>
G16_HD XYZZ<F> xyzz_mul_u32(uint32_t k, const XYZZ<F>& p) {
  XYZZ<F> acc = xyzz_inf<F>();
  if (k == 0) return acc;
  int top = 31;
  while (!((k >> top) & 1u)) top--;
#pragma unroll 1
  for (int i = top; i >= 0; i--) {
    xyzz_dbl_ni(acc, acc);
    if ((k >> i) & 1u) xyzz_add_ni(acc, acc, p);
  }
  return acc;
}

}  // namespace g16
