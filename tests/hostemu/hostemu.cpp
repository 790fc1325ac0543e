// TEST INFRASTRUCTURE: compiles the product's field/EC headers with g++ (portable host path of
// field.cuh) so the formulas can be checked against the Python oracle without a GPU.
// Not part of the shipped library.
#include <string.h>
#include "field.cuh"
#include "ec.cuh"
#include "msm_digits.cuh"
using namespace g16;

extern "C" {
void he_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* c) { *(Fr*)c = fmul(*(const Fr*)a, *(const Fr*)b); }
void he_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* c) { *(Fp*)c = fmul(*(const Fp*)a, *(const Fp*)b); }
void he_fp_add(const uint32_t* a, const uint32_t* b, uint32_t* c) { *(Fp*)c = fadd(*(const Fp*)a, *(const Fp*)b); }
void he_fp_sub(const uint32_t* a, const uint32_t* b, uint32_t* c) { *(Fp*)c = fsub(*(const Fp*)a, *(const Fp*)b); }
void he_fr_inv(const uint32_t* a, uint32_t* c) { *(Fr*)c = finv(*(const Fr*)a); }
void he_fr_halve(const uint32_t* a, uint32_t* c) { *(Fr*)c = fhalve(*(const Fr*)a); }
void he_fr_to_mont(const uint32_t* a, uint32_t* c) { *(Fr*)c = to_mont(*(const Fr*)a); }
void he_fr_from_mont(const uint32_t* a, uint32_t* c) { *(Fr*)c = from_mont(*(const Fr*)a); }
void he_fp2_mul(const uint32_t* a, const uint32_t* b, uint32_t* c) { *(Fp2*)c = fmul(*(const Fp2*)a, *(const Fp2*)b); }
void he_fp2_sqr(const uint32_t* a, uint32_t* c) { *(Fp2*)c = fsqr(*(const Fp2*)a); }
void he_fp2_inv(const uint32_t* a, uint32_t* c) { *(Fp2*)c = finv(*(const Fp2*)a); }

// acc = sum of sign_i * pts[i] using madd, then affine
void he_g1_sum(const uint32_t* pts, const int* neg, int n, uint32_t* out) {
  G1XYZZ acc = xyzz_inf<Fp>();
  for (int i = 0; i < n; i++) {
    G1Affine p = ((const G1Affine*)pts)[i];
    if (neg[i]) p = aff_neg(p);
    acc = xyzz_madd(acc, p);
  }
  *(G1Affine*)out = xyzz_to_affine(acc);
}
void he_g2_sum(const uint32_t* pts, const int* neg, int n, uint32_t* out) {
  G2XYZZ acc = xyzz_inf<Fp2>();
  for (int i = 0; i < n; i++) {
    G2Affine p = ((const G2Affine*)pts)[i];
    if (neg[i]) p = aff_neg(p);
    acc = xyzz_madd(acc, p);
  }
  *(G2Affine*)out = xyzz_to_affine(acc);
}
// tree sum with full adds (exercises xyzz_add incl. doubling / cancellation)
void he_g1_treesum(const uint32_t* pts, int n, uint32_t* out) {
  G1XYZZ* v = new G1XYZZ[n > 0 ? n : 1];
  for (int i = 0; i < n; i++) v[i] = xyzz_from_affine(((const G1Affine*)pts)[i]);
  for (int s = 1; s < n; s *= 2)
    for (int i = 0; i + s < n; i += 2 * s) v[i] = xyzz_add(v[i], v[i + s]);
  *(G1Affine*)out = n ? xyzz_to_affine(v[0]) : aff_inf<Fp>();
  delete[] v;
}
void he_g2_treesum(const uint32_t* pts, int n, uint32_t* out) {
  G2XYZZ* v = new G2XYZZ[n > 0 ? n : 1];
  for (int i = 0; i < n; i++) v[i] = xyzz_from_affine(((const G2Affine*)pts)[i]);
  for (int s = 1; s < n; s *= 2)
    for (int i = 0; i + s < n; i += 2 * s) v[i] = xyzz_add(v[i], v[i + s]);
  *(G2Affine*)out = n ? xyzz_to_affine(v[0]) : aff_inf<Fp2>();
  delete[] v;
}
void he_g1_scalar_mul(const uint32_t* k, const uint32_t* p, uint32_t* out) {
  *(G1Affine*)out = xyzz_to_affine(xyzz_scalar_mul(k, *(const G1Affine*)p));
}
void he_g2_scalar_mul(const uint32_t* k, const uint32_t* p, uint32_t* out) {
  *(G2Affine*)out = xyzz_to_affine(xyzz_scalar_mul(k, *(const G2Affine*)p));
}
void he_g1_mul_u32(uint32_t k, const uint32_t* p, uint32_t* out) {
  *(G1Affine*)out = xyzz_to_affine(xyzz_mul_u32(k, xyzz_from_affine(*(const G1Affine*)p)));
}
// signed-digit decomposition used by the MSM (msm_digits.cuh)
void he_digits(const uint32_t* k, int c, int nwin, int* out) {
  for (int w = 0, carry = 0; w < nwin; w++) out[w] = msm_signed_digit(k, c, w, nwin, carry);
}
}
