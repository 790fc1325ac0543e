// TEST INFRASTRUCTURE: compiles the product's field/EC headers with g++ (portable host path of
// field.cuh) so the formulas can be checked against the Python oracle without a GPU.
// Not part of the shipped library.
#include <string.h>
#include "field.cuh"
#include "ec.cuh"
#include "msm_digits.cuh"
#include "ntt_plan.cuh"
#include "msm_reduce_plan.cuh"
#include <vector>
using namespace g16;

// ---- host emulation of the pass structure of ntt.cu (same index helpers, same stage order) ----
static Fr gen28_mont() {
  static const uint32_t g[8] = {0x725b19f0u, 0x9bd61b6eu, 0x41112ed4u, 0x402d111eu,
                                0x8ef62abcu, 0x00e0a7ebu, 0xa58a7e85u, 0x2a3c09f0u};
  Fr x;
  for (int i = 0; i < 8; i++) x.v[i] = g[i];
  return to_mont(x);
}
struct EmuTables {
  std::vector<Fr> tw_fwd, tw_inv, coset;
  Fr n_inv;
};
static EmuTables emu_tables(int log_n) {
  EmuTables t;
  size_t n = (size_t)1 << log_n;
  Fr eta = gen28_mont();
  for (int i = 0; i < 28 - log_n - 1; i++) eta = fsqr(eta);
  Fr omega = fsqr(eta), omega_inv = finv(omega);
  Fr nn = Fr::zero();
  nn.v[0] = 1u << log_n;
  t.n_inv = finv(to_mont(nn));
  t.tw_fwd.resize(n / 2 ? n / 2 : 1);
  t.tw_inv.resize(n / 2 ? n / 2 : 1);
  t.coset.resize(n);
  Fr a = Fr::one(), b = Fr::one(), c = t.n_inv;
  for (size_t j = 0; j < n; j++) {
    if (j < n / 2) { t.tw_fwd[j] = a; t.tw_inv[j] = b; }
    t.coset[j] = c;
    a = fmul(a, omega); b = fmul(b, omega_inv); c = fmul(c, eta);
  }
  return t;
}
// one pass over all tiles; scale_mode 0 none, 1 const, 2 table[bitrev]
static void emu_pass(bool dif, std::vector<Fr>& src, std::vector<Fr>& dst, const std::vector<Fr>& tw, int log_n,
                     NttPass ps, int scale_mode, const Fr* scale, bool bitrev_store) {
  uint32_t E = 1u << (ps.k + ps.logC);
  uint32_t tiles = (1u << log_n) / E;
  std::vector<Fr> out(dst.size());
  bool alias = (&src == &dst);
  for (uint32_t tile = 0; tile < tiles; tile++) {
    std::vector<Fr> sm(E);
    for (uint32_t p = 0; p < E; p++) sm[p] = src[ntt_global_index(tile, p, ps.t_lo, ps.k, ps.logC)];
    for (int it = 0; it < ps.k; it++) {
      int s = dif ? (ps.k - 1 - it) : it;
      bool trivial = (ps.t_lo + s) == 0;
      for (uint32_t q = 0; q < E / 2; q++) {
        uint32_t pu, pv, e;
        ntt_butterfly_index(tile, q, s, ps.t_lo, ps.logC, log_n, pu, pv, e);
        Fr u = sm[pu], v = sm[pv];
        if (dif) {
          Fr d = fsub(u, v);
          if (!trivial) d = fmul(d, tw[e]);
          sm[pu] = fadd(u, v);
          sm[pv] = d;
        } else {
          if (!trivial) v = fmul(v, tw[e]);
          sm[pu] = fadd(u, v);
          sm[pv] = fsub(u, v);
        }
      }
    }
    for (uint32_t p = 0; p < E; p++) {
      uint32_t g = ntt_global_index(tile, p, ps.t_lo, ps.k, ps.logC);
      uint32_t gr = ntt_bitrev(g, log_n);
      Fr x = sm[p];
      if (scale_mode == 1) x = fmul(x, *scale);
      else if (scale_mode == 2) x = fmul(x, scale[gr]);
      (alias ? out : dst)[bitrev_store ? gr : g] = x;
    }
  }
  if (alias) dst = out;
}
static void emu_dif(std::vector<Fr>& v, std::vector<Fr>& dst, const std::vector<Fr>& tw, int log_n, int scale_mode,
                    const Fr* scale, bool bitrev_store) {
  NttPlan pl = ntt_make_plan(log_n);
  for (int i = pl.npass - 1; i >= 0; i--) {
    bool last = (i == 0);
    if (last) emu_pass(true, v, dst, tw, log_n, pl.pass[i], scale_mode, scale, bitrev_store);
    else emu_pass(true, v, v, tw, log_n, pl.pass[i], 0, nullptr, false);
  }
}
static void emu_dit(std::vector<Fr>& v, const std::vector<Fr>& tw, int log_n) {
  NttPlan pl = ntt_make_plan(log_n);
  for (int i = 0; i < pl.npass; i++) emu_pass(false, v, v, tw, log_n, pl.pass[i], 0, nullptr, false);
}

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
// natural-order NTT as ntt_natural() in ntt.cu
void he_ntt(const uint32_t* in, uint32_t* out, int log_n, int inverse) {
  size_t n = (size_t)1 << log_n;
  EmuTables t = emu_tables(log_n);
  std::vector<Fr> v((const Fr*)in, (const Fr*)in + n), dst(n);
  if (inverse) emu_dif(v, dst, t.tw_inv, log_n, 1, &t.n_inv, true);
  else emu_dif(v, dst, t.tw_fwd, log_n, 0, nullptr, true);
  memcpy(out, dst.data(), n * sizeof(Fr));
}
// shiftEvalDomain as quotient() in ntt.cu: DIF inverse with coset[bitrev] scaling, then DIT forward
void he_shift_eval(const uint32_t* in, uint32_t* out, int log_n) {
  size_t n = (size_t)1 << log_n;
  EmuTables t = emu_tables(log_n);
  std::vector<Fr> v((const Fr*)in, (const Fr*)in + n);
  emu_dif(v, v, t.tw_inv, log_n, 2, t.coset.data(), false);
  emu_dit(v, t.tw_fwd, log_n);
  memcpy(out, v.data(), n * sizeof(Fr));
}
}

// ---- host emulation of the bucket reduction (k_reduce_level x2, k_reduce_bits, k_reduce_final of msm_impl.cuh)
// driven by the product's own plan (msm_reduce_plan.cuh).  The group is Z / (2^61 - 1): the staging, the partial
// sum layout and the powers of two are what is checked, not the curve arithmetic.
static const uint64_t EMU_M = (1ull << 61) - 1;
static uint64_t emu_add(uint64_t a, uint64_t b) { return (a + b) % EMU_M; }
static uint64_t emu_dbl_n(uint64_t a, unsigned n) {
  for (unsigned i = 0; i < n; i++) a = emu_add(a, a);
  return a;
}
extern "C" {
// returns sum_k (k+1) * v[k] mod 2^61-1 computed through the staged plan; info[0..5] = L1, L2, nbits, nsmall,
// nchunks, launches
uint64_t he_reduce_emulate(const uint64_t* v, uint32_t nb, uint32_t* info) {
  ReducePlan p = msm_reduce_plan(nb);
  if (!p.ok) return ~0ull;
  // level 1 (weights j+1), one R partial per block of tpb1 threads
  std::vector<uint64_t> S1(p.n1), R1(p.blocks1, 0);
  for (uint32_t t = 0; t < p.n1; t++) {
    uint64_t running = 0, sum = 0;
    for (int k = (int)p.L1 - 1; k >= 0; k--) {
      running = emu_add(running, v[(size_t)t * p.L1 + k] % EMU_M);
      sum = emu_add(sum, running);
    }
    S1[t] = running;
    R1[t / p.tpb1] = emu_add(R1[t / p.tpb1], sum);
  }
  // small-level array [(level) * stride + k]
  std::vector<uint64_t> small((size_t)p.nsmall * p.stride + 1, 0);
  std::vector<uint64_t> S2;
  const std::vector<uint64_t>* S = &S1;
  if (p.L2 > 1) {                                  // level 2 (weights j), R partials into small level 0
    S2.resize(p.n2);
    for (uint32_t t = 0; t < p.n2; t++) {
      uint64_t running = 0, sum = 0;
      for (int k = (int)p.L2 - 1; k >= 0; k--) {
        running = emu_add(running, S1[(size_t)t * p.L2 + k]);
        if (k > 0) sum = emu_add(sum, running);
      }
      S2[t] = running;
      small[0 * p.stride + t / p.tpb2] = emu_add(small[0 * p.stride + t / p.tpb2], sum);
    }
    S = &S2;
  }
  for (uint32_t j = 0; j < p.nbits; j++)           // bit slices, one partial per chunk of 1024 entries
    for (uint32_t c = 0; c < p.nchunks; c++) {
      uint64_t acc = 0;
      for (uint32_t i = 0; i < REDUCE_BITS_CHUNK; i++) {
        uint32_t t = c * REDUCE_BITS_CHUNK + i;
        if (t < p.n2 && ((t >> j) & 1u)) acc = emu_add(acc, (*S)[t]);
      }
      small[(size_t)(p.first_bit_level + j) * p.stride + c] = acc;
    }
  // final: R1 + sum over small levels of 2^shift * (sum of count partials)
  uint64_t total = 0;
  for (uint32_t k = 0; k < p.blocks1; k++) total = emu_add(total, R1[k]);
  for (uint32_t l = 0; l < p.nsmall; l++) {
    uint64_t lv = 0;
    for (uint32_t k = 0; k < p.count[l]; k++) lv = emu_add(lv, small[(size_t)l * p.stride + k]);
    total = emu_add(total, emu_dbl_n(lv, p.shift[l]));
  }
  if (info) {
    info[0] = p.L1; info[1] = p.L2; info[2] = p.nbits; info[3] = p.nsmall; info[4] = p.nchunks;
    info[5] = 1 + (p.L2 > 1 ? 1 : 0) + (p.nbits ? 1 : 0) + 1;
  }
  return total;
}
}
