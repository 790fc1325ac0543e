// Resident Groth16 proving context on one B200.
//
// Replaces generateProofWithMask (groth16/prover.nim:215-304): buildABC (:245), the quotient (:250-260),
// the five MSMs (:279-302) and the proof assembly (:278-304).  At g16_ctx_create the zkey's prover points
// are expanded ONCE into window tables 2^(c w) P_i that stay resident in HBM (about 13x the point bytes --
// 5 GB at 2^20, 21 GB at 2^22 of the 180 GB), and the coefficient list is sorted into CSR rows; per proof
// only the witness travels.  The four witness MSMs (A1, B1, B2, C1) read the same scalars, so they share
// one digit/sort pass (C1 is padded with npubs+1 infinities so its indices are witness indices) and the
// three G1 sets are accumulated by the same launches; B2 and the H chain (ABC -> quotient -> sort -> H1)
// run on their own streams; the mask terms that need no MSM result run on a fourth.  A context may own only
// the point range [N*k/G, N*(k+1)/G) of every MSM (msm.nim:107-115): run_msms() then yields partial sums
// that the host side all-gathers between GPUs.
#include "prover.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include "ntt.cuh"
#include "glv.h"

namespace g16 {

template <class T>
static __device__ __forceinline__ T ldv(const T* p) {
  T r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = q[i];
  return r;
}
template <class T>
static __device__ __forceinline__ void stv(T* p, const T& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) q[i] = s[i];
}

// -(r*s) mod the group order, standard form in and out
static __device__ void neg_rs(const uint32_t r[8], const uint32_t s[8], uint32_t out[8]) {
  Fr a, b;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a.v[i] = r[i];
    b.v[i] = s[i];
  }
  Fr p = fneg(fmul(to_mont(a), to_mont(b)));   // prover.nim:300 negFr(r*s)
  p = from_mont(p);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = p.v[i];
}

// Tables 2^j * delta1 (G1) and 2^j * delta2 (G2), j < 256, built once per context: the mask terms
// r ** delta1, s ** delta1, negFr(r*s) ** delta1, s ** delta2 (prover.nim:281,287,293,300) then need no
// doubling chain -- a block adds the table entries selected by the scalar's bits in a tree.
__global__ void k_delta_tables(const SpecPointsDev* spec, G1XYZZ* t1, G2XYZZ* t2, G1XYZZ* ta, G1XYZZ* tb) {
  if (threadIdx.x & 31) return;
  const int w = threadIdx.x >> 5;
  if (w == 1) {
    G2XYZZ p = xyzz_from_affine(ldv(&spec->delta2));
    for (int j = 0; j < 256; j++) {
      stv(t2 + j, p);
      xyzz_dbl_ni(p, p);
    }
  } else {
    G1XYZZ p = xyzz_from_affine(ldv(w == 0 ? &spec->delta1 : w == 2 ? &spec->alpha1 : &spec->beta1));
    G1XYZZ* t = w == 0 ? t1 : w == 2 ? ta : tb;
    for (int j = 0; j < 256; j++) {
      stv(t + j, p);
      xyzz_dbl_ni(p, p);
    }
  }
}

template <class F>
static __device__ void block_bits_sum(const uint32_t k[8], const XYZZ<F>* table, XYZZ<F>* red, XYZZ<F>& out) {
  const uint32_t j = threadIdx.x;                    // 256 threads, one per scalar bit
  XYZZ<F> acc = xyzz_inf<F>();
  if ((k[j >> 5] >> (j & 31)) & 1u) acc = ldv(table + j);
  red[j] = acc;
  __syncthreads();
  for (uint32_t s = 128; s > 0; s >>= 1) {
    if (j < s) {
      XYZZ<F> o = red[j + s];
      xyzz_add_ni(acc, acc, o);
      red[j] = acc;
    }
    __syncthreads();
  }
  out = acc;
}

// blocks 0..2: G1 terms over the delta1 table; block 3: the G2 term; blocks 4, 5: s ** alpha1 and r ** beta1 (used by
// the masked-partials finish only).  256 threads per block.
__global__ void __launch_bounds__(256) k_mask_terms(const SpecPointsDev* spec, const G1XYZZ* t1, const G2XYZZ* t2,
                                                    const G1XYZZ* ta, const G1XYZZ* tb, MaskTerms* m) {
  extern __shared__ uint4 red_raw[];
  __shared__ uint32_t k[8];
  const int term = blockIdx.x;
  if (threadIdx.x == 0) {
    if (term == 2) neg_rs(m->r, m->s, k);             // negFr(r*s)                 prover.nim:300
    else
      for (int i = 0; i < 8; i++) k[i] = (term == 0 || term == 5) ? m->r[i] : m->s[i];
  }
  __syncthreads();
  if (term != 3) {
    G1XYZZ* red = reinterpret_cast<G1XYZZ*>(red_raw);
    G1XYZZ acc;
    block_bits_sum<Fp>(k, term == 4 ? ta : term == 5 ? tb : t1, red, acc);
    if (threadIdx.x == 0) {
      if (term == 0) {                                // alpha1 + r ** delta1       prover.nim:280-281
        xyzz_madd_ni(acc, acc, ldv(&spec->alpha1));
        stv(&m->t_a, acc);
      } else if (term == 1) {                         // beta1 + s ** delta1        prover.nim:286-287
        xyzz_madd_ni(acc, acc, ldv(&spec->beta1));
        stv(&m->t_b1, acc);
      } else if (term == 2) {
        stv(&m->t_c, acc);
      } else if (term == 4) {
        stv(&m->t_sa, acc);
      } else {
        stv(&m->t_rb, acc);
      }
    }
  } else {
    G2XYZZ* red = reinterpret_cast<G2XYZZ*>(red_raw);
    G2XYZZ acc;
    block_bits_sum<Fp2>(k, t2, red, acc);
    if (threadIdx.x == 0) {                           // beta2 + s ** delta2        prover.nim:292-293
      xyzz_madd_ni(acc, acc, ldv(&spec->beta2));
      stv(&m->t_b2, acc);
    }
  }
}

// proof assembly, prover.nim:278-304, in two kernels so that the expensive half overlaps with the MSMs that
// finish later.
// early (needs only the G1 witness MSMs A1, B1, C1): pi_a, rho, s ** pi_a, r ** rho and
//   partial_c = s**pi_a + r**rho + (-rs)**delta1 + MSM(zs, C1)
// The two scalar multiplications do not walk a double-and-add chain of ~380 dependent point operations: one
// thread per half writes the 254 doublings 2^i P into a scratch table, then 128 threads add the entries selected
// by the scalar's bits in a tree (same scheme as the delta tables of k_mask_terms).  256 threads = 2 halves.
// k * P for the two witness-dependent products of a proof, 128 threads per product (h = which half of the block):
// one thread writes the doublings 2^i P into `tbl`, then every thread adds the entries its scalar bits select and a
// tree sums the 128 partial results.  With the GLV split k = k1 + k2 lambda, |k_i| < 2^127 (glv.h), the table needs 127
// instead of 254 SEQUENTIAL doublings -- the latency of this kernel -- and thread j adds (+-) 2^j P for bit j of k1 and
// (+-) phi(2^j P) = (beta X, Y, ZZ, ZZZ) for bit j of k2.  Returns the sum in thread 0 of the half (red[h * 128]).
static __device__ void glv_scalar_mul_half(const G1XYZZ& base, bool base_ready_in_thread0, const uint32_t* k256,
                                           const uint32_t* glv, bool glv_ok, G1XYZZ* tbl, G1XYZZ* red, G1XYZZ& out) {
  const uint32_t j = threadIdx.x & 127u;
  const int ndbl = glv_ok ? 128 : 256;
  if (j == 0 && base_ready_in_thread0) {
    G1XYZZ p = base;
    for (int i = 0; i < ndbl; i++) {
      stv(tbl + i, p);
      if (i < ndbl - 1) xyzz_dbl_ni(p, p);
    }
  }
  __syncthreads();
  G1XYZZ acc = xyzz_inf<Fp>();
  if (glv_ok) {
    if ((glv[j >> 5] >> (j & 31)) & 1u) {
      acc = ldv(tbl + j);
      if (glv[8]) acc.y = fneg(acc.y);
    }
    if ((glv[4 + (j >> 5)] >> (j & 31)) & 1u) {
      // beta = 0x59e26bcea0d48bacd4f263f1acdb5c4f5763473177fffffe: the cube root of unity with phi(P) = lambda P
      Fp beta;
      const uint32_t bw[8] = {0x77fffffeu, 0x57634731u, 0xacdb5c4fu, 0xd4f263f1u, 0xa0d48bacu, 0x59e26bceu, 0u, 0u};
#pragma unroll
      for (int i = 0; i < 8; i++) beta.v[i] = bw[i];
      beta = to_mont(beta);
      G1XYZZ t = ldv(tbl + j);
      t.x = fmul(t.x, beta);
      if (glv[9]) t.y = fneg(t.y);
      xyzz_add_ni(acc, acc, t);
    }
  } else {
    if ((k256[j >> 5] >> (j & 31)) & 1u) acc = ldv(tbl + j);
    const uint32_t j2 = j + 128;
    if ((k256[j2 >> 5] >> (j2 & 31)) & 1u) xyzz_add_ni(acc, acc, ldv(tbl + j2));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t st = 64; st > 0; st >>= 1) {
    if (j < st) {
      G1XYZZ o = red[threadIdx.x + st];
      xyzz_add_ni(acc, acc, o);
      red[threadIdx.x] = acc;
    }
    __syncthreads();
  }
  out = acc;
}

__global__ void __launch_bounds__(256) k_assemble_early(const MsmResults* res, const MaskTerms* m, g16_proof* proof,
                                                        G1XYZZ* partial_c, G1XYZZ* scratch) {
  __shared__ G1XYZZ red[256];
  const uint32_t h = threadIdx.x >> 7, j = threadIdx.x & 127u;
  G1XYZZ base = xyzz_inf<Fp>();
  if (j == 0) {
    G1XYZZ t;
    if (h == 0) xyzz_add_ni(t, ldv(&m->t_a), ldv(&res->a1));        // pi_a                 prover.nim:282
    else xyzz_add_ni(t, ldv(&m->t_b1), ldv(&res->b1));              // rho                  prover.nim:288
    G1Affine pa;
    xyzz_to_affine_ni(pa, t);
    if (h == 0) stv(reinterpret_cast<G1Affine*>(proof->pi_a), pa);
    base = xyzz_from_affine(pa);
  }
  G1XYZZ acc;                                                       // s ** pi_a, r ** rho  prover.nim:298,299
  glv_scalar_mul_half(base, true, h == 0 ? m->s : m->r, m->glv[h == 0 ? 1 : 0], m->glv_ok != 0, scratch + h * 256, red,
                      acc);
  if (threadIdx.x == 0) {                             // + negFr(r*s) ** delta1 + MSM(zs, C1)   prover.nim:300,302
    G1XYZZ t;
    xyzz_add_ni(t, acc, red[128]);
    xyzz_add_ni(t, t, ldv(&m->t_c));
    xyzz_add_ni(t, t, ldv(&res->c1));
    stv(partial_c, t);
  }
}

// final (needs B2 and H1): pi_b and pi_c
__global__ void __launch_bounds__(64) k_assemble_final(const MsmResults* res, const MaskTerms* m,
                                                       const G1XYZZ* partial_c, g16_proof* proof) {
  if (threadIdx.x & 31) return;
  if ((threadIdx.x >> 5) == 0) {                      // pi_c = partial_c + MSM(qs, H1)          prover.nim:301
    G1XYZZ t;
    xyzz_add_ni(t, ldv(partial_c), ldv(&res->h1));
    G1Affine pc;
    xyzz_to_affine_ni(pc, t);
    stv(reinterpret_cast<G1Affine*>(proof->pi_c), pc);
  } else {                                            // pi_b                                     prover.nim:294
    G2XYZZ t;
    xyzz_add_ni(t, ldv(&m->t_b2), ldv(&res->b2));
    G2Affine pb;
    xyzz_to_affine_ni(pb, t);
    stv(reinterpret_cast<G2Affine*>(proof->pi_b), pb);
  }
}

// Masked partials (Prover::set_mask): with pi_a = alpha1 + r delta1 + sum_k A_k and rho = beta1 + s delta1 + sum_k B1_k,
//   pi_c = sum_k (C_k + s A_k + r B1_k) + sum_k H_k + s alpha1 + r beta1 + (r s) delta1          (prover.nim:298-302)
// so every shard multiplies ITS OWN partial sums by s and r -- overlapped with its B2 / H work, like the early
// assembly of the single-GPU path -- and ships c1' = C_k + s A_k + r B1_k; the finish is additions only.
__global__ void __launch_bounds__(256) k_shard_early(MsmResults* res, const MaskTerms* m, G1XYZZ* scratch) {
  __shared__ G1XYZZ red[256];
  const uint32_t h = threadIdx.x >> 7, j = threadIdx.x & 127u;
  G1XYZZ base = xyzz_inf<Fp>();
  if (j == 0) base = ldv(h == 0 ? &res->a1 : &res->b1);
  G1XYZZ acc;
  glv_scalar_mul_half(base, true, h == 0 ? m->s : m->r, m->glv[h == 0 ? 1 : 0], m->glv_ok != 0, scratch + h * 256, red,
                      acc);
  if (threadIdx.x == 0) {
    G1XYZZ t;
    xyzz_add_ni(t, acc, red[128]);
    xyzz_add_ni(t, t, ldv(&res->c1));
    stv(&res->c1, t);
  }
}

// finish of masked partials: res holds the sums over all shards (c1 = sum of the c1' records)
__global__ void __launch_bounds__(96) k_assemble_final_masked(const MsmResults* res, const MaskTerms* m,
                                                              g16_proof* proof) {
  if (threadIdx.x & 31) return;
  const int w = threadIdx.x >> 5;
  if (w == 0) {                                       // pi_a                                    prover.nim:282
    G1XYZZ t;
    xyzz_add_ni(t, ldv(&m->t_a), ldv(&res->a1));
    G1Affine a;
    xyzz_to_affine_ni(a, t);
    stv(reinterpret_cast<G1Affine*>(proof->pi_a), a);
  } else if (w == 1) {                                // pi_b                                    prover.nim:294
    G2XYZZ t;
    xyzz_add_ni(t, ldv(&m->t_b2), ldv(&res->b2));
    G2Affine b;
    xyzz_to_affine_ni(b, t);
    stv(reinterpret_cast<G2Affine*>(proof->pi_b), b);
  } else {                                            // pi_c
    G1XYZZ t;
    xyzz_add_ni(t, ldv(&res->c1), ldv(&res->h1));
    xyzz_add_ni(t, t, ldv(&m->t_sa));
    xyzz_add_ni(t, t, ldv(&m->t_rb));
    xyzz_add_ni(t, t, xyzz_neg(ldv(&m->t_c)));        // + (r s) delta1 = -((-r s) delta1)
    G1Affine c;
    xyzz_to_affine_ni(c, t);
    stv(reinterpret_cast<G1Affine*>(proof->pi_c), c);
  }
}

// XYZZ results -> affine partial sums (the per-chunk prj.affine of msm.nim:54,81)
__global__ void __launch_bounds__(160) k_partials_to_affine(const MsmResults* res, PartialsAffine* out, uint64_t masked,
                                                            uint64_t mask_hash) {
  if (threadIdx.x == 1) {
    out->tag[0] = masked;
    out->tag[1] = mask_hash;
  }
  if (threadIdx.x & 31) return;
  int w = threadIdx.x >> 5;
  if (w < 4) {
    const G1XYZZ* src = w == 0 ? &res->a1 : w == 1 ? &res->b1 : w == 2 ? &res->h1 : &res->c1;
    G1Affine* dst = w == 0 ? &out->a1 : w == 1 ? &out->b1 : w == 2 ? &out->h1 : &out->c1;
    G1Affine a;
    xyzz_to_affine_ni(a, ldv(src));
    stv(dst, a);
  } else {
    G2Affine a;
    xyzz_to_affine_ni(a, ldv(&res->b2));
    stv(&out->b2, a);
  }
}

// gathered affine partial sums -> XYZZ totals (res += sync pending[k], msm.nim:117-119)
// The records must agree with the finishing context on the mask convention (all masked with this (r, s), or none):
// a mix would give a wrong pi_c, so it is reported through the status word instead.
__global__ void __launch_bounds__(160) k_sum_partials(const PartialsAffine* parts, int count, MsmResults* res,
                                                      uint64_t masked, uint64_t mask_hash, uint32_t* status) {
  if (threadIdx.x == 1) {
    uint32_t bad = 0;
    for (int i = 0; i < count; i++)
      if (parts[i].tag[0] != masked || parts[i].tag[1] != mask_hash) bad = 1;
    *status = bad;
  }
  if (threadIdx.x & 31) return;
  int w = threadIdx.x >> 5;
  if (w < 4) {
    G1XYZZ acc = xyzz_inf<Fp>();
    for (int i = 0; i < count; i++) {
      const PartialsAffine* p = parts + i;
      const G1Affine* src = w == 0 ? &p->a1 : w == 1 ? &p->b1 : w == 2 ? &p->h1 : &p->c1;
      xyzz_madd_ni(acc, acc, ldv(src));
    }
    G1XYZZ* dst = w == 0 ? &res->a1 : w == 1 ? &res->b1 : w == 2 ? &res->h1 : &res->c1;
    stv(dst, acc);
  } else {
    G2XYZZ acc = xyzz_inf<Fp2>();
    for (int i = 0; i < count; i++) xyzz_madd_ni(acc, acc, ldv(&parts[i].b2));
    stv(&res->b2, acc);
  }
}

// ---------------------------------------------------------------------------------------
// Shard plan: which points of which MSM a rank owns.
//
// The reference chunks every MSM alike, [N*k/G, N*(k+1)/G) per thread (msm.nim:107-111).  Across GPUs that makes
// every rank run all five MSMs on a G-times smaller input: five bucket-set reductions, five latency-bound tails and
// a smaller window (more pairs per point) on every rank -- the fixed costs that capped 8-GPU efficiency at 0.55.
// The default policy "line" therefore places the MSMs themselves: the work of one proof is laid
// on a line, e.g.  [H | A1 | B1 | C1 | B2], and rank k owns the k-th segment, i.e. whole MSMs plus at most a tail of
// one array and a head of another.  buildABC and the quotient cannot be split by points, so every rank that owns H
// points computes them.  The planner tries a handful of line orders and every number m of H ranks, cuts each
// candidate so that no rank costs more than a bound L (bisection on L), and keeps the plan whose slowest rank is
// cheapest according to `rank_cost`, a model fitted to measured busy times of single shard shapes
// (tools/shape_probe.py, tools/fit_shard_model.py, profiles/r2_v9_shape_probe.log).
// Environment G16_SHARD_POLICY: "line" (default), "uniform" (the reference's equal chunks of every array), "g2own"
// (BASELINE.json configs[3]: the G2 MSM alone on the last rank, the rest as "line").
// Every rank evaluates the same arithmetic, so the plan needs no communication.
// ---------------------------------------------------------------------------------------
static void shard_range(size_t N, int k, int G, size_t& lo, size_t& hi) {   // msm.nim:107-111
  lo = (N * (size_t)k) / (size_t)G;
  hi = (k == G - 1) ? N : (N * (size_t)(k + 1)) / (size_t)G;
}

// Busy time of one rank with two proofs in flight, in units of one 254-bit Montgomery multiplication at the measured
// peak rate (6.76e10 /s on a B200 -- only ratios matter).  Least-squares fit to 30 shard shapes at 2^20 (rms error
// 0.09 ms, worst 4.7 %): a sorted pair costs 1.24, an accumulated pair 10.3 (G1, 10 multiplications at 0.97 of peak)
// or 30.7 (G2), a reduced bucket 24 / 84.5; every rank pays a fixed 0.46 ms of latency-bound tails once (they hide
// behind the rank's other work, so a second group adds only 0.05 ms), the quotient runs at 0.79 of the modmul peak,
// the witness-dependent scalar multiplications of a rank that owns A1 or B1 points cost 0.21 ms, a G2 bucket set
// 0.22 ms, and a G2 MSM that shares the rank with G1 work of another range costs 0.6 ms extra.
static double rank_cost(const ShardPlan& p, size_t nvars, size_t n) {
  (void)nvars;
  struct Grp { size_t lo, hi; int n1; bool g2; };
  Grp grp[5];
  int ng = 0;
  const size_t r[4][2] = {{p.a1_lo, p.a1_hi}, {p.b1_lo, p.b1_hi}, {p.c1_lo, p.c1_hi}, {p.b2_lo, p.b2_hi}};
  for (int a = 0; a < 4; a++) {
    if (r[a][1] <= r[a][0]) continue;
    int g = -1;
    for (int i = 0; i < ng; i++)
      if (grp[i].lo == r[a][0] && grp[i].hi == r[a][1]) g = i;
    if (g < 0) {
      g = ng++;
      grp[g] = {r[a][0], r[a][1], 0, false};
    }
    if (a == 3) grp[g].g2 = true;
    else grp[g].n1++;
  }
  const bool has_h = p.h_hi > p.h_lo, has_g2 = r[3][1] > r[3][0];
  if (has_h) grp[ng++] = {p.h_lo, p.h_hi, 1, false};
  if (!ng) return 0.0;
  double cost = 33.4e6 + 17.6e6 * (ng - 1);
  bool any_g1 = false;
  for (int i = 0; i < ng; i++) {
    const size_t npts = grp[i].hi - grp[i].lo;
    const int c = msm_pick_window(npts, true);
    const double pairs = (double)npts * (double)msm_num_windows(c), nb = (double)((size_t)1 << (c - 1));
    cost += pairs * 0.95 + grp[i].n1 * (pairs * 10.9 + nb * 15.0);
    if (grp[i].g2) cost += pairs * 31.4 + nb * 74.0;
    if (grp[i].n1) any_g1 = true;
  }
  if (has_g2) cost += 11.5e6 + (any_g1 ? 27.6e6 : 0.0);
  if (r[0][1] > r[0][0] || r[1][1] > r[1][0]) cost += 9.0e6;
  if (has_h) {
    const double lg = (double)ceil_log2_sz(n);
    // 6 NTTs + scaling + pointwise at 0.89 of the modmul peak; buildABC gathers ~3n witness values:
    // 0.085 ms at 2^20, where the witness stays in the 126 MB L2, 1.58 ms at 2^22, where it does not
    cost += (3.0 * ((double)n * lg + (double)n) + 2.0 * (double)n) / 0.89;
    cost += (double)n * ((double)nvars * 32.0 > 64.0e6 ? 25.0 : 5.5);
  }
  return cost;
}

static int shard_policy() {            // 0 = line, 1 = uniform, 2 = g2own
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("G16_SHARD_POLICY");
    v = (e && e[0] == 'u') ? 1 : (e && e[0] == 'g') ? 2 : 0;
  }
  return v;
}

// A unit of the line: arrays that are cut at the same witness indices, so that both sides of a cut keep one sort
// for all of them.  Arrays: 0 = A1, 1 = B1, 2 = C1, 3 = B2.
struct LineUnit { int arr[3]; int narr; };
struct LineOrder { LineUnit u[4]; int nu; };

static void set_range(ShardPlan& p, int arr, size_t lo, size_t hi) {
  size_t* l[4] = {&p.a1_lo, &p.b1_lo, &p.c1_lo, &p.b2_lo};
  size_t* h[4] = {&p.a1_hi, &p.b1_hi, &p.c1_hi, &p.b2_hi};
  *l[arr] = lo;
  *h[arr] = hi;
}

// Ranks [0, G) walk the line: rank k < m starts with the k-th of m equal pieces of H, then takes points of the
// current unit for as long as its modelled cost stays below L; the last rank takes what is left.  Returns the cost of
// the slowest rank.
static double fill_line(const LineOrder& ord, size_t nvars, size_t n, int G, int m, double L, std::vector<ShardPlan>& out) {
  const size_t snap = nvars / 8;        // pieces below an eighth of an array are not worth a sort and a reduction
  int cu = 0;                           // current unit
  size_t cp = 0;                        // next unassigned witness index of the current unit
  double worst = 0.0;
  for (int k = 0; k < G; k++) {
    ShardPlan& p = out[k];
    memset(&p, 0, sizeof(p));
    if (k < m) shard_range(n, k, m, p.h_lo, p.h_hi);
    const bool last = k == G - 1;
    while (cu < ord.nu) {
      const LineUnit& u = ord.u[cu];
      auto with = [&](size_t hi) {
        ShardPlan q = p;
        for (int i = 0; i < u.narr; i++) set_range(q, u.arr[i], cp, hi);
        return q;
      };
      size_t take_hi = nvars;
      if (!last && rank_cost(with(nvars), nvars, n) > L) {
        size_t lo = cp, hi = nvars;     // cost(with(lo)) <= L or nothing fits, cost(with(hi)) > L
        while (hi - lo > 1) {
          const size_t mid = lo + (hi - lo) / 2;
          if (rank_cost(with(mid), nvars, n) <= L) lo = mid;
          else hi = mid;
        }
        take_hi = lo;
        if (take_hi - cp <= snap) take_hi = cp;                               // a sliver: leave it to the next rank
        else if (nvars - take_hi <= snap) {                                   // would leave a sliver behind
          take_hi = rank_cost(with(nvars), nvars, n) <= L * 1.03 ? nvars : nvars - snap - 1;
          if (take_hi - cp <= snap) take_hi = cp;
        }
      }
      if (take_hi == cp) break;
      p = with(take_hi);
      cp = take_hi;
      if (cp == nvars) {
        cu++;
        cp = 0;
      } else break;                     // this rank is full
    }
    const double c = rank_cost(p, nvars, n);
    if (c > worst) worst = c;
  }
  return cu < ord.nu ? 1e300 : worst;
}

// best plan of `G` ranks for the arrays of `orders` (with_b2 = false: the G2 MSM is placed by the caller)
static void plan_line(size_t nvars, size_t n, int G, const bool with_b2, std::vector<ShardPlan>& out) {
  static const LineOrder kOrders[] = {
      {{{{0}, 1}, {{1}, 1}, {{2}, 1}, {{3}, 1}}, 4},          // H | A1 | B1 | C1 | B2
      {{{{0, 1}, 2}, {{2}, 1}, {{3}, 1}}, 3},                 // H | A1+B1 | C1 | B2
      {{{{2}, 1}, {{3}, 1}, {{0, 1}, 2}}, 3},                 // H | C1 | B2 | A1+B1
      {{{{2}, 1}, {{0, 1}, 2}, {{3}, 1}}, 3},                 // H | C1 | A1+B1 | B2
      {{{{0, 1, 2}, 3}, {{3}, 1}}, 2},                        // H | A1+B1+C1 | B2
      {{{{3}, 1}, {{2}, 1}, {{0, 1}, 2}}, 3},                 // H | B2 | C1 | A1+B1
  };
  ShardPlan whole;
  memset(&whole, 0, sizeof(whole));
  whole.a1_hi = whole.b1_hi = whole.c1_hi = nvars;
  if (with_b2) whole.b2_hi = nvars;
  whole.h_hi = n;
  const double all = rank_cost(whole, nvars, n);
  double best = 1e300;
  std::vector<ShardPlan> cand((size_t)G);
  for (const LineOrder& full : kOrders) {
    LineOrder ord = full;
    if (!with_b2) {                      // drop the B2 unit
      ord.nu = 0;
      for (int i = 0; i < full.nu; i++)
        if (full.u[i].arr[0] != 3) ord.u[ord.nu++] = full.u[i];
    }
    for (int m = 1; m <= G && m <= 8; m++) {
      double lo = all / (double)G * 0.5, hi = all * 1.05;
      for (int it = 0; it < 28; it++) {
        const double L = 0.5 * (lo + hi);
        const double worst = fill_line(ord, nvars, n, G, m, L, cand);
        // ties go to the earlier order and to fewer redundant quotients
        if (worst < best * 0.99) {
          best = worst;
          out = cand;
        }
        if (worst <= L * 1.03) hi = L;
        else lo = L;
      }
    }
  }
}

// the search costs a few hundred thousand model evaluations: keep the plans of the last few (key, G) seen
struct PlanCacheEntry { size_t nvars, n; int G, policy; std::vector<ShardPlan> plan; };
static std::mutex g_plan_mu;
static std::vector<PlanCacheEntry> g_plan_cache;
static bool plan_cache_get(size_t nvars, size_t n, int G, int policy, std::vector<ShardPlan>& out) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  for (auto& e : g_plan_cache)
    if (e.nvars == nvars && e.n == n && e.G == G && e.policy == policy) {
      out = e.plan;
      return true;
    }
  return false;
}
static void plan_cache_put(size_t nvars, size_t n, int G, int policy, const std::vector<ShardPlan>& plan) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  for (auto& e : g_plan_cache)
    if (e.nvars == nvars && e.n == n && e.G == G && e.policy == policy) return;
  if (g_plan_cache.size() >= 16) g_plan_cache.erase(g_plan_cache.begin());
  g_plan_cache.push_back({nvars, n, G, policy, plan});
}

void shard_plan(size_t nvars, size_t npubs, size_t n, int k, int G, ShardPlan& out) {
  (void)npubs;
  std::vector<ShardPlan> all((size_t)G);
  for (auto& p : all) memset(&p, 0, sizeof(p));
  const int policy = shard_policy();
  if (G == 1 || policy == 1) {
    for (int r = 0; r < G; r++) {
      ShardPlan& p = all[r];
      shard_range(nvars, r, G, p.a1_lo, p.a1_hi);
      p.b1_lo = p.c1_lo = p.b2_lo = p.a1_lo;
      p.b1_hi = p.c1_hi = p.b2_hi = p.a1_hi;
      shard_range(n, r, G, p.h_lo, p.h_hi);
    }
  } else if (plan_cache_get(nvars, n, G, policy, all)) {
  } else if (policy == 2) {
    std::vector<ShardPlan> head((size_t)(G - 1));
    plan_line(nvars, n, G - 1, false, head);
    for (int r = 0; r + 1 < G; r++) all[r] = head[r];
    all[G - 1].b2_lo = 0;
    all[G - 1].b2_hi = nvars;
  } else {
    plan_line(nvars, n, G, true, all);
  }
  if (G > 1 && policy != 1) plan_cache_put(nvars, n, G, policy, all);
  if (k == 0 && G > 1 && getenv("G16_PLAN_DEBUG"))       // modelled busy time per rank, ms at the B200 modmul peak
    for (int r = 0; r < G; r++) fprintf(stderr, "[g16] plan rank %d of %d: model %.3f ms\n", r, G, rank_cost(all[r], nvars, n) / 6.76e7);
  out = all[k];
  // experiment knob (tools/shape_probe.py): ten fractions a1_lo,a1_hi,b1_lo,b1_hi,c1_lo,c1_hi,b2_lo,b2_hi,h_lo,h_hi
  // replace the plan of whatever rank is asked for -- used to time single shard shapes when fitting the cost model
  if (const char* e = G > 1 ? getenv("G16_SHARD_SHAPE") : nullptr) {
    double f[10];
    if (sscanf(e, "%lf,%lf,%lf,%lf,%lf,%lf,%lf,%lf,%lf,%lf", f, f + 1, f + 2, f + 3, f + 4, f + 5, f + 6, f + 7, f + 8,
               f + 9) == 10) {
      size_t* dst[10] = {&out.a1_lo, &out.a1_hi, &out.b1_lo, &out.b1_hi, &out.c1_lo,
                         &out.c1_hi, &out.b2_lo, &out.b2_hi, &out.h_lo,  &out.h_hi};
      for (int i = 0; i < 10; i++) {
        const double v = f[i] < 0.0 ? 0.0 : f[i] > 1.0 ? 1.0 : f[i];
        *dst[i] = (size_t)(v * (double)(i < 8 ? nvars : n) + 0.5);
      }
      if (getenv("G16_PLAN_DEBUG")) fprintf(stderr, "[g16] shape model %.4f ms\n", rank_cost(out, nvars, n) / 6.76e7);
    }
  }
}

// raw points of [lo, hi) -> temporary device buffer.  The copy is issued on the consumer's stream: a plain
// cudaMemcpy from pageable memory returns once the data is staged and is ordered only against the legacy
// default stream, which non-blocking streams do not wait for.
static void upload(void* dst, const void* src, size_t elem, size_t lo, size_t hi, int mem_kind, cudaStream_t stream) {
  size_t bytes = (hi - lo) * elem;
  if (!bytes) return;
  G16_REQUIRE(src != nullptr, "zkey view: missing point array");
  const char* s = reinterpret_cast<const char*>(src) + lo * elem;
  G16_CUDA(cudaMemcpyAsync(dst, s, bytes, mem_kind == G16_MEM_DEVICE ? cudaMemcpyDefault : cudaMemcpyHostToDevice,
                           stream));
}

// Loader validation (groth16/bn128/io.nim:228-236 loadPointG1/G2 -> curves.nim:54-91 mkG1/mkG2): every point of a
// prover array must satisfy the curve equation (G1: y^2 = x^3 + 3; G2: the twist y^2 = x^3 + 3/(9+u)) or be the
// point at infinity (0,0).  bad[0] receives the smallest offending index (0xffffffff = none).
template <class F>
struct CurveB;
template <>
struct CurveB<Fp> {
  static __device__ __forceinline__ Fp b() {          // 3 in Montgomery form
    Fp o = Fp::one();
    return fadd(fadd(o, o), o);
  }
};
template <>
struct CurveB<Fp2> {
  static __device__ __forceinline__ Fp2 b() {         // curves.nim:75-77 twistCoeffB, standard form -> Montgomery
    Fp b1, bu;
    const uint32_t w1[8] = {0x24a138e5u, 0x3267e6dcu, 0x59dbefa3u, 0xb5b4c5e5u, 0x1be06ac3u, 0x81be1899u, 0xceb8aaaeu, 0x2b149d40u};
    const uint32_t wu[8] = {0x85c315d2u, 0xe4a2bd06u, 0xe52d1852u, 0xa74fa084u, 0xeed8fdf4u, 0xcd2cafadu, 0x3af0fed4u, 0x009713b0u};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      b1.v[i] = w1[i];
      bu.v[i] = wu[i];
    }
    Fp2 r;
    r.c0 = to_mont(b1);
    r.c1 = to_mont(bu);
    return r;
  }
};
template <class P>
static __device__ __forceinline__ bool canonical(const Fe<P>& a) {     // a < modulus
#pragma unroll
  for (int k = 7; k >= 0; k--)
    if (a.v[k] != P::mod(k)) return a.v[k] < P::mod(k);
  return false;
}
static __device__ __forceinline__ bool canonical(const Fp2& a) { return canonical(a.c0) && canonical(a.c1); }

template <class F>
__global__ void __launch_bounds__(128) k_check_on_curve(const Affine<F>* __restrict__ pts, uint32_t n, uint32_t* bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = ldv(pts + i);
  if (aff_is_inf(p)) return;
  bool ok = canonical(p.x) && canonical(p.y);
  if (ok) {
    F lhs = fsqr(p.y);
    F rhs = fadd(fmul(fsqr(p.x), p.x), CurveB<F>::b());
    ok = feq(lhs, rhs);
  }
  if (!ok) atomicMin(bad, i);
}

struct TableJob {                       // validation results are read once, after all arrays have been enqueued
  const char* name;
  size_t lo;
  uint32_t* bad_dev;
};

// one prover array range -> window table (precomp) or padded plain points, validated on the way
template <class F>
static void make_table(DevBuf& table, const void* src, size_t lo, size_t hi, size_t pad_front, int c, bool precomp,
                       bool validate, const char* name, int mem_kind, DevBuf& raw, uint32_t* bad_dev,
                       std::vector<TableJob>& jobs, cudaStream_t copy_stream, cudaEvent_t copied, cudaStream_t stream) {
  const size_t n = (hi - lo) + pad_front;
  const size_t nwin = precomp ? (size_t)msm_num_windows(c) : 1;
  table.ensure(n ? nwin * n * sizeof(Affine<F>) : 16);
  if (!n) return;
  Affine<F>* pts;
  if (precomp) {
    raw.ensure((hi - lo) * sizeof(Affine<F>) + 16);
    pts = raw.as<Affine<F>>();
  } else {
    pts = table.as<Affine<F>>() + pad_front;
    if (pad_front) G16_CUDA(cudaMemsetAsync(table.p, 0, pad_front * sizeof(Affine<F>), stream));
  }
  // the copy runs on its own stream: the H2D of this array overlaps with the table kernel of the previous one
  upload(pts, src, sizeof(Affine<F>), lo, hi, mem_kind, copy_stream);
  G16_CUDA(cudaEventRecord(copied, copy_stream));
  G16_CUDA(cudaStreamWaitEvent(stream, copied, 0));
  if (validate && hi > lo) {
    G16_CUDA(cudaMemsetAsync(bad_dev, 0xff, 4, stream));
    k_check_on_curve<F><<<div_up(hi - lo, 128), 128, 0, stream>>>(pts, (uint32_t)(hi - lo), bad_dev);
    G16_LAUNCH_CHECK();
    jobs.push_back(TableJob{name, lo, bad_dev});
  }
  if (precomp) msm_build_table<F>(pts, hi - lo, pad_front, c, table.as<Affine<F>>(), stream);
}

Resident::Resident(const g16_zkey_view& zk, int shard_index_in, int shard_count_in)
    : shard_index(shard_index_in), shard_count(shard_count_in) {
  G16_REQUIRE(shard_count >= 1 && shard_index >= 0 && shard_index < shard_count, "bad shard index/count");
  G16_REQUIRE(zk.log_domain >= 1 && zk.log_domain <= 26, "domain size must be 2^1 .. 2^26 (prover.nim:101)");
  G16_REQUIRE(zk.flavour == G16_FLAVOUR_JENSGROTH || zk.flavour == G16_FLAVOUR_SNARKJS, "unknown flavour");
  G16_REQUIRE(zk.nvars >= zk.npubs + 1, "nvars must be at least npubs + 1");
  nvars = zk.nvars;
  npubs = zk.npubs;
  log_n = zk.log_domain;
  flavour = zk.flavour;
  n = (size_t)1 << log_n;
  precomp = !(zk.flags & G16_ZKEY_ONE_SHOT);
  if (const char* e = getenv("G16_LAYOUT")) precomp = e[0] != 'p';     // "plain" / "table": A/B runs
  if (precomp) {
    // The window tables are ~13x the points this context owns (5.2 GB at 2^20, 84 GB at 2^24 on one GPU).  When they
    // do not fit the device -- next to the staging copies of the upload and the per-proof workspaces -- the context
    // falls back to the plain layout (1x the key, proofs ~25 % slower) instead of failing.  G16_TABLE_BUDGET_MB
    // overrides the budget (default: 80 % of the free device memory).
    shard_plan(nvars, npubs, n, shard_index, shard_count, plan);
    const double pts_g1 = (double)(plan.a1_hi - plan.a1_lo) + (double)(plan.b1_hi - plan.b1_lo) +
                          (double)(plan.c1_hi - plan.c1_lo) + (double)(plan.h_hi - plan.h_lo);
    const double pts_g2 = (double)(plan.b2_hi - plan.b2_lo);
    const double need = (14.0 + 1.0) * (64.0 * pts_g1 + 128.0 * pts_g2)          // tables + staging copies
                        + 2.0 * 16.0 * 14.0 * ((double)nvars + (double)n);      // sorter / bucket workspaces of two slots
    size_t fr = 0, tot = 0;
    double budget = 0.0;
    if (const char* e = getenv("G16_TABLE_BUDGET_MB")) budget = atof(e) * 1048576.0;
    else if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
      // memory cached by the library's pool (freed by earlier contexts) is available to this one as well
      double cached = 0.0;
      int dev = 0;
      cudaMemPool_t pool = nullptr;
      uint64_t reserved = 0, used = 0;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
          cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
          cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
        cached = (double)(reserved - used);
      cudaGetLastError();
      budget = 0.8 * ((double)fr + cached);
    }
    if (budget > 0.0 && need > budget) precomp = false;
  }
  const bool validate = !(zk.flags & G16_ZKEY_TRUSTED);
  struct Scoped {                       // streams and events of the build, released on every exit path
    cudaStream_t main = nullptr, copy = nullptr;
    cudaEvent_t copied[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Scoped() {
      if (main) {
        cudaStreamSynchronize(main);
        cudaStreamDestroy(main);
      }
      if (copy) {
        cudaStreamSynchronize(copy);
        cudaStreamDestroy(copy);
      }
      for (auto& e : copied)
        if (e) cudaEventDestroy(e);
    }
  } sc;
  G16_CUDA(cudaStreamCreateWithFlags(&sc.main, cudaStreamNonBlocking));
  G16_CUDA(cudaStreamCreateWithFlags(&sc.copy, cudaStreamNonBlocking));
  for (auto& e : sc.copied) G16_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaStream_t main_ = sc.main, copy_ = sc.copy;
  cudaEvent_t* copied = sc.copied;
  int ncopied = 0;
  shard_plan(nvars, npubs, n, shard_index, shard_count, plan);

  auto pick_c = [&](size_t npts) {
    int c = msm_pick_window(npts, precomp);
    if (const char* e = getenv("G16_C_DELTA")) {             // experiment knob: wider / narrower windows
      c += atoi(e);
      if (c < 4) c = 4;
      if (c > 22) c = 22;
    }
    return c;
  };
  // witness-indexed pieces grouped by range (same range -> one sorter run, fused G1 launches)
  struct Piece { int which; size_t lo, hi; };
  const Piece pieces[4] = {{0, plan.a1_lo, plan.a1_hi}, {1, plan.b1_lo, plan.b1_hi}, {2, plan.c1_lo, plan.c1_hi},
                           {3, plan.b2_lo, plan.b2_hi}};
  for (const Piece& pc : pieces) {
    if (pc.hi <= pc.lo) continue;
    G16_REQUIRE(pc.hi <= nvars, "shard plan out of range");
    WitnessGroup* g = nullptr;
    for (auto& q : groups)
      if (q->lo == pc.lo && q->hi == pc.hi) g = q.get();
    if (!g) {
      groups.emplace_back(new WitnessGroup());
      g = groups.back().get();
      g->lo = pc.lo;
      g->hi = pc.hi;
      g->geom = msm_geometry(pc.hi - pc.lo, pick_c(pc.hi - pc.lo), precomp);
    }
    if (pc.which == 3) g->has_b2 = true;
    else g->which1[g->nsets1++] = pc.which;
    if (pc.which < 2) owns_ab = true;
  }
  // the group with the G2 MSM (the longest kernel) is sorted and launched first
  for (size_t i = 1; i < groups.size(); i++)
    if (groups[i]->has_b2) std::swap(groups[0], groups[i]);
  const size_t nh = plan.h_hi - plan.h_lo;
  if (nh) gh = msm_geometry(nh, pick_c(nh), precomp);

  // uploads, validation and table building; one raw staging buffer per array and the copies on their own stream, so
  // that the H2D of the next array overlaps with the table kernel of the previous one (the buffers are released at
  // the end)
  std::vector<TableJob> jobs;
  DevBuf bad;
  bad.ensure(8 * 4);
  std::vector<std::unique_ptr<DevBuf>> raws;
  int nbad = 0;
  auto raw = [&]() -> DevBuf& {
    raws.emplace_back(new DevBuf());
    return *raws.back();
  };
  static const char* names[5] = {"pointsA1", "pointsB1", "pointsC1", "pointsB2", "pointsH1"};
  for (auto& gp : groups) {
    WitnessGroup& g = *gp;
    for (int s = 0; s < g.nsets1; s++) {
      const int which = g.which1[s];
      if (which < 2) {
        make_table<Fp>(g.tab1[s], which == 0 ? zk.points_a1 : zk.points_b1, g.lo, g.hi, 0, g.geom.c, precomp, validate,
                       names[which], zk.mem_kind, raw(), bad.as<uint32_t>() + nbad++, jobs, copy_, copied[ncopied++], main_);
      } else {
        // C1[j - npubs - 1] multiplies witness[j] (prover.nim:262-264): pad so that table index == witness index
        const size_t first = (size_t)npubs + 1;
        const size_t from = g.lo > first ? g.lo : first;            // first witness index of this piece with a C point
        const size_t pad = g.hi > from ? from - g.lo : g.hi - g.lo;
        const size_t c_lo = from - first, c_hi = g.hi > from ? g.hi - first : c_lo;
        make_table<Fp>(g.tab1[s], zk.points_c1, c_lo, c_hi, pad, g.geom.c, precomp, validate, names[2], zk.mem_kind,
                       raw(), bad.as<uint32_t>() + nbad++, jobs, copy_, copied[ncopied++], main_);
      }
    }
    if (g.has_b2)
      make_table<Fp2>(g.tabB2, zk.points_b2, g.lo, g.hi, 0, g.geom.c, precomp, validate, names[3], zk.mem_kind, raw(),
                      bad.as<uint32_t>() + nbad++, jobs, copy_, copied[ncopied++], main_);
  }
  make_table<Fp>(tabH1, zk.points_h1, plan.h_lo, plan.h_hi, 0, gh.c ? gh.c : 4, precomp, validate, names[4], zk.mem_kind,
                 raw(), bad.as<uint32_t>() + nbad++, jobs, copy_, copied[ncopied++], main_);

  // witness intervals this shard reads: everything when it runs buildABC, else the union of its groups
  if (nh) witness_needs.emplace_back((size_t)0, (size_t)nvars);
  else {
    std::vector<std::pair<size_t, size_t>> iv;
    for (auto& gp : groups) iv.emplace_back(gp->lo, gp->hi);
    std::sort(iv.begin(), iv.end());
    for (auto& x : iv) {
      if (!witness_needs.empty() && x.first <= witness_needs.back().second) {
        if (x.second > witness_needs.back().second) witness_needs.back().second = x.second;
      } else witness_needs.push_back(x);
    }
  }

  // coefficient list -> CSR rows (once per zkey; only ranks that own H points run buildABC)
  if (nh) {
    size_t rec = zk.coeff_format == G16_COEFF_PACKED44_R2 ? 44 : 48;
    DevBuf& rawc = raw();
    rawc.ensure(zk.ncoeffs * rec + 16);
    if (zk.ncoeffs) {
      G16_REQUIRE(zk.coeffs != nullptr, "zkey view: missing coefficient list");
      G16_CUDA(cudaMemcpyAsync(rawc.p, zk.coeffs, zk.ncoeffs * rec,
                               zk.mem_kind == G16_MEM_DEVICE ? cudaMemcpyDefault : cudaMemcpyHostToDevice, copy_));
      G16_CUDA(cudaEventRecord(copied[ncopied], copy_));
      G16_CUDA(cudaStreamWaitEvent(main_, copied[ncopied], 0));
    }
    coeffs_to_csr(csr, rawc.p, zk.ncoeffs, (int)zk.coeff_format, (int)log_n, nvars, main_);
  }

  SpecPointsDev sp;
  memcpy(&sp.alpha1, zk.alpha1, 64);
  memcpy(&sp.beta1, zk.beta1, 64);
  memcpy(&sp.delta1, zk.delta1, 64);
  memcpy(&sp.beta2, zk.beta2, 128);
  memcpy(&sp.delta2, zk.delta2, 128);
  spec.ensure(sizeof(SpecPointsDev));
  G16_CUDA(cudaMemcpyAsync(spec.p, &sp, sizeof(sp), cudaMemcpyHostToDevice, main_));
  G16_CUDA(cudaFuncSetAttribute(k_mask_terms, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(256 * sizeof(G2XYZZ))));
  dtab1.ensure(256 * sizeof(G1XYZZ));
  dtab2.ensure(256 * sizeof(G2XYZZ));
  atab1.ensure(256 * sizeof(G1XYZZ));
  btab1.ensure(256 * sizeof(G1XYZZ));
  k_delta_tables<<<1, 128, 0, main_>>>(spec.as<SpecPointsDev>(), dtab1.as<G1XYZZ>(), dtab2.as<G2XYZZ>(),
                                       atab1.as<G1XYZZ>(), btab1.as<G1XYZZ>());
  G16_LAUNCH_CHECK();
  if (nh) ntt_prepare((int)log_n, main_);
  // validation verdicts (one small copy), then the staging buffers go away
  uint32_t bad_host[8];
  G16_CUDA(cudaMemcpyAsync(bad_host, bad.p, sizeof(bad_host), cudaMemcpyDeviceToHost, main_));
  G16_CUDA(cudaStreamSynchronize(main_));
  for (const TableJob& j : jobs) {
    const uint32_t idx = bad_host[j.bad_dev - bad.as<uint32_t>()];
    // curves.nim:95-107 mkG1 / mkG2 assert texts
    if (idx != 0xffffffffu)
      throw Error(G16_ERR_ARG, std::string(j.name) + "[" + std::to_string(j.lo + idx) + "]: " +
                                   (strcmp(j.name, "pointsB2") == 0 ? "mkG2: not a G2 curve point"
                                                                     : "mkG1: not a G1 curve point"));
  }
}


size_t Resident::bytes() const {
  size_t t = tabH1.bytes + csr.ptr.bytes + csr.other.bytes + csr.vals.bytes + dtab1.bytes + dtab2.bytes + atab1.bytes +
             btab1.bytes;
  for (auto& g : groups) t += g->tab1[0].bytes + g->tab1[1].bytes + g->tab1[2].bytes + g->tabB2.bytes;
  return t;
}

void Prover::init_slot() {
  for (int i = 0; i < 24; i++) ev_[i] = nullptr;
  if (const char* e = getenv("G16_GRAPH")) use_graph_ = atoi(e);
  // Priorities order the kernels that compete for the SMs: the witness sort and the G2 MSM (longest
  // latency-bound reduction tail) first, then the fused G1 witness MSMs (their result feeds the early
  // assembly), the H chain last -- so the single-warp tails of one MSM overlap with the accumulation of another.
  int prio_lo = 0, prio_hi = 0;
  G16_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));      // hi is numerically smaller
  int p_mid = prio_hi + (prio_lo - prio_hi) / 2;
  G16_CUDA(cudaStreamCreateWithPriority(&main_, cudaStreamNonBlocking, prio_hi));
  // the latency-bound tails (bucket-reduction ends, assembly) sit above every throughput-bound kernel: a one-block
  // kernel must not queue behind the pending blocks of another proof's accumulation (see MsmAccumulator::run)
  const int p_b2 = prio_hi < prio_lo ? prio_hi + 1 : prio_hi;
  int pr[3] = {prio_lo, p_mid > p_b2 ? p_mid : p_b2, p_b2};   // ABC+quotient+H1 | witness sort, A1+B1+C1 | B2
  if (const char* e = getenv("G16_STREAM_PRIO"))      // experiment knob: three letters of l/m/h
    for (int i = 0; i < 3 && e[i]; i++) pr[i] = e[i] == 'h' ? prio_hi : e[i] == 'm' ? p_mid : prio_lo;
  G16_CUDA(cudaStreamCreateWithPriority(&st_[0], cudaStreamNonBlocking, pr[0]));
  G16_CUDA(cudaStreamCreateWithPriority(&st_[1], cudaStreamNonBlocking, pr[1]));
  G16_CUDA(cudaStreamCreateWithPriority(&st_[2], cudaStreamNonBlocking, pr[2]));
  G16_CUDA(cudaStreamCreateWithPriority(&st_mask_, cudaStreamNonBlocking, prio_hi));
  for (int i = 0; i < 24; i++) G16_CUDA(cudaEventCreate(&ev_[i]));
  for (auto& g : R->groups) {
    (void)g;
    gw_.emplace_back(new GroupWork());
  }
  for (int i = 0; i < 6; i++) {
    G16_CUDA(cudaStreamCreateWithPriority(&tail_[i], cudaStreamNonBlocking, prio_hi));
    G16_CUDA(cudaEventCreateWithFlags(&tdone_[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < 4; i++) {
    G16_CUDA(cudaEventCreateWithFlags(&gev_[i], cudaEventDisableTiming));
    G16_CUDA(cudaEventCreateWithFlags(&gdone_[i], cudaEventDisableTiming));
    st_g_[i] = nullptr;
    if (i > 0 && i < (int)R->groups.size()) G16_CUDA(cudaStreamCreateWithPriority(&st_g_[i], cudaStreamNonBlocking, pr[1]));
  }

  witness_.ensure((size_t)R->nvars * sizeof(Fr));
  if (R->plan.h_hi > R->plan.h_lo) {                 // only ranks that own H points run buildABC and the quotient
    abc_.ensure(3 * R->n * sizeof(Fr));
    qs_.ensure(R->n * sizeof(Fr));
  }
  results_.ensure(sizeof(MsmResults));
  G16_CUDA(cudaMemset(results_.p, 0, sizeof(MsmResults)));   // all-zero XYZZ == infinity (empty shards)
  mask_.ensure(sizeof(MaskTerms));
  proof_.ensure(sizeof(ProofOut));
  G16_CUDA(cudaMemset(proof_.p, 0, sizeof(ProofOut)));
  early_.ensure(sizeof(G1XYZZ) * (1 + 512));      // partial pi_c + the two doubling tables of k_assemble_early
  G16_CUDA(cudaMallocHost(reinterpret_cast<void**>(&proof_pinned_), sizeof(ProofOut)));
}

Prover::Prover(const g16_zkey_view& zk, int shard_index, int shard_count)
    : R(std::make_shared<Resident>(zk, shard_index, shard_count)) {
  init_slot();
}

Prover::Prover(std::shared_ptr<Resident> resident) : R(std::move(resident)) { init_slot(); }

size_t Prover::resident_bytes() const {
  size_t t = R->bytes() + witness_.bytes + abc_.bytes + qs_.bytes + sortH_.workspace_bytes() + accH_.workspace_bytes();
  for (auto& g : gw_) t += g->sort.workspace_bytes() + g->acc1.workspace_bytes() + g->acc2.workspace_bytes();
  return t;
}

Prover::~Prover() {
  cudaDeviceSynchronize();
  for (auto& g : graphs_)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (int i = 0; i < 24; i++)
    if (ev_[i]) cudaEventDestroy(ev_[i]);
  for (int i = 0; i < 6; i++) {
    if (tail_[i]) cudaStreamDestroy(tail_[i]);
    if (tdone_[i]) cudaEventDestroy(tdone_[i]);
  }
  for (int i = 0; i < 4; i++) {
    if (gev_[i]) cudaEventDestroy(gev_[i]);
    if (gdone_[i]) cudaEventDestroy(gdone_[i]);
    if (st_g_[i]) cudaStreamDestroy(st_g_[i]);
  }
  for (int i = 0; i < 3; i++)
    if (st_[i]) cudaStreamDestroy(st_[i]);
  if (st_mask_) cudaStreamDestroy(st_mask_);
  if (main_) cudaStreamDestroy(main_);
  if (proof_pinned_) cudaFreeHost(proof_pinned_);
  for (int i = 0; i < 2; i++)
    if (tev_[i]) cudaEventDestroy(tev_[i]);
}

void Prover::sync() { G16_CUDA(cudaStreamSynchronize(main_)); }

void Prover::timer_start() {
  for (int i = 0; i < 2; i++)
    if (!tev_[i]) G16_CUDA(cudaEventCreate(&tev_[i]));
  G16_CUDA(cudaEventRecord(tev_[0], main_));
}
float Prover::timer_stop() {
  G16_REQUIRE(tev_[0] && tev_[1], "timer_stop without timer_start");
  G16_CUDA(cudaEventRecord(tev_[1], main_));
  G16_CUDA(cudaEventSynchronize(tev_[1]));
  float ms = 0.f;
  G16_CUDA(cudaEventElapsedTime(&ms, tev_[0], tev_[1]));
  return ms;
}

// Only the witness intervals this shard reads travel (Resident::witness_needs): everything on a rank that runs
// buildABC, else the ranges of its MSM pieces -- on 8 GPUs most ranks read an eighth to a third of the witness.
void Prover::load_witness(const void* w, int form, int mem_kind) {
  G16_REQUIRE(w != nullptr, "witness is null");
  G16_REQUIRE(form == G16_FORM_MONT || form == G16_FORM_STD, "unknown witness form");
  cudaMemcpyKind kind = mem_kind == G16_MEM_DEVICE ? cudaMemcpyDefault : cudaMemcpyHostToDevice;
  G16_CUDA(cudaEventRecord(ev_[20], main_));
  if (form == G16_FORM_MONT) staging_.ensure((size_t)R->nvars * sizeof(Fr));
  h2d_bytes_ = 0;
  for (const auto& iv : R->witness_needs) {
    const size_t cnt = iv.second - iv.first, bytes = cnt * sizeof(Fr);
    if (!cnt) continue;
    const Fr* src = reinterpret_cast<const Fr*>(w) + iv.first;
    Fr* dst = witness_.as<Fr>() + iv.first;
    if (form == G16_FORM_STD) {
      if (src != dst) G16_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, main_));
      fr_reduce_std(dst, cnt, main_);                // values >= r (malformed .wtns) are reduced like io.nim:141-145
    } else {
      Fr* stg = staging_.as<Fr>() + iv.first;
      G16_CUDA(cudaMemcpyAsync(stg, src, bytes, kind, main_));
      fr_from_mont(stg, dst, cnt, main_);
    }
    h2d_bytes_ += bytes;
  }
  G16_CUDA(cudaEventRecord(ev_[21], main_));
}

void Prover::witness_begin(int form) {
  G16_REQUIRE(form == G16_FORM_MONT || form == G16_FORM_STD, "unknown witness form");
  G16_CUDA(cudaEventRecord(ev_[20], main_));
  if (form == G16_FORM_MONT) staging_.ensure((size_t)R->nvars * sizeof(Fr));
}
void Prover::witness_upload(const void* w_host, int form, size_t lo, size_t hi) {
  if (hi > lo)
    G16_CUDA(cudaMemcpyAsync(witness_raw(form) + lo, reinterpret_cast<const Fr*>(w_host) + lo, (hi - lo) * sizeof(Fr),
                             cudaMemcpyHostToDevice, main_));
  // standard-form values are canonicalised before the slice is exposed to the peers (later reductions of the same
  // elements then never write, so a peer copy cannot observe a half-written element)
  if (hi > lo && form == G16_FORM_STD) fr_reduce_std(witness_.as<Fr>() + lo, hi - lo, main_);
  G16_CUDA(cudaEventRecord(ev_[17], main_));
}
void Prover::witness_finish(int form, size_t h2d_bytes) {
  for (const auto& iv : R->witness_needs) {
    const size_t cnt = iv.second - iv.first;
    if (!cnt) continue;
    if (form == G16_FORM_STD) fr_reduce_std(witness_.as<Fr>() + iv.first, cnt, main_);
    else fr_from_mont(staging_.as<Fr>() + iv.first, witness_.as<Fr>() + iv.first, cnt, main_);
  }
  h2d_bytes_ = h2d_bytes;
  G16_CUDA(cudaEventRecord(ev_[21], main_));
}

// The per-proof DAG (about 70 launches on up to a dozen streams) is the same every time for a given context slot and
// mode.  With G16_GRAPH=1 it is captured once -- on the slot's second proof, when every workspace exists -- and
// replayed as ONE graph launch: no per-launch host cost, no launch gaps between the small dependent kernels.
// Per-phase timing events are not available in that mode (g16_stats phases read 0).
void Prover::run_msms(g16_stats* stats) {
  (void)stats;   // phase times are read by collect_stats() once the work has completed
  // an announced mask serves ONE set of partial sums: without a new g16_ctx_set_mask the records are plain again
  if (masked_partials_ && mask_used_) {
    masked_partials_ = false;
    mask_started_ = false;
  }
  mask_used_ = true;
  const int mode = masked_partials_ ? 1 : (mask_started_ && R->shard_count == 1) ? 2 : 0;
  GraphSlot& gs = graphs_[mode];
  if (use_graph_ && runs_ >= 1 && !gs.failed) {
    // the mask terms come from a stream outside the graph: order them before it
    if (mode != 0) G16_CUDA(cudaStreamWaitEvent(main_, ev_[23], 0));
    if (!gs.exec) {
      cudaGraph_t graph = nullptr;
      const uint64_t l0 = launches_so_far();
      cudaError_t e = cudaStreamBeginCapture(main_, cudaStreamCaptureModeRelaxed);
      if (e == cudaSuccess) {
        try {
          run_msms_body(true);
        } catch (...) {
          cudaStreamEndCapture(main_, &graph);
          if (graph) cudaGraphDestroy(graph);
          cudaGetLastError();
          gs.failed = true;
          throw;
        }
        e = cudaStreamEndCapture(main_, &graph);
      }
      if (e == cudaSuccess) e = cudaGraphInstantiate(&gs.exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (e != cudaSuccess) {               // capture not possible here: fall back to plain launches for good
        cudaGetLastError();
        gs.exec = nullptr;
        gs.failed = true;
      } else {
        gs.launches = launches_so_far() - l0;
      }
    }
    if (gs.exec) {
      G16_CUDA(cudaGraphLaunch(gs.exec, main_));
      count_launches(gs.launches);
      if (mode == 2) early_done_ = true;
      runs_++;
      return;
    }
  }
  run_msms_body(false);
  runs_++;
}

void Prover::run_msms_body(bool capturing) {
  // empty shards write nothing: start every proof from infinity (all-zero XYZZ)
  G16_CUDA(cudaMemsetAsync(results_.p, 0, sizeof(MsmResults), main_));
  // ev_[0]: witness ready on main_; every worker stream waits for it
  G16_CUDA(cudaEventRecord(ev_[0], main_));
  for (int i = 0; i < 3; i++) G16_CUDA(cudaStreamWaitEvent(st_[i], ev_[0], 0));
  MsmResults* res = results_.as<MsmResults>();
  const Fr* w = witness_.as<Fr>();
  const size_t nh = R->plan.h_hi - R->plan.h_lo;

  // stream 0: ABC -> quotient -> sort of qs -> MSM over the H points   (prover.nim:245-260, 301)
  G16_CUDA(cudaEventRecord(ev_[1], st_[0]));
  if (nh) build_abc(R->csr, w, abc_.as<Fr>(), (int)R->log_n, st_[0]);      // ranks without H points skip the chain
  G16_CUDA(cudaEventRecord(ev_[2], st_[0]));
  if (nh) quotient(abc_.as<Fr>(), qs_.as<Fr>(), (int)R->log_n, (int)R->flavour, st_[0]);
  G16_CUDA(cudaEventRecord(ev_[3], st_[0]));
  if (nh) {
    sortH_.run(qs_.as<Fr>() + R->plan.h_lo, true, R->gh, st_[0]);
    MsmPointSet<Fp> hs;
    hs.points = R->tabH1.as<G1Affine>();
    hs.result = &res->h1;
    accH_.run(sortH_, &hs, 1, st_[0], tail_[0]);
  } else {
    G16_CUDA(cudaEventRecord(ev_[12], st_[0]));
    G16_CUDA(cudaStreamWaitEvent(tail_[0], ev_[12], 0));
  }
  G16_CUDA(cudaEventRecord(ev_[4], tail_[0]));

  // stream 1: per group one digit/sort pass over its witness range, then its G1 sets in the same launches
  // (prover.nim:282, 288, 302; zs = witness[npubs+1 ..] through the padded C1 table); stream 2: the G2 set
  // (prover.nim:294) over the same sorted pairs.  Every group writes distinct result fields.
  G16_CUDA(cudaEventRecord(ev_[5], st_[1]));
  for (size_t gi = 0; gi < R->groups.size(); gi++) {
    const WitnessGroup& g = *R->groups[gi];
    GroupWork& gw = *gw_[gi];
    // every group has its own stream: the latency-bound reduction tail of one group overlaps with the sort and
    // the accumulation of the next instead of delaying them
    cudaStream_t sg = gi == 0 ? st_[1] : st_g_[gi];
    if (gi > 0) G16_CUDA(cudaStreamWaitEvent(sg, ev_[0], 0));
    gw.sort.run(w + g.lo, false, g.geom, sg);
    if (gi == 0) G16_CUDA(cudaEventRecord(ev_[6], st_[1]));
    if (g.has_b2) {
      G16_CUDA(cudaEventRecord(gev_[gi], sg));
      G16_CUDA(cudaStreamWaitEvent(st_[2], gev_[gi], 0));
      G16_CUDA(cudaEventRecord(ev_[8], st_[2]));
      MsmPointSet<Fp2> bs;
      bs.points = g.tabB2.as<G2Affine>();
      bs.result = &res->b2;
      gw.acc2.run(gw.sort, &bs, 1, st_[2], tail_[1]);
      G16_CUDA(cudaEventRecord(ev_[9], tail_[1]));
    }
    if (g.nsets1) {
      MsmPointSet<Fp> ws[3];
      for (int k = 0; k < g.nsets1; k++) {
        ws[k].points = g.tab1[k].as<G1Affine>();
        ws[k].result = g.which1[k] == 0 ? &res->a1 : g.which1[k] == 1 ? &res->b1 : &res->c1;
      }
      gw.acc1.run(gw.sort, ws, g.nsets1, sg, tail_[2 + gi]);
      if (gi > 0) {                                        // join the group's tail into the tail of group 0
        G16_CUDA(cudaEventRecord(gdone_[gi], tail_[2 + gi]));
        G16_CUDA(cudaStreamWaitEvent(tail_[2], gdone_[gi], 0));
      }
    }
  }
  // tail_[2] now follows every G1 group of this shard; it also has to follow stream 1 itself when no group has
  // G1 sets (nothing was handed over to it)
  G16_CUDA(cudaEventRecord(ev_[16], st_[1]));
  G16_CUDA(cudaStreamWaitEvent(tail_[2], ev_[16], 0));
  if (R->groups.empty()) G16_CUDA(cudaEventRecord(ev_[6], st_[1]));
  bool any_b2 = false;
  for (auto& g : R->groups) any_b2 = any_b2 || g->has_b2;
  if (!any_b2) {
    G16_CUDA(cudaEventRecord(ev_[8], st_[2]));
    G16_CUDA(cudaStreamWaitEvent(tail_[1], ev_[8], 0));
    G16_CUDA(cudaEventRecord(ev_[9], tail_[1]));
  }
  G16_CUDA(cudaEventRecord(ev_[7], tail_[2]));
  if (masked_partials_) {
    // this shard's share of s ** pi_a + r ** rho, folded into its c1 partial while B2 / H are still in flight
    if (R->owns_ab) {
      if (!capturing) G16_CUDA(cudaStreamWaitEvent(tail_[2], ev_[23], 0));
      k_shard_early<<<1, 256, 0, tail_[2]>>>(results_.as<MsmResults>(), mask_.as<MaskTerms>(), early_.as<G1XYZZ>() + 1);
      G16_LAUNCH_CHECK();
    }
  } else if (mask_started_ && R->shard_count == 1) {
    // the MSM-dependent scalar multiplications start now and overlap with the B2 / H work still in flight
    if (!capturing) G16_CUDA(cudaStreamWaitEvent(tail_[2], ev_[23], 0));
    k_assemble_early<<<1, 256, 0, tail_[2]>>>(results_.as<MsmResults>(), mask_.as<MaskTerms>(), proof_.as<g16_proof>(),
                                            early_.as<G1XYZZ>(), early_.as<G1XYZZ>() + 1);
    G16_LAUNCH_CHECK();
    early_done_ = true;
  }

  for (int i = 0; i < 3; i++) {                        // H chain, B2, G1 groups (+ early assembly): join the tails
    G16_CUDA(cudaEventRecord(ev_[13 + i], tail_[i]));
    G16_CUDA(cudaStreamWaitEvent(main_, ev_[13 + i], 0));
  }
  G16_CUDA(cudaEventRecord(ev_[18], main_));
}

// valid after the main stream has been synchronised past the events of the last run_msms()
void Prover::collect_stats(g16_stats* stats) {
  if (!stats) return;
  if (use_graph_ && runs_ >= 2) {         // the phase events live inside the replayed graph: only the totals are timed
    cudaEventElapsedTime(&stats->ms_h2d, ev_[20], ev_[21]);
    cudaGetLastError();
    return;
  }
  cudaEventElapsedTime(&stats->ms_abc, ev_[1], ev_[2]);
  cudaEventElapsedTime(&stats->ms_quotient, ev_[2], ev_[3]);
  cudaEventElapsedTime(&stats->ms_msm_h, ev_[3], ev_[4]);
  cudaEventElapsedTime(&stats->ms_sort_witness, ev_[5], ev_[6]);
  cudaEventElapsedTime(&stats->ms_msm_g1_witness, ev_[6], ev_[7]);
  cudaEventElapsedTime(&stats->ms_msm_b2, ev_[8], ev_[9]);
  cudaEventElapsedTime(&stats->ms_h2d, ev_[20], ev_[21]);
}

void Prover::partials_to_affine_async(void* partials_dev) {
  k_partials_to_affine<<<1, 160, 0, main_>>>(results_.as<MsmResults>(), reinterpret_cast<PartialsAffine*>(partials_dev),
                                             masked_partials_ ? 1ull : 0ull, masked_partials_ ? mask_hash_ : 0ull);
  G16_LAUNCH_CHECK();
  G16_CUDA(cudaEventRecord(ev_[10], main_));
}
void Prover::order_stream(cudaStream_t ext, int direction) {
  if (direction == 0) G16_CUDA(cudaStreamWaitEvent(ext, ev_[10], 0));
  else {
    G16_CUDA(cudaEventRecord(ev_[11], ext));
    G16_CUDA(cudaStreamWaitEvent(main_, ev_[11], 0));
  }
}
void Prover::partials_wait(g16_stats* stats) {
  G16_CUDA(cudaEventSynchronize(ev_[10]));
  collect_stats(stats);
}
void Prover::partials_to_affine(void* partials_dev) {
  partials_to_affine_async(partials_dev);
  partials_wait(nullptr);
}

void Prover::sum_partials(const void* gathered_dev, int count) {
  G16_REQUIRE(count >= 1, "need at least one partial record");
  k_sum_partials<<<1, 160, 0, main_>>>(reinterpret_cast<const PartialsAffine*>(gathered_dev), count,
                                       results_.as<MsmResults>(), masked_partials_ ? 1ull : 0ull,
                                       masked_partials_ ? mask_hash_ : 0ull, &proof_.as<ProofOut>()->status);
  G16_LAUNCH_CHECK();
}

// r*delta1, s*delta1, s*delta2, -rs*delta1 need no MSM result: started early on their own stream
void Prover::start_mask(const uint64_t r[4], const uint64_t s[4]) {
  G16_REQUIRE(r != nullptr && s != nullptr, "mask is null");
  MaskTerms* m = mask_.as<MaskTerms>();
  uint32_t rs[16 + 20 + 4];                            // r, s, glv[2][10], glv_ok + padding: contiguous in MaskTerms
  memcpy(rs, r, 32);
  memcpy(rs + 8, s, 32);
  bool ok_r = false, ok_s = false;
  const GlvSplit gr = glv_decompose(r, &ok_r), gs = glv_decompose(s, &ok_s);
  const GlvSplit* sp[2] = {&gr, &gs};
  for (int i = 0; i < 2; i++) {
    uint32_t* g = rs + 16 + 10 * i;
    memcpy(g, sp[i]->k1, 16);
    memcpy(g + 4, sp[i]->k2, 16);
    g[8] = sp[i]->neg1;
    g[9] = sp[i]->neg2;
  }
  static int glv_on = -1;
  if (glv_on < 0) {
    const char* e = getenv("G16_GLV");                 // A/B knob: 0 = plain 254-step doubling chains
    glv_on = (e && e[0] == '0') ? 0 : 1;
  }
  rs[36] = (glv_on && ok_r && ok_s) ? 1u : 0u;
  rs[37] = rs[38] = rs[39] = 0;
  static_assert(offsetof(MaskTerms, glv) == offsetof(MaskTerms, r) + 64 &&
                    offsetof(MaskTerms, glv_ok) == offsetof(MaskTerms, r) + 64 + 80,
                "MaskTerms: r, s, glv, glv_ok must be contiguous");
  G16_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(m) + offsetof(MaskTerms, r), rs, sizeof(rs), cudaMemcpyHostToDevice,
                           st_mask_));
  k_mask_terms<<<6, 256, 256 * sizeof(G2XYZZ), st_mask_>>>(R->spec.as<SpecPointsDev>(), R->dtab1.as<G1XYZZ>(),
                                                           R->dtab2.as<G2XYZZ>(), R->atab1.as<G1XYZZ>(),
                                                           R->btab1.as<G1XYZZ>(), m);
  G16_LAUNCH_CHECK();
  G16_CUDA(cudaEventRecord(ev_[23], st_mask_));
  mask_started_ = true;
  early_done_ = false;
  masked_partials_ = false;
}

void Prover::set_mask(const uint64_t r[4], const uint64_t s[4]) {
  start_mask(r, s);
  memcpy(mask_host_, r, 32);
  memcpy(mask_host_ + 4, s, 32);
  uint64_t h = 1469598103934665603ull;               // FNV-1a over the 64 mask bytes: the tag of masked records
  const unsigned char* b = reinterpret_cast<const unsigned char*>(mask_host_);
  for (int i = 0; i < 64; i++) h = (h ^ b[i]) * 1099511628211ull;
  mask_hash_ = h;
  masked_partials_ = true;
  mask_used_ = false;
}
bool Prover::same_mask(const uint64_t r[4], const uint64_t s[4]) const {
  return r && s && memcmp(mask_host_, r, 32) == 0 && memcmp(mask_host_ + 4, s, 32) == 0;
}

// enqueue the rest of the proof (assembly + 256-byte copy to pinned memory); returns without waiting
void Prover::finish_async() {
  G16_REQUIRE(mask_started_, "finish without start_mask");
  MaskTerms* m = mask_.as<MaskTerms>();
  G16_CUDA(cudaEventRecord(ev_[19], main_));
  G16_CUDA(cudaStreamWaitEvent(main_, ev_[23], 0));
  if (masked_partials_) {
    k_assemble_final_masked<<<1, 96, 0, main_>>>(results_.as<MsmResults>(), m, proof_.as<g16_proof>());
    G16_LAUNCH_CHECK();
  } else {
  if (!early_done_) {                                  // multi-GPU path: the sums arrive only now
    k_assemble_early<<<1, 256, 0, main_>>>(results_.as<MsmResults>(), m, proof_.as<g16_proof>(), early_.as<G1XYZZ>(),
                                           early_.as<G1XYZZ>() + 1);
    G16_LAUNCH_CHECK();
  }
  k_assemble_final<<<1, 64, 0, main_>>>(results_.as<MsmResults>(), m, early_.as<G1XYZZ>(), proof_.as<g16_proof>());
  G16_LAUNCH_CHECK();
  }
  mask_started_ = false;
  early_done_ = false;
  masked_partials_ = false;
  G16_CUDA(cudaMemcpyAsync(proof_pinned_, proof_.p, sizeof(ProofOut), cudaMemcpyDeviceToHost, main_));
  G16_CUDA(cudaEventRecord(ev_[22], main_));
  in_flight_ = true;
}

// wait for the proof enqueued by finish_async()
void Prover::wait(g16_proof* proof, g16_stats* stats) {
  G16_REQUIRE(proof != nullptr, "proof output is null");
  G16_REQUIRE(in_flight_, "no proof in flight on this context");
  G16_CUDA(cudaEventSynchronize(ev_[22]));
  in_flight_ = false;
  G16_REQUIRE(proof_pinned_->status == 0,
              "partial records disagree on the mask: g16_ctx_set_mask must be called by every rank of a proof with "
              "the same r, s, or by none");
  memcpy(proof, &proof_pinned_->proof, sizeof(g16_proof));
  if (stats) {
    collect_stats(stats);
    cudaEventElapsedTime(&stats->ms_assemble, ev_[19], ev_[22]);
    cudaEventElapsedTime(&stats->ms_total, ev_[20], ev_[22]);
  }
}

void Prover::finish(g16_proof* proof, g16_stats* stats) {
  finish_async();
  wait(proof, stats);
}

}  // namespace g16
