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
#include <stdlib.h>
#include "ntt.cuh"

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
__global__ void __launch_bounds__(256) k_assemble_early(const MsmResults* res, const MaskTerms* m, g16_proof* proof,
                                                        G1XYZZ* partial_c, G1XYZZ* scratch) {
  __shared__ G1XYZZ red[256];
  const uint32_t h = threadIdx.x >> 7, j = threadIdx.x & 127u;
  G1XYZZ* tbl = scratch + h * 256;
  if (j == 0) {
    G1XYZZ t;
    if (h == 0) xyzz_add_ni(t, ldv(&m->t_a), ldv(&res->a1));        // pi_a                 prover.nim:282
    else xyzz_add_ni(t, ldv(&m->t_b1), ldv(&res->b1));              // rho                  prover.nim:288
    G1Affine pa;
    xyzz_to_affine_ni(pa, t);
    if (h == 0) stv(reinterpret_cast<G1Affine*>(proof->pi_a), pa);
    G1XYZZ p = xyzz_from_affine(pa);
    for (int i = 0; i < 256; i++) {
      stv(tbl + i, p);
      if (i < 255) xyzz_dbl_ni(p, p);
    }
  }
  __syncthreads();
  const uint32_t* k = h == 0 ? m->s : m->r;                         // s ** pi_a, r ** rho  prover.nim:298,299
  G1XYZZ acc = xyzz_inf<Fp>();
  if ((k[j >> 5] >> (j & 31)) & 1u) acc = ldv(tbl + j);
  const uint32_t j2 = j + 128;
  if ((k[j2 >> 5] >> (j2 & 31)) & 1u) xyzz_add_ni(acc, acc, ldv(tbl + j2));
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
  G1XYZZ* tbl = scratch + h * 256;
  if (j == 0) {
    G1XYZZ p = ldv(h == 0 ? &res->a1 : &res->b1);
    for (int i = 0; i < 256; i++) {
      stv(tbl + i, p);
      if (i < 255) xyzz_dbl_ni(p, p);
    }
  }
  __syncthreads();
  const uint32_t* k = h == 0 ? m->s : m->r;
  G1XYZZ acc = xyzz_inf<Fp>();
  if ((k[j >> 5] >> (j & 31)) & 1u) acc = ldv(tbl + j);
  const uint32_t j2 = j + 128;
  if ((k[j2 >> 5] >> (j2 & 31)) & 1u) xyzz_add_ni(acc, acc, ldv(tbl + j2));
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
__global__ void __launch_bounds__(160) k_partials_to_affine(const MsmResults* res, PartialsAffine* out) {
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
__global__ void __launch_bounds__(160) k_sum_partials(const PartialsAffine* parts, int count, MsmResults* res) {
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
static void shard_range(size_t N, int k, int G, size_t& lo, size_t& hi) {   // msm.nim:107-111
  lo = (N * (size_t)k) / (size_t)G;
  hi = (k == G - 1) ? N : (N * (size_t)(k + 1)) / (size_t)G;
}

// Point ranges of rank k of G: out = {v_lo, v_hi, h_lo, h_hi} (witness-indexed arrays A1/B1/C1/B2, H array).
// "uniform" is the reference's chunking of every MSM (msm.nim:107-111).  The default "hgroup" keeps contiguous
// ranges but gives the H array -- and with it buildABC and the quotient, which cannot be sharded -- to the
// first m ranks only, and compensates them with a smaller share of the witness arrays: the other ranks skip the
// quotient altogether and every rank reduces fewer bucket sets.  Shares follow the measured costs at 2^20
// (witness MSMs 13.0, H MSM 2.05, buildABC + quotient 1.4 ms).  Every rank evaluates the same formula.
void shard_ranges(size_t nvars, size_t n, int k, int G, size_t out[4]) {
  static int uniform = -1;
  if (uniform < 0) {
    const char* e = getenv("G16_SHARD_POLICY");
    uniform = (e && e[0] == 'u') ? 1 : 0;
  }
  if (G < 4 || uniform) {                      // measured: 13.1 vs 12.9 ms at G = 2, 8.1 vs 8.6 at 4, 5.5 vs 5.9 at 8
    shard_range(nvars, k, G, out[0], out[1]);
    shard_range(n, k, G, out[2], out[3]);
    return;
  }
  const int m = G >= 8 ? G / 4 : 1;
  const double cw = 13.0, ch = 2.05, cq = 1.4;
  const double d = (cq + ch / m) / cw;
  double fh = (1.0 - (G - m) * d) / G;
  if (fh < 0) fh = 0;
  const double fo = (1.0 - m * fh) / (G - m);
  auto bound = [&](int idx) -> size_t {
    if (idx >= G) return nvars;
    double f = idx <= m ? idx * fh : m * fh + (idx - m) * fo;
    size_t b = (size_t)(f * (double)nvars + 0.5);
    return b > nvars ? nvars : b;
  };
  out[0] = bound(k);
  out[1] = bound(k + 1);
  if (out[1] < out[0]) out[1] = out[0];
  if (k < m) shard_range(n, k, m, out[2], out[3]);
  else out[2] = out[3] = n;
}

// raw points of [lo, hi) -> temporary device buffer.  The copy is issued on the consumer's stream: a plain
// cudaMemcpy from pageable memory returns once the data is staged and is ordered only against the legacy
// default stream, which non-blocking streams do not wait for.
static void upload(DevBuf& dst, const void* src, size_t elem, size_t lo, size_t hi, int mem_kind,
                   cudaStream_t stream) {
  size_t bytes = (hi - lo) * elem;
  dst.ensure(bytes ? bytes : 16);
  if (!bytes) return;
  G16_REQUIRE(src != nullptr, "zkey view: missing point array");
  const char* s = reinterpret_cast<const char*>(src) + lo * elem;
  G16_CUDA(cudaMemcpyAsync(dst.p, s, bytes,
                           mem_kind == G16_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
}

template <class F>
static void make_table(DevBuf& table, const void* src, size_t lo, size_t hi, size_t pad_front, int c, int mem_kind,
                       cudaStream_t stream) {
  size_t n = (hi - lo) + pad_front;
  table.ensure(n ? (size_t)msm_num_windows(c) * n * sizeof(Affine<F>) : 16);
  if (!n) return;
  DevBuf raw;
  upload(raw, src, sizeof(Affine<F>), lo, hi, mem_kind, stream);
  msm_build_table<F>(raw.as<Affine<F>>(), hi - lo, pad_front, c, table.as<Affine<F>>(), stream);
  G16_CUDA(cudaStreamSynchronize(stream));   // raw is released on return
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
  cudaStream_t main_ = nullptr;
  G16_CUDA(cudaStreamCreateWithFlags(&main_, cudaStreamNonBlocking));
  // this context's contiguous ranges (msm.nim:107-111): witness-indexed arrays and the H array
  {
    size_t rg[4];
    shard_ranges(nvars, n, shard_index, shard_count, rg);
    v_lo = rg[0];
    v_hi = rg[1];
    h_lo = rg[2];
    h_hi = rg[3];
  }
  const size_t nv = v_hi - v_lo, nh = h_hi - h_lo;
  if (nv) gw = msm_geometry(nv, msm_pick_window(nv, true), true);
  if (nh) gh = msm_geometry(nh, msm_pick_window(nh, true), true);
  make_table<Fp>(tabA1, zk.points_a1, v_lo, v_hi, 0, gw.c ? gw.c : 4, zk.mem_kind, main_);
  make_table<Fp>(tabB1, zk.points_b1, v_lo, v_hi, 0, gw.c ? gw.c : 4, zk.mem_kind, main_);
  make_table<Fp2>(tabB2, zk.points_b2, v_lo, v_hi, 0, gw.c ? gw.c : 4, zk.mem_kind, main_);
  {
    // C1[j - npubs - 1] multiplies witness[j] (prover.nim:262-264): pad so that table index == witness index
    size_t first = (size_t)npubs + 1;
    size_t from = v_lo > first ? v_lo : first;            // first witness index of this shard with a C point
    size_t pad = v_hi > from ? from - v_lo : nv;
    size_t c_lo = from - first, c_hi = v_hi > from ? v_hi - first : c_lo;
    make_table<Fp>(tabC1, zk.points_c1, c_lo, c_hi, pad, gw.c ? gw.c : 4, zk.mem_kind, main_);
  }
  make_table<Fp>(tabH1, zk.points_h1, h_lo, h_hi, 0, gh.c ? gh.c : 4, zk.mem_kind, main_);

  // coefficient list -> CSR rows (once per zkey)
  {
    size_t rec = zk.coeff_format == G16_COEFF_PACKED44_R2 ? 44 : 48;
    DevBuf raw;
    raw.ensure(zk.ncoeffs * rec + 16);
    if (zk.ncoeffs) {
      G16_REQUIRE(zk.coeffs != nullptr, "zkey view: missing coefficient list");
      G16_CUDA(cudaMemcpyAsync(raw.p, zk.coeffs, zk.ncoeffs * rec,
                               zk.mem_kind == G16_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               main_));
    }
    coeffs_to_csr(csr, raw.p, zk.ncoeffs, (int)zk.coeff_format, (int)log_n, nvars, main_);
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
  G16_CUDA(cudaStreamSynchronize(main_));

  ntt_prepare((int)log_n, main_);
  G16_CUDA(cudaStreamSynchronize(main_));
  cudaStreamDestroy(main_);
}


size_t Resident::bytes() const {
  return tabA1.bytes + tabB1.bytes + tabC1.bytes + tabH1.bytes + tabB2.bytes + csr.ptr.bytes + csr.other.bytes +
         csr.vals.bytes + dtab1.bytes + dtab2.bytes + atab1.bytes + btab1.bytes;
}

void Prover::init_slot() {
  for (int i = 0; i < 24; i++) ev_[i] = nullptr;
  // Priorities order the kernels that compete for the SMs: the witness sort and the G2 MSM (longest
  // latency-bound reduction tail) first, then the fused G1 witness MSMs (their result feeds the early
  // assembly), the H chain last -- so the single-warp tails of one MSM overlap with the accumulation of another.
  int prio_lo = 0, prio_hi = 0;
  G16_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));      // hi is numerically smaller
  int p_mid = prio_hi + (prio_lo - prio_hi) / 2;
  G16_CUDA(cudaStreamCreateWithPriority(&main_, cudaStreamNonBlocking, prio_hi));
  int pr[3] = {prio_lo, p_mid, prio_hi};              // ABC+quotient+H1 | witness sort, A1+B1+C1 | B2
  if (const char* e = getenv("G16_STREAM_PRIO"))      // experiment knob: three letters of l/m/h
    for (int i = 0; i < 3 && e[i]; i++) pr[i] = e[i] == 'h' ? prio_hi : e[i] == 'm' ? p_mid : prio_lo;
  G16_CUDA(cudaStreamCreateWithPriority(&st_[0], cudaStreamNonBlocking, pr[0]));
  G16_CUDA(cudaStreamCreateWithPriority(&st_[1], cudaStreamNonBlocking, pr[1]));
  G16_CUDA(cudaStreamCreateWithPriority(&st_[2], cudaStreamNonBlocking, pr[2]));
  G16_CUDA(cudaStreamCreateWithPriority(&st_mask_, cudaStreamNonBlocking, prio_hi));
  for (int i = 0; i < 24; i++) G16_CUDA(cudaEventCreate(&ev_[i]));

  witness_.ensure((size_t)R->nvars * sizeof(Fr));
  abc_.ensure(3 * R->n * sizeof(Fr));
  qs_.ensure(R->n * sizeof(Fr));
  results_.ensure(sizeof(MsmResults));
  G16_CUDA(cudaMemset(results_.p, 0, sizeof(MsmResults)));   // all-zero XYZZ == infinity (empty shards)
  mask_.ensure(sizeof(MaskTerms));
  proof_.ensure(sizeof(g16_proof));
  early_.ensure(sizeof(G1XYZZ) * (1 + 512));      // partial pi_c + the two doubling tables of k_assemble_early
  G16_CUDA(cudaMallocHost(reinterpret_cast<void**>(&proof_pinned_), sizeof(g16_proof)));
}

Prover::Prover(const g16_zkey_view& zk, int shard_index, int shard_count)
    : R(std::make_shared<Resident>(zk, shard_index, shard_count)) {
  init_slot();
}

Prover::Prover(std::shared_ptr<Resident> resident) : R(std::move(resident)) { init_slot(); }

size_t Prover::resident_bytes() const {
  return R->bytes() + witness_.bytes + abc_.bytes + qs_.bytes + sortW_.workspace_bytes() +
         sortH_.workspace_bytes() + accW_.workspace_bytes() + accH_.workspace_bytes() + accB2_.workspace_bytes();
}

Prover::~Prover() {
  cudaDeviceSynchronize();
  for (int i = 0; i < 24; i++)
    if (ev_[i]) cudaEventDestroy(ev_[i]);
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

void Prover::load_witness(const void* w, int form, int mem_kind) {
  G16_REQUIRE(w != nullptr, "witness is null");
  G16_REQUIRE(form == G16_FORM_MONT || form == G16_FORM_STD, "unknown witness form");
  size_t bytes = (size_t)R->nvars * sizeof(Fr);
  cudaMemcpyKind kind = mem_kind == G16_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  G16_CUDA(cudaEventRecord(ev_[20], main_));
  if (form == G16_FORM_STD) {
    G16_CUDA(cudaMemcpyAsync(witness_.p, w, bytes, kind, main_));
  } else {
    staging_.ensure(bytes);
    G16_CUDA(cudaMemcpyAsync(staging_.p, w, bytes, kind, main_));
    fr_from_mont(staging_.as<Fr>(), witness_.as<Fr>(), R->nvars, main_);
  }
  G16_CUDA(cudaEventRecord(ev_[21], main_));
}

void Prover::run_msms(g16_stats* stats) {
  // empty shards write nothing: start every proof from infinity (all-zero XYZZ)
  G16_CUDA(cudaMemsetAsync(results_.p, 0, sizeof(MsmResults), main_));
  // ev_[0]: witness ready on main_; every worker stream waits for it
  G16_CUDA(cudaEventRecord(ev_[0], main_));
  for (int i = 0; i < 3; i++) G16_CUDA(cudaStreamWaitEvent(st_[i], ev_[0], 0));
  MsmResults* res = results_.as<MsmResults>();
  const Fr* w = witness_.as<Fr>();
  const size_t nv = R->v_hi - R->v_lo, nh = R->h_hi - R->h_lo;

  // stream 0: ABC -> quotient -> sort of qs -> MSM over the H table   (prover.nim:245-260, 301)
  G16_CUDA(cudaEventRecord(ev_[1], st_[0]));
  if (nh) build_abc(R->csr, w, abc_.as<Fr>(), (int)R->log_n, st_[0]);      // ranks without H points skip the chain
  G16_CUDA(cudaEventRecord(ev_[2], st_[0]));
  if (nh) quotient(abc_.as<Fr>(), qs_.as<Fr>(), (int)R->log_n, (int)R->flavour, st_[0]);
  G16_CUDA(cudaEventRecord(ev_[3], st_[0]));
  if (nh) {
    sortH_.run(qs_.as<Fr>() + R->h_lo, true, R->gh, st_[0]);
    MsmPointSet<Fp> hs;
    hs.points = R->tabH1.as<G1Affine>();
    hs.result = &res->h1;
    accH_.run(sortH_, &hs, 1, st_[0]);
  }
  G16_CUDA(cudaEventRecord(ev_[4], st_[0]));

  // stream 1: one digit/sort pass over the witness, then A1, B1, C1 in the same launches
  // (prover.nim:282, 288, 302; zs = witness[npubs+1 ..] through the padded C1 table)
  G16_CUDA(cudaEventRecord(ev_[5], st_[1]));
  if (nv) sortW_.run(w + R->v_lo, false, R->gw, st_[1]);
  G16_CUDA(cudaEventRecord(ev_[6], st_[1]));
  G16_CUDA(cudaStreamWaitEvent(st_[2], ev_[6], 0));
  if (nv) {
    MsmPointSet<Fp> ws[3];
    ws[0].points = R->tabA1.as<G1Affine>();
    ws[0].result = &res->a1;
    ws[1].points = R->tabB1.as<G1Affine>();
    ws[1].result = &res->b1;
    ws[2].points = R->tabC1.as<G1Affine>();
    ws[2].result = &res->c1;
    accW_.run(sortW_, ws, 3, st_[1]);
  }
  G16_CUDA(cudaEventRecord(ev_[7], st_[1]));
  if (masked_partials_) {
    // this shard's share of s ** pi_a + r ** rho, folded into its c1 partial while B2 / H are still in flight
    G16_CUDA(cudaStreamWaitEvent(st_[1], ev_[23], 0));
    k_shard_early<<<1, 256, 0, st_[1]>>>(results_.as<MsmResults>(), mask_.as<MaskTerms>(), early_.as<G1XYZZ>() + 1);
    G16_LAUNCH_CHECK();
  } else if (mask_started_ && R->shard_count == 1) {
    // the MSM-dependent scalar multiplications start now and overlap with the B2 / H work still in flight
    G16_CUDA(cudaStreamWaitEvent(st_[1], ev_[23], 0));
    k_assemble_early<<<1, 256, 0, st_[1]>>>(results_.as<MsmResults>(), mask_.as<MaskTerms>(), proof_.as<g16_proof>(),
                                            early_.as<G1XYZZ>(), early_.as<G1XYZZ>() + 1);
    G16_LAUNCH_CHECK();
    early_done_ = true;
  }

  // stream 2: pi_B MSM in G2 over the same sorted pairs (prover.nim:294)
  G16_CUDA(cudaEventRecord(ev_[8], st_[2]));
  if (nv) {
    MsmPointSet<Fp2> bs;
    bs.points = R->tabB2.as<G2Affine>();
    bs.result = &res->b2;
    accB2_.run(sortW_, &bs, 1, st_[2]);
  }
  G16_CUDA(cudaEventRecord(ev_[9], st_[2]));

  for (int i = 0; i < 3; i++) {
    G16_CUDA(cudaEventRecord(ev_[13 + i], st_[i]));
    G16_CUDA(cudaStreamWaitEvent(main_, ev_[13 + i], 0));
  }
  G16_CUDA(cudaEventRecord(ev_[18], main_));
  (void)stats;   // phase times are read by collect_stats() once the work has completed
}

// valid after the main stream has been synchronised past the events of the last run_msms()
void Prover::collect_stats(g16_stats* stats) {
  if (!stats) return;
  cudaEventElapsedTime(&stats->ms_abc, ev_[1], ev_[2]);
  cudaEventElapsedTime(&stats->ms_quotient, ev_[2], ev_[3]);
  cudaEventElapsedTime(&stats->ms_msm_h, ev_[3], ev_[4]);
  cudaEventElapsedTime(&stats->ms_sort_witness, ev_[5], ev_[6]);
  cudaEventElapsedTime(&stats->ms_msm_g1_witness, ev_[6], ev_[7]);
  cudaEventElapsedTime(&stats->ms_msm_b2, ev_[8], ev_[9]);
  cudaEventElapsedTime(&stats->ms_h2d, ev_[20], ev_[21]);
}

void Prover::partials_to_affine_async(void* partials_dev) {
  k_partials_to_affine<<<1, 160, 0, main_>>>(results_.as<MsmResults>(), reinterpret_cast<PartialsAffine*>(partials_dev));
  G16_LAUNCH_CHECK();
  G16_CUDA(cudaEventRecord(ev_[10], main_));
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
                                       results_.as<MsmResults>());
  G16_LAUNCH_CHECK();
}

// r*delta1, s*delta1, s*delta2, -rs*delta1 need no MSM result: started early on their own stream
void Prover::start_mask(const uint64_t r[4], const uint64_t s[4]) {
  G16_REQUIRE(r != nullptr && s != nullptr, "mask is null");
  MaskTerms* m = mask_.as<MaskTerms>();
  uint32_t rs[16];
  memcpy(rs, r, 32);
  memcpy(rs + 8, s, 32);
  G16_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(m) + offsetof(MaskTerms, r), rs, 64, cudaMemcpyHostToDevice,
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
  masked_partials_ = true;
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
  G16_CUDA(cudaMemcpyAsync(proof_pinned_, proof_.p, sizeof(g16_proof), cudaMemcpyDeviceToHost, main_));
  G16_CUDA(cudaEventRecord(ev_[22], main_));
  in_flight_ = true;
}

// wait for the proof enqueued by finish_async()
void Prover::wait(g16_proof* proof, g16_stats* stats) {
  G16_REQUIRE(proof != nullptr, "proof output is null");
  G16_REQUIRE(in_flight_, "no proof in flight on this context");
  G16_CUDA(cudaEventSynchronize(ev_[22]));
  in_flight_ = false;
  memcpy(proof, proof_pinned_, sizeof(g16_proof));
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
