// CPU oracle, compiled part -- TEST INFRASTRUCTURE AND CPU BASELINE ONLY (never linked into the product).
//
// A C++ restatement of the reference's CPU prover *with the reference's own decomposition*, because the
// reference itself (Nim + un-vendored mratsim/constantine @5f7ba18f) cannot be built here:
//   - fields: 4 x u64 Montgomery, R = 2^256 (constantine's representation; io.nim:87-92)
//   - MSM G1: per-thread contiguous chunks, bucket-method Pippenger per chunk, chunk results converted to
//     affine and summed sequentially (msm.nim:35-59, 89-124)
//   - MSM G2: same threading; plain unsigned-window bucket method ("reference" variant, msm.nim:74-76)
//   - NTT: literal recursive radix-2 workers, natural order, 1/2 folded into the inverse (ntt.nim:17-161)
//   - quotient: three tasks (A, B, C chains), single-threaded pointwise (prover.nim:96-181)
//   - buildABC: single-threaded scatter-add (prover.nim:56-73)
//   - proof assembly (prover.nim:278-304)
// PARITY UNPINNED by reference outputs (see oracle/g16_oracle.py header); pinned against the Python oracle
// and the golden vectors by tests/test_oracle_cpu.py.
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;

// ------------------------------------------------------------------------------------------- fields
struct ModP {
  static const u64 M[4];
  static const u64 INV;
  static const u64 ONE[4];
  static const u64 R2[4];
};
struct ModR {
  static const u64 M[4];
  static const u64 INV;
  static const u64 ONE[4];
  static const u64 R2[4];
};
const u64 ModP::M[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
const u64 ModP::INV = 0x87d20782e4866389ull;
const u64 ModP::ONE[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
const u64 ModP::R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
const u64 ModR::M[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
const u64 ModR::INV = 0xc2e1f593efffffffull;
const u64 ModR::ONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
const u64 ModR::R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};

template <class Q>
struct F {
  u64 v[4];
  static F zero() { F r; memset(r.v, 0, 32); return r; }
  static F one() { F r; memcpy(r.v, Q::ONE, 32); return r; }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  bool operator==(const F& o) const { return memcmp(v, o.v, 32) == 0; }
};

template <class Q>
static inline bool geq_mod(const u64* a) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > Q::M[i]) return true;
    if (a[i] < Q::M[i]) return false;
  }
  return true;
}
template <class Q>
static inline void sub_mod(u64* a) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - Q::M[i] - (u64)b;
    a[i] = (u64)t;
    b = (t >> 64) & 1;
  }
}
template <class Q>
static inline F<Q> add(const F<Q>& a, const F<Q>& b) {
  F<Q> r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.v[i] + b.v[i];
    r.v[i] = (u64)c;
    c >>= 64;
  }
  if (geq_mod<Q>(r.v)) sub_mod<Q>(r.v);
  return r;
}
template <class Q>
static inline F<Q> sub(const F<Q>& a, const F<Q>& b) {
  F<Q> r;
  u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a.v[i] - b.v[i] - (u64)bw;
    r.v[i] = (u64)t;
    bw = (t >> 64) & 1;
  }
  if (bw) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)r.v[i] + Q::M[i];
      r.v[i] = (u64)c;
      c >>= 64;
    }
  }
  return r;
}
template <class Q>
static inline F<Q> neg(const F<Q>& a) { return a.is_zero() ? a : sub(F<Q>::zero(), a); }
template <class Q>
static inline F<Q> dbl(const F<Q>& a) { return add(a, a); }

// CIOS Montgomery multiplication on 64-bit limbs
template <class Q>
static inline F<Q> mul(const F<Q>& a, const F<Q>& b) {
  u64 t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.v[j] * b.v[i] + t[j];
      t[j] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (u64)c;
    t[5] = (u64)(c >> 64);
    u64 m = t[0] * Q::INV;
    c = (u128)m * Q::M[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * Q::M[j] + t[j];
      t[j - 1] = (u64)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (u64)c;
    t[4] = t[5] + (u64)(c >> 64);
  }
  F<Q> r;
  memcpy(r.v, t, 32);
  if (t[4] || geq_mod<Q>(r.v)) sub_mod<Q>(r.v);
  return r;
}
template <class Q>
static inline F<Q> sqr(const F<Q>& a) { return mul(a, a); }
template <class Q>
static F<Q> pow_limbs(const F<Q>& a, const u64 e[4]) {
  F<Q> acc = F<Q>::one();
  for (int i = 255; i >= 0; i--) {
    acc = sqr(acc);
    if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
  }
  return acc;
}
template <class Q>
static F<Q> inv(const F<Q>& a) {
  u64 e[4];
  memcpy(e, Q::M, 32);
  e[0] -= 2;
  return pow_limbs(a, e);
}
template <class Q>
static F<Q> to_mont(const F<Q>& a) { F<Q> r2; memcpy(r2.v, Q::R2, 32); return mul(a, r2); }
template <class Q>
static F<Q> from_mont(const F<Q>& a) { F<Q> o = F<Q>::zero(); o.v[0] = 1; return mul(a, o); }

typedef F<ModP> Fp;
typedef F<ModR> Fr;

struct Fp2 {
  Fp c0, c1;
  static Fp2 zero() { Fp2 r; r.c0 = Fp::zero(); r.c1 = Fp::zero(); return r; }
  static Fp2 one() { Fp2 r; r.c0 = Fp::one(); r.c1 = Fp::zero(); return r; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
};
static inline Fp2 add(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = add(a.c0, b.c0); r.c1 = add(a.c1, b.c1); return r; }
static inline Fp2 sub(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = sub(a.c0, b.c0); r.c1 = sub(a.c1, b.c1); return r; }
static inline Fp2 neg(const Fp2& a) { Fp2 r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
static inline Fp2 dbl(const Fp2& a) { return add(a, a); }
static inline Fp2 mul(const Fp2& a, const Fp2& b) {   // schoolbook: 4 Fp multiplications (u^2 = -1)
  Fp2 r;
  r.c0 = sub(mul(a.c0, b.c0), mul(a.c1, b.c1));
  r.c1 = add(mul(a.c0, b.c1), mul(a.c1, b.c0));
  return r;
}
static inline Fp2 sqr(const Fp2& a) { return mul(a, a); }
static Fp2 inv(const Fp2& a) {
  Fp d = inv(add(sqr(a.c0), sqr(a.c1)));
  Fp2 r;
  r.c0 = mul(a.c0, d);
  r.c1 = neg(mul(a.c1, d));
  return r;
}

// ------------------------------------------------------------------------------------------- curves
// Jacobian coordinates (X/Z^2, Y/Z^3); infinity <=> Z = 0.  Affine infinity = (0,0) (curves.nim:49-50).
template <class T>
struct Aff { T x, y; bool is_inf() const { return x.is_zero() && y.is_zero(); } };
template <class T>
struct Jac { T x, y, z; bool is_inf() const { return z.is_zero(); } };

template <class T>
static Jac<T> jac_inf() { Jac<T> r; r.x = T::zero(); r.y = T::one(); r.z = T::zero(); return r; }

template <class T>
static Jac<T> jac_dbl(const Jac<T>& p) {        // dbl-2009-l, a = 0
  if (p.is_inf() || p.y.is_zero()) return jac_inf<T>();
  T A = sqr(p.x), B = sqr(p.y), C = sqr(B);
  T t = add(p.x, B);
  T D = dbl(sub(sub(sqr(t), A), C));
  T E = add(dbl(A), A);
  T Fq = sqr(E);
  Jac<T> r;
  r.x = sub(Fq, dbl(D));
  T c8 = dbl(dbl(dbl(C)));
  r.y = sub(mul(E, sub(D, r.x)), c8);
  r.z = dbl(mul(p.y, p.z));
  return r;
}
template <class T>
static Jac<T> jac_madd(const Jac<T>& p, const Aff<T>& q) {   // mixed addition with all special cases
  if (q.is_inf()) return p;
  if (p.is_inf()) { Jac<T> r; r.x = q.x; r.y = q.y; r.z = T::one(); return r; }
  T Z1Z1 = sqr(p.z);
  T U2 = mul(q.x, Z1Z1);
  T S2 = mul(mul(q.y, p.z), Z1Z1);
  T H = sub(U2, p.x), rr = sub(S2, p.y);
  if (H.is_zero()) {
    if (rr.is_zero()) return jac_dbl(p);
    return jac_inf<T>();
  }
  T HH = sqr(H), HHH = mul(H, HH), V = mul(p.x, HH);
  Jac<T> r;
  r.x = sub(sub(sqr(rr), HHH), dbl(V));
  r.y = sub(mul(rr, sub(V, r.x)), mul(p.y, HHH));
  r.z = mul(p.z, H);
  return r;
}
template <class T>
static Jac<T> jac_add(const Jac<T>& p, const Jac<T>& q) {
  if (q.is_inf()) return p;
  if (p.is_inf()) return q;
  T Z1Z1 = sqr(p.z), Z2Z2 = sqr(q.z);
  T U1 = mul(p.x, Z2Z2), U2 = mul(q.x, Z1Z1);
  T S1 = mul(mul(p.y, q.z), Z2Z2), S2 = mul(mul(q.y, p.z), Z1Z1);
  T H = sub(U2, U1), rr = sub(S2, S1);
  if (H.is_zero()) {
    if (rr.is_zero()) return jac_dbl(p);
    return jac_inf<T>();
  }
  T HH = sqr(H), HHH = mul(H, HH), V = mul(U1, HH);
  Jac<T> r;
  r.x = sub(sub(sqr(rr), HHH), dbl(V));
  r.y = sub(mul(rr, sub(V, r.x)), mul(S1, HHH));
  r.z = mul(mul(p.z, q.z), H);
  return r;
}
template <class T>
static Aff<T> jac_to_aff(const Jac<T>& p) {     // prj.affine (msm.nim:54,81)
  Aff<T> r;
  if (p.is_inf()) { r.x = T::zero(); r.y = T::zero(); return r; }
  T zi = inv(p.z), zi2 = sqr(zi);
  r.x = mul(p.x, zi2);
  r.y = mul(p.y, mul(zi2, zi));
  return r;
}
template <class T>
static Aff<T> aff_neg(const Aff<T>& p) { Aff<T> r; r.x = p.x; r.y = neg(p.y); return r; }
template <class T>
static Aff<T> aff_add(const Aff<T>& p, const Aff<T>& q) {   // addG1/addG2 (curves.nim:136-154): via projective
  Jac<T> a = jac_madd(jac_inf<T>(), p);
  return jac_to_aff(jac_madd(a, q));
}
template <class T>
static Aff<T> scalar_mul(const u64 k[4], const Aff<T>& p) {  // `**` (curves.nim:182-196)
  Jac<T> acc = jac_inf<T>();
  for (int i = 255; i >= 0; i--) {
    acc = jac_dbl(acc);
    if ((k[i >> 6] >> (i & 63)) & 1) acc = jac_madd(acc, p);
  }
  return jac_to_aff(acc);
}

// ------------------------------------------------------------------------------------------- MSM
static inline uint32_t get_bits(const u64 k[4], int pos, int c) {
  if (pos >= 256) return 0;
  int limb = pos >> 6, sh = pos & 63;
  u64 w = k[limb] >> sh;
  if (sh + c > 64 && limb + 1 < 4) w |= k[limb + 1] << (64 - sh);
  return (uint32_t)(w & ((1ull << c) - 1));
}
static int pick_window(size_t n) {
  int best = 2;
  double bc = 1e300;
  for (int c = 2; c <= 16; c++) {
    double W = (254 + c - 1) / c;
    double cost = W * ((double)n + 2.0 * (double)(1u << c));
    if (cost < bc) { bc = cost; best = c; }
  }
  return best;
}
// unsigned-window bucket method over one chunk (the per-chunk MSM of msm.nim:49 / :76); scalars standard form
template <class T>
static Jac<T> msm_chunk(const u64* scalars, const Aff<T>* pts, size_t n) {
  if (n == 0) return jac_inf<T>();
  int c = pick_window(n);
  int W = (254 + c - 1) / c;
  std::vector<Jac<T>> buckets((size_t)1 << c);
  Jac<T> total = jac_inf<T>();
  for (int w = W - 1; w >= 0; w--) {
    for (int j = 0; j < c; j++) total = jac_dbl(total);
    for (auto& b : buckets) b = jac_inf<T>();
    for (size_t i = 0; i < n; i++) {
      uint32_t d = get_bits(scalars + 4 * i, w * c, c);
      if (d) buckets[d] = jac_madd(buckets[d], pts[i]);
    }
    Jac<T> run = jac_inf<T>(), sum = jac_inf<T>();
    for (size_t d = buckets.size() - 1; d >= 1; d--) {
      run = jac_add(run, buckets[d]);
      sum = jac_add(sum, run);
    }
    total = jac_add(total, sum);
  }
  return total;
}

template <class T>
struct MsmTask {
  const u64* scalars;
  const Aff<T>* pts;
  size_t n;
  Aff<T> out;
};
template <class T>
static void* msm_worker(void* arg) {
  MsmTask<T>* t = (MsmTask<T>*)arg;
  t->out = jac_to_aff(msm_chunk<T>(t->scalars, t->pts, t->n));     // msm.nim:49-54
  return nullptr;
}
// msmMultiThreadedG1/G2 (msm.nim:89-158): scalars in standard form (the reference converts with toBig, :44)
template <class T>
static Aff<T> msm_multithreaded(int nthreads_hint, int ncpu, const u64* scalars, const Aff<T>* pts, size_t N) {
  int target = nthreads_hint <= 0 ? ncpu : (nthreads_hint < 256 ? nthreads_hint : 256);   // msm.nim:98
  long byn = (long)(N / 128);
  int nthreads = (int)(byn < target ? byn : target);                                       // msm.nim:99
  if (nthreads < 1) nthreads = 1;
  int ntasks = nthreads > 1 ? nthreads : 1;
  std::vector<MsmTask<T>> tasks(ntasks);
  std::vector<pthread_t> th(ntasks);
  size_t a = 0;
  for (int k = 0; k < ntasks; k++) {
    size_t b = (k < ntasks - 1) ? (N * (size_t)(k + 1)) / (size_t)ntasks : N;              // msm.nim:107-111
    tasks[k].scalars = scalars + 4 * a;
    tasks[k].pts = pts + a;
    tasks[k].n = b - a;
    a = b;
  }
  if (ntasks == 1) msm_worker<T>(&tasks[0]);
  else {
    for (int k = 0; k < ntasks; k++) pthread_create(&th[k], nullptr, msm_worker<T>, &tasks[k]);
    for (int k = 0; k < ntasks; k++) pthread_join(th[k], nullptr);
  }
  Aff<T> res;
  res.x = T::zero();
  res.y = T::zero();
  for (int k = 0; k < ntasks; k++) res = aff_add(res, tasks[k].out);                       // msm.nim:117-119
  return res;
}

// ------------------------------------------------------------------------------------------- NTT (literal)
static void fwd_worker(int m, size_t stride, const Fr* gp, const Fr* src, size_t so, Fr* buf, size_t bo, Fr* tgt,
                       size_t to) {   // ntt.nim:17-50
  if (m == 0) { tgt[to] = src[so]; return; }
  if (m == 1) {
    tgt[to] = add(src[so], src[so + stride]);
    tgt[to + 1] = sub(src[so], src[so + stride]);
    return;
  }
  size_t N = (size_t)1 << m, h = N >> 1;
  fwd_worker(m - 1, stride << 1, gp, src, so, buf, bo + N, buf, bo);
  fwd_worker(m - 1, stride << 1, gp, src, so + stride, buf, bo + N, buf, bo + h);
  for (size_t j = 0; j < h; j++) {
    Fr y = mul(gp[j * stride], buf[bo + j + h]);
    tgt[to + j] = add(buf[bo + j], y);
    tgt[to + j + h] = sub(buf[bo + j], y);
  }
}
static const u64 HALF_STD[4] = {0xa1f0fac9f8000001ull, 0x9419f4243cdcb848ull, 0xdc2822db40c0ac2eull, 0x183227397098d014ull};
static Fr div2(const Fr& a, const Fr& half) { return mul(a, half); }
static void inv_worker(int m, size_t stride, const Fr* gp, const Fr& half, const Fr* src, size_t so, Fr* buf,
                       size_t bo, Fr* tgt, size_t to) {   // ntt.nim:97-135
  if (m == 0) { tgt[to] = src[so]; return; }
  if (m == 1) {
    tgt[to] = div2(add(src[so], src[so + 1]), half);
    tgt[to + stride] = div2(sub(src[so], src[so + 1]), half);
    return;
  }
  size_t N = (size_t)1 << m, h = N >> 1;
  for (size_t j = 0; j < h; j++) {
    buf[bo + j] = div2(add(src[so + j], src[so + j + h]), half);
    buf[bo + j + h] = mul(sub(src[so + j], src[so + j + h]), gp[j * stride]);
  }
  inv_worker(m - 1, stride << 1, gp, half, buf, bo, buf, bo + N, tgt, to);
  inv_worker(m - 1, stride << 1, gp, half, buf, bo + h, buf, bo + N, tgt, to + stride);
}
static const u64 GEN28_STD[4] = {0x9bd61b6e725b19f0ull, 0x402d111e41112ed4ull, 0x00e0a7eb8ef62abcull, 0x2a3c09f0a58a7e85ull};
static Fr domain_gen(int log_n) {   // domain.nim:32-33
  Fr g;
  memcpy(g.v, GEN28_STD, 32);
  g = to_mont(g);
  for (int i = 0; i < 28 - log_n; i++) g = sqr(g);
  return g;
}
static void ntt_forward(const Fr* src, Fr* tgt, int log_n) {   // ntt.nim:55-77
  size_t N = (size_t)1 << log_n;
  std::vector<Fr> buf(2 * N), gp(N / 2 ? N / 2 : 1);
  Fr x = Fr::one(), gen = domain_gen(log_n);
  for (size_t i = 0; i < N / 2; i++) { gp[i] = x; x = mul(x, gen); }
  fwd_worker(log_n, 1, gp.data(), src, 0, buf.data(), 0, tgt, 0);
}
static void ntt_inverse(const Fr* src, Fr* tgt, int log_n) {   // ntt.nim:139-161
  size_t N = (size_t)1 << log_n;
  std::vector<Fr> buf(2 * N), gp(N / 2 ? N / 2 : 1);
  Fr half;
  memcpy(half.v, HALF_STD, 32);
  half = to_mont(half);
  Fr x = half, ginv = inv(domain_gen(log_n));
  for (size_t i = 0; i < N / 2; i++) { gp[i] = x; x = mul(x, ginv); }
  inv_worker(log_n, 1, gp.data(), half, src, 0, buf.data(), 0, tgt, 0);
}

// shiftEvalDomain (prover.nim:109-113) with multiplyByPowers (:96-106)
struct ShiftTask { const Fr* in; Fr* out; int log_n; Fr eta; };
static void* shift_worker(void* arg) {
  ShiftTask* t = (ShiftTask*)arg;
  size_t n = (size_t)1 << t->log_n;
  std::vector<Fr> cs(n), ds(n);
  ntt_inverse(t->in, cs.data(), t->log_n);
  ds[0] = cs[0];
  if (n > 1) ds[1] = mul(t->eta, cs[1]);
  Fr sp = t->eta;
  for (size_t i = 2; i < n; i++) { sp = mul(sp, t->eta); ds[i] = mul(sp, cs[i]); }
  ntt_forward(ds.data(), t->out, t->log_n);
  return nullptr;
}
static void quotient_cpu(const Fr* Az, const Fr* Bz, const Fr* Cz, int log_n, int flavour, int nthreads, Fr* qs) {
  size_t n = (size_t)1 << log_n;
  Fr eta = domain_gen(log_n + 1);                              // prover.nim:127,163
  std::vector<Fr> A1(n), B1(n), C1(n);
  ShiftTask t[3] = {{Az, A1.data(), log_n, eta}, {Bz, B1.data(), log_n, eta}, {Cz, C1.data(), log_n, eta}};
  if (nthreads > 1) {                                          // prover.nim:132-138, 167-173: exactly 3 tasks
    pthread_t th[3];
    for (int i = 0; i < 3; i++) pthread_create(&th[i], nullptr, shift_worker, &t[i]);
    for (int i = 0; i < 3; i++) pthread_join(th[i], nullptr);
  } else {
    for (int i = 0; i < 3; i++) shift_worker(&t[i]);
  }
  if (flavour == 1) {                                          // Snarkjs, prover.nim:176
    for (size_t j = 0; j < n; j++) qs[j] = sub(mul(A1[j], B1[j]), C1[j]);
    return;
  }
  Fr etan = eta;                                               // JensGroth, prover.nim:128,141-143
  for (int i = 0; i < log_n; i++) etan = sqr(etan);
  Fr invz = inv(sub(etan, Fr::one()));
  std::vector<Fr> ys(n), q1(n);
  for (size_t j = 0; j < n; j++) ys[j] = mul(sub(mul(A1[j], B1[j]), C1[j]), invz);
  ntt_inverse(ys.data(), q1.data(), log_n);
  Fr einv = inv(eta);
  qs[0] = q1[0];
  if (n > 1) qs[1] = mul(einv, q1[1]);
  Fr sp = einv;
  for (size_t i = 2; i < n; i++) { sp = mul(sp, einv); qs[i] = mul(sp, q1[i]); }
}

// ------------------------------------------------------------------------------------------- C interface
extern "C" {

// scalars standard form (n x 4 limbs); points affine Montgomery
void ora_msm_g1(const u64* scalars, const u64* points, size_t n, int nthreads_hint, int ncpu, u64* out) {
  Aff<Fp> r = msm_multithreaded<Fp>(nthreads_hint, ncpu, scalars, (const Aff<Fp>*)points, n);
  memcpy(out, &r, 64);
}
void ora_msm_g2(const u64* scalars, const u64* points, size_t n, int nthreads_hint, int ncpu, u64* out) {
  Aff<Fp2> r = msm_multithreaded<Fp2>(nthreads_hint, ncpu, scalars, (const Aff<Fp2>*)points, n);
  memcpy(out, &r, 128);
}
// Montgomery in / out, natural order
void ora_ntt(const u64* in, u64* out, int log_n, int inverse) {
  if (inverse) ntt_inverse((const Fr*)in, (Fr*)out, log_n);
  else ntt_forward((const Fr*)in, (Fr*)out, log_n);
}
// coeff records: packed 44 bytes (u32 m,row,col + value*R^2); witness standard form; outputs Montgomery
int ora_build_abc(const uint8_t* coeffs, size_t nnz, const u64* witness, int log_n, u64* az, u64* bz, u64* cz) {
  size_t n = (size_t)1 << log_n;
  Fr* A = (Fr*)az; Fr* B = (Fr*)bz; Fr* C = (Fr*)cz;
  for (size_t i = 0; i < n; i++) { A[i] = Fr::zero(); B[i] = Fr::zero(); }
  for (size_t k = 0; k < nnz; k++) {                          // prover.nim:63-67
    const uint8_t* rec = coeffs + 44 * k;
    uint32_t m, row, col;
    memcpy(&m, rec, 4); memcpy(&row, rec + 4, 4); memcpy(&col, rec + 8, 4);
    Fr v, w;
    memcpy(v.v, rec + 12, 32);
    memcpy(w.v, witness + 4 * (size_t)col, 32);
    Fr prod = mul(v, w);                                      // (c R^2)(w)/R = c w R
    if (m == 0) A[row] = add(A[row], prod);
    else if (m == 1) B[row] = add(B[row], prod);
    else return 1;                                            // "fatal error"
  }
  for (size_t i = 0; i < n; i++) C[i] = mul(A[i], B[i]);     // prover.nim:69-71
  return 0;
}
void ora_quotient(const u64* az, const u64* bz, const u64* cz, int log_n, int flavour, int nthreads, u64* qs) {
  quotient_cpu((const Fr*)az, (const Fr*)bz, (const Fr*)cz, log_n, flavour, nthreads, (Fr*)qs);
}

struct ora_zkey {
  uint32_t nvars, npubs, log_n, flavour;
  uint64_t ncoeffs;
  const uint8_t* coeffs;
  const u64 *a1, *b1, *b2, *c1, *h1;
  const u64 *alpha1, *beta1, *beta2, *delta1, *delta2;
};
// generateProofWithMask (prover.nim:215-304); witness, r, s standard form; proof = pi_a(8) pi_b(16) pi_c(8) limbs.
// phase_seconds[6]: ABC, quotient, pi_A, rho, pi_B, pi_C  (the reference's timing labels, prover.nim:244-297)
int ora_prove(const ora_zkey* zk, const u64* witness, const u64* r, const u64* s, int nthreads, int ncpu, u64* proof,
              double* phase_seconds) {
  struct timespec t0, t1;
  auto tick = [&] { clock_gettime(CLOCK_MONOTONIC, &t0); };
  auto tock = [&](int i) {
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (phase_seconds) phase_seconds[i] = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  };
  size_t n = (size_t)1 << zk->log_n;
  std::vector<Fr> A(n), B(n), C(n), qs(n), qstd(n);
  tick();
  if (ora_build_abc(zk->coeffs, zk->ncoeffs, witness, zk->log_n, (u64*)A.data(), (u64*)B.data(), (u64*)C.data())) return 1;
  tock(0);
  tick();
  quotient_cpu(A.data(), B.data(), C.data(), zk->log_n, zk->flavour, nthreads, qs.data());
  tock(1);
  for (size_t i = 0; i < n; i++) qstd[i] = from_mont(qs[i]);   // msm.nim:44 toBig
  const Aff<Fp>* alpha1 = (const Aff<Fp>*)zk->alpha1;
  const Aff<Fp>* beta1 = (const Aff<Fp>*)zk->beta1;
  const Aff<Fp>* delta1 = (const Aff<Fp>*)zk->delta1;
  const Aff<Fp2>* beta2 = (const Aff<Fp2>*)zk->beta2;
  const Aff<Fp2>* delta2 = (const Aff<Fp2>*)zk->delta2;
  tick();
  Aff<Fp> pi_a = aff_add(aff_add(*alpha1, scalar_mul(r, *delta1)),
                         msm_multithreaded<Fp>(nthreads, ncpu, witness, (const Aff<Fp>*)zk->a1, zk->nvars));
  tock(2);
  tick();
  Aff<Fp> rho = aff_add(aff_add(*beta1, scalar_mul(s, *delta1)),
                        msm_multithreaded<Fp>(nthreads, ncpu, witness, (const Aff<Fp>*)zk->b1, zk->nvars));
  tock(3);
  tick();
  Aff<Fp2> pi_b = aff_add(aff_add(*beta2, scalar_mul(s, *delta2)),
                          msm_multithreaded<Fp2>(nthreads, ncpu, witness, (const Aff<Fp2>*)zk->b2, zk->nvars));
  tock(4);
  tick();
  Fr rm, sm;
  memcpy(rm.v, r, 32);
  memcpy(sm.v, s, 32);
  Fr nrs = from_mont(neg(mul(to_mont(rm), to_mont(sm))));      // prover.nim:300
  Aff<Fp> pi_c = scalar_mul(s, pi_a);
  pi_c = aff_add(pi_c, scalar_mul(r, rho));
  pi_c = aff_add(pi_c, scalar_mul(nrs.v, *delta1));
  pi_c = aff_add(pi_c, msm_multithreaded<Fp>(nthreads, ncpu, (const u64*)qstd.data(), (const Aff<Fp>*)zk->h1, n));
  size_t nz = (size_t)zk->nvars - zk->npubs - 1;
  pi_c = aff_add(pi_c, msm_multithreaded<Fp>(nthreads, ncpu, witness + 4 * ((size_t)zk->npubs + 1),
                                             (const Aff<Fp>*)zk->c1, nz));
  tock(5);
  memcpy(proof, &pi_a, 64);
  memcpy(proof + 8, &pi_b, 128);
  memcpy(proof + 24, &pi_c, 64);
  return 0;
}
}
