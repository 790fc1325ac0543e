// Fake trusted setup on the GPU and fixed-base scalar multiplication.
//
// Replaces groth16/fake_setup.nim:201-326 (fakeCircuitSetup) as the fixture generator behind every
// benchmark configuration: Lagrange evaluations L_k(tau) (fake_setup.nim:255, poly.nim:242-250), the
// per-wire column sums a_j, b_j, c_j (fake_setup.nim:264-266), the K / IC scalars (:276-280), the H
// scalars of both flavours (:285-304) and the group elements k * g1, k * g2 (:268-271; curves.nim:182-196).
// The reference performs one double-and-add per point on the CPU; here every point is one thread using a
// fixed-base byte-window table of the generator.
#include <map>
#include <memory>
#include <mutex>
#include "../../include/g16b200.h"
#include "abc.cuh"
#include "common.cuh"
#include "ec.cuh"

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

// ---------------------------------------------------------------------------------------
// generators (curves.nim:112-124), standard form limbs
// ---------------------------------------------------------------------------------------
__device__ __constant__ uint32_t c_g2[4][8] = {
    {0x1f149701u, 0xbde23fabu, 0x70acf5b0u, 0x98aa68a5u, 0x55e3808fu, 0x7040f466u, 0x10df9cb8u, 0x1adcd0edu},
    {0x1c13b23bu, 0xfa15d21cu, 0xf7f31269u, 0xfbfbe620u, 0x0a3a82e6u, 0xc3cd2a1du, 0xf05a6082u, 0x09e847e9u},
    {0x0b7f6fc8u, 0xe6f91525u, 0xdbfc4cbeu, 0x1c7cdf52u, 0x19d4fcfdu, 0x1f7ca7aau, 0x8a531946u, 0x056c0116u},
    {0xa623235cu, 0xaaa86456u, 0xfc3c0dadu, 0xf553b878u, 0x9f30895du, 0xf5f40132u, 0x2d02dd77u, 0x0efe500au}};

template <class F>
struct Gen;
template <>
struct Gen<Fp> {
  static __device__ Affine<Fp> get() {
    Affine<Fp> g;
    g.x = Fp::one();
    g.y = fdbl(Fp::one());
    return g;
  }
};
template <>
struct Gen<Fp2> {
  static __device__ Affine<Fp2> get() {
    Fp c[4];
    for (int k = 0; k < 4; k++) {
      for (int i = 0; i < 8; i++) c[k].v[i] = c_g2[k][i];
      c[k] = to_mont(c[k]);
    }
    Affine<Fp2> g;
    g.x.c0 = c[0];
    g.x.c1 = c[1];
    g.y.c0 = c[2];
    g.y.c1 = c[3];
    return g;
  }
};

// table[w * 255 + (d - 1)] = (d << 8w) * G, w < 32, d in 1..255
template <class F>
__global__ void k_fb_table(Affine<F>* table) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 32 * 255) return;
  uint32_t w = t / 255, d = t % 255 + 1;
  uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  k[w >> 2] = d << (8 * (w & 3));
  XYZZ<F> p = xyzz_scalar_mul(k, Gen<F>::get());
  Affine<F> a;
  xyzz_to_affine_ni(a, p);
  stv(table + t, a);
}

template <class F>
__global__ void __launch_bounds__(128) k_fixed_base(const Fr* __restrict__ scalars, uint32_t n,
                                                    const Affine<F>* __restrict__ table, Affine<F>* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = ldv(scalars + i);
  XYZZ<F> acc = xyzz_inf<F>();
#pragma unroll 1
  for (int w = 0; w < 32; w++) {
    uint32_t b = (s.v[w >> 2] >> (8 * (w & 3))) & 255u;
    if (b) xyzz_madd_ni(acc, acc, ldv(table + w * 255 + (b - 1)));
  }
  Affine<F> a;
  xyzz_to_affine_ni(a, acc);
  stv(out + i, a);
}

static std::mutex g_fb_mutex;
static std::map<std::pair<int, int>, std::unique_ptr<DevBuf>> g_fb_tables;   // (device, g2)

template <class F>
static const Affine<F>* fb_table(cudaStream_t stream) {
  int dev = 0;
  G16_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_fb_mutex);
  auto key = std::make_pair(dev, (int)(sizeof(F) > sizeof(Fp)));
  auto it = g_fb_tables.find(key);
  if (it != g_fb_tables.end()) return it->second->template as<Affine<F>>();
  std::unique_ptr<DevBuf> b(new DevBuf());
  b->ensure((size_t)32 * 255 * sizeof(Affine<F>));
  k_fb_table<F><<<div_up(32 * 255, 64), 64, 0, stream>>>(b->template as<Affine<F>>());
  G16_LAUNCH_CHECK();
  G16_CUDA(cudaStreamSynchronize(stream));
  const Affine<F>* p = b->template as<Affine<F>>();
  g_fb_tables[key] = std::move(b);
  return p;
}

void fb_release_tables() {
  std::lock_guard<std::mutex> lock(g_fb_mutex);
  g_fb_tables.clear();
}

// scalars: standard form, device; out: device
template <class F>
void fixed_base_mul(const Fr* scalars_dev, size_t n, Affine<F>* out_dev, cudaStream_t stream) {
  if (!n) return;
  const Affine<F>* table = fb_table<F>(stream);
  k_fixed_base<F><<<div_up(n, 128), 128, 0, stream>>>(scalars_dev, (uint32_t)n, table, out_dev);
  G16_LAUNCH_CHECK();
}
template void fixed_base_mul<Fp>(const Fr*, size_t, Affine<Fp>*, cudaStream_t);
template void fixed_base_mul<Fp2>(const Fr*, size_t, Affine<Fp2>*, cudaStream_t);

// ---------------------------------------------------------------------------------------
// field side of the fake setup
// ---------------------------------------------------------------------------------------
struct alignas(16) SetupConsts {
  Fr omega, eta, n_inv, n2_inv, tau, tau_n_m1, tau_2n_m1, alpha, beta, gamma_inv, delta_inv;
};

__device__ __constant__ uint32_t c_gen28s[8] = {0x725b19f0u, 0x9bd61b6eu, 0x41112ed4u, 0x402d111eu,
                                                0x8ef62abcu, 0x00e0a7ebu, 0xa58a7e85u, 0x2a3c09f0u};

__global__ void k_setup_consts(const g16_toxic* tox, int log_n, SetupConsts* c) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  auto load = [](const uint64_t* p) {
    Fr x;
    for (int i = 0; i < 4; i++) {
      x.v[2 * i] = (uint32_t)p[i];
      x.v[2 * i + 1] = (uint32_t)(p[i] >> 32);
    }
    return to_mont(x);
  };
  Fr g;
  for (int i = 0; i < 8; i++) g.v[i] = c_gen28s[i];
  g = to_mont(g);
  Fr eta = g;
  for (int i = 0; i < 28 - log_n - 1; i++) eta = fsqr(eta);   // omega_{2n}
  c->eta = eta;
  c->omega = fsqr(eta);                                        // domain.nim:32-33
  Fr nn = Fr::zero();
  nn.v[0] = 1u << log_n;
  Fr ninv = finv(to_mont(nn));
  c->n_inv = ninv;
  Fr two = fdbl(Fr::one());
  c->n2_inv = fmul(ninv, finv(two));
  Fr tau = load(tox->tau);
  c->tau = tau;
  Fr tn = tau;
  for (int i = 0; i < log_n; i++) tn = fsqr(tn);
  c->tau_n_m1 = fsub(tn, Fr::one());
  c->tau_2n_m1 = fsub(fsqr(tn), Fr::one());
  c->alpha = load(tox->alpha);
  c->beta = load(tox->beta);
  c->gamma_inv = finv(load(tox->gamma));
  c->delta_inv = finv(load(tox->delta));
}

// L[i] = L_i(tau) on the size-n domain (poly.nim:242-250), Montgomery
__global__ void k_setup_lagrange(const SetupConsts* c, uint32_t n, Fr* L, int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr wk = fpow_u64(c->omega, i);
  Fr den = fsub(c->tau, wk);
  if (fis_zero(den)) atomicOr(err, 1);     // "point should be outside the domain"
  stv(L + i, fmul(fmul(fmul(wk, c->tau_n_m1), c->n_inv), finv(den)));
}

// H scalars: Snarkjs: delta^-1 * L^(2n)_{2i+1}(tau) (fake_setup.nim:301-304); JensGroth: delta^-1 tau^i Z(tau) (:293-295)
__global__ void k_setup_h(const SetupConsts* c, uint32_t n, int flavour, Fr* h, int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr out;
  if (flavour == G16_FLAVOUR_SNARKJS) {
    Fr wk = fpow_u64(c->eta, 2ull * i + 1ull);
    Fr den = fsub(c->tau, wk);
    if (fis_zero(den)) atomicOr(err, 1);
    out = fmul(fmul(fmul(fmul(wk, c->tau_2n_m1), c->n2_inv), finv(den)), c->delta_inv);
  } else {
    out = fmul(fmul(fpow_u64(c->tau, i), c->tau_n_m1), c->delta_inv);
  }
  stv(h + i, out);
}

// one thread per wire j: column sums of A, B, C against L, then the K / IC combination
__global__ void k_setup_columns(const SetupConsts* c, const Fr* __restrict__ L, uint32_t nvars, uint32_t npubs,
                                uint32_t neqs, const uint32_t* pA, const uint32_t* rA, const Fr* vA,
                                const uint32_t* pB, const uint32_t* rB, const Fr* vB, const uint32_t* pC,
                                const uint32_t* rC, const Fr* vC, Fr* a_out, Fr* b_out, Fr* k_out, Fr* ic_out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nvars) return;
  Fr a = Fr::zero(), b = Fr::zero(), cc = Fr::zero();
  for (uint32_t t = pA[j]; t < pA[j + 1]; t++) a = fadd(a, fmul(to_mont(ldv(vA + t)), ldv(L + rA[t])));
  for (uint32_t t = pB[j]; t < pB[j + 1]; t++) b = fadd(b, fmul(to_mont(ldv(vB + t)), ldv(L + rB[t])));
  for (uint32_t t = pC[j]; t < pC[j + 1]; t++) cc = fadd(cc, fmul(to_mont(ldv(vC + t)), ldv(L + rC[t])));
  if (j <= npubs) a = fadd(a, ldv(L + neqs + j));              // dummy rows, fake_setup.nim:182-185
  Fr comb = fadd(fadd(fmul(c->beta, a), fmul(c->alpha, b)), cc);
  stv(a_out + j, a);
  stv(b_out + j, b);
  if (j <= npubs) stv(ic_out + j, fmul(comb, c->gamma_inv));   // fake_setup.nim:276-277
  else stv(k_out + (j - npubs - 1), fmul(comb, c->delta_inv)); // fake_setup.nim:279-280
}

__global__ void k_spec_scalars(const g16_toxic* tox, Fr* out) {   // alpha, beta, delta, beta, gamma, delta (std)
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint64_t* src[6] = {tox->alpha, tox->beta, tox->delta, tox->beta, tox->gamma, tox->delta};
  for (int k = 0; k < 6; k++)
    for (int i = 0; i < 4; i++) {
      out[k].v[2 * i] = (uint32_t)src[k][i];
      out[k].v[2 * i + 1] = (uint32_t)(src[k][i] >> 32);
    }
}

static void d2h(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (dst && bytes) G16_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
}

void fake_setup(const g16_r1cs_view& r, const g16_toxic& toxic, uint32_t* log_domain_out, g16_setup_out& out) {
  G16_REQUIRE(r.nvars >= r.npubs + 1, "nvars must be at least npubs + 1");
  size_t rows = (size_t)r.neqs + r.npubs + 1;
  int log_n = ceil_log2_sz(rows);                               // fake_setup.nim:205
  if (log_n < 1) log_n = 1;
  G16_REQUIRE(log_n <= 26, "domain too large");
  size_t n = (size_t)1 << log_n;
  if (log_domain_out) *log_domain_out = (uint32_t)log_n;
  cudaStream_t s = nullptr;
  G16_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  try {
    DevBuf tox, consts, err, L;
    tox.ensure(sizeof(g16_toxic));
    consts.ensure(sizeof(SetupConsts));
    err.ensure(4);
    L.ensure(n * sizeof(Fr));
    G16_CUDA(cudaMemcpyAsync(tox.p, &toxic, sizeof(toxic), cudaMemcpyHostToDevice, s));
    G16_CUDA(cudaMemsetAsync(err.p, 0, 4, s));
    k_setup_consts<<<1, 32, 0, s>>>(tox.as<g16_toxic>(), log_n, consts.as<SetupConsts>());
    G16_LAUNCH_CHECK();
    k_setup_lagrange<<<div_up(n, 128), 128, 0, s>>>(consts.as<SetupConsts>(), (uint32_t)n, L.as<Fr>(), err.as<int>());
    G16_LAUNCH_CHECK();

    SparseCsr csr[3];
    for (int m = 0; m < 3; m++) {
      size_t nnz = r.nnz[m];
      DevBuf dk, dr, dv;
      dk.ensure(nnz * 4 + 4);
      dr.ensure(nnz * 4 + 4);
      dv.ensure(nnz * sizeof(Fr) + sizeof(Fr));
      if (nnz) {
        G16_REQUIRE(r.rows[m] && r.cols[m] && r.vals[m], "r1cs view: missing matrix arrays");
        G16_CUDA(cudaMemcpyAsync(dk.p, r.cols[m], nnz * 4, cudaMemcpyHostToDevice, s));
        G16_CUDA(cudaMemcpyAsync(dr.p, r.rows[m], nnz * 4, cudaMemcpyHostToDevice, s));
        G16_CUDA(cudaMemcpyAsync(dv.p, r.vals[m], nnz * sizeof(Fr), cudaMemcpyHostToDevice, s));
      }
      coo_to_csr(csr[m], dk.as<uint32_t>(), dr.as<uint32_t>(), dv.as<Fr>(), nnz, r.nvars, s);
    }
    size_t nk = (size_t)r.nvars - r.npubs - 1;
    DevBuf da, db, dkk, dic, dh, dstd;
    da.ensure((size_t)r.nvars * sizeof(Fr));
    db.ensure((size_t)r.nvars * sizeof(Fr));
    dkk.ensure(nk * sizeof(Fr) + sizeof(Fr));
    dic.ensure(((size_t)r.npubs + 1) * sizeof(Fr));
    dh.ensure(n * sizeof(Fr));
    k_setup_columns<<<div_up(r.nvars, 128), 128, 0, s>>>(
        consts.as<SetupConsts>(), L.as<Fr>(), r.nvars, r.npubs, r.neqs, csr[0].ptr.as<uint32_t>(),
        csr[0].other.as<uint32_t>(), csr[0].vals.as<Fr>(), csr[1].ptr.as<uint32_t>(), csr[1].other.as<uint32_t>(),
        csr[1].vals.as<Fr>(), csr[2].ptr.as<uint32_t>(), csr[2].other.as<uint32_t>(), csr[2].vals.as<Fr>(),
        da.as<Fr>(), db.as<Fr>(), dkk.as<Fr>(), dic.as<Fr>());
    G16_LAUNCH_CHECK();
    k_setup_h<<<div_up(n, 128), 128, 0, s>>>(consts.as<SetupConsts>(), (uint32_t)n, (int)r.flavour, dh.as<Fr>(),
                                             err.as<int>());
    G16_LAUNCH_CHECK();
    int herr = 0;
    G16_CUDA(cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, s));
    G16_CUDA(cudaStreamSynchronize(s));
    G16_REQUIRE(herr == 0, "point should be outside the domain (poly.nim:247)");

    // discrete logs to standard form (in place), then the group elements
    fr_from_mont(da.as<Fr>(), da.as<Fr>(), r.nvars, s);
    fr_from_mont(db.as<Fr>(), db.as<Fr>(), r.nvars, s);
    fr_from_mont(dkk.as<Fr>(), dkk.as<Fr>(), nk, s);
    fr_from_mont(dic.as<Fr>(), dic.as<Fr>(), (size_t)r.npubs + 1, s);
    fr_from_mont(dh.as<Fr>(), dh.as<Fr>(), n, s);
    d2h(out.dlog_a, da.p, (size_t)r.nvars * sizeof(Fr), s);
    d2h(out.dlog_b, db.p, (size_t)r.nvars * sizeof(Fr), s);
    d2h(out.dlog_k, dkk.p, nk * sizeof(Fr), s);
    d2h(out.dlog_ic, dic.p, ((size_t)r.npubs + 1) * sizeof(Fr), s);
    d2h(out.dlog_h, dh.p, n * sizeof(Fr), s);

    DevBuf pts;
    size_t maxn = n > r.nvars ? n : r.nvars;
    pts.ensure(maxn * sizeof(G2Affine));
    auto emit_g1 = [&](const DevBuf& sc, size_t cnt, uint64_t* host) {
      if (!host || !cnt) return;
      fixed_base_mul<Fp>(sc.as<Fr>(), cnt, pts.as<G1Affine>(), s);
      G16_CUDA(cudaMemcpyAsync(host, pts.p, cnt * sizeof(G1Affine), cudaMemcpyDeviceToHost, s));
      G16_CUDA(cudaStreamSynchronize(s));
    };
    emit_g1(da, r.nvars, out.points_a1);
    emit_g1(db, r.nvars, out.points_b1);
    emit_g1(dkk, nk, out.points_c1);
    emit_g1(dh, n, out.points_h1);
    emit_g1(dic, (size_t)r.npubs + 1, out.points_ic);
    if (out.points_b2) {
      fixed_base_mul<Fp2>(db.as<Fr>(), r.nvars, pts.as<G2Affine>(), s);
      G16_CUDA(cudaMemcpyAsync(out.points_b2, pts.p, (size_t)r.nvars * sizeof(G2Affine), cudaMemcpyDeviceToHost, s));
      G16_CUDA(cudaStreamSynchronize(s));
    }
    if (out.spec) {
      dstd.ensure(6 * sizeof(Fr));
      k_spec_scalars<<<1, 32, 0, s>>>(tox.as<g16_toxic>(), dstd.as<Fr>());
      G16_LAUNCH_CHECK();
      fixed_base_mul<Fp>(dstd.as<Fr>(), 3, pts.as<G1Affine>(), s);
      G16_CUDA(cudaMemcpyAsync(out.spec, pts.p, 3 * sizeof(G1Affine), cudaMemcpyDeviceToHost, s));
      G16_CUDA(cudaStreamSynchronize(s));
      fixed_base_mul<Fp2>(dstd.as<Fr>() + 3, 3, pts.as<G2Affine>(), s);
      G16_CUDA(cudaMemcpyAsync(out.spec + 24, pts.p, 3 * sizeof(G2Affine), cudaMemcpyDeviceToHost, s));
      G16_CUDA(cudaStreamSynchronize(s));
    }
    G16_CUDA(cudaStreamSynchronize(s));
  } catch (...) {
    cudaStreamDestroy(s);
    throw;
  }
  cudaStreamDestroy(s);
}

}  // namespace g16
