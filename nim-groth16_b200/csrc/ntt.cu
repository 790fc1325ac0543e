// Fr NTT and quotient kernels for sm_100a.
//
// Replaces groth16/math/ntt.nim:55-77 (forwardNTT), :139-161 (inverseNTT), groth16/math/domain.nim:28-46
// (createDomain) and the quotient procs groth16/prover.nim:96-113 (multiplyByPowers, shiftEvalDomain),
// :118-148 (computeQuotientPointwise) and :158-181 (computeSnarkjsScalarCoeffs).
//
// Design (DESIGN.md "NTT"): multi-pass radix-2^k transform, each pass running k butterfly stages on a
// 64 KiB shared-memory tile (limb-major, bank-conflict-free); inverse transforms are
// decimation-in-frequency (natural in, bit-reversed out), forward transforms inside the quotient are
// decimation-in-time (bit-reversed in, natural out), so shiftEvalDomain needs no permutation pass:
// the coset factor eta^i / n is applied, indexed by bit-reversal, in the last inverse pass.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <map>
#include <memory>
#include <mutex>
#include "common.cuh"
#include "field.cuh"
#include <atomic>
#include "ntt.cuh"
#include "ntt_plan.cuh"

namespace g16 {

// ---------------------------------------------------------------------------------------
// domain constants and tables
// ---------------------------------------------------------------------------------------
// gen28 (domain.nim:26), standard form, little-endian limbs
__device__ __constant__ uint32_t c_gen28[8] = {0x725b19f0u, 0x9bd61b6eu, 0x41112ed4u, 0x402d111eu,
                                               0x8ef62abcu, 0x00e0a7ebu, 0xa58a7e85u, 0x2a3c09f0u};

// consts[0]=omega_n  [1]=omega_n^-1  [2]=1/n  [3]=eta=omega_2n  [4]=eta^-1  [5]=1/(eta^n-1)  [6]=1  [7]=omega_n*... spare
__global__ void k_domain_consts(int log_n, Fr* consts) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Fr g;
#pragma unroll
  for (int i = 0; i < 8; i++) g.v[i] = c_gen28[i];
  g = to_mont(g);
  Fr eta = g;                                   // omega_{2n} = gen28^(2^(28-log_n-1))   (prover.nim:127,163)
  for (int i = 0; i < 28 - log_n - 1; i++) eta = fsqr(eta);
  Fr omega = fsqr(eta);                         // domain.nim:32-33
  Fr nn = Fr::zero();
  nn.v[0] = 1u << log_n;
  consts[0] = omega;
  consts[1] = finv(omega);                      // domain.nim:43
  consts[2] = finv(to_mont(nn));                // domain.nim:44
  consts[3] = eta;
  consts[4] = finv(eta);
  Fr etan = eta;
  for (int i = 0; i < log_n; i++) etan = fsqr(etan);
  consts[5] = finv(fsub(etan, Fr::one()));      // prover.nim:128 invZ1
  consts[6] = Fr::one();
  consts[7] = Fr::zero();
}

// out[j] = scale * base^j for j < count; each thread produces 8 consecutive powers.
__global__ void k_gen_powers(Fr* out, uint32_t count, const Fr* base_p, const Fr* scale_p) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t j0 = t * 8u;
  if (j0 >= count) return;
  Fr base = *base_p;
  Fr x = fmul(fpow_u64(base, j0), *scale_p);
  for (uint32_t j = j0; j < j0 + 8u && j < count; j++) {
    out[j] = x;
    x = fmul(x, base);
  }
}

struct NttTables {
  int log_n = -1;
  DevBuf consts;      // 8 Fr
  DevBuf tw_fwd;      // omega^j,   j < n/2
  DevBuf tw_inv;      // omega^-j,  j < n/2
  DevBuf coset;       // eta^j / n, j < n      (prover.nim:96-106 fused with the 1/n of ntt.nim:139)
  DevBuf coset_inv;   // eta^-j / n, j < n     (prover.nim:143 fused with 1/n), built lazily
  bool have_coset_inv = false;
};

static std::mutex g_tab_mutex;
static std::map<std::pair<int, int>, std::unique_ptr<NttTables>> g_tables;   // (device, log_n)

static NttTables& ntt_tables(int log_n, cudaStream_t stream) {
  G16_REQUIRE(log_n >= 1 && log_n <= 26, "NTT domain must be 2^1 .. 2^26");
  int dev = 0;
  G16_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  auto key = std::make_pair(dev, log_n);
  auto it = g_tables.find(key);
  if (it != g_tables.end()) return *it->second;
  std::unique_ptr<NttTables> t(new NttTables());
  t->log_n = log_n;
  size_t n = (size_t)1 << log_n;
  t->consts.ensure(8 * sizeof(Fr));
  t->tw_fwd.ensure((n / 2) * sizeof(Fr));
  t->tw_inv.ensure((n / 2) * sizeof(Fr));
  t->coset.ensure(n * sizeof(Fr));
  Fr* c = t->consts.as<Fr>();
  k_domain_consts<<<1, 32, 0, stream>>>(log_n, c);
  G16_LAUNCH_CHECK();
  uint32_t half = (uint32_t)(n / 2);
  k_gen_powers<<<div_up(div_up(half, 8), 128), 128, 0, stream>>>(t->tw_fwd.as<Fr>(), half, c + 0, c + 6);
  G16_LAUNCH_CHECK();
  k_gen_powers<<<div_up(div_up(half, 8), 128), 128, 0, stream>>>(t->tw_inv.as<Fr>(), half, c + 1, c + 6);
  G16_LAUNCH_CHECK();
  k_gen_powers<<<div_up(div_up(n, 8), 128), 128, 0, stream>>>(t->coset.as<Fr>(), (uint32_t)n, c + 3, c + 2);
  G16_LAUNCH_CHECK();
  G16_CUDA(cudaStreamSynchronize(stream));
  NttTables& ref = *t;
  g_tables[key] = std::move(t);
  return ref;
}

static void ntt_ensure_coset_inv(NttTables& t, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  if (t.have_coset_inv) return;
  size_t n = (size_t)1 << t.log_n;
  t.coset_inv.ensure(n * sizeof(Fr));
  Fr* c = t.consts.as<Fr>();
  k_gen_powers<<<div_up(div_up(n, 8), 128), 128, 0, stream>>>(t.coset_inv.as<Fr>(), (uint32_t)n, c + 4, c + 2);
  G16_LAUNCH_CHECK();
  G16_CUDA(cudaStreamSynchronize(stream));
  t.have_coset_inv = true;
}

void ntt_release_tables() {
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  g_tables.clear();
}

// ---------------------------------------------------------------------------------------
// the pass kernel
// ---------------------------------------------------------------------------------------
enum { SC_NONE = 0, SC_CONST = 1, SC_TABLE_BITREV = 2, SC_TABLE = 3 };

struct NttPassArgs {
  const Fr* src;        // batch b reads src + b * src_stride
  Fr* dst;              // batch b writes dst + b * dst_stride (dst == src for in-place passes)
  size_t src_stride, dst_stride;
  const Fr* tw;         // n/2 powers of the root for this direction
  const Fr* scale;      // SC_CONST: one element; SC_TABLE_BITREV: n elements
  int log_n, t_lo, k, logC;
  int scale_mode;
  int bitrev_store;     // store element g at dst[bitrev(g)]
  int bitrev_load;      // load element g from src[bitrev(g)]  (gather: natural-order input of a DIT transform)
};

#ifndef G16_NTT_THREADS
#define G16_NTT_THREADS 256
#endif
#ifndef G16_NTT_MINB
#define G16_NTT_MINB 1
#endif
constexpr int NTT_THREADS = G16_NTT_THREADS;

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// shared-memory tile, limb-major: limb l of element p at sm[l * E + p]
__device__ __forceinline__ Fr sm_get(const uint32_t* sm, uint32_t E, uint32_t p) {
  Fr r;
#pragma unroll
  for (int l = 0; l < 8; l++) r.v[l] = sm[l * E + p];
  return r;
}
__device__ __forceinline__ void sm_put(uint32_t* sm, uint32_t E, uint32_t p, const Fr& r) {
#pragma unroll
  for (int l = 0; l < 8; l++) sm[l * E + p] = r.v[l];
}

template <bool DIF>
__global__ void __launch_bounds__(NTT_THREADS, G16_NTT_MINB) k_ntt_pass(NttPassArgs a) {
  extern __shared__ uint32_t sm[];
  const uint32_t E = 1u << (a.k + a.logC);
  const uint32_t tile = blockIdx.x;
  const Fr* src = a.src + (size_t)blockIdx.y * a.src_stride;
  Fr* dst = a.dst + (size_t)blockIdx.y * a.dst_stride;

  for (uint32_t p = threadIdx.x; p < E; p += NTT_THREADS) {
    uint32_t g = ntt_global_index(tile, p, a.t_lo, a.k, a.logC);
    sm_put(sm, E, p, ld_fr(src + (a.bitrev_load ? (__brev(g) >> (32 - a.log_n)) : g)));
  }
  __syncthreads();

  for (int it = 0; it < a.k; it++) {
    const int s = DIF ? (a.k - 1 - it) : it;
    const bool trivial = (a.t_lo + s) == 0;           // all twiddles are omega^0
    for (uint32_t q = threadIdx.x; q < E / 2; q += NTT_THREADS) {
      uint32_t pu, pv, e;
      ntt_butterfly_index(tile, q, s, a.t_lo, a.logC, a.log_n, pu, pv, e);
      Fr u = sm_get(sm, E, pu);
      Fr v = sm_get(sm, E, pv);
      if (DIF) {
        Fr d = fsub(u, v);
        if (!trivial) d = fmul(d, ldg_fr(a.tw + e));
        sm_put(sm, E, pu, fadd(u, v));
        sm_put(sm, E, pv, d);
      } else {
        if (!trivial) v = fmul(v, ldg_fr(a.tw + e));
        sm_put(sm, E, pu, fadd(u, v));
        sm_put(sm, E, pv, fsub(u, v));
      }
    }
    __syncthreads();
  }

  for (uint32_t p = threadIdx.x; p < E; p += NTT_THREADS) {
    uint32_t g = ntt_global_index(tile, p, a.t_lo, a.k, a.logC);
    Fr x = sm_get(sm, E, p);
    uint32_t gr = __brev(g) >> (32 - a.log_n);
    if (a.scale_mode == SC_CONST) x = fmul(x, ldg_fr(a.scale));
    else if (a.scale_mode == SC_TABLE_BITREV) x = fmul(x, ldg_fr(a.scale + gr));
    else if (a.scale_mode == SC_TABLE) x = fmul(x, ldg_fr(a.scale + g));
    st_fr(dst + (a.bitrev_store ? gr : g), x);
  }
}

// ---------------------------------------------------------------------------------------
// the pass kernel for full tiles (k + logC == 11): radix-8 groups in registers
//
// A thread holds 8 elements (64 registers) and runs up to three butterfly stages on them -- 12 butterflies, 12
// twiddle multiplications, no shared memory and no barrier -- then the CTA regroups through shared memory for the
// next three index bits.  An 11-bit pass is 3+3+3+2 stages with three exchanges where the radix-2 kernel above has
// eleven shared-memory round trips and barriers; shared memory moves 16-byte halves (LDS.128 / STS.128, XOR-swizzled
// so that a quarter-warp hits eight distinct bank groups).  Global stores are sector-complete: the two lanes of a
// pair swap halves by shuffle so that one store instruction writes the 32 bytes of an element contiguously (no
// partial-sector writes, which is what amplified the DRAM writes of the radix-2 kernel 3-5x).
// 256 threads x 8 elements = one 2048-element tile, 2 CTAs per SM: 512 tiles of a 2^20 transform on 296 slots.
// ---------------------------------------------------------------------------------------
constexpr int NTT8_THREADS = 256;

// tile position of element j (0..7) of thread t in a group that owns position bits [b, b + r): the low r bits of
// j are the group's stage bits, the other 3 - r bits select one of 2^(3-r) independent sets
__device__ __forceinline__ uint32_t ntt8_pos(uint32_t t, uint32_t j, int b, int r) {
  const uint32_t jlow = j & ((1u << r) - 1u), jset = j >> r;
  const uint32_t u = (jset << 8) | t;
  return ((u >> b) << (b + r)) | (jlow << b) | (u & ((1u << b) - 1u));
}
__device__ __forceinline__ uint32_t ntt8_phys(uint32_t p) { return p ^ ((p >> 3) & 7u); }

__device__ __forceinline__ uint4 shfl_xor4(uint4 v, int m) {
  v.x = __shfl_xor_sync(0xffffffffu, v.x, m);
  v.y = __shfl_xor_sync(0xffffffffu, v.y, m);
  v.z = __shfl_xor_sync(0xffffffffu, v.z, m);
  v.w = __shfl_xor_sync(0xffffffffu, v.w, m);
  return v;
}

template <bool DIF>
__global__ void __launch_bounds__(NTT8_THREADS, 2) k_ntt_pass8(NttPassArgs a) {
  extern __shared__ uint4 sm4[];
  uint4* const sm_lo = sm4;
  uint4* const sm_hi = sm4 + (1u << NTT_TILE_LOG);
  const uint32_t t = threadIdx.x, tile = blockIdx.x;
  const Fr* src = a.src + (size_t)blockIdx.y * a.src_stride;
  Fr* dst = a.dst + (size_t)blockIdx.y * a.dst_stride;
  const int ng = (a.k + 2) / 3;
  const int rlast = a.k - 3 * (ng - 1);
  Fr x[8];
  uint32_t g[8];

#pragma unroll 1
  for (int gi = 0; gi < ng; gi++) {
    // DIF walks the stage bits from the top, DIT from the bottom; the group with fewer than 3 bits comes last
    const int r = gi < ng - 1 ? 3 : rlast;
    const int b = DIF ? (gi < ng - 1 ? a.logC + a.k - 3 * (gi + 1) : a.logC) : a.logC + 3 * gi;
    if (gi == 0) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        g[j] = ntt_global_index(tile, ntt8_pos(t, j, b, r), a.t_lo, a.k, a.logC);
        x[j] = ld_fr(src + (a.bitrev_load ? (__brev(g[j]) >> (32 - a.log_n)) : g[j]));
      }
    } else {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t p = ntt8_pos(t, j, b, r), P = ntt8_phys(p);
        g[j] = ntt_global_index(tile, p, a.t_lo, a.k, a.logC);
        const uint4 lo = sm_lo[P], hi = sm_hi[P];
        x[j].v[0] = lo.x; x[j].v[1] = lo.y; x[j].v[2] = lo.z; x[j].v[3] = lo.w;
        x[j].v[4] = hi.x; x[j].v[5] = hi.y; x[j].v[6] = hi.z; x[j].v[7] = hi.w;
      }
    }
    const int tg0 = a.t_lo + (b - a.logC);                // global index bit of the group's stage 0
#pragma unroll
    for (int qq = 0; qq < 3; qq++) {
      const int q = DIF ? 2 - qq : qq;
      if (q < r) {
        const int tg = tg0 + q;
        const uint32_t mask = (1u << tg) - 1u;
        const int sh = a.log_n - 1 - tg;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (j & (1 << q)) continue;
          Fr& u = x[j];
          Fr& v = x[j | (1 << q)];
          if (DIF) {
            Fr d = fsub(u, v);
            if (tg != 0) d = fmul(d, ldg_fr(a.tw + ((g[j] & mask) << sh)));
            u = fadd(u, v);
            v = d;
          } else {
            Fr w = v;
            if (tg != 0) w = fmul(w, ldg_fr(a.tw + ((g[j] & mask) << sh)));
            v = fsub(u, w);
            u = fadd(u, w);
          }
        }
      }
    }
    if (gi < ng - 1) {
      if (gi > 0) __syncthreads();                        // everyone has read the previous exchange
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t P = ntt8_phys(ntt8_pos(t, j, b, r));
        sm_lo[P] = make_uint4(x[j].v[0], x[j].v[1], x[j].v[2], x[j].v[3]);
        sm_hi[P] = make_uint4(x[j].v[4], x[j].v[5], x[j].v[6], x[j].v[7]);
      }
    }
  }

  const bool odd = threadIdx.x & 1u;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    Fr v = x[j];
    const uint32_t gr = __brev(g[j]) >> (32 - a.log_n);
    if (a.scale_mode == SC_CONST) v = fmul(v, ldg_fr(a.scale));
    else if (a.scale_mode == SC_TABLE_BITREV) v = fmul(v, ldg_fr(a.scale + gr));
    else if (a.scale_mode == SC_TABLE) v = fmul(v, ldg_fr(a.scale + g[j]));
    const uint32_t di = a.bitrev_store ? gr : g[j];
    const uint4 lo = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]), hi = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
    // lane pair (even, odd): the even lane hands over its high half and takes the odd lane's low half, so that
    // store 1 writes the even lane's element and store 2 the odd lane's element as 32 contiguous bytes
    const uint4 got = shfl_xor4(odd ? lo : hi, 1);
    const uint32_t dj = __shfl_xor_sync(0xffffffffu, di, 1);
    uint4* const p1 = reinterpret_cast<uint4*>(dst + (odd ? dj : di)) + (odd ? 1 : 0);
    uint4* const p2 = reinterpret_cast<uint4*>(dst + (odd ? di : dj)) + (odd ? 1 : 0);
    *p1 = odd ? got : lo;
    *p2 = odd ? hi : got;
  }
}

static void launch_pass(bool dif, const NttPassArgs& a, int batch, cudaStream_t stream) {
  // function attributes are per device: a process that switches devices (g16_set_device) needs them on each
  static std::atomic<uint64_t> attr_devices{0};
  int dev = 0;
  G16_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  if (!(attr_devices.load(std::memory_order_relaxed) & bit)) {
    G16_CUDA(cudaFuncSetAttribute(k_ntt_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    G16_CUDA(cudaFuncSetAttribute(k_ntt_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    G16_CUDA(cudaFuncSetAttribute(k_ntt_pass8<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    G16_CUDA(cudaFuncSetAttribute(k_ntt_pass8<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_devices.fetch_or(bit, std::memory_order_relaxed);
  }
  size_t E = (size_t)1 << (a.k + a.logC);
  size_t tiles = ((size_t)1 << a.log_n) / E;
  dim3 grid((unsigned)tiles, (unsigned)batch);
  size_t smem = E * sizeof(Fr);
  static int use8 = -1;
  if (use8 < 0) {
    const char* e = getenv("G16_NTT_RADIX2");           // A/B knob: force the radix-2 shared-memory kernel
    use8 = (e && e[0] == '1') ? 0 : 1;
  }
  if (use8 && a.k + a.logC == NTT_TILE_LOG) {
    if (dif) k_ntt_pass8<true><<<grid, NTT8_THREADS, smem, stream>>>(a);
    else k_ntt_pass8<false><<<grid, NTT8_THREADS, smem, stream>>>(a);
  } else {
    if (dif) k_ntt_pass<true><<<grid, NTT_THREADS, smem, stream>>>(a);
    else k_ntt_pass<false><<<grid, NTT_THREADS, smem, stream>>>(a);
  }
  G16_LAUNCH_CHECK();
}

// Decimation in frequency over all bits: natural-order input, bit-reversed output order.
// The last pass can scale (SC_*) and/or store to bit-reversed addresses (=> natural order in dst).
static void run_dif(const Fr* src, Fr* work, Fr* dst, size_t stride, size_t dst_stride, int batch, int log_n,
                    const Fr* tw, int scale_mode, const Fr* scale, bool bitrev_store, cudaStream_t stream) {
  NttPlan pl = ntt_make_plan(log_n);
  for (int i = pl.npass - 1; i >= 0; i--) {
    bool first = (i == pl.npass - 1), last = (i == 0);
    NttPassArgs a;
    a.src = first ? src : work;
    a.dst = last ? dst : work;
    a.src_stride = stride;
    a.dst_stride = last ? dst_stride : stride;
    a.tw = tw;
    a.scale = scale;
    a.log_n = log_n;
    a.t_lo = pl.pass[i].t_lo;
    a.k = pl.pass[i].k;
    a.logC = pl.pass[i].logC;
    a.scale_mode = last ? scale_mode : SC_NONE;
    a.bitrev_store = (last && bitrev_store) ? 1 : 0;
    a.bitrev_load = 0;
    launch_pass(true, a, batch, stream);
  }
}

// Decimation in time over all bits: bit-reversed input order, natural-order output; in place when src == data.
// With gather = true the first pass reads src[bitrev(g)] instead -- natural-order input, natural-order output, the
// permutation being 32-byte gather READS (the scattered sector writes of a bit-reversed store are what DRAM
// handles worst) -- and the last pass can scale (SC_CONST, or SC_TABLE indexed by the natural output index).
static void run_dit(const Fr* src, Fr* data, size_t src_stride, size_t stride, int batch, int log_n, const Fr* tw,
                    bool gather, int scale_mode, const Fr* scale, cudaStream_t stream) {
  NttPlan pl = ntt_make_plan(log_n);
  for (int i = 0; i < pl.npass; i++) {
    const bool first = i == 0, last = i == pl.npass - 1;
    NttPassArgs a;
    a.src = first ? src : data;
    a.dst = data;
    a.src_stride = first ? src_stride : stride;
    a.dst_stride = stride;
    a.tw = tw;
    a.scale = scale;
    a.log_n = log_n;
    a.t_lo = pl.pass[i].t_lo;
    a.k = pl.pass[i].k;
    a.logC = pl.pass[i].logC;
    a.scale_mode = last ? scale_mode : SC_NONE;
    a.bitrev_store = 0;
    a.bitrev_load = (first && gather) ? 1 : 0;
    launch_pass(false, a, batch, stream);
  }
}

// ---------------------------------------------------------------------------------------
// element-wise kernels
// ---------------------------------------------------------------------------------------
__global__ void k_pointwise_mul(const Fr* a, const Fr* b, Fr* c, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    st_fr(c + i, fmul(ld_fr(a + i), ld_fr(b + i)));
}

// out = a*b - c   (prover.nim:176), optionally * invZ (prover.nim:141)
__global__ void k_quotient_pointwise(const Fr* a, const Fr* b, const Fr* c, Fr* out, uint32_t n, const Fr* invz) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Fr x = fsub(fmul(ld_fr(a + i), ld_fr(b + i)), ld_fr(c + i));
    if (invz) x = fmul(x, ldg_fr(invz));
    st_fr(out + i, x);
  }
}

static unsigned ew_grid(size_t n) {
  size_t g = (n + 255) / 256;
  size_t cap = 148 * 16;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

// ---------------------------------------------------------------------------------------
// public (library-internal) entry points
// ---------------------------------------------------------------------------------------
void ntt_prepare(int log_n, cudaStream_t stream) { (void)ntt_tables(log_n, stream); }

void ntt_natural(const Fr* in, Fr* out, Fr* work, int log_n, bool inverse, cudaStream_t stream) {
  G16_REQUIRE(in != out, "ntt_natural: output must not alias the input");
  (void)work;
  NttTables& t = ntt_tables(log_n, stream);
  size_t n = (size_t)1 << log_n;
  if (inverse)
    run_dit(in, out, n, n, 1, log_n, t.tw_inv.as<Fr>(), true, SC_CONST, t.consts.as<Fr>() + 2, stream);
  else
    run_dit(in, out, n, n, 1, log_n, t.tw_fwd.as<Fr>(), true, SC_NONE, nullptr, stream);
}

// abc: 3n elements [Az | Bz | scratch]; on return abc is clobbered and qs holds the n scalars that
// multiply the H points.  flavour: 0 = JensGroth (prover.nim:118-148), 1 = Snarkjs (prover.nim:158-181).
void quotient(Fr* abc, Fr* qs, int log_n, int flavour, cudaStream_t stream) {
  G16_REQUIRE(log_n >= 1, "quotient: the reference needs a domain of at least 2 (prover.nim:101)");
  NttTables& t = ntt_tables(log_n, stream);
  size_t n = (size_t)1 << log_n;
  Fr* A = abc;
  Fr* B = abc + n;
  Fr* C = abc + 2 * n;
  k_pointwise_mul<<<ew_grid(n), 256, 0, stream>>>(A, B, C, (uint32_t)n);        // prover.nim:69-71
  G16_LAUNCH_CHECK();
  // shiftEvalDomain (prover.nim:109-113) on the three vectors at once
  run_dif(abc, abc, abc, n, n, 3, log_n, t.tw_inv.as<Fr>(), SC_TABLE_BITREV, t.coset.as<Fr>(), false, stream);
  run_dit(abc, abc, n, n, 3, log_n, t.tw_fwd.as<Fr>(), false, SC_NONE, nullptr, stream);
  if (flavour == 1) {
    k_quotient_pointwise<<<ew_grid(n), 256, 0, stream>>>(A, B, C, qs, (uint32_t)n, nullptr);
    G16_LAUNCH_CHECK();
  } else {
    ntt_ensure_coset_inv(t, stream);
    k_quotient_pointwise<<<ew_grid(n), 256, 0, stream>>>(A, B, C, A, (uint32_t)n, t.consts.as<Fr>() + 5);
    G16_LAUNCH_CHECK();
    // inverse NTT (prover.nim:142) then * eta^-i (prover.nim:143), natural order in and out
    run_dit(A, qs, n, n, 1, log_n, t.tw_inv.as<Fr>(), true, SC_TABLE, t.coset_inv.as<Fr>(), stream);
  }
}

void pointwise_mul(const Fr* a, const Fr* b, Fr* c, size_t n, cudaStream_t stream) {
  k_pointwise_mul<<<ew_grid(n), 256, 0, stream>>>(a, b, c, (uint32_t)n);
  G16_LAUNCH_CHECK();
}

}  // namespace g16
