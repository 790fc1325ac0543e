// Diagnostics: device self-test of the field / curve layer and the integer-pipe microbenchmark that
// provides the IMAD roofline denominator ("this peak is not in MEASURED_PEAKS.json").
#include <vector>
#define G16_FP2_WHOLE_CALL   // the self-test exercises the lazily reduced Fp2 multiplication of the G2 MSM kernels
#include "common.cuh"
#include "ec.cuh"
#ifdef G16_EXPERIMENTS
#include "experiments/field_fp64.cuh"   // FP64-pipe multiplier, measured and rejected
#endif

namespace g16 {

// ---------------------------------------------------------------------------------------
// self-test: the same __host__ __device__ formulas run on the GPU (PTX carry chains) and on the host
// (portable path of field.cuh); any mismatch means the device arithmetic is wrong.
// ---------------------------------------------------------------------------------------
struct alignas(16) SelfCase {
  Fr ra, rb;
  Fp pa, pb;
};
struct alignas(16) SelfOut {
  Fr rmul, radd, rsub, rinv;
  Fp pmul, psub, pdmul;
  Fp2 qmul, qsqr;
  G1Affine g1;
  G2Affine g2;
};

static __host__ __device__ void self_eval(const SelfCase& c, SelfOut& o) {
  o.rmul = fmul(c.ra, c.rb);
  o.radd = fadd(c.ra, c.rb);
  o.rsub = fsub(c.ra, c.rb);
  o.rinv = finv(c.ra);
  o.pmul = fmul(c.pa, c.pb);
  o.psub = fsub(c.pa, c.pb);
#if defined(__CUDA_ARCH__) && defined(G16_EXPERIMENTS)
  o.pdmul = fd_to_fp(dfmul(fd_from_fp(c.pa), fd_from_fp(c.pb)));   // FP64-pipe multiplier: a*b*2^-260
#else
  {
    Fp k = Fp::zero();
    k.v[7] = 1u << 28;                                              // 2^252: fmul(x, k) = x * 2^-4
    o.pdmul = fmul(fmul(c.pa, c.pb), k);
  }
#endif
  Fp2 x, y;
  x.c0 = c.pa;
  x.c1 = c.pb;
  y.c0 = c.pb;
  y.c1 = fadd(c.pa, c.pa);
  o.qmul = fmul(x, y);
  o.qsqr = fsqr(x);
  // curve: (ra * G + rb * G + G) in G1 via scalar-mul, madd, add, dbl
  G1Affine g;
  g.x = Fp::one();
  g.y = fdbl(Fp::one());
  G1XYZZ p = xyzz_scalar_mul(c.ra.v, g);
  G1XYZZ q = xyzz_scalar_mul(c.rb.v, g);
  G1XYZZ s = xyzz_add(p, q);
  s = xyzz_madd(s, g);
  s = xyzz_add(s, xyzz_dbl(s));
  o.g1 = xyzz_to_affine(s);
  // G2 on the point (x, y) = (qsqr, qmul): not on the twist, but the formulas are polynomial identities
  G2Affine h;
  h.x = o.qsqr;
  h.y = o.qmul;
  G2XYZZ u = xyzz_dbl_affine(h);
  u = xyzz_madd(u, h);
  G2XYZZ v = xyzz_add(u, xyzz_dbl(u));
  o.g2 = xyzz_to_affine(v);
}

__global__ void k_selftest(const SelfCase* in, SelfOut* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SelfCase c = in[i];
  SelfOut o;
  self_eval(c, o);
  out[i] = o;
}

static uint64_t sm64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <class P>
static Fe<P> rand_fe(uint64_t& s) {
  Fe<P> x;
  for (int i = 0; i < 4; i++) {
    uint64_t v = sm64(s);
    x.v[2 * i] = (uint32_t)v;
    x.v[2 * i + 1] = (uint32_t)(v >> 32);
  }
  x.v[7] &= 0x0fffffffu;   // < 2^252 < modulus
  return x;
}

int selftest_run(uint32_t seed, uint32_t cases) {
  if (cases == 0) cases = 64;
  std::vector<SelfCase> in(cases);
  uint64_t s = seed;
  for (uint32_t i = 0; i < cases; i++) {
    in[i].ra = rand_fe<FrParams>(s);
    in[i].rb = rand_fe<FrParams>(s);
    in[i].pa = rand_fe<FpParams>(s);
    in[i].pb = rand_fe<FpParams>(s);
  }
  // edge cases: modulus-1, 0, 1
  if (cases >= 4) {
    in[0].ra = Fr::modulus();
    in[0].ra.v[0] -= 1;
    in[0].rb = in[0].ra;
    in[0].pa = Fp::modulus();
    in[0].pa.v[0] -= 1;
    in[0].pb = in[0].pa;
    in[1].rb = Fr::zero();
    in[1].pb = Fp::zero();
    in[2].ra = Fr::one();
    in[2].pa = Fp::one();
  }
  DevBuf din, dout;
  din.ensure(cases * sizeof(SelfCase));
  dout.ensure(cases * sizeof(SelfOut));
  G16_CUDA(cudaMemcpyAsync(din.p, in.data(), cases * sizeof(SelfCase), cudaMemcpyHostToDevice, 0));
  k_selftest<<<div_up(cases, 32), 32>>>(din.as<SelfCase>(), dout.as<SelfOut>(), cases);   // same (default) stream
  G16_LAUNCH_CHECK();
  G16_CUDA(cudaDeviceSynchronize());
  std::vector<SelfOut> got(cases);
  G16_CUDA(cudaMemcpy(got.data(), dout.p, cases * sizeof(SelfOut), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (uint32_t i = 0; i < cases; i++) {
    SelfOut want;
    self_eval(in[i], want);
    if (memcmp(&want, &got[i], sizeof(SelfOut)) != 0) {
      if (!bad) {
        const char* names[] = {"rmul", "radd", "rsub", "rinv", "pmul", "psub", "pdmul", "qmul", "qsqr", "g1", "g2"};
        size_t offs[] = {offsetof(SelfOut, rmul), offsetof(SelfOut, radd), offsetof(SelfOut, rsub),
                         offsetof(SelfOut, rinv), offsetof(SelfOut, pmul), offsetof(SelfOut, psub), offsetof(SelfOut, pdmul),
                         offsetof(SelfOut, qmul), offsetof(SelfOut, qsqr), offsetof(SelfOut, g1),
                         offsetof(SelfOut, g2), sizeof(SelfOut)};
        std::string msg = "selftest mismatch in case " + std::to_string(i) + ":";
        for (int k = 0; k < 11; k++)
          if (memcmp((char*)&want + offs[k], (char*)&got[i] + offs[k], offs[k + 1] - offs[k]) != 0)
            msg += std::string(" ") + names[k];
        set_last_error(msg);
      }
      bad++;
    }
  }
  return bad;
}

// ---------------------------------------------------------------------------------------
// integer pipe microbenchmark
// ---------------------------------------------------------------------------------------
constexpr int IP_ITERS = 2048;

__global__ void __launch_bounds__(256) k_ip_madlo(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
#pragma unroll 1
  for (int it = 0; it < IP_ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_ip_madhi(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0x9e3779b9u + i;
#pragma unroll 1
  for (int it = 0; it < IP_ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// lo/hi carry pairs exactly as in fmul rows: 4 wide MACs per asm block, 4 independent accumulators
__global__ void __launch_bounds__(256) k_ip_wide(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[4][8];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 8; i++) x[j][i] = threadIdx.x + i + j;
#if defined(__CUDA_ARCH__)
  uint32_t p1 = a, p3 = a + 2, p5 = a + 4, p7 = a + 6;
#pragma unroll 1
  for (int it = 0; it < IP_ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int j = 0; j < 4; j++) mad_row_nc(x[j], p1, p3, p5, p7, b + j);
  }
#else
  (void)a;
  (void)b;
#endif
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_ip_fmul_lazy(Fp* out, uint32_t a) {
  Fp x = Fp::one(), y = Fp::rsquared(), z = Fp::one();
  x.v[0] += threadIdx.x;
  z.v[0] += a;
#pragma unroll 1
  for (int it = 0; it < IP_ITERS / 8; it++) {
    x = fmul_core<true>(x, y);
    z = fmul_core<true>(z, y);
    x = fmul_core<true>(x, z);
    z = fmul_core<true>(z, x);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = fadd(x, z);
}
__global__ void __launch_bounds__(256) k_ip_fmul(Fp* out, uint32_t a) {
  Fp x = Fp::one(), y = Fp::rsquared(), z = Fp::one();
  x.v[0] += threadIdx.x;
  z.v[0] += a;
#pragma unroll 1
  for (int it = 0; it < IP_ITERS / 8; it++) {
    x = fmul(x, y);
    z = fmul(z, y);
    x = fmul(x, z);
    z = fmul(z, x);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = fadd(x, z);
}

#ifdef G16_EXPERIMENTS
// FP64-pipe experiment: the DFMA Montgomery multiplier of field_fp64.cuh, same dependency pattern as k_ip_fmul
// mode bit 0: even warps run the IMAD Montgomery multiply loop; bit 1: odd warps run the DFMA primitive loop
__global__ void __launch_bounds__(256) k_ip_mix(Fp* out, uint32_t a, int mode, int fmul_iters, int dfma_iters,
                                                uint32_t dfma_warps) {
  const bool odd = (dfma_warps >> (threadIdx.x >> 5)) & 1;
  Fp r = Fp::zero();
  if (!odd) {
    if (!(mode & 1)) return;
    Fp x = Fp::one(), y = Fp::rsquared(), z = Fp::one();
    x.v[0] += threadIdx.x;
    z.v[0] += a;
#pragma unroll 1
    for (int it = 0; it < fmul_iters; it++) {
      x = fmul(x, y);
      z = fmul(z, y);
      x = fmul(x, z);
      z = fmul(z, x);
    }
    r = fadd(x, z);
  } else {
    if (!(mode & 2)) return;
    Fp x0 = Fp::one(), z0 = Fp::one();
    x0.v[0] += threadIdx.x;
    z0.v[0] += a;
    Fd x = fd_from_fp(x0), y = fd_from_fp(Fp::rsquared()), z = fd_from_fp(z0);
#pragma unroll 1
    for (int it = 0; it < dfma_iters; it++) {
      x = dfmul(x, y);
      z = dfmul(z, y);
      x = dfmul(x, z);
      z = dfmul(z, x);
    }
    r = fadd(fd_to_fp(x), fd_to_fp(z));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

#endif

// Instruction-mix probe for a Karatsuba multiplier: per "multiply" 28 rows of 4 wide MACs (112 IMAD.WIDE instead
// of 128) plus `adds8` carry chains of 8 IADD3 on the ALU pipe; wide = 32 rows and adds8 = 5 models today's fmul.
template <int ROWS, int ADDS8>
__global__ void __launch_bounds__(256) k_ip_kmix(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t x[4][8], y[3][8];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 8; i++) x[j][i] = threadIdx.x + i + j;
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 8; i++) y[j][i] = threadIdx.x * 3 + i + j;
#if defined(__CUDA_ARCH__)
  uint32_t p1 = a, p3 = a + 2, p5 = a + 4, p7 = a + 6;
#pragma unroll 1
  for (int it = 0; it < IP_ITERS / 4; it++) {
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      mad_row_nc(x[r & 3], p1, p3, p5, p7, b + r);
      if ((r * ADDS8) / ROWS != ((r + 1) * ADDS8) / ROWS) {
        const int k = (r * ADDS8) / ROWS;
        add8_ip(y[k % 3], x[(r + 2) & 3]);   // consumes a product row like the Karatsuba glue would
      }
    }
  }
#else
  (void)a;
  (void)b;
#endif
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[j][i];
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= y[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

void bench_int_pipe(int kind, double* ops_per_sec, float* ms) {
  int dev = 0, sms = 0;
  G16_CUDA(cudaGetDevice(&dev));
  G16_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256;
  DevBuf out;
  out.ensure((size_t)blocks * threads * sizeof(Fp));
  cudaEvent_t e0, e1;
  G16_CUDA(cudaEventCreate(&e0));
  G16_CUDA(cudaEventCreate(&e1));
  auto launch = [&] {
    if (kind == 0) k_ip_madlo<<<blocks, threads>>>(out.as<uint32_t>(), 3, 5);
    else if (kind == 1) k_ip_madhi<<<blocks, threads>>>(out.as<uint32_t>(), 0x9e3779b9u, 5);
    else if (kind == 2) k_ip_wide<<<blocks, threads>>>(out.as<uint32_t>(), 0x9e3779b9u, 0x7f4a7c15u);
    else if (kind == 3) k_ip_fmul<<<blocks, threads>>>(out.as<Fp>(), 7);
    else if (kind == 13) k_ip_fmul_lazy<<<blocks, threads>>>(out.as<Fp>(), 7);
    else if (kind == 10) k_ip_kmix<32, 5><<<blocks, threads>>>(out.as<uint32_t>(), 0x9e3779b9u, 0x7f4a7c15u);
    else if (kind == 11) k_ip_kmix<28, 19><<<blocks, threads>>>(out.as<uint32_t>(), 0x9e3779b9u, 0x7f4a7c15u);
    else if (kind == 12) k_ip_kmix<28, 25><<<blocks, threads>>>(out.as<uint32_t>(), 0x9e3779b9u, 0x7f4a7c15u);
    else {
#ifdef G16_EXPERIMENTS
      k_ip_mix<<<blocks, threads>>>(out.as<Fp>(), 7, kind == 4 ? 2 : kind == 5 ? 1 : 3, IP_ITERS / 8, IP_ITERS / 8,
                                    kind <= 6 ? 0xAAu : kind == 7 ? 0xFFu : kind == 8 ? 0x88u : 0xEEu);
#else
      G16_REQUIRE(false, "kinds 4..9 (FP64-pipe experiment) need a library built with `make EXPERIMENTS=1`");
#endif
    }
    G16_LAUNCH_CHECK();
  };
  for (int w = 0; w < 2; w++) launch();
  G16_CUDA(cudaDeviceSynchronize());
  const int reps = 5;
  G16_CUDA(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) launch();
  G16_CUDA(cudaEventRecord(e1));
  G16_CUDA(cudaEventSynchronize(e1));
  float t = 0.f;
  G16_CUDA(cudaEventElapsedTime(&t, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double per_thread;
  if (kind == 0 || kind == 1) per_thread = (double)IP_ITERS * 32.0;        // MAC32 per thread
  else if (kind == 2) per_thread = (double)IP_ITERS * 2.0 * 4.0 * 4.0;      // wide MAC32 (lo/hi pair = 1)
  else if (kind == 3) per_thread = (double)(IP_ITERS / 8) * 4.0;             // modmuls
  else if (kind == 13) per_thread = (double)(IP_ITERS / 8) * 4.0;
  else if (kind >= 10) per_thread = (double)(IP_ITERS / 4);                   // modelled multiplies
  else if (kind >= 7) per_thread = (double)(IP_ITERS / 8) * 4.0;             // all warps, either multiplier
  else if (kind == 4) per_thread = (double)(IP_ITERS / 8) * 4.0 * 0.5;       // FP64 modmuls, odd warps only
  else per_thread = (double)(IP_ITERS / 8) * 4.0 * 0.5;                      // modmuls, even warps only
  double total = per_thread * (double)blocks * threads * reps;
  *ms = t / reps;
  *ops_per_sec = total / ((double)t * 1e-3);
}

}  // namespace g16
